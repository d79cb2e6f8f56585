"""Host-side orchestration of the sm_100a kernels: the ODE loop and the HiFi-GAN generator as sequences of C-ABI
calls on the current stream, captured once per (batch, frames, dt, truncation) bucket into a CUDA graph.

Mirrors, step for step, the reference call stack:
ConditionalFlowMatchingModel.sample (src/flow_matching/models.py:132-189) and
FastSpeech2ConformerHifiGan.forward (HF:1451-1491).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _native as nat
from .packing import (RESBLOCK_DILATIONS, RESBLOCK_KERNELS, UPSAMPLE_KERNELS, UPSAMPLE_RATES, PackedCFM, PackedVocoder,
                      pack_cfm, pack_vocoder)

P = nat.ptr


def pad_value_f32() -> float:
    """log(float32(1e-5)) -- hifigan/data.py:9-10 evaluated the way the reference does (float32 tensor math)."""
    return float(torch.log(torch.clamp(torch.tensor(0.0, dtype=torch.float32), min=1e-5)))


def ode_times(dt: float) -> torch.Tensor:
    """The reference iterates the float32 tensor torch.arange(0, 1, dt) (models.py:172); only its values matter."""
    return torch.arange(0, 1, dt, dtype=torch.float32)


def _i32(vals) -> "nat.ctypes.Array":
    import ctypes

    return (ctypes.c_int32 * len(vals))(*vals)


class _Fork:
    """Run independent kernel chains on side streams (captured as parallel branches of the CUDA graph): the
    persistent GEMM kernels leave SMs idle in their last wave, and an independent chain fills those tails."""

    def __init__(self, device, n: int):
        with torch.cuda.device(device):
            self.side = [torch.cuda.Stream(device=device) for _ in range(n)]
        self.enabled = True

    def run(self, branches) -> None:
        """branches: list of callables; branch 0 runs on the current stream, the others on side streams."""
        if not self.enabled or len(branches) == 1:
            for fn in branches:
                fn()
            return
        main = torch.cuda.current_stream()
        start = torch.cuda.Event()
        start.record(main)
        done = []
        for fn, st in zip(branches[1:], self.side):
            st.wait_event(start)
            with torch.cuda.stream(st):
                fn()
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        branches[0]()
        for ev in done:
            main.wait_event(ev)


class CFMSampler:
    """ODE sampler over the flow-matching transformer velocity field (kernels: srb_cfm_*)."""

    def __init__(self, packed: PackedCFM, depth: int, mean: float, std: float):
        self.w = packed
        self.depth = depth
        self.mean = float(mean)
        self.std = float(std)
        self.device = packed.w_embed.device
        self._rot: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._cond_cache: Dict[Tuple[float, ...], torch.Tensor] = {}
        self.fork = _Fork(self.device, 1)

    # -- tables -----------------------------------------------------------------------------------------------
    def rotary(self, rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._rot is None or self._rot[0].shape[0] < rows:
            n = max(rows, 1024)
            cs = torch.empty(n, 64, dtype=torch.float32, device=self.device)
            sn = torch.empty_like(cs)
            nat.call("srb_rotary_table", P(self.w.inv_freq), n, P(cs), P(sn))
            self._rot = (cs, sn)
        return self._rot

    def cond_table(self, times: torch.Tensor) -> torch.Tensor:
        """g[nfe][2*depth][256] = sqrt(H) (W_gamma c(t) + 1): batch independent, cached per time grid."""
        key = tuple(float(t) for t in times)
        g = self._cond_cache.get(key)
        if g is None:
            nfe = len(key)
            t_dev = times.to(self.device)
            temb = torch.empty(nfe, 256, dtype=torch.float32, device=self.device)
            g = torch.empty(nfe, 2 * self.depth, 256, dtype=torch.float32, device=self.device)
            nat.call("srb_time_cond_table", P(t_dev), nfe, P(self.w.four_w), P(self.w.lin_w), P(self.w.lin_b),
                     P(self.w.gamma_w), 2 * self.depth, P(temb), P(g))
            torch.cuda.current_stream().synchronize()  # t_dev must outlive the launch
            self._cond_cache[key] = g
            self._last_time_emb = temb
        return g

    # -- duration-prediction variant ---------------------------------------------------------------------------
    def regulate(self, input_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """models.py:157-164: predict frames per unit and expand the ids (length_regulator on ids: the pad id's embedding
        is the zero row).  Returns (expanded ids (B, max_len) int64, durations (B, N) int32).  One host sync: the
        expanded length decides the shapes of everything downstream, exactly like `lengths.max()` in the reference."""
        if self.w.dur_table is None:
            raise RuntimeError("this checkpoint has no duration predictor (config.predict_duration is False)")
        ids = input_ids.contiguous()
        b, n = ids.shape
        dur = torch.empty(b, n, dtype=torch.int32, device=self.device)
        tot = torch.empty(b, dtype=torch.int32, device=self.device)
        nat.call("srb_duration_predict", P(ids), P(self.w.dur_table), self.w.dur_bias, P(dur), P(tot), b, n,
                 self.w.dur_table.shape[1])
        totals = tot.cpu()
        all_one = int(totals.sum()) == 0          # the regulator's all-zero rule (HF:113-114)
        n_out = n if all_one else int(totals.max())
        out = torch.empty(b, n_out, dtype=torch.int64, device=self.device)
        nat.call("srb_length_regulate", P(ids), P(dur), P(out), b, n, n_out, 1 if all_one else 0)
        return out, dur

    # -- workspace --------------------------------------------------------------------------------------------
    def workspace(self, batch: int, frames: int) -> Dict[str, torch.Tensor]:
        """`frames` must be a multiple of 8 (see `padded_frames`): the transposed-V operand of the attention kernel
        is addressed by TMA per utterance and TMA needs 16-byte aligned box origins."""
        assert frames % 8 == 0, "CFMSampler.workspace: frames must be padded to a multiple of 8"
        dev, m = self.device, batch * frames
        m_pad = (m + 255) // 256 * 256
        f32 = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        b16 = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)
        return dict(
            ids=torch.zeros(batch, frames, dtype=torch.int64, device=dev),
            lengths=torch.zeros(batch, dtype=torch.int32, device=dev),
            cond=f32(m, 256), xt=torch.zeros(batch, frames, 80, dtype=torch.float32, device=dev),
            xt_b=b16(batch, frames, 80), x0=f32(m, 256), x=f32(m, 256),
            # xn is also the B operand of the V^T GEMM, read in 256-row tiles: rows >= m stay zero forever
            xn=torch.zeros(m_pad, 256, dtype=torch.bfloat16, device=dev),
            qk=b16(m, 512), vt=b16(256, m_pad), o=b16(m, 256), h=b16(m, 896),
            mel=f32(batch, frames, 80), mel_b=b16(batch, frames, 80),
            # max |q|^2, |k|^2 per (utterance, head), recorded by qk_rope, read by the attention kernel; two buffers used
            # alternately by successive layers (each projection clears the other one for its successor)
            qkmax=torch.zeros(2, batch, 2, 2, 2, dtype=torch.float32, device=dev),
        )

    # -- the loop ---------------------------------------------------------------------------------------------
    def prepare(self, ws: Dict[str, torch.Tensor], truncation: Optional[float]) -> None:
        """mask/lengths (models.py:152), hoisted conditioning gather (:154,:175-176), prior clamp (:169-170)."""
        b, n = ws["ids"].shape
        nat.call("srb_unit_lengths", P(ws["ids"]), P(ws["lengths"]), b, n)
        nat.call("srb_embed_gather", P(self.w.cond_table), P(ws["ids"]), P(ws["cond"]), b * n,
                 self.w.cond_table.shape[0], 256, nbytes=b * n * (2 * 256 * 4 + 8))
        tv = float(truncation) if truncation is not None else 0.0
        nat.call("srb_prior_prepare", P(ws["xt"]), P(ws["xt_b"]), b * n * 80, tv)
        ws["qkmax"].zero_()
        self._qk_calls = 0

    def step(self, ws: Dict[str, torch.Tensor], g_step: torch.Tensor, dt: float, last: bool) -> None:
        """One velocity evaluation + Euler update (models.py:173-184); `last` adds :186-187."""
        b, n = ws["ids"].shape
        w, L = self.w, ws["lengths"]
        cs, sn = self.rotary(n)
        m = b * n
        nat.call("srb_cfm_embed", P(ws["xt_b"]), P(w.w_embed), P(ws["cond"]), P(ws["x0"]), b, n, flops=2.0 * m * 80 * 256)
        nat.call("srb_cfm_posconv_norm", P(ws["x0"]), P(w.dw_w), P(w.dw_b), P(g_step[0]), P(L), P(ws["x"]), P(ws["xn"]), b, n,
                 flops=2.0 * m * 31 * 256, nbytes=m * 256 * (4 + 4 + 2))
        for i in range(self.depth):
            m_pad = ws["vt"].shape[1]
            qk_cur, qk_next = ws["qkmax"][self._qk_calls % 2], ws["qkmax"][(self._qk_calls + 1) % 2]
            self._qk_calls += 1
            # q|k projection and the transposed-v projection both read xn only: two parallel graph branches
            self.fork.run([
                lambda: nat.call("srb_cfm_qk_rope", P(ws["xn"]), P(w.w_qkv[i]), P(cs), P(sn), P(ws["qk"]),
                                 P(qk_cur), P(qk_next), b, n, flops=2.0 * m * 256 * 512),
                lambda: nat.call("srb_cfm_v_transposed", P(ws["xn"]), P(w.w_qkv[i][512:]), P(ws["vt"]), m_pad,
                                 flops=2.0 * m * 256 * 256),
            ])
            nat.call("srb_cfm_attention_tc", P(ws["qk"]), 512, P(ws["vt"]), m_pad, P(L), P(qk_cur), P(ws["o"]), b, n,
                     flops=4.0 * m * n * 256)
            nat.call("srb_cfm_attn_out_norm", P(ws["o"]), P(w.w_out[i]), P(g_step[2 * i + 1]), P(L), P(ws["x"]), P(ws["xn"]), b, n,
                     flops=2.0 * m * 256 * 256)
            nat.call("srb_cfm_ffn_glu", P(ws["xn"]), P(w.w_ff1[i]), P(w.b_ff1[i]), P(L), P(ws["h"]), b, n,
                     flops=2.0 * m * 768 * 1792)
            if i + 1 < self.depth:
                g_next, mode = g_step[2 * i + 2], 1
            else:
                g_next, mode = w.final_norm_w, 2
            nat.call("srb_cfm_ffn_out_norm", P(ws["h"]), P(w.w_ff2[i]), P(w.b_ff2[i]), P(g_next), mode, P(L), P(ws["x"]),
                     P(ws["xn"]), b, n, flops=2.0 * m * 2688 * 256)
        mel = P(ws["mel"]) if last else None
        mel_b = P(ws["mel_b"]) if last else None
        nat.call("srb_cfm_pred_euler", P(ws["xn"]), P(w.w_pred), float(dt), P(ws["xt"]), P(ws["xt_b"]), mel, mel_b,
                 self.std, self.mean, pad_value_f32(), P(L), b, n, flops=2.0 * m * 256 * 80)

    def run(self, ws: Dict[str, torch.Tensor], dt: float, truncation: Optional[float]) -> None:
        """ids and the prior sample must already be in ws['ids'] / ws['xt']; result lands in ws['mel'], ws['mel_b']."""
        times = ode_times(dt)
        g = self.cond_table(times)
        self.prepare(ws, truncation)
        nfe = len(times)
        for s in range(nfe):
            self.step(ws, g[s], dt, last=(s == nfe - 1))


def padded_frames(frames: int) -> int:
    """Frames are padded to a multiple of 8 inside the sampler (extra rows are ordinary pad frames: masked
    everywhere, never visible to valid frames, dropped before the vocoder)."""
    return (frames + 7) // 8 * 8


def waveform_rows(frames: int) -> int:
    """_get_waveform_lengths (models.py:211-221) for the padded frame count: 320*T + 80."""
    rows = frames
    for k, s in zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES):
        rows = (rows - 1) * s - 2 * ((k - s) // 2) + k
    return rows


class HifiGanGenerator:
    """mel (B, T, 80) bf16 -> waveform (B, 320 T + 80) fp32 (kernels: srb_hifigan_*)."""

    def __init__(self, packed: PackedVocoder, slope: float = 0.1, fuse_mrf: bool = True):
        self.w = packed
        self.slope = float(slope)
        self.fuse_mrf = fuse_mrf
        # SRB_PAIR_UPSAMPLE=0 runs every up-sampler through the polyphase kernel instead (A/B runs)
        self.pair_upsample = os.environ.get("SRB_PAIR_UPSAMPLE", "1") != "0"
        self.fork = _Fork(packed.w_pre.device, 2)
        self.device = packed.w_pre.device

    def workspace(self, batch: int, frames: int) -> Dict[str, object]:
        dev = self.device
        b16 = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)
        ws: Dict[str, object] = {"pre": b16(batch, frames, 512)}
        rows, c = frames, 512
        stages = []
        for k, s in zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES):
            rows = (rows - 1) * s - 2 * ((k - s) // 2) + k
            c //= 2
            st = dict(rows=rows, c=c, u_raw=b16(batch, rows, c), u_act=b16(batch, rows, c), out=b16(batch, rows, c),
                      t=[b16(batch, rows, c) for _ in range(3)], xr=[b16(batch, rows, c) for _ in range(3)],
                      xa=[b16(batch, rows, c) for _ in range(3)])
            stages.append(st)
        ws["stages"] = stages
        ws["wav"] = torch.empty(batch, rows, dtype=torch.float32, device=dev)
        return ws

    def run(self, mel_b: torch.Tensor, ws: Dict[str, object]) -> torch.Tensor:
        b, t, _ = mel_b.shape
        w = self.w
        one = _i32([7])
        dil1 = _i32([1])
        # conv_pre (HF:1470); its only consumer is leaky_relu -> upsampler, so only the activated copy is stored
        nat.call("srb_hifigan_conv", P(mel_b), None, None, 1, one, dil1, P(w.w_pre), P(w.b_pre), None, None, None, None,
                 P(ws["pre"]), b, t, 80, 512, 1.0, self.slope, flops=2.0 * b * t * 7 * 80 * 512)
        x_act, rows_in, c_in = ws["pre"], t, 512
        n_stage = len(UPSAMPLE_RATES)
        for i, (k, s) in enumerate(zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES)):
            st = ws["stages"][i]
            rows, c = st["rows"], st["c"]
            fused = self.fuse_mrf and i in w.w_mrf
            # upsampler (HF:1472-1473): raw copy = residual of the three resblocks, activated copy = their input
            if i in w.up_pair and self.pair_upsample:
                # L_out = s L: all s output phases of an input row from one 3-tap conv (packing.upsampler_as_row_group_conv);
                # (B, rows_in, s c) is the (B, rows, c) result.  FLOPs reported are the transposed conv's own.
                wp, bp = w.up_pair[i]
                nat.call("srb_hifigan_conv", P(x_act), None, None, 1, _i32([3]), dil1, P(wp), P(bp), None, None, None,
                         P(st["u_raw"]), None if fused else P(st["u_act"]), b, rows_in, c_in, s * c, 1.0, self.slope,
                         flops=2.0 * b * rows_in * k * c_in * c)
            else:
                nat.call("srb_hifigan_upsample", P(x_act), P(w.w_up[i]), P(w.b_up[i]), P(st["u_raw"]),
                         None if fused else P(st["u_act"]), b, rows_in, c_in, c, k, s, self.slope,
                         flops=2.0 * b * rows_in * k * c_in * c)
            if fused:
                # narrow stages: the whole MRF block (18 convs + residuals + mean + next leaky_relu) in one kernel
                slope_next = self.slope if i + 1 < n_stage else 0.01
                nat.call("srb_hifigan_mrf_fused", P(st["u_raw"]), P(w.w_mrf[i]), P(w.b_mrf[i]), P(st["out"]), b, rows, c,
                         self.slope, slope_next, flops=252.0 * c * c * b * rows)
                x_act, rows_in, c_in = st["out"], rows, c
                continue
            res: List[Optional[torch.Tensor]] = [None, None, None]

            def chain(j: int, rk: int, st=st, rows=rows, c=c, i=i) -> None:
                kk = _i32([rk])
                xr, xa = st["u_raw"], st["u_act"]
                for q, dil in enumerate(RESBLOCK_DILATIONS):
                    # conv1 with dilation (HF:1361-1363), output only needed activated
                    nat.call("srb_hifigan_conv", P(xa), None, None, 1, kk, _i32([dil]), P(w.w_c1[i][j][q]),
                             P(w.b_c1[i][j][q]), None, None, None, None, P(st["t"][j]), b, rows, c, c, 1.0, self.slope,
                             flops=2.0 * b * rows * rk * c * c)
                    if q < 2:
                        # conv2 + residual (HF:1364-1366): raw (next residual) and activated (next conv1 input)
                        nat.call("srb_hifigan_conv", P(st["t"][j]), None, None, 1, kk, dil1, P(w.w_c2[i][j][q]),
                                 P(w.b_c2[i][j][q]), P(xr), None, None, P(st["xr"][j]), P(st["xa"][j]), b, rows, c, c,
                                 1.0, self.slope, flops=2.0 * b * rows * rk * c * c)
                        xr, xa = st["xr"][j], st["xa"][j]
                res[j] = xr

            # the three resblocks of a stage are independent until the MRF mean: three parallel graph branches
            # (longest chain, k = 11, on the main stream)
            self.fork.run([lambda: chain(2, RESBLOCK_KERNELS[2]), lambda: chain(1, RESBLOCK_KERNELS[1]),
                           lambda: chain(0, RESBLOCK_KERNELS[0])])
            # fused MRF tail: the three last conv2's + their residuals + mean (HF:1475-1478) + the next leaky_relu
            # (slope 0.1 before an upsampler, torch default 0.01 before conv_post, HF:1480)
            slope_next = self.slope if i + 1 < n_stage else 0.01
            nat.call("srb_hifigan_conv", P(st["t"][0]), P(st["t"][1]), P(st["t"][2]), 3, _i32(list(RESBLOCK_KERNELS)),
                     _i32([1, 1, 1]), P(w.w_tail[i]), P(w.b_tail[i]), P(res[0]), P(res[1]), P(res[2]), None, P(st["out"]),
                     b, rows, c, c, 1.0 / 3.0, slope_next, flops=2.0 * b * rows * sum(RESBLOCK_KERNELS) * c * c)
            x_act, rows_in, c_in = st["out"], rows, c
        nat.call("srb_hifigan_post", P(x_act), P(w.w_post), w.b_post, P(ws["wav"]), b, rows_in,
                 flops=2.0 * b * rows_in * 7 * 16, nbytes=b * rows_in * (16 * 2 + 4))
        return ws["wav"]


@dataclass
class _Plan:
    cfm_ws: Dict[str, torch.Tensor]
    voc_ws: Dict[str, object]
    graph: Optional[torch.cuda.CUDAGraph]
    body: object
    n_launches: int


def build_sampler(state_dict: Dict[str, torch.Tensor], device, depth: int = 4, mean: float = -5.8843,
                  std: float = 2.2615) -> CFMSampler:
    """state_dict uses the top-level key names ("model.*")."""
    nat.require_blackwell()
    device = torch.device(device)
    with torch.cuda.device(device):
        return CFMSampler(pack_cfm(state_dict, device, depth=depth), depth, mean, std)


def build_vocoder(state_dict: Dict[str, torch.Tensor], device, slope: float = 0.1) -> HifiGanGenerator:
    """state_dict uses the top-level key names ("vocoder.*")."""
    nat.require_blackwell()
    device = torch.device(device)
    with torch.cuda.device(device):
        return HifiGanGenerator(pack_vocoder(state_dict, device), slope)


class ResynthEngine:
    """units -> waveform for one device.  One CUDA graph per (batch, frames, dt, truncation) bucket."""

    def __init__(self, sampler: Optional[CFMSampler], vocoder: Optional[HifiGanGenerator], use_graphs: bool = True):
        nat.require_blackwell()
        self.sampler = sampler
        self.vocoder = vocoder
        self.device = (sampler or vocoder).device
        self.use_graphs = use_graphs
        self._plans: Dict[Tuple, _Plan] = {}
        self._voc_plans: Dict[Tuple, Tuple] = {}

    def _plan(self, batch: int, frames: int, dt: float, truncation: Optional[float], with_vocoder: bool) -> _Plan:
        key = (batch, frames, float(dt), truncation, with_vocoder)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        n8 = padded_frames(frames)
        cfm_ws = self.sampler.workspace(batch, n8)
        voc_ws = self.vocoder.workspace(batch, frames) if with_vocoder else {}
        self.sampler.cond_table(ode_times(dt))
        self.sampler.rotary(n8)
        if with_vocoder and n8 != frames:
            # the vocoder must see exactly the caller's frames (pad frames are vocoded)
            voc_ws["mel_in"] = torch.empty(batch, frames, 80, dtype=torch.bfloat16, device=self.device)

        def body():
            self.sampler.run(cfm_ws, dt, truncation)
            if with_vocoder:
                mel_b = cfm_ws["mel_b"]
                if n8 != frames:
                    voc_ws["mel_in"].copy_(mel_b[:, :frames])
                    mel_b = voc_ws["mel_in"]
                self.vocoder.run(mel_b, voc_ws)

        graph, n_launches = None, 0
        if self.use_graphs:
            # warm-up run outside capture (configures kernel attributes, fills caches), then capture
            cfm_ws["ids"].zero_()
            cfm_ws["ids"][:, :frames].fill_(1)
            # (private generator: the warm-up must not advance the global CUDA RNG, or the first call of a new shape
            # would draw a different prior than the reference's torch.randn with the same seed)
            cfm_ws["xt"].normal_(generator=torch.Generator(device=self.device).manual_seed(0))
            body()
            torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            c0 = nat.launch_count
            with torch.cuda.graph(graph):
                body()
            n_launches = nat.launch_count - c0   # recorded, not executed: counted again at every replay
            nat.launch_count = c0
        plan = _Plan(cfm_ws, voc_ws, graph, body, n_launches)
        self._plans[key] = plan
        return plan

    def _launch(self, plan: _Plan) -> None:
        if plan.graph is not None:
            plan.graph.replay()
            nat.launch_count += plan.n_launches
        else:
            plan.body()

    def _run(self, input_ids, dt, truncation, noise, with_vocoder) -> _Plan:
        assert input_ids.dim() == 2 and input_ids.dtype == torch.int64
        b, n = input_ids.shape
        with torch.cuda.device(self.device):
            plan = self._plan(b, n, dt, truncation, with_vocoder)
            plan.cfm_ws["ids"][:, :n].copy_(input_ids, non_blocking=True)
            if noise is None:
                # same call as the reference (models.py:168) so a seeded run draws the same prior on the same device
                noise = torch.randn(b, n, 80, device=self.device)
            plan.cfm_ws["xt"].zero_()
            plan.cfm_ws["xt"][:, :n].copy_(noise, non_blocking=True)
            self._launch(plan)
        return plan

    def lengths_async(self, input_ids: torch.Tensor):
        """Valid-frame counts of a batch on their way to the host, enqueued BEFORE the big graph so that the caller can wait
        for them (event) while the graph is still running -- the public forward then never drains the GPU between calls.
        srb_unit_lengths writes straight into pinned host memory (device-accessible under unified addressing): a D2H
        copy would queue on the copy engine behind the previous call's 41 MB waveform read-back and hold up the graph
        enqueued after it (measured: +0.8 ms per call).  Returns (pinned int32 tensor, event)."""
        b, n = input_ids.shape
        slot = getattr(self, "_len_slot", 0)
        self._len_slot = slot ^ 1
        bufs = getattr(self, "_len_bufs", None)
        if bufs is None or bufs[0].shape[0] < b:
            bufs = [torch.zeros(max(b, 64), dtype=torch.int32).pin_memory() for _ in range(2)]
            self._len_bufs = bufs
        host_buf = bufs[slot]
        with torch.cuda.device(self.device):
            nat.call("srb_unit_lengths", P(input_ids.contiguous()), host_buf.data_ptr(), b, n)
            ev = torch.cuda.Event()
            ev.record()
        return host_buf[:b], ev

    def sample(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float],
               noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ConditionalFlowMatchingModel.sample: (B, N) int64 -> mel (B, N, 80) fp32 (a fresh tensor)."""
        plan = self._run(input_ids, dt, truncation, noise, with_vocoder=False)
        return plan.cfm_ws["mel"][:, : input_ids.shape[1]].clone()

    def resynthesize(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float],
                     noise: Optional[torch.Tensor] = None):
        """Returns (wav (B, 320 N + 80) fp32 [plan-owned buffer], lengths (B,) int32 device tensor, mel)."""
        plan = self._run(input_ids, dt, truncation, noise, with_vocoder=True)
        return plan.voc_ws["wav"], plan.cfm_ws["lengths"], plan.cfm_ws["mel"][:, : input_ids.shape[1]]

    def vocode(self, mel: torch.Tensor) -> torch.Tensor:
        """decoder.vocoder(mel): (B, T, 80) float -> (B, 320 T + 80) fp32 (a fresh tensor)."""
        squeeze = mel.dim() == 2
        if squeeze:
            mel = mel.unsqueeze(0)
        b, t, _ = mel.shape
        with torch.cuda.device(self.device):
            key = (b, t)
            entry = self._voc_plans.get(key)
            if entry is None:
                ws = self.vocoder.workspace(b, t)
                mel_b = torch.empty(b, t, 80, dtype=torch.bfloat16, device=self.device)
                graph = None
                if self.use_graphs:
                    mel_b.zero_()
                    self.vocoder.run(mel_b, ws)
                    torch.cuda.current_stream().synchronize()
                    graph = torch.cuda.CUDAGraph()
                    c0 = nat.launch_count
                    with torch.cuda.graph(graph):
                        self.vocoder.run(mel_b, ws)
                    n_launch = nat.launch_count - c0
                else:
                    n_launch = 0
                entry = (ws, mel_b, graph, n_launch)
                self._voc_plans[key] = entry
            ws, mel_b, graph, n_launch = entry
            mel_b.copy_(mel)  # fp32 -> bf16 (the reference's autocast would do the same cast at conv_pre)
            if graph is not None:
                graph.replay()
                nat.launch_count += n_launch
            else:
                self.vocoder.run(mel_b, ws)
            wav = ws["wav"].clone()
        return wav[0] if squeeze else wav
