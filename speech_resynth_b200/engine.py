"""Host-side orchestration of the sm_100a kernels: the ODE loop and the HiFi-GAN generator as sequences of C-ABI
calls on the current stream.

Mirrors, step for step, the reference call stack:
ConditionalFlowMatchingModel.sample (src/flow_matching/models.py:132-189) and
FastSpeech2ConformerHifiGan.forward (HF:1451-1491).

Memory and launch model
  * ONE arena per device (`Arena`): every workspace of every (batch, frames) shape is carved from offset 0 of the
    same buffer, so the footprint is that of the largest shape seen, not the sum over shapes.  Shapes run one after
    the other on one stream, so aliasing them is safe; nothing in a workspace survives a call (the per-call staging
    kernel re-establishes the few regions the loop needs zeroed).
  * a shape's FIRST call runs its kernels eagerly (the host enqueues while the GPU still works on the previous
    call); from its SECOND call on it is a CUDA graph.  Graphs live in an LRU of bounded size; a graph is only valid
    for the arena generation it was captured in (growing the arena drops them all).
  * results leave the arena inside the call: the last kernel (conv_post + tanh) is launched after the graph and
    writes either the plan's dense (B, 320 N + 80) buffer or -- ragged form -- a fresh tensor holding the cropped
    utterances back to back, which is what the public `forward` hands out.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple

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


# ------------------------------------------------------------------------------------------------------ arena
class Arena:
    """The device buffer all workspaces are carved from (one per device, shared by every engine on it)."""

    _by_device: Dict[int, "Arena"] = {}

    def __init__(self, device: torch.device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None
        self.generation = 0

    @classmethod
    def get(cls, device: torch.device) -> "Arena":
        idx = device.index if device.index is not None else torch.cuda.current_device()
        a = cls._by_device.get(idx)
        if a is None:
            a = cls._by_device[idx] = Arena(torch.device("cuda", idx))
        return a

    @property
    def nbytes(self) -> int:
        return 0 if self.buf is None else self.buf.numel()

    def ensure(self, nbytes: int) -> None:
        """Grow to at least `nbytes`.  Growing invalidates every pointer handed out before (generation += 1): callers
        compare generations and re-carve / re-capture."""
        if self.nbytes >= nbytes:
            return
        torch.cuda.synchronize(self.device)      # kernels still in flight use the old buffer
        if self.buf is not None:
            self.buf = None                      # release before allocating: never hold both
            torch.cuda.empty_cache()             # ... and hand the old block back to the driver (growth is rare)
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.generation += 1


class Carver:
    """Bump allocation of 1 KB-aligned tensors; with `buf=None` it only measures."""

    def __init__(self, buf: Optional[torch.Tensor], device: torch.device):
        self.buf, self.device, self.off = buf, device, 0

    def take(self, shape: Tuple[int, ...], dtype: torch.dtype) -> Optional[torch.Tensor]:
        n = 1
        for s in shape:
            n *= int(s)
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        off = self.off
        self.off = (off + nbytes + 1023) // 1024 * 1024
        if self.buf is None:
            return None
        return self.buf[off: off + nbytes].view(dtype).view(*shape)


class _Fork:
    """Run independent kernel chains on side streams (captured as parallel branches of the CUDA graph): the
    persistent GEMM kernels leave SMs idle in their last wave, and an independent chain fills those tails."""

    def __init__(self, device, n: int):
        with torch.cuda.device(device):
            self.side = [torch.cuda.Stream(device=device) for _ in range(n)]
        # SRB_FORK: bit 0 = the sampler's q|k / v^T branches, bit 1 = the vocoder's three resblock chains (A/B knob, default 3)
        self.enabled = bool(int(os.environ.get("SRB_FORK", "3")) & (1 if n == 1 else 2))

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


# ------------------------------------------------------------------------------------------------------ sampler
class CFMSampler:
    """ODE sampler over the flow-matching transformer velocity field (kernels: srb_cfm_*)."""

    def __init__(self, packed: PackedCFM, depth: int, mean: float, std: float, tight: bool = False):
        self.w = packed
        self.depth = depth
        # tight=True: the tight-precision library (include/srb.h: srb_split_factor) -- bf16 activations travel as
        # [hi | lo | hi] (three times as wide), weights are packed [Wh | Wh | Wl], attention runs in fp32 on CUDA cores
        self.tight = tight
        self.split = 3 if tight else 1
        self.mean = float(mean)
        self.std = float(std)
        self.device = packed.w_embed.device
        self._rot: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
        self._cond_cache: Dict[Tuple[float, ...], torch.Tensor] = {}
        self.fork = _Fork(self.device, 1)
        # SRB_FUSED_QKV=1: the whole to_qkv GEMM as one launch with a transposing V epilogue (srb_cfm_qk_rope_vt) instead of
        # the q|k projection and the transposed-v GEMM on two parallel branches.  Bit-identical results, 64 launches fewer per
        # call; same-box A/B at config 2: 28.98 / 29.25 ms split vs 29.25 / 29.73 ms fused -- the two branches overlap, the
        # fused launch (27.6 us) is no shorter than they are together, so the split stays the default
        self.fused_qkv = os.environ.get("SRB_FUSED_QKV", "0") == "1"

    def call(self, name: str, *args, **kw) -> None:
        nat.call(name, *args, tight=self.tight, **kw)

    # -- tables -----------------------------------------------------------------------------------------------
    def rotary(self, rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._rot is None or self._rot[0].shape[0] < rows:
            n = max(rows, 4096)
            cs = torch.empty(n, 64, dtype=torch.float32, device=self.device)
            sn = torch.empty_like(cs)
            self.call("srb_rotary_table", P(self.w.inv_freq), n, P(cs), P(sn))
            # captured graphs hold raw pointers into earlier tables: those stay alive
            self._rot_keep = getattr(self, "_rot_keep", []) + [(cs, sn)]
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
            self.call("srb_time_cond_table", P(t_dev), nfe, P(self.w.four_w), P(self.w.lin_w), P(self.w.lin_b),
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
        self.call("srb_duration_predict", P(ids), P(self.w.dur_table), self.w.dur_bias, P(dur), P(tot), b, n,
                 self.w.dur_table.shape[1])
        totals = tot.cpu()
        all_one = int(totals.sum()) == 0          # the regulator's all-zero rule (HF:113-114)
        n_out = n if all_one else int(totals.max())
        out = torch.empty(b, n_out, dtype=torch.int64, device=self.device)
        self.call("srb_length_regulate", P(ids), P(dur), P(out), b, n, n_out, 1 if all_one else 0)
        return out, dur

    # -- workspace --------------------------------------------------------------------------------------------
    def workspace(self, batch: int, frames: int, mel_rows: Optional[int] = None, carver: Optional[Carver] = None
                  ) -> Dict[str, object]:
        """Buffers of one (batch, frames) shape.  `frames` must be a multiple of 8 (see `padded_frames`): the
        transposed-V operand of the attention kernel is addressed by TMA per utterance and TMA needs 16-byte aligned box
        origins.  `mel_rows` (<= frames) is the caller's frame count: the mel outputs are compact (batch, mel_rows, 80).
        With a `carver` the tensors are views of the arena (a measuring carver returns None entries); without one they are
        stand-alone allocations (tests, tools)."""
        assert frames % 8 == 0, "CFMSampler.workspace: frames must be padded to a multiple of 8"
        mel_rows = frames if mel_rows is None else mel_rows
        assert 0 < mel_rows <= frames
        dev, m = self.device, batch * frames
        m_pad = (m + 255) // 256 * 256
        c = carver if carver is not None else Carver(None, dev)
        if carver is None:
            take = lambda shape, dtype: torch.empty(*shape, dtype=dtype, device=dev)
        else:
            take = c.take
        f32, b16 = torch.float32, torch.bfloat16
        S = self.split      # bf16 activation tensors are S times as wide in the tight-precision format
        ws: Dict[str, object] = dict(
            frames=frames, mel_rows=mel_rows, batch=batch, qk_calls=0,
            ids=take((batch, frames), torch.int64), lengths=take((batch,), torch.int32),
            cond=take((m, 256), f32), xt=take((batch, frames, 80), f32), xt_b=take((batch, frames, 80 * S), b16),
            x0=take((m, 256), f32), x=take((m, 256), f32),
            # xn is also the B operand of the V^T GEMM, read in 256-row tiles: rows >= m are cleared by the staging kernel
            xn=take((m_pad, 256 * S), b16),
            o=take((m, 256 * S), b16), h=take((m, 896 * S), b16),
            mel=take((batch, mel_rows, 80), f32), mel_b=take((batch, mel_rows, 80 * S), b16),
            # max |q|^2, |k|^2 per (utterance, head), recorded by qk_rope, read by the attention kernel; two buffers used
            # alternately by successive layers (each projection clears the other one for its successor)
            qkmax=take((2, batch, 2, 2, 2), f32),
        )
        if self.tight:
            ws["qkv"] = take((m, 768 * S), b16)      # [q | k | v] after rotary, read by the fp32 attention kernel
        else:
            ws["qk"], ws["vt"] = take((m, 512), b16), take((256, m_pad), b16)
        return ws

    # -- the loop ---------------------------------------------------------------------------------------------
    def stage(self, ws: Dict[str, object], input_ids: torch.Tensor, noise: torch.Tensor,
              truncation: Optional[float]) -> None:
        """Per-call staging, ahead of the (captured) loop: the caller's ids / prior sample go into the static buffers
        (rows padded to the workspace's frame count), the prior is clamped (models.py:169-170) and the zero regions the
        loop relies on are re-established (the arena is shared between shapes)."""
        b, n = input_ids.shape
        n8 = ws["frames"]
        assert b == ws["batch"] and n == ws["mel_rows"] and tuple(noise.shape) == (b, n, 80)
        assert noise.dtype == torch.float32 and input_ids.dtype == torch.int64
        xn, m = ws["xn"], b * n8
        tail = xn[m:]
        self.call("srb_stage_inputs", P(input_ids.contiguous()), P(noise.contiguous()), P(ws["ids"]), P(ws["xt"]), P(ws["xt_b"]),
                 P(tail) if tail.numel() else None, tail.numel() * 2, P(ws["qkmax"]), ws["qkmax"].numel() * 4, b, n, n8,
                 float(truncation) if truncation is not None else 0.0, 0 if truncation is None else 1)
        ws["qk_calls"] = 0

    def prepare(self, ws: Dict[str, object]) -> None:
        """mask/lengths (models.py:152) and the hoisted conditioning gather (:154,:175-176)."""
        b, n = ws["ids"].shape
        self.call("srb_unit_lengths", P(ws["ids"]), P(ws["lengths"]), b, n)
        self.call("srb_embed_gather", P(self.w.cond_table), P(ws["ids"]), P(ws["cond"]), b * n,
                 self.w.cond_table.shape[0], 256, nbytes=b * n * (2 * 256 * 4 + 8))
        ws["qk_calls"] = 0

    def step(self, ws: Dict[str, object], g_step: torch.Tensor, dt: float, last: bool) -> None:
        """One velocity evaluation + Euler update (models.py:173-184); `last` adds :186-187."""
        b, n = ws["ids"].shape
        w, L = self.w, ws["lengths"]
        cs, sn = self.rotary(n)
        m = b * ws["mel_rows"]      # algorithmic work (profiling log only): the caller's frames, not the 8-row padding
        self.call("srb_cfm_embed", P(ws["xt_b"]), P(w.w_embed), P(ws["cond"]), P(ws["x0"]), b, n, flops=2.0 * m * 80 * 256,
                 nbytes=m * (80 * 2 + 256 * 8))
        self.call("srb_cfm_posconv_norm", P(ws["x0"]), P(w.dw_w), P(w.dw_b), P(g_step[0]), P(L), P(ws["x"]), P(ws["xn"]), b, n,
                 flops=2.0 * m * 31 * 256, nbytes=m * 256 * (4 + 4 + 2))
        for i in range(self.depth):
            qk_cur, qk_next = ws["qkmax"][ws["qk_calls"] % 2], ws["qkmax"][(ws["qk_calls"] + 1) % 2]
            ws["qk_calls"] += 1
            if self.tight:
                # reference-style chain: one q|k|v projection with rotary, then fp32 softmax attention on CUDA cores
                self.call("srb_cfm_qkv_rope", P(ws["xn"]), P(w.w_qkv[i]), P(cs), P(sn), P(ws["qkv"]), None, None, b, n,
                          flops=2.0 * m * 256 * 768)
                self.call("srb_cfm_attention_simt", P(ws["qkv"]), P(L), P(ws["o"]), b, n, flops=4.0 * m * ws["mel_rows"] * 256)
            elif self.fused_qkv:
                m_pad = ws["vt"].shape[1]
                # the whole to_qkv GEMM in one launch: q | k with rotary, v stored transposed by the epilogue
                self.call("srb_cfm_qk_rope_vt", P(ws["xn"]), P(w.w_qkv[i]), P(cs), P(sn), P(ws["qk"]), P(ws["vt"]), m_pad,
                         P(qk_cur), P(qk_next), b, n, flops=2.0 * m * 256 * 768)
            else:
                # q|k projection and the transposed-v projection both read xn only: two parallel graph branches
                m_pad = ws["vt"].shape[1]
                self.fork.run([
                    lambda: self.call("srb_cfm_qk_rope", P(ws["xn"]), P(w.w_qkv[i]), P(cs), P(sn), P(ws["qk"]),
                                     P(qk_cur), P(qk_next), b, n, flops=2.0 * m * 256 * 512),
                    lambda: self.call("srb_cfm_v_transposed", P(ws["xn"]), P(w.w_qkv[i][512:]), P(ws["vt"]), m_pad,
                                     flops=2.0 * m * 256 * 256),
                ])
            if not self.tight:
                self.call("srb_cfm_attention_tc", P(ws["qk"]), 512, P(ws["vt"]), ws["vt"].shape[1], P(L), P(qk_cur), P(ws["o"]), b, n,
                          flops=4.0 * m * ws["mel_rows"] * 256)
            self.call("srb_cfm_attn_out_norm", P(ws["o"]), P(w.w_out[i]), P(g_step[2 * i + 1]), P(L), P(ws["x"]), P(ws["xn"]), b, n,
                     flops=2.0 * m * 256 * 256)
            # every utterance ends in a pad row whenever the caller's frame count is not a multiple of 8 (lengths <= mel_rows < n)
            self.call("srb_cfm_ffn_glu", P(ws["xn"]), P(w.w_ff1[i]), P(w.b_ff1[i]), P(L), P(ws["h"]), b, n,
                     1 if ws["mel_rows"] < n else 0,
                     flops=2.0 * m * 768 * 1792)
            if i + 1 < self.depth:
                g_next, mode = g_step[2 * i + 2], 1
            else:
                g_next, mode = w.final_norm_w, 2
            self.call("srb_cfm_ffn_out_norm", P(ws["h"]), P(w.w_ff2[i]), P(w.b_ff2[i]), P(g_next), mode, P(L), P(ws["x"]),
                     P(ws["xn"]), b, n, flops=2.0 * m * 2688 * 256)
        mel = P(ws["mel"]) if last else None
        mel_b = P(ws["mel_b"]) if last else None
        self.call("srb_cfm_pred_euler", P(ws["xn"]), P(w.w_pred), float(dt), P(ws["xt"]), P(ws["xt_b"]), mel, mel_b,
                 ws["mel_rows"], self.std, self.mean, pad_value_f32(), P(L), b, n, flops=2.0 * m * 256 * 80,
                 nbytes=m * (256 * 2 + 80 * (4 + 4 + 2)))

    def run(self, ws: Dict[str, object], dt: float) -> None:
        """ids and the (clamped) prior sample must already be staged (`stage`); result lands in ws['mel'], ws['mel_b']."""
        times = ode_times(dt)
        g = self.cond_table(times)
        self.prepare(ws)
        nfe = len(times)
        for s in range(nfe):
            self.step(ws, g[s], dt, last=(s == nfe - 1))


# ------------------------------------------------------------------------------------------------------ vocoder
class HifiGanGenerator:
    """mel (B, T, 80) bf16 -> waveform (B, 320 T + 80) fp32 (kernels: srb_hifigan_*)."""

    def __init__(self, packed: PackedVocoder, slope: float = 0.1, fuse_mrf: bool = True, tight: bool = False):
        self.w = packed
        self.slope = float(slope)
        # tight=True: the tight-precision library (see CFMSampler); every stage runs conv by conv there
        self.tight = tight
        self.split = 3 if tight else 1
        self.fuse_mrf = fuse_mrf and not tight
        # SRB_PAIR_UPSAMPLE=0 runs every up-sampler through the polyphase kernel instead (A/B runs)
        self.pair_upsample = os.environ.get("SRB_PAIR_UPSAMPLE", "1") != "0" and not tight
        self.fork = _Fork(packed.w_pre.device, 2)
        self.device = packed.w_pre.device
        # One copy of every tensor of the unfused resblock chains: only the ACTIVATED tensor (what the next conv reads) is
        # stored; the residual adds recover the raw value from it in the epilogue (srb_hifigan_conv_res_act: leaky_relu is
        # invertible, and the recovered value is as accurate as a bf16 copy of the raw one).  These launches run at the
        # HBM roofline, so a third less traffic is a third less time.  SRB_SINGLE_COPY=0 keeps both copies (A/B knob).
        self.single_copy = os.environ.get("SRB_SINGLE_COPY", "1") != "0" and not tight and self.slope > 0
        # The first two (conv1, conv2) pairs of the k = 3 resblock of the C = 64 stage as ONE launch each
        # (srb_hifigan_pair_fused: the conv1 output never leaves the SM; two tensor passes instead of five).  Needs the
        # single-copy form.  Measured at config 2: 232-246 us per pair against 110 + 178 = 288 us unfused; the k = 7 pair
        # is slower fused (410-424 us against 140 + 203 = 343: its 56 MMAs per tile run serially with the rest of the tile's
        # shared-memory traffic) and keeps its two launches.  SRB_PAIR_FUSED=0: never, =2: k = 7 too (A/B knob).
        pf = os.environ.get("SRB_PAIR_FUSED", "1")
        self.pair_fused_kernels = () if (pf == "0" or not self.single_copy) else ((3, 7) if pf == "2" else (3,))

    def call(self, name: str, *args, **kw) -> None:
        nat.call(name, *args, tight=self.tight, **kw)

    def workspace(self, batch: int, frames: int, carver: Optional[Carver] = None) -> Dict[str, object]:
        """Stage tensors of one (batch, frames) shape.  A stage run by the fused MRF kernel needs only the up-sampler
        output and the stage output; the others add the activated copy and nine conv intermediates."""
        dev, S = self.device, self.split
        # (bf16 activation tensors are S times as wide in the tight-precision format)
        if carver is None:
            take = lambda shape: torch.empty(*shape[:-1], shape[-1] * S, dtype=torch.bfloat16, device=dev)
        else:
            take = lambda shape: carver.take((*shape[:-1], shape[-1] * S), torch.bfloat16)
        ws: Dict[str, object] = {"batch": batch, "frames": frames, "pre": take((batch, frames, 512))}
        rows, c = frames, 512
        stages = []
        for i, (k, s) in enumerate(zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES)):
            rows = (rows - 1) * s - 2 * ((k - s) // 2) + k
            c //= 2
            fused = self.fuse_mrf and i in self.w.w_mrf
            st = dict(rows=rows, c=c, out=take((batch, rows, c)))
            if fused or not self.single_copy:
                st["u_raw"] = take((batch, rows, c))
            if not fused:
                st.update(u_act=take((batch, rows, c)), t=[take((batch, rows, c)) for _ in range(3)],
                          xa=[take((batch, rows, c)) for _ in range(3)])
                if not self.single_copy:
                    st["xr"] = [take((batch, rows, c)) for _ in range(3)]
            stages.append(st)
        ws["stages"] = stages
        ws["rows"] = rows
        if carver is None:
            ws["wav"] = torch.empty(batch, rows, dtype=torch.float32, device=dev)
        else:
            ws["wav"] = carver.take((batch, rows), torch.float32)
        return ws

    def run(self, mel_b: torch.Tensor, ws: Dict[str, object]) -> torch.Tensor:
        """Everything up to (not including) conv_post: returns the leaky_relu(0.01)'ed last stage (B, 320 T + 80, 16)."""
        b, t, _ = mel_b.shape
        assert mel_b.shape[2] == 80 * self.split
        w = self.w
        one = _i32([7])
        dil1 = _i32([1])
        # conv_pre (HF:1470); its only consumer is leaky_relu -> upsampler, so only the activated copy is stored
        self.call("srb_hifigan_conv", P(mel_b), None, None, 1, one, dil1, P(w.w_pre), P(w.b_pre), None, None, None, None,
                 P(ws["pre"]), b, t, 80, 512, 1.0, self.slope, flops=2.0 * b * t * 7 * 80 * 512)
        x_act, rows_in, c_in = ws["pre"], t, 512
        n_stage = len(UPSAMPLE_RATES)
        for i, (k, s) in enumerate(zip(UPSAMPLE_KERNELS, UPSAMPLE_RATES)):
            st = ws["stages"][i]
            rows, c = st["rows"], st["c"]
            fused = self.fuse_mrf and i in w.w_mrf
            single = self.single_copy and not fused
            u_raw = None if single else P(st["u_raw"])
            # upsampler (HF:1472-1473): raw copy = residual of the three resblocks, activated copy = their input
            # (single-copy form: the activated copy serves as both)
            if i in w.up_pair and self.pair_upsample:
                # L_out = s L: all s output phases of an input row from one 3-tap conv (packing.upsampler_as_row_group_conv);
                # (B, rows_in, s c) is the (B, rows, c) result.  FLOPs reported are the transposed conv's own.
                wp, bp = w.up_pair[i]
                self.call("srb_hifigan_conv", P(x_act), None, None, 1, _i32([3]), dil1, P(wp), P(bp), None, None, None,
                         u_raw, None if fused else P(st["u_act"]), b, rows_in, c_in, s * c, 1.0, self.slope,
                         flops=2.0 * b * rows_in * k * c_in * c)
            else:
                self.call("srb_hifigan_upsample", P(x_act), P(w.w_up[i]), P(w.b_up[i]), u_raw,
                         None if fused else P(st["u_act"]), b, rows_in, c_in, c, k, s, self.slope,
                         flops=2.0 * b * rows_in * k * c_in * c)
            if fused:
                # narrow stages: the whole MRF block (18 convs + residuals + mean + next leaky_relu) in one kernel
                slope_next = self.slope if i + 1 < n_stage else 0.01
                self.call("srb_hifigan_mrf_fused", P(st["u_raw"]), P(w.w_mrf[i]), P(w.b_mrf[i]), P(st["out"]), b, rows, c,
                         self.slope, slope_next, flops=252.0 * c * c * b * rows)
                x_act, rows_in, c_in = st["out"], rows, c
                continue
            res: List[Optional[torch.Tensor]] = [None, None, None]
            srcs: List[torch.Tensor] = list(st["t"])       # the tail's conv inputs (the last conv1 output of each chain)

            res_slope = self.slope if single else 0.0

            def chain(j: int, rk: int, st=st, rows=rows, c=c, i=i, single=single, res_slope=res_slope) -> None:
                kk = _i32([rk])
                xr, xa = (st["u_act"], st["u_act"]) if single else (st["u_raw"], st["u_act"])
                if single and c == 64 and rk in self.pair_fused_kernels:
                    # pairs 0 and 1 fused: u_act -> xa[j] -> t[j]; then the last conv1 t[j] -> xa[j] (the tail's source) with
                    # t[j] as the tail's residual
                    bufs = (st["u_act"], st["xa"][j], st["t"][j])
                    for q in range(2):
                        self.call("srb_hifigan_pair_fused", P(bufs[q]), P(w.w_c1[i][j][q]), P(w.b_c1[i][j][q]), P(w.w_c2[i][j][q]),
                                 P(w.b_c2[i][j][q]), P(bufs[q + 1]), b, rows, c, rk, RESBLOCK_DILATIONS[q], self.slope,
                                 flops=4.0 * b * rows * rk * c * c)
                    self.call("srb_hifigan_conv", P(st["t"][j]), None, None, 1, kk, _i32([RESBLOCK_DILATIONS[2]]),
                             P(w.w_c1[i][j][2]), P(w.b_c1[i][j][2]), None, None, None, None, P(st["xa"][j]), b, rows, c, c, 1.0,
                             self.slope, flops=2.0 * b * rows * rk * c * c)
                    res[j] = st["t"][j]
                    srcs[j] = st["xa"][j]
                    return
                for q, dil in enumerate(RESBLOCK_DILATIONS):
                    # conv1 with dilation (HF:1361-1363), output only needed activated
                    self.call("srb_hifigan_conv", P(xa), None, None, 1, kk, _i32([dil]), P(w.w_c1[i][j][q]),
                             P(w.b_c1[i][j][q]), None, None, None, None, P(st["t"][j]), b, rows, c, c, 1.0, self.slope,
                             flops=2.0 * b * rows * rk * c * c)
                    if q < 2 or self.tight:
                        # conv2 + residual (HF:1364-1366): raw (next residual) and activated (next conv1 input)
                        self.call("srb_hifigan_conv_res_act", P(st["t"][j]), None, None, 1, kk, dil1, P(w.w_c2[i][j][q]),
                                 P(w.b_c2[i][j][q]), P(xr), None, None, res_slope, None if single else P(st["xr"][j]),
                                 P(st["xa"][j]) if q < 2 else None, b, rows, c, c, 1.0, self.slope,
                                 flops=2.0 * b * rows * rk * c * c)
                        xr, xa = (st["xa"][j], st["xa"][j]) if single else (st["xr"][j], st["xa"][j])
                res[j] = xr

            # the three resblocks of a stage are independent until the MRF mean: three parallel graph branches
            # (longest chain, k = 11, on the main stream)
            self.fork.run([lambda: chain(2, RESBLOCK_KERNELS[2]), lambda: chain(1, RESBLOCK_KERNELS[1]),
                           lambda: chain(0, RESBLOCK_KERNELS[0])])
            # fused MRF tail: the three last conv2's + their residuals + mean (HF:1475-1478) + the next leaky_relu
            # (slope 0.1 before an upsampler, torch default 0.01 before conv_post, HF:1480)
            slope_next = self.slope if i + 1 < n_stage else 0.01
            if self.tight:
                # every resblock finished conv by conv above: the MRF mean + next leaky_relu as its own kernel
                self.call("srb_hifigan_mean3", P(res[0]), P(res[1]), P(res[2]), P(st["out"]), b * rows, c, 1.0 / 3.0, slope_next)
                x_act, rows_in, c_in = st["out"], rows, c
                continue
            self.call("srb_hifigan_conv_res_act", P(srcs[0]), P(srcs[1]), P(srcs[2]), 3, _i32(list(RESBLOCK_KERNELS)),
                     _i32([1, 1, 1]), P(w.w_tail[i]), P(w.b_tail[i]), P(res[0]), P(res[1]), P(res[2]), res_slope, None, P(st["out"]),
                     b, rows, c, c, 1.0 / 3.0, slope_next, flops=2.0 * b * rows * sum(RESBLOCK_KERNELS) * c * c)
            x_act, rows_in, c_in = st["out"], rows, c
        return x_act

    def post(self, x_act: torch.Tensor, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> None:
        """conv_post + tanh (HF:1480-1482).  lengths=None: dense (B, rows) output; else ragged (see srb_hifigan_post)."""
        b, rows, _ = x_act.shape      # (B, rows, 16 * split)
        self.call("srb_hifigan_post", P(x_act), P(self.w.w_post), self.w.b_post, P(wav), b, rows,
                 P(lengths) if lengths is not None else None, flops=2.0 * b * rows * 7 * 16, nbytes=b * rows * (16 * 2 + 4))


# ------------------------------------------------------------------------------------------------------ engine
@dataclass
class _Plan:
    key: Tuple
    generation: int
    cfm_ws: Dict[str, object]
    voc_ws: Dict[str, object]
    body: Callable[[], None]
    graph: Optional[torch.cuda.CUDAGraph] = None
    n_launches: int = 0
    uses: int = 0
    x_last: Optional[torch.Tensor] = None     # last stage's activated output (input of conv_post)
    extra: Dict[str, object] = field(default_factory=dict)


def build_sampler(state_dict: Dict[str, torch.Tensor], device, depth: int = 4, mean: float = -5.8843,
                  std: float = 2.2615, tight: bool = False) -> CFMSampler:
    """state_dict uses the top-level key names ("model.*").  tight=True: the tight-precision library."""
    nat.require_blackwell()
    device = torch.device(device)
    with torch.cuda.device(device):
        return CFMSampler(pack_cfm(state_dict, device, depth=depth, split=tight), depth, mean, std, tight=tight)


def build_vocoder(state_dict: Dict[str, torch.Tensor], device, slope: float = 0.1, tight: bool = False) -> HifiGanGenerator:
    """state_dict uses the top-level key names ("vocoder.*").  tight=True: the tight-precision library."""
    nat.require_blackwell()
    device = torch.device(device)
    with torch.cuda.device(device):
        return HifiGanGenerator(pack_vocoder(state_dict, device, split=tight), slope, tight=tight)


class ResynthEngine:
    """units -> waveform for one device.

    use_graphs: "auto" (default; a shape's first call is eager, later calls replay a CUDA graph), True (capture at the
    first call), False (always eager).  max_graphs bounds the LRU of captured shapes."""

    def __init__(self, sampler: Optional[CFMSampler], vocoder: Optional[HifiGanGenerator], use_graphs="auto",
                 max_graphs: int = 48):
        nat.require_blackwell()
        self.sampler = sampler
        self.vocoder = vocoder
        self.device = (sampler or vocoder).device
        self.tight = (sampler or vocoder).tight
        assert sampler is None or vocoder is None or sampler.tight == vocoder.tight, "sampler and vocoder precision modes differ"
        env = os.environ.get("SRB_GRAPHS")
        if env is not None:
            use_graphs = {"0": False, "1": True}.get(env, "auto")
        self.use_graphs = use_graphs
        self.max_graphs = max_graphs
        self.arena = Arena.get(self.device)
        self._plans: "OrderedDict[Tuple, _Plan]" = OrderedDict()
        self.stats = {"graphs_captured": 0, "eager_runs": 0, "graph_replays": 0, "plans_built": 0, "plans_evicted": 0}

    # -- memory -----------------------------------------------------------------------------------------------
    def _layout(self, batch: int, frames: int, with_cfm: bool, with_vocoder: bool, carver: Carver):
        n8 = padded_frames(frames)
        cfm_ws = self.sampler.workspace(batch, n8, mel_rows=frames, carver=carver) if with_cfm else {}
        voc_ws: Dict[str, object] = {}
        if with_vocoder:
            voc_ws = self.vocoder.workspace(batch, frames, carver=carver)
            if not with_cfm:
                voc_ws["mel_b"] = carver.take((batch, frames, 80 * self.vocoder.split), torch.bfloat16)
        return cfm_ws, voc_ws

    def workspace_bytes(self, batch: int, frames: int, with_cfm: bool = True, with_vocoder: bool = True) -> int:
        c = Carver(None, self.device)
        self._layout(batch, frames, with_cfm, with_vocoder, c)
        return c.off

    def reserve(self, batch: int, frames: int, with_vocoder: bool = True) -> int:
        """Size the arena for a (batch, frames) shape up front (a driver that knows its largest bucket calls this once
        so that no later shape has to grow the arena and drop the captured graphs).  Returns the arena size."""
        with torch.cuda.device(self.device):
            self.arena.ensure(self.workspace_bytes(batch, frames, self.sampler is not None, with_vocoder and self.vocoder is not None))
        return self.arena.nbytes

    # -- plans ------------------------------------------------------------------------------------------------
    def _plan(self, batch: int, frames: int, dt: Optional[float], with_cfm: bool, with_vocoder: bool) -> _Plan:
        key = (batch, frames, None if dt is None else float(dt), with_cfm, with_vocoder)
        plan = self._plans.get(key)
        if plan is not None and plan.generation == self.arena.generation:
            self._plans.move_to_end(key)
            return plan
        self.arena.ensure(self.workspace_bytes(batch, frames, with_cfm, with_vocoder))
        # if the arena moved (now or through another engine on this device), older plans point into freed memory
        for k in [k for k, p in self._plans.items() if p.generation != self.arena.generation]:
            self._drop(k)
        cfm_ws, voc_ws = self._layout(batch, frames, with_cfm, with_vocoder, Carver(self.arena.buf, self.device))
        if with_cfm:
            self.sampler.cond_table(ode_times(dt))
            self.sampler.rotary(cfm_ws["frames"])
        plan = _Plan(key, self.arena.generation, cfm_ws, voc_ws, body=lambda: None)

        def body():
            if with_cfm:
                self.sampler.run(cfm_ws, dt)
            if with_vocoder:
                plan.x_last = self.vocoder.run(cfm_ws["mel_b"] if with_cfm else voc_ws["mel_b"], voc_ws)

        plan.body = body
        self._plans[key] = plan
        self.stats["plans_built"] += 1
        while len(self._plans) > self.max_graphs:
            self._drop(next(iter(self._plans)))
        return plan

    def _drop(self, key) -> None:
        plan = self._plans.pop(key, None)
        if plan is not None:
            if plan.graph is not None:
                plan.graph.reset()
            self.stats["plans_evicted"] += 1

    def _launch(self, plan: _Plan) -> None:
        plan.uses += 1
        if plan.graph is None and (self.use_graphs is True or (self.use_graphs == "auto" and plan.uses >= 2)):
            # the previous (eager) call of this shape -- or, with use_graphs=True, the run below -- has configured the
            # kernels' attributes; capture allocates nothing
            if plan.uses == 1:
                plan.body()
                torch.cuda.current_stream().synchronize()
            graph = torch.cuda.CUDAGraph()
            c0 = nat.launch_count
            with torch.cuda.graph(graph):
                plan.body()
            plan.n_launches = nat.launch_count - c0   # recorded, not executed: counted again at every replay
            nat.launch_count = c0
            plan.graph = graph
            self.stats["graphs_captured"] += 1
        if plan.graph is not None:
            plan.graph.replay()
            nat.launch_count += plan.n_launches
            self.stats["graph_replays"] += 1
        else:
            plan.body()
            self.stats["eager_runs"] += 1

    def _run(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float], noise: Optional[torch.Tensor],
             with_vocoder: bool) -> _Plan:
        assert input_ids.dim() == 2 and input_ids.dtype == torch.int64 and input_ids.is_cuda
        b, n = input_ids.shape
        with torch.cuda.device(self.device):
            plan = self._plan(b, n, dt, True, with_vocoder)
            if noise is None:
                # same call as the reference (models.py:168) so a seeded run draws the same prior on the same device
                noise = torch.randn(b, n, 80, device=self.device)
            self.sampler.stage(plan.cfm_ws, input_ids, noise, truncation)
            self._launch(plan)
        return plan

    def lengths_async(self, input_ids: torch.Tensor):
        """Valid-frame counts (and extents, see srb_unit_extents) of a batch on their way to the host, enqueued BEFORE the
        big graph so that the caller can wait for them (event) while the graph is still running -- the public forward
        then never drains the GPU between calls.  The kernel writes straight into pinned host memory (device-accessible
        under unified addressing): a D2H copy would queue on the copy engine behind the previous call's waveform
        read-back and hold up the graph enqueued after it (measured: +0.8 ms per call).
        Returns (pinned int32 (2, B) tensor [counts; extents], event)."""
        b, n = input_ids.shape
        slot = getattr(self, "_len_slot", 0)
        self._len_slot = (slot + 1) % 4
        bufs = getattr(self, "_len_bufs", None)
        if bufs is None or bufs[0].shape[1] < b:
            bufs = [torch.zeros(2, max(b, 64), dtype=torch.int32).pin_memory() for _ in range(4)]
            self._len_bufs = bufs
        host_buf = bufs[slot]
        with torch.cuda.device(self.device):
            nat.call("srb_unit_extents", P(input_ids.contiguous()), host_buf[0].data_ptr(), host_buf[1].data_ptr(), b, n,
                     tight=self.tight)
            ev = torch.cuda.Event()
            ev.record()
        return host_buf[:, :b], ev

    def sample(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float],
               noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ConditionalFlowMatchingModel.sample: (B, N) int64 -> mel (B, N, 80) fp32 (a fresh tensor)."""
        plan = self._run(input_ids, dt, truncation, noise, with_vocoder=False)
        return plan.cfm_ws["mel"].clone()

    def resynthesize(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float],
                     noise: Optional[torch.Tensor] = None):
        """Returns (wav (B, 320 N + 80) fp32, lengths (B,) int32 device tensor, mel (B, N, 80)) -- all three are views
        of the arena, valid until the next call on this device."""
        plan = self._run(input_ids, dt, truncation, noise, with_vocoder=True)
        with torch.cuda.device(self.device):
            self.vocoder.post(plan.x_last, plan.voc_ws["wav"])
        return plan.voc_ws["wav"], plan.cfm_ws["lengths"], plan.cfm_ws["mel"]

    def resynthesize_ragged(self, input_ids: torch.Tensor, dt: float, truncation: Optional[float], total_samples: int,
                            noise: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The public forward's form: a 1-D tensor (fresh, or the caller's `out`) holding every utterance cropped to
        320 len + 80 samples, back to back (the crop loop of models.py:252-256 happens in the last kernel's store).
        `total_samples` = sum of those lengths (the caller knows the valid-frame counts on the host)."""
        plan = self._run(input_ids, dt, truncation, noise, with_vocoder=True)
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(total_samples, dtype=torch.float32, device=self.device)
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= total_samples
            self.vocoder.post(plan.x_last, out, plan.cfm_ws["lengths"])
        return out

    def vocode(self, mel: torch.Tensor) -> torch.Tensor:
        """decoder.vocoder(mel): (B, T, 80) float -> (B, 320 T + 80) fp32 (a fresh tensor)."""
        squeeze = mel.dim() == 2
        if squeeze:
            mel = mel.unsqueeze(0)
        b, t, _ = mel.shape
        with torch.cuda.device(self.device):
            plan = self._plan(b, t, None, False, True)
            mel32 = mel.to(device=self.device, dtype=torch.float32).contiguous()
            # fp32 -> bf16 (the reference's autocast would do the same cast at conv_pre)
            nat.call("srb_prior_prepare", P(mel32), P(plan.voc_ws["mel_b"]), mel32.numel(), 0.0, 0, tight=self.tight)
            self._launch(plan)
            wav = torch.empty(b, plan.voc_ws["rows"], dtype=torch.float32, device=self.device)
            self.vocoder.post(plan.x_last, wav)
        return wav[0] if squeeze else wav
