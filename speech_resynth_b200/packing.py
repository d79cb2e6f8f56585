"""Load-time repacking of the reference state-dict into the layouts the sm_100a kernels consume.

Runs once per model (``.cuda()`` / first call).  GEMM operands become bf16 ``[n][k]`` matrices with ``k`` ordered
(tap, channel) so a conv tap is a contiguous K slab; everything read by CUDA-core epilogues stays fp32.
Shapes/keys: SURVEY.md section 8(b).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

import torch

UPSAMPLE_RATES = (5, 4, 4, 2, 2)
UPSAMPLE_KERNELS = (10, 9, 8, 4, 4)
RESBLOCK_KERNELS = (3, 7, 11)
RESBLOCK_DILATIONS = (1, 3, 5)
FUSED_MRF_CHANNELS = (16, 32)     # stages whose whole MRF block runs in srb_hifigan_mrf_fused


def block_k_for(c_in: int) -> int:
    """K block (channels per TMA box) the conv-GEMM core uses for an input width (mirrors srb_convgemm.cu)."""
    return 64 if c_in >= 64 else (32 if c_in >= 32 else 16)


def split_operand(w: torch.Tensor, dim: int) -> torch.Tensor:
    """Tight-precision (split) packing of a GEMM weight along its input-channel dimension `dim`: [Wh | Wh | Wl] with
    Wh = bf16(W), Wl = bf16(W - Wh).  Activations travel as [xh | xl | xh] (include/srb.h: srb_split_factor), so the
    bf16 tensor-core loop accumulates xh Wh + xl Wh + xh Wl in fp32.  Returned in fp32 (every value is a bf16 number)."""
    w = w.float()
    hi = w.to(torch.bfloat16).float()
    lo = (w - hi).to(torch.bfloat16).float()
    return torch.cat([hi, hi, lo], dim=dim)


def pack_conv_weight(w: torch.Tensor, block_k: int) -> torch.Tensor:
    """Conv1d weight (C_out, C_in, k) -> bf16 [C_out][k * C_in_pad], k-major then channel, channels zero-padded."""
    c_out, c_in, k = w.shape
    c_pad = (c_in + block_k - 1) // block_k * block_k
    p = torch.zeros(c_out, k, c_pad, dtype=torch.float32, device=w.device)
    p[:, :, :c_in] = w.float().permute(0, 2, 1)
    return p.reshape(c_out, k * c_pad).to(torch.bfloat16).contiguous()


def pack_upsampler_weight(w: torch.Tensor, stride: int, split: bool = False) -> torch.Tensor:
    """ConvTranspose1d weight (C_in, C_out, k) -> polyphase bf16 [C_out][k * C_in].

    Phase r (output rows q*stride + r) uses taps j = j0 + m*stride, j0 = (r + pad) % stride, in increasing m
    (the kernel reads input row q + (r + pad)//stride - m for tap m); phases are concatenated along K.
    """
    c_in, c_out, k = w.shape
    pad = (k - stride) // 2
    cols: List[torch.Tensor] = []
    for r in range(stride):
        j0 = (r + pad) % stride
        for j in range(j0, k, stride):
            tap = w[:, :, j].float().t()  # (C_out, C_in)
            cols.append(split_operand(tap, 1) if split else tap)
    return torch.cat(cols, dim=1).to(torch.bfloat16).contiguous()


def upsampler_as_row_group_conv(w: torch.Tensor, bias: torch.Tensor, stride: int):
    """ConvTranspose1d (C_in, C_out, k) with k - 2 pad = stride (HF:1392-1402: (8, 4, 2) and (4, 2, 1)) as an ordinary
    3-tap Conv1d that produces ALL `stride` output phases of an input row at once:
        out[stride r + ph, co] = y[r, ph * C_out + co],   y[r] = Wv[:, :, 0] x[r-1] + Wv[:, :, 1] x[r] + Wv[:, :, 2] x[r+1]
    (zero padding = the transposed conv's skipped out-of-range inputs).  From o = stride i - pad + j: tap j feeds phase ph
    iff (ph + pad - j) % stride == 0, reading x[r + d] with d = (ph + pad - j) / stride in {-1, 0, 1}.
    The (B, L, stride C_out) result IS the (B, stride L, C_out) up-sampled tensor (L_out = stride L exactly), so the
    launch writes whole contiguous rows and issues N = stride C_out MMAs (narrow tiles pay per MMA, DESIGN.md section
    5.2) at the price of some zero blocks.  Returns (Conv1d weight (stride C_out, C_in, 3), bias (stride C_out,))."""
    c_in, c_out, k = w.shape
    pad = (k - stride) // 2
    assert k - 2 * pad == stride, "row-group form needs L_out = stride * L"
    wv = torch.zeros(stride * c_out, c_in, 3, dtype=torch.float32, device=w.device)
    wt = w.float()
    for ph in range(stride):
        for j in range(k):
            if (ph + pad - j) % stride == 0:
                d = (ph + pad - j) // stride
                assert -1 <= d <= 1
                wv[ph * c_out:(ph + 1) * c_out, :, d + 1] = wt[:, :, j].t()
    return wv, bias.float().repeat(stride).contiguous()


def pack_operand_taps(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight (C_out, C_in, k) -> bf16 [k][C_in/8][C_out][8]: per tap the B operand of the fused-MRF kernel
    in the un-swizzled K-major layout (16-byte K chunks, rows 16 bytes apart)."""
    c_out, c_in, k = w.shape
    t = w.float().permute(2, 1, 0).reshape(k, c_in // 8, 8, c_out).permute(0, 1, 3, 2)
    return t.to(torch.bfloat16).contiguous()


def mrf_phases(channels: int) -> int:
    """Time phases the fused-MRF kernel uses for this width (srb_hifigan_mrf_phases): 2 at C = 16, else 1."""
    from . import _native as nat

    return int(nat.load().srb_hifigan_mrf_phases(int(channels)))


def pack_mrf_conv(w: torch.Tensor, dilation: int, phases: int) -> torch.Tensor:
    """One conv of the fused-MRF kernel, flat bf16 in tcgen05 operand layout.

    Dilated convs (and every conv when phases == 1): the k tap matrices of pack_operand_taps.
    Dilation-1 convs with D = phases > 1: k - 1 + D "phase matrices" [D*C_out][C_in]; matrix pi (input time offset
    o = pi - (k-1)/2 relative to the D-row group) holds, in row block d', tap pi - d' (zero when outside [0, k)):
    out[q D + d'] = sum_tap W[tap] x[q D + d' + tap - (k-1)/2]."""
    if dilation != 1 or phases == 1:
        return pack_operand_taps(w).reshape(-1)
    c_out, c_in, k = w.shape
    mats = torch.zeros(k - 1 + phases, phases * c_out, c_in, dtype=torch.float32, device=w.device)
    for pi in range(k - 1 + phases):
        for d in range(phases):
            tap = pi - d
            if 0 <= tap < k:
                mats[pi, d * c_out:(d + 1) * c_out] = w[:, :, tap].float()
    # (n, ci) -> chunk ci // 8, row n, lane ci % 8
    t = mats.reshape(k - 1 + phases, phases * c_out, c_in // 8, 8).permute(0, 2, 1, 3)
    return t.to(torch.bfloat16).contiguous().reshape(-1)


def pack_mrf_weights(sd: Dict[str, torch.Tensor], stage: int, device):
    """All 18 convs of one stage's three resblocks, in the order the fused kernel walks them:
    (resblock k=3,7,11) x (pair 0..2) x (convs1, convs2).  Returns (weights bf16 flat, bias fp32 [18][C])."""
    ws, bs = [], []
    phases = None
    for j in range(3):
        pre = f"vocoder.resblocks.{stage * 3 + j}."
        for q in range(3):
            for name in ("convs1", "convs2"):
                w = sd[pre + f"{name}.{q}.weight"].detach().to(device)
                if phases is None:
                    phases = mrf_phases(w.shape[0])
                dil = RESBLOCK_DILATIONS[q] if name == "convs1" else 1
                ws.append(pack_mrf_conv(w, dil, phases))
                bs.append(sd[pre + f"{name}.{q}.bias"].detach().to(device=device, dtype=torch.float32))
    return torch.cat(ws).contiguous(), torch.stack(bs).contiguous()


@dataclass
class PackedCFM:
    cond_table: torch.Tensor          # (vocab+1, 256) fp32 : E @ W_embed[:, 80:]^T + b_embed (hoisted, loop invariant)
    emb_table: torch.Tensor           # (vocab+1, 768) fp32 : raw embedding (API parity / duration predictor)
    w_embed: torch.Tensor             # bf16 [256][128]  (xt part of to_embed, K padded 80 -> 128)
    dw_w: torch.Tensor                # fp32 [31][256] (tap-major)
    dw_b: torch.Tensor
    four_w: torch.Tensor
    lin_w: torch.Tensor
    lin_b: torch.Tensor
    gamma_w: torch.Tensor             # fp32 [2*depth][256][256]
    inv_freq: torch.Tensor
    w_qkv: List[torch.Tensor] = field(default_factory=list)
    w_out: List[torch.Tensor] = field(default_factory=list)
    w_ff1: List[torch.Tensor] = field(default_factory=list)   # bf16 [1792][768], value/gate rows interleaved per 256 block
    b_ff1: List[torch.Tensor] = field(default_factory=list)
    w_ff2: List[torch.Tensor] = field(default_factory=list)   # bf16 [256][2688]
    b_ff2: List[torch.Tensor] = field(default_factory=list)
    final_norm_w: torch.Tensor = None
    w_pred: torch.Tensor = None       # bf16 [80][256]
    # duration-prediction variant only: per-unit dot products of the Conv1d(768 -> 1, k 3) taps, (3, vocab+1) fp32
    dur_table: torch.Tensor = None
    dur_bias: float = 0.0


@dataclass
class PackedVocoder:
    w_pre: torch.Tensor               # bf16 [512][7*128]
    b_pre: torch.Tensor
    w_up: List[torch.Tensor] = field(default_factory=list)
    b_up: List[torch.Tensor] = field(default_factory=list)
    # stage -> (packed weight, bias) of the up-samplers with L_out = stride L in row-group conv form
    # (upsampler_as_row_group_conv)
    up_pair: Dict[int, tuple] = field(default_factory=dict)
    # [stage][resblock j][pair q] -> (w1, b1, w2, b2); for q == 2 the conv2 lives in w_tail instead
    w_c1: List[List[List[torch.Tensor]]] = field(default_factory=list)
    b_c1: List[List[List[torch.Tensor]]] = field(default_factory=list)
    w_c2: List[List[List[torch.Tensor]]] = field(default_factory=list)
    b_c2: List[List[List[torch.Tensor]]] = field(default_factory=list)
    w_tail: List[torch.Tensor] = field(default_factory=list)  # bf16 [C][(3+7+11)*C]: last conv2 of the 3 resblocks
    b_tail: List[torch.Tensor] = field(default_factory=list)  # fp32 [C] = sum of their biases
    w_mrf: Dict[int, torch.Tensor] = field(default_factory=dict)   # stage -> fused-MRF operand-layout weights
    b_mrf: Dict[int, torch.Tensor] = field(default_factory=dict)   # stage -> fp32 [18][C]
    w_post: torch.Tensor = None       # fp32 [7][16]
    b_post: float = 0.0


def glu_row_permutation(inter: int = 896, device=None) -> torch.Tensor:
    """Row order of the packed conv1 weight: per 256-row tile, 128 value rows then the matching 128 gate rows
    (SIGLU: value = first half of the channels, gate = second half; fastspeech/modules.py:29)."""
    idx = []
    for t in range(inter // 128):
        idx.append(torch.arange(t * 128, (t + 1) * 128))
        idx.append(torch.arange(inter + t * 128, inter + (t + 1) * 128))
    return torch.cat(idx).to(device)


def pack_cfm(sd: Dict[str, torch.Tensor], device, depth: int = 4, dim_in: int = 80, split: bool = False) -> PackedCFM:
    """split=True packs the GEMM weights for the tight-precision library (split_operand)."""
    f = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    sp = (lambda w, dim=1: split_operand(w, dim)) if split else (lambda w, dim=1: w)
    w_emb = f("model.to_embed.weight")
    emb = f("model.to_cond_emb.weight").contiguous()
    hidden = w_emb.shape[0]
    inter = sd["model.transformer.layers.0.4.conv2.weight"].shape[1]
    # hoisted loop-invariant conditioning projection, in fp64 then rounded once to fp32
    cond_table = (emb.double() @ w_emb[:, dim_in:].double().t() + f("model.to_embed.bias").double()).float().contiguous()
    w_in = sp(w_emb[:, :dim_in])                       # (hidden, 80) or, split, (hidden, 240)
    k_pad = (w_in.shape[1] + 63) // 64 * 64
    w_x = torch.zeros(hidden, k_pad, dtype=torch.float32, device=device)
    w_x[:, :w_in.shape[1]] = w_in
    gam = []
    for i in range(depth):
        gam.append(f(f"model.transformer.layers.{i}.1.to_weight.weight"))
        gam.append(f(f"model.transformer.layers.{i}.3.to_weight.weight"))
    p = PackedCFM(
        cond_table=cond_table,
        emb_table=emb,
        w_embed=w_x.to(torch.bfloat16).contiguous(),
        dw_w=f("model.conv_embed.dw_conv1d.0.weight").reshape(hidden, -1).t().contiguous(),   # tap-major [31][256]
        dw_b=f("model.conv_embed.dw_conv1d.0.bias").contiguous(),
        four_w=f("model.time_cond_mlp.0.weights").contiguous(),
        lin_w=f("model.time_cond_mlp.1.weight").contiguous(),
        lin_b=f("model.time_cond_mlp.1.bias").contiguous(),
        gamma_w=torch.stack(gam).contiguous(),
        inv_freq=f("model.transformer.rotary_emb.inv_freq").contiguous(),
        final_norm_w=f("model.transformer.final_norm.weight").contiguous(),
        w_pred=sp(f("model.to_pred.weight")).to(torch.bfloat16).contiguous(),
    )
    if "model.duration_predictor.conv.weight" in sd:
        # conv over embedding rows = three table lookups (fastspeech/modules.py:86,103); fp64 product, rounded once
        wd = f("model.duration_predictor.conv.weight")[0].double()            # (768, 3)
        p.dur_table = (wd.t() @ emb.double().t()).float().contiguous()        # (3, vocab+1); column 0 = pad row = 0
        p.dur_bias = float(sd["model.duration_predictor.conv.bias"].reshape(-1)[0])
    perm = glu_row_permutation(inter, device)
    for i in range(depth):
        pre = f"model.transformer.layers.{i}."
        p.w_qkv.append(sp(f(pre + "2.to_qkv.weight")).to(torch.bfloat16).contiguous())
        p.w_out.append(sp(f(pre + "2.to_out.weight")).to(torch.bfloat16).contiguous())
        w1 = f(pre + "4.conv1.weight")[perm]
        p.w_ff1.append(pack_conv_weight(sp(w1), 64))
        p.b_ff1.append(f(pre + "4.conv1.bias")[perm].contiguous())
        p.w_ff2.append(pack_conv_weight(sp(f(pre + "4.conv2.weight")), 64))
        p.b_ff2.append(f(pre + "4.conv2.bias").contiguous())
    return p


def pack_vocoder(sd: Dict[str, torch.Tensor], device, split: bool = False) -> PackedVocoder:
    """split=True packs for the tight-precision library: split GEMM weights, no fused-MRF / row-group forms (that
    library runs every stage through the plain conv and polyphase kernels)."""
    f = lambda k: sd[k].detach().to(device=device, dtype=torch.float32)
    sp = (lambda w: split_operand(w, 1)) if split else (lambda w: w)
    v = PackedVocoder(w_pre=pack_conv_weight(sp(f("vocoder.conv_pre.weight")), 64), b_pre=f("vocoder.conv_pre.bias").contiguous())
    c = 512
    for i, (s, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        v.w_up.append(pack_upsampler_weight(f(f"vocoder.upsampler.{i}.weight"), s, split=split))
        v.b_up.append(f(f"vocoder.upsampler.{i}.bias").contiguous())
        if k - 2 * ((k - s) // 2) == s and not split:
            wv, bv = upsampler_as_row_group_conv(f(f"vocoder.upsampler.{i}.weight"), f(f"vocoder.upsampler.{i}.bias"), s)
            v.up_pair[i] = (pack_conv_weight(wv, block_k_for(wv.shape[1])), bv)
        c //= 2
        bk = block_k_for(c)
        w1s, b1s, w2s, b2s, tails, tail_b = [], [], [], [], [], None
        for j in range(len(RESBLOCK_KERNELS)):
            pre = f"vocoder.resblocks.{i * 3 + j}."
            w1s.append([pack_conv_weight(sp(f(pre + f"convs1.{q}.weight")), bk) for q in range(3)])
            b1s.append([f(pre + f"convs1.{q}.bias").contiguous() for q in range(3)])
            # (split: all three conv2's individually -- the tight engine does not use the fused tail)
            w2s.append([pack_conv_weight(sp(f(pre + f"convs2.{q}.weight")), bk) for q in range(3 if split else 2)])
            b2s.append([f(pre + f"convs2.{q}.bias").contiguous() for q in range(3 if split else 2)])
            tails.append(pack_conv_weight(f(pre + "convs2.2.weight"), bk))
            tb = f(pre + "convs2.2.bias")
            tail_b = tb if tail_b is None else tail_b + tb
        v.w_c1.append(w1s)
        v.b_c1.append(b1s)
        v.w_c2.append(w2s)
        v.b_c2.append(b2s)
        v.w_tail.append(torch.cat(tails, dim=1).contiguous())
        v.b_tail.append(tail_b.contiguous())
        if c in FUSED_MRF_CHANNELS and not split:
            v.w_mrf[i], v.b_mrf[i] = pack_mrf_weights(sd, i, device)
    v.w_post = f("vocoder.conv_post.weight")[0].t().contiguous()  # (7, 16)
    v.b_post = float(sd["vocoder.conv_post.bias"].detach().float().reshape(-1)[0])
    return v
