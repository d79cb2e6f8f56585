"""Per-kernel parity checks: each function runs one C-ABI op on cuda:0 and compares it with the CPU oracle
(oracle/cfm_hifigan_oracle.py) or, for the generic conv core, with a plain torch fp32 conv of the same bf16-rounded
operands.  Used by tests/test_gpu_kernels.py (pytest -m gpu) and tools/gpu_check.py (prints every result).

Tolerances (relative L2 unless noted):
  * integer / gather work: bit exact;
  * fp32 CUDA-core kernels: 1e-5;
  * bf16 tensor-core kernels (fp32 accumulate, one bf16 rounding of the output): 6e-3  (bf16 eps = 7.8e-3, the
    rounding error of a single store is <= 3.9e-3 relative per element).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Tuple

import torch
import torch.nn.functional as F

from oracle import cfm_hifigan_oracle as oracle
from speech_resynth_b200 import _native as nat
from speech_resynth_b200 import packing, synthetic
from speech_resynth_b200.engine import _i32

DEV = "cuda:0"
_KEEP: list = []


def P(t):
    """device pointer of `t`, keeping the tensor alive until the check ends: the launches are asynchronous and an
    expression like P(x.to(DEV)) would otherwise free (and let the allocator recycle) the buffer before the kernel
    has even been enqueued"""
    if t is None:
        return None
    _KEEP.append(t)
    return nat.ptr(t)
BF16_TOL = 6e-3
F32_TOL = 1e-5



def _guarded(fn):
    def run():
        _KEEP.clear()
        try:
            res = fn()
            torch.cuda.synchronize()
            return res
        finally:
            torch.cuda.synchronize()
            _KEEP.clear()

    run.__name__ = fn.__name__
    return run


class _Registry(dict):
    def __setitem__(self, key, fn):
        super().__setitem__(key, _guarded(fn))


CHECKS = _Registry()


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def bf(x: torch.Tensor) -> torch.Tensor:
    """round to bf16 and back (what the kernels see)"""
    return x.to(torch.bfloat16).float()


_SD = None


def sd():
    global _SD
    if _SD is None:
        _SD = synthetic.make_state_dict(0)
    return _SD


def g(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------------------ small ops
@check
def gather_bit_exact():
    table = sd()["model.to_cond_emb.weight"]
    ids = synthetic.make_units(4, 33, seed=21, lengths=[33, 20, 1, 7])
    out = torch.empty(4, 33, 768, device=DEV)
    nat.call("srb_embed_gather", P(table.to(DEV)), P(ids.to(DEV)), P(out), ids.numel(), 2001, 768)
    exact = torch.equal(out.cpu(), oracle.embed_gather(table, ids))
    return (0.0 if exact else 1.0), 0.0


@check
def unit_lengths_exact():
    ids = synthetic.make_units(5, 300, seed=3, lengths=[300, 1, 299, 128, 57])
    out = torch.empty(5, dtype=torch.int32, device=DEV)
    nat.call("srb_unit_lengths", P(ids.to(DEV)), P(out), 5, 300)
    return (0.0 if out.cpu().tolist() == [300, 1, 299, 128, 57] else 1.0), 0.0


@check
def time_cond_table():
    s = sd()
    times = oracle.ode_times(0.1)
    pk = packing.pack_cfm(s, DEV)
    nfe = len(times)
    temb = torch.empty(nfe, 256, device=DEV)
    gt = torch.empty(nfe, 8, 256, device=DEV)
    t_dev = times.to(DEV)
    nat.call("srb_time_cond_table", P(t_dev), nfe, P(pk.four_w), P(pk.lin_w), P(pk.lin_b), P(pk.gamma_w), 8, P(temb), P(gt))
    torch.cuda.synchronize()
    s64 = oracle.to_dtype(s, torch.float64)
    ref_t = torch.stack([oracle.time_embedding(s64, t.double()) for t in times])
    ref_g = []
    for t_row in ref_t:
        rows = []
        for i in range(4):
            for j in (1, 3):
                rows.append(16.0 * (F.linear(t_row, s64[f"model.transformer.layers.{i}.{j}.to_weight.weight"]) + 1.0))
        ref_g.append(torch.stack(rows))
    e1 = rel_l2(temb, ref_t)
    e2 = rel_l2(gt, torch.stack(ref_g))
    return max(e1, e2), 2e-5


@check
def rotary_table():
    inv = sd()["model.transformer.rotary_emb.inv_freq"]
    cs = torch.empty(1500, 64, device=DEV)
    sn = torch.empty(1500, 64, device=DEV)
    nat.call("srb_rotary_table", P(inv.to(DEV)), 1500, P(cs), P(sn))
    ang = torch.arange(1500).float()[:, None] * inv[None, :]   # fp32 product, as the reference
    err = max(float((cs.cpu() - ang.double().cos()).abs().max()), float((sn.cpu() - ang.double().sin()).abs().max()))
    return err, 2e-6   # absolute


@check
def prior_prepare():
    x = torch.randn(2, 37, 80, generator=g(1))
    xt = x.clone().to(DEV)
    xb = torch.empty(2, 37, 80, dtype=torch.bfloat16, device=DEV)
    nat.call("srb_prior_prepare", P(xt), P(xb), x.numel(), 1.0, 1)
    ref = x.clamp(-1, 1)
    ok = torch.equal(xt.cpu(), ref) and torch.equal(xb.cpu(), ref.to(torch.bfloat16))
    return (0.0 if ok else 1.0), 0.0


# ------------------------------------------------------------------------------------------------ conv-GEMM core
def _conv_case(c_in, c_out, k, dil, batch=2, rows=300, with_res=True, seed=0):
    gen = g(seed)
    x = bf(torch.randn(batch, rows, c_in, generator=gen))
    w = bf(torch.randn(c_out, c_in, k, generator=gen) / math.sqrt(c_in * k))
    bias = torch.randn(c_out, generator=gen)
    res = bf(torch.randn(batch, rows, c_out, generator=gen)) if with_res else None
    y = F.conv1d(x.transpose(1, 2).double(), w.double(), bias.double(), dilation=dil, padding=(k - 1) // 2 * dil).transpose(1, 2)
    if with_res:
        y = y + res.double()
    y = y * 0.5
    wp = packing.pack_conv_weight(w.to(DEV), packing.block_k_for(c_in))
    out_raw = torch.full((batch, rows, c_out), float("nan"), dtype=torch.bfloat16, device=DEV)
    out_act = torch.full_like(out_raw, float("nan"))
    xd = x.to(DEV).to(torch.bfloat16).contiguous()
    rd = res.to(DEV).to(torch.bfloat16).contiguous() if with_res else None
    nat.call("srb_hifigan_conv", P(xd), None, None, 1, _i32([k]), _i32([dil]), P(wp), P(bias.to(DEV)), P(rd), None, None,
             P(out_raw), P(out_act), batch, rows, c_in, c_out, 0.5, 0.1)
    torch.cuda.synchronize()
    e1 = rel_l2(out_raw.float(), y)
    e2 = rel_l2(out_act.float(), F.leaky_relu(y, 0.1))
    return max(e1, e2), BF16_TOL


def _conv_res_act_case(c, k, rows=300, batch=2, seed=3, n_src=1):
    """srb_hifigan_conv_res_act: residuals handed over as bf16(leaky_relu(r, 0.1)); the launch must add the RAW r (recovered
    in the epilogue) and, with only the activated output requested, equal the two-copy form to one bf16 rounding -- for one
    source (TMA epilogue, in place: the output overwrites the residual tensor, as the resblock chains do) and for the
    three-source fused tail (register epilogue, three residuals)."""
    gen = g(seed)
    ks = [k] if n_src == 1 else [3, 7, 11]
    xs = [bf(torch.randn(batch, rows, c, generator=gen)) for _ in ks]
    ws = [bf(torch.randn(c, c, kk, generator=gen) / math.sqrt(c * kk * n_src)) for kk in ks]
    bias = torch.randn(c, generator=gen)
    raws = [torch.randn(batch, rows, c, generator=gen) * 2.0 for _ in ks]
    acts = [bf(F.leaky_relu(r, 0.1)) for r in raws]                       # what the chain stores
    recovered = [torch.where(a < 0, a * 10.0, a) for a in acts]           # what the epilogue must add
    y = bias.double()[None, None, :].expand(batch, rows, c).clone()
    for x, w, kk, r in zip(xs, ws, ks, recovered):
        y = y + F.conv1d(x.transpose(1, 2).double(), w.double(), None, padding=(kk - 1) // 2).transpose(1, 2) + r.double()
    y = y * (1.0 / n_src)
    wp = torch.cat([packing.pack_conv_weight(w.to(DEV), packing.block_k_for(c)) for w in ws], dim=1).contiguous()
    xd = [x.to(DEV).to(torch.bfloat16).contiguous() for x in xs]
    rd = [a.to(DEV).to(torch.bfloat16).contiguous() for a in acts]
    out = rd[0] if n_src == 1 else torch.full((batch, rows, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    nat.call("srb_hifigan_conv_res_act", P(xd[0]), P(xd[1]) if n_src == 3 else None, P(xd[2]) if n_src == 3 else None, n_src,
             _i32(ks), _i32([1] * n_src), P(wp), P(bias.to(DEV)), P(rd[0]), P(rd[1]) if n_src == 3 else None,
             P(rd[2]) if n_src == 3 else None, 0.1, None, P(out), batch, rows, c, c, 1.0 / n_src, 0.1)
    torch.cuda.synchronize()
    # the recovered residual itself is within one bf16 rounding of the raw one
    assert max(rel_l2(rc, r) for rc, r in zip(recovered, raws)) < 3e-3
    return rel_l2(out.float(), F.leaky_relu(y, 0.1)), BF16_TOL


for _name, _args in (("conv_res_act_c64_k3", dict(c=64, k=3)), ("conv_res_act_c128_k7", dict(c=128, k=7)),
                     ("conv_res_act_c256_k11", dict(c=256, k=11, rows=400)), ("conv_res_act_tail_c64", dict(c=64, k=0, n_src=3))):
    def _fn(_args=_args):
        return _conv_res_act_case(**_args)
    _fn.__name__ = _name
    CHECKS[_name] = _fn


def _pair_fused_case(k, dil, rows=300, batch=2, seed=5):
    """srb_hifigan_pair_fused (C = 64): conv1_{k,dil} -> leaky_relu -> conv2_{k,1} -> + residual -> leaky_relu in one launch,
    input handed over activated (single-copy form).  Against a float64 reference that rounds t to bf16 where the kernel
    does, and against the two-launch form (srb_hifigan_conv + srb_hifigan_conv_res_act) on the same operands."""
    c = 64
    gen = g(seed)
    raw = torch.randn(batch, rows, c, generator=gen) * 1.5
    xa = bf(F.leaky_relu(raw, 0.1))
    rec = torch.where(xa < 0, xa * 10.0, xa)
    w1 = bf(torch.randn(c, c, k, generator=gen) / math.sqrt(c * k))
    w2 = bf(torch.randn(c, c, k, generator=gen) / math.sqrt(c * k))
    b1, b2 = torch.randn(c, generator=gen) * 0.1, torch.randn(c, generator=gen) * 0.1
    t = F.conv1d(xa.transpose(1, 2).double(), w1.double(), b1.double(), dilation=dil, padding=(k - 1) // 2 * dil)
    t = bf(F.leaky_relu(t, 0.1).float()).double()
    y = F.conv1d(t, w2.double(), b2.double(), padding=(k - 1) // 2).transpose(1, 2) + rec.double()
    ref = F.leaky_relu(y, 0.1)
    wp1 = packing.pack_conv_weight(w1.to(DEV), 64)
    wp2 = packing.pack_conv_weight(w2.to(DEV), 64)
    xd = xa.to(DEV).to(torch.bfloat16).contiguous()
    out = torch.full((batch, rows, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    nat.call("srb_hifigan_pair_fused", P(xd), P(wp1), P(b1.to(DEV)), P(wp2), P(b2.to(DEV)), P(out), batch, rows, c, k, dil, 0.1)
    # the two-launch form
    td = torch.empty_like(out)
    out2 = torch.empty_like(out)
    nat.call("srb_hifigan_conv", P(xd), None, None, 1, _i32([k]), _i32([dil]), P(wp1), P(b1.to(DEV)), None, None, None, None,
             P(td), batch, rows, c, c, 1.0, 0.1)
    nat.call("srb_hifigan_conv_res_act", P(td), None, None, 1, _i32([k]), _i32([1]), P(wp2), P(b2.to(DEV)), P(xd), None, None, 0.1,
             None, P(out2), batch, rows, c, c, 1.0, 0.1)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.float()).all()), "pair_fused left rows unwritten"
    e_two = rel_l2(out.float(), out2.float().double())
    assert e_two < 2e-3, f"pair_fused differs from the two-launch form: {e_two}"
    return rel_l2(out.float(), ref), BF16_TOL


for _k, _d, _rows, _b in ((3, 1, 300, 2), (3, 3, 126, 1), (3, 5, 1000, 3), (7, 1, 300, 2), (7, 3, 122, 1), (7, 5, 1000, 3), (3, 1, 5, 2)):
    def _fn(_k=_k, _d=_d, _rows=_rows, _b=_b):
        return _pair_fused_case(_k, _d, rows=_rows, batch=_b)
    _fn.__name__ = f"pair_fused_c64_k{_k}_d{_d}_r{_rows}"
    CHECKS[_fn.__name__] = _fn


def _make_conv_check(name, *args, **kw):
    def fn():
        return _conv_case(*args, **kw)

    fn.__name__ = name
    CHECKS[name] = fn


# every (block_n, block_k) instantiation the vocoder uses, at the real kernel sizes / dilations
_make_conv_check("conv_c256_k3_d1", 256, 256, 3, 1)
_make_conv_check("conv_c256_k11_d5", 256, 256, 11, 5, rows=400)
_make_conv_check("conv_c128_k7_d3", 128, 128, 7, 3)
_make_conv_check("conv_c64_k11_d1", 64, 64, 11, 1)
_make_conv_check("conv_c32_k7_d5", 32, 32, 7, 5)
_make_conv_check("conv_c16_k3_d3", 16, 16, 3, 3)
_make_conv_check("conv_c16_k11_d5", 16, 16, 11, 5, rows=1000, batch=3)
_make_conv_check("conv_pre_80_512_k7", 80, 512, 7, 1, with_res=False, rows=130)
_make_conv_check("conv_rows_lt_tile", 64, 64, 3, 1, rows=17, batch=1)
# weight-stationary form of the C = 64 convs: more row tiles than resident CTAs (several tiles per CTA), halo boxes
_make_conv_check("conv_c64_k3_d1_many_tiles", 64, 64, 3, 1, rows=5000, batch=9)
_make_conv_check("conv_c64_k7_d5_many_tiles", 64, 64, 7, 5, rows=4099, batch=10, seed=3)
_make_conv_check("conv_c64_k11_d3_many_tiles", 64, 64, 11, 3, rows=3001, batch=7, seed=4)


@check
def conv_mrf_tail_fused():
    """three sources (k = 3, 7, 11) accumulated in one launch + three residuals + mean + leaky_relu"""
    gen = g(5)
    c, rows, batch = 64, 333, 2
    xs = [bf(torch.randn(batch, rows, c, generator=gen)) for _ in range(3)]
    ws = [bf(torch.randn(c, c, k, generator=gen) / math.sqrt(c * k)) for k in (3, 7, 11)]
    bs = [torch.randn(c, generator=gen) for _ in range(3)]
    rs = [bf(torch.randn(batch, rows, c, generator=gen)) for _ in range(3)]
    y = 0
    for x, w, b_, r, k in zip(xs, ws, bs, rs, (3, 7, 11)):
        y = y + F.conv1d(x.transpose(1, 2).double(), w.double(), b_.double(), padding=k // 2).transpose(1, 2) + r.double()
    y = y / 3
    wp = torch.cat([packing.pack_conv_weight(w.to(DEV), 64) for w in ws], dim=1).contiguous()
    out = torch.empty(batch, rows, c, dtype=torch.bfloat16, device=DEV)
    d16 = lambda t: t.to(DEV).to(torch.bfloat16).contiguous()
    xd, rd = [d16(x) for x in xs], [d16(r) for r in rs]
    nat.call("srb_hifigan_conv", P(xd[0]), P(xd[1]), P(xd[2]), 3, _i32([3, 7, 11]), _i32([1, 1, 1]), P(wp),
             P(sum(bs).to(DEV)), P(rd[0]), P(rd[1]), P(rd[2]), None, P(out), batch, rows, c, c, 1.0 / 3.0, 0.01)
    torch.cuda.synchronize()
    return rel_l2(out.float(), F.leaky_relu(y, 0.01)), BF16_TOL


def _upsample_case(c_in, k, s, rows_in=77, batch=2, seed=0):
    gen = g(seed)
    c_out = c_in // 2
    x = bf(torch.randn(batch, rows_in, c_in, generator=gen))
    w = bf(torch.randn(c_in, c_out, k, generator=gen) / math.sqrt(c_in * k / s))
    bias = torch.randn(c_out, generator=gen)
    pad = (k - s) // 2
    with torch.backends.mkldnn.flags(enabled=False):
        y = F.conv_transpose1d(x.transpose(1, 2).double(), w.double(), bias.double(), stride=s, padding=pad).transpose(1, 2)
    rows_out = y.shape[1]
    wp = packing.pack_upsampler_weight(w.to(DEV), s)
    out_raw = torch.full((batch, rows_out, c_out), float("nan"), dtype=torch.bfloat16, device=DEV)
    out_act = torch.full_like(out_raw, float("nan"))
    xd = x.to(DEV).to(torch.bfloat16).contiguous()
    nat.call("srb_hifigan_upsample", P(xd), P(wp), P(bias.to(DEV)), P(out_raw), P(out_act), batch, rows_in, c_in, c_out, k, s, 0.1)
    torch.cuda.synchronize()
    e1 = rel_l2(out_raw.float(), y)
    e2 = rel_l2(out_act.float(), F.leaky_relu(y, 0.1))
    return max(e1, e2), BF16_TOL


for _i, (_c, _k, _s) in enumerate(zip((512, 256, 128, 64, 32), packing.UPSAMPLE_KERNELS, packing.UPSAMPLE_RATES)):
    def _fn(c=_c, k=_k, s=_s, i=_i):
        return _upsample_case(c, k, s, rows_in=77 + 60 * i)
    CHECKS[f"upsample_stage{_i}_c{_c}_k{_k}_s{_s}"] = _fn


def _upsample_row_group_case(c_in, k, s, rows_in, batch=2, seed=0):
    """the production form of the up-samplers with L_out = stride L: srb_hifigan_conv on the weights of
    packing.upsampler_as_row_group_conv, (B, L, s C_out) output read as (B, s L, C_out), vs conv_transpose1d"""
    gen = g(seed)
    c_out = c_in // 2
    x = bf(torch.randn(batch, rows_in, c_in, generator=gen))
    w = bf(torch.randn(c_in, c_out, k, generator=gen) / math.sqrt(c_in * k / s))
    bias = torch.randn(c_out, generator=gen)
    pad = (k - s) // 2
    with torch.backends.mkldnn.flags(enabled=False):
        y = F.conv_transpose1d(x.transpose(1, 2).double(), w.double(), bias.double(), stride=s, padding=pad).transpose(1, 2)
    assert y.shape[1] == s * rows_in
    wv, bv = packing.upsampler_as_row_group_conv(w.to(DEV), bias.to(DEV), s)
    wp = packing.pack_conv_weight(wv, packing.block_k_for(c_in))
    out_raw = torch.full((batch, s * rows_in, c_out), float("nan"), dtype=torch.bfloat16, device=DEV)
    out_act = torch.full_like(out_raw, float("nan"))
    xd = x.to(DEV).to(torch.bfloat16).contiguous()
    nat.call("srb_hifigan_conv", P(xd), None, None, 1, _i32([3]), _i32([1]), P(wp), P(bv), None, None, None, P(out_raw),
             P(out_act), batch, rows_in, c_in, s * c_out, 1.0, 0.1)
    torch.cuda.synchronize()
    e1 = rel_l2(out_raw.float(), y)
    e2 = rel_l2(out_act.float(), F.leaky_relu(y, 0.1))
    return max(e1, e2), BF16_TOL


for _i, (_c, _k, _s) in enumerate(zip((512, 256, 128, 64, 32), packing.UPSAMPLE_KERNELS, packing.UPSAMPLE_RATES)):
    if _k - 2 * ((_k - _s) // 2) == _s:
        def _fn(c=_c, k=_k, s=_s, i=_i):
            return _upsample_row_group_case(c, k, s, rows_in=77 + 60 * i)
        CHECKS[f"upsample_row_group_stage{_i}_c{_c}_k{_k}_s{_s}"] = _fn


def _mrf_case(c, rows, batch=2, seed=0):
    """whole fused MRF stage vs a float64 evaluation of HF:1359-1367,1475-1480 on the same bf16-rounded weights"""
    gen = g(seed)
    u = bf(torch.randn(batch, rows, c, generator=gen))
    ws, bs, ref_sum = [], [], 0
    for j, k in enumerate((3, 7, 11)):
        x = u.double().transpose(1, 2)
        for q, dil in enumerate((1, 3, 5)):
            w1 = bf(torch.randn(c, c, k, generator=gen) / math.sqrt(c * k))
            b1 = torch.randn(c, generator=gen) * 0.1
            w2 = bf(torch.randn(c, c, k, generator=gen) / math.sqrt(c * k))
            b2 = torch.randn(c, generator=gen) * 0.1
            ph = packing.mrf_phases(c)
            ws += [packing.pack_mrf_conv(w1, dil, ph), packing.pack_mrf_conv(w2, 1, ph)]
            bs += [b1, b2]
            r = x
            t = F.conv1d(F.leaky_relu(x, 0.1), w1.double(), b1.double(), dilation=dil, padding=(k - 1) // 2 * dil)
            x = F.conv1d(F.leaky_relu(t, 0.1), w2.double(), b2.double(), padding=(k - 1) // 2) + r
        ref_sum = ref_sum + x
    ref = F.leaky_relu(ref_sum / 3, 0.01).transpose(1, 2)
    out = torch.full((batch, rows, c), float("nan"), dtype=torch.bfloat16, device=DEV)
    nat.call("srb_hifigan_mrf_fused", P(u.to(DEV).to(torch.bfloat16).contiguous()), P(torch.cat(ws).to(DEV).contiguous()),
             P(torch.stack(bs).to(DEV).contiguous()), P(out), batch, rows, c, 0.1, 0.01)
    torch.cuda.synchronize()
    return rel_l2(out.float(), ref), 1.2e-2   # two bf16 operand roundings per pair, six pairs deep, fp32 residual stream


@check
def mrf_fused_c16():
    return _mrf_case(16, 3000)


@check
def mrf_fused_c32():
    return _mrf_case(32, 1500, seed=1)


@check
def mrf_fused_c16_short():
    return _mrf_case(16, 41, batch=3, seed=2)


@check
def post_tanh():
    gen = g(9)
    x = bf(torch.randn(2, 1000, 16, generator=gen))
    w = torch.randn(1, 16, 7, generator=gen) / 10
    b = 0.05
    y = torch.tanh(F.conv1d(x.transpose(1, 2).double(), w.double(), torch.tensor([b]).double(), padding=3)).squeeze(1)
    wav = torch.empty(2, 1000, device=DEV)
    nat.call("srb_hifigan_post", P(x.to(DEV).to(torch.bfloat16).contiguous()), P(w[0].t().contiguous().to(DEV)), b, P(wav), 2, 1000, None)
    return rel_l2(wav, y), F32_TOL


# ------------------------------------------------------------------------------------------------ transformer ops
def _cfm_inputs(batch=2, frames=150, lengths=(150, 97)):
    ids = synthetic.make_units(batch, frames, seed=7, lengths=list(lengths))
    mask = ids.ne(0)
    L = torch.tensor(lengths, dtype=torch.int32, device=DEV)
    return ids, mask, L


@check
def cfm_embed():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs()
    b, n = ids.shape
    xt = bf(torch.randn(b, n, 80, generator=g(2)))
    cond = torch.empty(b * n, 256, device=DEV)
    nat.call("srb_embed_gather", P(pk.cond_table), P(ids.to(DEV)), P(cond), b * n, 2001, 256)
    x0 = torch.empty(b * n, 256, device=DEV)
    nat.call("srb_cfm_embed", P(xt.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_embed), P(cond), P(x0), b, n)
    s64 = oracle.to_dtype(s, torch.float64)
    w = s64["model.to_embed.weight"].clone()
    w[:, :80] = bf(s["model.to_embed.weight"][:, :80]).double()  # the xt part of the weight is bf16 in the kernel
    hs = oracle.embed_gather(s64["model.to_cond_emb.weight"], ids)
    ref = F.linear(torch.cat([xt.double(), hs], dim=-1), w, s64["model.to_embed.bias"])
    return rel_l2(x0.view(b, n, 256), ref), 1e-5


@check
def cfm_posconv_norm():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs()
    b, n = ids.shape
    x0 = torch.randn(b, n, 256, generator=g(3))
    gvec = 16.0 * (1.0 + 0.1 * torch.randn(256, generator=g(4)))
    x = torch.empty(b, n, 256, device=DEV)
    xn = torch.empty(b, n, 256, dtype=torch.bfloat16, device=DEV)
    nat.call("srb_cfm_posconv_norm", P(x0.to(DEV)), P(pk.dw_w), P(pk.dw_b), P(gvec.to(DEV)), P(L), P(x), P(xn), b, n)
    s64 = oracle.to_dtype(s, torch.float64)
    ref_x = oracle.conv_pos_embed(s64, x0.double(), mask) + x0.double()
    ref_n = ref_x / ref_x.norm(dim=-1, keepdim=True).clamp_min(1e-12) * gvec.double()
    ref_n = ref_n * mask[..., None]
    e1 = rel_l2(x.cpu()[mask], ref_x[mask])
    e2 = rel_l2(xn.float(), ref_n)
    return max(e1, e2 / 400), 1e-5   # xn carries one bf16 rounding (4e-3): scaled so both share a bound


@check
def cfm_qkv_rope():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    b, n = 2, 150
    xn = bf(torch.randn(b, n, 256, generator=g(5)))
    sampler_cs = torch.empty(1024, 64, device=DEV)
    sampler_sn = torch.empty(1024, 64, device=DEV)
    nat.call("srb_rotary_table", P(pk.inv_freq), 1024, P(sampler_cs), P(sampler_sn))
    qkv = torch.empty(b, n, 768, dtype=torch.bfloat16, device=DEV)
    nat.call("srb_cfm_qkv_rope", P(xn.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_qkv[1]), P(sampler_cs), P(sampler_sn), P(qkv), None, None, b, n)
    w = bf(s["model.transformer.layers.1.2.to_qkv.weight"]).double()
    r = F.linear(xn.double(), w)
    q, k, v = r.chunk(3, dim=-1)
    rot = oracle.rotary_table(s["model.transformer.rotary_emb.inv_freq"], n).double()
    hd = lambda z: z.reshape(b, n, 2, 128).permute(0, 2, 1, 3)
    un = lambda z: z.permute(0, 2, 1, 3).reshape(b, n, 256)
    q, k = un(oracle.apply_rotary(rot, hd(q))), un(oracle.apply_rotary(rot, hd(k)))
    return rel_l2(qkv.float(), torch.cat([q, k, v], dim=-1)), BF16_TOL


def _norm_bounds(qk: torch.Tensor) -> torch.Tensor:
    """(B, N, 512) -> (B, 2 [q|k], 2 [head], 2 [f]): max over rows of the squared norm over columns i, i + 64 of the
    head with i in [32 f, 32 f + 32)"""
    b, n, _ = qk.shape
    z = qk.reshape(b, n, 2, 2, 2, 2, 32).double().pow(2)      # (b, n, q|k, head, lo|hi, f, 32)
    return z.sum(dim=(4, 6)).amax(dim=1).float().contiguous()


def _attention_tc_case(bounds: str, scale: float = 1.0):
    """tcgen05 attention (q|k buffer + transposed v) against the float64 softmax-attention definition.
    bounds: "none" -> two-pass path; "given" -> exact norm bounds (single pass when they allow it)"""
    worst = 0.0
    for b, n, lengths in ((3, 200, (200, 131, 64)), (2, 504, (500, 1)), (1, 136, (129,)), (2, 1024, (1024, 700)), (1, 128, (77,)),
                          (2, 1160, (1160, 385)), (1, 2200, (2100,))):
        L = torch.tensor(lengths, dtype=torch.int32, device=DEV)
        qkv = torch.randn(b, n, 768, generator=g(6 + n))
        qkv[..., :512] *= scale
        qkv = bf(qkv)
        m_pad = (b * n + 255) // 256 * 256
        qk = qkv[..., :512].contiguous().to(DEV).to(torch.bfloat16)
        vt = torch.zeros(256, m_pad, dtype=torch.bfloat16, device=DEV)
        vt[:, : b * n] = qkv[..., 512:].reshape(b * n, 256).t().to(DEV).to(torch.bfloat16)
        o = torch.full((b, n, 256), float("nan"), dtype=torch.bfloat16, device=DEV)
        nb = None
        if bounds == "given":
            # (B, 2 [q|k], 2 [head], 2 [frequency half]) maxima of the partial squared row norms
            nb = _norm_bounds(qkv[..., :512]).to(DEV)
        nat.call("srb_cfm_attention_tc", P(qk), 512, P(vt), m_pad, P(L), P(nb), P(o), b, n)
        torch.cuda.synchronize()
        q, k, v = (z.reshape(b, n, 2, 128).permute(0, 2, 1, 3).double() for z in qkv.chunk(3, dim=-1))
        mask = torch.arange(n)[None, :] < torch.tensor(lengths)[:, None]
        sc = torch.einsum("bhid,bhjd->bhij", q, k) / math.sqrt(128)
        sc = sc.masked_fill(~mask[:, None, None, :], float("-inf"))
        ref = torch.einsum("bhij,bhjd->bhid", sc.softmax(-1), v).permute(0, 2, 1, 3).reshape(b, n, 256)
        worst = max(worst, rel_l2(o.float(), ref))
    return worst, 1e-2


@check
def cfm_qkv_rope_records_norm_bounds():
    """fused projection: q | k rotated, v plain, norm bounds recorded / the other buffer cleared"""
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    b, n = 2, 150
    xn_host = bf(torch.randn(b, n, 256, generator=g(5)))
    xn = xn_host.to(DEV).to(torch.bfloat16).contiguous()
    cs = torch.empty(1024, 64, device=DEV)
    sn = torch.empty(1024, 64, device=DEV)
    nat.call("srb_rotary_table", P(pk.inv_freq), 1024, P(cs), P(sn))
    qkv = torch.empty(b, n, 768, dtype=torch.bfloat16, device=DEV)
    nb = torch.zeros(b, 2, 2, 2, device=DEV)
    nb_other = torch.full((b, 2, 2, 2), 7.0, device=DEV)
    nat.call("srb_cfm_qkv_rope", P(xn), P(pk.w_qkv[1]), P(cs), P(sn), P(qkv), P(nb), P(nb_other), b, n)
    w = bf(s["model.transformer.layers.1.2.to_qkv.weight"]).double()
    q, k, v = F.linear(xn_host.double(), w).chunk(3, dim=-1)
    rot = oracle.rotary_table(s["model.transformer.rotary_emb.inv_freq"], n).double()
    hd = lambda z: z.reshape(b, n, 2, 128).permute(0, 2, 1, 3)
    un = lambda z: z.permute(0, 2, 1, 3).reshape(b, n, 256)
    q, k = un(oracle.apply_rotary(rot, hd(q))), un(oracle.apply_rotary(rot, hd(k)))
    e1 = rel_l2(qkv.float(), torch.cat([q, k, v], dim=-1))
    e3 = rel_l2(nb, _norm_bounds(torch.cat([q, k], dim=-1)))
    cleared = bool((nb_other == 0).all())
    return (max(e1, e3 * 100) if cleared else 1.0), BF16_TOL


@check
def cfm_attention_tc():
    """no bounds supplied: two-pass path"""
    return _attention_tc_case("none")


@check
def cfm_attention_tc_single_pass():
    """unit-variance q, k: |q||k| log2(e)/sqrt(128) ~ 20 -> single pass, shift 0"""
    return _attention_tc_case("given")


@check
def cfm_attention_tc_bound_fallback():
    """q, k scaled x3: bound ~ 190 > 100 -> the kernel must choose the two-pass path by itself (peaky softmax)"""
    return _attention_tc_case("given", scale=3.0)


def _qk_rope_vt_case(b, n, m_pad):
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    xn_host = bf(torch.randn(b, n, 256, generator=g(5)))
    xn = torch.zeros(m_pad, 256, dtype=torch.bfloat16, device=DEV)
    xn[: b * n] = xn_host.reshape(b * n, 256).to(DEV).to(torch.bfloat16)
    cs = torch.empty(max(1024, n), 64, device=DEV)
    sn = torch.empty(max(1024, n), 64, device=DEV)
    nat.call("srb_rotary_table", P(pk.inv_freq), max(1024, n), P(cs), P(sn))
    qk = torch.empty(b, n, 512, dtype=torch.bfloat16, device=DEV)
    vt = torch.full((256, m_pad), float("nan"), dtype=torch.bfloat16, device=DEV)
    wq = pk.w_qkv[1]
    nb = torch.zeros(b, 2, 2, 2, device=DEV)
    nb_other = torch.full((b, 2, 2, 2), 7.0, device=DEV)
    nat.call("srb_cfm_qk_rope", P(xn), P(wq), P(cs), P(sn), P(qk), P(nb), P(nb_other), b, n)
    nat.call("srb_cfm_v_transposed", P(xn), P(wq[512:]), P(vt), m_pad)
    w = bf(s["model.transformer.layers.1.2.to_qkv.weight"]).double()
    r = F.linear(xn_host.double(), w)
    q, k, v = r.chunk(3, dim=-1)
    rot = oracle.rotary_table(s["model.transformer.rotary_emb.inv_freq"], n).double()
    hd = lambda z: z.reshape(b, n, 2, 128).permute(0, 2, 1, 3)
    un = lambda z: z.permute(0, 2, 1, 3).reshape(b, n, 256)
    q, k = un(oracle.apply_rotary(rot, hd(q))), un(oracle.apply_rotary(rot, hd(k)))
    e1 = rel_l2(qk.float(), torch.cat([q, k], dim=-1))
    e2 = rel_l2(vt[:, : b * n].float(), v.reshape(b * n, 256).t())
    tail_zero = bool((vt[:, b * n:].float() == 0).all())
    # recorded bounds: max squared row norm per (utterance, q|k, head), of the fp32 values before the bf16 rounding
    nb_ref = _norm_bounds(torch.cat([q, k], dim=-1))
    e3 = rel_l2(nb, nb_ref)
    cleared = bool((nb_other == 0).all())
    return (max(e1, e2, e3 * 100) if tail_zero and cleared else 1.0), BF16_TOL


@check
def cfm_qk_rope_vt_fused():
    """the whole to_qkv GEMM in one launch: q | k rotated into (B, N, 512), v transposed by the epilogue; bit-equal to the
    two-launch form (same MMAs, same roundings) on ragged frame counts, untouched beyond batch * frames"""
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    worst = 0.0
    for b, n in ((2, 152), (3, 504), (1, 40)):
        m_pad = (b * n + 255) // 256 * 256 + 256
        xn = torch.zeros(m_pad, 256, dtype=torch.bfloat16, device=DEV)
        xn[: b * n] = bf(torch.randn(b * n, 256, generator=g(5 + n))).to(DEV).to(torch.bfloat16)
        cs = torch.empty(1024, 64, device=DEV)
        sn = torch.empty(1024, 64, device=DEV)
        nat.call("srb_rotary_table", P(pk.inv_freq), 1024, P(cs), P(sn))
        wq = pk.w_qkv[2]
        qk1 = torch.empty(b, n, 512, dtype=torch.bfloat16, device=DEV)
        vt1 = torch.full((256, m_pad), 5.0, dtype=torch.bfloat16, device=DEV)
        nb1, nc1 = torch.zeros(b, 2, 2, 2, device=DEV), torch.full((b, 2, 2, 2), 7.0, device=DEV)
        nat.call("srb_cfm_qk_rope", P(xn), P(wq), P(cs), P(sn), P(qk1), P(nb1), P(nc1), b, n)
        nat.call("srb_cfm_v_transposed", P(xn), P(wq[512:]), P(vt1), m_pad)
        qk2 = torch.empty(b, n, 512, dtype=torch.bfloat16, device=DEV)
        vt2 = torch.full((256, m_pad), 5.0, dtype=torch.bfloat16, device=DEV)
        nb2, nc2 = torch.zeros(b, 2, 2, 2, device=DEV), torch.full((b, 2, 2, 2), 7.0, device=DEV)
        nat.call("srb_cfm_qk_rope_vt", P(xn), P(wq), P(cs), P(sn), P(qk2), P(vt2), m_pad, P(nb2), P(nc2), b, n)
        torch.cuda.synchronize()
        ok = torch.equal(qk1, qk2) and torch.equal(vt1[:, : b * n], vt2[:, : b * n]) and torch.equal(nb1, nb2)
        ok = ok and bool((nc2 == 0).all()) and bool((vt2[:, b * n:] == 5.0).all())
        worst = max(worst, 0.0 if ok else 1.0)
    return worst, 0.0


@check
def cfm_qk_rope_and_v_transposed():
    return _qk_rope_vt_case(2, 150, 512)


@check
def cfm_qk_rope_long_positions():
    """60 s utterance (config 5): rotary angles up to 3000 rad through the angle-addition tables"""
    return _qk_rope_vt_case(1, 3000, 3072)


@check
def cfm_attn_out_norm():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs()
    b, n = ids.shape
    o = bf(torch.randn(b, n, 256, generator=g(7)))
    x = torch.randn(b, n, 256, generator=g(8))
    gvec = 16.0 * (1.0 + 0.1 * torch.randn(256, generator=g(4)))
    xd = x.clone().to(DEV)
    xn = torch.empty(b, n, 256, dtype=torch.bfloat16, device=DEV)
    nat.call("srb_cfm_attn_out_norm", P(o.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_out[2]), P(gvec.to(DEV)), P(L), P(xd), P(xn), b, n)
    w = bf(s["model.transformer.layers.2.2.to_out.weight"]).double()
    ref_x = F.linear(o.double(), w) + x.double()
    ref_n = ref_x / ref_x.norm(dim=-1, keepdim=True).clamp_min(1e-12) * gvec.double() * mask[..., None]
    e1 = rel_l2(xd, ref_x)
    e2 = rel_l2(xn.float(), ref_n)
    return max(e1 * 400, e2), BF16_TOL   # x is fp32 (bound 1.5e-5), xn carries the bf16 rounding


def _ffn_glu_case(lengths, frames, pad_separated):
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs(batch=len(lengths), frames=frames, lengths=lengths)
    b, n = ids.shape
    xn = bf(torch.randn(b, n, 256, generator=g(9))) * mask[..., None]
    h = torch.empty(b, n, 896, dtype=torch.bfloat16, device=DEV)
    nat.call("srb_cfm_ffn_glu", P(xn.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_ff1[0]), P(pk.b_ff1[0]), P(L), P(h), b, n,
             pad_separated)
    w = bf(s["model.transformer.layers.0.4.conv1.weight"]).double()
    y = F.conv1d(xn.double().transpose(1, 2), w, s["model.transformer.layers.0.4.conv1.bias"].double(), padding=1)
    val, gate = y.chunk(2, dim=1)
    ref = (F.silu(gate) * val).transpose(1, 2) * mask[..., None]
    return h, ref


@check
def cfm_ffn_glu():
    h, ref = _ffn_glu_case((150, 97), 150, 0)
    return rel_l2(h.float(), ref), BF16_TOL


@check
def cfm_ffn_glu_rows_as_one_sequence():
    """pad-separated utterances convolved as one sequence (row tiles span utterances): same values as the per-utterance
    tiling, bit for bit, and within the bf16 tolerance of the float64 evaluation"""
    lengths = (151, 97, 1, 150)
    h_flat, ref = _ffn_glu_case(lengths, 152, 1)
    h_tile, _ = _ffn_glu_case(lengths, 152, 0)
    if not torch.equal(h_flat, h_tile):
        return float("inf"), BF16_TOL
    return rel_l2(h_flat.float(), ref), BF16_TOL


@check
def cfm_ffn_out_norm():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs()
    b, n = ids.shape
    h = bf(torch.randn(b, n, 896, generator=g(10))) * mask[..., None]
    x = torch.randn(b, n, 256, generator=g(11))
    worst = 0.0
    for mode in (1, 2):
        gvec = 16.0 * (1.0 + 0.1 * torch.randn(256, generator=g(4))) if mode == 1 else s["model.transformer.final_norm.weight"]
        xd = x.clone().to(DEV)
        xn = torch.empty(b, n, 256, dtype=torch.bfloat16, device=DEV)
        nat.call("srb_cfm_ffn_out_norm", P(h.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_ff2[3]), P(pk.b_ff2[3]),
                 P(gvec.to(DEV)), mode, P(L), P(xd), P(xn), b, n)
        w = bf(s["model.transformer.layers.3.4.conv2.weight"]).double()
        ref_x = F.conv1d(h.double().transpose(1, 2), w, s["model.transformer.layers.3.4.conv2.bias"].double(), padding=1).transpose(1, 2) + x.double()
        if mode == 1:
            ref_n = ref_x / ref_x.norm(dim=-1, keepdim=True).clamp_min(1e-12) * gvec.double()
        else:
            ref_n = ref_x * torch.rsqrt(ref_x.pow(2).mean(-1, keepdim=True) + torch.finfo(torch.float32).eps) * gvec.double()
        ref_n = ref_n * mask[..., None]
        worst = max(worst, rel_l2(xd, ref_x) * 400, rel_l2(xn.float(), ref_n))
    return worst, BF16_TOL


@check
def cfm_pred_euler():
    s = sd()
    pk = packing.pack_cfm(s, DEV)
    ids, mask, L = _cfm_inputs()
    b, n = ids.shape
    xn = bf(torch.randn(b, n, 256, generator=g(12)))
    xt = torch.randn(b, n, 80, generator=g(13))
    xtd = xt.clone().to(DEV)
    xtb = torch.empty(b, n, 80, dtype=torch.bfloat16, device=DEV)
    # the mel outputs are compact: mel_rows = n - 3 rows per utterance (the caller's frame count), the rest is dropped
    mr = n - 3
    mel = torch.full((b, mr + 1, 80), 7.0, device=DEV)[:, :mr].contiguous()
    melb = torch.empty(b, mr, 80, dtype=torch.bfloat16, device=DEV)
    guard = torch.full((64,), 7.0, device=DEV)
    pv = oracle.pad_value()
    nat.call("srb_cfm_pred_euler", P(xn.to(DEV).to(torch.bfloat16).contiguous()), P(pk.w_pred), 0.0625, P(xtd), P(xtb),
             P(mel), P(melb), mr, 2.2615, -5.8843, pv, P(L), b, n)
    w = bf(s["model.to_pred.weight"]).double()
    ref_xt = xt.double() + F.linear(xn.double(), w) * 0.0625
    ref_mel = (ref_xt * 2.2615 + (-5.8843))[:, :mr]
    mk = mask[:, :mr]
    pad_ok = bool((mel.cpu()[~mk] == pv).all()) and bool((guard == 7.0).all())
    e = max(rel_l2(xtd, ref_xt), rel_l2(mel.cpu()[mk], ref_mel[mk]), rel_l2(xtb.float(), ref_xt) / 400,
            rel_l2(melb.float().cpu()[mk], ref_mel[mk]) / 400)
    return (e if pad_ok else 1.0), 2e-5


@check
def stage_inputs_bit_exact():
    """per-call staging: ids / prior padded from n to n8 rows, clamp as torch.clamp, bf16 copy, cleared regions"""
    b, n, n8 = 3, 37, 40
    ids = synthetic.make_units(b, n, seed=2, lengths=[37, 20, 1])
    x = torch.randn(b, n, 80, generator=g(1)) * 2
    ok = True
    for tv, has in ((1.0, 1), (0.0, 0), (0.0, 1), (-0.5, 1)):
        ids_ws = torch.full((b, n8), 9, dtype=torch.int64, device=DEV)
        xt = torch.full((b, n8, 80), 9.0, device=DEV)
        xb = torch.full((b, n8, 80), 9.0, dtype=torch.bfloat16, device=DEV)
        za = torch.full((48,), 9.0, device=DEV)
        zb = torch.full((16,), 9.0, device=DEV)
        nat.call("srb_stage_inputs", P(ids.to(DEV)), P(x.to(DEV)), P(ids_ws), P(xt), P(xb), P(za), 32 * 4, P(zb), 16 * 4, b, n, n8, tv, has)
        ref = torch.zeros(b, n8, 80)
        ref[:, :n] = torch.clamp(x, -tv, tv) if has else x
        ref_ids = torch.zeros(b, n8, dtype=torch.int64)
        ref_ids[:, :n] = ids
        ok = ok and torch.equal(xt.cpu(), ref) and torch.equal(xb.cpu(), ref.to(torch.bfloat16)) and torch.equal(ids_ws.cpu(), ref_ids)
        ok = ok and bool((za[:32] == 0).all()) and bool((za[32:] == 9).all()) and bool((zb == 0).all())
    return (0.0 if ok else 1.0), 0.0


@check
def unit_extents_exact():
    ids = synthetic.make_units(4, 50, seed=3, lengths=[50, 1, 49, 17])
    ids[2, 10] = 0                       # an interior pad: count 48, extent 49
    cnt = torch.empty(4, dtype=torch.int32, device=DEV)
    ext = torch.empty(4, dtype=torch.int32, device=DEV)
    nat.call("srb_unit_extents", P(ids.to(DEV)), P(cnt), P(ext), 4, 50)
    ok = cnt.cpu().tolist() == [50, 1, 48, 17] and ext.cpu().tolist() == [50, 1, 49, 17]
    return (0.0 if ok else 1.0), 0.0


@check
def hifigan_post_ragged_equals_dense_rows():
    """conv_post + tanh: the ragged store (utterances cropped to 320 len + 80, back to back) == the dense rows cropped"""
    s = sd()
    pk = packing.pack_vocoder(s, DEV)
    b, rows = 5, 320 * 30 + 80
    x = (torch.randn(b, rows, 16, generator=g(4)) * 0.5).to(torch.bfloat16).to(DEV)
    dense = torch.empty(b, rows, device=DEV)
    nat.call("srb_hifigan_post", P(x), P(pk.w_post), pk.b_post, P(dense), b, rows, None)
    lens = [30, 7, 29, 1, 30]
    L = torch.tensor(lens, dtype=torch.int32, device=DEV)
    total = sum(320 * n + 80 for n in lens)
    flat = torch.full((total + 8,), 7.0, device=DEV)
    nat.call("srb_hifigan_post", P(x), P(pk.w_post), pk.b_post, P(flat), b, rows, P(L))
    ok, off = bool((flat[total:] == 7.0).all()), 0
    for i, n in enumerate(lens):
        k = 320 * n + 80
        ok = ok and torch.equal(flat[off: off + k], dense[i, :k])
        off += k
    w = s["vocoder.conv_post.weight"].double()
    ref = torch.tanh(F.conv1d(x.cpu().double().transpose(1, 2), w, s["vocoder.conv_post.bias"].double(), padding=3))[:, 0]
    return (rel_l2(dense, ref) if ok else 1.0), 1e-5


def run_all(verbose: bool = True) -> List[Tuple[str, float, float, str]]:
    results = []
    for name, fn in CHECKS.items():
        try:
            err, tol = fn()
            torch.cuda.synchronize()
            status = "ok" if (err <= tol and err == err) else "FAIL"
        except Exception as exc:  # noqa: BLE001 - report and continue
            err, tol, status = float("nan"), float("nan"), f"ERROR {type(exc).__name__}: {str(exc)[:200]}"
        results.append((name, err, tol, status))
        if verbose:
            print(f"{name:40s} err={err:.3e} tol={tol:.1e} {status}", flush=True)
    return results
