"""Deterministic synthetic ("random-init") weights and inputs.

No checkpoint of the reference is reachable offline, so parity and the benchmark run
on a seeded random state-dict with the reference's key names and shapes.  The recipe
is self-contained (CPU ``torch.Generator``), so the build container, the GPU box, the live reference, the
oracle and the CUDA path all see bit-identical weights.

Scales follow torch's default Conv/Linear initialisation (uniform +-1/sqrt(fan_in)); the adaptive-norm
``to_weight`` matrices, zero in the reference constructor (norm.py:35), get a small random value so the
time conditioning is visible in parity tests (SURVEY.md "random-init degeneracy").
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

UPSAMPLE_RATES = (5, 4, 4, 2, 2)
UPSAMPLE_KERNELS = (10, 9, 8, 4, 4)
RESBLOCK_KERNELS = (3, 7, 11)


def state_dict_spec(depth: int = 4, hidden: int = 256, inter: int = 896, dim_in: int = 80, dim_cond: int = 768,
                    vocab: int = 2000, heads: int = 2) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for all 239 tensors of model.safetensors, in a fixed order."""
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    add = lambda k, s, kind: spec.append((k, tuple(s), kind))
    add("model.time_cond_mlp.0.weights", (hidden // 2,), "normal")
    add("model.time_cond_mlp.1.weight", (hidden, hidden + 1), "fan")
    add("model.time_cond_mlp.1.bias", (hidden,), "fan:%d" % (hidden + 1))
    add("model.to_cond_emb.weight", (vocab + 1, dim_cond), "embedding")
    add("model.to_embed.weight", (hidden, dim_in + dim_cond), "fan")
    add("model.to_embed.bias", (hidden,), "fan:%d" % (dim_in + dim_cond))
    add("model.conv_embed.dw_conv1d.0.weight", (hidden, 1, 31), "fan")
    add("model.conv_embed.dw_conv1d.0.bias", (hidden,), "fan:31")
    add("model.transformer.rotary_emb.inv_freq", (hidden // heads // 2,), "inv_freq")
    for i in range(depth):
        p = f"model.transformer.layers.{i}."
        add(p + "1.to_weight.weight", (hidden, hidden), "gamma")
        add(p + "2.to_qkv.weight", (3 * hidden, hidden), "fan")
        add(p + "2.to_out.weight", (hidden, hidden), "fan")
        add(p + "3.to_weight.weight", (hidden, hidden), "gamma")
        add(p + "4.conv1.weight", (2 * inter, hidden, 3), "fan")
        add(p + "4.conv1.bias", (2 * inter,), "fan:%d" % (hidden * 3))
        add(p + "4.conv2.weight", (hidden, inter, 3), "fan")
        add(p + "4.conv2.bias", (hidden,), "fan:%d" % (inter * 3))
    add("model.transformer.final_norm.weight", (hidden,), "norm_weight")
    add("model.to_pred.weight", (dim_in, hidden), "fan")
    add("vocoder.mean", (dim_in,), "zeros")
    add("vocoder.scale", (dim_in,), "ones")
    add("vocoder.conv_pre.weight", (512, dim_in, 7), "fan")
    add("vocoder.conv_pre.bias", (512,), "fan:%d" % (dim_in * 7))
    c = 512
    for i, (s, k) in enumerate(zip(UPSAMPLE_RATES, UPSAMPLE_KERNELS)):
        # ConvTranspose1d weight is (C_in, C_out, k); torch computes its fan_in from dim 1 (C_out * k)
        add(f"vocoder.upsampler.{i}.weight", (c, c // 2, k), "fan:%d" % (c // 2 * k))
        add(f"vocoder.upsampler.{i}.bias", (c // 2,), "fan:%d" % (c // 2 * k))
        c //= 2
        for j, rk in enumerate(RESBLOCK_KERNELS):
            for q in range(3):
                for name in ("convs1", "convs2"):
                    p = f"vocoder.resblocks.{i * 3 + j}.{name}.{q}."
                    add(p + "weight", (c, c, rk), "fan")
                    add(p + "bias", (c,), "fan:%d" % (c * rk))
    add("vocoder.conv_post.weight", (1, c, 7), "fan")
    add("vocoder.conv_post.bias", (1,), "fan:%d" % (c * 7))
    return spec


def make_state_dict(seed: int = 0, dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape, kind in state_dict_spec():
        if kind == "zeros":
            t = torch.zeros(shape)
        elif kind == "ones":
            t = torch.ones(shape)
        elif kind == "normal":
            t = torch.randn(shape, generator=g)
        elif kind == "embedding":
            t = torch.randn(shape, generator=g)
            t[0].zero_()  # padding_idx=0 row (models.py:51)
        elif kind == "gamma":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "norm_weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "inv_freq":
            d = shape[0] * 2  # transformer.py:47
            t = 1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))
        else:
            fan_in = int(kind.split(":")[1]) if ":" in kind else math.prod(shape[1:])
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[key] = t.to(dtype)
    return sd


def duration_predictor_state(seed: int = 0, dim_cond: int = 768) -> Dict[str, torch.Tensor]:
    """Extra tensors of the duration-prediction variant (models.py:71, fastspeech/modules.py:76-86): Conv1d(768 -> 1, k 3).
    Scaled so that round(exp(x) - 1) spreads over 0..4 frames per unit on the synthetic embedding table."""
    g = torch.Generator(device="cpu").manual_seed(1000 + seed)
    w = torch.randn(1, dim_cond, 3, generator=g) * (0.45 / math.sqrt(dim_cond * 3))
    return {"model.duration_predictor.conv.weight": w, "model.duration_predictor.conv.bias": torch.tensor([0.85])}


def make_units(batch: int, frames: int, seed: int = 7, lengths=None, vocab: int = 2000) -> torch.Tensor:
    """Random unit ids in [1, vocab] (ids = unit + 1, 0 = pad; synthesize.py:39-42), right-padded to ``frames``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ids = torch.randint(1, vocab + 1, (batch, frames), generator=g)
    if lengths is not None:
        for b, n in enumerate(lengths):
            ids[b, int(n):] = 0
    return ids


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Cheap fingerprint used by the golden fixtures to detect an RNG/recipe drift."""
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return tot
