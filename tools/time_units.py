"""Time the unit quantiser (srb_kmeans_assign: split copy, score GEMM + arg-max epilogue, decode) at the headline shape:
64 x 500 frames of 768-wide features against the 2000-entry codebook.  Prints per-launch CUDA-event times."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_resynth_b200 import _native as nat  # noqa: E402
from speech_resynth_b200.units import UnitQuantizer  # noqa: E402


def main():
    gen = torch.Generator().manual_seed(0)
    cents = torch.randn(2000, 768, generator=gen)
    q = UnitQuantizer(cents, device="cuda")
    x = (cents[torch.randint(0, 2000, (64, 500), generator=gen)] + 0.3 * torch.randn(64, 500, 768, generator=gen)).cuda()
    rows = 64 * 500
    split_ws = torch.empty(rows, 3 * 768, dtype=torch.bfloat16, device="cuda")
    keys = torch.empty(rows, dtype=torch.int64, device="cuda")
    units = torch.empty(rows, dtype=torch.int64, device="cuda")
    P = nat.ptr
    steps = [
        ("srb_split_bf16", lambda: nat.call("srb_split_bf16", P(x), P(split_ws), rows, 768, P(keys))),
        ("srb_kmeans_scores_argmax", lambda: nat.call("srb_kmeans_scores_argmax", P(split_ws), P(q.packed), P(q.bias), P(keys), rows, 768, q.packed.shape[0])),
        ("srb_kmeans_decode", lambda: nat.call("srb_kmeans_decode", P(keys), P(units), rows, 1, None, 1)),
        ("srb_kmeans_assign (all three)", lambda: q.predict(x)),
    ]
    out = {}
    for name, fn in steps:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 20 * 1e3
    flops_alg = 2.0 * rows * 2000 * 768
    us = out["srb_kmeans_scores_argmax"]
    out["score_gemm_tflops_executed_bf16"] = 2.0 * rows * 2048 * 2304 / (us * 1e-6) / 1e12
    out["score_gemm_tflops_algorithmic_fp32_equiv"] = flops_alg / (us * 1e-6) / 1e12
    out["bytes_split_kernel_GBps"] = rows * 768 * (4 + 6) / (out["srb_split_bf16"] * 1e-6) / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
