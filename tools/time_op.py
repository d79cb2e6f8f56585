"""Time one CUDA-core op of the path at config-2 size with CUDA events (A/B runs of env knobs; GPU box only).

    SRB_POSCONV_ROWS=32 python tools/time_op.py posconv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from speech_resynth_b200 import _native as nat  # noqa: E402
from speech_resynth_b200 import packing, synthetic  # noqa: E402

P = nat.ptr


def timed(fn, reps=50):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


if __name__ == "__main__":
    what = sys.argv[1]
    b, n = 64, 504
    dev = "cuda"
    pk = packing.pack_cfm(synthetic.make_state_dict(0), dev)
    L = torch.full((b,), 500, dtype=torch.int32, device=dev)
    if what == "posconv":
        x0 = torch.randn(b, n, 256, device=dev)
        g = torch.full((256,), 16.0, device=dev)
        x = torch.empty_like(x0)
        xn = torch.empty(b, n, 256, dtype=torch.bfloat16, device=dev)
        us = timed(lambda: nat.call("srb_cfm_posconv_norm", P(x0), P(pk.dw_w), P(pk.dw_b), P(g), P(L), P(x), P(xn), b, n))
    elif what == "post":
        rows = 160080
        xa = torch.randn(b, rows, 16, device=dev).to(torch.bfloat16)
        w = torch.randn(7 * 16, device=dev)
        wav = torch.empty(b, rows, device=dev)
        us = timed(lambda: nat.call("srb_hifigan_post", P(xa), P(w), 0.1, P(wav), b, rows, None))
    elif what == "pair":
        # srb_hifigan_pair_fused at the C = 64 stage's shape (SRB_PAIR_DEBUG=1: no MMAs, =2: no output stores -- timing only)
        rows = 40020
        k = int(sys.argv[2]) if len(sys.argv) > 2 else 3
        xa = torch.randn(b, rows, 64, device=dev).to(torch.bfloat16)
        out = torch.empty_like(xa)
        w1 = (torch.randn(64, k * 64, device=dev) * 0.05).to(torch.bfloat16)
        w2 = (torch.randn(64, k * 64, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.zeros(64, device=dev)
        us = timed(lambda: nat.call("srb_hifigan_pair_fused", P(xa), P(w1), P(bias), P(w2), P(bias), P(out), b, rows, 64, k, 1, 0.1), reps=20)
        what = f"pair k={k} SRB_PAIR_DEBUG={os.environ.get('SRB_PAIR_DEBUG', '0')}"
    else:
        raise SystemExit(f"unknown op {what}")
    print(f"{what} {os.environ.get('SRB_POSCONV_ROWS', '')}: {us:.1f} us (L2 flushed between launches)")
