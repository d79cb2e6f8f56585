"""In-kernel timeline of the fused MRF kernel (debug build: bash speech_resynth_b200/csrc/build.sh trace).

    SRB_DEBUG_LIB=build/libsrb_trace.so python tools/trace_mrf.py [channels]

CTA 0, first windows: the first epilogue warp's stamps (window start, raw rows loaded, then one per conv of the window)
and issuer 0's (raw rows visible, then one per conv issued), microseconds from the first stamp.
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from speech_resynth_b200 import _native as nat  # noqa: E402
from speech_resynth_b200 import packing, synthetic  # noqa: E402

P = nat.ptr

if __name__ == "__main__":
    c = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    lib = nat.load()
    lib.srb_debug_set_trace.argtypes = [ctypes.c_void_p]
    lib.srb_debug_set_trace.restype = None
    voc = packing.pack_vocoder(synthetic.make_state_dict(0), "cuda")
    stage = {32: 3, 16: 4}[c]
    rows = {32: 80040, 16: 160080}[c]
    b = 64
    u = (torch.randn(b, rows, c, device="cuda") * 0.3).to(torch.bfloat16)
    out = torch.empty_like(u)
    args = (P(u), P(voc.w_mrf[stage]), P(voc.b_mrf[stage]), P(out), b, rows, c, 0.1, 0.1 if c == 32 else 0.01)
    nat.call("srb_hifigan_mrf_fused", *args)
    torch.cuda.synchronize()
    trace = torch.zeros(3 * 256, dtype=torch.int64, device="cuda")
    lib.srb_debug_set_trace(trace.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.call("srb_hifigan_mrf_fused", *args)
    e1.record()
    torch.cuda.synchronize()
    lib.srb_debug_set_trace(None)
    t = trace.cpu().view(3, 256)
    t0 = int(t[1, 0])
    print(f"mrf C={c}: {e0.elapsed_time(e1) * 1e3:.0f} us")
    # SM clock while the kernel runs: cycles / wall time between the first and the last stamp of the epilogue warp
    n = int((t[1] != 0).sum())
    if n > 1:
        print(f" SM clock over the first {n} stamps: {(int(t[2, n - 1]) - int(t[2, 0])) / (int(t[1, n - 1]) - int(t[1, 0])) * 1e3:.0f} MHz")
    for role, nm in ((1, "epilogue warp 0"), (0, "issuer 0")):
        vals = [(int(v) - t0) / 1e3 for v in t[role] if int(v) != 0]
        per = 20 if role == 1 else 19
        for wi in range(0, min(len(vals), 4 * per), per):
            print(f" {nm} window {wi // per}: " + " ".join(f"{v:.2f}" for v in vals[wi:wi + per]))
