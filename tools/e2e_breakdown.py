"""Where does the end-to-end step spend its extra time over the device-resident step?  (GPU box only)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import synthetic  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.to(dev)
    eng = decoder.engine()
    ids_host = synthetic.make_units(64, 500, seed=7).pin_memory()
    ids_dev = ids_host.to(dev)
    rows = 320 * 500 + 80
    wav_host = torch.empty(64, rows, dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)

    def timed(fn, steps=10, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(copy_stream)
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, 1e3 * t_enq / steps

    def a():   # device-resident engine call
        eng.resynthesize(ids_dev, 0.0625, 1.0)

    def b():   # public forward, device ids, no read-back
        decoder(ids_dev, 0.0625, 1.0)

    def c():   # + H2D of ids
        decoder(ids_host.to(dev, non_blocking=True), 0.0625, 1.0)

    def d():   # + D2H on the side stream as one copy of the whole (cloned) buffer
        wavs = decoder(ids_host.to(dev, non_blocking=True), 0.0625, 1.0)
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            full = torch.empty(0, dtype=torch.float32, device=dev).set_(wavs[0].untyped_storage(), 0, (64, rows), (rows, 1))
            full.record_stream(copy_stream)
            wav_host.copy_(full, non_blocking=True)

    def e():   # + D2H per utterance (bench.py's e2e step)
        wavs = decoder(ids_host.to(dev, non_blocking=True), 0.0625, 1.0)
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            for i, w in enumerate(wavs):
                w.record_stream(copy_stream)
                wav_host[i, : w.shape[-1]].copy_(w[0], non_blocking=True)

    for name, fn in (("engine", a), ("forward", b), ("forward+h2d", c), ("forward+h2d+d2h(one copy)", d), ("forward+h2d+d2h(per utt)", e)):
        ms, enq = timed(fn)
        print(f"{name:32s} {ms:7.3f} ms/step on the device clock, {enq:6.3f} ms/step of host enqueue time", flush=True)
