"""Time the decoder on a grid of bucket shapes (GPU box): eager first call (host enqueue time and device time), graph
capture, graph replay.  Feeds the per-bucket cost model of sharding.py and shows what a one-shot shape costs."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import sharding, synthetic  # noqa: E402

if __name__ == "__main__":
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.cuda()
    eng = decoder.engine()
    eng.reserve(64, 1000)
    shapes = [(64, 500), (36, 1000), (74, 500), (37, 500), (18, 500), (98, 384), (148, 250), (74, 250), (37, 250),
              (148, 128), (74, 128), (296, 128), (42, 852), (49, 720), (59, 600), (8, 500), (2, 687), (160, 200)]
    rows = []
    torch.cuda.synchronize()
    for b, n in shapes:
        ids = synthetic.make_units(b, n, seed=3)
        rec = {"batch": b, "frames": n, "tiles": b * sharding.tiles_per_utterance(n),
               "tflop": b * sharding.utterance_cost(n, 16) / 1e12}
        for phase in ("eager", "capture", "replay", "replay2"):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()
            decoder.resynthesize_flat(ids, 0.0625, 1.0)
            e1.record()
            host = (time.perf_counter() - t0) * 1e3
            torch.cuda.synchronize()
            rec[phase + "_host_ms"] = round(host, 3)
            rec[phase + "_dev_ms"] = round(e0.elapsed_time(e1), 3)
        rec["tflops_replay"] = round(rec["tflop"] / (rec["replay2_dev_ms"] / 1e3), 1)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
