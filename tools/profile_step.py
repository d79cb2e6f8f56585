"""One eager (un-captured) decoder pass at a given size -- the short command ncu wraps.  GPU box only.

    python tools/profile_step.py [batch frames nfe]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import engine as eng  # noqa: E402
from speech_resynth_b200 import synthetic  # noqa: E402

if __name__ == "__main__":
    b, n, nfe = (int(a) for a in (sys.argv[1:4] + ["64", "500", "2"][len(sys.argv) - 1:]))
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.cuda()
    e = eng.ResynthEngine(decoder.model.sampler(), decoder.vocoder.generator(), use_graphs=False)
    ids = synthetic.make_units(b, n, seed=7).cuda()
    wav, _, _ = e.resynthesize(ids, 1.0 / nfe, 1.0)
    torch.cuda.synchronize()
    print("ok", tuple(wav.shape), float(wav.abs().max()))
