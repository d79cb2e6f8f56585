"""Bitwise run-to-run determinism of the path at several sizes (GPU box only): same units + same prior twice."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import synthetic  # noqa: E402

if __name__ == "__main__":
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.cuda()
    eng = decoder.engine()
    bad = 0
    for b, n, lengths in ((64, 500, None), (12, 960, [960, 955, 951, 950, 949, 940, 930, 920, 915, 910, 905, 900]), (37, 333, None), (3, 3000, [3000, 2500, 1])):
        ids = synthetic.make_units(b, n, seed=5, lengths=lengths).cuda()
        x0 = torch.randn(b, n, 80, generator=torch.Generator().manual_seed(2)).cuda()
        mels, wavs = [], []
        for rep in range(3):
            mel = eng.sample(ids, 0.25, 1.0, noise=x0)
            wav, _, mel2 = eng.resynthesize(ids, 0.25, 1.0, noise=x0)
            mels.append(mel.clone())
            wavs.append(wav.clone())
            other = synthetic.make_units(5, 200, seed=rep).cuda()   # unrelated work in between
            eng.resynthesize(other, 0.5, 1.0)
        em = [bool(torch.equal(mels[0], m)) for m in mels[1:]]
        ew = [bool(torch.equal(wavs[0], w)) for w in wavs[1:]]
        dm = max(float((mels[0] - m).abs().max()) for m in mels[1:])
        dw = max(float((wavs[0] - w).abs().max()) for w in wavs[1:])
        print(f"B={b} N={n}: mel equal {em} (max diff {dm:.3e})  wav equal {ew} (max diff {dw:.3e})", flush=True)
        bad += (not all(em)) + (not all(ew))
    sys.exit(1 if bad else 0)
