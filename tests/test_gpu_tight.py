"""Tight-precision mode (`set_precision("tight")`, libsrb_tight.so) against the live reference's fp32 goldens.

BASELINE.json north_star asks for parity "tight in fp32/TF32, looser in bf16".  The tight mode runs the SAME tcgen05 kernels
over split bf16 operands (x = hi + lo, W = Wh + Wl; xh Wh + xl Wh + xh Wl accumulated in fp32: include/srb.h,
srb_split_factor) and fp32 CUDA-core attention, so its error against the reference's fp32 CPU run is that of fp32
arithmetic in a different summation order.  Tolerances (relative L2) are about 3x what a B200 measures; every comparison
appends its figure to gpurun_out/parity_errors.txt.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cfm_hifigan_oracle as oracle
from speech_resynth_b200 import synthetic
from tests.test_gpu_e2e import MEAN, STD, rel_l2, within

pytestmark = pytest.mark.gpu

TIGHT_MEL_TOL_NORM, TIGHT_MEL_TOL_RAW, TIGHT_WAV_TOL = 2.0e-5, 1.0e-5, 1.0e-5


@pytest.fixture(scope="module")
def tight_decoder(state_dict):
    import speech_resynth_b200 as srb

    m = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda()
    m.set_precision("tight")
    assert m.precision == m.model.precision == m.vocoder.precision == "tight"
    return m


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz")), json.load(open(os.path.join(golden_dir, "MANIFEST.json")))["cases"][name]


def test_tight_library_reports_its_format():
    from speech_resynth_b200 import _native as nat

    assert nat.load(tight=True).srb_split_factor() == 3
    assert nat.load().srb_split_factor() == 1


@pytest.mark.parametrize("name", ["resynth_b2_n40", "resynth_b1_n64_dt01", "resynth_b3_n150"])
def test_tight_resynthesis_matches_reference_golden(tight_decoder, golden_dir, name):
    z, info = _load(golden_dir, name)
    ids = torch.from_numpy(z["ids"]).cuda()
    x0 = torch.from_numpy(z["x0"]).cuda()
    wav, lengths, mel = tight_decoder.engine().resynthesize(ids, info["dt"], info["truncation"], noise=x0)
    torch.cuda.synchronize()
    ref_mel = torch.from_numpy(z["mel"])
    valid = ids.ne(0).cpu()
    mel = mel.cpu()
    assert torch.equal(lengths.cpu().long(), valid.sum(1))
    assert bool((mel[~valid] == oracle.pad_value()).all())
    within(f"tight {name} mel raw", rel_l2(mel[valid], ref_mel[valid]), TIGHT_MEL_TOL_RAW)
    within(f"tight {name} mel normalised", rel_l2((mel[valid] - MEAN) / STD, (ref_mel[valid] - MEAN) / STD), TIGHT_MEL_TOL_NORM)
    ref_wavs = np.split(z["wav_flat"], np.cumsum(z["wav_lengths"])[:-1])
    worst = 0.0
    for b, ref_w in enumerate(ref_wavs):
        n = len(ref_w)
        worst = max(worst, rel_l2(wav[b, :n], torch.from_numpy(ref_w)))
    within(f"tight {name} waveform", worst, TIGHT_WAV_TOL)


def test_tight_vocoder_golden(tight_decoder, golden_dir):
    """The vocoder alone (decoder.vocoder(mel), train.py:59) against the live reference's waveform for a given mel
    (HF:1451-1491)."""
    z, _ = _load(golden_dir, "vocoder_b2_t30")
    wav = tight_decoder.vocoder(torch.from_numpy(z["mel"]).cuda()).cpu()
    within("tight vocoder_b2_t30 waveform", rel_l2(wav, torch.from_numpy(z["wav"])), TIGHT_WAV_TOL)


def test_tight_sample_matches_oracle_ragged(tight_decoder, state_dict):
    """Ragged batch whose frame count is not a multiple of 8 (pad rows inside the padded workspace), NFE 8, against the
    fp32 CPU oracle; then the public call: list of cropped waveforms, same seed => same output."""
    ids = synthetic.make_units(4, 90, seed=31, lengths=[90, 50, 7, 1])
    x0 = torch.randn(4, 90, 80, generator=torch.Generator().manual_seed(5))
    mel = tight_decoder.engine().sample(ids.cuda(), 0.125, 1.0, noise=x0.cuda()).cpu()
    ref = oracle.sample(state_dict, ids, x0, 0.125, 1.0)
    valid = ids.ne(0)
    assert bool((mel[~valid] == oracle.pad_value()).all())
    within("tight sample 4x90 NFE 8 mel normalised", rel_l2((mel[valid] - MEAN) / STD, (ref[valid] - MEAN) / STD), TIGHT_MEL_TOL_NORM)
    torch.manual_seed(3)
    a = tight_decoder(ids.cuda(), 0.25, 1.0)
    torch.manual_seed(3)
    b = tight_decoder(ids.cuda(), 0.25, 1.0)
    assert [tuple(w.shape) for w in a] == [(1, 320 * n + 80) for n in [90, 50, 7, 1]]
    assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_precision_switch_rebuilds_and_modes_agree(state_dict):
    """set_precision switches libraries on a live model; the two modes agree with each other to the bf16 tolerance and the
    tight one is the closer of the two to the oracle."""
    import speech_resynth_b200 as srb

    m = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda()
    ids = synthetic.make_units(2, 64, seed=4, lengths=[64, 33])
    x0 = torch.randn(2, 64, 80, generator=torch.Generator().manual_seed(6))
    ref = oracle.sample(state_dict, ids, x0, 0.25, 1.0)
    valid = ids.ne(0)
    assert m.precision == os.environ.get("SRB_PRECISION", "bf16")
    m.set_precision("bf16")
    mel_bf16 = m.engine().sample(ids.cuda(), 0.25, 1.0, noise=x0.cuda()).cpu()
    m.set_precision("tight")
    assert m.engine().tight and m.engine().sampler.tight and m.engine().vocoder.tight
    mel_tight = m.engine().sample(ids.cuda(), 0.25, 1.0, noise=x0.cuda()).cpu()
    e_bf16 = rel_l2((mel_bf16[valid] - MEAN) / STD, (ref[valid] - MEAN) / STD)
    e_tight = rel_l2((mel_tight[valid] - MEAN) / STD, (ref[valid] - MEAN) / STD)
    within("precision switch: tight vs oracle", e_tight, TIGHT_MEL_TOL_NORM)
    assert e_tight < 0.2 * e_bf16, (e_tight, e_bf16)
    m.set_precision("bf16")
    again = m.engine().sample(ids.cuda(), 0.25, 1.0, noise=x0.cuda()).cpu()
    assert torch.equal(again, mel_bf16)
    with pytest.raises(ValueError):
        m.set_precision("fp8")
