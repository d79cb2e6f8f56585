"""CPU: the oracle restatement against the golden vectors minted from the live reference (tests/golden/*.npz),
its internal consistency (polyphase form, float64 anchor), and -- when /root/reference is mounted -- directly
against the live reference classes."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import cfm_hifigan_oracle as oracle
from oracle import ref_loader
from speech_resynth_b200 import synthetic


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _golden(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz")), json.load(open(os.path.join(golden_dir, "MANIFEST.json")))


def test_weight_recipe_is_stable(state_dict, golden_dir):
    man = json.load(open(os.path.join(golden_dir, "MANIFEST.json")))
    assert len(state_dict) == 239
    assert abs(synthetic.state_dict_checksum(state_dict) - man["weights_checksum"]) <= 1e-6 * man["weights_checksum"]
    assert float(state_dict["model.to_cond_emb.weight"][0].abs().sum()) == 0.0
    assert sum(v.numel() for k, v in state_dict.items() if k.startswith("vocoder.") and k not in ("vocoder.mean", "vocoder.scale")) == 12_977_473


def test_pad_value_is_the_float32_log():
    assert oracle.pad_value() == -11.512925148010254


@pytest.mark.parametrize("name", ["resynth_b2_n40", "resynth_b1_n64_dt01"])
def test_oracle_resynthesis_matches_reference_golden(state_dict, golden_dir, name):
    z, man = _golden(golden_dir, name)
    info = man["cases"][name]
    ids, x0 = torch.from_numpy(z["ids"]), torch.from_numpy(z["x0"])
    mel = oracle.sample(state_dict, ids, x0, info["dt"], info["truncation"])
    valid = ids.ne(0)
    assert rel_l2(mel[valid], torch.from_numpy(z["mel"])[valid]) <= 1e-5
    assert bool((mel[~valid] == oracle.pad_value()).all())
    wavs = oracle.resynthesize(state_dict, ids, x0, info["dt"], info["truncation"])
    assert [w.shape[-1] for w in wavs] == info["wav_lengths"] == [320 * int(n) + 80 for n in valid.sum(1)]
    ref = np.split(z["wav_flat"], np.cumsum(z["wav_lengths"])[:-1])
    for w, r in zip(wavs, ref):
        assert tuple(w.shape) == (1, len(r))
        assert rel_l2(w[0], torch.from_numpy(r)) <= 1e-4


def test_oracle_vocoder_and_polyphase_match_reference_golden(state_dict, golden_dir):
    z, _ = _golden(golden_dir, "vocoder_b2_t30")
    mel = torch.from_numpy(z["mel"])
    ref = torch.from_numpy(z["wav"])
    assert rel_l2(oracle.hifigan(state_dict, mel), ref) <= 1e-5
    sd64 = oracle.to_dtype(state_dict, torch.float64)
    assert rel_l2(oracle.hifigan(sd64, mel.double(), polyphase=True), ref) <= 1e-5


def test_oracle_velocity_and_time_embedding_match_reference_golden(state_dict, golden_dir):
    z, _ = _golden(golden_dir, "velocity_b2_n40")
    ids, xt = torch.from_numpy(z["ids"]), torch.from_numpy(z["xt"])
    t = torch.tensor(float(z["t"]))
    mask = ids.ne(0)
    v = oracle.velocity(state_dict, xt, oracle.embed_gather(state_dict["model.to_cond_emb.weight"], ids), mask, t)
    assert rel_l2(v[mask], torch.from_numpy(z["v"])[mask]) <= 1e-5
    assert rel_l2(oracle.time_embedding(state_dict, t), torch.from_numpy(z["time_emb"])) <= 1e-6


def test_oracle_gather_fingerprint(state_dict, golden_dir):
    z, man = _golden(golden_dir, "gather_b4_n33")
    emb = oracle.embed_gather(state_dict["model.to_cond_emb.weight"], torch.from_numpy(z["ids"]))
    assert hashlib.sha256(emb.numpy().tobytes()).hexdigest() == man["cases"]["gather_b4_n33"]["sha256"]


def test_polyphase_transposed_conv_edge_cases():
    """phase/tap bookkeeping of the polyphase form for every up-sampler geometry, tiny and ragged lengths"""
    g = torch.Generator().manual_seed(0)
    for k, s in zip(oracle.UPSAMPLE_KERNELS, oracle.UPSAMPLE_RATES):
        for lin in (1, 2, 3, 17):
            x = torch.randn(2, 6, lin, generator=g, dtype=torch.float64)
            w = torch.randn(6, 4, k, generator=g, dtype=torch.float64)
            b = torch.randn(4, generator=g, dtype=torch.float64)
            ref = torch.nn.functional.conv_transpose1d(x, w, b, stride=s, padding=(k - s) // 2)
            got = oracle.conv_transpose1d_polyphase(x, w, b, s, (k - s) // 2)
            assert got.shape == ref.shape and rel_l2(got, ref) <= 1e-12


def test_waveform_lengths_and_flop_model():
    assert oracle.waveform_lengths(torch.tensor([1, 25, 500])).tolist() == [400, 8080, 160080]
    assert oracle.vocoder_flops(500) == 160_305_440_256      # SURVEY.md section 8(a) row V*
    assert oracle.vocoder_flops(250) == 80_190_240_256
    assert oracle.transformer_flops(500, 16, hoisted=False) == 16 * 500 * 21_151_232


def test_ode_time_grid_matches_torch_arange_semantics():
    assert len(oracle.ode_times(0.0625)) == 16 and len(oracle.ode_times(0.1)) == 10
    assert float(oracle.ode_times(0.1)[3]) == float(torch.arange(0, 1, 0.1)[3]) != 0.3


def test_batch_composition_independence_of_valid_frames(state_dict):
    """an utterance's valid mel frames do not depend on what it is batched with (masks isolate utterances)"""
    ids = synthetic.make_units(2, 24, seed=3, lengths=[24, 9])
    x0 = torch.randn(2, 24, 80, generator=torch.Generator().manual_seed(2))
    sd64 = oracle.to_dtype(state_dict, torch.float64)
    both = oracle.sample(sd64, ids, x0.double(), 0.5, 1.0)
    alone = oracle.sample(sd64, ids[1:, :9], x0[1:, :9].double(), 0.5, 1.0)
    assert rel_l2(both[1, :9], alone[0]) <= 1e-10


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference is only mounted in the build container")
def test_oracle_against_live_reference(state_dict):
    ref = ref_loader.build_reference_model(state_dict)
    ids = synthetic.make_units(2, 20, seed=17, lengths=[20, 11])
    with torch.backends.mkldnn.flags(enabled=False):   # see DESIGN.md: oneDNN fp32 deconvolution defect
        torch.manual_seed(4)
        with torch.inference_mode():
            ref_wavs = ref(ids, 0.25, 1.0)
        torch.manual_seed(4)
        x0 = torch.randn(2, 20, 80)
        wavs = oracle.resynthesize(state_dict, ids, x0, 0.25, 1.0)
    for a, r in zip(wavs, ref_wavs):
        assert a.shape == r.shape and rel_l2(a, r) <= 1e-5


def test_oracle_duration_variant_matches_reference_golden(state_dict, golden_dir):
    """Duration-prediction variant (models.py:157-164): integer durations and the expanded ids are bit exact against the
    live reference, the mel sampled on the expanded sequence within fp32 noise."""
    from speech_resynth_b200 import synthetic

    z = np.load(os.path.join(golden_dir, "duration_b3_n48.npz"))
    sd = dict(state_dict, **synthetic.duration_predictor_state(0))
    ids = torch.from_numpy(z["ids"])
    dur = oracle.duration_predict(sd, ids)
    assert torch.equal(dur, torch.from_numpy(z["durations"]))
    exp_ids, lengths = oracle.length_regulate_ids(ids, dur)
    assert torch.equal(exp_ids, torch.from_numpy(z["expanded_ids"]))
    assert [320 * int(n) + 80 for n in lengths] == z["wav_lengths"].tolist()
    mel = oracle.sample(sd, exp_ids, torch.from_numpy(z["x0"]), 0.25, 1.0)
    valid = exp_ids.ne(0)
    ref = torch.from_numpy(z["mel"])
    assert float((mel[valid] - ref[valid]).norm() / ref[valid].norm()) <= 1e-5


def test_length_regulator_all_zero_rule():
    """HF:113-114: when every predicted duration of the batch is 0, all of them become 1 (pads included)."""
    ids = torch.tensor([[5, 9, 0], [7, 0, 0]])
    out, lengths = oracle.length_regulate_ids(ids, torch.zeros(2, 3, dtype=torch.long))
    assert torch.equal(out, ids) and lengths.tolist() == [3, 3]


def test_oracle_log_mel_matches_reference_golden(golden_dir):
    """mel_spectrogram (hifigan/data.py:17-53): the golden is the live reference function run with the restated librosa
    filter bank injected (librosa is not installed; the bank itself is the unpinned piece)."""
    z = np.load(os.path.join(golden_dir, "logmel_b2.npz"))
    mel = oracle.mel_spectrogram(torch.from_numpy(z["y"]))
    assert tuple(mel.shape) == (2, 80, 20)
    assert float((mel - torch.from_numpy(z["mel"])).abs().max()) <= 1e-5


def test_mel_filter_bank_properties():
    """Both restatements of librosa.filters.mel (oracle: scalar loops, package: vectorised numpy) agree, every band is a
    non-negative triangle, and Slaney's area normalisation holds: sum(weights) * (sr / n_fft) ~ 1 per band."""
    from speech_resynth_b200.features import mel_filter_bank

    a = oracle.librosa_mel_filter_bank()
    b = torch.from_numpy(mel_filter_bank())
    assert a.shape == (80, 201) and float((a - b).abs().max()) <= 1e-7
    assert bool((a >= 0).all()) and bool((a.sum(1) > 0).all())
    area = a.double().sum(1) * (16000 / 400)
    assert float((area[5:] - 1).abs().max()) <= 0.15      # discretisation of narrow triangles on a 40 Hz grid
    peak = a.argmax(1)
    assert bool((peak[1:] >= peak[:-1]).all())


def test_mel_filter_bank_matches_third_party_librosa_compatible_banks():
    """librosa (the reference's source of the bank, hifigan/data.py:6,33) is not installed here, so the restated bank is
    pinned against the two librosa-compatible implementations that are: transformers.audio_utils.mel_filter_bank
    (norm="slaney", mel_scale="slaney": the form HF's own feature extractors use in place of librosa.filters.mel; transformers
    is a pinned dependency of the reference) and, when importable, torchaudio.functional.melscale_fbanks."""
    from transformers.audio_utils import mel_filter_bank as hf_bank

    from speech_resynth_b200.features import mel_filter_bank

    ours = np.asarray(oracle.librosa_mel_filter_bank(), dtype=np.float64)            # (80, 201)
    hf = hf_bank(num_frequency_bins=201, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0, sampling_rate=16000,
                 norm="slaney", mel_scale="slaney").T
    assert ours.shape == hf.shape == (80, 201)
    assert float(np.abs(ours - hf).max()) <= 1e-8 * 1.0 + 1e-8
    assert float(np.abs(np.asarray(mel_filter_bank(), dtype=np.float64) - hf).max()) <= 1e-7
    try:
        import torchaudio
    except Exception:
        return
    ta = torchaudio.functional.melscale_fbanks(201, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="slaney").T.double().numpy()
    assert float(np.abs(ours - ta).max()) <= 2e-7


# ------------------------------------------------------------------------------------------------ unit quantiser oracle
def test_kmeans_oracle_matches_sklearn_golden_and_live_predict(golden_dir):
    """oracle/kmeans_oracle.py against scikit-learn's own KMeans.predict -- the call textlesslib's quantiser makes
    (utils/textless.py:9-21): the committed golden, then a live predict on a fresh codebook."""
    from oracle import kmeans_oracle as ko

    z = np.load(os.path.join(golden_dir, "kmeans_k300_d64.npz"))
    assert np.array_equal(ko.assign(z["feats"], z["centroids"]), z["labels"])
    assert float(ko.margins(z["feats"], z["centroids"]).min()) > 1e-4      # the golden has no near-ties
    from sklearn.cluster import KMeans

    rng = np.random.default_rng(1)
    train = rng.normal(size=(500, 16)).astype(np.float32)
    km = KMeans(n_clusters=40, n_init=1, max_iter=3, random_state=0).fit(train)
    x = rng.normal(size=(700, 16)).astype(np.float32)
    got, ref = ko.assign(x, km.cluster_centers_), km.predict(x)
    decisive = ko.margins(x, km.cluster_centers_) > 1e-5
    assert np.array_equal(got[decisive], ref[decisive]) and decisive.mean() > 0.99


def test_kmeans_oracle_deduplication_follows_torch_unique_consecutive(golden_dir):
    from oracle import kmeans_oracle as ko

    z = np.load(os.path.join(golden_dir, "kmeans_k300_d64.npz"))
    lengths = z["lengths"].tolist()
    ids, counts, n_out = ko.encode(z["feats"], lengths, z["centroids"], deduplicate=True)
    for b, n in enumerate(lengths):
        u, c = torch.unique_consecutive(torch.from_numpy(z["labels"][b, :n] + 1), return_counts=True)
        assert n_out[b] == len(u) and np.array_equal(ids[b, : len(u)], u.numpy()) and np.array_equal(counts[b, : len(u)], c.numpy())
        assert not ids[b, len(u):].any() and not counts[b, len(u):].any() and counts[b].sum() == n
    assert n_out[0] < lengths[0]           # the golden really has runs
    u, c = ko.unique_consecutive(np.array([], dtype=np.int64))
    assert len(u) == 0 and len(c) == 0
