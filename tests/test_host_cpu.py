"""CPU: host-side logic -- drop-in API surface (config / state-dict / save-load), weight packing layouts, the
bucketing + sharding plan (incl. a world_size-2 gloo run), and the C-ABI library's exported symbols."""
import ctypes
import os
import re
import tempfile

import pytest
import torch
import torch.nn.functional as F

import speech_resynth_b200 as srb
from speech_resynth_b200 import _native, packing, sharding, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------ drop-in API
@pytest.fixture(scope="module")
def model(state_dict):
    m = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    m.load_state_dict(state_dict, strict=True)
    return m


def test_state_dict_keys_and_shapes_match_the_reference_checkpoint_layout(model, state_dict):
    sd = model.state_dict()
    assert set(sd) == set(state_dict) and len(sd) == 239
    assert all(sd[k].shape == state_dict[k].shape for k in sd)
    assert "model.time_cond_mlp.0.weights" in sd and "model.transformer.rotary_emb.inv_freq" in sd  # persistent buffers


def test_constructor_initialisation_follows_the_reference():
    m = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config())
    assert float(m.model.transformer.layers[0][1].to_weight.weight.abs().sum()) == 0.0      # norm.py:35
    assert float(m.model.to_cond_emb.weight[0].abs().sum()) == 0.0                           # padding_idx row
    inv = m.model.transformer.rotary_emb.inv_freq
    assert torch.allclose(inv, 1.0 / (10000 ** (torch.arange(0, 128, 2).float() / 128)))
    assert m.model.transformer.layers[0][0] is None                                          # no U-Net skip combiner


def test_save_pretrained_from_pretrained_round_trip(model, state_dict):
    with tempfile.TemporaryDirectory() as d:
        model.save_pretrained(d)
        assert sorted(os.listdir(d)) == ["config.json", "model.safetensors"]
        again = srb.ConditionalFlowMatchingWithHifiGan.from_pretrained(d)
        assert all(torch.equal(v, state_dict[k]) for k, v in again.state_dict().items())
        assert again.config.model_config.hidden_size == 256 and again.config.vocoder_config.upsample_rates == [5, 4, 4, 2, 2]
        model.model.save_pretrained(os.path.join(d, "cfm"))
        model.vocoder.save_pretrained(os.path.join(d, "voc"))
        joined = srb.ConditionalFlowMatchingWithHifiGan.load_pretrained(os.path.join(d, "cfm"), os.path.join(d, "voc"))
        assert all(torch.equal(v, state_dict[k]) for k, v in joined.state_dict().items())


def test_waveform_length_rule(model):
    assert model._get_waveform_lengths(torch.tensor([1, 25, 500])).tolist() == [400, 8080, 160080]


def test_no_cpu_fallback(model):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_native.NativeLibraryError):
        model(torch.ones(1, 8, dtype=torch.long))
    with pytest.raises(RuntimeError):
        model.model.conv_embed(torch.zeros(1, 4, 256))   # parameter containers have no PyTorch forward


def test_unsupported_variants_are_rejected():
    cfg = srb.ConditionalFlowMatchingConfig(use_unet_skip_connection=True)
    with pytest.raises(NotImplementedError):
        srb.ConditionalFlowMatchingModel(cfg)


def test_duration_variant_has_the_reference_parameter_names(state_dict):
    """predict_duration adds exactly `duration_predictor.conv.{weight (1, 768, 3), bias (1,)}` (models.py:71,
    fastspeech/modules.py:86) and the extended state-dict loads strictly."""
    from speech_resynth_b200 import synthetic

    m = srb.ConditionalFlowMatchingModel(srb.ConditionalFlowMatchingConfig(predict_duration=True))
    extra = synthetic.duration_predictor_state(0)
    keys = set(m.state_dict())
    assert {k[len("model."):] for k in extra} <= keys
    assert tuple(m.state_dict()["duration_predictor.conv.weight"].shape) == (1, 768, 3)
    sd = {k[len("model."):]: v for k, v in dict(state_dict, **extra).items() if k.startswith("model.")}
    m.load_state_dict(sd, strict=True)


# ------------------------------------------------------------------------------------------------ packing
def test_conv_weight_packing_is_tap_major_with_channel_padding():
    w = torch.randn(6, 80, 7)
    p = packing.pack_conv_weight(w, 64).float()
    assert p.shape == (6, 7 * 128)
    p = p.view(6, 7, 128)
    assert torch.equal(p[:, :, :80], w.permute(0, 2, 1).to(torch.bfloat16).float()) and float(p[:, :, 80:].abs().sum()) == 0.0


def test_glu_row_permutation_pairs_value_and_gate_rows():
    perm = packing.glu_row_permutation(896)
    assert sorted(perm.tolist()) == list(range(1792))
    for t in range(7):
        blk = perm[t * 256:(t + 1) * 256]
        assert blk[:128].tolist() == list(range(t * 128, (t + 1) * 128))
        assert blk[128:].tolist() == list(range(896 + t * 128, 896 + (t + 1) * 128))


@pytest.mark.parametrize("k,s", list(zip(packing.UPSAMPLE_KERNELS, packing.UPSAMPLE_RATES)))
def test_polyphase_packing_reproduces_conv_transpose(k, s):
    """emulate the kernel's gather form from the packed matrix and compare with F.conv_transpose1d (fp64 operands
    rounded to bf16 first, so the packed bf16 weights are exact)"""
    g = torch.Generator().manual_seed(k * 10 + s)
    c_in, c_out, lin = 8, 4, 19
    w = torch.randn(c_in, c_out, k, generator=g).to(torch.bfloat16).float()
    x = torch.randn(1, lin, c_in, generator=g)
    pad = (k - s) // 2
    ref = F.conv_transpose1d(x.transpose(1, 2).double(), w.double(), None, stride=s, padding=pad).transpose(1, 2)[0]
    wp = packing.pack_upsampler_weight(w, s).double()          # (c_out, k * c_in), phases concatenated along K
    lout = ref.shape[0]
    out = torch.zeros(lout, c_out, dtype=torch.float64)
    col = 0
    for r in range(s):
        j0, c_r = (r + pad) % s, (r + pad) // s
        for m, _ in enumerate(range(j0, k, s)):
            wt = wp[:, col:col + c_in]
            col += c_in
            for q in range((lout - r + s - 1) // s):
                src = q + c_r - m
                if 0 <= src < lin:
                    out[q * s + r] += wt @ x[0, src].double()
    assert col == k * c_in
    assert float((out - ref).abs().max()) <= 1e-12


def test_config_accepts_the_reference_positional_order():
    """The reference's config takes its 15 fields positionally (src/flow_matching/configs.py:7-24)."""
    from speech_resynth_b200.configs import ConditionalFlowMatchingConfig

    c = ConditionalFlowMatchingConfig(1000, 80, 768, 256, 4, 2, 896, 0.0, False, 31, 256, 0.0, -5.0, 2.0, True)
    assert (c.vocab_size, c.mean, c.std, c.predict_duration, c.depth) == (1000, -5.0, 2.0, True, 4)
    assert ConditionalFlowMatchingConfig(500, dim_in=80).vocab_size == 500
    with pytest.raises(TypeError):
        ConditionalFlowMatchingConfig(500, vocab_size=2000)
    if os.path.isdir("/root/reference/src/flow_matching"):
        from oracle import ref_loader

        _, _, ref_cfg = ref_loader.load_reference()
        assert ref_cfg(1000, 80, 768, 256, 4, 2, 896, 0.0, False, 31, 256, 0.0, -5.0, 2.0, True).to_dict().items() >= {
            k: getattr(c, k) for k, _ in __import__("speech_resynth_b200.configs", fromlist=["_CFM_FIELDS"])._CFM_FIELDS}.items()


# ------------------------------------------------------------------------------------------------ sharding
def test_buckets_partition_and_pad_like_pad_sequence():
    g = torch.Generator().manual_seed(11)
    lengths = torch.randint(100, 1001, (300,), generator=g).tolist()
    buckets = sharding.bucket_by_length(lengths, granularity=64, max_batch=32)
    seen = sorted(i for b in buckets for i in b.indices)
    assert seen == list(range(300))
    for b in buckets:
        ls = [lengths[i] for i in b.indices]
        assert b.frames == max(ls) and b.batch <= 32
        assert len({(n + 63) // 64 for n in ls}) == 1
    plan = sharding.assign_buckets(buckets, 8)
    assert sorted(j for r in plan for j in r) == list(range(len(buckets)))
    loads = [sum(buckets[j].batch * sharding.utterance_cost(buckets[j].frames, 16) for j in r) for r in plan]
    assert max(loads) / (sum(loads) / 8) < 1.15


def test_empty_utterance_is_rejected():
    with pytest.raises(ValueError):
        sharding.bucket_by_length([5, 0, 3])


def _fake_synth(ids):
    # deterministic stand-in for the decoder: waveform i = (sum of its ids) ramp of the reference's output length
    out = []
    for row in ids:
        n = int(row.ne(0).sum())
        out.append((torch.arange(320 * n + 80, dtype=torch.float32) * 1e-3 + float(row.sum())).unsqueeze(0))
    return out


def _fake_synth_into(ids, out):
    off = 0
    for w in _fake_synth(ids):
        out[off: off + w.shape[-1]].copy_(w[0])
        off += w.shape[-1]
    assert off == out.numel()


def test_single_rank_sharded_call_restores_caller_order():
    units = [torch.randint(1, 2001, (n,)) for n in (7, 300, 64, 65, 1)]
    for kw in (dict(strategy="lpt", granularity=64), dict(strategy="contiguous"), dict(strategy="contiguous", max_batch=2)):
        outs = sharding.resynthesize_sharded(units, _fake_synth, **kw)
        for u, w in zip(units, outs):
            assert tuple(w.shape) == (1, 320 * u.numel() + 80) and float(w[0, 0]) == float(u.sum())


def test_contiguous_plan_balances_cost_and_keeps_buckets_homogeneous():
    """BASELINE configs[2] (1024 utterances of 2-20 s, seed 11): cost-balanced contiguous ranges of the sorted list."""
    g = torch.Generator().manual_seed(11)
    lengths = torch.randint(100, 1001, (1024,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        plan = sharding.plan_shards(lengths, world, nfe=16)
        assert sorted(i for b in plan.buckets for i in b.indices) == list(range(1024))
        assert sorted(j for r in plan.per_rank for j in r) == list(range(len(plan.buckets)))
        for b in plan.buckets:
            ls = [lengths[i] for i in b.indices]
            assert b.frames == max(ls) == ls[0] and ls == sorted(ls, reverse=True)
            assert b.batch * sharding.tiles_per_utterance(b.frames) <= 296 and b.batch <= 160
            assert b.frames - min(ls) <= max(64, int(0.12 * b.frames))
        # rank r's utterances are all at least as long as rank r + 1's
        mins = [min(lengths[i] for j in r for i in plan.buckets[j].indices) for r in plan.per_rank]
        maxs = [max(lengths[i] for j in r for i in plan.buckets[j].indices) for r in plan.per_rank]
        assert all(mins[r] >= maxs[r + 1] for r in range(world - 1))
        assert plan.imbalance < 1.03, (world, plan.imbalance)    # 8 ranks: 1.050 before the boundary search, 1.027 with it
        # every rank derives the same plan from the lengths alone (the search is deterministic)
        again = sharding.plan_shards(lengths, world, nfe=16)
        assert [b.indices for b in again.buckets] == [b.indices for b in plan.buckets] and again.per_rank == plan.per_rank
        # padding overhead of the whole plan (padded frames / real frames)
        padded = sum(b.batch * b.frames for b in plan.buckets)
        assert padded / sum(lengths) < 1.08


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    units = [torch.randint(1, 2001, (int(n),), generator=g) for n in torch.randint(1, 200, (23,), generator=g)]
    outs = sharding.resynthesize_sharded(units, _fake_synth, rank=rank, world=world, strategy="lpt", granularity=32, max_batch=4)
    outs2 = sharding.resynthesize_sharded(units, None, rank=rank, world=world, max_batch=3, synth_into=_fake_synth_into)
    if rank == 0:
        assert all(torch.equal(a, b) for a, b in zip(outs, outs2))
    if rank == 0:
        ok = all(tuple(w.shape) == (1, 320 * u.numel() + 80) and float(w[0, 0]) == float(u.sum())
                 and float(w[0, -1]) == pytest.approx(float(u.sum()) + (320 * u.numel() + 79) * 1e-3, rel=1e-6)
                 for u, w in zip(units, outs))
        q.put(ok)
    else:
        assert outs is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_symbol_declared_in_the_header():
    header = open(os.path.join(ROOT, "include", "srb.h")).read()
    declared = set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_native.EXPORTED_SYMBOLS)
    if not os.path.exists(_native.LIB_PATH):
        pytest.skip("libsrb.so not built (run __graft_entry__.build())")
    # the product library and the tight-precision build of the same sources (include/srb.h: srb_split_factor)
    for path, split in ((_native.LIB_PATH, 1), (_native.TIGHT_LIB_PATH, 3)):
        assert os.path.exists(path), path
        lib = ctypes.CDLL(path)
        for name in declared:
            assert hasattr(lib, name), (path, name)
        assert lib.srb_version() == 100
        assert lib.srb_split_factor() == split


# ------------------------------------------------------------------------------------------------ synthesize driver
@pytest.mark.parametrize("bits", [32, 16])
def test_wav_writer_round_trip(tmp_path, bits):
    """The driver's RIFF writer (float32 like torchaudio.save of a float tensor, or 16-bit PCM) read back by scipy."""
    from scipy.io import wavfile

    from speech_resynth_b200.synthesize import write_wav

    x = (torch.rand(16123, generator=torch.Generator().manual_seed(3)) * 2 - 1).numpy()
    path = str(tmp_path / "sub" / f"a{bits}.wav")
    write_wav(path, x, 16000, bits)
    rate, y = wavfile.read(path)
    assert rate == 16000 and y.shape == x.shape
    if bits == 32:
        assert y.dtype.kind == "f" and (y == x).all()
    else:
        assert y.dtype == "int16" and abs(y.astype("float64") / 32768.0 - x).max() <= 0.5 / 32768 + 1e-9


def test_cost_model_matches_the_oracle_flop_count():
    """The package's own FLOP formulas (rank balancing, bench.py's roofline) against the oracle's and SURVEY.md 8(a)."""
    from oracle import cfm_hifigan_oracle as oracle

    for n in (1, 40, 250, 500, 3000):
        assert sharding.vocoder_flops(n) == oracle.vocoder_flops(n)
        for hoisted in (False, True):
            assert sharding.transformer_flops(n, 16, hoisted) == oracle.transformer_flops(n, 16, hoisted)
    assert sharding.vocoder_flops(500) == 160_305_440_256
    assert sharding.utterance_cost(500, 16) == sharding.transformer_flops(500, 16) + sharding.vocoder_flops(500)


def test_product_package_never_reaches_for_the_oracle_or_the_reference_tree():
    """The oracle is test infrastructure: nothing under speech_resynth_b200/ may import it, and nothing there may
    read /root/reference (absent on the GPU box)."""
    pkg = os.path.dirname(os.path.abspath(srb.__file__))
    offenders = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                continue
            text = open(os.path.join(root, f), encoding="utf-8").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "/root/reference" in text:
                offenders.append(os.path.relpath(os.path.join(root, f), pkg))
    assert offenders == []


@pytest.mark.parametrize("c_in,c_out,k,s", [(128, 64, 8, 4), (64, 32, 4, 2), (32, 16, 4, 2)])
def test_upsampler_as_row_group_conv_reproduces_conv_transpose(c_in, c_out, k, s):
    """ConvTranspose1d with k - 2 pad = stride == a 3-tap Conv1d producing all output phases per input row (HF:1392-1402)."""
    g = torch.Generator().manual_seed(3)
    w = torch.randn(c_in, c_out, k, generator=g, dtype=torch.float64)
    b = torch.randn(c_out, generator=g, dtype=torch.float64)
    x = torch.randn(2, c_in, 37, generator=g, dtype=torch.float64)
    ref = F.conv_transpose1d(x, w, b, stride=s, padding=(k - s) // 2)            # (2, c_out, s * 37)
    wv, bv = packing.upsampler_as_row_group_conv(w, b, s)
    y = F.conv1d(x, wv.double(), bv.double(), padding=1)                          # (2, s c_out, 37)
    got = y.view(2, s, c_out, 37).permute(0, 2, 3, 1).reshape(2, c_out, s * 37)  # [b, co, s r + ph]
    assert ref.shape == got.shape
    assert float((ref - got).abs().max()) < 1e-5


def test_bucketing_and_rank_assignment_properties():
    """Property test (hypothesis) of the host-side sharding: buckets partition the utterances, are padded to their longest
    member, respect the batch / frame caps, and every bucket lands on exactly one rank with a balanced load (LPT bound)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=1, max_value=3000), min_size=1, max_size=200),
           st.sampled_from([16, 64, 128]), st.sampled_from([1, 8, 64]), st.integers(min_value=1, max_value=8))
    def check(lengths, granularity, max_batch, world):
        buckets = sharding.bucket_by_length(lengths, granularity=granularity, max_batch=max_batch)
        seen = sorted(i for b in buckets for i in b.indices)
        assert seen == list(range(len(lengths)))
        for b in buckets:
            member = [lengths[i] for i in b.indices]
            assert b.frames == max(member) and 1 <= b.batch <= max_batch
            assert len({(n + granularity - 1) // granularity for n in member}) == 1
        plan = sharding.assign_buckets(buckets, world)
        assert sorted(j for r in plan for j in r) == list(range(len(buckets)))
        cost = [b.batch * sharding.utterance_cost(b.frames, 16) for b in buckets]
        load = [sum(cost[j] for j in r) for r in plan]
        # greedy LPT: no rank exceeds the mean by more than the largest single bucket
        assert max(load) <= sum(cost) / world + max(cost) + 1e-6

    check()


# ------------------------------------------------------------------------------------------------ tight-precision packing
def test_split_operand_packing_recovers_fp32_products():
    """The tight-precision format (include/srb.h: srb_split_factor): activations [xh | xl | xh], weights [Wh | Wh | Wl]
    along K, every value a bf16 number -- a bf16 GEMM with fp32 accumulation over them gives x W up to the dropped
    xl Wl term (~2^-17 relative), against ~2^-8 for plain bf16 operands."""
    from speech_resynth_b200.packing import pack_conv_weight, split_operand

    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 48, generator=g)
    w = torch.randn(24, 48, generator=g)
    xh = x.to(torch.bfloat16).float()
    xl = (x - xh).to(torch.bfloat16).float()
    xs = torch.cat([xh, xl, xh], dim=1)
    ws = split_operand(w, 1)
    assert ws.shape == (24, 144) and torch.equal(ws, ws.to(torch.bfloat16).float())
    ref = x.double() @ w.double().t()
    tight = xs.double() @ ws.double().t()
    plain = xh.double() @ w.to(torch.bfloat16).double().t()
    e_tight = float((tight - ref).norm() / ref.norm())
    e_plain = float((plain - ref).norm() / ref.norm())
    assert e_tight < 2e-5 and e_plain > 1e-3, (e_tight, e_plain)
    # conv weights: the split runs along input channels, then the usual tap-major packing with channel padding
    cw = torch.randn(8, 20, 3, generator=g)
    p = pack_conv_weight(split_operand(cw, 1), 64)
    assert p.shape == (8, 3 * 64) and p.dtype == torch.bfloat16
    assert torch.equal(p[:, 64:124].float(), split_operand(cw[:, :, 1], 1))
    assert bool((p[:, 60:64] == 0).all())


def test_centroid_packing_for_the_unit_quantiser():
    """units.pack_centroids: split rows [Ch | Ch | Cl] padded to 64 columns, codebook padded to 256 rows that can never win
    (bias -inf), bias = -|c|^2 / 2 -- so argmax_j (x . c_j + bias_j) over the packed operands is KMeans.predict's argmin."""
    import numpy as np

    from oracle import kmeans_oracle as ko
    from speech_resynth_b200.packing import split_operand
    from speech_resynth_b200.units import pack_centroids

    g = torch.Generator().manual_seed(3)
    c = torch.randn(300, 24, generator=g)
    packed, bias = pack_centroids(c)
    assert packed.shape == (512, 128) and packed.dtype == torch.bfloat16 and bias.shape == (512,)
    assert torch.equal(packed[:300, :72].float(), split_operand(c, 1)) and not bool(packed[:300, 72:].any()) and not bool(packed[300:].any())
    assert bool(torch.isinf(bias[300:]).all()) and bool((bias[300:] < 0).all())
    assert torch.allclose(bias[:300], -0.5 * c.pow(2).sum(1), rtol=1e-6)
    x = torch.randn(50, 24, generator=g)
    xh = x.to(torch.bfloat16).float()
    xs = torch.cat([xh, (x - xh).to(torch.bfloat16).float(), xh], dim=1)
    scores = xs.double() @ packed[:, :72].double().t() + bias.double()
    assert np.array_equal(scores.argmax(1).numpy(), ko.assign(x.numpy(), c.numpy()))
    with pytest.raises(ValueError):
        pack_centroids(torch.randn(10, 20))        # width not a multiple of 8
