"""Unit quantiser (speech_resynth_b200.units: srb_kmeans_assign, srb_unique_consecutive) on the GPU against scikit-learn's
KMeans.predict golden and the CPU oracle.  Integer work: bit exact wherever the assignment is decisive (the relative gap
between the two nearest centroids exceeds 1e-5; fp32-grade scores cannot order a closer pair reliably, nor can sklearn)."""
import os

import numpy as np
import pytest
import torch

from oracle import kmeans_oracle as ko
from speech_resynth_b200 import synthetic

pytestmark = pytest.mark.gpu


def test_kmeans_assign_matches_sklearn_golden(golden_dir):
    from speech_resynth_b200.units import UnitQuantizer

    z = np.load(os.path.join(golden_dir, "kmeans_k300_d64.npz"))
    q = UnitQuantizer(torch.from_numpy(z["centroids"]), device="cuda")
    labels = q.predict(torch.from_numpy(z["feats"]).cuda()).cpu().numpy()
    assert labels.dtype == np.int64 and np.array_equal(labels, z["labels"])          # K = 300 -> 512 padded columns never win
    lengths = torch.from_numpy(z["lengths"])
    ids = q.encode(torch.from_numpy(z["feats"]).cuda(), lengths).cpu().numpy()
    ref = ko.encode(z["feats"], z["lengths"].tolist(), z["centroids"], deduplicate=False)
    assert np.array_equal(ids, ref)
    out, counts, n_out = (t.cpu().numpy() for t in q.encode(torch.from_numpy(z["feats"]).cuda(), lengths, deduplicate=True))
    r_out, r_counts, r_n = ko.encode(z["feats"], z["lengths"].tolist(), z["centroids"], deduplicate=True)
    assert np.array_equal(out, r_out) and np.array_equal(counts, r_counts) and np.array_equal(n_out, r_n)


def test_decoder_codebook_round_trip_and_full_size(state_dict):
    """The decoder's embedding table is the codebook (utils/textless.py:33-35): quantising its own rows returns their
    indices, and at the real size -- 2000 x 768, a 64 x 500 frame batch -- noisy rows agree with the float64 oracle on
    every decisive frame.  Then units -> decoder: encode() output is a valid decoder input."""
    import speech_resynth_b200 as srb
    from speech_resynth_b200.units import UnitQuantizer

    m = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    m.load_state_dict(state_dict, strict=True)
    m = m.cuda()
    q = UnitQuantizer.from_decoder(m)
    table = m.model.to_cond_emb.weight.detach()
    assert (q.n_clusters, q.dim) == (2000, 768)
    assert torch.equal(q.predict(table[1:]), torch.arange(2000, device="cuda"))
    gen = torch.Generator().manual_seed(2)
    units = torch.randint(0, 2000, (64, 500), generator=gen)
    cents = table[1:].cpu()
    feats = cents[units] + 0.35 * cents.std() * torch.randn(64, 500, 768, generator=gen)
    got = q.predict(feats.cuda()).cpu().numpy()
    sub = slice(0, 4)                                                   # the float64 oracle on 2000 frames
    ref = ko.assign(feats[sub].numpy(), cents.numpy())
    decisive = ko.margins(feats[sub].numpy(), cents.numpy()).reshape(ref.shape) > 1e-5
    assert decisive.mean() > 0.999 and np.array_equal(got[sub][decisive], ref[decisive])
    # every frame: the chosen centroid is (within fp32 noise) as close as any other -- checked through the distance itself
    x = feats.cuda()
    d_got = (x - table[1:][torch.from_numpy(got).cuda()]).pow(2).sum(-1)
    d_true = (x - table[1:][units.cuda()]).pow(2).sum(-1)
    assert bool((d_got <= d_true * (1 + 1e-5)).all())
    lengths = torch.tensor([500] * 60 + [123, 7, 1, 64])
    ids = q.encode(x, lengths)
    assert bool((ids.cpu() == 0).eq(torch.arange(500)[None, :] >= lengths[:, None]).all()) and int(ids.max()) <= 2000
    wavs = m(ids, 0.25, 1.0)
    assert [w.shape[-1] for w in wavs] == [320 * n + 80 for n in lengths.tolist()]


def test_ragged_shapes_and_ties():
    """Row counts that are not tile multiples, a width that needs K padding (24 -> 72 split columns -> 128), exact ties
    (duplicate centroids: the smaller index wins, numpy / sklearn argmin), single rows; de-duplication without lengths."""
    from speech_resynth_b200 import _native as nat
    from speech_resynth_b200.units import UnitQuantizer

    gen = torch.Generator().manual_seed(9)
    cents = torch.randn(37, 24, generator=gen)
    cents[20] = cents[5]                                                # duplicate: index 5 must win
    q = UnitQuantizer(cents, device="cuda")
    for rows in (1, 127, 129, 1000):
        x = torch.randn(rows, 24, generator=gen)
        x[0] = cents[5]
        got = q.predict(x.cuda()).cpu().numpy()
        ref = ko.assign(x.numpy(), cents.numpy())
        decisive = ko.margins(x.numpy(), np.delete(cents.numpy(), 20, axis=0)) > 1e-5
        assert got[0] == 5 and np.array_equal(got[decisive], ref[decisive]) and 20 not in got
    ids = torch.tensor([[3, 3, 3, 7, 7, 1, 3, 3], [5, 5, 5, 5, 5, 5, 5, 5]], dtype=torch.int64).cuda()
    out, counts, n_out = torch.empty_like(ids), torch.empty(2, 8, dtype=torch.int32, device="cuda"), torch.empty(2, dtype=torch.int32, device="cuda")
    nat.call("srb_unique_consecutive", nat.ptr(ids), None, nat.ptr(out), nat.ptr(counts), nat.ptr(n_out), 2, 8)
    assert out.cpu().tolist() == [[3, 7, 1, 3, 0, 0, 0, 0], [5, 0, 0, 0, 0, 0, 0, 0]]
    assert counts.cpu().tolist() == [[3, 2, 1, 2, 0, 0, 0, 0], [8, 0, 0, 0, 0, 0, 0, 0]] and n_out.cpu().tolist() == [4, 1]
    long_ids = torch.randint(1, 4, (3, 1000), generator=gen)           # runs that cross the 256-thread chunks
    lens = torch.tensor([1000, 513, 256], dtype=torch.int32)
    o2, c2, n2 = torch.empty(3, 1000, dtype=torch.int64, device="cuda"), torch.empty(3, 1000, dtype=torch.int32, device="cuda"), torch.empty(3, dtype=torch.int32, device="cuda")
    nat.call("srb_unique_consecutive", nat.ptr(long_ids.cuda()), nat.ptr(lens.cuda()), nat.ptr(o2), nat.ptr(c2), nat.ptr(n2), 3, 1000)
    for b in range(3):
        u, c = torch.unique_consecutive(long_ids[b, : int(lens[b])], return_counts=True)
        assert int(n2[b]) == len(u) and torch.equal(o2[b, : len(u)].cpu(), u) and torch.equal(c2[b, : len(u)].cpu().long(), c)
        assert not bool(o2[b, len(u):].any())
