"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the unit quantiser on the input side of the path.

The reference quantises dense features with textlesslib's ``SpeechEncoder`` (src/flow_matching/utils/textless.py:9-21):
``units = kmeans_model.predict(features)`` on a scikit-learn ``KMeans`` (joblib checkpoint), optionally followed by
``torch.unique_consecutive(units, return_counts=True)`` (``deduplicate=True``); callers add 1 (README.md:43, 0 = pad).
textlesslib / fairseq are not installed here; scikit-learn IS, and ``KMeans.predict`` is the algorithm itself, so the
restatement below (float64 arg-min of squared Euclidean distances, first minimum on ties) is pinned against sklearn's
own ``predict`` on the same codebook (tests/test_oracle_cpu.py, tests/golden/kmeans_k300_d64.npz minted by
oracle/make_golden_kmeans.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def assign(features: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """argmin_j |x - c_j|^2 in float64 (KMeans.predict / _labels_inertia: squared Euclidean distance, argmin)."""
    x = np.asarray(features, dtype=np.float64)
    c = np.asarray(centroids, dtype=np.float64)
    flat = x.reshape(-1, x.shape[-1])
    out = np.empty(flat.shape[0], dtype=np.int64)
    c2 = (c * c).sum(axis=1)
    for s in range(0, flat.shape[0], 4096):
        blk = flat[s: s + 4096]
        d = (blk * blk).sum(axis=1, keepdims=True) - 2.0 * blk @ c.T + c2[None, :]
        out[s: s + 4096] = d.argmin(axis=1)
    return out.reshape(x.shape[:-1])


def margins(features: np.ndarray, centroids: np.ndarray) -> np.ndarray:
    """Relative gap between the two smallest squared distances of every row (how decisive the assignment is)."""
    x = np.asarray(features, dtype=np.float64).reshape(-1, features.shape[-1])
    c = np.asarray(centroids, dtype=np.float64)
    d = (x * x).sum(axis=1, keepdims=True) - 2.0 * x @ c.T + (c * c).sum(axis=1)[None, :]
    part = np.partition(d, 1, axis=1)
    return (part[:, 1] - part[:, 0]) / np.maximum(part[:, 1], 1e-30)


def unique_consecutive(units: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """torch.unique_consecutive(units, return_counts=True) for one utterance."""
    units = np.asarray(units)
    if units.size == 0:
        return units[:0], np.zeros(0, dtype=np.int64)
    starts = np.flatnonzero(np.concatenate([[True], units[1:] != units[:-1]]))
    counts = np.diff(np.concatenate([starts, [units.size]]))
    return units[starts], counts


def encode(features: np.ndarray, lengths: List[int], centroids: np.ndarray, deduplicate: bool):
    """SpeechEncoder post-processing for a padded (B, T, D) batch: labels + 1, 0 at pads; with `deduplicate` the collapsed ids,
    run lengths and run counts, right-padded with 0."""
    b, t, _ = features.shape
    ids = np.zeros((b, t), dtype=np.int64)
    for i, n in enumerate(lengths):
        ids[i, :n] = assign(features[i, :n], centroids) + 1
    if not deduplicate:
        return ids
    out = np.zeros_like(ids)
    counts = np.zeros((b, t), dtype=np.int32)
    n_out = np.zeros(b, dtype=np.int32)
    for i, n in enumerate(lengths):
        u, c = unique_consecutive(ids[i, :n])
        out[i, : len(u)] = u
        counts[i, : len(u)] = c
        n_out[i] = len(u)
    return out, counts, n_out
