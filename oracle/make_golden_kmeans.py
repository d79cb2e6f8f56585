"""Mint tests/golden/kmeans_k300_d64.npz from scikit-learn's own KMeans.predict (the call textlesslib's quantiser makes,
src/flow_matching/utils/textless.py:9-21).  Run in the build container: python -m oracle.make_golden_kmeans"""
import os

import numpy as np
from sklearn.cluster import KMeans

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rng = np.random.default_rng(5)
    k, d = 300, 64
    # clustered training data so that the fitted centres are well separated, like a real codebook
    centres = rng.normal(size=(k, d)).astype(np.float32) * 2.0
    train = (centres[rng.integers(0, k, size=6000)] + 0.3 * rng.normal(size=(6000, d))).astype(np.float32)
    km = KMeans(n_clusters=k, n_init=1, max_iter=5, random_state=0).fit(train)
    feats = (km.cluster_centers_[rng.integers(0, k, size=(3, 64))] + 0.5 * rng.normal(size=(3, 64, d))).astype(np.float32)
    # runs of repeated frames (speech units repeat): every frame copied 1-3 times
    reps = rng.integers(1, 4, size=64)
    feats = np.repeat(feats, reps, axis=1)[:, :96]
    assert feats.shape[1] == 96
    labels = km.predict(feats.reshape(-1, d)).reshape(feats.shape[:2]).astype(np.int64)
    out = os.path.join(ROOT, "tests", "golden", "kmeans_k300_d64.npz")
    np.savez_compressed(out, centroids=km.cluster_centers_.astype(np.float32), feats=feats, labels=labels,
                        lengths=np.array([96, 61, 7], dtype=np.int32))
    print("wrote", out, os.path.getsize(out), "bytes; sklearn", __import__("sklearn").__version__)


if __name__ == "__main__":
    main()
