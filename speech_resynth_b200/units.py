"""Unit quantiser: the k-means half of the reference's upstream unit encoder, on the GPU.

The reference turns speech into units with textlesslib's ``SpeechEncoder`` (src/flow_matching/utils/textless.py:9-21):
dense mHuBERT features -> ``kmeans_model.predict`` (scikit-learn KMeans, 2000 clusters) -> optional de-duplication
(``torch.unique_consecutive``) -> ``units + 1`` as the decoder's input ids (README.md:43, 0 = pad).  The decoder's own
embedding table IS that codebook: ``to_cond_emb.weight = [0; kmeans.cluster_centers_]`` (utils/textless.py:33-35), so a
quantiser needs nothing beyond the decoder checkpoint.  The dense encoder itself (mHuBERT, fairseq) is outside this
package: features come from the caller.

Kernels: ``srb_kmeans_assign`` (score GEMM on the tensor cores over split bf16 operands with the arg-max in the epilogue;
include/srb.h) and ``srb_unique_consecutive``.  No CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as nat
from .packing import split_operand

P = nat.ptr


def pack_centroids(centroids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(K, D) fp32 cluster centres -> (packed (K_pad, k_pad) bf16 = [Ch | Ch | Cl] per row with k_pad = 3 D rounded up to 64 and
    K_pad = K rounded up to 256, zero filled;  neg_half_norm2 (K_pad,) fp32 = -|c|^2 / 2 computed in fp64, -inf for padding rows)."""
    c = centroids.detach().float()
    k, d = c.shape
    if d % 8 != 0:
        raise ValueError("the feature width must be a multiple of 8")
    k_pad = (3 * d + 63) // 64 * 64
    n_pad = (k + 255) // 256 * 256
    packed = torch.zeros(n_pad, k_pad, dtype=torch.float32, device=c.device)
    packed[:k, : 3 * d] = split_operand(c, 1)
    bias = torch.full((n_pad,), float("-inf"), dtype=torch.float32, device=c.device)
    bias[:k] = (-0.5 * c.double().pow(2).sum(dim=1)).float()
    return packed.to(torch.bfloat16).contiguous(), bias.contiguous()


class UnitQuantizer:
    """Nearest-centroid unit assignment (sklearn ``KMeans.predict`` semantics: squared Euclidean distance, ties to the
    smallest index) and the ``SpeechEncoder`` post-processing (de-duplication, +1 id offset, right padding)."""

    def __init__(self, centroids: torch.Tensor, device=None):
        nat.require_blackwell()
        device = torch.device(device if device is not None else (centroids.device if centroids.is_cuda else "cuda"))
        self.device = device
        self.n_clusters, self.dim = centroids.shape
        with torch.cuda.device(device):
            self.packed, self.bias = pack_centroids(centroids.to(device))

    @classmethod
    def from_decoder(cls, decoder) -> "UnitQuantizer":
        """Codebook = the decoder's embedding table without its pad row (utils/textless.py:33-35)."""
        model = getattr(decoder, "model", decoder)
        table = model.to_cond_emb.weight.detach()
        return cls(table[1:], device=table.device)

    def _assign(self, feats: torch.Tensor, id_offset: int, lengths: Optional[torch.Tensor], frames: int) -> torch.Tensor:
        x = feats.to(device=self.device, dtype=torch.float32).contiguous()
        if x.shape[-1] != self.dim:
            raise ValueError(f"features are {x.shape[-1]} wide, the codebook {self.dim}")
        rows = x.numel() // self.dim
        units = torch.empty(x.shape[:-1], dtype=torch.int64, device=self.device)
        if rows == 0:
            return units
        with torch.cuda.device(self.device):
            split_ws = torch.empty(rows, 3 * self.dim, dtype=torch.bfloat16, device=self.device)
            keys = torch.empty(rows, dtype=torch.int64, device=self.device)
            nat.call("srb_kmeans_assign", P(x), P(self.packed), P(self.bias), P(split_ws), P(keys), P(units), rows, self.dim,
                     self.packed.shape[0], id_offset, P(lengths) if lengths is not None else None, frames,
                     flops=2.0 * rows * self.packed.shape[0] * 3 * self.dim)
        return units

    @torch.inference_mode()
    def predict(self, feats: torch.Tensor) -> torch.Tensor:
        """(..., D) features -> (...) int64 cluster labels in [0, K): ``kmeans_model.predict``."""
        return self._assign(feats, 0, None, 1)

    @torch.inference_mode()
    def encode(self, feats: torch.Tensor, lengths: Optional[torch.Tensor] = None, deduplicate: bool = False):
        """(B, T, D) features (+ valid frame counts) -> the decoder's input ids (B, T) int64: label + 1, 0 at pads.
        deduplicate=True (the duration-prediction checkpoints): returns (ids, durations (B, T) int32, lengths (B,) int32) with
        every run of equal units collapsed to one id + its run length (``torch.unique_consecutive(return_counts=True)`` per
        utterance), right-padded with 0."""
        if feats.dim() != 3:
            raise ValueError("encode takes (batch, frames, dim) features")
        b, t, _ = feats.shape
        lens = None
        if lengths is not None:
            lens = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        ids = self._assign(feats, 1, lens, t)
        if not deduplicate:
            return ids
        out = torch.empty_like(ids)
        counts = torch.empty(b, t, dtype=torch.int32, device=self.device)
        n_out = torch.empty(b, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            nat.call("srb_unique_consecutive", P(ids), P(lens) if lens is not None else None, P(out), P(counts), P(n_out), b, t)
        return out, counts, n_out
