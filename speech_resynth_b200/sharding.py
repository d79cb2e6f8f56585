"""Utterance-level sharding across the GPUs of one box.

Utterances are independent (pads are masked everywhere inside the ODE and never reach a valid frame), so the path
shards with NO collective inside the ODE loop or the vocoder: sort by length, cut into buckets of equal padded
length, deal whole buckets to ranks balancing the FLOP cost model, run each bucket through the local decoder and
gather the cropped waveforms on rank 0 once at the end (lengths first, then one padded float buffer per rank).

Host-side logic only: it works with any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch

SAMPLE_RATE = 16000


@dataclass
class Bucket:
    indices: List[int]      # positions in the caller's utterance list
    frames: int             # padded length = longest member (what pad_sequence gives the reference, synthesize.py:42)

    @property
    def batch(self) -> int:
        return len(self.indices)


# ---- cost model: algorithmic FLOPs of the path (multiply-add = 2) --------------------------
# generator geometry of the shipped vocoder config (HF config defaults + src/hifigan/train.py:36-42)
_UP_RATES, _UP_KERNELS = (5, 4, 4, 2, 2), (10, 9, 8, 4, 4)
_PER_FRAME_LINEAR = 19_103_232      # embed + 4 x (qkv, out, both FFN convs) + to_pred, per frame per velocity evaluation
_HOISTED_COND = 393_216             # the loop-invariant conditioning half of to_embed, computed once per call here


def transformer_flops(frames: int, nfe: int, hoisted: bool = True) -> int:
    """ODE loop over the transformer for one utterance padded to `frames`: NFE N (19 103 232 + 4096 N), attention being
    the 4096 N term (4 layers x 2 heads x 128 x 4 N); `hoisted` drops the conditioning projection this path runs once."""
    per_frame = _PER_FRAME_LINEAR + 4096 * frames - (_HOISTED_COND if hoisted else 0)
    return nfe * frames * per_frame


def vocoder_flops(frames: int) -> int:
    """HiFi-GAN generator on `frames` mel frames: conv_pre, five (transposed conv + three-resblock MRF) stages, conv_post."""
    total = 2 * 7 * 80 * 512 * frames
    channels, rows = 512, frames
    for rate, kernel in zip(_UP_RATES, _UP_KERNELS):
        out_rows = (rows - 1) * rate - 2 * ((kernel - rate) // 2) + kernel
        total += 2 * channels * (channels // 2) * kernel * rows          # transposed conv, per input row
        total += 2 * (3 + 7 + 11) * 6 * (channels // 2) ** 2 * out_rows   # 18 convs of the three resblocks
        channels, rows = channels // 2, out_rows
    return total + 2 * 16 * 7 * rows


def utterance_cost(frames: int, nfe: int) -> float:
    """Algorithmic FLOPs of one utterance padded to `frames`: what the rank assignment balances."""
    return float(transformer_flops(frames, nfe) + vocoder_flops(frames))


def bucket_by_length(lengths: Sequence[int], granularity: int = 64, max_batch: int = 64,
                     max_frames_per_bucket: int = 64 * 1024) -> List[Bucket]:
    """Sort by length (descending) and cut into buckets of utterances whose lengths round up to the same multiple of
    `granularity`.  A bucket is padded to its longest member only, exactly like the reference's pad_sequence batch,
    so "the vocoder sees the padded batch" semantics are those of the reference."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    buckets: List[Bucket] = []
    cur: Optional[Bucket] = None
    cur_class = -1
    for i in order:
        n = int(lengths[i])
        if n <= 0:
            raise ValueError("every utterance needs at least one unit (the reference yields NaN for empty rows)")
        klass = (n + granularity - 1) // granularity
        if cur is None or klass != cur_class or cur.batch >= max_batch or (cur.batch + 1) * cur.frames > max_frames_per_bucket:
            cur, cur_class = Bucket([], n), klass   # descending order: the first member is the longest
            buckets.append(cur)
        cur.indices.append(i)
    return buckets


def assign_buckets(buckets: Sequence[Bucket], world: int, nfe: int = 16) -> List[List[int]]:
    """Greedy longest-processing-time assignment of whole buckets to ranks; returns bucket ids per rank."""
    cost = [b.batch * utterance_cost(b.frames, nfe) for b in buckets]
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for j in sorted(range(len(buckets)), key=lambda j: -cost[j]):
        r = min(range(world), key=lambda r: (load[r], r))
        out[r].append(j)
        load[r] += cost[j]
    for r in range(world):
        out[r].sort()
    return out


def pad_bucket(units: Sequence[torch.Tensor], bucket: Bucket) -> torch.Tensor:
    """Right-pad with 0 (ids are unit + 1, 0 = pad; synthesize.py:39-42) to the bucket's padded length."""
    ids = torch.zeros(bucket.batch, bucket.frames, dtype=torch.int64)
    for row, i in enumerate(bucket.indices):
        u = units[i].reshape(-1).to(torch.int64)
        ids[row, : u.numel()] = u
    return ids


def resynthesize_sharded(units: Sequence[torch.Tensor], synth: Callable[[torch.Tensor], List[torch.Tensor]],
                         rank: int = 0, world: int = 1, nfe: int = 16, granularity: int = 64, max_batch: int = 64,
                         device: Optional[torch.device] = None, group=None) -> Optional[List[torch.Tensor]]:
    """Run `synth(ids (B, N) int64) -> list of (1, L_i) waveforms` on this rank's buckets and gather on rank 0.

    Every rank passes the same `units` list (length-only metadata is enough to agree on the plan).  Returns the
    waveforms in the caller's order on rank 0 and None elsewhere.
    """
    import torch.distributed as dist

    lengths = [int(u.numel()) for u in units]
    buckets = bucket_by_length(lengths, granularity, max_batch)
    plan = assign_buckets(buckets, world, nfe)
    mine = plan[rank]
    wav_len = [320 * n + 80 for n in lengths]

    local_idx: List[int] = []
    local_wavs: List[torch.Tensor] = []
    for j in mine:
        b = buckets[j]
        ids = pad_bucket(units, b)
        if device is not None:
            ids = ids.to(device)
        outs = synth(ids)
        for i, w in zip(b.indices, outs):
            assert w.shape[-1] == wav_len[i], (w.shape, wav_len[i])
            local_idx.append(i)
            local_wavs.append(w.reshape(-1))
    if world == 1:
        out: List[Optional[torch.Tensor]] = [None] * len(units)
        for i, w in zip(local_idx, local_wavs):
            out[i] = w.unsqueeze(0)
        return out  # type: ignore[return-value]

    # final gather: every rank knows every rank's utterance list from the shared plan, so only samples move
    per_rank_idx = [[i for j in plan[r] for i in buckets[j].indices] for r in range(world)]
    sizes = [sum(wav_len[i] for i in idx) for idx in per_rank_idx]
    cap = max(max(sizes), 1)
    dev = local_wavs[0].device if local_wavs else (device or torch.device("cpu"))
    flat = torch.zeros(cap, dtype=torch.float32, device=dev)
    if local_wavs:
        torch.cat(local_wavs, out=flat[: sizes[rank]])
    gathered = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(flat, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out = [None] * len(units)
    for r in range(world):
        off = 0
        for i in per_rank_idx[r]:
            out[i] = gathered[r][off: off + wav_len[i]].unsqueeze(0)
            off += wav_len[i]
    return out  # type: ignore[return-value]
