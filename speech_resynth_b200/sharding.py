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


def contiguous_partition(lengths: Sequence[int], world: int, nfe: int = 16,
                         shares: Optional[Sequence[float]] = None) -> List[List[int]]:
    """Sort by length (descending) and cut the sorted list into `world` contiguous ranges whose modelled cost follows
    `shares` (default: equal): rank 0 gets the longest utterances, the last rank the shortest.  Balanced to within one
    utterance, and every rank's utterances are as homogeneous in length as the input allows (little padding inside
    its buckets)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    cost = [utterance_cost(int(lengths[i]), nfe) for i in order]
    total = sum(cost)
    shares = [1.0 / world] * world if shares is None else [s / sum(shares) for s in shares]
    edges, run = [], 0.0
    for s in shares:
        run += s
        edges.append(run * total)
    out: List[List[int]] = [[] for _ in range(world)]
    acc, r = 0.0, 0
    for i, c in zip(order, cost):
        # an utterance goes to the rank whose share contains its midpoint
        while r + 1 < world and acc + 0.5 * c >= edges[r]:
            r += 1
        out[r].append(i)
        acc += c
    return out


def tiles_per_utterance(frames: int) -> int:
    """128-row GEMM tiles an utterance of `frames` (padded to a multiple of 8) occupies: tiles never span utterances."""
    return ((frames + 7) // 8 * 8 + 127) // 128


def _greedy_buckets(order: Sequence[int], lengths: Sequence[int], tile_budget: int, max_batch: int, max_waste: float,
                    min_window: int) -> List[Bucket]:
    buckets: List[Bucket] = []
    cur: Optional[Bucket] = None
    for i in order:
        n = int(lengths[i])
        if n <= 0:
            raise ValueError("every utterance needs at least one unit (the reference yields NaN for empty rows)")
        if cur is not None:
            full = (cur.batch + 1) * tiles_per_utterance(cur.frames) > tile_budget or cur.batch >= max_batch
            if full or cur.frames - n > max(min_window, int(max_waste * cur.frames)):
                cur = None
        if cur is None:
            cur = Bucket([], n)
            buckets.append(cur)
        cur.indices.append(i)
    return buckets


def bucket_sorted(order: Sequence[int], lengths: Sequence[int], tile_budget: int = 296, max_batch: int = 160,
                  max_waste: float = 0.12, min_window: int = 64) -> List[Bucket]:
    """Buckets over utterances already sorted by descending length.  A bucket is padded to its first (longest) member
    and closes when the next utterance would (a) push it past the tile budget (296 row tiles = two full waves of the
    148 SMs for the kernels whose tile spans the whole width), (b) exceed `max_batch`, or (c) be padded by more than
    max(min_window, max_waste * frames) frames.  The budget actually used is the smallest one that needs no more
    buckets than `tile_budget` does: the same number of buckets, but evenly filled instead of full ones plus a
    nearly empty straggler (every bucket costs ~1500 kernel launches however small it is)."""
    order = list(order)
    best = _greedy_buckets(order, lengths, tile_budget, max_batch, max_waste, min_window)
    lo, hi = 1, tile_budget
    while lo < hi:
        mid = (lo + hi) // 2
        trial = _greedy_buckets(order, lengths, mid, max_batch, max_waste, min_window)
        if len(trial) <= len(best):
            best, hi = trial, mid
        else:
            lo = mid + 1
    return best


@dataclass
class ShardPlan:
    buckets: List[Bucket]
    per_rank: List[List[int]]          # bucket ids per rank, in execution order
    cost: List[float]                  # modelled FLOPs per rank (padded buckets)

    @property
    def imbalance(self) -> float:
        """max / mean of the modelled per-rank cost"""
        mean = sum(self.cost) / max(len(self.cost), 1)
        return max(self.cost) / mean if mean > 0 else 1.0


def bucket_cost(bucket: Bucket, nfe: int = 16) -> float:
    """Modelled cost of a bucket: the FLOPs of its PADDED shape (pad frames are computed like any other) plus a fixed
    term for the ~500 kernel launches every bucket pays however small it is.  Measured on a B200
    (tools/calibrate_buckets.py, profiles/r02_bucket_calibration.jsonl): time = 5.0 ms + 1.17 ms / TFLOP from 2 x 687 to
    64 x 500 frames, i.e. the fixed part is worth 4.3 TFLOP of marginal work."""
    return bucket.batch * utterance_cost(bucket.frames, nfe) + BUCKET_FIXED_FLOPS


BUCKET_FIXED_FLOPS = 4.3e12


def plan_shards(lengths: Sequence[int], world: int, nfe: int = 16, strategy: str = "contiguous", **bucket_args) -> ShardPlan:
    """Every rank calls this with the same lengths and gets the same plan.
    strategy "contiguous": cost-balanced contiguous ranges of the length-sorted list, bucketed per rank (`bucket_sorted`);
    the range boundaries are refined for a few rounds against the cost of the buckets they produce (padding and the
    per-bucket fixed cost differ between long and short ranges);
    strategy "lpt": global buckets by length class (`bucket_by_length`) dealt to ranks longest-processing-time first."""
    if strategy == "lpt":
        buckets = bucket_by_length(lengths, **bucket_args)
        per_rank = assign_buckets(buckets, world, nfe)
        cost = [sum(bucket_cost(buckets[j], nfe) for j in ids) for ids in per_rank]
        return ShardPlan(buckets, per_rank, cost)
    if strategy != "contiguous":
        raise ValueError(f"unknown sharding strategy {strategy!r}")
    best: Optional[ShardPlan] = None
    best_parts: List[List[int]] = []
    shares = [1.0] * world
    for _ in range(8 if world > 1 else 1):
        parts = contiguous_partition(lengths, world, nfe, shares)
        plan = _plan_of_parts(parts, lengths, nfe, bucket_args)
        if best is None or plan.imbalance < best.imbalance:
            best, best_parts = plan, parts
        mean = sum(plan.cost) / world
        shares = [s * (mean / c if c > 0 else 1.0) ** 0.7 for s, c in zip(shares, plan.cost)]
    if world > 1:
        best = _refine_boundaries(best, best_parts, lengths, nfe, bucket_args)
    return best


def _plan_of_parts(parts: Sequence[Sequence[int]], lengths: Sequence[int], nfe: int, bucket_args: dict) -> ShardPlan:
    buckets: List[Bucket] = []
    per_rank: List[List[int]] = []
    for part in parts:
        mine = bucket_sorted(part, lengths, **bucket_args) if part else []
        per_rank.append(list(range(len(buckets), len(buckets) + len(mine))))
        buckets.extend(mine)
    cost = [sum(bucket_cost(buckets[j], nfe) for j in ids) for ids in per_rank]
    return ShardPlan(buckets, per_rank, cost)


def _refine_boundaries(plan: ShardPlan, parts: Sequence[Sequence[int]], lengths: Sequence[int], nfe: int,
                       bucket_args: dict, max_passes: int = 12) -> ShardPlan:
    """Local search on the rank boundaries of the length-sorted list.  The share iteration above moves all boundaries at
    once and stalls where a rank's cost jumps (one utterance more opens another bucket = another fixed cost): at 8 ranks on
    BASELINE configs[2] it leaves max / mean = 1.050, and the measured per-rank times follow the model (1.053).  Here single
    boundaries move by +-1 ... +-32 utterances; a move is kept when it lowers the larger of the two ranks it touches (never
    raising the global maximum), until no move helps.  Deterministic, so every rank still derives the same plan."""
    order = [i for part in parts for i in part]
    bounds = [0]
    for part in parts:
        bounds.append(bounds[-1] + len(part))
    world = len(parts)

    def rank_cost(lo: int, hi: int) -> float:
        if hi <= lo:
            return 0.0
        return sum(bucket_cost(b, nfe) for b in bucket_sorted(order[lo:hi], lengths, **bucket_args))

    cost = [rank_cost(bounds[r], bounds[r + 1]) for r in range(world)]
    for _ in range(max_passes):
        improved = False
        for k in range(1, world):                      # boundary between ranks k - 1 and k
            pair_max = max(cost[k - 1], cost[k])
            best_move = None
            for step in (1, 2, 4, 8, 16, 32):
                for delta in (-step, step):
                    nb = bounds[k] + delta
                    if nb <= bounds[k - 1] or nb >= bounds[k + 1]:
                        continue
                    c0, c1 = rank_cost(bounds[k - 1], nb), rank_cost(nb, bounds[k + 1])
                    if max(c0, c1) < pair_max * (1.0 - 1e-9):
                        pair_max, best_move = max(c0, c1), (nb, c0, c1)
            if best_move is not None:
                bounds[k], cost[k - 1], cost[k] = best_move
                improved = True
        if not improved:
            break
    refined = _plan_of_parts([order[bounds[r]: bounds[r + 1]] for r in range(world)], lengths, nfe, bucket_args)
    return refined if refined.imbalance < plan.imbalance else plan


def pad_bucket(units: Sequence[torch.Tensor], bucket: Bucket, pin: bool = False) -> torch.Tensor:
    """Right-pad with 0 (ids are unit + 1, 0 = pad; synthesize.py:39-42) to the bucket's padded length."""
    ids = torch.zeros(bucket.batch, bucket.frames, dtype=torch.int64)
    for row, i in enumerate(bucket.indices):
        u = units[i].reshape(-1).to(torch.int64)
        ids[row, : u.numel()] = u
    return ids.pin_memory() if pin else ids


def resynthesize_sharded(units: Sequence[torch.Tensor], synth: Optional[Callable[[torch.Tensor], List[torch.Tensor]]] = None,
                         rank: int = 0, world: int = 1, nfe: int = 16, device: Optional[torch.device] = None, group=None,
                         plan: Optional[ShardPlan] = None, strategy: str = "contiguous",
                         synth_into: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None,
                         on_plan: Optional[Callable[[ShardPlan], None]] = None, stats: Optional[dict] = None,
                         **bucket_args) -> Optional[List[torch.Tensor]]:
    """Run this rank's buckets through the local decoder and gather the cropped waveforms on rank 0.

    Every rank passes the same `units` list (length-only metadata is enough to agree on the plan).  Either
    `synth(ids (B, N) int64) -> list of (1, L_i) waveforms` (the decoder's public call), or
    `synth_into(ids, out)` which writes the bucket's cropped waveforms back to back into the 1-D view `out` of this
    rank's gather buffer (no per-utterance tensors, no concatenation).  `ids` are host tensors (pinned when a device
    is given): the decoder takes the valid-frame counts from them without waiting for the GPU.
    `on_plan(plan)` runs before the first bucket (a driver sizes its arena there).  `stats`, when given, receives
    `plan`, `buckets_run`, `local_samples` and (CUDA) `local_done`, an event recorded after this rank's last bucket, before the gather.  Returns the waveforms in the caller's order on rank 0, None elsewhere.
    """
    import torch.distributed as dist

    lengths = [int(u.numel()) for u in units]
    if plan is None:
        plan = plan_shards(lengths, world, nfe, strategy=strategy, **bucket_args)
    if on_plan is not None:
        on_plan(plan)
    buckets = plan.buckets
    wav_len = [320 * n + 80 for n in lengths]
    per_rank_idx = [[i for j in plan.per_rank[r] for i in buckets[j].indices] for r in range(world)]
    sizes = [sum(wav_len[i] for i in idx) for idx in per_rank_idx]
    cap = max(max(sizes), 1)
    dev = device if device is not None else torch.device("cpu")
    # this rank's waveforms, back to back in plan order; padded to the largest rank's size for the gather
    flat = torch.zeros(cap if world > 1 else sizes[rank], dtype=torch.float32, device=dev)

    off = 0
    for j in plan.per_rank[rank]:
        b = buckets[j]
        ids = pad_bucket(units, b, pin=dev.type == "cuda")
        n_b = sum(wav_len[i] for i in b.indices)
        if synth_into is not None:
            synth_into(ids, flat[off: off + n_b])
        else:
            outs = synth(ids)
            o = off
            for i, w in zip(b.indices, outs):
                assert w.shape[-1] == wav_len[i], (w.shape, wav_len[i])
                flat[o: o + wav_len[i]].copy_(w.reshape(-1))
                o += wav_len[i]
        off += n_b
    if stats is not None:
        stats.update(plan=plan, buckets_run=len(plan.per_rank[rank]), local_samples=sizes[rank])
        if dev.type == "cuda":
            # this rank's own work ends here (the gather below waits for the slowest rank): lets a driver measure the load balance
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            stats["local_done"] = ev

    if world == 1:
        gathered = [flat]
    else:
        gathered = [torch.empty(cap, dtype=torch.float32, device=dev) for _ in range(world)] if rank == 0 else None
        dist.gather(flat, gathered, dst=0, group=group)
        if rank != 0:
            return None
    out: List[Optional[torch.Tensor]] = [None] * len(units)
    for r in range(world):
        o = 0
        for i in per_rank_idx[r]:
            out[i] = gathered[r][o: o + wav_len[i]].unsqueeze(0)
            o += wav_len[i]
    return out  # type: ignore[return-value]
