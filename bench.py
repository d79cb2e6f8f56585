"""Headline benchmark: seconds of audio synthesised per wall-second (RTF^-1) of the unit-to-speech path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager-gpu] [--config 2|4|5]

Default (--config 2): one "step" = one decoder(units) call over one synthetic batch: BASELINE.json configs[1] -- 64
utterances x 500 units (10 s each), dt = 0.0625 (NFE 16, the reference's step count), truncation 1.0, bf16 tensor-core
compute.  With N > 1 (torchrun, one process per GPU) every rank synthesises its own batch of that size (weak scaling,
no collective inside the ODE loop or the vocoder) and the waveforms are gathered on rank 0 at the end of each step.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (units already in HBM); `e2e` = through the
public API from pinned host memory with the waveforms copied back to the host; `roofline` = the dominant kernel's
achieved bf16 TFLOP/s (algorithmic FLOPs / CUDA-event time) against MEASURED_PEAKS.json, with the whole step beside
it; `hbm_rooflines` = the CUDA-core kernels' achieved GB/s; `cpu_baseline` = the reference's CPU path timed on this
box's host cores on a bounded sample; `config3` = BASELINE.json configs[2] (1024 ragged utterances, 2-20 s) through
the product's sharded driver at this N (strong scaling: the same 1024 utterances at every N).

--config 4: HiFi-GAN alone, 256 x 10 s mel -> waveform (vocoder conv roofline, per-stage table).
--config 5: ODE step sweep NFE 1/4/8/16/32 on 16 x 60 s utterances (attention-heavy).
--impl reference: the reference's own CPU implementation alone (live classes from baseline/_ref when installed,
else the oracle port), same metric / unit / config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "audio_seconds_per_second"
UNIT = "s_audio/s"
BATCH, FRAMES, DT, TRUNC = 64, 500, 0.0625, 1.0
SAMPLE_RATE = 16000
C3_UTTS, C3_SEED, C3_LO, C3_HI = 1024, 11, 100, 1000


def audio_seconds(lengths) -> float:
    return sum(320 * int(n) + 80 for n in lengths) / SAMPLE_RATE


def workload_config(n_gpus: int) -> dict:
    return {
        "workload": "configs[1]: ConditionalFlowMatchingWithHifiGan mhubert-expresso-2000, random-init, 64 x 500 units "
                    "(10 s) per GPU, dt=0.0625 (NFE 16), truncation 1.0",
        "batch_per_gpu": BATCH, "frames": FRAMES, "nfe": 16, "global_batch": BATCH * n_gpus,
        "parallelism": f"utterance-sharded x{n_gpus}, final gather only",
        "l2_policy": "per-step working set (~4.5 GB of activations) exceeds the 126 MB L2; no explicit flush",
    }


# ------------------------------------------------------------------------------------------------ CPU arm
REF_SAMPLE_BATCH = 8


def cpu_reference_run(steps: int, warmup: int, n_gpus: int, as_arm: bool, batch: int = REF_SAMPLE_BATCH):
    """The reference's CPU path on the host cores, bounded sample: `batch` utterances x 500 units per step (the
    per-utterance workload of configs[1]; the reference's cost is linear in the batch).  Runs the LIVE reference classes
    from baseline/_ref (installed by oracle/install_reference.py; stock code path, oneDNN on) when present, else the
    oracle port."""
    from speech_resynth_b200 import synthetic

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_state_dict(0)
    ids = synthetic.make_units(batch, FRAMES, seed=7)
    secs = audio_seconds([FRAMES] * batch)
    kind, how = "port", "fp32 torch CPU ops (oracle port of the reference)"
    run = None
    try:
        from oracle import ref_loader

        if ref_loader.reference_root() is not None:
            model = ref_loader.build_reference_model(sd)

            def run():
                torch.manual_seed(1)
                return model(ids, DT, TRUNC)

            kind = "reference"
            how = ("live reference classes (unmodified src/flow_matching/models.py from baseline/_ref), fp32 eager, stock code "
                   f"path, oneDNN {'on' if torch.backends.mkldnn.is_available() and torch.backends.mkldnn.enabled else 'off'}")
    except Exception as e:  # noqa: BLE001 - fall back to the port, say why
        how += f" [live reference unavailable: {type(e).__name__}: {str(e)[:80]}]"
        run = None
    if run is None:
        from oracle import cfm_hifigan_oracle as oracle

        x0 = torch.randn(batch, FRAMES, 80, generator=torch.Generator().manual_seed(1))

        def run():
            return oracle.resynthesize(sd, ids, x0, DT, TRUNC)

    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            run()
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
    total = sum(times)
    value = secs * len(times) / total
    sample = f"{batch} x {FRAMES} units per step (1/{BATCH // batch} of the configs[1] batch), NFE 16, {how}, {len(times)} steps"
    base = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample}
    if not as_arm:
        return base
    cfg = workload_config(n_gpus)
    cfg["reference_sample"] = sample
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": base, "gpu_launches": 0,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ------------------------------------------------------------------------------------------------ eager-PyTorch bar
def eager_gpu_run(steps: int, warmup: int):
    """The "existing Blackwell library kernels" bar of SURVEY.md section 8(d): the same algorithm as plain eager PyTorch
    ops on the B200 (cuBLAS / cuDNN / SDPA kernels through the oracle port -- the live reference cannot travel to the GPU
    box), config 2, in fp32 and under bf16 autocast (the reference's own low-precision mode, train.py:174).  Reported
    beside the product numbers in profiles/; not part of the default bench line."""
    from oracle import cfm_hifigan_oracle as oracle
    from speech_resynth_b200 import synthetic

    dev = torch.device("cuda", 0)
    sd = {k: v.to(dev) for k, v in synthetic.make_state_dict(0).items()}
    ids = synthetic.make_units(BATCH, FRAMES, seed=7).to(dev)
    x0 = torch.randn(BATCH, FRAMES, 80, device=dev)
    secs = audio_seconds([FRAMES] * BATCH)
    out = {"impl": "eager-pytorch-on-gpu", "metric": METRIC, "unit": UNIT, "config": workload_config(1), "modes": {}}
    for mode in ("fp32", "fp32_tf32", "bf16_autocast"):
        torch.backends.cuda.matmul.allow_tf32 = mode == "fp32_tf32"
        torch.backends.cudnn.allow_tf32 = mode == "fp32_tf32"
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16_autocast" else torch.autocast("cuda", enabled=False)
        with torch.inference_mode(), ctx:
            for _ in range(warmup):
                oracle.resynthesize(sd, ids, x0, DT, TRUNC)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                oracle.resynthesize(sd, ids, x0, DT, TRUNC)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["modes"][mode] = {"ms_per_step": ms, "value": secs / (ms / 1e3)}
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(float(r[1])) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [int(float(r[2])) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"], "src": "measured"}
    return {"bf16": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "src": "fallback"}


def profile_body(engine, plan, setup):
    """One un-captured pass of a plan's kernels with CUDA events around every launch: per-op device time, algorithmic
    FLOPs / bytes.  `setup()` stages the inputs."""
    from speech_resynth_b200 import _native as nat

    # serialise the parallel graph branches for this pass so every kernel is timed alone
    if engine.sampler is not None:
        engine.sampler.fork.enabled = False
    if engine.vocoder is not None:
        engine.vocoder.fork.enabled = False
    setup()
    torch.cuda.synchronize()
    nat.profile_log = []
    # keep the GPU busy while the host enqueues the whole pass: an eager launch costs the host 20-30 us, more than
    # many of these kernels run, and an idle GPU would add that wait to the event-to-event time of the next kernel
    torch.cuda._sleep(int(60e-3 * 1.9e9))
    plan.body()
    if engine.vocoder is not None and plan.x_last is not None:
        engine.vocoder.post(plan.x_last, plan.voc_ws["wav"])
    torch.cuda.synchronize()
    log, nat.profile_log = nat.profile_log, None
    if engine.sampler is not None:
        engine.sampler.fork.enabled = True
    if engine.vocoder is not None:
        engine.vocoder.fork.enabled = True
    agg = {}
    for name, tag, e0, e1, flops, nbytes in log:
        key = (name, tag + (int(flops),))   # launches of one op with different kernel sizes are different rows
        a = agg.setdefault(key, {"ms": 0.0, "n": 0, "flops": flops, "bytes": nbytes})
        a["ms"] += e0.elapsed_time(e1)
        a["n"] += 1
    return agg


def op_profile(engine, ids, x0):
    plan = engine._plan(ids.shape[0], ids.shape[1], DT, True, True)
    return profile_body(engine, plan, lambda: engine.sampler.stage(plan.cfm_ws, ids, x0, TRUNC))


def write_ops(path, agg):
    total_ms = sum(a["ms"] for a in agg.values())
    rowsout = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
    with open(path, "w") as f:
        f.write("op,tag,launches,total_ms,avg_ms,share,tflops,gbs\n")
        for (n_, t_), a in rowsout:
            avg_s = a["ms"] / a["n"] / 1e3 if a["ms"] > 0 else 0
            tf = a["flops"] / avg_s / 1e12 if avg_s else 0
            gb = a["bytes"] / avg_s / 1e9 if avg_s and a["bytes"] else 0
            f.write(f"{n_},\"{list(t_)}\",{a['n']},{a['ms']:.4f},{a['ms'] / a['n']:.4f},{a['ms'] / total_ms:.4f},{tf:.1f},{gb:.0f}\n")


def hbm_rooflines(agg, pk):
    """The CUDA-core / small-N kernels of the step against the measured HBM copy bandwidth: algorithmic bytes (stated per
    op in engine.py) / CUDA-event time."""
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath))
    out = []
    for (name, tag), a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if not a["bytes"]:
            continue
        avg_s = a["ms"] / a["n"] / 1e3
        gbs = a["bytes"] / avg_s / 1e9
        out.append({"bound": "hbm", "kernel": f"{name}{list(tag[:-1])}", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": gbs / pk["hbm"], "algorithmic_bytes_per_launch": a["bytes"], "avg_launch_ms": a["ms"] / a["n"],
                    "launches_per_step": a["n"], "traffic": traffic.get(f"{name}{list(tag[:-1])}")})
    return out


def make_config3_units():
    gen = torch.Generator().manual_seed(C3_SEED)
    lengths = torch.randint(C3_LO, C3_HI + 1, (C3_UTTS,), generator=gen).tolist()
    units = [torch.randint(1, 2001, (n,), generator=gen) for n in lengths]
    return lengths, units


def config3_run(decoder, dev, rank, world, passes: int = 4, bucket_args=None):
    """BASELINE configs[2] through the product API: sharding.resynthesize_sharded over the same 1024 utterances (seed 11,
    100-1000 frames, NFE 16) at every world size -- STRONG scaling.  Pass 1 is cold (every bucket shape runs eagerly),
    pass 2 captures the graphs, later passes replay them; the headline is the best later pass, the others are listed."""
    import torch.distributed as dist

    from speech_resynth_b200 import sharding

    lengths, units = make_config3_units()
    eng = decoder.engine()
    plan = sharding.plan_shards(lengths, world, nfe=16, **(bucket_args or {}))
    mine = [plan.buckets[j] for j in plan.per_rank[rank]]
    big = max(mine, key=lambda b: eng.workspace_bytes(b.batch, b.frames))
    eng.reserve(big.batch, big.frames)

    def synth_into(ids, out):
        decoder.resynthesize_flat(ids, DT, TRUNC, out=out)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    rec = []
    for p in range(passes):
        s0 = dict(eng.stats)
        sync_all()
        e0, e1, el = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0 = time.perf_counter()
        e0.record()
        stats = {}

        def on_plan(_):
            pass

        outs = sharding.resynthesize_sharded(units, None, rank=rank, world=world, nfe=16, device=dev, plan=plan,
                                             synth_into=synth_into, stats=stats)
        e1.record()
        torch.cuda.synchronize()
        host_ms = (time.perf_counter() - t0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), host_ms, e0.elapsed_time(stats["local_done"])], device=dev)
        allms = [torch.zeros_like(ms) for _ in range(world)] if world > 1 else [ms]
        if world > 1:
            dist.all_gather(allms, ms)
        dev_ms = [float(m[0]) for m in allms]
        rec.append({"ms": max(dev_ms), "local_ms_per_rank": [float(m[2]) for m in allms], "host_ms_per_rank": [float(m[1]) for m in allms],
                    "graphs_captured": eng.stats["graphs_captured"] - s0["graphs_captured"],
                    "eager_runs": eng.stats["eager_runs"] - s0["eager_runs"],
                    "graph_replays": eng.stats["graph_replays"] - s0["graph_replays"]})
        if rank == 0 and p == passes - 1:
            assert [o.shape[-1] for o in outs] == [320 * n + 80 for n in lengths]
            assert all(bool(torch.isfinite(o).all()) for o in outs[:: 37])
        del outs
    secs = audio_seconds(lengths)
    steady = min(r["ms"] for r in rec[2:]) if len(rec) > 2 else rec[-1]["ms"]
    best = min(rec[2:], key=lambda r: r["ms"]) if len(rec) > 2 else rec[-1]
    mean_t = sum(best["local_ms_per_rank"]) / world
    flops = sum(sharding.transformer_flops(n, 16) + sharding.vocoder_flops(n) for n in lengths)
    padded_flops = sum(b.batch * (sharding.transformer_flops(b.frames, 16) + sharding.vocoder_flops(b.frames)) for b in plan.buckets)
    return {
        "workload": f"configs[2]: {C3_UTTS} utterances, lengths randint({C3_LO}, {C3_HI + 1}) frames (2-20 s, seed {C3_SEED}), "
                    "NFE 16, truncation 1.0, length-bucketed and sharded by sharding.resynthesize_sharded through "
                    "decoder.resynthesize_flat (host ids in, cropped waveforms gathered on rank 0); strong scaling",
        "n_gpus": world, "audio_seconds": secs, "value": secs / (steady / 1e3), "unit": UNIT, "ms": steady,
        "cold_value": secs / (rec[0]["ms"] / 1e3), "passes": rec,
        "bucket_args": bucket_args or "defaults of sharding.bucket_sorted",
        "buckets": len(plan.buckets), "buckets_per_rank": [len(r) for r in plan.per_rank],
        "bucket_shapes_rank0": [[b.batch, b.frames] for b in mine][:12],
        "load": {"model_max_over_mean": plan.imbalance, "measured_max_over_mean": max(best["local_ms_per_rank"]) / mean_t},
        "algorithmic_tflop": flops / 1e12, "padded_tflop": padded_flops / 1e12,
        "achieved_tflops_per_gpu": flops / (steady / 1e3) / 1e12 / world,
        "arena_gb": eng.arena.nbytes / 2 ** 30,
    }


def gpu_run(args):
    import torch.distributed as dist

    import speech_resynth_b200 as srb
    from speech_resynth_b200 import _native as nat
    from speech_resynth_b200 import sharding, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.to(dev)
    if args.precision == "tight":
        # tight-precision mode (libsrb_tight.so: split bf16 operands, fp32 attention): a verification mode; this line
        # records what it costs at the headline shape (no configs[2] record)
        decoder.set_precision("tight")
        args.no_config3 = True
    engine = decoder.engine()

    if args.config == 4:
        return config4_run(args, decoder, dev) if rank == 0 else None
    if args.config == 5:
        return config5_run(args, decoder, dev) if rank == 0 else None

    ids_host = synthetic.make_units(BATCH, FRAMES, seed=7 + rank).pin_memory()
    ids_dev = ids_host.to(dev)
    rows = 320 * FRAMES + 80
    wav_host = torch.empty(BATCH, rows, dtype=torch.float32).pin_memory()
    gather_buf = [torch.empty(BATCH, rows, dtype=torch.float32, device=dev) for _ in range(world)] if (world > 1 and rank == 0) else None
    secs_per_step = audio_seconds([FRAMES] * BATCH) * world

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def device_step():
        wav, _, _ = engine.resynthesize(ids_dev, DT, TRUNC)
        if world > 1:
            dist.gather(wav, gather_buf, dst=0)

    copy_stream = torch.cuda.Stream(device=dev)
    wav_hosts = [wav_host, torch.empty_like(wav_host).pin_memory()]
    e2e_count = [0]

    def e2e_step():
        """One serving step through the public API: decoder(units) with the units in pinned HOST memory (the H2D copy is
        the decoder's first action), D2H of every waveform.  The read-back runs on a second stream (double-buffered pinned
        host rows) so it overlaps the next step's kernels, as a serving loop would do; all of it is inside the timed
        region (both streams are synchronised at its end)."""
        wavs = decoder(ids_host, DT, TRUNC)                   # public API (list of per-utterance waveforms, fresh storage)
        host = wav_hosts[e2e_count[0] & 1]
        e2e_count[0] += 1
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            for i, w in enumerate(wavs):                      # D2H of the result
                w.record_stream(copy_stream)
                host[i, : w.shape[-1]].copy_(w[0], non_blocking=True)
        if world > 1:
            flat = wavs[0]._base if wavs[0]._base is not None else torch.cat([w.reshape(-1) for w in wavs])
            dist.gather(flat.view(BATCH, rows), gather_buf, dst=0)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = nat.launch_count
        e0.record()
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(copy_stream)   # the e2e read-backs belong to the timed region
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), nat.launch_count - c0

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_dev, launches = timed(device_step, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    # same warm-up as the device loop: the first few public calls still grow the caching allocator (fresh waveform
    # storage per call, held by the copy stream), which is start-up cost, not serving cost
    ms_e2e, _ = timed(e2e_step, args.steps, max(3, args.warmup))

    value = secs_per_step * args.steps / (ms_dev / 1e3)
    e2e_value = secs_per_step * args.steps / (ms_e2e / 1e3)

    agg = None
    if rank == 0:
        x0 = torch.randn(BATCH, FRAMES, 80, device=dev)
        agg = op_profile(engine, ids_dev, x0)
    c3 = None
    if not args.no_config3:
        ba = None
        if args.c3_buckets:
            tb, mb, mw = args.c3_buckets.split(",")
            ba = dict(tile_budget=int(tb), max_batch=int(mb), max_waste=float(mw))
        c3 = config3_run(decoder, dev, rank, world, bucket_args=ba)

    out = None
    if rank == 0:
        pk = peaks()
        total_ms = sum(a["ms"] for a in agg.values())
        (top_name, top_tag), top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        avg_s = top["ms"] / top["n"] / 1e3
        achieved = top["flops"] / avg_s / 1e12
        flops_step = BATCH * (sharding.transformer_flops(FRAMES, 16, hoisted=True) + sharding.vocoder_flops(FRAMES))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            # dram read+write bytes per launch from an ncu --set full capture of exactly this launch shape (else null)
            traffic = json.load(open(tpath)).get(f"{top_name}{list(top_tag[:-1])}")
        step_tflops = flops_step / (ms_dev / args.steps / 1e3) / 1e12
        roofline = {
            "bound": "tensor", "kernel": f"{top_name}{list(top_tag[:-1])}", "achieved": achieved, "peak": pk["bf16_sustained"],
            "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
            "algorithmic_flops_per_launch": top["flops"],
            "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside the step)",
            "launches_per_step": top["n"], "avg_launch_ms": top["ms"] / top["n"], "share_of_step": top["ms"] / total_ms,
            "whole_step": {"algorithmic_tflop": flops_step / 1e12, "achieved_tflops": step_tflops,
                           "frac_of_peak": step_tflops / pk["bf16_sustained"]},
        }
        cpu = cpu_reference_run(steps=2, warmup=1, n_gpus=1, as_arm=False) if world == 1 else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16 hi+lo split operands, fp32 accumulation (tight mode)",
            "data": "synthetic", "config": dict(workload_config(world), precision=args.precision), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8,
                    "d2h_bytes_per_step": BATCH * rows * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "hbm_rooflines": hbm_rooflines(agg, pk), "cpu_baseline": cpu,
            "config3": c3, "engine_stats": dict(engine.stats),
        }
        if args.ops:
            write_ops(args.ops, agg)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------ config 4 / 5
def config4_run(args, decoder, dev):
    """BASELINE configs[3]: HiFi-GAN generator alone (decoder.vocoder(mel), the call at src/flow_matching/train.py:59),
    mel = randn(256, 500, 80) * 2.26 - 5.88 (seed 13) -> 256 x 10.005 s of audio."""
    from speech_resynth_b200 import _native as nat
    from speech_resynth_b200 import sharding

    b, t = 256, FRAMES
    mel = (torch.randn(b, t, 80, generator=torch.Generator().manual_seed(13)) * 2.26 - 5.88).to(dev)
    secs = audio_seconds([t] * b)
    voc = decoder.vocoder
    clocks = ClockSampler(dev.index or 0)
    for _ in range(max(3, args.warmup)):
        voc(mel)
    torch.cuda.synchronize()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = nat.launch_count
    e0.record()
    for _ in range(args.steps):
        voc(mel)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.stop()
    launches = nat.launch_count - c0
    ms = e0.elapsed_time(e1) / args.steps
    eng = voc._engine
    plan = eng._plan(b, t, None, False, True)
    mel32 = mel.float().contiguous()
    agg = profile_body(eng, plan, lambda: nat.call("srb_prior_prepare", nat.ptr(mel32), nat.ptr(plan.voc_ws["mel_b"]), mel32.numel(), 0.0, 0))
    if args.ops:
        write_ops(args.ops, agg)
    pk = peaks()
    flops = b * sharding.vocoder_flops(t)
    # per-stage table: a stage is identified by its channel width (256, 128, 64, 32, 16); an up-sampler is accounted to the
    # stage it produces.  Profile tags hold the integer arguments of the call (engine.py).
    def stage_of(name, tag):
        if name == "srb_hifigan_post":
            return "conv_post"
        if name == "srb_hifigan_upsample":
            return f"stage{(256 // tag[3]).bit_length() - 1}_c{tag[3]}"
        if name == "srb_hifigan_mrf_fused":
            return f"stage{(256 // tag[2]).bit_length() - 1}_c{tag[2]}"
        c_in, c_out = tag[3], tag[4]
        if c_in == 80:
            return "conv_pre"
        c = c_in if c_in == c_out else c_in // 2     # resblock conv / fused tail, else an up-sampler in row-group form
        return f"stage{(256 // c).bit_length() - 1}_c{c}"

    stages = {}
    for (name, tag), a in agg.items():
        s_ = stages.setdefault(stage_of(name, tag), {"ms": 0.0, "tflop": 0.0, "launches": 0})
        s_["ms"] += a["ms"]
        s_["tflop"] += a["flops"] * a["n"] / 1e12
        s_["launches"] += a["n"]
    for s_ in stages.values():
        s_["tflops"] = s_["tflop"] / (s_["ms"] / 1e3) if s_["ms"] > 0 else 0.0
        s_["frac_of_peak"] = s_["tflops"] / pk["bf16_sustained"]
    tf = flops / (ms / 1e3) / 1e12
    return {
        "metric": METRIC, "value": secs / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "configs[3]: HiFi-GAN Generator alone, mel randn(256, 500, 80)*2.26-5.88 (seed 13) -> waveform, "
                               "decoder.vocoder(mel) with the mel resident in HBM", "batch": b, "frames": t,
                   "l2_policy": "working set (~30 GB of stage tensors) exceeds L2; no explicit flush"},
        "clocks": clk, "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "whole vocoder (conv_pre + 5 x (transposed conv + MRF) + conv_post)",
                     "achieved": tf, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": tf / pk["bf16_sustained"],
                     "traffic": None, "algorithmic_tflop": flops / 1e12, "stages_serialised": stages},
        "arena_gb": eng.arena.nbytes / 2 ** 30,
    }


def config5_run(args, decoder, dev):
    """BASELINE configs[4]: ODE step sweep on long-form utterances: 16 x 3000 frames (60 s each), NFE 1/4/8/16/32."""
    from speech_resynth_b200 import _native as nat
    from speech_resynth_b200 import sharding, synthetic

    b, n = 16, 3000
    ids = synthetic.make_units(b, n, seed=41).to(dev)
    secs = audio_seconds([n] * b)
    eng = decoder.engine()
    pk = peaks()
    sweep = []
    launches = 0
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    for nfe in (1, 4, 8, 16, 32):
        dt = 1.0 / nfe
        for _ in range(3):
            eng.resynthesize(ids, dt, TRUNC)
        torch.cuda.synchronize()
        steps = max(2, min(args.steps, 160 // max(nfe, 4)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = nat.launch_count
        e0.record()
        for _ in range(steps):
            eng.resynthesize(ids, dt, TRUNC)
        e1.record()
        torch.cuda.synchronize()
        launches += nat.launch_count - c0
        ms = e0.elapsed_time(e1) / steps
        fl_t = b * sharding.transformer_flops(n, nfe, hoisted=True)
        fl_v = b * sharding.vocoder_flops(n)
        tf = (fl_t + fl_v) / (ms / 1e3) / 1e12
        sweep.append({"nfe": nfe, "dt": dt, "steps": steps, "ms_per_step": ms, "value": secs / (ms / 1e3),
                      "transformer_tflop": fl_t / 1e12, "vocoder_tflop": fl_v / 1e12, "achieved_tflops": tf,
                      "frac_of_peak": tf / pk["bf16_sustained"]})
    clk = clocks.stop()
    # per-op table at NFE 16
    x0 = torch.randn(b, n, 80, device=dev)
    plan = eng._plan(b, n, DT, True, True)
    agg = profile_body(eng, plan, lambda: eng.sampler.stage(plan.cfm_ws, ids, x0, TRUNC))
    if args.ops:
        write_ops(args.ops, agg)
    total = sum(a["ms"] for a in agg.values())
    att_ops = ("srb_cfm_attention_tc", "srb_cfm_attention_qkv")
    att = sum(a["ms"] for (nm, _), a in agg.items() if nm in att_ops)
    att_fl = sum(a["flops"] * a["n"] for (nm, _), a in agg.items() if nm in att_ops)
    att_name = next(nm for (nm, _) in agg if nm in att_ops)
    head = next(s for s in sweep if s["nfe"] == 16)
    # slope of time against NFE = the transformer's cost per step; intercept = vocoder + fixed
    per_step_ms = (sweep[-1]["ms_per_step"] - sweep[1]["ms_per_step"]) / (sweep[-1]["nfe"] - sweep[1]["nfe"])
    per_step_tf = b * sharding.transformer_flops(n, 1, hoisted=True) / (per_step_ms / 1e3) / 1e12
    return {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": head["steps"], "warmup": 3,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": "configs[4]: ODE step sweep NFE 1/4/8/16/32 on 16 x 3000 units (60 s each), truncation 1.0; "
                               "headline = NFE 16", "batch": b, "frames": n,
                   "l2_policy": "per-step working set (~7 GB) exceeds L2; no explicit flush"},
        "clocks": clk, "gpu_launches": launches, "sweep": sweep,
        "transformer_per_step": {"ms": per_step_ms, "achieved_tflops": per_step_tf, "frac_of_peak": per_step_tf / pk["bf16_sustained"]},
        "roofline": {"bound": "tensor", "kernel": f"{att_name} (64 launches at NFE 16)", "achieved": att_fl / (att / 1e3) / 1e12,
                     "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": att_fl / (att / 1e3) / 1e12 / pk["bf16_sustained"],
                     "traffic": None, "share_of_step": att / total, "avg_launch_ms": att / 64,
                     "whole_step": {"achieved_tflops": head["achieved_tflops"], "frac_of_peak": head["frac_of_peak"]}},
        "arena_gb": eng.arena.nbytes / 2 ** 30,
    }


def _claim_stdout():
    """Keep the process's stdout for the ONE JSON line: libraries (NCCL prints its version banner at communicator
    creation) write to file descriptor 1 directly, so fd 1 is pointed at stderr and the line goes to a duplicate of
    the original descriptor."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager-gpu"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4, 5], help="BASELINE.json configs[] (1-based): 2 = headline "
                    "(with the configs[2] sharded record inside), 4 = vocoder alone, 5 = NFE sweep at 60 s")
    ap.add_argument("--no-config3", action="store_true", help="skip the configs[2] sharded record of the default run")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tight"], help="tight = the tight-precision library (fp32-grade "
                    "results, several times slower); the line then says so in `dtype` and `config.precision`")
    ap.add_argument("--c3-buckets", default=None, help="A/B knob for the configs[2] record: 'tile_budget,max_batch,max_waste' "
                    "passed to sharding.plan_shards instead of its defaults")
    ap.add_argument("--ops", default=None, help="write the per-op device-time table (csv) here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        print(json.dumps(cpu_reference_run(args.steps, args.warmup, args.gpus, as_arm=True)), flush=True)
        return
    if args.impl == "eager-gpu":
        print(json.dumps(eager_gpu_run(args.steps, args.warmup)), flush=True)
        return
    real_stdout = _claim_stdout()
    out = gpu_run(args)
    if out is not None:
        print(json.dumps(out), file=real_stdout, flush=True)


if __name__ == "__main__":
    main()
