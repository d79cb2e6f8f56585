"""Headline benchmark: seconds of audio synthesised per wall-second (RTF^-1) of the unit-to-speech path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one decoder(units) call over one synthetic batch: BASELINE.json configs[1] -- 64 utterances x 500
units (10 s each), dt = 0.0625 (NFE 16, the reference's step count), truncation 1.0, bf16 tensor-core compute.
With N > 1 (torchrun, one process per GPU) every rank synthesises its own batch of that size (weak scaling, no
collective inside the ODE loop or the vocoder) and the waveforms are gathered on rank 0 at the end of each step.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (units already in HBM); `e2e` = through the
public API from pinned host memory with the waveforms copied back to the host; `roofline` = the dominant kernel's
achieved bf16 TFLOP/s (algorithmic FLOPs / CUDA-event time) against MEASURED_PEAKS.json; `cpu_baseline` = the CPU
oracle port timed on this box's host cores on a bounded sample.  `--impl reference` times that CPU path alone.
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
def cpu_reference_run(steps: int, warmup: int, n_gpus: int, as_arm: bool):
    """The reference's algorithm on the host cores (oracle port; the live reference cannot travel to the GPU box).
    Bounded sample: 2 utterances x 500 units per step (same per-utterance workload as the GPU arm)."""
    from oracle import cfm_hifigan_oracle as oracle
    from speech_resynth_b200 import synthetic

    torch.set_num_threads(os.cpu_count() or 1)
    sd = synthetic.make_state_dict(0)
    b = 2
    ids = synthetic.make_units(b, FRAMES, seed=7)
    x0 = torch.randn(b, FRAMES, 80, generator=torch.Generator().manual_seed(1))
    secs = audio_seconds([FRAMES] * b)
    times = []
    with torch.inference_mode():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.resynthesize(sd, ids, x0, DT, TRUNC)
            t1 = time.perf_counter()
            if i >= warmup:
                times.append(t1 - t0)
    total = sum(times)
    value = secs * len(times) / total
    base = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{b} x {FRAMES} units, NFE 16, fp32 torch CPU ops (oracle port of the reference), {len(times)} steps"}
    if not as_arm:
        return base
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(n_gpus),
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


def op_profile(engine, ids, x0):
    """One un-captured pass with CUDA events around every kernel: per-op device time and algorithmic FLOPs."""
    from speech_resynth_b200 import _native as nat

    plan = engine._plan(ids.shape[0], ids.shape[1], DT, TRUNC, True)
    # serialise the parallel graph branches for this pass so every kernel is timed alone
    engine.sampler.fork.enabled = False
    engine.vocoder.fork.enabled = False
    plan.cfm_ws["ids"][:, : ids.shape[1]].copy_(ids)
    plan.cfm_ws["xt"].zero_()
    plan.cfm_ws["xt"][:, : ids.shape[1]].copy_(x0)
    torch.cuda.synchronize()
    nat.profile_log = []
    # keep the GPU busy while the host enqueues the whole pass: an eager launch costs the host 20-30 us, more than
    # many of these kernels run, and an idle GPU would add that wait to the event-to-event time of the next kernel
    torch.cuda._sleep(int(60e-3 * 1.9e9))
    plan.body()
    torch.cuda.synchronize()
    log, nat.profile_log = nat.profile_log, None
    engine.sampler.fork.enabled = True
    engine.vocoder.fork.enabled = True
    agg = {}
    for name, tag, e0, e1, flops, nbytes in log:
        key = (name, tag + (int(flops),))   # launches of one op with different kernel sizes are different rows
        a = agg.setdefault(key, {"ms": 0.0, "n": 0, "flops": flops, "bytes": nbytes})
        a["ms"] += e0.elapsed_time(e1)
        a["n"] += 1
    return agg


def gpu_run(args):
    import torch.distributed as dist

    import speech_resynth_b200 as srb
    from speech_resynth_b200 import _native as nat
    from speech_resynth_b200 import synthetic

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
    engine = decoder.engine()

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
        """One serving step through the public API: H2D of the units, decoder(units), D2H of every waveform.  The
        read-back runs on a second stream (double-buffered pinned host rows) so it overlaps the next step's kernels, as
        a serving loop would do; all of it is inside the timed region (both streams are synchronised at its end)."""
        ids = ids_host.to(dev, non_blocking=True)           # H2D of this step's units
        wavs = decoder(ids, DT, TRUNC)                        # public API (list of per-utterance waveforms, fresh storage)
        host = wav_hosts[e2e_count[0] & 1]
        e2e_count[0] += 1
        copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(copy_stream):
            for i, w in enumerate(wavs):                      # D2H of the result
                w.record_stream(copy_stream)
                host[i, : w.shape[-1]].copy_(w[0], non_blocking=True)
        if world > 1:
            dist.gather(engine._plans[(BATCH, FRAMES, DT, TRUNC, True)].voc_ws["wav"], gather_buf, dst=0)

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

    out = None
    if rank == 0:
        pk = peaks()
        x0 = torch.randn(BATCH, FRAMES, 80, device=dev)
        agg = op_profile(engine, ids_dev, x0)
        total_ms = sum(a["ms"] for a in agg.values())
        (top_name, top_tag), top = max(agg.items(), key=lambda kv: kv[1]["ms"])
        avg_s = top["ms"] / top["n"] / 1e3
        achieved = top["flops"] / avg_s / 1e12
        from speech_resynth_b200 import sharding
        flops_step = BATCH * (sharding.transformer_flops(FRAMES, 16, hoisted=True) + sharding.vocoder_flops(FRAMES))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            # dram read+write bytes per launch from an ncu --set full capture of exactly this launch shape (else null)
            traffic = json.load(open(tpath)).get(f"{top_name}{list(top_tag[:-1])}")
        roofline = {
            "bound": "tensor", "kernel": f"{top_name}{list(top_tag[:-1])}", "achieved": achieved, "peak": pk["bf16_sustained"],
            "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"], "traffic": traffic,
            "algorithmic_flops_per_launch": top["flops"],
            "peak_source": f"{pk['src']} sustained bf16 (kernel timed inside the step)",
            "launches_per_step": top["n"], "avg_launch_ms": top["ms"] / top["n"], "share_of_step": top["ms"] / total_ms,
            "whole_step": {"algorithmic_tflop": flops_step / 1e12,
                           "achieved_tflops": flops_step * world / (ms_dev / args.steps / 1e3) / 1e12 / world,
                           "frac_of_peak": flops_step / (ms_dev / args.steps / 1e3) / 1e12 / pk["bf16_sustained"]},
        }
        cpu = cpu_reference_run(steps=2, warmup=1, n_gpus=1, as_arm=False) if world == 1 else None
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(world), "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": ids_host.numel() * 8,
                    "d2h_bytes_per_step": BATCH * rows * 4 + BATCH * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
        }
        if args.ops:
            rowsout = sorted(agg.items(), key=lambda kv: -kv[1]["ms"])
            with open(args.ops, "w") as f:
                f.write("op,tag,launches,total_ms,avg_ms,share,tflops\n")
                for (n_, t_), a in rowsout:
                    tf = a["flops"] / (a["ms"] / a["n"] / 1e3) / 1e12 if a["ms"] > 0 else 0
                    f.write(f"{n_},\"{list(t_)}\",{a['n']},{a['ms']:.4f},{a['ms'] / a['n']:.4f},{a['ms'] / total_ms:.4f},{tf:.1f}\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


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
