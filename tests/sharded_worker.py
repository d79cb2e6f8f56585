"""One rank of the multi-GPU sharded-resynthesis check (launched by tests/test_gpu_multi.py under torchrun, one process
per GPU, NCCL): BASELINE configs[2] scaled down -- ragged utterances through sharding.resynthesize_sharded with the
product decoder; rank 0 compares sampled buckets of EVERY rank with the CPU oracle and prints `SHARDED-OK`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def bucket_seed(ids):
    return 1000 + int(ids.sum()) % 100003


def main():
    import speech_resynth_b200 as srb
    from oracle import cfm_hifigan_oracle as oracle
    from speech_resynth_b200 import sharding, synthetic

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = synthetic.make_state_dict(0)
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(sd, strict=True)
    decoder = decoder.to(dev)

    n_utts, dt, tv, nfe = int(os.environ.get("SRB_TEST_UTTS", "160")), 0.25, 1.0, 4
    gen = torch.Generator().manual_seed(11)
    lengths = torch.randint(100, 1001, (n_utts,), generator=gen).tolist()
    units = [torch.randint(1, 2001, (n,), generator=gen) for n in lengths]

    def synth_into(ids, out):
        torch.manual_seed(bucket_seed(ids))      # one reproducible prior per bucket, the same on whichever GPU it runs
        decoder.resynthesize_flat(ids, dt, tv, out=out)

    stats = {}
    outs = sharding.resynthesize_sharded(units, None, rank=rank, world=world, nfe=nfe, device=dev, synth_into=synth_into,
                                         stats=stats, on_plan=lambda p: decoder.engine().reserve(
                                             *max(((b.batch, b.frames) for j in p.per_rank[rank] for b in [p.buckets[j]]),
                                                  key=lambda bf: bf[0] * bf[1])))
    plan = stats["plan"]
    assert len(plan.per_rank) == world and all(len(r) >= 1 for r in plan.per_rank)
    ok = True
    if rank == 0:
        assert [o.shape[-1] for o in outs] == [320 * n + 80 for n in lengths]
        assert all(bool(torch.isfinite(o).all()) for o in outs)
        worst = 0.0
        for r in range(world):
            # the last (shortest, cheapest for the oracle) bucket of every rank
            bk = plan.buckets[plan.per_rank[r][-1]]
            ids = sharding.pad_bucket(units, bk)
            torch.manual_seed(bucket_seed(ids))
            x0 = torch.randn(bk.batch, bk.frames, 80, device=dev).cpu()
            ref = oracle.resynthesize(sd, ids, x0, dt, tv)
            for i, w in zip(bk.indices, ref):
                e = float((outs[i].double().cpu() - w.double()).norm() / w.double().norm())
                worst = max(worst, e)
        print(f"[parity] sharded world={world}: worst waveform rel-L2 vs oracle over one bucket per rank = {worst:.3e}", flush=True)
        ok = worst <= 5e-3
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and ok:
        print("SHARDED-OK", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
