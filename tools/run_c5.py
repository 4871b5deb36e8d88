#!/usr/bin/env python
"""BASELINE config C5: jump_mala, d = 1000, 2^20 chains over the GPUs of one node, mixture target, flow refitted every outer
iteration (fit_nf=True) on rows of each rank's shard with NCCL gradient all-reduce.  Launch with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 tools/run_c5.py [--chains N] [--iters T]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import nfmc_b200
from nfmc_b200.dist import sample_sharded, shard_range
from nfmc_b200.potentials import make_potential


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=1 << 20)
    ap.add_argument("--dim", type=int, default=1000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--inner", type=int, default=100)
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--strategy", default="jump_mala", help="jump_mala (C5) or imh / adaptive_imh (C4: --dim 100 --potential rb)")
    ap.add_argument("--potential", default="gm")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, n = a.dim, a.chains
    torch.manual_seed(0)
    pk = {"n_iterations": a.iters, "store_samples": False}
    jump = a.strategy.startswith("jump_")
    if jump and not a.no_fit:
        pk.update(fit_nf=True, n_jumps_before_training=0)
    extra = {"inner_param_kwargs": {"n_iterations": a.inner}} if jump else {}
    s = nfmc_b200.create_sampler(make_potential(a.potential, (d,)), event_shape=(d,), flow="realnvp", strategy=a.strategy,
                                 param_kwargs=pk, device=dev, **extra)
    first, count = shard_range(n, rank, world)
    g = torch.Generator(device=dev).manual_seed(rank)
    x0_shard = torch.randn(count, d, device=dev, generator=g)

    class _Global:                      # sample_sharded slices rows [first, first+count) of the "global" x0
        shape = (n, d)
        def __getitem__(self, sl):
            return x0_shard
    sample_sharded(s, _Global())        # untimed: module load, NCCL channels, workspaces
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = sample_sharded(s, _Global())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    st = out.statistics
    if rank == 0:
        print(json.dumps({"config": "%s %s d=%d chains=%d gpus=%d fit_nf=%s" % (a.strategy, a.potential, d, n, world, jump and not a.no_fit), "seconds": dt,
                          "chain_steps_per_s": st.expectations.n_seen / dt, "device_seconds": st.elapsed_time_seconds,
                          "acc_rate": st.acceptance_rate, "jump_acc_rate": getattr(st, "jump_acceptance_rate", None),
                          "n_seen": st.expectations.n_seen}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
