"""Chain-batch data parallelism: one process per GPU, chains sharded contiguously, no data-path collective.

Chains are independent given a frozen flow and kernel parameters (every operator of the path is row-wise over
the chain axis), so rank ``g`` of ``G`` owns rows ``[g*n/G, (g+1)*n/G)`` and the only exchange is the pooled
statistics: one ``all_reduce(SUM)`` of ``[sum_x (d), sum_x2 (d), n_seen, counters]`` per run -- it replaces
``MCMCExpectation.update`` (/root/reference/nfmc/algorithms/sampling/base.py:75-95) and
``MCMCStatistics.update_counters`` (base.py:139-149, jump.py:52-58) across devices.  The Philox streams are keyed
by GLOBAL chain index (``sampler.chain0``), so the pooled result does not depend on G.
Backend: NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .records import MCMCOutput

COUNTER_FIELDS = ("n_accepted_trajectories", "n_attempted_trajectories", "n_divergences", "n_target_gradient_calls",
                  "n_target_calls", "n_accepted_jumps", "n_attempted_jumps")


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global chain, number of chains) of ``rank``; the remainder goes to the first ranks."""
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def pool_statistics(out: MCMCOutput, group=None, device: Optional[torch.device] = None) -> MCMCOutput:
    """All-reduce the moments sums and integer counters of ``out`` in place (one fused fp64 + one int64 message)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return out
    st = out.statistics
    ex = st.expectations
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    d = ex.sum_x.numel()
    fbuf = torch.empty(2 * d + 1, dtype=torch.float64, device=dev)
    fbuf[:d] = ex.sum_x.to(dev)
    fbuf[d:2 * d] = ex.sum_x2.to(dev)
    fbuf[2 * d] = st.elapsed_time_seconds
    ibuf = torch.tensor([ex.n_seen] + [int(getattr(st, f, 0)) for f in COUNTER_FIELDS], dtype=torch.int64, device=dev)
    dist.all_reduce(fbuf[:2 * d], op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(fbuf[2 * d:], op=dist.ReduceOp.MAX, group=group)      # elapsed time: slowest rank
    dist.all_reduce(ibuf, op=dist.ReduceOp.SUM, group=group)
    fb = fbuf.cpu()
    ib = ibuf.cpu().tolist()
    ex.sum_x, ex.sum_x2, ex.n_seen = fb[:d].clone(), fb[d:2 * d].clone(), int(ib[0])
    st.elapsed_time_seconds = float(fb[2 * d])
    for f, v in zip(COUNTER_FIELDS, ib[1:]):
        if hasattr(st, f):
            setattr(st, f, int(v))
    return out


def sample_sharded(sampler, x0_global: torch.Tensor, show_progress: bool = False, time_limit_seconds=None, group=None,
                   **inject) -> MCMCOutput:
    """Run ``sampler`` on this rank's shard of ``x0_global`` and pool the statistics.  ``samples`` /
    ``last_sample`` of the returned output are this rank's shard (rows ``shard_range(...)``)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_range(x0_global.shape[0], rank, world)
    sampler.chain0 = first
    if hasattr(sampler, "inner_sampler"):
        sampler.inner_sampler.chain0 = first
    # Every rank must use the same Philox seed (chains are keyed by global index).  A seed drawn here is valid for THIS call
    # only: each call restarts its step counters at 0, so keeping it would replay the same noise on the next call
    # (warm-up then sampling, or a loop continuing from last_sample).  A seed the user fixed is advanced per call by
    # Sampler.session_seed().
    targets = [sampler] + ([sampler.inner_sampler] if hasattr(sampler, "inner_sampler") else [])
    drawn = sampler.seed is None
    if drawn:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
        if world > 1:
            t = seed.cuda() if dist.get_backend(group) == "nccl" else seed
            dist.broadcast(t, src=0, group=group)
            seed = t.cpu()
        saved = [(s_, s_.seed, getattr(s_, "_n_sessions", 0)) for s_ in targets]
        for s_ in targets:
            s_.seed, s_._n_sessions = int(seed), 0
    try:
        out = sampler.sample(x0_global[first:first + count], show_progress=show_progress,
                             time_limit_seconds=time_limit_seconds, **inject)
    finally:
        if drawn:
            for s_, old_seed, old_n in saved:
                s_.seed, s_._n_sessions = old_seed, old_n
    return pool_statistics(out, group)
