"""CPU, world_size 2 over gloo: the N>1 host logic (sharding + pooled statistics)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nfmc_b200.dist import pool_statistics, shard_range
    from nfmc_b200.records import JumpNFMCOutput
    torch.manual_seed(0)
    n, d, k = 37, 5, 4
    x = torch.randn(k, n, d)                       # the "global" run every rank can reproduce
    first, count = shard_range(n, rank, world)
    out = JumpNFMCOutput((d,), store_samples=False)
    out.statistics.expectations.update(x[:, first:first + count])
    out.statistics.update_counters(n_accepted_trajectories=10 * (rank + 1), n_attempted_trajectories=count * k,
                                   n_accepted_jumps=rank + 1, n_attempted_jumps=count, n_target_calls=2 * count * k)
    out.statistics.update_elapsed_time(0.5 + rank)
    pool_statistics(out)
    st = out.statistics
    ok = (torch.allclose(st.running_first_moment, x.mean((0, 1)), atol=1e-6)
          and torch.allclose(st.running_second_moment, (x ** 2).mean((0, 1)), atol=1e-6)
          and st.expectations.n_seen == n * k and st.n_attempted_trajectories == n * k and st.n_accepted_trajectories == 30
          and st.n_accepted_jumps == 3 and st.n_attempted_jumps == n and st.n_target_calls == 2 * n * k
          and abs(st.elapsed_time_seconds - 1.5) < 1e-12)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_shard_range_covers_everything():
    from nfmc_b200.dist import shard_range
    for n in [1, 7, 8, 100, 1 << 20]:
        for w in [1, 2, 3, 8]:
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            for (f0, c0), (f1, _) in zip(parts, parts[1:]):
                assert f0 + c0 == f1


@pytest.mark.timeout(120)
def test_pool_statistics_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res == [(0, True), (1, True)]


def _warmup_worker(rank, world, port, q):
    """The warm-up collective of SURVEY 8(e)-3 over gloo: every rank reduces [sum x (d), sum x^2 (d), accepted, n] of its own
    chains (what nfmc_chain_sums writes), one all-reduce pools them, and the update rule of nfmc_tune_inv_mass / the dual
    averaging (restated here in torch -- the kernels themselves need a GPU) gives every rank the SAME inverse mass and step
    size, equal to what a single process gets from all chains (mcmc/base.py:142-161)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nfmc_b200.dist import shard_range
    from nfmc_b200.records import DualAveraging, DualAveragingParams
    torch.manual_seed(3)
    n, d, iters, c = 51, 6, 5, 1e-3
    xs = torch.randn(iters, n, d) * torch.linspace(0.5, 3.0, d)          # the chains after each warm-up iteration
    masks = torch.rand(iters, n) < 0.6
    first, count = shard_range(n, rank, world)

    def run(lo, hi, pooled):
        imd, da = torch.ones(d), DualAveraging(0.1, DualAveragingParams())
        steps = []
        for i in range(iters):
            x, m = xs[i, lo:hi].double(), masks[i, lo:hi]
            sums = torch.cat([x.sum(0), (x * x).sum(0), torch.tensor([float(m.sum()), float(hi - lo)], dtype=torch.float64)])
            if pooled:
                dist.all_reduce(sums)
            N_ = sums[2 * d + 1]
            mean = sums[:d] / N_
            var = (sums[d:2 * d] - N_ * mean * mean) / (N_ - 1)
            imd = (c * var + (1 - c) * imd.double()).float()
            da.step(0.651 - float(sums[2 * d] / N_))
            steps.append(da.value)
        return imd, steps

    imd_p, steps_p = run(first, first + count, True)
    imd_1, steps_1 = run(0, n, False)
    ref_var = torch.var(xs[-1], dim=0)                                     # the reference's own estimator on all chains
    x_last = xs[-1].double()
    ok = (torch.allclose(imd_p, imd_1, rtol=1e-6) and all(abs(a - b) < 1e-12 * abs(b) for a, b in zip(steps_p, steps_1))
          and torch.allclose(((x_last ** 2).sum(0) - n * x_last.mean(0) ** 2) / (n - 1), ref_var.double(), rtol=1e-5))
    q.put((rank, bool(ok), [float(v) for v in imd_p], steps_p[-1]))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_warmup_collective_world_size_2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 7) % 2000)
    procs = [ctx.Process(target=_warmup_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert res[0][1] and res[1][1]
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]          # identical kernel parameters on both ranks


def test_sample_sharded_does_not_keep_a_drawn_seed():
    """A seed drawn inside sample_sharded is valid for that call only (each call restarts its step counters at 0); a seed
    the user fixed is advanced per call."""
    from nfmc_b200.dist import sample_sharded

    class FakeSampler:
        def __init__(self):
            self.seed, self._n_sessions, self.chain0, self.seen = None, 0, 0, []

        def sample(self, x0, show_progress=False, time_limit_seconds=None):
            from nfmc_b200.samplers import Sampler
            self.seen.append(Sampler.session_seed(self))
            from nfmc_b200.records import MCMCOutput
            return MCMCOutput((x0.shape[1],), store_samples=False)

    s = FakeSampler()
    x0 = torch.zeros(8, 3)
    sample_sharded(s, x0)
    sample_sharded(s, x0)
    assert s.seed is None and s.seen[0] is not None and s.seen[0] != s.seen[1]
    s.seed = 42
    sample_sharded(s, x0)
    sample_sharded(s, x0)
    assert s.seed == 42 and s.seen[2] != s.seen[3] and s.seen[2] == 42
