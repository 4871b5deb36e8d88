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
