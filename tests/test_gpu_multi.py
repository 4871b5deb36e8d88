"""B200 x >= 2 (skipped on a single-GPU box): one process per GPU over NCCL.

* chain sharding: the pooled moments / counters of a jump_mala run on 2 GPUs equal the single-GPU run over the same
  global chains (Philox is keyed by the global chain index);
* data-parallel flow training: each rank trains on its own rows, gradients are all-reduced between the gradient and the
  AdamW kernels, so the ranks end with identical parameters that fit the pooled data.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import nfmc_b200
        from nfmc_b200.dist import sample_sharded
        from nfmc_b200.flow import Flow, RealNVP, create_flow_object
        from nfmc_b200.potentials import make_potential
        d, n = 20, 4096
        # ---- sharded sampling ------------------------------------------------------------------------------------------
        torch.manual_seed(0)
        flow = create_flow_object("realnvp", (d,))
        with torch.no_grad():
            for p in flow.parameters():
                p.add_(0.05 * torch.randn_like(p))
        x0 = torch.randn(n, d)
        s = nfmc_b200.create_sampler(make_potential("g1", (d,)), flow=flow, strategy="jump_mala",
                                     param_kwargs={"n_iterations": 3, "store_samples": False},
                                     inner_param_kwargs={"n_iterations": 7}, device=torch.device("cuda", rank))
        s.seed = 1234
        s.inner_sampler.seed = 1234
        out = sample_sharded(s, x0)
        stats = out.statistics
        res = dict(mean=np.asarray(out.mean), second=np.asarray(out.second_moment),
                   counters=[stats.n_accepted_trajectories, stats.n_attempted_trajectories, stats.n_accepted_jumps,
                             stats.n_attempted_jumps, stats.n_target_calls])
        # ---- data-parallel training ------------------------------------------------------------------------------------
        g = torch.Generator().manual_seed(100 + rank)
        x = (torch.linspace(0.5, 2.0, d) * torch.randn(2000, d, generator=g) + 0.5).cuda()
        torch.manual_seed(7)                                   # same initial flow and same shuffles on every rank
        f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
        before = float(-f.log_prob(x).mean())
        f.fit(x, n_epochs=8, lr=0.02, batch_size=500)
        after = float(-f.log_prob(x).mean())
        theta = torch.cat([p.detach().reshape(-1) for p in f.parameters()])
        gathered = [torch.empty_like(theta) for _ in range(world)]
        dist.all_gather(gathered, theta)
        res.update(before=before, after=after, theta_spread=float((gathered[0] - gathered[1]).abs().max()))
        # ---- sharded warm-up: the per-iteration all-reduce keeps every rank's kernel identical --------------------
        from nfmc_b200.records import LangevinKernel, LangevinParameters
        from nfmc_b200.samplers import MALA
        w = MALA((d,), make_potential("g1", (d,)), LangevinKernel(event_size=d, step_size=0.05), LangevinParameters(n_iterations=12))
        w.device = torch.device("cuda", rank)
        w.seed = 77
        w.params.tuning_mode()
        sample_sharded(w, x0)
        res.update(tuned_step=float(w.kernel.step_size), tuned_imd=w.kernel.inv_mass_diag.cpu().numpy())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_sharding_and_training():
    _need_two_gpus()
    import nfmc_b200
    from nfmc_b200.flow import create_flow_object
    from nfmc_b200.potentials import make_potential
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=500) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    # single-GPU run over the same global chains
    d, n = 20, 4096
    torch.manual_seed(0)
    flow = create_flow_object("realnvp", (d,))
    with torch.no_grad():
        for p in flow.parameters():
            p.add_(0.05 * torch.randn_like(p))
    x0 = torch.randn(n, d)
    s = nfmc_b200.create_sampler(make_potential("g1", (d,)), flow=flow, strategy="jump_mala",
                                 param_kwargs={"n_iterations": 3, "store_samples": False},
                                 inner_param_kwargs={"n_iterations": 7})
    s.seed = 1234
    s.inner_sampler.seed = 1234
    ref = s.sample(x0, show_progress=False)
    st = ref.statistics
    ref_counters = [st.n_accepted_trajectories, st.n_attempted_trajectories, st.n_accepted_jumps, st.n_attempted_jumps,
                    st.n_target_calls]
    for rank in (0, 1):
        r = res[rank]
        assert r["counters"] == ref_counters, (rank, r["counters"], ref_counters)
        np.testing.assert_allclose(r["mean"], np.asarray(ref.mean), atol=1e-6)
        np.testing.assert_allclose(r["second"], np.asarray(ref.second_moment), rtol=1e-6, atol=1e-6)
        assert r["theta_spread"] == 0.0
        assert r["after"] < r["before"] - 1.0
    # warm-up: both ranks hold the same tuned kernel, and it equals the single-GPU warm-up over all chains
    from nfmc_b200.records import LangevinKernel, LangevinParameters
    from nfmc_b200.samplers import MALA
    w = MALA((d,), make_potential("g1", (d,)), LangevinKernel(event_size=d, step_size=0.05), LangevinParameters(n_iterations=12))
    w.seed = 77
    w.params.tuning_mode()
    w.sample(x0, show_progress=False)
    assert res[0]["tuned_step"] == res[1]["tuned_step"]
    np.testing.assert_array_equal(res[0]["tuned_imd"], res[1]["tuned_imd"])
    assert abs(res[0]["tuned_step"] - float(w.kernel.step_size)) <= 1e-9 * float(w.kernel.step_size)
    np.testing.assert_allclose(res[0]["tuned_imd"], w.kernel.inv_mass_diag.cpu().numpy(), rtol=1e-6)
