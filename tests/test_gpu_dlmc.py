"""B200: deterministic Langevin Monte Carlo (dlmc) against the reference's golden outputs and the oracle.

Reference: nfmc/dlmc.py:44-119.  The goldens were made with the per-iteration flow refit switched off (the refit is an
optimiser run of the absent torchflows); draw order per iteration: normal(n,d) = base draw of flow.sample, uniform(n).
"""
import numpy as np
import pytest
import torch

from golden_util import load_case, oracle_flow
from oracle import samplers_ref as R
from oracle.potentials_ref import make_potential_ref
from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu


def close(a, b, atol):
    a = torch.as_tensor(np.asarray(a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert err <= atol, f"max abs err {err} > {atol}"


@pytest.mark.parametrize("name", ["dlmc_gm", "dlmc_latent_gm"])
def test_golden_dlmc(name):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import DLMCKernel, DLMCParameters
    from nfmc_b200.samplers import DLMC
    g = load_case(name)
    n, d = g["x0"].shape
    T = int(g["T"])
    s = DLMC((d,), product_target(g["pot"], d), product_target(g["nll"], d),
             DLMCKernel((d,), flow=product_flow_from_oracle(oracle_flow(g)), step_size=float(g["step"])),
             DLMCParameters(n_iterations=T, latent_updates=bool(int(g["latent"]))))
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, z=torch.stack(g["normals"]), uniforms=torch.stack(g["uniforms"]),
                   refit=False)
    ref = torch.from_numpy(g["samples"])
    close(out.samples, ref, atol=5e-5 * max(1.0, float(ref.abs().max())))
    close(out.running_samples.last_sample, g["last"], atol=5e-5 * max(1.0, float(ref.abs().max())))
    close(out.mean, g["mean"], atol=5e-5)
    acc, att, div, grads, calls, _, _ = (int(x) for x in g["counters"])
    st = out.statistics
    assert (st.n_accepted_trajectories, st.n_attempted_trajectories, st.n_divergences) == (acc, att, div)
    assert (st.n_target_gradient_calls, st.n_target_calls) == (grads, calls)


@pytest.mark.parametrize("pot,d,Lc,n,T,latent", [("g1", 100, 2, 517, 3, False), ("gm", 25, 3, 300, 3, False), ("rb", 26, 1, 129, 3, True),
                                                 ("fn", 100, 2, 257, 2, True), ("g0", 1000, 2, 35, 2, False)])
def test_dlmc_against_oracle(pot, d, Lc, n, T, latent):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import DLMCKernel, DLMCParameters
    from nfmc_b200.samplers import DLMC
    torch.manual_seed(n + d)
    oflow = make_flow((d,), n_layers=Lc, perturb=0.05, seed=d)
    x0 = 0.3 * torch.randn(n, d)
    z = torch.randn(T, n, d)
    un = torch.rand(T, n)
    eps = 0.002 if pot == "g1" else 0.02
    tape_n, tape_u = list(z), list(un)
    run = R.run_dlmc(x0, make_potential_ref(pot, (d,)), make_potential_ref("g0", (d,)), oflow, T, R.TapeDraws(tape_n, tape_u),
                     step_size=eps, latent_updates=latent)
    s = DLMC((d,), product_target(pot, d), product_target("g0", d), DLMCKernel((d,), flow=product_flow_from_oracle(oflow), step_size=eps),
             DLMCParameters(n_iterations=T, latent_updates=latent))
    out = s.sample(x0, show_progress=False, z=z, uniforms=un, refit=False)
    ref = run.samples
    scale = max(1.0, float(ref.abs().max()))
    per_chain = (out.samples - ref).abs().amax(dim=(0, 2))
    ok = per_chain <= 2e-4 * scale
    assert ok.float().mean() >= 0.97, float(ok.float().mean())          # an accept decided within rounding error may flip
    assert out.statistics.n_attempted_trajectories == n * T
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~ok).sum()) * T
    assert out.statistics.n_target_calls == run.n_target_calls and out.statistics.n_target_gradient_calls == run.n_grad_calls


def test_dlmc_through_sample_api_with_refits():
    """The full loop with a flow refit every iteration: the particles move towards the target (a shifted Gaussian)."""
    import nfmc_b200
    from nfmc_b200.potentials import DiagonalGaussian
    d, n = 8, 1024
    torch.manual_seed(3)
    target = DiagonalGaussian((d,), precision=torch.full((d,), 4.0), mean=torch.full((d,), 1.5))
    out = nfmc_b200.sample(target, strategy="dlmc", negative_log_likelihood=target, n_chains=n, n_iterations=12,
                           show_progress=False, kernel_kwargs=dict(step_size=0.05),
                           param_kwargs=dict(flow_fit_kwargs=dict(n_epochs=20, lr=0.05, batch_size="adaptive")))
    assert out.samples.shape == (12, n, d)
    last = out.running_samples.last_sample
    assert abs(float(last.mean()) - 1.5) < 0.25, float(last.mean())
    assert float(last.var(dim=0).mean()) < 0.8
    st = out.statistics
    assert st.n_attempted_trajectories == 12 * n
    assert st.n_target_calls == n + 12 * 3 * n and st.n_target_gradient_calls == n + 12 * n
    with pytest.raises(ValueError):
        nfmc_b200.sample(target, strategy="dlmc", n_chains=n, n_iterations=1, show_progress=False)
