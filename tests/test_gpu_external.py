"""B200: the external-target path (a plain Python callable as ``target``, the reference's own contract: sample.py:34-36,
README.md:45-46) against the fused path and the oracle.  The golden replays with a callable target live in
test_gpu_parity.py (``ext=True`` parametrisation); here: Philox-mode equivalence with the fused kernels, the README
example through the public API, the sample sink / thinning, and the stationary moments."""
import ctypes as C

import numpy as np
import pytest
import torch

import nfmc_b200
from nfmc_b200 import _native as N
from nfmc_b200.records import (HMCKernel, HMCParameters, JumpNFMCParameters, LangevinKernel, LangevinParameters, MHKernel,
                               MHParameters, NFMCKernel, IMHKernel, IMHParameters)
from nfmc_b200.samplers import HMC, MALA, MH, FixedIMH, JumpMALA

pytestmark = pytest.mark.gpu


def readme_target(x):                                      # README.md:45-46
    return torch.sum(x ** 2, dim=1)


def _pair(cls, d, kernel, params):
    """The same sampler twice: built-in StandardGaussian (fused kernels) and the README lambda (external path)."""
    from copy import deepcopy
    a = cls((d,), nfmc_b200.potentials.StandardGaussian((d,)), deepcopy(kernel), deepcopy(params))
    b = cls((d,), readme_target, deepcopy(kernel), deepcopy(params))
    assert not a.target.external and b.target.external
    a.seed = b.seed = 1234
    return a, b


@pytest.mark.parametrize("d,n", [(25, 300), (100, 517), (7, 33)])
def test_mala_callable_equals_fused_in_philox_mode(d, n):
    """No injected noise: the external path draws its numbers with nfmc_rng_fill from the same Philox counters the fused
    kernel uses, so both runs take the same decisions and end (within fp32 reduction order) in the same state."""
    K = 12
    imd = torch.linspace(0.5, 1.5, d)
    a, b = _pair(MALA, d, LangevinKernel(event_size=d, inv_mass_diag=imd, step_size=0.2 * d ** (-1 / 3)), LangevinParameters(n_iterations=K))
    x0 = torch.randn(n, d)
    oa = a.sample(x0, show_progress=False)
    ob = b.sample(x0, show_progress=False)
    assert oa.statistics.n_attempted_trajectories == ob.statistics.n_attempted_trajectories == n * K
    assert abs(oa.statistics.n_accepted_trajectories - ob.statistics.n_accepted_trajectories) <= max(2, n * K // 2000)
    assert float(torch.quantile((oa.samples - ob.samples).abs().flatten(), 0.99)) < 1e-4
    np.testing.assert_allclose(oa.mean.numpy(), ob.mean.numpy(), atol=2e-3)
    assert (oa.statistics.n_target_calls, oa.statistics.n_target_gradient_calls) == (ob.statistics.n_target_calls, ob.statistics.n_target_gradient_calls)


@pytest.mark.parametrize("cls,kernel,params", [
    (HMC, HMCKernel(event_size=25, step_size=0.05, n_leapfrog_steps=7), HMCParameters(n_iterations=6)),
    (MH, MHKernel(event_size=25, inv_mass_diag=torch.full((25,), 0.15)), MHParameters(n_iterations=9)),
])
def test_hmc_mh_callable_equals_fused_in_philox_mode(cls, kernel, params):
    d, n = 25, 411
    a, b = _pair(cls, d, kernel, params)
    x0 = 0.7 * torch.randn(n, d)
    oa, ob = a.sample(x0, show_progress=False), b.sample(x0, show_progress=False)
    assert oa.statistics.n_attempted_trajectories == ob.statistics.n_attempted_trajectories
    assert abs(oa.statistics.n_accepted_trajectories - ob.statistics.n_accepted_trajectories) <= 3
    assert float(torch.quantile((oa.samples - ob.samples).abs().flatten(), 0.99)) < 1e-4


def test_jump_and_imh_callable_equal_fused_in_philox_mode():
    from gpu_util import product_flow_from_oracle
    from oracle.realnvp_ref import make_flow
    d, n, T, K = 25, 389, 3, 4
    flow = product_flow_from_oracle(make_flow((d,), perturb=0.1, seed=5))
    x0 = 0.7 * torch.randn(n, d)
    outs = []
    for tgt in (nfmc_b200.potentials.StandardGaussian((d,)), readme_target):
        s = JumpMALA((d,), tgt, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                     inner_kernel=LangevinKernel(event_size=d, step_size=0.1), inner_params=LangevinParameters(n_iterations=K))
        s.seed = 99
        outs.append(s.sample(x0, show_progress=True))      # progress on: the per-iteration loop on both sides
    a, b = outs
    assert a.samples.shape == b.samples.shape == (T * (K + 1), n, d)
    assert a.statistics.n_attempted_jumps == b.statistics.n_attempted_jumps == n * T
    assert abs(a.statistics.n_accepted_jumps - b.statistics.n_accepted_jumps) <= 2
    assert float(torch.quantile((a.samples - b.samples).abs().flatten(), 0.98)) < 1e-4
    outs = []
    for tgt in (nfmc_b200.potentials.StandardGaussian((d,)), readme_target):
        s = FixedIMH((d,), tgt, IMHKernel((d,), flow=flow), IMHParameters(n_iterations=7))
        s.seed = 7
        outs.append(s.sample(x0, show_progress=False))
    a, b = outs
    assert a.samples.shape == b.samples.shape == (7, n, d)
    assert abs(a.statistics.n_accepted_trajectories - b.statistics.n_accepted_trajectories) <= 2
    assert float(torch.quantile((a.samples - b.samples).abs().flatten(), 0.98)) < 1e-4
    assert a.statistics.n_target_calls == b.statistics.n_target_calls


def test_ess_callable_equals_fused_in_philox_mode():
    """ESS with a lambda likelihood draws nu (stream 0) and its 2 + M scalar uniforms (stream 2) from the same Philox counters
    as the fused kernel: same prior restart, same brackets, same states up to fp32 reduction order."""
    from nfmc_b200.records import ESSKernel, ESSParameters
    from nfmc_b200.samplers import ESS
    d, n, K, M = 25, 333, 6, 5
    outs = []
    for nll in (nfmc_b200.potentials.StandardGaussian((d,)), readme_target):
        s = ESS((d,), nll, nll, ESSKernel(event_shape=(d,)), ESSParameters(n_iterations=K, max_ess_step_iterations=M))
        s.seed = 21
        outs.append(s.sample(torch.randn(n, d), show_progress=False))
    a, b = outs
    assert a.samples.shape == b.samples.shape == (K, n, d)
    assert float(torch.quantile((a.samples - b.samples).abs().flatten(), 0.99)) < 1e-4
    assert a.statistics.n_accepted_trajectories == b.statistics.n_accepted_trajectories == n * K
    assert a.statistics.n_target_calls == b.statistics.n_target_calls


def test_readme_example_runs_with_a_lambda_target():
    """The reference README's own call (README.md:40-52), here with fewer iterations."""
    torch.manual_seed(0)
    out = nfmc_b200.sample(lambda x: torch.sum(x ** 2, dim=1), event_shape=(25,), strategy="jump_mala", n_chains=100,
                           n_iterations=20, show_progress=False, inner_param_kwargs=dict(n_iterations=10))
    assert out.samples.shape == (20 * 11, 100, 25)
    assert bool(torch.isfinite(out.samples).all())
    assert out.statistics.n_attempted_jumps == 20 * 100 and 0 < out.statistics.acceptance_rate <= 1
    assert out.running_samples.last_sample.shape == (100, 25)


def test_callable_target_with_a_matrix_event_shape_and_non_unit_mass():
    """event_shape = (4, 5): the callable sees [n, 4, 5]; HMC with a non-identity inverse mass; jump_hmc through the API."""
    from copy import deepcopy
    es = (4, 5)
    fn = lambda x: (x ** 2).sum(dim=(1, 2))                                               # noqa: E731
    imd = torch.linspace(0.5, 2.0, 20)
    k = HMCKernel(event_size=20, inv_mass_diag=imd, step_size=0.05, n_leapfrog_steps=4)
    a = HMC(es, nfmc_b200.potentials.StandardGaussian(es), deepcopy(k), HMCParameters(n_iterations=5))
    b = HMC(es, fn, deepcopy(k), HMCParameters(n_iterations=5))
    a.seed = b.seed = 11
    x0 = 0.7 * torch.randn(257, *es)
    oa, ob = a.sample(x0, show_progress=False), b.sample(x0, show_progress=False)
    assert oa.samples.shape == ob.samples.shape == (5, 257, 4, 5)
    assert abs(oa.statistics.n_accepted_trajectories - ob.statistics.n_accepted_trajectories) <= 2
    assert float(torch.quantile((oa.samples - ob.samples).abs().flatten(), 0.99)) < 1e-4
    out = nfmc_b200.sample(fn, event_shape=es, strategy="jump_hmc", n_chains=64, n_iterations=3, show_progress=False)
    assert out.samples.shape == (3 * 6, 64, 4, 5) and bool(torch.isfinite(out.samples).all())


def test_callable_target_thinning_and_moments():
    """Sample sink with thinning on the external path, and the stationary moments of N(0, I/2) under MALA."""
    d, n, K = 10, 2048, 60
    s = MALA((d,), readme_target, LangevinKernel(event_size=d, step_size=0.15), LangevinParameters(n_iterations=K))
    out = s.sample(torch.randn(n, d) * 0.7, show_progress=False)
    assert out.samples.shape == (K, n, d)
    late = out.samples[K // 2:]
    assert abs(float(late.mean())) < 0.02 and abs(float(late.var()) - 0.5) < 0.03
    np.testing.assert_allclose(out.mean.numpy(), out.samples.mean(dim=(0, 1)).numpy(), atol=1e-4)
    np.testing.assert_allclose(out.second_moment.numpy(), (out.samples ** 2).mean(dim=(0, 1)).numpy(), atol=1e-4)
    from nfmc_b200.records import MCMCOutput
    from nfmc_b200.samplers import DeviceSession
    out2 = MCMCOutput((d,), store_samples=True)
    out2.running_samples.thinning = 4
    ses = DeviceSession(torch.randn(64, d), (d,), None, seed=3)
    buf = s.run_steps(ses, out2, 10, True)
    assert buf.shape[0] == 3                                 # steps 0, 4, 8 of 10
    assert bool(torch.isfinite(buf).all())


@pytest.mark.parametrize("d,n_layers,ck", [(6, 2, None), (7, 3, dict(n_layers=3, n_hidden=6)), (100, 2, None), (26, 3, None)])
def test_neutra_pullback_against_autograd_through_the_oracle(d, n_layers, ck):
    """nfmc_neutra_pullback (inverse pass + reversible sweep seeded with an external grad U) against autograd of
    U(T^-1 z) - log|det dT^-1/dz| through the oracle flow, for a callable U."""
    from gpu_util import product_flow_from_oracle
    from nfmc_b200.external import LatentTarget
    from nfmc_b200.potentials import CallablePotential
    from oracle.realnvp_ref import make_flow
    torch.manual_seed(d)
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=0.1, seed=d)
    flow = product_flow_from_oracle(oflow)
    fn = lambda x: torch.sum(x ** 2, dim=1) + 0.3 * torch.sum(torch.sin(x), dim=1)     # noqa: E731
    z = (0.6 * torch.randn(53, d)).requires_grad_(True)
    x_ref, ld_ref = oflow.bijection.inverse(z)
    val_ref = fn(x_ref) - ld_ref
    (g_ref,) = torch.autograd.grad(val_ref.sum(), z)
    val, gz = LatentTarget(CallablePotential(fn, (d,)), flow).value_and_grad(z.detach().cuda())
    scale = 1.0 + float(val_ref.detach().abs().max())
    assert float((val.cpu() - val_ref.detach()).abs().max()) < 1e-4 * scale
    assert float((gz.cpu() - g_ref).abs().max()) < 1e-4 * (1.0 + float(g_ref.abs().max()))


def test_warmup_with_a_callable_target():
    """The warm-up paths that differentiate the target: IMH / NeuTra variational fit (reverse KL through the native wide
    trainer with autograd's grad U), MALA step-size tuning, and jump_mala's flow fit -- all with a lambda target."""
    torch.manual_seed(1)
    d = 8
    fn = lambda x: torch.sum(x ** 2, dim=1)                                              # noqa: E731
    out = nfmc_b200.sample(fn, event_shape=(d,), strategy="imh", n_chains=256, n_iterations=30, n_warmup_iterations=5, warmup=True,
                           show_progress=False, param_kwargs=dict(warmup_fit_kwargs=dict(n_epochs=60, n_samples=64)))
    assert out.samples.shape == (30, 256, d) and bool(torch.isfinite(out.samples).all())
    assert out.statistics.acceptance_rate > 0.2                                          # the fitted flow proposes well
    assert abs(float(out.samples[10:].var()) - 0.5) < 0.08
    out = nfmc_b200.sample(fn, event_shape=(d,), strategy="neutra_hmc", n_chains=128, n_iterations=10, n_warmup_iterations=5,
                           warmup=True, show_progress=False, param_kwargs=dict(warmup_fit_kwargs=dict(n_epochs=40, n_samples=64)),
                           inner_kernel_kwargs=dict(n_leapfrog_steps=5, step_size=0.1))
    assert out.samples.shape == (10, 128, d) and bool(torch.isfinite(out.samples).all())
    out = nfmc_b200.sample(fn, event_shape=(d,), strategy="jump_mala", n_chains=128, n_iterations=5, n_warmup_iterations=20,
                           warmup=True, show_progress=False, inner_param_kwargs=dict(n_iterations=5))
    assert out.samples.shape == (5 * 6, 128, d) and bool(torch.isfinite(out.samples).all())


def test_ext_entry_points_validate_arguments():
    lib = N.lib()
    x = torch.zeros(4, 3, device="cuda")
    assert lib.nfmc_ext_langevin_propose(N.ptr(x), None, N.ptr(x), None, 0.1, 0, 4, 3, N.ptr(x), None) != 0     # grad missing
    assert b"ext_langevin_propose" in lib.nfmc_last_error()
    assert lib.nfmc_ext_accept(N.ptr(x), N.ptr(x), None, None, 1, 4, 3, None, None, None, None, None, None, None, None, 0, None) != 0
    assert lib.nfmc_ext_hmc_leapfrog(N.ptr(x), N.ptr(x), N.ptr(x), None, 0.1, 3, 1, 4, 3, None) != 0              # kicks out of range
