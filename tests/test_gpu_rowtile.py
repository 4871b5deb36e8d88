"""Row-tile fp32 flow passes (csrc/train_wide.cu: nfmc_flow_wide_pass / nfmc_flow_wide_log_prob) and the NF jump / IMH step
composed from them (nfmc_jump_step_wide) -- the sampling path of every conditioner shape outside the register-resident
(M = 2, H <= 8) and tensor-core (M = 2, even d <= 128) kernels: deep conditioners such as the reference's
``n_layers=5, n_hidden=100`` (/root/reference/test/test_flow_kwargs.py:49), odd d, d > 128 with a wide conditioner.
Oracle: oracle/realnvp_ref.py and oracle/samplers_ref.py with the draws injected; tolerance rtol 1e-4 (fp32)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import samplers_ref as R                              # noqa: E402  (checker only)
from oracle.potentials_ref import make_potential_ref             # noqa: E402
from oracle.realnvp_ref import make_flow                          # noqa: E402

DEEP = [(37, 3, dict(n_layers=3, n_hidden=20)), (100, 2, dict(n_layers=5, n_hidden=100)), (101, 2, dict(n_layers=2, n_hidden=64)),
        (8, 1, dict(n_layers=1)), (200, 2, dict(n_layers=2, n_hidden=24))]


def _pair(d, n_layers, ck, seed=4, perturb=0.05):
    from gpu_util import product_flow_from_oracle
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=perturb, seed=seed)
    return oflow, product_flow_from_oracle(oflow)


def test_routing():
    from nfmc_b200.flow import RealNVP
    assert not RealNVP((100,)).uses_row_tile_pass()                                                        # default: registers
    assert not RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_row_tile_pass()      # tcgen05
    assert RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=64), conditioner_dtype="fp32").uses_row_tile_pass()
    assert RealNVP((100,), n_layers=10, conditioner_kwargs=dict(n_layers=5, n_hidden=100)).uses_row_tile_pass()   # the reference's deep shape
    assert RealNVP((101,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_row_tile_pass()          # odd d
    assert RealNVP((1000,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_row_tile_pass()         # d > 128


@pytest.mark.parametrize("d,n_layers,ck", DEEP)
def test_passes_and_log_prob_against_oracle(d, n_layers, ck):
    oflow, flow = _pair(d, n_layers, ck)
    assert flow.bijection.uses_row_tile_pass()
    n = 333                                              # not a multiple of any row tile
    torch.manual_seed(d)
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, ld_ref = oflow.bijection.forward(x)
        xi_ref, ldi_ref = oflow.bijection.inverse(x)
        lp_ref = oflow.log_prob(x)
    z, ld = flow.bijection.forward(x.cuda())
    xi, ldi = flow.bijection.inverse(x.cuda())
    lp = flow.log_prob(x.cuda())
    for got, want in ((z, z_ref), (ld, ld_ref), (xi, xi_ref), (ldi, ldi_ref), (lp, lp_ref)):
        np.testing.assert_allclose(got.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4 * max(1.0, float(want.abs().max())))


@pytest.mark.parametrize("d,n_layers,ck", DEEP[:3])
def test_generic_per_chain_path_still_agrees(monkeypatch, d, n_layers, ck):
    """NFMC_B200_NO_ROW_TILE=1 keeps the per-chain generic conditioner of flow.cuh (still what Flow.sample, the fused NeuTra
    kernel and tess use for these shapes): both paths give the same numbers."""
    oflow, flow = _pair(d, n_layers, ck)
    x = torch.randn(129, d, generator=torch.Generator().manual_seed(1)).cuda()
    z, ld = flow.bijection.forward(x)
    lp = flow.log_prob(x)
    monkeypatch.setenv("NFMC_B200_NO_ROW_TILE", "1")
    assert not flow.bijection.uses_row_tile_pass()
    z2, ld2 = flow.bijection.forward(x)
    lp2 = flow.log_prob(x)
    for a, b in ((z, z2), (ld, ld2), (lp, lp2)):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize("d,n_layers,ck", DEEP[:3])
def test_fixed_imh_against_oracle(d, n_layers, ck):
    from gpu_util import product_target
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import FixedIMH
    oflow, flow = _pair(d, n_layers, ck, perturb=0.03)
    n, T = 777, 4
    torch.manual_seed(5)
    x0 = torch.randn(n, d)
    z, u = torch.randn(T, n, d), torch.rand(T, n)
    run = R.run_fixed_imh(x0, make_potential_ref("g0", (d,)), oflow, T, R.TapeDraws(list(z), list(u)), trace=True)
    s = FixedIMH((d,), product_target("g0", d), IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    out = s.sample(x0, show_progress=False, z=z, uniforms=u)
    la = torch.stack(run.trace["log_alpha"])
    clear = ((la - torch.log(u)).abs().min(dim=0).values > 1e-3 * (1 + la.abs().max(dim=0).values))
    assert clear.float().mean() > 0.97
    np.testing.assert_allclose(out.samples[:, clear].numpy(), run.samples[:, clear].numpy(), rtol=1e-4,
                               atol=2e-5 * max(1.0, float(run.samples.abs().max())))
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~clear).sum()) * T
    assert out.statistics.n_attempted_trajectories == n * T


def test_jump_mala_deep_flow_philox_equals_injected_and_counters():
    """jump_mala with the reference's deep conditioner shape: the run drawing its own Philox numbers equals the run fed those
    numbers (nfmc_rng_fill, stream 1 for the base draw), and the counters follow the reference's formulas (jump.py:214-239)."""
    import nfmc_b200
    from nfmc_b200 import _native as N
    from nfmc_b200.potentials import StandardGaussian
    d, n, T, K = 100, 600, 3, 4
    oflow, flow = _pair(d, 3, dict(n_layers=5, n_hidden=100), perturb=0.01)
    torch.manual_seed(2)
    x0 = 0.7 * torch.randn(n, d)

    def make():
        s = nfmc_b200.create_sampler(StandardGaussian((d,)), flow=flow, strategy="jump_mala", param_kwargs=dict(n_iterations=T),
                                     inner_param_kwargs=dict(n_iterations=K))
        s.seed = 77
        return s
    out = make().sample(x0, show_progress=False)
    st = out.statistics
    assert st.n_attempted_jumps == n * T and st.n_attempted_trajectories == n * T * K
    assert out.samples.shape == (T * (K + 1), n, d) and bool(torch.isfinite(out.samples).all())
    dev = torch.device("cuda")
    jz, ju = torch.empty(T, n, d, device=dev), torch.empty(T, n, device=dev)
    for t in range(T):
        rng = N.rng_desc(77, t, None, None)
        N.check(N.lib().nfmc_rng_fill(C.byref(rng), 1, 0, d, n, 1, N.ptr(jz[t]), N.ptr(ju[t]), N.stream_ptr(dev)))
    out2 = make().sample(x0, show_progress=False, jump_z=jz, jump_uniforms=ju)
    assert torch.equal(out.samples, out2.samples)
    assert out2.statistics.n_accepted_jumps == st.n_accepted_jumps


@pytest.mark.parametrize("d,n_layers,ck,pot", [(37, 3, dict(n_layers=3, n_hidden=20), "gm"), (100, 2, dict(n_layers=5, n_hidden=100), "fn"),
                                               (101, 2, dict(n_layers=2, n_hidden=64), "g1")])
def test_neutra_latent_potential_and_gradient(d, n_layers, ck, pot):
    """U~(z) = U(T^-1 z) - log|det dT^-1/dz| and its gradient (neutra.py:58-68) composed from the row-tile inverse pass, the
    potential kernel and the row-tile backward sweep, against autograd through the oracle flow."""
    from gpu_util import product_target
    from nfmc_b200.external import LatentTarget
    oflow, flow = _pair(d, n_layers, ck, perturb=0.05)
    torch.manual_seed(d)
    z = 0.5 * torch.randn(131, d)
    u_ref, g_ref = R.value_and_grad(R.neutra_potential(oflow, make_potential_ref(pot, (d,))), z)
    lt = LatentTarget(product_target(pot, d), flow)
    u, g = lt.value_and_grad(z.cuda())
    np.testing.assert_allclose(u.cpu().numpy(), u_ref.numpy(), rtol=1e-4, atol=1e-5 * max(1.0, float(u_ref.abs().max())))
    np.testing.assert_allclose(g.cpu().numpy(), g_ref.numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(g_ref.abs().max())))
    np.testing.assert_allclose(lt.value(z.cuda()).cpu().numpy(), u_ref.numpy(), rtol=1e-4, atol=1e-5 * max(1.0, float(u_ref.abs().max())))


@pytest.mark.parametrize("kind", ["hmc", "mh"])
def test_neutra_deep_flow_against_oracle(kind):
    from gpu_util import product_target
    from nfmc_b200.records import HMCKernel, HMCParameters, MHKernel, MHParameters, NeuTraKernel, NeuTraParameters
    from nfmc_b200.samplers import NeuTraHMC, NeuTraMH
    d, n, T, L, tau = 37, 257, 3, 5, 0.05
    oflow, flow = _pair(d, 3, dict(n_layers=3, n_hidden=20), perturb=0.05)
    torch.manual_seed(3)
    z0 = 0.5 * torch.randn(n, d)
    normals, uniforms = torch.randn(T, n, d), torch.rand(T, n)
    tgt_ref = make_potential_ref("g0", (d,))
    if kind == "hmc":
        imd = torch.ones(d)
        run = R.run_neutra_hmc(z0, tgt_ref, oflow, T, R.TapeDraws(list(normals), list(uniforms)), tau, imd, n_leapfrog=L, trace=True)
        s = NeuTraHMC((d,), product_target("g0", d), HMCKernel(event_size=d, step_size=tau, n_leapfrog_steps=L), HMCParameters(),
                      NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    else:
        imd = torch.full((d,), 0.05)
        run = R.run_neutra_mh(z0, tgt_ref, oflow, T, R.TapeDraws(list(normals), list(uniforms)), imd, trace=True)
        s = NeuTraMH((d,), product_target("g0", d), MHKernel(event_size=d, inv_mass_diag=imd), MHParameters(),
                     NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    out = s.sample(z0, show_progress=False, normals=normals, uniforms=uniforms)
    lr = torch.stack(run.trace["log_ratio"])
    clear = (lr - torch.log(uniforms)).abs().min(dim=0).values > 1e-3 * (1.0 + lr.abs().max(dim=0).values)
    assert clear.float().mean() > 0.9
    np.testing.assert_allclose(out.samples[:, clear].numpy(), run.samples[:, clear].numpy(), rtol=1e-4,
                               atol=5e-5 * max(1.0, float(run.samples.abs().max())))
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~clear).sum()) * T
    assert out.statistics.n_target_calls == run.n_target_calls


@pytest.mark.parametrize("d,n_layers,ck", DEEP[:3])
def test_flow_sample_row_tile(monkeypatch, d, n_layers, ck):
    """Flow.sample(return_log_prob=True) (jump.py:205, imh.py:221): injected base draw against the oracle; the Philox draw is the
    one the per-chain kernel makes (same seed -> same x), and log q is consistent with Flow.log_prob."""
    oflow, flow = _pair(d, n_layers, ck)
    n = 515
    z = torch.randn(n, d, generator=torch.Generator().manual_seed(d))
    with torch.no_grad():
        x_ref, ld_ref = oflow.bijection.inverse(z)
        lq_ref = oflow.log_prob(x_ref)
    x, lq = flow.sample((n,), return_log_prob=True, z=z.cuda())
    np.testing.assert_allclose(x.cpu().numpy(), x_ref.numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(x_ref.abs().max())))
    np.testing.assert_allclose(lq.cpu().numpy(), lq_ref.numpy(), rtol=1e-4, atol=1e-4 * max(1.0, float(lq_ref.abs().max())))
    xs, lqs = flow.sample((n,), return_log_prob=True, seed=123)
    np.testing.assert_allclose(flow.log_prob(xs).cpu().numpy(), lqs.cpu().numpy(), rtol=1e-4, atol=1e-4 * max(1.0, float(lqs.abs().max())))
    # the Philox base draw is nfmc_rng_fill's stream 1 in LOGICAL order: feeding those numbers back reproduces the sample bitwise
    from nfmc_b200 import _native as N
    dev = torch.device("cuda")
    zf = torch.empty(n, d, device=dev)
    rng = N.rng_desc(123, 0, None, None)
    N.check(N.lib().nfmc_rng_fill(C.byref(rng), 1, 0, d, n, 1, N.ptr(zf), None, N.stream_ptr(dev)))
    xi, lqi = flow.sample((n,), return_log_prob=True, z=zf)
    assert torch.equal(xi, xs) and torch.equal(lqi, lqs)
    if n_layers % 2 == 0:
        # ... and, for an even number of reversals, the draw of the per-chain kernel (which draws in PHYSICAL order: with an
        # odd number of reversals its logical z is the mirror image of the same numbers -- an equally valid N(0, I) draw)
        monkeypatch.setenv("NFMC_B200_NO_ROW_TILE", "1")
        xg, lqg = flow.sample((n,), return_log_prob=True, seed=123)
        np.testing.assert_allclose(xs.cpu().numpy(), xg.cpu().numpy(), rtol=1e-4, atol=2e-5 * max(1.0, float(xg.abs().max())))
        np.testing.assert_allclose(lqs.cpu().numpy(), lqg.cpu().numpy(), rtol=1e-4, atol=1e-4 * max(1.0, float(lqg.abs().max())))


def test_neutra_hmc_tensor_core_flow_outside_the_tc_neutra_kernels():
    """d = 102, H = 64: the flow passes are tensor-core eligible, the tensor-core NeuTra kernels are not (d % 4 != 0).  NeuTra
    then runs the row-tile fp32 kernels (not the per-chain generic conditioner): fp32 agreement with the oracle."""
    from gpu_util import product_flow_from_oracle, product_target
    from nfmc_b200.records import HMCKernel, HMCParameters, NeuTraKernel, NeuTraParameters
    from nfmc_b200.samplers import NeuTraHMC
    d, n, T, L, tau = 102, 200, 2, 4, 0.03
    oflow = make_flow((d,), n_layers=2, conditioner_kwargs=dict(n_layers=2, n_hidden=64), perturb=0.03, seed=8)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="auto")
    bij = flow.bijection
    assert bij.uses_tensor_cores() and not bij.uses_tensor_cores_for_neutra(n) and bij.row_tile_supported() and not bij.uses_row_tile_pass()
    torch.manual_seed(4)
    z0 = 0.5 * torch.randn(n, d)
    normals, uniforms = torch.randn(T, n, d), torch.rand(T, n)
    run = R.run_neutra_hmc(z0, make_potential_ref("g0", (d,)), oflow, T, R.TapeDraws(list(normals), list(uniforms)), tau, torch.ones(d),
                           n_leapfrog=L, trace=True)
    s = NeuTraHMC((d,), product_target("g0", d), HMCKernel(event_size=d, step_size=tau, n_leapfrog_steps=L), HMCParameters(),
                  NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    out = s.sample(z0, show_progress=False, normals=normals, uniforms=uniforms)
    lr = torch.stack(run.trace["log_ratio"])
    clear = (lr - torch.log(uniforms)).abs().min(dim=0).values > 1e-3 * (1.0 + lr.abs().max(dim=0).values)
    assert clear.float().mean() > 0.9
    np.testing.assert_allclose(out.samples[:, clear].numpy(), run.samples[:, clear].numpy(), rtol=1e-4,
                               atol=5e-5 * max(1.0, float(run.samples.abs().max())))


def test_callable_target_with_deep_flow_equals_builtin_target():
    """A callable target (external path: autograd for U, flow proposal by nfmc_flow_wide_sample, log q by nfmc_flow_wide_log_prob)
    and the same function as a built-in potential (nfmc_jump_step_wide) make the same IMH run from the same Philox seed."""
    from gpu_util import product_target
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import FixedIMH
    d, n, T = 37, 500, 4
    oflow, flow = _pair(d, 3, dict(n_layers=3, n_hidden=20), perturb=0.03)
    x0 = torch.randn(n, d, generator=torch.Generator().manual_seed(6))
    outs = []
    for callable_target in (False, True):
        s = FixedIMH((d,), product_target("g0", d, callable_target=callable_target), IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
        s.seed = 31
        outs.append(s.sample(x0, show_progress=False))
    a, b = outs
    same = (a.samples - b.samples).abs().amax(dim=(0, 2)) < 1e-4
    assert same.float().mean() > 0.97              # a decision within rounding of its threshold may flip
    assert abs(a.statistics.n_accepted_trajectories - b.statistics.n_accepted_trajectories) <= int((~same).sum()) * T
