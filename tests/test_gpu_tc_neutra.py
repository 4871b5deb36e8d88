"""B200: NeuTra on the tensor cores (csrc/tc_neutra.cu) -- the latent potential U~(z) = U(T^-1 z) - log|det dT^-1/dz|
(nfmc/neutra.py:58-68), its gradient (conditioner forward AND input-VJP on tcgen05) and the HMC step built on them
(mcmc/hmc.py:51-58,96-126) -- against the fp32 CPU oracle.  Tolerance: the bf16-conditioner bound of the north star
(rtol 1e-2) on values; gradients to 2e-2 of the largest gradient entry (bf16 cotangents)."""
import ctypes as C

import pytest
import torch

from oracle import samplers_ref as R
from oracle.potentials_ref import make_potential_ref
from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu

CASES = [(100, 2, 64, "fn", 300), (100, 4, 256, "g1", 1000), (64, 3, 32, "rb", 129), (16, 1, 32, "g0", 5), (32, 2, 128, "gm", 777),
         (100, 3, 96, "g0", 260), (120, 2, 256, "fn", 131),
         (100, 2, 5, "fn", 300), (100, 2, 40, "g0", 200)]      # default conditioner (H = 5) and H = 40: padded to 32 / 64


def _setup(d, Lc, H, pot, perturb=0.05):
    from gpu_util import product_flow_from_oracle, product_target
    oflow = make_flow((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=perturb, seed=H + d)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="bf16")
    assert flow.bijection.uses_tensor_cores_for_neutra()
    return oflow, flow, make_potential_ref(pot, (d,)), product_target(pot, d)


def _value_grad_tc(flow, tgt, z):
    from nfmc_b200 import _native as N
    dev = torch.device("cuda")
    n, d = z.shape
    zd = z.to(dev).contiguous()
    x = torch.empty(n, d, device=dev)
    ld = torch.empty(n, device=dev)
    u = torch.empty(n, device=dev)
    g = torch.full((n, d), float("nan"), device=dev)
    pd, k1 = tgt.descriptor(dev)
    td, k2 = flow.bijection.tc_descriptor(dev)
    bt = flow.bijection.tc_transposed(dev)
    N.check(N.lib().nfmc_neutra_potential_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(zd), N.ptr(x), N.ptr(ld), N.ptr(u),
                                             N.ptr(g), n, N.stream_ptr(dev)))
    torch.cuda.synchronize()
    return u.cpu(), g.cpu(), x.cpu(), ld.cpu()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("d,Lc,H,pot,n", CASES)
def test_tc_neutra_potential_and_gradient(d, Lc, H, pot, n):
    oflow, flow, tgt_ref, tgt = _setup(d, Lc, H, pot)
    torch.manual_seed(d + n)
    z = 0.5 * torch.randn(n, d)
    u_ref, g_ref = R.value_and_grad(R.neutra_potential(oflow, tgt_ref), z)
    u, g, x, ld = _value_grad_tc(flow, tgt, z)
    assert bool(torch.isfinite(g).all()) and bool(torch.isfinite(u).all())
    assert float((u - u_ref).abs().max()) < 1e-2 * (1.0 + float(u_ref.abs().max())), float((u - u_ref).abs().max())
    gmax = float(g_ref.abs().max())
    err = float((g - g_ref).abs().max())
    assert err < 2e-2 * max(1.0, gmax), (err, gmax)
    # and in the mean much tighter than the worst entry
    assert float((g - g_ref).abs().mean()) < 3e-3 * max(1.0, float(g_ref.abs().mean()) * 10), float((g - g_ref).abs().mean())


@pytest.mark.timeout(300)
def test_tc_neutra_gradient_is_the_derivative_of_the_tc_value():
    """Directional finite differences of the kernel's OWN value (bf16 flow) agree with its gradient: the pair (U~, grad U~)
    is consistent, which is what makes the leapfrog efficient (exactness of HMC does not depend on it)."""
    d, Lc, H = 100, 4, 256
    oflow, flow, tgt_ref, tgt = _setup(d, Lc, H, "g1", perturb=0.03)
    torch.manual_seed(3)
    n = 256
    z = 0.4 * torch.randn(n, d)
    v = torch.randn(n, d)
    v = v / v.norm(dim=1, keepdim=True)
    u0, g, _, _ = _value_grad_tc(flow, tgt, z)
    # fp32 reference of the same directional derivative (the bf16 value is piecewise constant at the 1e-3 level, so the
    # finite difference is taken over a step that is large against the rounding but small against the curvature)
    _, g_ref = R.value_and_grad(R.neutra_potential(oflow, tgt_ref), z)
    dd_ref = (g_ref * v).sum(dim=1)
    dd = (g * v).sum(dim=1)
    h = 0.05
    up, _, _, _ = _value_grad_tc(flow, tgt, z + h * v)
    um, _, _, _ = _value_grad_tc(flow, tgt, z - h * v)
    fd = (up - um) / (2 * h)
    scale = float(dd_ref.abs().max())
    assert float((dd - dd_ref).abs().max()) < 2e-2 * scale, (float((dd - dd_ref).abs().max()), scale)
    # the finite difference carries the bf16 rounding of the VALUE (~1e-3 |U~| / h): loose bound on the worst chain, tight on the mean
    assert float((fd - dd).abs().max()) < 0.15 * scale, (float((fd - dd).abs().max()), scale)
    assert float((fd - dd).abs().mean()) < 2e-2 * scale, (float((fd - dd).abs().mean()), scale)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("d,Lc,H,pot,n,imd_scale", [(100, 2, 64, "g0", 1500, None), (64, 3, 32, "fn", 700, 0.5), (100, 4, 256, "g1", 2000, None)])
def test_tc_neutra_hmc_against_oracle(d, Lc, H, pot, n, imd_scale):
    """Two HMC steps of L = 5 leapfrogs with injected momenta / uniforms against oracle.run_neutra_hmc: proposals agree to the
    bf16 tolerance, decisions agree wherever the margin exceeds the bf16 error of the Hamiltonian, counters and moments follow."""
    from nfmc_b200 import _native as N
    oflow, flow, tgt_ref, tgt = _setup(d, Lc, H, pot, perturb=0.03)
    torch.manual_seed(5)
    T, L, tau = 2, 5, 0.05
    z0 = 0.5 * torch.randn(n, d)
    imd = torch.ones(d) if imd_scale is None else (imd_scale + torch.rand(d))
    draws = R.GlobalDraws(record=True)
    ref = R.run_neutra_hmc(z0, tgt_ref, oflow, T, draws, tau, imd, n_leapfrog=L, store=True, trace=True)
    nz = torch.stack(draws.normals).contiguous()
    un = torch.stack(draws.uniforms).contiguous()
    dev = torch.device("cuda")
    z = z0.to(dev).contiguous()
    mom = torch.zeros(2 * d, device=dev, dtype=torch.float64)
    cnt = torch.zeros(8, device=dev, dtype=torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    buf = torch.full((T, n, d), float("nan"), device=dev)
    sink = N.SinkDesc(buf.data_ptr(), 0, 1)
    pd, k1 = tgt.descriptor(dev)
    td, k2 = flow.bijection.tc_descriptor(dev)
    bt = flow.bijection.tc_transposed(dev)
    nb = N.lib().nfmc_neutra_tc_workspace_bytes(d, n)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    nzd, und = nz.to(dev), un.to(dev)
    imd_d = None if imd_scale is None else imd.to(dev)
    rng = N.rng_desc(0, 0, nzd, und)
    N.check(N.lib().nfmc_neutra_hmc_steps_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(z), n, T, tau, L, N.ptr(imd_d), 1,
                                             C.byref(rng), 0, C.byref(st), C.byref(sink), N.ptr(ws), nb, N.stream_ptr(dev)))
    torch.cuda.synchronize()
    got = buf.cpu()
    assert bool(torch.isfinite(got).all())
    same = ((got - ref.samples).abs().amax(dim=2) < 2e-2 * (1.0 + ref.samples.abs().amax(dim=2)))   # [T, n]
    frac = float(same.float().mean())
    assert frac > 0.97, frac                                   # decisions agree except near ties; proposals to bf16 tolerance
    acc = int(cnt[0])
    assert int(cnt[1]) == T * n
    assert abs(acc - ref.n_accepted) <= int((~same).sum()) + 2, (acc, ref.n_accepted)
    assert torch.equal(z.cpu(), got[-1])
    # moments of the stored states
    sx = got.double().sum(dim=(0, 1))
    assert float((mom[:d].cpu() - sx).abs().max()) < 1e-3 * (1.0 + float(sx.abs().max()))


@pytest.mark.timeout(600)
def test_tc_neutra_hmc_through_the_api_and_statistics():
    """neutra_hmc with a wide flow through nfmc_b200.sample: the tensor-core path is taken, counters follow the reference
    formulas (hmc.py:122-125), and with an identity-initialised flow the chain samples the target (moments of N(0, 1/2 I),
    SURVEY quirk Q9) in latent = data space."""
    import nfmc_b200
    from nfmc_b200 import potentials as P
    from nfmc_b200.samplers import NeuTraHMC
    d, n, T = 32, 4096, 60
    target = P.make_potential("g0", (d,))
    flow = 'realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}'
    s = nfmc_b200.create_sampler(target, event_shape=(d,), strategy="neutra_hmc", flow=flow, device="cuda",
                                 param_kwargs=dict(n_iterations=T, store_samples=False),
                                 inner_kernel_kwargs=dict(step_size=0.15, n_leapfrog_steps=8))
    assert isinstance(s, NeuTraHMC) and s.kernel.flow.bijection.uses_tensor_cores_for_neutra()
    torch.manual_seed(0)
    out = s.sample(torch.randn(n, d) * 0.7, show_progress=False)
    stats = out.statistics
    assert stats.n_attempted_trajectories == T * n
    assert stats.n_target_gradient_calls == 2 * 8 * n * T and stats.n_target_calls == (2 * 8 + 2) * n * T
    assert 0.6 < stats.acceptance_rate <= 1.0, stats.acceptance_rate
    mean, var = out.mean, out.variance
    assert float(mean.abs().max()) < 0.05, float(mean.abs().max())
    assert float((var - 0.5).abs().max()) < 0.06, float((var - 0.5).abs().max())


def test_tc_neutra_rejects_ineligible_shapes():
    from nfmc_b200 import _native as N
    from nfmc_b200.flow import RealNVP
    assert N.lib().nfmc_neutra_tc_transposed_bytes(100, 2, 48) == -1          # hidden not a multiple of 32
    assert N.lib().nfmc_neutra_tc_transposed_bytes(102, 2, 64) == -1          # d % 4 != 0
    assert N.lib().nfmc_neutra_tc_transposed_bytes(128, 2, 256) == -1         # shared-memory plan too large
    assert N.lib().nfmc_neutra_tc_transposed_bytes(100, 4, 256) > 0
    # the packed images pad the hidden width to a multiple of 32, so every tensor-core flow with d % 4 == 0 is eligible
    assert RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=40)).uses_tensor_cores_for_neutra()
    assert not RealNVP((102,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_tensor_cores_for_neutra()
    assert not RealNVP((100,)).uses_tensor_cores_for_neutra()                 # default conditioner: fp32 unless opted in
    assert RealNVP((100,), conditioner_dtype="bf16").uses_tensor_cores_for_neutra()
    # 'auto': a default conditioner goes to the tensor cores only for large batches (where it is 1.5x faster); 'fp32' never
    assert RealNVP((100,)).uses_tensor_cores_for_neutra(1 << 18) and not RealNVP((100,)).uses_tensor_cores_for_neutra(1000)
    assert not RealNVP((100,), conditioner_dtype="fp32").uses_tensor_cores_for_neutra(1 << 18)
