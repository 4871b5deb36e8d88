"""Edge cases and BASELINE-size properties of the hot path (through the C ABI).

* full-size, size-independent properties of the flow operators (2^20 chains, d = 100 -- config CT's shape; the oracle
  cannot finish that in seconds): forward -> inverse round trip, log-det antisymmetry, ``log_prob = log N(z) + log-det``,
  row independence (a row's result does not depend on which tile / CTA / slab it lands in);
* a single chain, the smallest event sizes, and the empty batch (the reference divides by zero there:
  ``ZeroDivisionError`` out of its acceptance-rate bookkeeping; ours refuses the call up front).
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle.samplers_ref as R                                   # noqa: E402  (checker only)
from oracle.potentials_ref import make_potential_ref             # noqa: E402
from oracle.realnvp_ref import make_flow                          # noqa: E402


def _flow(d, wide, seed=3):
    from gpu_util import product_flow_from_oracle
    kw = dict(conditioner_kwargs=dict(n_layers=2, n_hidden=64)) if wide else {}
    oflow = make_flow((d,), n_layers=2, perturb=0.05, seed=seed, **kw)
    return oflow, product_flow_from_oracle(oflow, conditioner_dtype="bf16" if wide else "fp32")


@pytest.mark.parametrize("wide", [False, True], ids=["fp32-cuda-cores", "bf16-tcgen05"])
def test_flow_full_size_properties(wide):
    d, n = 100, 1 << 20
    oflow, flow = _flow(d, wide)
    assert flow.bijection.uses_tensor_cores() == wide
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(n, d, device="cuda", generator=g) * 1.5
    z, ld_f = flow.bijection.forward(x)
    xr, ld_i = flow.bijection.inverse(z)
    # the two directions evaluate the conditioner on the same half, which differs between them only by the rounding of the
    # elementwise layers in between: a few fp32 ulps on the CUDA-core path; on the bf16 path such an ulp can flip the bf16
    # rounding of a conditioner input (2^-9 relative), so the round trip closes to the bf16 tolerance of the north star
    rt = 1e-2 if wide else 1e-4
    assert float((xr - x).abs().max()) < (2e-3 if wide else 2e-5) * float(x.abs().max())
    assert float((ld_f + ld_i).abs().max()) < rt * (1.0 + float(ld_f.abs().max()))
    lp = flow.log_prob(x)
    want = -0.5 * (z.double() ** 2).sum(dim=1) - 0.5 * d * math.log(2 * math.pi) + ld_f.double()
    assert float((lp.double() - want).abs().max()) < (2e-2 if wide else 1e-4) * (1.0 + float(want.abs().max()))
    assert bool(torch.isfinite(lp).all())
    # row independence: 1000 rows taken from everywhere in the batch and passed on their own give the same numbers
    idx = torch.randint(0, n, (1000,), generator=torch.Generator().manual_seed(1)).cuda()
    z2, ld2 = flow.bijection.forward(x[idx].contiguous())
    assert torch.equal(z2, z[idx]) and torch.equal(ld2, ld_f[idx])
    # and the oracle agrees on those rows
    with torch.no_grad():
        zo, ldo = oflow.bijection.forward(x[idx].cpu())
    np.testing.assert_allclose(z2.cpu().numpy(), zo.numpy(), rtol=rt, atol=rt * float(zo.abs().max()))
    np.testing.assert_allclose(ld2.cpu().numpy(), ldo.numpy(), rtol=rt, atol=rt * (1.0 + float(ldo.abs().max())))


@pytest.mark.parametrize("kind", ["mala", "hmc"])
@pytest.mark.parametrize("d", [1, 2, 3])
def test_single_chain_smallest_events(kind, d):
    """n = 1 and d in {1, 2, 3} against the oracle with injected draws (one partially filled warp, one lane group)."""
    from gpu_util import product_target, run_local_injected
    from nfmc_b200.records import LangevinKernel, LangevinParameters, HMCKernel, HMCParameters
    from nfmc_b200.samplers import MALA, HMC
    K, n = 6, 1
    torch.manual_seed(10 * d + len(kind))
    x0 = 0.5 * torch.randn(n, d)
    normals, uniforms = torch.randn(K, n, d), torch.rand(K, n)
    ref_t = make_potential_ref("g0", (d,))
    imd = torch.ones(d)
    if kind == "mala":
        run = R.run_mala(x0, ref_t, 0.3, imd, K, R.TapeDraws(list(normals), list(uniforms)), trace=True)
        s = MALA((d,), product_target("g0", d), LangevinKernel(event_size=d, step_size=0.3), LangevinParameters())
    else:
        run = R.run_hmc(x0, ref_t, 0.2, imd, 4, K, R.TapeDraws(list(normals), list(uniforms)), trace=True)
        s = HMC((d,), product_target("g0", d), HMCKernel(event_size=d, step_size=0.2, n_leapfrog_steps=4), HMCParameters())
    samples, ses, (sx, sx2, cnt) = run_local_injected(s, x0, normals, uniforms)
    lr = torch.stack(run.trace["log_ratio"])
    if float((lr - torch.log(uniforms)).abs().min()) > 1e-3:
        np.testing.assert_allclose(samples.numpy(), run.samples.numpy(), rtol=1e-4, atol=2e-5)
        assert cnt[0] == run.n_accepted
    assert cnt[1] == run.n_attempted == K


@pytest.mark.parametrize("strategy", ["mala", "hmc", "jump_mala", "imh", "neutra_hmc"])
def test_empty_batch_is_refused(strategy):
    """The reference fails on zero chains (ZeroDivisionError in its acceptance-rate bookkeeping, base.py:90-97); the CUDA
    path must not launch on an empty grid or return NaN statistics silently."""
    import nfmc_b200
    from nfmc_b200.potentials import StandardGaussian
    with pytest.raises((ValueError, ZeroDivisionError)):
        nfmc_b200.sample(StandardGaussian((4,)), event_shape=(4,), strategy=strategy, n_chains=0, n_iterations=2,
                         device=torch.device("cuda"), show_progress=False)


@pytest.mark.parametrize("strategy", ["jump_mala", "imh", "neutra_hmc"])
def test_single_chain_flow_strategies(strategy):
    import nfmc_b200
    from nfmc_b200.potentials import StandardGaussian
    extra = {"inner_param_kwargs": {"n_iterations": 2}} if strategy.startswith("jump") else {}
    out = nfmc_b200.sample(StandardGaussian((2,)), event_shape=(2,), strategy=strategy, n_chains=1, n_iterations=3,
                           device=torch.device("cuda"), show_progress=False, **extra)
    assert tuple(out.samples.shape)[-2:] == (1, 2) and bool(torch.isfinite(out.samples).all())
    assert 0.0 <= out.statistics.acceptance_rate <= 1.0


def test_one_dimensional_flow_is_refused_like_the_reference():
    """RealNVP on a 1-dimensional event has an empty source half; torchflows' hidden-width rule takes log10(0) there
    (the oracle restatement raises ValueError: math domain error)."""
    from nfmc_b200.flow import create_flow_object
    with pytest.raises(ValueError):
        make_flow((1,), n_layers=2)
    with pytest.raises(ValueError):
        create_flow_object("realnvp", (1,))
