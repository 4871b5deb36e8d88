"""B200: the tcgen05 / TMEM conditioner path (csrc/cond_tc.cu) against the fp32 CPU oracle.
Tolerance: the bf16-conditioner bound of the north star, rtol 1e-2."""
import pytest
import torch

from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu


def _err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float(((a - b).abs() / (1.0 + b.abs())).max())


@pytest.mark.timeout(180)
@pytest.mark.parametrize("d,n_layers,H,n", [(100, 2, 5, 700), (26, 3, 40, 100), (100, 2, 64, 300), (100, 4, 256, 1000), (64, 3, 16, 129), (128, 2, 128, 77),
                                           (16, 1, 32, 5), (100, 2, 256, 20000)])
def test_tc_forward_inverse_logprob(d, n_layers, H, n):
    from gpu_util import product_flow_from_oracle
    torch.manual_seed(d + H)
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=0.05, seed=H)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="bf16")
    assert flow.bijection.uses_tensor_cores()
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, ld_ref = oflow.bijection.forward(x)
        xi_ref, ldi_ref = oflow.bijection.inverse(x)
        lp_ref = oflow.log_prob(x)
    z, ld = flow.bijection.forward(x.cuda())
    torch.cuda.synchronize()
    assert _err(z, z_ref) < 1e-2, _err(z, z_ref)
    assert float((ld.cpu() - ld_ref).abs().max()) < 1e-2 * (1 + float(ld_ref.abs().max())), float((ld.cpu() - ld_ref).abs().max())
    xi, ldi = flow.bijection.inverse(x.cuda())
    assert _err(xi, xi_ref) < 1e-2, _err(xi, xi_ref)
    assert float((ldi.cpu() - ldi_ref).abs().max()) < 1e-2 * (1 + float(ldi_ref.abs().max()))
    lp = flow.log_prob(x.cuda())
    assert float((lp.cpu() - lp_ref).abs().max()) < 1e-2 * (1 + float(lp_ref.abs().max()))
    # the fp32 CUDA-core path on the same flow agrees with the tensor-core path to the same tolerance
    flow32 = product_flow_from_oracle(oflow, conditioner_dtype="fp32")
    z32, ld32 = flow32.bijection.forward(x.cuda())
    assert _err(z, z32) < 1e-2


def test_tc_rejects_unsupported_shapes():
    from nfmc_b200.flow import Flow, RealNVP
    with pytest.raises(ValueError):
        RealNVP((100,), conditioner_kwargs=dict(n_layers=3, n_hidden=32), conditioner_dtype="bf16")
    with pytest.raises(ValueError):
        RealNVP((101,), conditioner_kwargs=dict(n_layers=2, n_hidden=64), conditioner_dtype="bf16")
    assert not RealNVP((100,)).uses_tensor_cores()           # default H = 5: CUDA-core path
    assert RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_tensor_cores()


@pytest.mark.timeout(180)
def test_tc_jump_and_imh_against_oracle():
    """jump_mala and fixed IMH with a wide flow: the flow passes run on the tensor cores (bf16), the rest in fp32.
    Proposals agree to the bf16 tolerance; decisions agree wherever the margin exceeds the bf16 error of log alpha."""
    from gpu_util import product_flow_from_oracle, product_target
    from oracle import samplers_ref as R
    from oracle.potentials_ref import make_potential_ref
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import FixedIMH
    d, n, T, H = 100, 4096, 3, 64
    oflow = make_flow((d,), n_layers=2, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=0.02, seed=3)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="bf16")
    torch.manual_seed(9)
    x0 = 0.7 * torch.randn(n, d)
    z = torch.randn(T, n, d)
    u = torch.rand(T, n)
    run = R.run_fixed_imh(x0, make_potential_ref("g0", (d,)), oflow, T, R.TapeDraws(list(z), list(u)), trace=True)
    s = FixedIMH((d,), product_target("g0", d), IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    out = s.sample(x0, show_progress=False, z=z, uniforms=u)
    la = torch.stack(run.trace["log_alpha"])
    margin = (la - torch.log(u)).abs().min(dim=0).values
    clear = margin > 0.05 * (1.0 + la.abs().max(dim=0).values)
    assert clear.float().mean() > 0.6
    err = ((out.samples[:, clear] - run.samples[:, clear]).abs() / (1 + run.samples[:, clear].abs())).max()
    assert float(err) < 2e-2, float(err)
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~clear).sum()) * T
    assert out.statistics.n_attempted_trajectories == n * T and out.statistics.n_target_calls == 2 * n * T
    # every stored state is either the previous state or that iteration's proposal (bookkeeping is exact)
    xp = torch.stack(run.trace["x_prime"])
    prev = x0
    for k in range(T):
        cur = out.samples[k]
        same = (cur - prev).abs().amax(dim=1) == 0
        prop = ((cur - xp[k]).abs() / (1 + xp[k].abs())).amax(dim=1) < 2e-2
        assert bool((same | prop).all())
        prev = cur


@pytest.mark.timeout(180)
def test_tc_jump_mala_through_api():
    import nfmc_b200
    from nfmc_b200.potentials import StandardGaussian
    torch.manual_seed(0)
    out = nfmc_b200.sample(StandardGaussian((100,)), strategy="jump_mala", n_chains=2048, n_iterations=3, show_progress=False,
                           flow='realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}',
                           inner_param_kwargs=dict(n_iterations=5))
    assert out.kernel.flow.bijection.uses_tensor_cores()
    assert out.samples.shape == (3 * 6, 2048, 100) and bool(torch.isfinite(out.samples).all())
    st = out.statistics
    assert st.n_attempted_jumps == 3 * 2048 and st.n_attempted_trajectories == 15 * 2048
    # identity flow at initialisation: proposals are N(0, I) draws against N(0, I/2): some are accepted
    assert 0 < st.n_accepted_jumps < st.n_attempted_jumps


def test_auto_dtype_routes_mid_width_conditioners_to_tensor_cores():
    """H in 9..15 (2 linear layers, even d <= 128) is padded to 16 and runs on tcgen05 under conditioner_dtype='auto';
    the result matches the fp32 CUDA-core path to bf16 tolerance."""
    from nfmc_b200.flow import Flow, RealNVP
    d, n = 64, 513
    torch.manual_seed(5)
    ck = dict(n_layers=2, n_hidden=12)
    f_auto = Flow(RealNVP((d,), n_layers=2, conditioner_kwargs=ck))
    with torch.no_grad():
        for p in f_auto.parameters():
            p.add_(0.1 * torch.randn_like(p))
    f_fp32 = Flow(RealNVP((d,), n_layers=2, conditioner_kwargs=ck, conditioner_dtype="fp32"))
    f_fp32.load_state_dict(f_auto.state_dict())
    f_auto, f_fp32 = f_auto.to("cuda"), f_fp32.to("cuda")
    assert f_auto.bijection.uses_tensor_cores() and not f_fp32.bijection.uses_tensor_cores()
    x = torch.randn(n, d, device="cuda")
    za, la = f_auto.bijection.forward(x)
    zf, lf = f_fp32.bijection.forward(x)
    assert float((za - zf).abs().max()) < 2e-2 * max(1.0, float(zf.abs().max()))
    assert float((la - lf).abs().max()) < 5e-2 * max(1.0, float(lf.abs().max()))
