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


@pytest.mark.timeout(300)
@pytest.mark.parametrize("d,Lc,H,pot,n", [(100, 2, 64, "gm", 3000), (100, 3, 32, "rb", 1111), (64, 4, 256, "fn", 700),
                                          (100, 4, 256, "g1", 70000), (32, 2, 16, "g0", 257)])
def test_tc_fused_jump_equals_composed_launches(monkeypatch, d, Lc, H, pot, n):
    """The one-kernel tensor-core jump (csrc/tc_jump.cu) against the same jump composed from separate launches (TC log_prob,
    Philox fill, TC inverse, CUDA-core accept kernel): same Philox numbers, same pass arithmetic -- so the proposals are
    identical and the states differ only where the two evaluations of log alpha (different summation orders) straddle
    log u.  Counters, moments, the log q cache and the sample sink follow."""
    import ctypes as C
    from gpu_util import product_flow_from_oracle, product_target
    from nfmc_b200 import _native as N
    oflow = make_flow((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=0.03, seed=7)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="bf16")
    dev = torch.device("cuda")
    tgt = product_target(pot, d)
    pd, keep = tgt.descriptor(dev)
    fd, keep2 = flow.bijection.tc_descriptor(dev)
    torch.manual_seed(1)
    x0 = (0.6 * torch.randn(n, d)).to(dev)
    res = {}
    for mode in ("fused", "composed"):
        if mode == "composed":
            monkeypatch.setenv("NFMC_TC_NO_FUSED_JUMP", "1")
        else:
            monkeypatch.delenv("NFMC_TC_NO_FUSED_JUMP", raising=False)
        out = {}
        for use_cache in (False, True):
            x = x0.clone()
            mom = torch.zeros(2 * d, device=dev, dtype=torch.float64)
            cnt = torch.zeros(8, device=dev, dtype=torch.int64)
            st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
            sink_buf = torch.full((n, d), float("nan"), device=dev)
            sink = N.SinkDesc(sink_buf.data_ptr(), 0, 1)
            cache = flow.log_prob(x).contiguous() if use_cache else None
            nb = N.lib().nfmc_jump_tc_workspace_bytes(d, n)
            ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            # with an odd number of couplings the composed path treats the Philox fill as the LOGICAL latent and the fused
            # kernel (like the CUDA-core kernels, flow_args.cuh: draw_base) as the physical one -- the same distribution, but
            # not the same numbers: inject the draws there
            gen = torch.Generator().manual_seed(5)
            for step in range(2):
                zin = torch.randn(n, d, generator=gen).to(dev) if Lc % 2 else None
                uin = torch.rand(n, generator=gen).to(dev) if Lc % 2 else None
                rng = N.rng_desc(1234, step, zin, uin)
                N.check(N.lib().nfmc_jump_step_tc(C.byref(pd), C.byref(fd), N.ptr(x), None if cache is None else N.ptr(cache), 0, n, 1,
                                                  C.byref(rng), 17, C.byref(st), C.byref(sink), N.ptr(ws), nb, N.stream_ptr(dev)))
            torch.cuda.synchronize()
            out[use_cache] = (x.cpu(), mom.cpu(), cnt.cpu(), sink_buf.cpu(), None if cache is None else cache.cpu())
        res[mode] = out
    for use_cache in (False, True):
        xf, mf, cf, sf, qf = res["fused"][use_cache]
        xc, mc, cc, sc, qc = res["composed"][use_cache]
        same = (xf == xc).all(dim=1)
        assert float(same.float().mean()) > 0.995, float(same.float().mean())          # decisions agree except near-ties
        assert bool(torch.isfinite(xf).all())
        assert torch.equal(sf, xf)                                                       # the sink holds the post-jump state
        assert int(cf[1]) == 2 * n and abs(int(cf[0]) - int(cc[0])) <= int((~same).sum()) * 2
        assert int(cf[0]) < 2 * n and (int(cf[0]) > 0) == (int(cc[0]) > 0)
        # moments: identical up to the rows whose decision differs
        tol = 1e-6 * n + 2.0 * float((~same).sum()) * float(xc.abs().max()) ** 2 + 1e-3
        assert float((mf - mc).abs().max()) <= tol * max(1.0, float(xc.abs().max())), float((mf - mc).abs().max())
        if use_cache:
            assert float((qf[same] - qc[same]).abs().max()) < 1e-3 * (1 + float(qc.abs().max()))
