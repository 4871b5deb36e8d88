"""B200: the CUDA path (called through the C ABI via nfmc_b200) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star): flow outputs, log-dets, potentials and proposals within rtol 1e-4 (fp32);
accept / reject decisions agree except for ties within tolerance of the threshold.
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from golden_util import load_case, tape, oracle_flow, oracle_target, CASES
from oracle import samplers_ref as R
from oracle.philox_ref import step_noise
from oracle.potentials_ref import make_potential_ref
from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def close(a, b, rtol=RTOL, atol=1e-5):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"max err {float(err.max()):.3e}, worst tol ratio {float((err / tol).max()):.2f}"


# ------------------------------------------------------------------------------------------------------------
# potentials
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["g0", "g1", "fn", "rb", "gm"])
@pytest.mark.parametrize("d", [2, 6, 26, 100, 1000])
def test_potential_value_and_grad(name, d):
    from gpu_util import product_target
    torch.manual_seed(d)
    n = 37
    x = 0.7 * torch.randn(n, d)
    ref = make_potential_ref(name, (d,))
    u_ref, g_ref = R.value_and_grad(ref, x)
    u, g = product_target(name, d).value_and_grad(x.cuda())
    scale = 1e-5 * max(1.0, float(u_ref.abs().max()))
    close(u, u_ref, atol=scale)
    close(g, g_ref, atol=1e-5 * max(1.0, float(g_ref.abs().max())))


def test_potential_odd_dims():
    from gpu_util import product_target
    for name in ["g0", "g1", "fn", "gm"]:
        for d in [3, 7, 25, 101]:
            x = torch.randn(11, d)
            u_ref, g_ref = R.value_and_grad(make_potential_ref(name, (d,)), x)
            u, g = product_target(name, d).value_and_grad(x.cuda())
            close(u, u_ref, atol=1e-5 * max(1.0, float(u_ref.abs().max())))
            close(g, g_ref, atol=1e-5 * max(1.0, float(g_ref.abs().max())))


# ------------------------------------------------------------------------------------------------------------
# RealNVP
# ------------------------------------------------------------------------------------------------------------
FLOW_CASES = [
    (6, 2, None), (7, 3, dict(n_layers=3, n_hidden=6)), (8, 1, dict(n_layers=1)), (25, 2, None), (100, 2, None),
    (100, 4, dict(n_layers=2, n_hidden=64)), (101, 3, dict(n_layers=4, n_hidden=17)), (64, 2, dict(n_layers=2, n_hidden=32)),
    (1000, 2, None), (9, 3, dict(n_layers=1)), (1023, 3, None), (1024, 2, None), (2, 2, None), (3, 1, None), (33, 2, None),
    (57, 3, dict(n_layers=2, n_hidden=7)), (209, 2, None),
]


@pytest.mark.parametrize("d,n_layers,ck", FLOW_CASES)
def test_realnvp_forward_inverse_logprob(d, n_layers, ck):
    from gpu_util import product_flow_from_oracle
    torch.manual_seed(d + n_layers)
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=0.1, seed=d)
    flow = product_flow_from_oracle(oflow)
    n = 45
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, ld_ref = oflow.bijection.forward(x)
        xi_ref, ldi_ref = oflow.bijection.inverse(x)
        lp_ref = oflow.log_prob(x)
    z, ld = flow.bijection.forward(x.cuda())
    close(z, z_ref)
    close(ld, ld_ref, atol=1e-4)
    xi, ldi = flow.bijection.inverse(x.cuda())
    close(xi, xi_ref, atol=1e-5 * max(1.0, float(xi_ref.abs().max())))
    close(ldi, ldi_ref, atol=1e-4)
    close(flow.log_prob(x.cuda()), lp_ref, atol=1e-4 * max(1.0, float(lp_ref.abs().max()) / 100))
    # round trip on the device
    xr, ldr = flow.bijection.inverse(z)
    close(xr, x, atol=2e-5 * max(1.0, float(x.abs().max())))
    close(ld + ldr, torch.zeros(n), atol=2e-4)


@pytest.mark.parametrize("d,n_layers,ck", [(6, 2, None), (7, 3, dict(n_layers=3, n_hidden=6)), (100, 2, None), (25, 3, None)])
def test_flow_sample_with_injected_base_draw(d, n_layers, ck):
    from gpu_util import product_flow_from_oracle
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=0.1, seed=d + 1)
    flow = product_flow_from_oracle(oflow)
    torch.manual_seed(3)
    z = torch.randn(33, d)
    x_ref, lq_ref = R.flow_sample_with_logq(oflow, 33, R.TapeDraws([z], []))
    x, lq = flow.sample(33, return_log_prob=True, z=z)
    close(x, x_ref, atol=1e-5 * max(1.0, float(x_ref.abs().max())))
    close(lq, lq_ref, atol=1e-4 * max(1.0, float(lq_ref.abs().max()) / 100))


# ------------------------------------------------------------------------------------------------------------
# golden cases (outputs of the unmodified reference) with the reference's own random draws injected
# ------------------------------------------------------------------------------------------------------------
def _golden_local(name, kind, ext=False):
    from gpu_util import product_target, run_local_injected
    from nfmc_b200.records import LangevinKernel, LangevinParameters, HMCKernel, HMCParameters
    from nfmc_b200.samplers import MALA, HMC
    g = load_case(name)
    d = g["x0"].shape[1]
    K = int(g["K"])
    imd = torch.from_numpy(g["imd"])
    tgt = product_target(g["pot"], d, callable_target=ext)
    if kind == "mala":
        s = MALA((d,), tgt, LangevinKernel(event_size=d, inv_mass_diag=imd, step_size=float(g["step"])), LangevinParameters())
    elif kind == "ula":
        from nfmc_b200.samplers import ULA
        s = ULA((d,), tgt, LangevinKernel(event_size=d, inv_mass_diag=imd, step_size=float(g["step"])), LangevinParameters())
    elif kind == "uhmc":
        from nfmc_b200.samplers import UHMC
        s = UHMC((d,), tgt, HMCKernel(event_size=d, inv_mass_diag=imd, step_size=float(g["step"]), n_leapfrog_steps=int(g["L"])),
                 HMCParameters())
    elif kind == "mh":
        from nfmc_b200.records import MHKernel, MHParameters
        from nfmc_b200.samplers import MH
        s = MH((d,), tgt, MHKernel(event_size=d, inv_mass_diag=imd), MHParameters())
    else:
        s = HMC((d,), tgt, HMCKernel(event_size=d, inv_mass_diag=imd, step_size=float(g["step"]), n_leapfrog_steps=int(g["L"])),
                HMCParameters())
    normals = torch.stack(g["normals"])
    uniforms = torch.stack(g["uniforms"]) if g["uniforms"] else None         # the unadjusted kernels draw no uniforms
    samples, ses, (sx, sx2, cnt) = run_local_injected(s, torch.from_numpy(g["x0"]), normals, uniforms)
    ref = torch.from_numpy(g["samples"])
    close(samples, ref, atol=1e-5 * max(1.0, float(ref.abs().max())))
    acc, att = int(g["counters"][0]), int(g["counters"][1])
    assert (cnt[0], cnt[1]) == (acc, att)
    n_seen = ref.shape[0] * ref.shape[1]
    close(sx / n_seen, g["mean"], atol=1e-5)
    close(sx2 / n_seen, g["second_moment"], atol=1e-5 * max(1.0, float(np.abs(g["second_moment"]).max())))


# ext = True: the target is a plain Python callable (the reference's own contract) -> external-target path
EXT = pytest.mark.parametrize("ext", [False, True], ids=["fused", "callable"])


@EXT
@pytest.mark.parametrize("name", ["mala_g0", "mala_fn"])
def test_golden_mala(name, ext):
    _golden_local(name, "mala", ext)


@EXT
@pytest.mark.parametrize("name", ["hmc_g1", "hmc_rb"])
def test_golden_hmc(name, ext):
    _golden_local(name, "hmc", ext)


@EXT
def test_golden_mh(ext):
    _golden_local("mh_gm", "mh", ext)


@EXT
def test_golden_ula(ext):
    _golden_local("ula_g0", "ula", ext)          # langevin.py:131-134


@EXT
def test_golden_uhmc(ext):
    _golden_local("uhmc_gm", "uhmc", ext)        # hmc.py:129-132


@EXT
@pytest.mark.parametrize("name,kind", [("mala_tune_g1", "mala"), ("hmc_tune_fn", "hmc")])
def test_golden_warmup_trajectory(name, kind, ext):
    """Warm-up (mcmc/base.py:39-54,142-161; tuning.py:15-41) with the reference's draws injected: the step size and the
    inverse-mass diagonal after EVERY iteration follow the reference's trajectory (the across-chain variance comes from
    fp64 sums on the device, the reference's from torch.var in fp32)."""
    from gpu_util import product_target
    from nfmc_b200.records import LangevinKernel, LangevinParameters, HMCKernel, HMCParameters
    from nfmc_b200.samplers import MALA, HMC, _DeviceTuner
    g = load_case(name)
    d, K = g["x0"].shape[1], int(g["K"])
    tgt = product_target(g["pot"], d, callable_target=ext)
    if kind == "mala":
        s = MALA((d,), tgt, LangevinKernel(event_size=d, step_size=float(g["step"])), LangevinParameters(n_iterations=K))
    else:
        s = HMC((d,), tgt, HMCKernel(event_size=d, step_size=float(g["step"]), n_leapfrog_steps=int(g["L"])), HMCParameters(n_iterations=K))
    s.params.tuning_mode()
    steps, imds = [], []
    orig = _DeviceTuner.update

    def recording(self, ses, kernel, params, n_steps):
        orig(self, ses, kernel, params, n_steps)
        steps.append(float(kernel.step_size))
        imds.append(kernel.inv_mass_diag.detach().cpu().clone())

    _DeviceTuner.update = recording
    try:
        out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, normals=torch.stack(g["normals"]), uniforms=torch.stack(g["uniforms"]))
    finally:
        _DeviceTuner.update = orig
    _check_output(out, g)
    np.testing.assert_allclose(np.array(steps), g["step_traj"], rtol=2e-5)
    np.testing.assert_allclose(torch.stack(imds).numpy(), g["imd_traj"], rtol=2e-5, atol=1e-7)
    assert not s.kernel.inv_mass_diag.is_cuda and abs(float(s.kernel.step_size) - float(g["step_traj"][-1])) < 2e-5 * float(g["step_traj"][-1])


def _check_output(out, g, jump=False):
    ref = torch.from_numpy(g["samples"])
    close(out.samples, ref, atol=2e-5 * max(1.0, float(ref.abs().max())))
    close(out.running_samples.last_sample, g["last"], atol=2e-5 * max(1.0, float(ref.abs().max())))
    close(out.mean, g["mean"], atol=2e-5)
    close(out.second_moment, g["second_moment"], atol=2e-5 * max(1.0, float(np.abs(g["second_moment"]).max())))
    acc, att, div, grads, calls, jacc, jatt = (int(v) for v in g["counters"])
    st = out.statistics
    assert (st.n_accepted_trajectories, st.n_attempted_trajectories, st.n_divergences) == (acc, att, div)
    assert (st.n_target_gradient_calls, st.n_target_calls) == (grads, calls)
    if jump:
        assert (st.n_accepted_jumps, st.n_attempted_jumps) == (jacc, jatt)


@EXT
@pytest.mark.parametrize("name,inner", [("jump_mala_g0", "mala"), ("jump_hmc_gm", "hmc"), ("jump_mala_g1_d100", "mala"),
                                        ("jump_hmc_g1_d100", "hmc"), ("jump_mala_gm_d1000", "mala")])
def test_golden_jump(name, inner, ext):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import (LangevinKernel, LangevinParameters, HMCKernel, HMCParameters, NFMCKernel,
                                   JumpNFMCParameters)
    from nfmc_b200.samplers import JumpMALA, JumpHMC
    g = load_case(name)
    n, d = g["x0"].shape
    T, K = int(g["T"]), int(g["K"])
    tgt = product_target(g["pot"], d, callable_target=ext)
    flow = product_flow_from_oracle(oracle_flow(g))
    if inner == "mala":
        s = JumpMALA((d,), tgt, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                     inner_kernel=LangevinKernel(event_size=d, step_size=float(g["step"])),
                     inner_params=LangevinParameters(n_iterations=K))
    else:
        s = JumpHMC((d,), tgt, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                    inner_kernel=HMCKernel(event_size=d, step_size=float(g["step"]), n_leapfrog_steps=int(g["L"])),
                    inner_params=HMCParameters(n_iterations=K))
    # reference draw order per outer iteration: K x [normal(n,d), uniform(n)], normal(n,d) [flow base], uniform(n)
    nn, uu = g["normals"], g["uniforms"]
    normals = torch.stack([torch.stack(nn[i * (K + 1): i * (K + 1) + K]) for i in range(T)])
    uniforms = torch.stack([torch.stack(uu[i * (K + 1): i * (K + 1) + K]) for i in range(T)])
    jump_z = torch.stack([nn[i * (K + 1) + K] for i in range(T)])
    jump_u = torch.stack([uu[i * (K + 1) + K] for i in range(T)])
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, normals=normals, uniforms=uniforms, jump_z=jump_z,
                   jump_uniforms=jump_u)
    _check_output(out, g, jump=True)


@EXT
def test_golden_adaptive_imh(ext):
    """AdaptiveIMH.sample (imh.py:102-181) with the refit switched off as in the fixture: log q recomputed every iteration,
    2 n GRADIENT calls booked per iteration (imh.py:146), every iteration stored."""
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import AdaptiveIMH
    g = load_case("adaptive_imh_rb")
    n, d = g["x0"].shape
    T = int(g["T"])
    s = AdaptiveIMH((d,), product_target(g["pot"], d, callable_target=ext), IMHKernel((d,), flow=product_flow_from_oracle(oracle_flow(g))),
                    IMHParameters(n_iterations=T))
    s.adapt = False
    # tape per iteration: uniform(n) for the accept test, then one scalar uniform for the refit decision (imh.py:152)
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, z=torch.stack(g["normals"]), uniforms=torch.stack(g["uniforms"][0::2]))
    _check_output(out, g)


@EXT
@pytest.mark.parametrize("name", ["imh_rb", "imh_rb_d100"])
def test_golden_fixed_imh(name, ext):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import FixedIMH
    g = load_case(name)
    n, d = g["x0"].shape
    T = int(g["T"])
    s = FixedIMH((d,), product_target(g["pot"], d, callable_target=ext), IMHKernel((d,), flow=product_flow_from_oracle(oracle_flow(g))),
                 IMHParameters(n_iterations=T))
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, z=torch.stack(g["normals"]), uniforms=torch.stack(g["uniforms"]))
    _check_output(out, g)


@EXT
@pytest.mark.parametrize("name", ["neutra_hmc_fn", "neutra_hmc_fn_d100"])
def test_golden_neutra_hmc(name, ext):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import HMCKernel, HMCParameters, NeuTraKernel, NeuTraParameters
    from nfmc_b200.samplers import NeuTraHMC
    g = load_case(name)
    n, d = g["x0"].shape
    T = int(g["T"])
    s = NeuTraHMC((d,), product_target(g["pot"], d, callable_target=ext), HMCKernel(event_size=d, step_size=float(g["step"]), n_leapfrog_steps=int(g["L"])),
                  HMCParameters(), NeuTraKernel((d,), flow=product_flow_from_oracle(oracle_flow(g))), NeuTraParameters(n_iterations=T))
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, normals=torch.stack(g["normals"]),
                   uniforms=torch.stack(g["uniforms"]))
    _check_output(out, g)


@EXT
def test_golden_neutra_mh(ext):
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import MHKernel, MHParameters, NeuTraKernel, NeuTraParameters
    from nfmc_b200.samplers import NeuTraMH
    g = load_case("neutra_mh_gm")
    n, d = g["x0"].shape
    T = int(g["T"])
    s = NeuTraMH((d,), product_target(g["pot"], d, callable_target=ext), MHKernel(event_size=d, inv_mass_diag=torch.from_numpy(g["imd"])), MHParameters(),
                 NeuTraKernel((d,), flow=product_flow_from_oracle(oracle_flow(g))), NeuTraParameters(n_iterations=T))
    out = s.sample(torch.from_numpy(g["x0"]), show_progress=False, normals=torch.stack(g["normals"]),
                   uniforms=torch.stack(g["uniforms"]))
    _check_output(out, g)


def test_neutra_mh_against_oracle_large():
    """d = 100 (production layout), ragged tile, Philox-free: injected noise; decisions agree except ties."""
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import MHKernel, MHParameters, NeuTraKernel, NeuTraParameters
    from nfmc_b200.samplers import NeuTraMH
    from oracle.realnvp_ref import make_flow
    d, n, T = 100, 301, 4
    torch.manual_seed(2)
    oflow = make_flow((d,), n_layers=2, perturb=0.05, seed=9)
    z0 = 0.5 * torch.randn(n, d)
    normals, uniforms = torch.randn(T, n, d), torch.rand(T, n)
    imd = torch.full((d,), 0.05)
    run = R.run_neutra_mh(z0, make_potential_ref("g1", (d,)), oflow, T, R.TapeDraws(list(normals), list(uniforms)), imd, trace=True)
    s = NeuTraMH((d,), product_target("g1", d), MHKernel(event_size=d, inv_mass_diag=imd), MHParameters(),
                 NeuTraKernel((d,), flow=product_flow_from_oracle(oflow)), NeuTraParameters(n_iterations=T))
    out = s.sample(z0, show_progress=False, normals=normals, uniforms=uniforms)
    lr = torch.stack(run.trace["log_ratio"])
    margin = (lr - torch.log(uniforms)).abs().min(dim=0).values
    clear = margin > 1e-3 * (1.0 + lr.abs().max(dim=0).values)
    assert clear.float().mean() > 0.9
    ref = run.samples
    close(out.samples[:, clear], ref[:, clear], atol=5e-5 * max(1.0, float(ref.abs().max())))
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~clear).sum()) * T
    assert out.statistics.n_target_calls == run.n_target_calls and out.statistics.n_target_gradient_calls == 0


# ------------------------------------------------------------------------------------------------------------
# NeuTra latent potential and its hand-written gradient against autograd through the oracle flow
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d,n_layers,ck,pot", [(6, 2, None, "fn"), (7, 3, dict(n_layers=3, n_hidden=6), "gm"),
                                               (100, 2, None, "g1"), (25, 3, None, "g0"), (9, 2, dict(n_layers=1), "g0"),
                                               (64, 2, dict(n_layers=2, n_hidden=32), "rb")])
def test_neutra_potential_and_gradient(d, n_layers, ck, pot):
    from gpu_util import product_flow_from_oracle, product_target
    from nfmc_b200 import _native as N
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=0.1, seed=d + 7)
    flow = product_flow_from_oracle(oflow)
    tgt_ref = make_potential_ref(pot, (d,))
    tgt = product_target(pot, d)
    torch.manual_seed(d)
    n = 29
    z = 0.5 * torch.randn(n, d)
    u_ref, g_ref = R.value_and_grad(R.neutra_potential(oflow, tgt_ref), z)
    dev = torch.device("cuda")
    zd = z.to(dev).contiguous()
    u = torch.empty(n, device=dev)
    gr = torch.empty(n, d, device=dev)
    pd, k1 = tgt.descriptor(dev)
    fd, k2 = flow.bijection.descriptor(dev)
    N.check(N.lib().nfmc_neutra_potential(C.byref(pd), C.byref(fd), N.ptr(zd), N.ptr(u), N.ptr(gr), n, N.stream_ptr(dev)))
    close(u, u_ref, atol=1e-5 * max(1.0, float(u_ref.abs().max())))
    close(gr, g_ref, atol=2e-5 * max(1.0, float(g_ref.abs().max())))


# ------------------------------------------------------------------------------------------------------------
# Philox path
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("d", [6, 25, 100, 1000])
def test_philox_fill_matches_numpy(d):
    from nfmc_b200 import _native as N
    dev = torch.device("cuda")
    n, steps, seed, chain0, step0 = 50, 3, 0x1234567890ABCDEF, 1000, 5
    for stream in (0, 1):
        nz = torch.empty(steps, n, d, device=dev)
        un = torch.empty(steps, n, device=dev)
        rng = N.rng_desc(seed, step0, None, None)
        N.check(N.lib().nfmc_rng_fill(C.byref(rng), stream, chain0, d, n, steps, N.ptr(nz), N.ptr(un), N.stream_ptr(dev)))
        for k in range(steps):
            z_ref, u_ref = step_noise(seed, stream, step0 + k, chain0, n, d)
            np.testing.assert_array_equal(un[k].cpu().numpy(), u_ref)              # bits -> uniform: exact
            err = np.abs(nz[k].cpu().numpy().astype(np.float64) - z_ref)
            assert err.max() < 2e-5, err.max()                                        # MUFU approximations


def test_philox_mode_equals_injected_mode():
    """The fused kernel drawing its own Philox numbers == the same kernel fed those numbers (d=100, K=6)."""
    from gpu_util import product_target, run_local_injected
    from nfmc_b200 import _native as N
    from nfmc_b200.records import LangevinKernel, LangevinParameters, MCMCOutput
    from nfmc_b200.samplers import MALA, DeviceSession
    dev = torch.device("cuda")
    d, n, K, seed = 100, 300, 6, 99
    torch.manual_seed(0)
    x0 = torch.randn(n, d)
    s = MALA((d,), product_target("g1", d), LangevinKernel(event_size=d, step_size=0.002), LangevinParameters())
    out = MCMCOutput((d,), store_samples=True)
    ses = DeviceSession(x0, (d,), None, seed=seed)
    buf = s.run_steps(ses, out, K, True)
    nz = torch.empty(K, n, d, device=dev)
    un = torch.empty(K, n, device=dev)
    rng = N.rng_desc(seed, 0, None, None)
    N.check(N.lib().nfmc_rng_fill(C.byref(rng), 0, 0, d, n, K, N.ptr(nz), N.ptr(un), N.stream_ptr(dev)))
    samples_inj, ses2, stats2 = run_local_injected(s, x0, nz, un)
    assert torch.equal(buf.cpu(), samples_inj)
    assert ses.read_back()[2][:2] == stats2[2][:2]


def test_hmc_philox_mode_equals_injected_mode():
    """The packed-fp32x2 HMC kernel drawing its own Philox numbers == the generic kernel fed those numbers."""
    from gpu_util import product_target, run_local_injected
    from nfmc_b200 import _native as N
    from nfmc_b200.records import HMCKernel, HMCParameters, MCMCOutput
    from nfmc_b200.samplers import HMC, DeviceSession
    dev = torch.device("cuda")
    from nfmc_b200.potentials import DiagonalGaussian
    # "dg": a diagonal Gaussian WITH means (the packed kernel's staged-table path keeps the subtraction), d = 99 and 1000:
    # odd split, invalid lanes in the last slot
    for pot, d, tau in (("g1", 100, 0.02), ("fn", 26, 0.05), ("g0", 1000, 0.05), ("dg", 99, 0.05), ("dg", 1000, 0.05), ("gm", 25, 0.1),
                        ("g1", 2, 0.02)):
        n, K, L, seed = 300, 4, 7, 99
        torch.manual_seed(0)
        x0 = 0.3 * torch.randn(n, d)
        target = DiagonalGaussian((d,), torch.linspace(0.5, 4.0, d), torch.linspace(-1.0, 1.0, d)) if pot == "dg" else product_target(pot, d)
        s = HMC((d,), target, HMCKernel(event_size=d, step_size=tau, n_leapfrog_steps=L), HMCParameters())
        out = MCMCOutput((d,), store_samples=True)
        ses = DeviceSession(x0, (d,), None, seed=seed)
        buf = s.run_steps(ses, out, K, True)
        nz = torch.empty(K, n, d, device=dev)
        un = torch.empty(K, n, device=dev)
        rng = N.rng_desc(seed, 0, None, None)
        N.check(N.lib().nfmc_rng_fill(C.byref(rng), 0, 0, d, n, K, N.ptr(nz), N.ptr(un), N.stream_ptr(dev)))
        samples_inj, ses2, stats2 = run_local_injected(s, x0, nz, un)
        assert torch.equal(buf.cpu(), samples_inj), pot
        assert ses.read_back()[2][:2] == stats2[2][:2]


# ------------------------------------------------------------------------------------------------------------
# larger batches: decisions agree except ties, moments agree, ragged tail tiles
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,pot,d,n,K", [("mala", "g1", 100, 1031, 5), ("mala", "gm", 25, 517, 8), ("hmc", "g1", 100, 773, 3),
                                            ("hmc", "fn", 26, 300, 3), ("mala", "g0", 1000, 67, 3), ("mala", "gm", 1023, 37, 3),
                                            ("hmc", "g1", 1024, 33, 2), ("mala", "fn", 3, 129, 4), ("hmc", "rb", 2, 65, 3)])
def test_local_kernels_against_oracle(kind, pot, d, n, K):
    from gpu_util import product_target, run_local_injected
    from nfmc_b200.records import LangevinKernel, LangevinParameters, HMCKernel, HMCParameters
    from nfmc_b200.samplers import MALA, HMC
    torch.manual_seed(n)
    x0 = 0.3 * torch.randn(n, d)
    normals = torch.randn(K, n, d)
    uniforms = torch.rand(K, n)
    ref_t = make_potential_ref(pot, (d,))
    imd = torch.ones(d)
    if kind == "mala":
        tau = {"g1": 0.004, "gm": 0.1, "g0": 0.02, "fn": 0.05}[pot]
        run = R.run_mala(x0, ref_t, tau, imd, K, R.TapeDraws(list(normals), list(uniforms)), trace=True)
        s = MALA((d,), product_target(pot, d), LangevinKernel(event_size=d, step_size=tau), LangevinParameters())
    else:
        tau, L = (0.03, 6) if pot == "g1" else (0.01, 5) if pot == "rb" else (0.05, 5)
        run = R.run_hmc(x0, ref_t, tau, imd, L, K, R.TapeDraws(list(normals), list(uniforms)), trace=True)
        s = HMC((d,), product_target(pot, d), HMCKernel(event_size=d, step_size=tau, n_leapfrog_steps=L), HMCParameters())
    samples, ses, (sx, sx2, cnt) = run_local_injected(s, x0, normals, uniforms)
    ref = run.samples
    # chains whose every decision had a clear margin must match to tolerance; ties may flip a decision
    lr = torch.stack(run.trace["log_ratio"])                           # [K, n]
    margin = (lr - torch.log(uniforms)).abs().min(dim=0).values
    clear = margin > 1e-3 * (1.0 + lr.abs().max(dim=0).values)
    assert clear.float().mean() > 0.97
    close(samples[:, clear], ref[:, clear], atol=2e-5 * max(1.0, float(ref.abs().max())))
    assert abs(cnt[0] - run.n_accepted) <= int((~clear).sum()) * K
    assert cnt[1] == run.n_attempted
    if bool(clear.all()):
        close(sx / (n * K), run.mean, atol=1e-5 * max(1.0, float(ref.abs().max())))


def test_jump_and_imh_against_oracle_large():
    from gpu_util import product_target, product_flow_from_oracle
    from nfmc_b200.records import IMHKernel, IMHParameters
    from nfmc_b200.samplers import FixedIMH
    d, n, T = 100, 1500, 4
    oflow = make_flow((d,), n_layers=2, perturb=0.05, seed=11)
    flow = product_flow_from_oracle(oflow)
    torch.manual_seed(5)
    x0 = torch.randn(n, d)
    z = torch.randn(T, n, d)
    u = torch.rand(T, n)
    run = R.run_fixed_imh(x0, make_potential_ref("g0", (d,)), oflow, T, R.TapeDraws(list(z), list(u)), trace=True)
    s = FixedIMH((d,), product_target("g0", d), IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    out = s.sample(x0, show_progress=False, z=z, uniforms=u)
    la = torch.stack(run.trace["log_alpha"])
    clear = ((la - torch.log(u)).abs().min(dim=0).values > 1e-3 * (1 + la.abs().max(dim=0).values))
    assert clear.float().mean() > 0.97
    close(out.samples[:, clear], run.samples[:, clear], atol=2e-5 * max(1.0, float(run.samples.abs().max())))
    assert abs(out.statistics.n_accepted_trajectories - run.n_accepted) <= int((~clear).sum()) * T


# ------------------------------------------------------------------------------------------------------------
# statistical correctness at scale (size-independent properties)
# ------------------------------------------------------------------------------------------------------------
def test_mala_moments_converge_on_gaussian():
    """MALA leaves N(0, 1/w) invariant: after burn-in the pooled moments match the target's."""
    import nfmc_b200
    from nfmc_b200.potentials import DiagonalGaussian
    torch.manual_seed(0)
    d, n = 100, 16384
    w = torch.linspace(0.5, 4.0, d)
    out = nfmc_b200.sample(DiagonalGaussian((d,), w), strategy="mala", n_chains=n, n_iterations=400, show_progress=False,
                           x0=torch.randn(n, d) / w.sqrt(), param_kwargs=dict(store_samples=False), kernel_kwargs=dict(step_size=0.05))
    assert out.samples is None
    var = out.variance
    assert float((var * w - 1).abs().max()) < 0.03
    assert float(out.mean.abs().max()) < 0.02
    assert 0.5 < out.statistics.acceptance_rate < 1.0


def test_jump_mala_full_size_properties():
    """BASELINE-size run (2^17 chains, d=100): counters follow the reference formulas, state finite, G-invariance of
    the Philox stream (a shard run with chain0 offset reproduces the corresponding rows)."""
    import nfmc_b200
    from nfmc_b200.potentials import StandardGaussian
    from nfmc_b200.flow import create_flow_object
    d, n, T, K = 100, 1 << 17, 2, 10
    torch.manual_seed(1)
    x0 = torch.randn(n, d)
    flow = create_flow_object("realnvp", (d,))
    with torch.no_grad():
        for p in flow.parameters():
            p.add_(0.05 * torch.randn_like(p))
    s = nfmc_b200.create_sampler(StandardGaussian((d,)), flow=flow, strategy="jump_mala",
                                 param_kwargs=dict(n_iterations=T, store_samples=False), inner_param_kwargs=dict(n_iterations=K))
    s.seed = 1234
    out = s.sample(x0, show_progress=False)
    st = out.statistics
    assert st.n_attempted_trajectories == n * T * K and st.n_attempted_jumps == n * T
    assert st.n_target_calls == 2 * n * K * T + 2 * n * T and st.n_target_gradient_calls == 2 * n * K * T
    assert st.expectations.n_seen == n * T * (K + 1)
    last = out.running_samples.last_sample
    assert last.shape == (n, d) and bool(torch.isfinite(last).all())
    # shard [4096, 8192) run on its own with chain0 = 4096
    s2 = nfmc_b200.create_sampler(StandardGaussian((d,)), flow=flow, strategy="jump_mala",
                                  param_kwargs=dict(n_iterations=T, store_samples=False), inner_param_kwargs=dict(n_iterations=K))
    s2.seed, s2.chain0 = 1234, 4096
    out2 = s2.sample(x0[4096:8192], show_progress=False)
    assert torch.equal(out2.running_samples.last_sample, last[4096:8192])
