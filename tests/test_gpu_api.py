"""B200: API conformance -- the reference's own smoke tests (/root/reference/test/*.py) restated against nfmc_b200 for the
strategies on the accelerated path.  Same assertions: return type, sample shapes, finiteness, store_samples semantics,
flow-string parsing, moment shapes, N-D events, warm-up hand-off."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import nfmc_b200
from nfmc_b200 import sample, create_sampler
from nfmc_b200.potentials import StandardGaussian, DiagonalGaussian
from nfmc_b200.records import MCMCOutput


def _g(shape):
    return StandardGaussian(shape)


# /root/reference/test/test_samplers.py:20-40 (test_mcmc), restricted to the local kernels on the hot path
@pytest.mark.parametrize("name", ["HMC", "UHMC", "MALA", "ULA", "MH", "RandomWalk"])
def test_mcmc(name):
    from nfmc_b200 import samplers
    torch.manual_seed(0)
    s = getattr(samplers, name)(event_shape=(5,), target=_g((5,)))
    s.params.n_iterations = 3
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


# test_samplers.py:124-145 (test_jump_nfmc)
@pytest.mark.parametrize("name", ["JumpMALA", "JumpHMC", "JumpUHMC", "JumpULA", "JumpMH"])
def test_jump_nfmc(name):
    from nfmc_b200 import samplers
    torch.manual_seed(0)
    s = getattr(samplers, name)(event_shape=(5,), target=_g((5,)))
    s.params.n_iterations = 3
    s.inner_sampler.params.n_iterations = 4
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3 * (4 + 1), 4, 5) and bool(torch.isfinite(out.samples).all())


# test_samplers.py:148-172 (test_other_nfmc)
@pytest.mark.parametrize("name", ["NeuTraHMC", "FixedIMH", "AdaptiveIMH"])
def test_other_nfmc(name):
    from nfmc_b200 import samplers
    torch.manual_seed(0)
    s = getattr(samplers, name)(event_shape=(5,), target=_g((5,)))
    s.params.n_iterations = 3
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


# test_samplers.py:175-201 (test_sample_wrapper_no_jump)
@pytest.mark.parametrize("strategy", ["hmc", "uhmc", "ula", "mala", "mh", "imh", "neutra_hmc"])
def test_sample_wrapper_no_jump(strategy):
    torch.manual_seed(0)
    out = sample(_g((5,)), event_shape=(5,), strategy=strategy, n_chains=4, n_iterations=3, device=torch.device("cuda"),
                 show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


# test_samplers.py:227-248 (test_sample_wrapper_jump)
@pytest.mark.parametrize("strategy", ["jump_mala", "jump_ula", "jump_hmc", "jump_uhmc", "jump_mh"])
def test_sample_wrapper_jump(strategy):
    torch.manual_seed(0)
    out = sample(_g((5,)), event_shape=(5,), strategy=strategy, n_chains=4, n_iterations=3,
                 inner_param_kwargs={"n_iterations": 7}, device=torch.device("cuda"), show_progress=False)
    assert out.samples.shape == (3 * 8, 4, 5) and bool(torch.isfinite(out.samples).all())


def test_jump_hmc_default_inner_iterations():
    """create_sampler forces 5 inner HMC steps for jump_hmc (/root/reference/nfmc/sample.py:161-162)."""
    out = sample(_g((5,)), strategy="jump_hmc", n_chains=4, n_iterations=2, show_progress=False)
    assert out.samples.shape == (2 * 6, 4, 5)


# test_flow_kwargs.py
def test_flow_kwargs():
    torch.manual_seed(0)
    basic = sample(event_shape=(100,), target=_g((100,)), flow="realnvp", strategy="imh", n_iterations=3, n_warmup_iterations=3,
                   show_progress=False)
    adv = sample(event_shape=(100,), target=_g((100,)), flow='realnvp%{"n_layers": 10}', strategy="imh", n_iterations=3,
                 show_progress=False)
    adv2 = sample(event_shape=(100,), target=_g((100,)), strategy="imh", n_iterations=3, show_progress=False,
                  flow='realnvp%{"n_layers": 10, "conditioner_kwargs": {"n_layers": 5, "n_hidden": 100}}', n_chains=16)
    nb = len(basic.kernel.flow.bijection.layers)
    assert len(adv.kernel.flow.bijection.layers) > nb and len(adv2.kernel.flow.bijection.layers) > nb
    assert bool(torch.isfinite(adv2.samples).all())


# test_no_sample_storing.py:33-49 (test_sampling)
@pytest.mark.parametrize("strategy", ["hmc", "uhmc", "ula", "mala", "imh", "fixed_imh", "jump_mala", "jump_ula", "jump_hmc",
                                      "jump_uhmc", "neutra_hmc"])
def test_no_sample_storing(strategy):
    torch.manual_seed(0)
    s = create_sampler(target=_g((10,)), event_shape=(10,), strategy=strategy, param_kwargs={"store_samples": False})
    out = s.sample(torch.randn(20, 10), time_limit_seconds=1.0, show_progress=False)
    assert out.samples is None
    assert out.running_samples.last_sample is not None and out.running_samples.last_sample.shape == (20, 10)


def test_adaptive_imh_ignores_param_kwargs():
    """Quirk Q2 (/root/reference/nfmc/sample.py:129): adaptive_imh always runs 100 iterations and stores samples."""
    s = create_sampler(target=_g((10,)), event_shape=(10,), strategy="adaptive_imh", param_kwargs={"store_samples": False, "n_iterations": 3})
    out = s.sample(torch.randn(8, 10), show_progress=False)
    assert out.samples.shape == (100, 8, 10)
    assert out.statistics.n_target_gradient_calls == 2 * 8 * 100 and out.statistics.n_target_calls == 0   # quirk Q3


# test_moment_estimation.py
@pytest.mark.parametrize("strategy", ["hmc", "mala", "imh", "adaptive_imh", "jump_mala", "jump_hmc", "neutra_hmc"])
def test_moments(strategy):
    torch.manual_seed(0)
    out = sample(target=_g((10,)), event_shape=(10,), strategy=strategy, n_iterations=3, n_warmup_iterations=3, show_progress=False)
    for t in (out.mean, out.second_moment, out.variance, out.statistics.running_first_moment, out.statistics.running_second_moment):
        assert t.shape == (10,) and bool(t.isfinite().all())


def test_moments_diag_gaussian_100d():
    torch.manual_seed(0)
    target = DiagonalGaussian((100,), 1.0 / torch.linspace(1.0, 10.0, 100) ** 2)
    for cls in ("HMC", "NeuTraHMC", "JumpHMC", "AdaptiveIMH"):
        from nfmc_b200 import samplers
        s = getattr(samplers, cls)(target.event_shape, target)
        s.params.n_iterations = 3
        if cls == "JumpHMC":
            s.inner_sampler.params.n_iterations = 3
        out = s.sample(torch.randn(100, 100), show_progress=False)
        assert out.statistics.running_first_moment.shape == (100,) and bool(out.statistics.running_second_moment.isfinite().all())


# test_custom_shapes.py (N-D events)
@pytest.mark.parametrize("strategy", ["imh", "jump_hmc", "neutra_hmc", "hmc", "jump_mala", "mala"])
def test_image_shaped_events(strategy):
    torch.manual_seed(0)
    out = sample(_g((8, 8)), event_shape=(8, 8), strategy=strategy, n_iterations=2, n_warmup_iterations=2, n_chains=3, warmup=False,
                 inner_param_kwargs=dict(n_iterations=2), show_progress=False)
    rows = {"jump_hmc": 6, "jump_mala": 6}.get(strategy, 2)
    assert out.samples.shape == (rows, 3, 8, 8) and out.mean.shape == (8, 8)
    assert out.running_samples.last_sample.shape == (3, 8, 8)


# test_warmup.py (local samplers + jump): step size and mass adapt, sampling starts from the warm-up state
@pytest.mark.parametrize("strategy", ["mala", "hmc", "jump_mala", "jump_hmc"])
def test_warmup_then_sample(strategy):
    torch.manual_seed(0)
    s = create_sampler(target=DiagonalGaussian((10,), torch.linspace(0.5, 5.0, 10)), event_shape=(10,), strategy=strategy,
                       param_kwargs={"n_iterations": 5, "n_warmup_iterations": 30}, inner_param_kwargs={"n_iterations": 4, "n_warmup_iterations": 30})
    inner = getattr(s, "inner_sampler", s)
    step0 = inner.kernel.step_size
    w = s.warmup(torch.randn(64, 10), show_progress=False)
    assert isinstance(w, MCMCOutput) and w.running_samples.last_sample.shape == (64, 10)
    assert inner.kernel.step_size != step0 and not inner.kernel.has_unit_mass()
    assert not inner.params.tuning
    out = sample(DiagonalGaussian((10,), torch.linspace(0.5, 5.0, 10)), strategy=strategy, n_chains=32, n_iterations=3,
                 n_warmup_iterations=10, warmup=True, show_progress=False, inner_param_kwargs={"n_iterations": 4, "n_warmup_iterations": 10})
    assert bool(torch.isfinite(out.samples).all())


def test_time_limit_stops_early():
    s = create_sampler(target=_g((100,)), event_shape=(100,), strategy="jump_mala", param_kwargs={"n_iterations": 100000, "store_samples": False})
    out = s.sample(torch.randn(4096, 100), time_limit_seconds=0.2, show_progress=False)
    assert 0 < out.statistics.n_attempted_jumps < 100000 * 4096
    assert out.statistics.elapsed_time_seconds >= 0.2


def test_unsupported_strategy_and_target_raise():
    with pytest.raises(NotImplementedError):
        sample(_g((5,)), strategy="nuts", n_chains=4, n_iterations=2)
    assert sorted(nfmc_b200.get_supported_samplers()) == sorted([
        "hmc", "uhmc", "ula", "mala", "mh", "ess", "imh", "fixed_imh", "adaptive_imh", "jump_mala", "jump_ula", "jump_hmc",
        "jump_uhmc", "jump_ess", "jump_mh", "neutra_mh", "neutra_hmc", "tess", "dlmc"])        # reference: util.py:421-444
    with pytest.raises(NotImplementedError):       # transport ESS evaluates the likelihood inside the flow sweep: analytic only
        sample(_g((6,)), event_shape=(6,), strategy="tess", n_chains=4, n_iterations=2, show_progress=False,
               negative_log_likelihood=lambda x: (x ** 2).sum(-1))
    with pytest.raises(TypeError):
        sample(3.0, event_shape=(5,), strategy="mala", n_chains=4, n_iterations=2)
    with pytest.raises(ValueError):
        nfmc_b200.samplers.JumpMALA((5,), _g((5,)), inner_params=nfmc_b200.records.LangevinParameters(store_samples=False)).sample(torch.randn(3, 5))


def test_thinning_and_max_samples():
    from nfmc_b200.samplers import MALA
    s = MALA((6,), _g((6,)))
    s.params.n_iterations = 10
    out = s.sample(torch.randn(5, 6), show_progress=False)
    full = out.samples
    torch.manual_seed(1)
    # same chain, same seed, with thinning 3: rows 0, 3, 6, 9 of the un-thinned run
    s2 = MALA((6,), _g((6,)))
    s2.params.n_iterations = 10
    s.seed = s2.seed = 77
    full = s.sample(torch.zeros(5, 6), show_progress=False).samples
    from nfmc_b200.records import MCMCOutput as MO
    import nfmc_b200.samplers as S
    out2 = MO((6,), store_samples=True)
    out2.running_samples.thinning = 3
    ses = S.DeviceSession(torch.zeros(5, 6), (6,), None, seed=77)
    buf = s2.run_steps(ses, out2, 10, True)
    assert buf.shape[0] == 4 and torch.equal(buf.cpu(), full[0::3])


# ---- flow training hooks (nfmc_b200/flow_train.py; the default conditioners train on the native kernels) ----------------
def test_flow_fit_improves_likelihood_and_kernels_see_new_weights():
    from nfmc_b200.flow import create_flow_object
    torch.manual_seed(0)
    flow = create_flow_object("realnvp", (6,)).to("cuda")
    x = 0.3 * torch.randn(2048, 6, device="cuda") + 1.5
    before = float(flow.log_prob(x).mean())
    flow.fit(x, n_epochs=60, lr=0.05, batch_size=512)
    after = float(flow.log_prob(x).mean())                 # log_prob runs the CUDA kernel on the re-packed blob
    assert after > before + 1.0
    from gpu_util import oracle_flow_from_product
    with torch.no_grad():
        ref = float(oracle_flow_from_product(flow).log_prob(x).mean())      # the oracle flow with the fitted weights
    assert abs(after - ref) < 1e-3 * (1 + abs(ref))


def test_jump_mala_with_in_loop_flow_fitting():
    """fit_nf=True (config C5's shape of loop): the flow is refitted inside the sampling loop and jump acceptance rises."""
    torch.manual_seed(0)
    target = DiagonalGaussian((8,), torch.full((8,), 4.0), mean=torch.full((8,), 2.0))
    common = dict(strategy="jump_mala", n_chains=1024, show_progress=False, inner_param_kwargs={"n_iterations": 20},
                  x0=2.0 + 0.5 * torch.randn(1024, 8))
    frozen = sample(target, n_iterations=6, param_kwargs={"fit_nf": False}, **common)
    fitted = sample(target, n_iterations=6, param_kwargs={"fit_nf": True, "n_jumps_before_training": 1,
                                                           "flow_fit_kwargs": {"n_epochs": 30, "lr": 0.05, "batch_size": 1024}}, **common)
    assert fitted.statistics.jump_acceptance_rate > frozen.statistics.jump_acceptance_rate + 0.05
    assert bool(torch.isfinite(fitted.samples).all())


def test_imh_and_neutra_warmup_fit_the_flow():
    torch.manual_seed(0)
    target = DiagonalGaussian((6,), torch.full((6,), 1.0), mean=torch.full((6,), 1.0))
    s = create_sampler(target, strategy="imh", param_kwargs={"n_iterations": 5, "warmup_fit_kwargs": {"n_epochs": 150, "lr": 0.05, "n_samples": 256}})
    w = s.warmup(torch.randn(256, 6), show_progress=False)
    out = s.sample(w.running_samples.last_sample, show_progress=False)
    assert out.statistics.acceptance_rate > 0.3            # an untrained (identity) flow gets far less on a shifted target
    s2 = create_sampler(target, strategy="neutra_hmc", param_kwargs={"n_iterations": 3, "n_warmup_iterations": 10,
                                                                      "warmup_fit_kwargs": {"n_epochs": 50, "lr": 0.05, "n_samples": 128}})
    step0 = s2.inner_kernel.step_size
    w2 = s2.warmup(torch.randn(64, 6), show_progress=False)
    assert s2.inner_kernel.step_size != step0 and w2.running_samples.last_sample.shape == (64, 6)
    assert bool(torch.isfinite(s2.sample(w2.running_samples.last_sample, show_progress=False).samples).all())


def test_adaptive_imh_refits():
    torch.manual_seed(0)
    from nfmc_b200.samplers import AdaptiveIMH
    target = DiagonalGaussian((6,), torch.full((6,), 1.0), mean=torch.full((6,), 1.0))
    s = AdaptiveIMH((6,), target)
    s.params.n_iterations = 12
    before = {k: v.clone() for k, v in s.kernel.flow.state_dict().items()}
    out = s.sample(torch.randn(512, 6), show_progress=False)
    after = s.kernel.flow.state_dict()
    assert any(not torch.equal(before[k].cpu(), after[k].cpu()) for k in before)
    assert out.samples.shape == (12, 512, 6)


@pytest.mark.parametrize("strategy,inner", [("jump_mala", dict(n_iterations=6)), ("jump_hmc", dict(n_iterations=2)),
                                            ("jump_mh", dict(n_iterations=5)), ("jump_ula", dict(n_iterations=4))])
def test_fused_whole_run_equals_per_iteration_loop(strategy, inner):
    """store_samples=False runs go through nfmc_jump_sample_device (slabs over several streams); a time limit forces the
    per-iteration loop.  Same Philox steps, so states, moments and every counter agree exactly."""
    import nfmc_b200
    from nfmc_b200.flow import create_flow_object
    from nfmc_b200.potentials import make_potential
    d, n, T = 26, 70001, 3
    outs = []
    for limit in (None, 1e9):
        torch.manual_seed(3)
        flow = create_flow_object("realnvp", (d,))
        with torch.no_grad():
            for p in flow.parameters():
                p.add_(0.05 * torch.randn_like(p))
        s = nfmc_b200.create_sampler(make_potential("rb", (d,)), flow=flow, strategy=strategy,
                                     param_kwargs={"n_iterations": T, "store_samples": False}, inner_param_kwargs=dict(inner),
                                     inner_kernel_kwargs={} if strategy == "jump_mh" else {"step_size": 0.02})
        s.seed = 77
        x0 = 0.5 * torch.randn(n, d)
        outs.append(s.sample(x0, show_progress=False, time_limit_seconds=limit))
    a, b = outs
    la, lb = a.running_samples.last_sample, b.running_samples.last_sample
    assert torch.equal(torch.isnan(la), torch.isnan(lb))          # unadjusted chains may blow up -- identically
    assert torch.equal(torch.nan_to_num(la), torch.nan_to_num(lb))
    sa, sb = a.statistics, b.statistics
    for f in ("n_accepted_trajectories", "n_attempted_trajectories", "n_accepted_jumps", "n_attempted_jumps", "n_target_calls",
              "n_target_gradient_calls", "n_divergences"):
        assert getattr(sa, f) == getattr(sb, f), f
    if not bool(torch.isnan(la).any()):
        np.testing.assert_allclose(np.asarray(a.mean), np.asarray(b.mean), rtol=0, atol=1e-6)
        np.testing.assert_allclose(np.asarray(a.second_moment), np.asarray(b.second_moment), rtol=1e-6, atol=1e-6)


def test_full_pipeline_recovers_target_moments():
    """warm-up (tuning + flow fit) then jump_mala on an anisotropic Gaussian: the pooled moments land on the target's, and
    the fitted flow makes the NF jumps useful (a sizeable share is accepted)."""
    d, n = 10, 2048
    torch.manual_seed(12)
    sig = torch.logspace(-0.5, 0.5, d)                           # standard deviations 0.32 .. 3.2
    target = DiagonalGaussian((d,), precision=1.0 / sig ** 2, mean=torch.linspace(-1.0, 1.0, d))
    out = sample(target, strategy="jump_mala", flow="realnvp", n_chains=n, n_iterations=30, n_warmup_iterations=40,
                 warmup=True, show_progress=False, inner_param_kwargs=dict(n_iterations=20),
                 param_kwargs=dict(store_samples=False))
    mean, var = np.asarray(out.mean), np.asarray(out.variance)
    assert np.abs((mean - np.linspace(-1.0, 1.0, d)) / sig.numpy()).max() < 0.1, mean
    np.testing.assert_allclose(var, sig.numpy() ** 2, rtol=0.15)
    assert out.statistics.jump_acceptance_rate > 0.2, out.statistics.jump_acceptance_rate


# ---- the rest of the reference's sampler suite (test_samplers.py:57-140,204-262), restated ---------------------------------
def test_ess_class():                                                        # test_samplers.py:57-74
    from nfmc_b200.samplers import ESS
    torch.manual_seed(0)
    s = ESS(event_shape=(5,), target=_g((5,)), negative_log_likelihood=_g((5,)))
    s.params.n_iterations = 3
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


def test_jump_ess_class():                                                   # test_samplers.py:77-96
    from nfmc_b200.samplers import JumpESS
    torch.manual_seed(0)
    s = JumpESS(event_shape=(5,), target=_g((5,)), negative_log_likelihood=_g((5,)))
    s.params.n_iterations = 3
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert out.samples.shape == (3 * (s.inner_sampler.params.n_iterations + 1), 4, 5)
    assert bool(torch.isfinite(out.samples).all())


@pytest.mark.parametrize("name", ["TESS", "DLMC"])
def test_nfmc_with_nll(name):                                                # test_samplers.py:99-120
    from nfmc_b200 import samplers
    torch.manual_seed(0)
    s = getattr(samplers, name)(event_shape=(5,), target=_g((5,)), negative_log_likelihood=_g((5,)))
    s.params.n_iterations = 3
    out = s.sample(x0=torch.randn(4, 5), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


@pytest.mark.parametrize("strategy", ["dlmc", "tess", "ess"])
def test_sample_wrapper_nll(strategy):                                       # test_samplers.py:204-224
    torch.manual_seed(0)
    out = sample(_g((5,)), event_shape=(5,), strategy=strategy, negative_log_likelihood=_g((5,)), n_chains=4, n_iterations=3,
                 device=torch.device("cuda"), show_progress=False)
    assert isinstance(out, MCMCOutput)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())


def test_sample_wrapper_jump_ess():                                          # test_samplers.py:250-272
    torch.manual_seed(0)
    out = sample(_g((5,)), event_shape=(5,), strategy="jump_ess", negative_log_likelihood=_g((5,)), n_chains=4, n_iterations=3,
                 inner_param_kwargs={"n_iterations": 7}, device=torch.device("cuda"), show_progress=False)
    assert out.samples.shape == (3 * 8, 4, 5) and bool(torch.isfinite(out.samples).all())


def test_neutra_mh_wrapper():
    torch.manual_seed(0)
    out = sample(_g((5,)), event_shape=(5,), strategy="neutra_mh", n_chains=4, n_iterations=3, device=torch.device("cuda"),
                 show_progress=False)
    assert out.samples.shape == (3, 4, 5) and bool(torch.isfinite(out.samples).all())
