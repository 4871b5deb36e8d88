"""CPU: pin oracle/samplers_ref.py to outputs of the unmodified reference (tests/golden/, made by make_golden.py)."""
import numpy as np
import pytest
import torch

from golden_util import load_case, tape, oracle_flow, oracle_target
from oracle import samplers_ref as R


def _check(g, run, jump=False):
    # the restatement follows the reference op for op, so agreement is to the last bit on the same CPU
    np.testing.assert_allclose(run.samples.numpy(), g["samples"], rtol=0, atol=0)
    np.testing.assert_allclose(run.x.numpy(), g["last"], rtol=0, atol=0)
    np.testing.assert_allclose(np.asarray(run.mean), g["mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(np.asarray(run.second_moment), g["second_moment"], rtol=1e-6, atol=1e-7)
    acc, att, div, grads, calls, jacc, jatt = (int(v) for v in g["counters"])
    assert (run.n_accepted, run.n_attempted, run.n_divergences) == (acc, att, div)
    assert (run.n_grad_calls, run.n_target_calls) == (grads, calls)
    if jump:
        assert (run.n_accepted_jumps, run.n_attempted_jumps) == (jacc, jatt)


@pytest.mark.parametrize("name", ["mala_g0", "mala_fn"])
def test_mala(name):
    g = load_case(name)
    run = R.run_mala(torch.from_numpy(g["x0"]), oracle_target(g), float(g["step"]), torch.from_numpy(g["imd"]),
                     int(g["K"]), tape(g))
    _check(g, run)


@pytest.mark.parametrize("name", ["hmc_g1", "hmc_rb"])
def test_hmc(name):
    g = load_case(name)
    run = R.run_hmc(torch.from_numpy(g["x0"]), oracle_target(g), float(g["step"]), torch.from_numpy(g["imd"]),
                    int(g["L"]), int(g["K"]), tape(g))
    _check(g, run)


@pytest.mark.parametrize("name", ["jump_mala_g0", "jump_mala_g1_d100", "jump_mala_gm_d1000"])
def test_jump_mala(name):
    g = load_case(name)
    run = R.run_jump(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), "mala", int(g["T"]), int(g["K"]),
                     tape(g), float(g["step"]), torch.from_numpy(g["imd"]))
    _check(g, run, jump=True)


@pytest.mark.parametrize("name", ["jump_hmc_gm", "jump_hmc_g1_d100"])
def test_jump_hmc(name):
    g = load_case(name)
    run = R.run_jump(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), "hmc", int(g["T"]), int(g["K"]),
                     tape(g), float(g["step"]), torch.from_numpy(g["imd"]), n_leapfrog=int(g["L"]))
    _check(g, run, jump=True)


@pytest.mark.parametrize("name", ["imh_rb", "imh_rb_d100"])
def test_fixed_imh(name):
    g = load_case(name)
    run = R.run_fixed_imh(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), int(g["T"]), tape(g))
    _check(g, run)


@pytest.mark.parametrize("name", ["neutra_hmc_fn", "neutra_hmc_fn_d100"])
def test_neutra_hmc(name):
    g = load_case(name)
    run = R.run_neutra_hmc(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), int(g["T"]), tape(g),
                           float(g["step"]), torch.from_numpy(g["imd"]), n_leapfrog=int(g["L"]))
    _check(g, run)


def test_mh():
    g = load_case("mh_gm")
    run = R.run_mh(torch.from_numpy(g["x0"]), oracle_target(g), torch.from_numpy(g["imd"]), int(g["K"]), tape(g))
    _check(g, run)


def test_ess():
    g = load_case("ess_fn")
    run = R.run_ess(torch.from_numpy(g["x0"]), oracle_target(g), int(g["K"]), tape(g), max_iterations=int(g["M"]))
    _check(g, run)


def test_jump_ess():
    from oracle.potentials_ref import make_potential_ref
    g = load_case("jump_ess_gm")
    d = g["x0"].shape[1]
    run = R.run_jump(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), "ess", int(g["T"]), int(g["K"]),
                     tape(g), 0.0, torch.ones(d), nll=make_potential_ref(str(g["nll"]), (d,)),
                     max_ess_iterations=int(g["M"]))
    _check(g, run, jump=True)


def test_neutra_mh():
    g = load_case("neutra_mh_gm")
    run = R.run_neutra_mh(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), int(g["T"]), tape(g),
                          torch.from_numpy(g["imd"]))
    _check(g, run)


def test_tess():
    g = load_case("tess_fn")
    # draw order per iteration: normal(n,d) = v, uniform(n) = w, normal(n) = theta, then M x uniform(n,1)
    run = R.run_tess(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), int(g["T"]), tape(g), max_iterations=int(g["M"]))
    _check(g, run)


@pytest.mark.parametrize("name", ["dlmc_gm", "dlmc_latent_gm"])
def test_dlmc(name):
    from oracle.potentials_ref import make_potential_ref
    g = load_case(name)
    d = g["x0"].shape[1]
    run = R.run_dlmc(torch.from_numpy(g["x0"]), oracle_target(g), make_potential_ref(str(g["nll"]), (d,)), oracle_flow(g),
                     int(g["T"]), tape(g), step_size=float(g["step"]), latent_updates=bool(int(g["latent"])))
    _check(g, run)


# ---- round 2: warm-up trajectories, unadjusted kernels, adaptive IMH ----------------------------------------------------
@pytest.mark.parametrize("name,kind", [("mala_tune_g1", "mala"), ("hmc_tune_fn", "hmc")])
def test_warmup_trajectory(name, kind):
    """mcmc/base.py:142-161 + tuning.py:15-41: the step size and inverse-mass diagonal after EVERY warm-up iteration."""
    g = load_case(name)
    run = R.run_tuned(torch.from_numpy(g["x0"]), oracle_target(g), kind, float(g["step"]), torch.from_numpy(g["imd"]),
                      int(g["K"]), tape(g), n_leapfrog=int(g["L"]) if "L" in g else 20)
    _check(g, run)
    np.testing.assert_allclose(np.array(run.step_traj), g["step_traj"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(torch.stack(run.imd_traj).numpy(), g["imd_traj"], rtol=0, atol=0)


def test_ula():
    g = load_case("ula_g0")
    run = R.run_mala(torch.from_numpy(g["x0"]), oracle_target(g), float(g["step"]), torch.from_numpy(g["imd"]),
                     int(g["K"]), tape(g), adjusted=False)
    _check(g, run)


def test_uhmc():
    g = load_case("uhmc_gm")
    run = R.run_hmc(torch.from_numpy(g["x0"]), oracle_target(g), float(g["step"]), torch.from_numpy(g["imd"]),
                    int(g["L"]), int(g["K"]), tape(g), adjusted=False)
    _check(g, run)


def test_adaptive_imh():
    g = load_case("adaptive_imh_rb")
    run = R.run_adaptive_imh(torch.from_numpy(g["x0"]), oracle_target(g), oracle_flow(g), int(g["T"]), tape(g))
    _check(g, run)
