"""Round-2 additions to the golden fixtures (same rules as make_golden.py: the UNMODIFIED reference is run, every random
draw is recorded).  Run in the build container only:

    PYTHONPATH=/root/repo:/root/repo/oracle/shim:/root/reference python tests/golden/make_golden_r02.py

New cases:
* warm-up / tuning trajectories (mcmc/base.py:39-54,142-161, tuning.py:15-41): MALA and HMC with ``params.tuning``;
  ``step_traj[i]`` / ``imd_traj[i]`` are the kernel's step size and inverse-mass diagonal after iteration i
* ULA and UHMC (langevin.py:131-134, hmc.py:129-132)
* AdaptiveIMH with ``flow.fit`` stubbed (imh.py:102-181; the refit is torchflows' optimiser, absent here)
* small-n cases at the BASELINE config shapes: C2 jump_hmc / G1 / d = 100 / L = 20, C4 imh / RB / d = 100,
  C5 jump_mala / GM / d = 1000
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import Tape, pack, flow_arrays, make_flow, make_potential_ref          # noqa: E402  (sets sys.path)
from nfmc.algorithms.sampling.mcmc.hmc import HMC, UHMC, HMCKernel, HMCParameters      # noqa: E402
from nfmc.algorithms.sampling.mcmc.langevin import MALA, ULA, LangevinKernel, LangevinParameters  # noqa: E402
from nfmc.algorithms.sampling.nfmc.imh import AdaptiveIMH, FixedIMH, IMHKernel, IMHParameters     # noqa: E402
from nfmc.algorithms.sampling.nfmc.jump import JumpMALA, JumpHMC, JumpNFMCParameters   # noqa: E402
from nfmc.algorithms.sampling.base import NFMCKernel                                   # noqa: E402


def record_tuning(sampler):
    """Wrap ``update_kernel`` so that the kernel state after every warm-up iteration is kept."""
    steps, imds = [], []
    orig = sampler.update_kernel

    def wrapped(data):
        orig(data)
        steps.append(float(sampler.kernel.step_size))
        imds.append(sampler.kernel.inv_mass_diag.detach().clone().numpy())

    sampler.update_kernel = wrapped
    return steps, imds


def main():
    torch.set_num_threads(1)
    cases = {}

    # ---- MALA warm-up: dual averaging of the step size + inverse-mass EMA -------------------------------------------
    torch.manual_seed(31)
    d, n, K = 8, 16, 10
    target = make_potential_ref("g1", (d,))
    x0 = 0.5 * torch.randn(n, d)
    s = MALA((d,), target, LangevinKernel(event_size=d, step_size=0.05), LangevinParameters(n_iterations=K))
    s.params.tuning_mode()
    steps, imds = record_tuning(s)
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["mala_tune_g1"] = pack(out, t, x0, dict(pot="g1", step=0.05, imd=np.ones(d, np.float32), K=K,
                                                   step_traj=np.array(steps), imd_traj=np.stack(imds)))

    # ---- HMC warm-up -------------------------------------------------------------------------------------------------
    torch.manual_seed(32)
    d, n, K, L = 6, 12, 8, 4
    target = make_potential_ref("fn", (d,))
    x0 = 0.3 * torch.randn(n, d)
    s = HMC((d,), target, HMCKernel(event_size=d, step_size=0.02, n_leapfrog_steps=L), HMCParameters(n_iterations=K))
    s.params.tuning_mode()
    steps, imds = record_tuning(s)
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["hmc_tune_fn"] = pack(out, t, x0, dict(pot="fn", step=0.02, imd=np.ones(d, np.float32), K=K, L=L,
                                                  step_traj=np.array(steps), imd_traj=np.stack(imds)))

    # ---- ULA / UHMC ---------------------------------------------------------------------------------------------------
    torch.manual_seed(33)
    d, n, K = 7, 6, 5
    target = make_potential_ref("g0", (d,))
    imd = 0.5 + torch.rand(d)
    x0 = torch.randn(n, d)
    s = ULA((d,), target, LangevinKernel(event_size=d, inv_mass_diag=imd.clone(), step_size=0.1), LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["ula_g0"] = pack(out, t, x0, dict(pot="g0", step=0.1, imd=imd.numpy(), K=K))

    torch.manual_seed(34)
    d, n, K, L = 7, 5, 4, 3
    target = make_potential_ref("gm", (d,))
    x0 = torch.randn(n, d)
    s = UHMC((d,), target, HMCKernel(event_size=d, step_size=0.05, n_leapfrog_steps=L), HMCParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["uhmc_gm"] = pack(out, t, x0, dict(pot="gm", step=0.05, imd=np.ones(d, np.float32), K=K, L=L))

    # ---- adaptive IMH, refit stubbed -----------------------------------------------------------------------------------
    torch.manual_seed(35)
    d, n, T = 6, 10, 5
    target = make_potential_ref("rb", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.1, seed=110)
    flow.fit = lambda *a, **k: None
    x0 = torch.randn(n, d)
    s = AdaptiveIMH((d,), target, IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["adaptive_imh_rb"] = pack(out, t, x0, dict(pot="rb", T=T, **flow_arrays(flow, 2, 2, 4)))

    # ---- C2 shape: jump_hmc, ill-conditioned Gaussian, d = 100, L = 20 ------------------------------------------------
    torch.manual_seed(36)
    d, n, T, K, L = 100, 8, 1, 2, 20
    target = make_potential_ref("g1", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.05, seed=111)
    x0 = 0.1 * torch.randn(n, d)
    s = JumpHMC((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                inner_kernel=HMCKernel(event_size=d, step_size=0.01, n_leapfrog_steps=L), inner_params=HMCParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_hmc_g1_d100"] = pack(out, t, x0, dict(pot="g1", step=0.01, imd=np.ones(d, np.float32), K=K, T=T, L=L,
                                                       **flow_arrays(flow, 2, 2, 5)))

    # ---- C4 shape: fixed IMH, Rosenbrock, d = 100 ------------------------------------------------------------------------
    torch.manual_seed(37)
    d, n, T = 100, 24, 3
    target = make_potential_ref("rb", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.05, seed=112)
    x0 = torch.randn(n, d)
    s = FixedIMH((d,), target, IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["imh_rb_d100"] = pack(out, t, x0, dict(pot="rb", T=T, **flow_arrays(flow, 2, 2, 5)))

    # ---- C5 shape: jump_mala, Gaussian mixture, d = 1000 (default conditioner H = 8) ----------------------------------------
    torch.manual_seed(38)
    d, n, T, K = 1000, 6, 1, 2
    target = make_potential_ref("gm", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.02, seed=113)
    x0 = torch.randn(n, d)
    s = JumpMALA((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                 inner_kernel=LangevinKernel(event_size=d), inner_params=LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_mala_gm_d1000"] = pack(out, t, x0, dict(pot="gm", step=d ** (-1 / 3), imd=np.ones(d, np.float32), K=K, T=T,
                                                         **flow_arrays(flow, 2, 2, 8)))

    for name, arrays in cases.items():
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"wrote {path}  ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main()
