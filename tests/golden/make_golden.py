"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONPATH=/root/repo:/root/repo/oracle/shim:/root/reference python tests/golden/make_golden.py

The reference's sampler classes are imported from ``/root/reference``; the absent third-party packages
``torchflows`` / ``potentials`` are satisfied by ``oracle/shim`` (the oracle's RealNVP restatement -- the
flow arithmetic therefore stays "parity unpinned", see oracle/__init__.py).  Every random draw the reference
makes is recorded (the reference has no injection hook) and stored in the fixture together
with the inputs and the reference's outputs, so that

* ``tests/test_oracle_golden.py`` (CPU) replays the tape through ``oracle/samplers_ref.py`` and must
  reproduce the reference's samples / counters / moments, and
* ``tests/test_gpu_parity.py`` (B200) injects the same tape into the CUDA kernels.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle", "shim"), "/root/reference"):
    if p not in sys.path:
        sys.path.insert(0, p)

from nfmc.algorithms.sampling.mcmc.hmc import HMC, HMCKernel, HMCParameters          # noqa: E402
from nfmc.algorithms.sampling.mcmc.langevin import MALA, LangevinKernel, LangevinParameters  # noqa: E402
from nfmc.algorithms.sampling.mcmc.mh import MH, MHKernel, MHParameters                     # noqa: E402
from nfmc.algorithms.sampling.mcmc.ess import ESS, ESSKernel, ESSParameters                 # noqa: E402
from nfmc.algorithms.sampling.nfmc.imh import FixedIMH, IMHKernel, IMHParameters      # noqa: E402
from nfmc.algorithms.sampling.nfmc.jump import JumpMALA, JumpHMC, JumpESS, JumpNFMCParameters  # noqa: E402
from nfmc.algorithms.sampling.nfmc.neutra import NeuTraHMC, NeuTraMH, NeuTraKernel, NeuTraParameters  # noqa: E402
from nfmc.algorithms.sampling.base import NFMCKernel                                   # noqa: E402
from nfmc.algorithms.sampling.nfmc.tess import TESS, TESSKernel, TESSParameters       # noqa: E402
from nfmc.algorithms.sampling.nfmc.dlmc import DLMC, DLMCKernel, DLMCParameters       # noqa: E402

from oracle.potentials_ref import make_potential_ref                                   # noqa: E402
from oracle.realnvp_ref import make_flow                                               # noqa: E402


class Tape:
    """Record every torch.randn / randn_like / rand / rand_like call the reference makes."""

    def __enter__(self):
        self.normals, self.uniforms = [], []
        self._orig = {k: getattr(torch, k) for k in ("randn", "randn_like", "rand", "rand_like")}

        def randn(*a, **k):
            v = self._orig["randn"](*a, **k)
            self.normals.append(v.detach().clone())
            return v

        def randn_like(*a, **k):
            v = self._orig["randn_like"](*a, **k)
            self.normals.append(v.detach().clone())
            return v

        def rand(*a, **k):
            v = self._orig["rand"](*a, **k)
            self.uniforms.append(v.detach().clone())
            return v

        def rand_like(*a, **k):
            v = self._orig["rand_like"](*a, **k)
            self.uniforms.append(v.detach().clone())
            return v

        torch.randn, torch.randn_like, torch.rand, torch.rand_like = randn, randn_like, rand, rand_like
        return self

    def __exit__(self, *exc):
        for k, v in self._orig.items():
            setattr(torch, k, v)


def pack(out, tape, x0, extra):
    st = out.statistics
    d = dict(
        x0=x0.numpy(),
        samples=out.samples.numpy(),
        last=out.running_samples.last_sample.numpy(),
        mean=np.asarray(out.mean), second_moment=np.asarray(out.second_moment),
        counters=np.array([st.n_accepted_trajectories, st.n_attempted_trajectories, st.n_divergences,
                           st.n_target_gradient_calls, st.n_target_calls,
                           getattr(st, "n_accepted_jumps", -1), getattr(st, "n_attempted_jumps", -1)], dtype=np.int64),
        n_normals=len(tape.normals), n_uniforms=len(tape.uniforms),
    )
    for i, v in enumerate(tape.normals):
        d[f"normal_{i}"] = v.numpy()
    for i, v in enumerate(tape.uniforms):
        d[f"uniform_{i}"] = v.numpy()
    d.update(extra)
    return d


def flow_arrays(flow, n_layers, cond_layers, cond_hidden):
    d = {f"flow/{k}": v.detach().numpy() for k, v in flow.state_dict().items()}
    d["flow_cfg"] = np.array([n_layers, cond_layers, cond_hidden], dtype=np.int64)
    return d


def main():
    torch.set_num_threads(1)
    cases = {}

    # ---- MALA, non-trivial inverse mass (pins Q5: Langevin divides by imd) ------------------------------
    torch.manual_seed(11)
    d, n, K = 7, 5, 4
    target = make_potential_ref("g0", (d,))
    imd = 0.5 + torch.rand(d)
    x0 = torch.randn(n, d)
    s = MALA((d,), target, LangevinKernel(event_size=d, inv_mass_diag=imd.clone(), step_size=0.3),
             LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["mala_g0"] = pack(out, t, x0, dict(pot="g0", step=0.3, imd=imd.numpy(), K=K))

    # ---- MALA on the funnel with default kernel -----------------------------------------------------------
    torch.manual_seed(12)
    d, n, K = 8, 6, 5
    target = make_potential_ref("fn", (d,))
    x0 = 0.3 * torch.randn(n, d)
    kern = LangevinKernel(event_size=d, step_size=0.05)
    s = MALA((d,), target, kern, LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["mala_fn"] = pack(out, t, x0, dict(pot="fn", step=0.05, imd=np.ones(d, np.float32), K=K))

    # ---- HMC on the ill-conditioned Gaussian, non-trivial inverse mass (HMC multiplies by imd) -----------
    torch.manual_seed(13)
    d, n, K, L = 6, 5, 3, 4
    target = make_potential_ref("g1", (d,))
    imd = 0.5 + torch.rand(d)
    x0 = 0.1 * torch.randn(n, d)
    s = HMC((d,), target, HMCKernel(event_size=d, inv_mass_diag=imd.clone(), step_size=0.02, n_leapfrog_steps=L),
            HMCParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["hmc_g1"] = pack(out, t, x0, dict(pot="g1", step=0.02, imd=imd.numpy(), K=K, L=L))

    # ---- HMC on Rosenbrock -------------------------------------------------------------------------------
    torch.manual_seed(14)
    d, n, K, L = 6, 4, 3, 5
    target = make_potential_ref("rb", (d,))
    x0 = 0.5 * torch.randn(n, d)
    s = HMC((d,), target, HMCKernel(event_size=d, step_size=0.01, n_leapfrog_steps=L), HMCParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["hmc_rb"] = pack(out, t, x0, dict(pot="rb", step=0.01, imd=np.ones(d, np.float32), K=K, L=L))

    # ---- jump_mala, frozen perturbed flow ------------------------------------------------------------------
    torch.manual_seed(15)
    d, n, T, K = 6, 4, 2, 3
    target = make_potential_ref("g0", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.1, seed=100)
    x0 = torch.randn(n, d)
    s = JumpMALA((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                 inner_kernel=LangevinKernel(event_size=d), inner_params=LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_mala_g0"] = pack(out, t, x0, dict(pot="g0", step=d ** (-1 / 3), imd=np.ones(d, np.float32), K=K, T=T,
                                                   **flow_arrays(flow, 2, 2, 4)))

    # ---- jump_hmc on the mixture, 3 coupling layers, 3-layer conditioner ----------------------------------
    torch.manual_seed(16)
    d, n, T, K, L = 7, 4, 2, 2, 3
    target = make_potential_ref("gm", (d,))
    flow = make_flow((d,), n_layers=3, conditioner_kwargs=dict(n_layers=3, n_hidden=6), perturb=0.1, seed=101)
    x0 = torch.randn(n, d)
    s = JumpHMC((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                inner_kernel=HMCKernel(event_size=d, step_size=0.05, n_leapfrog_steps=L),
                inner_params=HMCParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_hmc_gm"] = pack(out, t, x0, dict(pot="gm", step=0.05, imd=np.ones(d, np.float32), K=K, T=T, L=L,
                                                  **flow_arrays(flow, 3, 3, 6)))

    # ---- fixed IMH on Rosenbrock ------------------------------------------------------------------------------
    torch.manual_seed(17)
    d, n, T = 6, 8, 4
    target = make_potential_ref("rb", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.1, seed=102)
    x0 = torch.randn(n, d)
    s = FixedIMH((d,), target, IMHKernel((d,), flow=flow), IMHParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["imh_rb"] = pack(out, t, x0, dict(pot="rb", T=T, **flow_arrays(flow, 2, 2, 4)))

    # ---- NeuTra HMC on the funnel -------------------------------------------------------------------------------
    torch.manual_seed(18)
    d, n, T, L = 6, 4, 2, 3
    target = make_potential_ref("fn", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.1, seed=103)
    x0 = 0.5 * torch.randn(n, d)
    s = NeuTraHMC((d,), target, HMCKernel(event_size=d, step_size=0.03, n_leapfrog_steps=L), HMCParameters(),
                  NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["neutra_hmc_fn"] = pack(out, t, x0, dict(pot="fn", step=0.03, imd=np.ones(d, np.float32), T=T, L=L,
                                                    **flow_arrays(flow, 2, 2, 4)))

    # ---- random-walk MH on the mixture, non-trivial proposal scale -------------------------------------------------
    torch.manual_seed(19)
    d, n, K = 7, 6, 5
    target = make_potential_ref("gm", (d,))
    imd = 0.2 + 0.3 * torch.rand(d)
    x0 = torch.randn(n, d)
    s = MH((d,), target, MHKernel(event_size=d, inv_mass_diag=imd.clone()), MHParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["mh_gm"] = pack(out, t, x0, dict(pot="gm", imd=imd.numpy(), K=K))

    # ---- elliptical slice sampling: N(0, I) prior x funnel "likelihood" (x0 only supplies n: ess.py:126) ----------
    torch.manual_seed(20)
    d, n, K, M = 7, 6, 4, 5
    nll = make_potential_ref("fn", (d,))
    x0 = torch.randn(n, d)
    s = ESS((d,), nll, nll, ESSKernel(event_shape=(d,)), ESSParameters(n_iterations=K, max_ess_step_iterations=M))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["ess_fn"] = pack(out, t, x0, dict(pot="fn", K=K, M=M))

    # ---- jump_ess: every local stage restarts from the prior; jump against the mixture target -----------------
    torch.manual_seed(21)
    d, n, T, K, M = 6, 5, 2, 3, 4
    target = make_potential_ref("gm", (d,))
    nll = make_potential_ref("rb", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.1, seed=104)
    x0 = torch.randn(n, d)
    s = JumpESS((d,), target, nll, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                inner_kernel=ESSKernel(event_shape=(d,)),
                inner_params=ESSParameters(n_iterations=K, max_ess_step_iterations=M))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_ess_gm"] = pack(out, t, x0, dict(pot="gm", nll="rb", K=K, T=T, M=M, **flow_arrays(flow, 2, 2, 4)))

    # ---- production shape (d = 100: the exact 4-lane x 13-slot layout, default conditioner H = 5) ------------------------
    torch.manual_seed(22)
    d, n, T, K = 100, 40, 2, 4
    target = make_potential_ref("g1", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.05, seed=105)
    x0 = 0.1 * torch.randn(n, d)
    s = JumpMALA((d,), target, kernel=NFMCKernel((d,), flow=flow), params=JumpNFMCParameters(n_iterations=T),
                 inner_kernel=LangevinKernel(event_size=d, step_size=0.004), inner_params=LangevinParameters(n_iterations=K))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["jump_mala_g1_d100"] = pack(out, t, x0, dict(pot="g1", step=0.004, imd=np.ones(d, np.float32), K=K, T=T,
                                                        **flow_arrays(flow, 2, 2, 5)))

    torch.manual_seed(23)
    d, n, T, L = 100, 12, 2, 3
    target = make_potential_ref("fn", (d,))
    flow = make_flow((d,), n_layers=2, perturb=0.05, seed=106)
    x0 = 0.3 * torch.randn(n, d)
    s = NeuTraHMC((d,), target, HMCKernel(event_size=d, step_size=0.01, n_leapfrog_steps=L), HMCParameters(),
                  NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["neutra_hmc_fn_d100"] = pack(out, t, x0, dict(pot="fn", step=0.01, imd=np.ones(d, np.float32), T=T, L=L,
                                                         **flow_arrays(flow, 2, 2, 5)))

    # ---- NeuTra MH: random-walk Metropolis in the latent space, non-trivial proposal scale ---------------------------------
    torch.manual_seed(24)
    d, n, T = 7, 6, 5
    target = make_potential_ref("gm", (d,))
    flow = make_flow((d,), n_layers=3, perturb=0.1, seed=107)
    imd = 0.2 + 0.3 * torch.rand(d)
    x0 = 0.5 * torch.randn(n, d)
    s = NeuTraMH((d,), target, MHKernel(event_size=d, inv_mass_diag=imd.clone()), MHParameters(),
                 NeuTraKernel((d,), flow=flow), NeuTraParameters(n_iterations=T))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["neutra_mh_gm"] = pack(out, t, x0, dict(pot="gm", imd=imd.numpy(), T=T, **flow_arrays(flow, 3, 2, 4)))

    # ---- transport elliptical slice sampling, frozen flow (3 couplings: the latent is flipped in physical order) ---------
    torch.manual_seed(25)
    d, n, T, M = 7, 6, 4, 5
    nll = make_potential_ref("fn", (d,))
    flow = make_flow((d,), n_layers=3, perturb=0.1, seed=108)
    x0 = 0.5 * torch.randn(n, d)
    s = TESS((d,), nll, nll, TESSKernel((d,), flow=flow), TESSParameters(n_iterations=T, max_ess_step_iterations=M))
    with Tape() as t:
        out = s.sample(x0.clone(), show_progress=False)
    cases["tess_fn"] = pack(out, t, x0, dict(pot="fn", T=T, M=M, **flow_arrays(flow, 3, 2, 4)))

    # ---- deterministic Langevin Monte Carlo with the per-iteration refit switched off (the refit is torchflows' optimiser,
    #      absent here; everything else of dlmc.py:44-119 is pinned), both update rules ---------------------------------
    for tag, latent, seed in (("dlmc_gm", False, 26), ("dlmc_latent_gm", True, 27)):
        torch.manual_seed(seed)
        d, n, T = 6, 7, 4
        target = make_potential_ref("gm", (d,))
        nll = make_potential_ref("g0", (d,))
        flow = make_flow((d,), n_layers=2, perturb=0.1, seed=109)
        flow.fit = lambda *a, **k: None
        x0 = torch.randn(n, d)
        s = DLMC((d,), target, nll, DLMCKernel((d,), flow=flow, step_size=0.05),
                 DLMCParameters(n_iterations=T, latent_updates=latent))
        with Tape() as t:
            out = s.sample(x0.clone(), show_progress=False)
        cases[tag] = pack(out, t, x0, dict(pot="gm", nll="g0", T=T, step=0.05, latent=int(latent), **flow_arrays(flow, 2, 2, 4)))

    for name, arrays in cases.items():
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **arrays)
        print(f"wrote {path}  ({os.path.getsize(path)} B)")


if __name__ == "__main__":
    main()
