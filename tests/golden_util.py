"""Helpers shared by the CPU and GPU parity tests: load a golden fixture and rebuild its objects."""
import os

import numpy as np
import torch

from oracle.potentials_ref import make_potential_ref
from oracle.realnvp_ref import FlowRef, RealNVPRef
from oracle.samplers_ref import TapeDraws

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["mala_g0", "mala_fn", "hmc_g1", "hmc_rb", "jump_mala_g0", "jump_hmc_gm", "imh_rb", "neutra_hmc_fn", "mh_gm",
         "ess_fn", "jump_ess_gm", "jump_mala_g1_d100", "neutra_hmc_fn_d100",
         "neutra_mh_gm", "tess_fn", "dlmc_gm", "dlmc_latent_gm",
         # round 2 (make_golden_r02.py): warm-up trajectories, unadjusted kernels, adaptive IMH, BASELINE config shapes
         "mala_tune_g1", "hmc_tune_fn", "ula_g0", "uhmc_gm", "adaptive_imh_rb", "jump_hmc_g1_d100", "imh_rb_d100",
         "jump_mala_gm_d1000"]


def load_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    g = {k: z[k] for k in z.files}
    g["normals"] = [torch.from_numpy(g[f"normal_{i}"]) for i in range(int(g["n_normals"]))]
    g["uniforms"] = [torch.from_numpy(g[f"uniform_{i}"]) for i in range(int(g["n_uniforms"]))]
    return g


def tape(g):
    # the reference draws some uniforms as [n, 1] (mcmc/ess.py:39,58): same numbers, flattened for the tape
    return TapeDraws(g["normals"], [u.reshape(-1) for u in g["uniforms"]])


def oracle_flow(g):
    n_layers, cond_layers, cond_hidden = (int(v) for v in g["flow_cfg"])
    d = g["x0"].shape[1]
    flow = FlowRef(RealNVPRef((d,), n_layers=n_layers,
                              conditioner_kwargs=dict(n_layers=cond_layers, n_hidden=cond_hidden)))
    sd = {k[len("flow/"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("flow/")}
    flow.load_state_dict(sd)
    return flow.eval()


def oracle_target(g):
    return make_potential_ref(str(g["pot"]), (g["x0"].shape[1],))
