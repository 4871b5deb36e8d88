import sys, time, torch
sys.path.insert(0, "/root/repo")
import nfmc_b200 as nfmc
from nfmc_b200.potentials import StandardGaussian
torch.manual_seed(0)
for show in (False, True):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = nfmc.sample(StandardGaussian((25,)), strategy='jump_mala', flow='realnvp', n_chains=100, n_iterations=1000, show_progress=show)
        s = out.samples
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print("show_progress", show, "wall", round(dt, 3), "s; samples", tuple(s.shape), "device s", round(out.statistics.elapsed_time_seconds, 3),
          "chain-steps/s", round(s.shape[0] * s.shape[1] / dt))
