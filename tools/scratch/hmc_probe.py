import sys, time, torch
sys.path.insert(0, "/root/repo")
import nfmc_b200
from nfmc_b200.potentials import make_potential
for pot in ("g1", "g0", "fn", "rb", "gm"):
    d, n, T = 100, 1 << 20, 10
    s = nfmc_b200.create_sampler(make_potential(pot, (d,)), flow=None, strategy="hmc",
                                 param_kwargs={"n_iterations": T, "store_samples": False}, kernel_kwargs={"step_size": 0.01})
    s.seed = 5
    x0 = torch.randn(n, d, device="cuda") * 0.3
    s.sample(x0, show_progress=False)
    out = s.sample(x0, show_progress=False)
    print(pot, "ms per HMC step (L=20, 2^20 chains):", round(1e3 * out.statistics.elapsed_time_seconds / T, 4))
