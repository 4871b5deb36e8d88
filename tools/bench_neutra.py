import sys, os, time, json
sys.path.insert(0, os.getcwd())
import torch, nfmc_b200
from nfmc_b200.potentials import make_potential
from nfmc_b200.flow import create_flow_object
torch.manual_seed(0)
d, n, T = 100, 262144, 4
f = create_flow_object("realnvp", (d,))
with torch.no_grad():
    for p in f.parameters(): p.add_(0.05 * torch.randn_like(p))
s = nfmc_b200.create_sampler(make_potential("fn", (d,)), flow=f, strategy="neutra_hmc", param_kwargs={"n_iterations": T, "store_samples": False}, inner_kernel_kwargs={"step_size": 0.01})
x0 = torch.randn(n, d, device="cuda") * 0.5
s.sample(x0, show_progress=False)
out = s.sample(x0, show_progress=False)
print(json.dumps({"neutra_hmc chain-steps/s": n * T / out.statistics.elapsed_time_seconds, "ms_per_iteration": 1e3 * out.statistics.elapsed_time_seconds / T, "acc": out.statistics.acceptance_rate}))
