import sys, time, torch, cProfile, pstats
sys.path.insert(0, "/root/repo")
import nfmc_b200
from nfmc_b200.potentials import make_potential
from nfmc_b200.flow import create_flow_object
torch.manual_seed(0)
d, n, T = 100, 1 << 17, 24
flow = create_flow_object("realnvp", (d,))
s = nfmc_b200.create_sampler(make_potential("rb", (d,)), flow=flow, strategy="adaptive_imh")
s.params.n_iterations = T
x0 = torch.randn(n, d, device="cuda") * 0.5
s.sample(x0, show_progress=False)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
out = s.sample(x0, show_progress=False)
torch.cuda.synchronize()
pr.disable()
print("wall per iteration ms", 1e3 * (time.perf_counter() - t0) / T)
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
