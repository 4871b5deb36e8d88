import ctypes as C, sys, time, torch
sys.path.insert(0, "/root/repo")
from nfmc_b200 import _native as N
from nfmc_b200.flow import Flow, RealNVP
from nfmc_b200.potentials import make_potential
dev = torch.device("cuda:0"); lib = N.lib()
d, n = 100, 1 << 20
torch.manual_seed(0)
flow = Flow(RealNVP((d,), n_layers=2))
with torch.no_grad():
    for p in flow.parameters(): p.add_(0.1 * torch.randn_like(p))
flow = flow.to(dev)
pd, kp = make_potential("g0", (d,)).descriptor(dev)
fd, kf = flow.bijection.descriptor(dev)
x = torch.randn(n, d, device=dev); xp = torch.empty_like(x)
lq = torch.empty(n, device=dev); lq2 = torch.empty(n, device=dev)
mom = torch.zeros(2 * d, device=dev, dtype=torch.float64); cnt = torch.zeros(8, device=dev, dtype=torch.int64)
st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
s = N.stream_ptr(dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
rng = N.rng_desc(1, 0, None, None)
print("fused jump        ms", timeit(lambda: N.check(lib.nfmc_jump_step(C.byref(pd), C.byref(fd), N.ptr(x), n, 1, C.byref(rng), 0, C.byref(st), None, s))))
print("flow_log_prob     ms", timeit(lambda: N.check(lib.nfmc_flow_log_prob(C.byref(fd), N.ptr(x), N.ptr(lq), n, s))))
print("flow_sample       ms", timeit(lambda: N.check(lib.nfmc_flow_sample(C.byref(fd), C.byref(rng), 0, N.ptr(xp), N.ptr(lq2), n, s))))
u = torch.empty(n, device=dev); g = torch.empty_like(x)
print("potential_eval    ms", timeit(lambda: N.check(lib.nfmc_potential_eval(C.byref(pd), N.ptr(x), N.ptr(u), None, n, s))))
