"""Does splitting the chains into slabs on several streams (kernels of different slabs overlapping) beat one stream?"""
import ctypes as C, sys, time, torch
sys.path.insert(0, "/root/repo")
from nfmc_b200 import _native as N
from nfmc_b200.flow import Flow, RealNVP
from nfmc_b200.potentials import make_potential
dev = torch.device("cuda:0")
lib = N.lib()
d, n, K, T = 100, 1 << 20, 100, 4
torch.manual_seed(0)
flow = Flow(RealNVP((d,), n_layers=2))
with torch.no_grad():
    for p in flow.parameters():
        p.add_(0.1 * torch.randn_like(p))
flow = flow.to(dev)
pot = make_potential("g0", (d,))
pd, kp = pot.descriptor(dev)
fd, kf = flow.bijection.descriptor(dev)
x = torch.randn(n, d, device=dev)
mom = torch.zeros(2 * d, device=dev, dtype=torch.float64)
cnt = torch.zeros(8, device=dev, dtype=torch.int64)
st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
sj = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr() + 32)
tau = d ** (-1 / 3)

def run(slab, n_streams):
    streams = [torch.cuda.Stream(dev) for _ in range(n_streams)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    first, si = 0, 0
    while first < n:
        c = min(slab, n - first)
        s = streams[si % n_streams]
        sp = C.c_void_p(s.cuda_stream)
        xp = C.c_void_p(x.data_ptr() + first * d * 4)
        for it in range(T):
            r1 = N.rng_desc(1, it * K, None, None)
            r2 = N.rng_desc(1, it, None, None)
            N.check(lib.nfmc_mala_steps(C.byref(pd), xp, c, K, tau, None, 1, C.byref(r1), first, C.byref(st), None, sp))
            N.check(lib.nfmc_jump_step(C.byref(pd), C.byref(fd), xp, c, 1, C.byref(r2), first, C.byref(sj), None, sp))
        first += c
        si += 1
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / T * 1e3

for slab, ns in [(n, 1), (n // 2, 2), (n // 4, 2), (n // 4, 4), (56832 * 2, 3), (56832, 3), (56832, 4), (56832 * 4, 3)]:
    run(slab, ns)
    print(f"slab {slab:8d} streams {ns}: {run(slab, ns):7.3f} ms per outer iteration")
