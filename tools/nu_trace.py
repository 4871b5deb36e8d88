#!/usr/bin/env python
"""Timeline of the tensor-core NeuTra unwind kernel on one SM (debug build with -DNFMC_NU_TRACE -> libnfmc_b200_nutrace.so):
clock of every phase boundary of epilogue thread 0 of CTA 0, first tiles.   python tools/nu_trace.py [H] [Lc]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NFMC_B200_LIB"] = os.path.join(ROOT, "nfmc_b200", "libnfmc_b200_nutrace.so")
sys.path.insert(0, ROOT)
import torch
from nfmc_b200 import _native as N
from nfmc_b200 import potentials as P
from nfmc_b200.flow import Flow, RealNVP

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Lc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d, n = 100, 262144
flow = Flow(RealNVP((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=2, n_hidden=H), conditioner_dtype="bf16")).cuda()
dev = torch.device("cuda")
tgt = P.make_potential("fn", (d,))
pd, k1 = tgt.descriptor(dev)
td, k2 = flow.bijection.tc_descriptor(dev)
bt = flow.bijection.tc_transposed(dev)
z = 0.5 * torch.randn(n, d, device=dev)
x, ld, u, g = torch.empty(n, d, device=dev), torch.empty(n, device=dev), torch.empty(n, device=dev), torch.empty(n, d, device=dev)
lib = N.lib()
FUSED = os.environ.get("NU_FUSED", "1") == "1"       # the production path: leapfrog update fused (one HMC step, L = 2)
if FUSED:
    nb = lib.nfmc_neutra_tc_workspace_bytes(d, n)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    mom = torch.zeros(2 * d, device=dev, dtype=torch.float64)
    cnt = torch.zeros(8, device=dev, dtype=torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    rng = N.rng_desc(1, 0, None, None)
    call = lambda: N.check(lib.nfmc_neutra_hmc_steps_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(z), n, 1, 0.01, 2, None, 1,
                                                        C.byref(rng), 0, C.byref(st), None, N.ptr(ws), nb, N.stream_ptr(dev)))
else:
    call = lambda: N.check(lib.nfmc_neutra_potential_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(z), N.ptr(x), N.ptr(ld), N.ptr(u), N.ptr(g), n, N.stream_ptr(dev)))
for _ in range(2):
    call()
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device=dev)
lib.nfmc_nu_trace_set.argtypes = [C.c_void_p]
lib.nfmc_nu_trace_set(buf.data_ptr())
call()
torch.cuda.synchronize()
b = buf.cpu().reshape(2048, 2)
names = {0: "tile begin", 1: "x tile landed", 2: "potential + affine0 done", 10: "A1 written", 11: "G1 done", 12: "epi1 done (tanh)", 13: "G2 done",
         14: "epi2 done (dU)", 15: "G3 done", 16: "epi3 done (dpre)", 17: "G4 done", 18: "epi4 + affine done", 3: "grad ready", 4: "p/z tiles landed", 5: "leapfrog done"}
t0 = int(b[0, 1]); last = t0
for i in range(int(os.environ.get("TRACE_ROWS", "110"))):
    if b[i, 1] == 0:
        break
    t = int(b[i, 1])
    print(f"{t - t0:9d} (+{t - last:6d}) {names.get(int(b[i, 0]), int(b[i, 0]))}")
    last = t
