"""Time nfmc_mala_steps (n = 2^20, d = 100, K = 100, iso Gaussian) for the library named by NFMC_B200_LIB."""
import ctypes as C
import sys
import torch
from nfmc_b200 import _native as N

lib = N.lib()
n, d, K = 1 << 20, 100, 100
x = torch.randn(n, d, device="cuda")
pot = N.PotentialDesc(kind=0, d=d, params=None)
pot.scalar[0] = 2.0
sx = torch.zeros(d, device="cuda", dtype=torch.float64); sx2 = torch.zeros_like(sx)
cnt = torch.zeros(8, device="cuda", dtype=torch.int64)
st = N.StatsDesc(sum_x=sx.data_ptr(), sum_x2=sx2.data_ptr(), counts=cnt.data_ptr())
rng = N.RngDesc(seed=1, step0=0, normals=None, uniforms=None)
sink = N.SinkDesc(samples=None, seen0=0, thinning=1)
s = torch.cuda.current_stream().cuda_stream
def go():
    rc = lib.nfmc_mala_steps(C.byref(pot), x.data_ptr(), n, K, d ** (-1 / 3), None, 1, C.byref(rng), 0, C.byref(st), C.byref(sink), s)
    assert rc == 0, lib.nfmc_last_error()
go(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(sys.argv[1] if len(sys.argv) > 1 else "", "ms/launch", " ".join(f"{t:.2f}" for t in ts), "acc", int(cnt[0]), "chk", float(x.double().sum()))
