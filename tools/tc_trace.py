#!/usr/bin/env python
"""Timeline of the tensor-core flow kernel on one SM (debug build: make -C nfmc_b200/csrc trace -> libnfmc_b200_trace.so).
Prints, for CTA 0, the clock of every control-lane and epilogue-thread-0 event of the first pairs."""
import ctypes as C, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["NFMC_B200_LIB"] = os.path.join(ROOT, "nfmc_b200", "libnfmc_b200_trace.so")
sys.path.insert(0, ROOT)
import torch
from nfmc_b200 import _native as N
from nfmc_b200.flow import Flow, RealNVP

H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Lc = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else (1 << 20)
flow = Flow(RealNVP((100,), n_layers=Lc, conditioner_kwargs=dict(n_layers=2, n_hidden=H), conditioner_dtype="bf16")).cuda()
x = torch.randn(n, 100, device="cuda")
for _ in range(2):
    flow.bijection.forward(x)
torch.cuda.synchronize()
buf = torch.zeros(2 * 2048 * 2, dtype=torch.int64, device="cuda")
lib = N.lib()
lib.nfmc_tc_trace_set.argtypes = [C.c_void_p]
lib.nfmc_tc_trace_set(buf.data_ptr())
flow.bijection.forward(x)
torch.cuda.synchronize()
b = buf.cpu().reshape(2, 2048, 2)
ev = []
for who in range(2):
    for i in range(2048):
        if b[who, i, 1] == 0:
            break
        ev.append((int(b[who, i, 1]), who, int(b[who, i, 0])))
ev.sort()
t0 = ev[0][0]
names = {0: "c.begin", 1: "c.w1_full", 2: "c.a1[0]", 3: "c.G1(0)issued", 4: "c.a1[1]", 5: "c.G1(1)issued", 6: "c.wl_full", 7: "c.hid[0]",
         8: "c.G2(0)issued", 9: "c.w1refill", 10: "c.hid[1]", 11: "c.G2(1)issued", 12: "c.end", 20: "e.wait g1[0]", 21: "e.wait g1[1]",
         22: "e.g1[0] ok", 23: "e.g1[1] ok", 24: "e.E1(0) done", 25: "e.E1(1) done", 26: "e.g2[0] ok", 27: "e.g2[1] ok", 28: "e.E2(0) done",
         29: "e.E2(1) done", 43: "e.xfull(0)", 44: "e.xfull(1)", 45: "e.xready(0)", 46: "e.xready(1)", 47: "e.begun(0)", 48: "e.begun(1)",
         49: "e.exchanged(0)", 50: "e.exchanged(1)", 51: "e.fenced(0)", 52: "e.fenced(1)", 53: "e.affine0(0)", 54: "e.affine0(1)",
         55: "e.a1 free0(0)", 56: "e.a1 free0(1)", 57: "e.a1 written(0)", 58: "e.a1 written(1)", 59: "e.a1 fenced(0)", 60: "e.a1 fenced(1)",
         39: "e.pair begin", 40: "e.loaded", 41: "e.pass done", 42: "e.pair done"}
last = t0
for t, who, i in ev[:int(os.environ.get("TRACE_ROWS", "160"))]:
    print(f"{t - t0:9d} (+{t - last:6d}) {'ctl' if who == 0 else 'epi'} {names.get(i, i)}")
    last = t
