#!/usr/bin/env python
"""Throughput of the RealNVP operators alone (flow.log_prob / forward / inverse) on one GPU.

    python tools/bench_flow.py --dim 100 --layers 4 --hidden 256 --dtype bf16     # tcgen05 path
    python tools/bench_flow.py --dim 100 --layers 2                              # default conditioner, CUDA cores
Prints one JSON line: chains/s, algorithmic TFLOP/s (2*Lc*(d/2*H + H*2*(d-d/2)) MAC per pass) and GB/s (8*d+4 B/chain).
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nfmc_b200.flow import Flow, RealNVP


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--dim", type=int, default=100)
    p.add_argument("--layers", type=int, default=2)
    p.add_argument("--hidden", type=int, default=None)
    p.add_argument("--dtype", default="auto")
    p.add_argument("--chains", type=int, default=1 << 20)
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--op", default="forward", choices=["forward", "inverse", "log_prob"])
    a = p.parse_args()
    ck = None if a.hidden is None else dict(n_layers=2, n_hidden=a.hidden)
    torch.manual_seed(0)
    flow = Flow(RealNVP((a.dim,), n_layers=a.layers, conditioner_kwargs=ck, conditioner_dtype=a.dtype))
    with torch.no_grad():
        for q in flow.parameters():
            q.add_(0.05 * torch.randn_like(q))
    flow = flow.cuda()
    x = torch.randn(a.chains, a.dim, device="cuda")
    fn = {"forward": flow.bijection.forward, "inverse": flow.bijection.inverse, "log_prob": flow.log_prob}[a.op]
    for _ in range(3):
        fn(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    M, H = flow.bijection.conditioner_shape()
    da, db = a.dim // 2, a.dim - a.dim // 2
    flops = 2.0 * a.layers * (da * H + H * 2 * db) * a.chains
    bytes_ = (8.0 * a.dim + 4) * a.chains
    print(json.dumps({"op": a.op, "d": a.dim, "Lc": a.layers, "H": H, "tensor_cores": flow.bijection.uses_tensor_cores(),
                      "chains": a.chains, "ms": ms, "chains_per_s": a.chains / ms * 1e3, "tflops": flops / ms / 1e9,
                      "gbs": bytes_ / ms / 1e6}))


if __name__ == "__main__":
    main()
