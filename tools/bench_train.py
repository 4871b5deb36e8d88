"""Time flow training: the native kernels (csrc/train_kernels.cu for the default conditioners, csrc/train_wide.cu for wide /
deep ones) against a torch-autograd + torch.optim.AdamW loop written HERE over the same torch restatement (the product
has no such loop), same data and settings.

    python tools/bench_train.py [--d 100] [--n 4096] [--epochs 30] [--hidden H] [--cond-layers M] [--layers Lc]

Settings follow the reference's fit call (jump.py:193-201): <= 4096 training rows, <= 4096 validation rows,
batch_size='adaptive', lr=0.05."""
import argparse
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nfmc_b200.flow import Flow, RealNVP  # noqa: E402
from nfmc_b200 import flow_train as FT    # noqa: E402
from nfmc_b200 import potentials as P     # noqa: E402


def autograd_fit(flow, x, xv, n_epochs, lr, batch_size):
    params = [p for p in flow.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=lr)
    n = len(x)
    flow.train()
    with torch.enable_grad():
        for _ in range(n_epochs):
            perm = torch.randperm(n, device=x.device)
            for i in range(0, n, batch_size):
                opt.zero_grad(set_to_none=True)
                loss = -FT.log_prob_autograd(flow, x[perm[i:i + batch_size]], training=True).mean()
                loss.backward()
                opt.step()
            with torch.no_grad():
                float(-FT.log_prob_autograd(flow, xv).mean())          # the epoch's validation score (one sync)
    flow.eval()


def autograd_kl(flow, potential, n_epochs, lr, n_samples):
    params = [p for p in flow.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=lr)
    d = flow.bijection.n_dim
    tlp = potential.log_prob_fn()
    with torch.enable_grad():
        for _ in range(n_epochs):
            opt.zero_grad(set_to_none=True)
            z = torch.randn(n_samples, d, device="cuda")
            x, ld = FT.inverse_autograd(flow.bijection, z)
            log_q = (-0.5 * z.square()).sum(dim=1) - 0.5 * d * math.log(2 * math.pi) - ld
            loss = (log_q - tlp(x)).mean()
            loss.backward()
            opt.step()
            float(loss.detach())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--hidden", type=int, default=None)
    ap.add_argument("--cond-layers", type=int, default=2)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--lr", type=float, default=0.05)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    x = (torch.linspace(0.5, 2.0, a.d) * torch.randn(a.n, a.d, generator=g)).cuda()
    xv = (torch.linspace(0.5, 2.0, a.d) * torch.randn(a.n, a.d, generator=g)).cuda()
    ck = None if a.hidden is None else dict(n_layers=a.cond_layers, n_hidden=a.hidden)
    res = {"lr": a.lr, "d": a.d, "n_train": a.n, "epochs": a.epochs, "Lc": a.layers, "conditioner": ck or "default"}
    bs = max(32, min(1024, a.n // 10 if a.n >= 320 else a.n))
    for mode in ("native", "autograd"):
        for rep in range(2):                       # first repetition warms up
            torch.manual_seed(1)
            f = Flow(RealNVP((a.d,), n_layers=a.layers, conditioner_kwargs=ck)).to("cuda")
            torch.cuda.synchronize()
            t0 = time.time()
            if mode == "native":
                f.fit(x, x_val=xv, n_epochs=a.epochs, lr=a.lr, batch_size="adaptive")
            else:
                autograd_fit(f, x, xv, a.epochs, a.lr, bs)
            torch.cuda.synchronize()
            dt = time.time() - t0
        res[f"{mode}_ms_per_epoch"] = 1e3 * dt / a.epochs
        f.bijection._packed.clear(); f.bijection._packed_tc.clear()
        res[f"{mode}_val_nll"] = float(-FT.log_prob_autograd(f, xv).detach().mean())
        pot = P.make_potential("g1", (a.d,))
        torch.manual_seed(2)
        f = Flow(RealNVP((a.d,), n_layers=a.layers, conditioner_kwargs=ck)).to("cuda")
        torch.cuda.synchronize()
        t0 = time.time()
        if mode == "native":
            f.variational_fit(pot.log_prob_fn(), n_epochs=100, lr=a.lr, n_samples=256)
        else:
            autograd_kl(f, pot, 100, a.lr, 256)
        torch.cuda.synchronize()
        res[f"{mode}_kl_ms_per_step"] = 1e3 * (time.time() - t0) / 100
    res["fit_speedup"] = res["autograd_ms_per_epoch"] / res["native_ms_per_epoch"]
    res["kl_speedup"] = res["autograd_kl_ms_per_step"] / res["native_kl_ms_per_step"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
