"""Time flow training: native kernels (csrc/train_kernels.cu) vs the torch-autograd loop, same data and settings.

    python tools/bench_train.py [--d 100] [--n 4096] [--epochs 30]

Settings follow the reference's fit call (jump.py:193-201): <= 4096 training rows, <= 4096 validation rows,
batch_size='adaptive', lr=0.05."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nfmc_b200.flow import Flow, RealNVP  # noqa: E402
from nfmc_b200 import potentials as P     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", type=int, default=100)
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=30)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(0)
    x = (torch.linspace(0.5, 2.0, a.d) * torch.randn(a.n, a.d, generator=g)).cuda()
    xv = (torch.linspace(0.5, 2.0, a.d) * torch.randn(a.n, a.d, generator=g)).cuda()
    res = {"d": a.d, "n_train": a.n, "epochs": a.epochs}
    for mode in ("native", "library"):
        os.environ["NFMC_B200_LIBRARY_TRAINING"] = "1" if mode == "library" else "0"
        for rep in range(2):                       # first repetition warms up
            torch.manual_seed(1)
            f = Flow(RealNVP((a.d,), n_layers=2)).to("cuda")
            torch.cuda.synchronize()
            t0 = time.time()
            f.fit(x, x_val=xv, n_epochs=a.epochs, lr=0.05, batch_size="adaptive")
            torch.cuda.synchronize()
            dt = time.time() - t0
        res[f"{mode}_ms_per_epoch"] = 1e3 * dt / a.epochs
        res[f"{mode}_val_nll"] = float(-f.log_prob(xv).mean())
        pot = P.make_potential("g1", (a.d,))
        torch.manual_seed(2)
        f = Flow(RealNVP((a.d,), n_layers=2)).to("cuda")
        torch.cuda.synchronize()
        t0 = time.time()
        f.variational_fit(pot.log_prob_fn(), n_epochs=100, lr=0.05, n_samples=256)
        torch.cuda.synchronize()
        res[f"{mode}_kl_ms_per_step"] = 1e3 * (time.time() - t0) / 100
    res["fit_speedup"] = res["library_ms_per_epoch"] / res["native_ms_per_epoch"]
    res["kl_speedup"] = res["library_kl_ms_per_step"] / res["native_kl_ms_per_step"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
