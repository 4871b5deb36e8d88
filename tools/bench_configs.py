#!/usr/bin/env python
"""Chain-steps/s of every BASELINE.json config that fits one GPU, through the public API (device-resident state,
store_samples=False).  One JSON line per config.  Not the driver's bench (that is bench.py); numbers go to DESIGN.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import nfmc_b200
from nfmc_b200.potentials import make_potential
from nfmc_b200.flow import create_flow_object


def flow_for(d, spec="realnvp"):
    torch.manual_seed(0)
    f = create_flow_object(spec, (d,))
    with torch.no_grad():
        for p in f.parameters():
            p.add_(0.05 * torch.randn_like(p))
    return f


def run(name, strategy, pot, d, n, T, K=None, flow_spec="realnvp", adapt=False, fit_nf=False, **kw):
    flow = flow_for(d, flow_spec)
    ik = {} if K is None else {"inner_param_kwargs": {"n_iterations": K}}
    pk = {"n_iterations": T, "store_samples": False}
    if fit_nf:
        pk.update(fit_nf=True, n_jumps_before_training=0)
    if strategy in ("ess", "jump_ess", "tess", "dlmc"):
        kw = dict(kw, negative_log_likelihood=make_potential(pot, (d,)))
    s = nfmc_b200.create_sampler(make_potential(pot, (d,)), flow=None if strategy == "ess" else flow, strategy=strategy,
                                 param_kwargs=pk, **ik, **kw)
    if hasattr(s, "adapt"):
        s.adapt = adapt                      # False: MH part only (all iterations fused into one launch)
    if strategy == "adaptive_imh":
        s.params.n_iterations = T            # the reference drops param_kwargs for adaptive_imh (quirk Q2)
    x0 = torch.randn(n, d, device="cuda") * 0.5
    s.params.n_iterations = max(1, T // 4)
    s.sample(x0, show_progress=False)                       # warm-up launch
    s.params.n_iterations = T
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = s.sample(x0, show_progress=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = out.statistics.expectations.n_seen
    st = out.statistics
    print(json.dumps({"config": name, "strategy": strategy, "potential": pot, "d": d, "chains": n, "iterations": T, "inner": K,
                      "flow": flow_spec, "tensor_cores": flow.bijection.uses_tensor_cores(), "seconds": dt,
                      "chain_steps_per_s": steps / dt, "device_seconds": st.elapsed_time_seconds,
                      "acc_rate": st.acceptance_rate, "jump_acc_rate": getattr(st, "jump_acceptance_rate", None)}), flush=True)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None, help="run only the configs whose name contains this string")
    only = ap.parse_args().only
    _run = run

    def run(name, *a, **k):      # noqa: F811
        if only is None or only in name:
            _run(name, *a, **k)

    wide = 'realnvp%{"n_layers": 4, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 256}}'
    run("C1 README jump_mala d=25 n=100", "jump_mala", "g0", 25, 100, 200, K=100)
    run("C2 jump_hmc d=100 n=65536", "jump_hmc", "g1", 100, 65536, 20, K=5)
    run("C3 neutra_hmc funnel d=100 n=262144 (default flow, conditioner_dtype='auto' -> tcgen05 at this batch size)", "neutra_hmc", "fn", 100,
        262144, 3, inner_kernel_kwargs={"step_size": 0.01})
    run("C3 neutra_hmc funnel d=100 n=262144, conditioner_dtype='fp32' (CUDA-core kernel)", "neutra_hmc", "fn", 100, 262144, 3,
        flow_spec='realnvp%{"conditioner_dtype": "fp32"}', inner_kernel_kwargs={"step_size": 0.01})
    run("C4 imh rosenbrock d=100 n=2^20", "imh", "rb", 100, 1 << 20, 20)
    run("C4 adaptive_imh(no refit) d=100 n=2^17", "adaptive_imh", "rb", 100, 1 << 17, 100)
    run("C4 adaptive_imh WITH per-iteration refit d=100 n=2^17", "adaptive_imh", "rb", 100, 1 << 17, 24, adapt=True)
    run("C5-shape jump_mala mixture d=1000 n=131072 fit_nf=True (flow refit every outer iteration)", "jump_mala", "gm", 1000,
        131072, 2, K=100, fit_nf=True)
    run("C5-shape jump_mala mixture d=1000 n=131072 (1/8 of 2^20, frozen flow)", "jump_mala", "gm", 1000, 131072, 2, K=100)
    run("CT jump_mala d=100 n=2^20", "jump_mala", "g0", 100, 1 << 20, 5, K=100)
    run("ess funnel-likelihood d=100 n=2^20", "ess", "fn", 100, 1 << 20, 20)
    run("tess funnel d=100 n=2^18 (frozen flow)", "tess", "fn", 100, 1 << 18, 10)
    run("neutra_mh funnel d=100 n=2^20", "neutra_mh", "fn", 100, 1 << 20, 20)
    run("dlmc mixture d=100 n=2^17 (refit every iteration)", "dlmc", "gm", 100, 1 << 17, 4)
    run("C3-wide neutra_hmc funnel d=100 n=262144 H=256 Lc=4 (tcgen05 forward + dgrad)", "neutra_hmc", "fn", 100, 262144, 3,
        flow_spec=wide, inner_kernel_kwargs={"step_size": 0.01})
    run("C3-wide neutra_hmc funnel d=100 n=262144 H=64 Lc=2 (tcgen05 forward + dgrad)", "neutra_hmc", "fn", 100, 262144, 3,
        flow_spec='realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}', inner_kernel_kwargs={"step_size": 0.01})
    run("C3 neutra_hmc funnel d=100 n=262144 DEFAULT conditioner opted into bf16 (tcgen05 forward + dgrad, H = 5 padded to 32)", "neutra_hmc",
        "fn", 100, 262144, 3, flow_spec='realnvp%{"conditioner_dtype": "bf16"}', inner_kernel_kwargs={"step_size": 0.01})
    run("C4 imh rosenbrock d=100 n=2^20 DEFAULT conditioner opted into bf16 (tcgen05 fused IMH iteration)", "imh", "rb", 100, 1 << 20, 20,
        flow_spec='realnvp%{"conditioner_dtype": "bf16"}')
    deep = 'realnvp%{"n_layers": 10, "conditioner_kwargs": {"n_layers": 5, "n_hidden": 100}}'
    run("deep-flow imh d=100 n=2^16 Lc=10 M=5 H=100 (the reference's test_flow_kwargs shape; row-tile fp32 passes)", "imh", "g0", 100,
        1 << 16, 5, flow_spec=deep)
    run("deep-flow jump_mala K=10 d=100 n=2^16 Lc=10 M=5 H=100 (row-tile fp32 passes)", "jump_mala", "g0", 100, 1 << 16, 3, K=10,
        flow_spec=deep)
    run("deep-flow neutra_hmc d=100 n=2^14 Lc=10 M=5 H=100 L=20 (row-tile fp32 pass + sweep)", "neutra_hmc", "fn", 100, 1 << 14, 2,
        flow_spec=deep, inner_kernel_kwargs={"step_size": 0.01})
    run("odd-d wide-flow imh d=101 n=2^18 Lc=2 H=64 (row-tile fp32 passes)", "imh", "g0", 101, 1 << 18, 5,
        flow_spec='realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}')
    run("wide-flow jump_mala d=100 n=2^20 H=256 Lc=4 (tcgen05)", "jump_mala", "g0", 100, 1 << 20, 5, K=100, flow_spec=wide)
    run("wide-flow imh d=100 n=2^20 H=256 Lc=4 (tcgen05)", "imh", "g0", 100, 1 << 20, 10, flow_spec=wide)
    run("wide-flow jump_mala K=1 (jump-dominated) d=100 n=2^20 H=256 Lc=4 (tcgen05)", "jump_mala", "g0", 100, 1 << 20, 10, K=1, flow_spec=wide)
