import time, torch, nfmc_b200
from nfmc_b200.potentials import StandardGaussian
torch.manual_seed(0)
n, d = 1 << 18, 100
for name, tgt in [("fused", StandardGaussian((d,))), ("callable", lambda x: torch.sum(x ** 2, dim=1))]:
    for strat, kw in [("jump_mala", dict(inner_param_kwargs=dict(n_iterations=20))), ("jump_hmc", dict(inner_param_kwargs=dict(n_iterations=2), inner_kernel_kwargs=dict(n_leapfrog_steps=10, step_size=0.05))), ("imh", {})]:
        s = nfmc_b200.create_sampler(tgt, event_shape=(d,), strategy=strat, param_kwargs=dict(n_iterations=3, store_samples=False), **kw)
        x0 = torch.randn(n, d, device="cuda") * 0.7
        s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        steps = out.statistics.expectations.n_seen
        print(f"{name:9s} {strat:10s} {steps / dt:.3e} chain-steps/s  acc {out.statistics.acceptance_rate:.3f}  mean|x| {float(out.mean.abs().max()):.3f} var {float(out.variance.mean()):.3f}  peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
