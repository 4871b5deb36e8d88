"""HMC alone vs jump_hmc at config C2's shape (ill-conditioned Gaussian, d = 100): where a jump_hmc iteration spends its time."""
import json, sys, time, torch, nfmc_b200
from nfmc_b200.potentials import make_potential
d = 100
if len(sys.argv) > 1:       # HMC alone on the named potentials, 2^20 chains
    for pot in sys.argv[1:]:
        s = nfmc_b200.create_sampler(make_potential(pot, (d,)), event_shape=(d,), strategy="hmc", param_kwargs=dict(n_iterations=30, store_samples=False))
        x0 = torch.randn(1 << 20, d, device="cuda") * 0.1
        s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"strategy": "hmc", "potential": pot, "chains": 1 << 20, "ms_per_step": 1e3 * dt / 30, "chain_steps_per_s": out.statistics.expectations.n_seen / dt}))
    sys.exit(0)
for n in (65536, 1 << 20):
    for strat, kw, T in [("hmc", {}, 50), ("jump_hmc", {"inner_param_kwargs": {"n_iterations": 5}}, 20), ("jump_hmc", {"inner_param_kwargs": {"n_iterations": 50}}, 4)]:
        s = nfmc_b200.create_sampler(make_potential("g1", (d,)), event_shape=(d,), strategy=strat, param_kwargs=dict(n_iterations=T, store_samples=False), **kw)
        x0 = torch.randn(n, d, device="cuda") * 0.1
        s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = s.sample(x0, show_progress=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        st = out.statistics
        print(json.dumps({"strategy": strat, "inner": kw.get("inner_param_kwargs", {}).get("n_iterations"), "chains": n, "iterations": T, "seconds": dt,
                          "chain_steps_per_s": st.expectations.n_seen / dt, "acc": st.acceptance_rate}))
