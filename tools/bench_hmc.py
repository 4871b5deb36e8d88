"""HMC alone vs jump_hmc at config C2's shape (ill-conditioned Gaussian, d = 100): where a jump_hmc iteration spends its time."""
import json, sys, time, torch, nfmc_b200
from nfmc_b200.potentials import make_potential
d = 100
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
