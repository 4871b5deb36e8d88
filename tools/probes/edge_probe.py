import torch, nfmc_b200, traceback
from nfmc_b200.potentials import StandardGaussian
for strat in ["mala", "hmc", "jump_mala", "imh", "neutra_hmc"]:
    for n, d in [(0, 4), (1, 4), (1, 2), (1, 1), (3, 1)]:
        try:
            out = nfmc_b200.sample(StandardGaussian((d,)), event_shape=(d,), strategy=strat, n_chains=n, n_iterations=3, device=torch.device("cuda"),
                                   **({"inner_param_kwargs": {"n_iterations": 2}} if strat.startswith("jump") else {}))
            print(strat, n, d, "ok", tuple(out.samples.shape), out.statistics.acceptance_rate, bool(torch.isfinite(out.samples).all()))
        except Exception as e:
            print(strat, n, d, "ERR", type(e).__name__, str(e)[:120])
