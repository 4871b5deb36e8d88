"""Flow passes for conditioner shapes outside the tensor-core path (deep: n_layers = 5, or odd d): the generic CUDA-core path of
flow.cuh against the row-tile fp32 pass of train_wide.cu (nfmc_flow_wide_pass)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nfmc_b200.flow import create_flow_object
from nfmc_b200.flow_train import WideTrainer


def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


for name, d, spec, n in [("deep M=5 H=100 Lc=10 d=100", 100, 'realnvp%{"n_layers": 10, "conditioner_kwargs": {"n_layers": 5, "n_hidden": 100}}', 1 << 16),
                         ("deep M=3 H=64 Lc=2 d=100", 100, 'realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 3, "n_hidden": 64}}', 1 << 18),
                         ("wide M=2 H=64 Lc=2 d=101 (odd)", 101, 'realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}', 1 << 18),
                         ("wide M=2 H=64 Lc=2 d=1000", 1000, 'realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 64}}', 1 << 16)]:
    torch.manual_seed(0)
    f = create_flow_object(spec, (d,)).to("cuda")
    with torch.no_grad():
        for p in f.parameters():
            p.add_(0.02 * torch.randn_like(p))
    x = torch.randn(n, d, device="cuda")
    assert not f.bijection.uses_tensor_cores()
    tr = WideTrainer(f, torch.device("cuda"), 1e-3)
    assert f.bijection.uses_row_tile_pass()
    zs, ls = f.bijection.forward(x)                      # sampling path: row-tile kernel, weights packed transposed, 2 CTAs/SM
    zw, lw = tr.run_pass(x, False)                       # trainer's pass: same kernel on the module-order vector
    rec = {"shape": name, "rows": n, "row_tile_transposed_forward_ms": t(lambda: f.bijection.forward(x)),
           "row_tile_transposed_inverse_ms": t(lambda: f.bijection.inverse(x)),
           "row_tile_module_order_forward_ms": t(lambda: tr.run_pass(x, False)), "row_tile_module_order_inverse_ms": t(lambda: tr.run_pass(x, True))}
    os.environ["NFMC_B200_NO_ROW_TILE"] = "1"            # per-chain generic conditioner of flow.cuh
    zg, lg = f.bijection.forward(x)
    rec.update(generic_forward_ms=t(lambda: f.bijection.forward(x)), generic_inverse_ms=t(lambda: f.bijection.inverse(x)),
               max_abs_diff_z_vs_generic=float((zg - zs).abs().max()), max_abs_diff_logdet_vs_generic=float((lg - ls).abs().max()),
               max_abs_diff_z_vs_module_order=float((zw - zs).abs().max()))
    del os.environ["NFMC_B200_NO_ROW_TILE"]
    print(json.dumps(rec), flush=True)
