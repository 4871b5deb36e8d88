"""B200: the tcgen05 / TMEM conditioner path (csrc/cond_tc.cu) against the fp32 CPU oracle.
Tolerance: the bf16-conditioner bound of the north star, rtol 1e-2."""
import pytest
import torch

from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu


def _err(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float(((a - b).abs() / (1.0 + b.abs())).max())


@pytest.mark.timeout(180)
@pytest.mark.parametrize("d,n_layers,H,n", [(100, 2, 64, 300), (100, 4, 256, 1000), (64, 3, 16, 129), (128, 2, 128, 77),
                                           (16, 1, 32, 5), (100, 2, 256, 20000)])
def test_tc_forward_inverse_logprob(d, n_layers, H, n):
    from gpu_util import product_flow_from_oracle
    torch.manual_seed(d + H)
    oflow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=0.05, seed=H)
    flow = product_flow_from_oracle(oflow, conditioner_dtype="bf16")
    assert flow.bijection.uses_tensor_cores()
    x = torch.randn(n, d)
    with torch.no_grad():
        z_ref, ld_ref = oflow.bijection.forward(x)
        xi_ref, ldi_ref = oflow.bijection.inverse(x)
        lp_ref = oflow.log_prob(x)
    z, ld = flow.bijection.forward(x.cuda())
    torch.cuda.synchronize()
    assert _err(z, z_ref) < 1e-2, _err(z, z_ref)
    assert float((ld.cpu() - ld_ref).abs().max()) < 1e-2 * (1 + float(ld_ref.abs().max())), float((ld.cpu() - ld_ref).abs().max())
    xi, ldi = flow.bijection.inverse(x.cuda())
    assert _err(xi, xi_ref) < 1e-2, _err(xi, xi_ref)
    assert float((ldi.cpu() - ldi_ref).abs().max()) < 1e-2 * (1 + float(ldi_ref.abs().max()))
    lp = flow.log_prob(x.cuda())
    assert float((lp.cpu() - lp_ref).abs().max()) < 1e-2 * (1 + float(lp_ref.abs().max()))
    # the fp32 CUDA-core path on the same flow agrees with the tensor-core path to the same tolerance
    flow32 = product_flow_from_oracle(oflow, conditioner_dtype="fp32")
    z32, ld32 = flow32.bijection.forward(x.cuda())
    assert _err(z, z32) < 1e-2


def test_tc_rejects_unsupported_shapes():
    from nfmc_b200.flow import Flow, RealNVP
    with pytest.raises(ValueError):
        RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=20), conditioner_dtype="bf16")
    with pytest.raises(ValueError):
        RealNVP((101,), conditioner_kwargs=dict(n_layers=2, n_hidden=64), conditioner_dtype="bf16")
    assert not RealNVP((100,)).uses_tensor_cores()           # default H = 5: CUDA-core path
    assert RealNVP((100,), conditioner_kwargs=dict(n_layers=2, n_hidden=64)).uses_tensor_cores()
