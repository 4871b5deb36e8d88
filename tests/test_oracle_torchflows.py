"""Pin for SURVEY.md section 8 row a11: the moment the real ``torchflows`` package (github davidnabergoj/torchflows; an
unpinned, un-vendored dependency of the reference: /root/reference/setup.py:51-56) becomes importable, these tests compare
``oracle/realnvp_ref.py`` -- the specification every CUDA flow kernel is tested against -- with it.  Here, and on the GPU
box, the package is absent and the whole module is skipped: the flow arithmetic stays "parity unpinned" (oracle/__init__.py).

Assumptions the comparison checks: ``Flow(RealNVP(event_shape, n_layers=..))`` has
``len(bijection.layers) == 3 * n_layers + 3``; parameters appear in the same order with the same shapes, so they can be
copied positionally; ``bijection.forward / inverse`` return ``(y, log_det)``; ``log_prob`` and ``sample(return_log_prob=True)``
agree; ActNorm parameters are trainable ``nn.Parameter``s that are data-initialised on the first training pass."""
import pytest
import torch

torchflows = pytest.importorskip("torchflows")


def _pair(d, n_layers, ck):
    from torchflows.flows import Flow
    from torchflows.architectures import RealNVP
    from oracle.realnvp_ref import FlowRef, RealNVPRef
    torch.manual_seed(0)
    kw = {} if ck is None else dict(conditioner_kwargs=ck)
    real = Flow(RealNVP((d,), n_layers=n_layers, **kw))
    ref = FlowRef(RealNVPRef((d,), n_layers=n_layers, conditioner_kwargs=ck))
    pr, po = list(real.parameters()), list(ref.parameters())
    assert [tuple(p.shape) for p in pr] == [tuple(p.shape) for p in po], "parameter layout differs from torchflows"
    with torch.no_grad():
        for a, b in zip(pr, po):
            a.add_(0.1 * torch.randn_like(a))
            b.copy_(a)
    return real.eval(), ref.eval()


@pytest.mark.parametrize("d,n_layers,ck", [(6, 2, None), (25, 2, None), (100, 2, None), (7, 3, dict(n_layers=3, n_hidden=10)),
                                           (10, 10, dict(n_layers=5, n_hidden=100))])
def test_realnvp_restatement_matches_torchflows(d, n_layers, ck):
    real, ref = _pair(d, n_layers, ck)
    assert len(real.bijection.layers) == len(ref.bijection.layers) == 3 * n_layers + 3     # test/test_flow_kwargs.py:18,28,30
    x = torch.randn(64, d)
    with torch.no_grad():
        z0, l0 = real.bijection.forward(x)
        z1, l1 = ref.bijection.forward(x)
        torch.testing.assert_close(z1, z0, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(l1, l0, rtol=1e-5, atol=1e-5)
        x0, m0 = real.bijection.inverse(x)
        x1, m1 = ref.bijection.inverse(x)
        torch.testing.assert_close(x1, x0, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(m1, m0, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(ref.log_prob(x), real.log_prob(x), rtol=1e-5, atol=1e-5)


def test_actnorm_is_trainable_and_data_initialised_like_torchflows():
    real, ref = _pair(8, 2, None)
    names_real = [n for n, p in real.named_parameters() if p.requires_grad]
    names_ref = [n for n, p in ref.named_parameters() if p.requires_grad]
    assert len(names_real) == len(names_ref), "trainable parameter sets differ (is ActNorm frozen in torchflows?)"
    x = 3.0 + 2.0 * torch.randn(512, 8)
    real.train(); ref.train()
    with torch.no_grad():
        a = real.log_prob(x)
        b = ref.log_prob(x)
    torch.testing.assert_close(b, a, rtol=1e-4, atol=1e-4)
