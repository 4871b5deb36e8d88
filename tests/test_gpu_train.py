"""B200: native flow training (csrc/train_kernels.cu, csrc/train_api.cu) against torch autograd through the ORACLE RealNVP
(oracle/realnvp_ref.py: FlowRef.log_prob / bijection.inverse, carrying the same parameters) and torch.optim.AdamW.

Reference call sites of the training these kernels replace: nfmc/jump.py:139-151,201, nfmc/imh.py:67-72,171-175,
nfmc/neutra.py:84-91.  Tolerances: gradients agree to 2e-4 of the largest gradient entry (fp32 sums over the batch in a
different order, MUFU exp/log/tanh in the kernel).
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(6, 2, 4), (7, 3, 6), (25, 2, None), (100, 2, None), (9, 1, 8), (64, 4, 8), (1000, 1, None), (33, 0, 4),
          (1023, 2, None), (2, 2, 4), (3, 3, 5)]


def _flow(d, Lc, H, seed=0, scale=0.15):
    from nfmc_b200.flow import Flow, RealNVP
    torch.manual_seed(seed)
    ck = None if H is None else dict(n_layers=2, n_hidden=H)
    f = Flow(RealNVP((d,), n_layers=Lc, conditioner_kwargs=ck, conditioner_dtype="fp32"))
    with torch.no_grad():
        for p in f.parameters():
            p.add_(scale * torch.randn_like(p))
        for l in f.bijection.layers:
            if hasattr(l, "initialised"):
                l.initialised.fill_(True)
    return f.to("cuda").eval()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


@pytest.mark.parametrize("d,Lc,H", SHAPES)
def test_device_pack_matches_host_pack(d, Lc, H):
    from nfmc_b200.flow import pack_realnvp
    from nfmc_b200.flow_train import NativeTrainer
    f = _flow(d, Lc, H)
    tr = NativeTrainer(f, torch.device("cuda"), 0.05)
    tr.pack()
    ref = pack_realnvp(f.bijection)
    got = tr.blob.cpu()
    assert got.shape == ref.shape
    np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("d,Lc,H", SHAPES)
def test_nll_gradient_matches_autograd(d, Lc, H):
    from nfmc_b200 import _native as N
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import NativeTrainer
    f = _flow(d, Lc, H, seed=d)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    n = 137
    x = torch.randn(n, d, device=dev) * 1.3 + 0.2
    rows = torch.randperm(n, device=dev)[:101]
    params = list(oflow.bijection.parameters())
    with torch.enable_grad():
        loss = -oflow.log_prob(x[rows]).sum()
        grads = torch.autograd.grad(loss, params)
    ref = torch.cat([g.reshape(-1) for g in grads])
    tr = NativeTrainer(f, dev, 0.05)
    tr.pack()
    desc = tr.desc()
    N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), rows.data_ptr(), rows.numel(), N.ptr(tr.gblob),
                                       N.ptr(tr.loss), 0, tr.stream))
    tr.unpack(1.0)
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    assert _rel(tr.gtheta, ref) < 2e-4, _rel(tr.gtheta, ref)
    # per-tensor check so that small gradients (biases, act-norms) are not hidden behind the weights
    off = 0
    for p, g in zip(params, grads):
        k = p.numel()
        assert _rel(tr.gtheta[off:off + k], g.reshape(-1)) < 1e-3, (tuple(p.shape), off)
        off += k


@pytest.mark.parametrize("d,Lc,H,pot", [(6, 2, 4, "fn"), (7, 3, 6, "gm"), (100, 2, None, "g1"), (26, 2, 8, "rb"),
                                        (25, 1, None, "g0"), (1000, 1, None, "gm")])
def test_reverse_kl_gradient_matches_autograd(d, Lc, H, pot):
    from nfmc_b200 import _native as N
    from nfmc_b200 import potentials as P
    from gpu_util import oracle_flow_from_product
    from oracle.potentials_ref import make_potential_ref
    from nfmc_b200.flow_train import NativeTrainer
    f = _flow(d, Lc, H, seed=3 * d, scale=0.1)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    n = 77
    z = torch.randn(n, d, device=dev)
    potential = P.make_potential(pot, (d,))
    upot = make_potential_ref(pot, (d,))
    tlp = lambda x_: -upot(x_.cpu()).to(x_.device)      # the oracle potentials keep their parameters on the host
    params = list(oflow.bijection.parameters())
    with torch.enable_grad():
        x, ld = oflow.bijection.inverse(z)
        log_q = (-0.5 * z.square()).sum(dim=1) - 0.5 * d * math.log(2 * math.pi) - ld
        loss = (log_q - tlp(x)).sum()
        grads = torch.autograd.grad(loss, params)
    ref = torch.cat([g.reshape(-1) for g in grads])
    tr = NativeTrainer(f, dev, 0.05)
    tr.pack()
    desc = tr.desc()
    pd, keep = potential.descriptor(dev)
    rng = N.rng_desc(0, 0, z, None)
    N.check(N.lib().nfmc_flow_kl_grad(C.byref(pd), C.byref(desc), C.byref(rng), 0, n, N.ptr(tr.gblob), N.ptr(tr.loss), 0,
                                      tr.stream))
    tr.unpack(1.0)
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 2e-4 * max(1.0, abs(float(loss.detach())))
    assert _rel(tr.gtheta, ref) < 5e-4, _rel(tr.gtheta, ref)


def test_nll_gradient_large_batch_shared_accumulation():
    """Batches of many tiles per CTA accumulate in shared memory and flush once (train_api.cu: shared_grad)."""
    from nfmc_b200 import _native as N
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import NativeTrainer
    d, n = 100, 70001
    f = _flow(d, 2, None, seed=5)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    x = torch.randn(n, d, device=dev)
    params = list(oflow.bijection.parameters())
    with torch.enable_grad():
        loss = -oflow.log_prob(x).sum()
        grads = torch.autograd.grad(loss, params)
    ref = torch.cat([g.reshape(-1) for g in grads])
    tr = NativeTrainer(f, dev, 0.05)
    tr.pack()
    desc = tr.desc()
    N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), None, n, N.ptr(tr.gblob), N.ptr(tr.loss), 0, tr.stream))
    tr.unpack(1.0)
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    assert _rel(tr.gtheta, ref) < 5e-4, _rel(tr.gtheta, ref)


def test_adamw_matches_torch():
    from nfmc_b200 import _native as N
    dev = torch.device("cuda")
    torch.manual_seed(1)
    n = 5000
    theta = torch.randn(n, device=dev)
    p = torch.nn.Parameter(theta.clone())
    opt = torch.optim.AdamW([p], lr=0.05)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 6):
        g = torch.randn(n, device=dev) * (0.1 * step)
        p.grad = g.clone()
        opt.step()
        N.check(N.lib().nfmc_adamw_step(N.ptr(theta), N.ptr(g), N.ptr(m), N.ptr(v), n, 0.05, 0.9, 0.999, 1e-8, 0.01, step,
                                        N.stream_ptr(dev)))
    np.testing.assert_allclose(theta.cpu().numpy(), p.detach().cpu().numpy(), rtol=2e-5, atol=2e-6)


def _gaussian_data(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    mu = torch.linspace(-1.0, 1.0, d)
    sd = torch.linspace(0.5, 2.0, d)
    return mu + sd * torch.randn(n, d, generator=g)


def test_native_fit_tracks_an_autograd_adamw_loop():
    """Same data, same minibatches: a few epochs of the native loop land where torch autograd + torch.optim.AdamW over the
    ORACLE flow land (the autograd loop is test code; the product has none)."""
    from nfmc_b200.flow import Flow, RealNVP
    from gpu_util import oracle_flow_from_product
    d = 20
    x = _gaussian_data(3000, d, 0).cuda()
    xv = _gaussian_data(1000, d, 1).cuda()
    torch.manual_seed(11)
    f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
    for l in f.bijection.layers:
        if hasattr(l, "initialised"):
            l.initialised.fill_(True)               # identical starting point for both loops
    oflow = oracle_flow_from_product(f)
    before = float(-f.log_prob(xv).mean())
    f.fit(x, x_val=xv, n_epochs=6, lr=0.02, batch_size=500, shuffle=False, keep_best_weights=False)
    after = float(-f.log_prob(xv).mean())
    assert after < before - 1.0
    params = list(oflow.bijection.parameters())
    opt = torch.optim.AdamW(params, lr=0.02)
    oflow.train()
    for _ in range(6):
        for i in range(0, 3000, 500):
            opt.zero_grad(set_to_none=True)
            with torch.enable_grad():
                loss = -oflow.log_prob(x[i:i + 500]).mean()
            loss.backward()
            opt.step()
    oflow.eval()
    with torch.no_grad():
        ref = float(-oflow.log_prob(xv).mean())
    assert abs(after - ref) < 0.02 * abs(ref) + 0.05, (after, ref)


def test_native_fit_early_stopping_and_rollback():
    from nfmc_b200.flow import Flow, RealNVP
    d = 12
    x = _gaussian_data(2000, d, 2).cuda()
    torch.manual_seed(0)
    f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
    f.fit(x, n_epochs=400, lr=0.05, batch_size="adaptive", early_stopping=True, early_stopping_threshold=5)
    nll = float(-f.log_prob(x).mean())
    # entropy of the generating Gaussian: the fitted flow must be close to it
    sd = torch.linspace(0.5, 2.0, d)
    h = float(0.5 * d * (1 + math.log(2 * math.pi)) + sd.log().sum())
    assert nll < h + 0.6, (nll, h)
    with pytest.raises(ValueError):
        bad = x.clone()
        bad[0, 0] = float("nan")
        Flow(RealNVP((d,), n_layers=2)).to("cuda").fit(bad, n_epochs=3, lr=0.05, batch_size=500)


def test_native_variational_fit_reduces_reverse_kl():
    from nfmc_b200 import potentials as P
    from nfmc_b200.flow import Flow, RealNVP
    d = 10
    potential = P.make_potential("g1", (d,))            # sigma_i from 0.1 to 100: far from the standard-normal start
    torch.manual_seed(4)
    f = Flow(RealNVP((d,), n_layers=2)).to("cuda")

    def reverse_kl():
        xs, lq = f.sample(4096, return_log_prob=True, seed=7)
        return float((lq + potential(xs)).mean())

    before = reverse_kl()
    f.variational_fit(potential.log_prob_fn(), n_epochs=300, lr=0.05, n_samples=64)
    after = reverse_kl()
    assert after < before - 5.0, (before, after)


def test_training_kernels_edge_cases():
    """Single row, repeated row indices, minimal event size, accumulate = 1, Philox reproducibility of the reverse-KL draw."""
    from nfmc_b200 import _native as N
    from nfmc_b200 import potentials as P
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import NativeTrainer
    dev = torch.device("cuda")
    for d, Lc, H in [(2, 1, 4), (3, 2, 5)]:
        f = _flow(d, Lc, H, seed=d)
        oflow = oracle_flow_from_product(f)
        x = torch.randn(5, d, device=dev)
        params = list(oflow.bijection.parameters())
        tr = NativeTrainer(f, dev, 0.05)
        tr.pack()
        desc = tr.desc()
        # one row
        rows = torch.tensor([3], device=dev)
        with torch.enable_grad():
            ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad(-oflow.log_prob(x[rows]).sum(), params)])
        N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), rows.data_ptr(), 1, N.ptr(tr.gblob), N.ptr(tr.loss), 0, tr.stream))
        tr.unpack(1.0)
        assert _rel(tr.gtheta, ref) < 5e-4
        # repeated indices, in two accumulating launches == one launch over all of them
        rows = torch.tensor([0, 0, 4, 1, 4, 4, 2], device=dev)
        with torch.enable_grad():
            loss = -oflow.log_prob(x[rows]).sum()
            ref = torch.cat([g.reshape(-1) for g in torch.autograd.grad(loss, params)])
        N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), rows.data_ptr(), 3, N.ptr(tr.gblob), N.ptr(tr.loss), 0, tr.stream))
        N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), rows[3:].contiguous().data_ptr(), 4, N.ptr(tr.gblob),
                                           N.ptr(tr.loss), 1, tr.stream))
        tr.unpack(1.0)
        assert _rel(tr.gtheta, ref) < 5e-4
        assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    # reverse KL with the kernel's own Philox draw: same (seed, step) -> same loss and gradient, next step -> different
    d = 10
    f = _flow(d, 2, None, seed=1)
    tr = NativeTrainer(f, dev, 0.05)
    tr.pack()
    desc = tr.desc()
    pd, keep = P.make_potential("fn", (d,)).descriptor(dev)
    out = []
    for step0 in (5, 5, 6):
        rng = N.rng_desc(1234, step0, None, None)
        N.check(N.lib().nfmc_flow_kl_grad(C.byref(pd), C.byref(desc), C.byref(rng), 0, 64, N.ptr(tr.gblob), N.ptr(tr.loss), 0, tr.stream))
        out.append((float(tr.loss[0]), tr.gblob.clone()))
    assert out[0][0] == out[1][0] and torch.allclose(out[0][1], out[1][1], rtol=1e-5, atol=1e-6)
    assert out[0][0] != out[2][0]
    # errors are reported
    bad = N.RealNVPDesc(d, 2, 3, 16, tr.blob.data_ptr(), tr.blob.numel())
    assert N.lib().nfmc_flow_nll_grad(C.byref(bad), N.ptr(torch.zeros(4, d, device=dev)), None, 4, N.ptr(tr.gblob), None, 0, tr.stream) != 0
    assert N.lib().nfmc_flow_param_count(d, 2, 3, 16) == -1
