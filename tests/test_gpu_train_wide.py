"""B200: native training of wide / deep conditioners (csrc/train_wide.cu) against torch autograd through the ORACLE RealNVP
(oracle/realnvp_ref.py) -- gradients of both objectives, the fp32 pass, input cotangents, and that `fit` /
`variational_fit` / adaptive IMH with such a flow never enter torch autograd.

Reference call sites of the training these kernels replace: nfmc/jump.py:139-151,201, nfmc/imh.py:67-72,171-175,
nfmc/neutra.py:84-91; the reference's own deep shape is test/test_flow_kwargs.py:49 (n_layers=10, conditioner n_layers=5,
n_hidden=100).  Tolerance: 2e-4 of the largest gradient entry overall, 2e-3 per parameter tensor (fp32 sums in a
different order).
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

# (d, Lc, M, H)
SHAPES = [(6, 2, 2, 16), (7, 3, 3, 10), (100, 2, 2, 64), (100, 4, 2, 256), (25, 2, 5, 100), (9, 1, 1, 4), (33, 0, 2, 12),
          (101, 2, 2, 32), (1000, 1, 2, 16), (10, 10, 5, 100), (2, 2, 2, 9), (3, 3, 4, 7), (100, 2, 2, 5)]


def _flow(d, Lc, M, H, seed=0, scale=0.12):
    from nfmc_b200.flow import Flow, RealNVP
    torch.manual_seed(seed)
    f = Flow(RealNVP((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=M, n_hidden=H), conditioner_dtype="fp32"))
    with torch.no_grad():
        for p in f.parameters():
            p.add_(scale * torch.randn_like(p) / max(1.0, math.sqrt(p.shape[-1] / 8.0)))
        for l in f.bijection.layers:
            if hasattr(l, "initialised"):
                l.initialised.fill_(True)
    return f.to("cuda").eval()


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _check_grads(got, params, grads, tol_all=2e-4, tol_each=2e-3):
    ref = torch.cat([g.reshape(-1) for g in grads])
    assert _rel(got, ref) < tol_all, _rel(got, ref)
    off = 0
    scale = float(ref.abs().max())
    for p, g in zip(params, grads):
        k = p.numel()
        err = float((got[off:off + k].double() - g.reshape(-1).double()).abs().max())
        # small tensors (biases, act-norms) are checked against their own scale, floored at 1e-3 of the global one
        assert err < tol_each * max(float(g.abs().max()), 1e-3 * scale), (tuple(p.shape), off, err)
        off += k


@pytest.mark.parametrize("d,Lc,M,H", SHAPES)
def test_wide_nll_gradient_matches_autograd(d, Lc, M, H, monkeypatch):
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import WideTrainer
    f = _flow(d, Lc, M, H, seed=d + M)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    n = 137
    x = torch.randn(n, d, device=dev) * 1.3 + 0.2
    rows = torch.randperm(n, device=dev)[:101]
    params = list(oflow.bijection.parameters())
    xr = x[rows].clone().requires_grad_(True)
    with torch.enable_grad():
        loss = -oflow.log_prob(xr).sum()
        grads = torch.autograd.grad(loss, params + [xr])
    for R in ("8", "16", "32"):
        monkeypatch.setenv("NFMC_WIDE_ROWS", R)
        tr = WideTrainer(f, dev, 0.05)
        gx = torch.full((rows.numel(), d), float("nan"), device=dev)
        tr.nll_grad(x, rows, rows.numel(), grad_x=gx)
        assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
        _check_grads(tr.gtheta, params, grads[:-1])
        assert _rel(gx, grads[-1]) < 5e-4, _rel(gx, grads[-1])


@pytest.mark.parametrize("d,Lc,M,H", [(6, 2, 2, 16), (7, 3, 3, 10), (100, 4, 2, 256), (25, 2, 5, 40), (9, 1, 1, 4), (101, 3, 2, 32)])
def test_wide_pass_matches_oracle(d, Lc, M, H):
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import WideTrainer
    f = _flow(d, Lc, M, H, seed=2 * d)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    tr = WideTrainer(f, dev, 0.05)
    x = torch.randn(301, d, device=dev)
    with torch.no_grad():
        z_ref, ld_ref = oflow.bijection.forward(x)
        xi_ref, ldi_ref = oflow.bijection.inverse(x)
    z, ld = tr.run_pass(x, inverse=False)
    xi, ldi = tr.run_pass(x, inverse=True)
    torch.testing.assert_close(z, z_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ld, ld_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(xi, xi_ref, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(ldi, ldi_ref, rtol=1e-4, atol=1e-4)
    back, ldb = tr.run_pass(z, inverse=True)
    torch.testing.assert_close(back, x, rtol=1e-4, atol=2e-4)
    torch.testing.assert_close(ldb, -ld, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("d,Lc,M,H,pot", [(6, 2, 2, 16, "fn"), (7, 3, 3, 12, "gm"), (100, 4, 2, 256, "g1"), (26, 2, 2, 64, "rb"),
                                          (25, 1, 4, 20, "g0"), (100, 2, 2, 5, "fn")])
def test_wide_reverse_kl_gradient_matches_autograd(d, Lc, M, H, pot):
    from nfmc_b200 import potentials as P
    from gpu_util import oracle_flow_from_product
    from oracle.potentials_ref import make_potential_ref
    from nfmc_b200.flow_train import WideTrainer
    f = _flow(d, Lc, M, H, seed=3 * d, scale=0.08)
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
    tr = WideTrainer(f, dev, 0.05)
    theta0 = tr.theta.clone()
    tr.lr = 0.0                                          # keep theta; the step still leaves the gradient in tr.gtheta
    tr.kl_step(potential, n, 0, 0, z=z)
    assert torch.equal(tr.theta, theta0 * (1.0 - 0.0))
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 2e-4 * max(1.0, abs(float(loss.detach())))
    _check_grads(tr.gtheta, params, grads, tol_all=5e-4, tol_each=5e-3)


def test_wide_large_batch_and_accumulate():
    """n = 70001 rows (32-row tiles, many tiles per CTA, ragged tail) and accumulate = 1 over two launches."""
    import ctypes as C
    from nfmc_b200 import _native as N
    from gpu_util import oracle_flow_from_product
    from nfmc_b200.flow_train import WideTrainer
    d, n = 100, 70001
    f = _flow(d, 2, 2, 64, seed=5)
    oflow = oracle_flow_from_product(f)
    dev = torch.device("cuda")
    x = torch.randn(n, d, device=dev)
    params = list(oflow.bijection.parameters())
    with torch.enable_grad():
        loss = -oflow.log_prob(x).sum()
        grads = torch.autograd.grad(loss, params)
    tr = WideTrainer(f, dev, 0.05)
    tr.nll_grad(x, None, n)
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    _check_grads(tr.gtheta, params, grads, tol_all=5e-4, tol_each=5e-3)
    one = tr.gtheta.clone()
    k = 30000
    sh = tr._shape()
    N.check(N.lib().nfmc_flow_wide_nll_grad(*sh, N.ptr(tr.theta), N.ptr(x), None, k, N.ptr(tr.gtheta), N.ptr(tr.loss), None, 0, tr.stream))
    N.check(N.lib().nfmc_flow_wide_nll_grad(*sh, N.ptr(tr.theta), N.ptr(x[k:]), None, n - k, N.ptr(tr.gtheta), N.ptr(tr.loss), None, 1,
                                            tr.stream))
    assert _rel(tr.gtheta, one) < 1e-3
    assert abs(float(tr.loss[0]) - float(loss.detach())) <= 1e-4 * abs(float(loss.detach()))
    # errors are reported
    assert N.lib().nfmc_flow_wide_param_count(d, 2, 0, 16) == -1
    assert N.lib().nfmc_flow_wide_nll_grad(d, 2, 2, 64, N.ptr(tr.theta), None, None, 4, N.ptr(tr.gtheta), None, None, 0, tr.stream) != 0


def _gaussian_data(n, d, seed):
    g = torch.Generator().manual_seed(seed)
    mu = torch.linspace(-1.0, 1.0, d)
    sd = torch.linspace(0.5, 2.0, d)
    return mu + sd * torch.randn(n, d, generator=g)


class _NoAutograd:
    """Context: any call into torch autograd raises (the product must train through its own kernels)."""

    def __enter__(self):
        self.saved = (torch.Tensor.backward, torch.autograd.grad, torch.autograd.backward)

        def boom(*a, **k):
            raise AssertionError("torch autograd was entered")
        torch.Tensor.backward = boom
        torch.autograd.grad = boom
        torch.autograd.backward = boom
        return self

    def __exit__(self, *exc):
        torch.Tensor.backward, torch.autograd.grad, torch.autograd.backward = self.saved
        return False


@pytest.mark.parametrize("ck", [dict(n_layers=2, n_hidden=64), dict(n_layers=3, n_hidden=24)])
def test_wide_fit_matches_an_autograd_adamw_loop(ck):
    """Same data, same minibatches: a few epochs of the native wide loop land where torch autograd + torch.optim.AdamW land
    (the loop below is TEST code over the oracle flow; the product has no such loop)."""
    from nfmc_b200.flow import Flow, RealNVP
    from gpu_util import oracle_flow_from_product
    d = 20
    x = _gaussian_data(3000, d, 0).cuda()
    xv = _gaussian_data(1000, d, 1).cuda()
    torch.manual_seed(11)
    f = Flow(RealNVP((d,), n_layers=2, conditioner_kwargs=ck, conditioner_dtype="fp32")).to("cuda")
    for l in f.bijection.layers:
        if hasattr(l, "initialised"):
            l.initialised.fill_(True)               # identical starting point for both loops
    oflow = oracle_flow_from_product(f)
    before = float(-f.log_prob(xv).mean())
    with _NoAutograd():
        f.fit(x, x_val=xv, n_epochs=6, lr=0.02, batch_size=500, shuffle=False, keep_best_weights=False)
    after = float(-f.log_prob(xv).mean())
    assert after < before - 1.0, (before, after)
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
    got = torch.cat([p.detach().reshape(-1) for p in f.bijection.parameters()])
    want = torch.cat([p.detach().reshape(-1) for p in params])
    assert float((got - want).abs().max()) < 0.05, float((got - want).abs().max())


def test_wide_variational_fit_reduces_reverse_kl_without_autograd():
    from nfmc_b200 import potentials as P
    from nfmc_b200.flow import Flow, RealNVP
    d = 10
    potential = P.make_potential("g1", (d,))
    torch.manual_seed(4)
    f = Flow(RealNVP((d,), n_layers=2, conditioner_kwargs=dict(n_layers=2, n_hidden=32), conditioner_dtype="fp32")).to("cuda")

    def reverse_kl():
        xs, lq = f.sample(4096, return_log_prob=True, seed=7)
        return float((lq + potential(xs)).mean())

    before = reverse_kl()
    with _NoAutograd():
        f.variational_fit(potential.log_prob_fn(), n_epochs=300, lr=0.05, n_samples=64)
    after = reverse_kl()
    assert after < before - 5.0, (before, after)


def test_adaptive_imh_and_fit_nf_with_a_wide_flow_never_enter_autograd():
    """adaptive_imh (imh.py:152-175: one full-batch epoch per iteration) and jump_mala with fit_nf=True (jump.py:193-201) on
    a tensor-core-eligible wide flow: refits run on csrc/train_wide.cu, sampling on the tcgen05 kernels."""
    import nfmc_b200
    from nfmc_b200 import potentials as P
    d = 16
    target = P.make_potential("g0", (d,))
    flow = 'realnvp%{"n_layers": 2, "conditioner_kwargs": {"n_layers": 2, "n_hidden": 32}}'
    torch.manual_seed(0)
    with _NoAutograd():
        out = nfmc_b200.sample(target, event_shape=(d,), strategy="adaptive_imh", flow=flow, n_chains=512, n_iterations=100,
                               device="cuda", show_progress=False)
        assert out.samples.shape == (100, 512, d) and bool(torch.isfinite(out.samples).all())
        out = nfmc_b200.sample(target, event_shape=(d,), strategy="jump_mala", flow=flow, n_chains=256, n_iterations=13,
                               device="cuda", show_progress=False,
                               param_kwargs=dict(fit_nf=True, n_jumps_before_training=2,
                                                 flow_fit_kwargs=dict(n_epochs=3, batch_size="adaptive")),
                               inner_param_kwargs=dict(n_iterations=5))
        assert bool(torch.isfinite(out.samples).all())
        assert out.statistics.n_attempted_jumps == 13 * 256


def test_default_flow_through_the_wide_kernel_equals_the_register_kernel(monkeypatch):
    """The two native training paths on the same default flow and data end at the same place (AdamW turns fp32 rounding
    differences of near-zero gradients into O(lr) weight differences, hence the loose weight tolerance)."""
    from nfmc_b200.flow import Flow, RealNVP
    d = 20
    x = _gaussian_data(3000, d, 0).cuda()
    xv = _gaussian_data(1000, d, 1).cuda()
    res, score = {}, {}
    for mode in ("register", "wide"):
        if mode == "wide":
            monkeypatch.setenv("NFMC_B200_WIDE_TRAINING", "1")
        torch.manual_seed(11)
        f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
        f.fit(x, n_epochs=4, lr=0.02, batch_size=500, shuffle=False, keep_best_weights=False)
        res[mode] = torch.cat([p.detach().reshape(-1) for p in f.bijection.parameters()])
        score[mode] = float(-f.log_prob(xv).mean())
    assert abs(score["register"] - score["wide"]) < 0.01 * abs(score["register"]), score
    assert float((res["register"] - res["wide"]).abs().max()) < 0.1, float((res["register"] - res["wide"]).abs().max())
    # one step from the same start: the gradients themselves agree tightly
    from nfmc_b200.flow_train import NativeTrainer, WideTrainer
    import ctypes as C
    from nfmc_b200 import _native as N
    torch.manual_seed(3)
    f = Flow(RealNVP((d,), n_layers=2)).to("cuda")
    with torch.no_grad():
        for p in f.parameters():
            p.add_(0.1 * torch.randn_like(p))
    dev = torch.device("cuda")
    a, b = NativeTrainer(f, dev, 0.05), WideTrainer(f, dev, 0.05)
    a.pack()
    desc = a.desc()
    N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), None, 3000, N.ptr(a.gblob), N.ptr(a.loss), 0, a.stream))
    a.unpack(1.0)
    b.nll_grad(x, None, 3000)
    assert _rel(b.gtheta, a.gtheta) < 2e-4, _rel(b.gtheta, a.gtheta)
    assert abs(float(a.loss[0]) - float(b.loss[0])) < 1e-5 * abs(float(a.loss[0]))
