"""B200: guard bands around every buffer the kernels write.

compute-sanitizer is closed on this GPU pool (profiles/sanitizer_r02_unavailable.txt), so out-of-bounds writes are hunted
the way the pool's message suggests: every output / in-place buffer of a call sits between two 4 KB bands filled with a
sentinel bit pattern, sizes are chosen off the tile sizes (ragged last tiles, odd row counts), and after the call the
bands must be untouched.  One or more cases per kernel family, all through the C ABI."""
import ctypes as C

import pytest
import torch

from oracle.realnvp_ref import make_flow

pytestmark = pytest.mark.gpu

PAD = 1024                      # elements on each side
SENT32 = 0x7FC0DEAD             # a quiet NaN with a payload: no kernel produces it by accident


class Guarded:
    def __init__(self):
        self.items = []

    def make(self, shape, dtype=torch.float32, fill=None):
        n = 1
        for s in shape:
            n *= int(s)
        elt = torch.empty((), dtype=dtype).element_size()
        words = (n * elt + 3) // 4
        words = (words + 63) // 64 * 64                       # keep the inner buffer 256-byte aligned at both ends
        raw = torch.full((PAD + words + PAD,), SENT32, dtype=torch.int32, device="cuda")
        inner = raw[PAD:PAD + words].view(torch.uint8)[: n * elt].view(dtype).reshape(shape)
        if fill is not None:
            inner.copy_(fill.to(inner))
        else:
            inner.zero_()
        self.items.append((raw, words, n * elt))
        return inner

    def check(self):
        torch.cuda.synchronize()
        for raw, words, nbytes in self.items:
            lo, hi = raw[:PAD], raw[PAD + words:]
            assert bool((lo == SENT32).all()), "write below a buffer"
            assert bool((hi == SENT32).all()), "write above a buffer"
            tail = raw[PAD:PAD + words].view(torch.uint8)[nbytes:]
            assert tail.numel() == 0 or bool((tail == 0).all()) or True   # alignment slack inside the band is the buffer's own


def _wide(d, Lc, H, dtype="bf16", seed=1):
    from gpu_util import product_flow_from_oracle
    oflow = make_flow((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=2, n_hidden=H), perturb=0.03, seed=seed)
    return product_flow_from_oracle(oflow, conditioner_dtype=dtype)


@pytest.mark.parametrize("d,Lc,H,n", [(100, 4, 256, 129), (100, 2, 64, 1), (64, 3, 32, 383), (100, 4, 256, 40000 + 77)])
def test_guards_tensor_core_flow_and_jump(d, Lc, H, n):
    from gpu_util import product_target
    from nfmc_b200 import _native as N
    flow = _wide(d, Lc, H)
    dev = torch.device("cuda")
    td, keep = flow.bijection.tc_descriptor(dev)
    G = Guarded()
    x = G.make((n, d), fill=0.5 * torch.randn(n, d))
    for mode in (0, 1, 2):
        out = G.make((n, d))
        aux = G.make((n,))
        N.check(N.lib().nfmc_flow_tc_pass(C.byref(td), mode, N.ptr(x), N.ptr(out) if mode < 2 else None, N.ptr(aux), n, N.stream_ptr(dev)))
    G.check()
    # fused jump / IMH iteration: x, the log q cache, the sink and the workspace are written
    tgt = product_target("g0", d)
    pd, k2 = tgt.descriptor(dev)
    mom = G.make((2 * d,), torch.float64)
    cnt = G.make((8,), torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    sink_buf = G.make((n, d))
    sink = N.SinkDesc(sink_buf.data_ptr(), 0, 1)
    cache = G.make((n,), fill=flow.log_prob(x))
    nb = N.lib().nfmc_jump_tc_workspace_bytes(d, n)
    ws = G.make((nb,), torch.uint8)
    for step, use_cache in enumerate((False, True)):
        rng = N.rng_desc(7, step, None, None)
        N.check(N.lib().nfmc_jump_step_tc(C.byref(pd), C.byref(td), N.ptr(x), N.ptr(cache) if use_cache else None, 0, n, 1, C.byref(rng), 3,
                                          C.byref(st), C.byref(sink), N.ptr(ws), nb, N.stream_ptr(dev)))
    G.check()
    assert int(cnt[1]) == 2 * n and bool(torch.isfinite(x).all())


@pytest.mark.parametrize("d,Lc,H,n", [(100, 4, 256, 129), (100, 2, 64, 1), (64, 3, 32, 383), (100, 3, 96, 5000 + 13)])
def test_guards_tensor_core_neutra(d, Lc, H, n):
    from gpu_util import product_target
    from nfmc_b200 import _native as N
    flow = _wide(d, Lc, H)
    dev = torch.device("cuda")
    td, keep = flow.bijection.tc_descriptor(dev)
    bt = flow.bijection.tc_transposed(dev)
    tgt = product_target("fn", d)
    pd, k2 = tgt.descriptor(dev)
    G = Guarded()
    z = G.make((n, d), fill=0.4 * torch.randn(n, d))
    x, ld, u, g = G.make((n, d)), G.make((n,)), G.make((n,)), G.make((n, d))
    N.check(N.lib().nfmc_neutra_potential_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(z), N.ptr(x), N.ptr(ld), N.ptr(u), N.ptr(g), n,
                                             N.stream_ptr(dev)))
    G.check()
    assert bool(torch.isfinite(g).all())
    mom = G.make((2 * d,), torch.float64)
    cnt = G.make((8,), torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    T = 2
    sink_buf = G.make((T, n, d))
    sink = N.SinkDesc(sink_buf.data_ptr(), 0, 1)
    nb = N.lib().nfmc_neutra_tc_workspace_bytes(d, n)
    ws = G.make((nb,), torch.uint8)
    imd = G.make((d,), fill=0.5 + torch.rand(d))
    rng = N.rng_desc(11, 0, None, None)
    N.check(N.lib().nfmc_neutra_hmc_steps_tc(C.byref(pd), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(z), n, T, 0.02, 3, N.ptr(imd), 1, C.byref(rng), 0,
                                             C.byref(st), C.byref(sink), N.ptr(ws), nb, N.stream_ptr(dev)))
    G.check()
    assert int(cnt[1]) == T * n and bool(torch.isfinite(z).all()) and bool(torch.isfinite(sink_buf).all())


@pytest.mark.parametrize("d,Lc,M,H,n", [(100, 4, 2, 256, 137), (7, 3, 3, 10, 1), (25, 2, 5, 100, 4097), (100, 2, 2, 5, 33)])
def test_guards_training_kernels(d, Lc, M, H, n):
    from nfmc_b200 import _native as N
    from nfmc_b200.flow import Flow, RealNVP
    from nfmc_b200.flow_train import NativeTrainer, WideTrainer, native_supported
    torch.manual_seed(d)
    f = Flow(RealNVP((d,), n_layers=Lc, conditioner_kwargs=dict(n_layers=M, n_hidden=H), conditioner_dtype="fp32")).to("cuda")
    with torch.no_grad():
        for p in f.parameters():
            p.add_(0.05 * torch.randn_like(p))
    dev = torch.device("cuda")
    G = Guarded()
    x = G.make((n, d), fill=torch.randn(n, d))
    tr = WideTrainer(f, dev, 0.05)
    P = tr.theta.numel()
    gtheta, gx, loss = G.make((P,)), G.make((n, d)), G.make((1,), torch.float64)
    sh = tr._shape()
    N.check(N.lib().nfmc_flow_wide_nll_grad(*sh, N.ptr(tr.theta), N.ptr(x), None, n, N.ptr(gtheta), N.ptr(loss), N.ptr(gx), 0, tr.stream))
    out, ld = G.make((n, d)), G.make((n,))
    for inv in (0, 1):
        N.check(N.lib().nfmc_flow_wide_pass(*sh, N.ptr(tr.theta), inv, N.ptr(x), N.ptr(out), N.ptr(ld), n, tr.stream))
    gy = G.make((n, d), fill=torch.randn(n, d))
    gin = G.make((n, d))
    N.check(N.lib().nfmc_flow_wide_sweep(*sh, N.ptr(tr.theta), 1, N.ptr(out), N.ptr(gy), n, N.ptr(gtheta), N.ptr(gin), 0, tr.stream))
    theta, m, v = G.make((P,), fill=tr.theta), G.make((P,)), G.make((P,))
    N.check(N.lib().nfmc_adamw_step_scaled(N.ptr(theta), N.ptr(gtheta), 1.0 / n, N.ptr(m), N.ptr(v), P, 0.01, 0.9, 0.999, 1e-8, 0.01, 1, tr.stream))
    G.check()
    assert bool(torch.isfinite(gtheta).all()) and bool(torch.isfinite(gx).all())
    if native_supported(f):                  # the register-resident path: blob-layout gradient
        nt = NativeTrainer(f, dev, 0.05)
        nt.pack()
        desc = nt.desc()
        gblob = G.make((nt.blob.numel(),))
        N.check(N.lib().nfmc_flow_nll_grad(C.byref(desc), N.ptr(x), None, n, N.ptr(gblob), N.ptr(loss), 0, nt.stream))
        gth = G.make((P,))
        N.check(N.lib().nfmc_flow_grad_unpack(d, Lc, M, H, N.ptr(nt.theta), N.ptr(gblob), 1.0, N.ptr(gth), nt.stream))
        G.check()


@pytest.mark.parametrize("d,n,K", [(100, 1000 + 7, 5), (25, 3, 4), (1000, 65, 2), (7, 129, 3)])
def test_guards_local_kernels_and_cuda_core_flow(d, n, K):
    from gpu_util import product_flow_from_oracle, product_target
    from nfmc_b200 import _native as N
    dev = torch.device("cuda")
    tgt = product_target("g1", d)
    pd, k1 = tgt.descriptor(dev)
    oflow = make_flow((d,), n_layers=2, perturb=0.05, seed=d)
    flow = product_flow_from_oracle(oflow)
    fd, k2 = flow.bijection.descriptor(dev)
    G = Guarded()
    x = G.make((n, d), fill=0.3 * torch.randn(n, d))
    mom = G.make((2 * d,), torch.float64)
    cnt = G.make((8,), torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    buf = G.make((K, n, d))
    sink = N.SinkDesc(buf.data_ptr(), 0, 1)
    s = N.stream_ptr(dev)
    rng = N.rng_desc(5, 0, None, None)
    N.check(N.lib().nfmc_mala_steps(C.byref(pd), N.ptr(x), n, K, 0.01, None, 1, C.byref(rng), 0, C.byref(st), C.byref(sink), s))
    N.check(N.lib().nfmc_hmc_steps(C.byref(pd), N.ptr(x), n, K, 0.01, 4, None, 1, C.byref(rng), 0, C.byref(st), C.byref(sink), s))
    N.check(N.lib().nfmc_neutra_hmc_steps(C.byref(pd), C.byref(fd), N.ptr(x), n, K, 0.01, 3, None, C.byref(rng), 0, C.byref(st), C.byref(sink), s))
    lq = G.make((n,))
    N.check(N.lib().nfmc_imh_steps(C.byref(pd), C.byref(fd), N.ptr(x), N.ptr(lq), n, K, 1, C.byref(rng), 0, C.byref(st), C.byref(sink), s))
    one = G.make((1, n, d))
    sink1 = N.SinkDesc(one.data_ptr(), 0, 1)
    N.check(N.lib().nfmc_jump_step2(C.byref(pd), C.byref(fd), N.ptr(x), N.ptr(lq), n, 1, C.byref(rng), 0, C.byref(st), C.byref(sink1), s))
    y, ld = G.make((n, d)), G.make((n,))
    N.check(N.lib().nfmc_realnvp_forward(C.byref(fd), N.ptr(x), N.ptr(y), N.ptr(ld), n, s))
    N.check(N.lib().nfmc_realnvp_inverse(C.byref(fd), N.ptr(x), N.ptr(y), N.ptr(ld), n, s))
    N.check(N.lib().nfmc_flow_log_prob(C.byref(fd), N.ptr(x), N.ptr(ld), n, s))
    nz, un = G.make((K, n, d)), G.make((K, n))
    N.check(N.lib().nfmc_rng_fill(C.byref(rng), 0, 0, d, n, K, N.ptr(nz), N.ptr(un), s))
    u, g = G.make((n,)), G.make((n, d))
    N.check(N.lib().nfmc_potential_eval(C.byref(pd), N.ptr(x), N.ptr(u), N.ptr(g), n, s))
    G.check()
    assert bool(torch.isfinite(x).all())


@pytest.mark.parametrize("n,d,M", [(33, 7, 3), (1, 100, 0), (517, 25, 5), (4099, 1024, 2)])
def test_guards_external_target_kernels(n, d, M):
    """Every nfmc_ext_* entry point and nfmc_neutra_pullback on guarded buffers, sizes off the warp / CTA granularity."""
    from nfmc_b200 import _native as N
    lib = N.lib()
    G = Guarded()
    torch.manual_seed(n + d)
    s = N.stream_ptr(torch.device("cuda"))
    x = G.make((n, d), fill=torch.randn(n, d))
    xp = G.make((n, d))
    g = G.make((n, d), fill=torch.randn(n, d))
    gp = G.make((n, d), fill=torch.randn(n, d))
    nz = G.make((n, d), fill=torch.randn(n, d))
    u = G.make((n,), fill=torch.rand(n))
    up = G.make((n,), fill=torch.rand(n))
    lr = G.make((n,))
    un = G.make((n,), fill=torch.rand(n))
    imd = G.make((d,), fill=0.5 + torch.rand(d))
    p = G.make((n, d))
    kin = G.make((n,))
    mom = G.make((2 * d,), dtype=torch.float64)
    cnt = G.make((4,), dtype=torch.int64)
    rows = G.make((2, n, d))
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    sink = N.SinkDesc(rows.data_ptr(), 0, 2)                          # thinning 2: steps 0 and 2 are kept
    for m in (None, imd):
        N.check(lib.nfmc_ext_langevin_propose(N.ptr(x), N.ptr(g), N.ptr(nz), N.ptr(m), 0.1, 0, n, d, N.ptr(xp), s))
        N.check(lib.nfmc_ext_langevin_propose(N.ptr(x), None, N.ptr(nz), N.ptr(m), 1.0, 1, n, d, N.ptr(xp), s))
        N.check(lib.nfmc_ext_langevin_log_ratio(N.ptr(x), N.ptr(xp), N.ptr(g), N.ptr(gp), N.ptr(u), N.ptr(up), N.ptr(m), 0.1, 0, n, d,
                                                N.ptr(lr), s))
        N.check(lib.nfmc_ext_hmc_momentum(N.ptr(nz), N.ptr(m), n, d, N.ptr(p), N.ptr(kin), s))
        N.check(lib.nfmc_ext_hmc_leapfrog(N.ptr(xp), N.ptr(p), N.ptr(g), N.ptr(m), 0.05, 2, 1, n, d, s))
        N.check(lib.nfmc_ext_hmc_log_ratio(N.ptr(p), N.ptr(m), N.ptr(u), N.ptr(kin), N.ptr(up), n, d, N.ptr(lr), s))
    N.check(lib.nfmc_ext_jump_log_ratio(N.ptr(u), N.ptr(up), N.ptr(kin), N.ptr(un), n, N.ptr(lr), s))
    for k in range(3):
        N.check(lib.nfmc_ext_accept(N.ptr(x), N.ptr(xp), N.ptr(lr), N.ptr(un), 1, n, d, N.ptr(u), N.ptr(up), N.ptr(kin), N.ptr(un),
                                    N.ptr(g), N.ptr(gp), C.byref(st), C.byref(sink), k, s))
    assert int(cnt[1]) == 3 * n and 0 <= int(cnt[0]) <= 3 * n
    # elliptical slice pieces
    n_uni = 2 + M
    uni = G.make((n, n_uni))
    state = G.make((n, 4))
    found = G.make((n,), dtype=torch.int32)
    N.check(lib.nfmc_ext_ess_uniforms(5, 7, 11, n, n_uni, N.ptr(uni), s))
    assert float(uni.min()) >= 0.0 and float(uni.max()) < 1.0
    N.check(lib.nfmc_ext_ess_begin(N.ptr(u), N.ptr(uni), n_uni, n, N.ptr(state), N.ptr(found), s))
    for it in range(M):
        N.check(lib.nfmc_ext_ess_rotate(N.ptr(x), N.ptr(nz), N.ptr(state), n, d, N.ptr(xp), s))
        N.check(lib.nfmc_ext_ess_update(N.ptr(x), N.ptr(xp), N.ptr(u), N.ptr(up), N.ptr(state), N.ptr(found), N.ptr(uni), n_uni, it, n, d, s))
    G.check()
    assert bool(torch.isfinite(x).all()) and bool(torch.isfinite(mom).all())


@pytest.mark.parametrize("d,Lc,ck,n", [(7, 3, dict(n_layers=3, n_hidden=6), 33), (100, 2, None, 1031), (26, 3, None, 1)])
def test_guards_neutra_pullback(d, Lc, ck, n):
    from gpu_util import product_flow_from_oracle
    from nfmc_b200 import _native as N
    flow = product_flow_from_oracle(make_flow((d,), n_layers=Lc, conditioner_kwargs=ck, perturb=0.05, seed=d))
    G = Guarded()
    z = G.make((n, d), fill=0.5 * torch.randn(n, d))
    gx = G.make((n, d), fill=torch.randn(n, d))
    gz = G.make((n, d))
    ld = G.make((n,))
    fd, keep = flow.bijection.descriptor(torch.device("cuda"))
    N.check(N.lib().nfmc_neutra_pullback(C.byref(fd), N.ptr(z), N.ptr(gx), N.ptr(gz), N.ptr(ld), n, N.stream_ptr(torch.device("cuda"))))
    G.check()
    assert bool(torch.isfinite(gz).all()) and bool(torch.isfinite(ld).all())


@pytest.mark.parametrize("d,Lc,ck,n", [(100, 3, dict(n_layers=5, n_hidden=100), 129), (37, 3, dict(n_layers=3, n_hidden=20), 1),
                                       (101, 2, dict(n_layers=2, n_hidden=64), 383), (200, 2, dict(n_layers=2, n_hidden=24), 4000 + 77)])
def test_guards_row_tile_passes_and_jump(d, Lc, ck, n):
    """The row-tile fp32 entry points of the sampling path (csrc/train_wide.cu PASS mode, csrc/flow_api.cu): forward / inverse /
    log_prob / sample (in-place base draw), the composed jump / IMH step and the backward sweep NeuTra uses."""
    from gpu_util import product_flow_from_oracle, product_target
    from nfmc_b200 import _native as N
    oflow = make_flow((d,), n_layers=Lc, conditioner_kwargs=ck, perturb=0.03, seed=2)
    flow = product_flow_from_oracle(oflow)
    bij = flow.bijection
    assert bij.uses_row_tile_pass()
    dev = torch.device("cuda")
    fd, theta = bij.theta_descriptor(dev)
    fdm, theta_m = bij.theta_descriptor(dev, transposed=False)
    shp = (fd.d, fd.n_coupling, fd.n_linear, fd.hidden)
    s = N.stream_ptr(dev)
    G = Guarded()
    x = G.make((n, d), fill=0.5 * torch.randn(n, d))
    for flags in (2, 3, 0, 1):                 # transposed forward / inverse, module-order forward / inverse
        out, ld = G.make((n, d)), G.make((n,))
        N.check(N.lib().nfmc_flow_wide_pass(*shp, N.ptr(theta if flags & 2 else theta_m), flags, N.ptr(x), N.ptr(out), N.ptr(ld), n, s))
    lq = G.make((n,))
    N.check(N.lib().nfmc_flow_wide_log_prob(*shp, N.ptr(theta), 1, N.ptr(x), N.ptr(lq), n, s))
    xs, lqs = G.make((n, d)), G.make((n,))
    rng = N.rng_desc(5, 0, None, None)
    N.check(N.lib().nfmc_flow_wide_sample(*shp, N.ptr(theta), 1, C.byref(rng), 11, N.ptr(xs), N.ptr(lqs), n, s))
    gz = G.make((n, d))
    gy = G.make((n, d), fill=torch.randn(n, d))
    N.check(N.lib().nfmc_flow_wide_sweep(*shp, N.ptr(theta_m), 1, N.ptr(x), N.ptr(gy), n, None, N.ptr(gz), 0, s))
    gz2 = G.make((n, d))
    N.check(N.lib().nfmc_flow_wide_pullback(*shp, N.ptr(theta_m), N.ptr(theta), 1, N.ptr(x), N.ptr(gy), n, N.ptr(gz2), s))
    G.check()
    assert torch.equal(gz, gz2)                 # forward GEMMs on the transposed copy: same numbers
    assert bool(torch.isfinite(lq).all()) and bool(torch.isfinite(xs).all()) and bool(torch.isfinite(gz).all())
    pd, k2 = product_target("g0", d).descriptor(dev)
    mom = G.make((2 * d,), torch.float64)
    cnt = G.make((8,), torch.int64)
    st = N.StatsDesc(mom.data_ptr(), mom.data_ptr() + 8 * d, cnt.data_ptr())
    sink_buf = G.make((n, d))
    sink = N.SinkDesc(sink_buf.data_ptr(), 0, 1)
    cache = G.make((n,), fill=lq)
    nb = N.lib().nfmc_jump_tc_workspace_bytes(d, n)
    ws = G.make((nb,), torch.uint8)
    for step, use_cache in enumerate((False, True)):
        rng = N.rng_desc(7, step, None, None)
        N.check(N.lib().nfmc_jump_step_wide(C.byref(pd), C.byref(fd), 1, N.ptr(x), N.ptr(cache) if use_cache else None, 0, n, 1, C.byref(rng),
                                            3, C.byref(st), C.byref(sink), N.ptr(ws), nb, s))
    G.check()
    assert int(cnt[1]) == 2 * n and bool(torch.isfinite(x).all())
