"""External-target path: samplers for a target that is an arbitrary Python callable (``potentials.CallablePotential``).

The reference's contract is ``target: callable [n, *event] -> [n]`` differentiated with autograd
(/root/reference/nfmc/sample.py:34-36; mcmc/langevin.py:66-68; mcmc/hmc.py:40-48).  The fused step kernels need an
analytic potential, so for a callable every step is split: the callable (and ``torch.autograd.grad``) supplies U and
grad U on the device, and the ``nfmc_ext_*`` kernels (csrc/ext_kernels.cu) do the proposal, the proposal potentials, the
leapfrog updates, the Hamiltonians, the log-ratio, the accept test, the masked overwrite, the running moments, the
counters and the sample sink -- with the same random numbers (Philox counters or injected tensors) and the same roundings
as the fused kernels.  One local step costs one U / grad U evaluation (the value at the current state travels with the
state), where the reference spends two.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N


def _step_noise(ses, stream_id: int, step: int, normals: Optional[torch.Tensor], uniforms: Optional[torch.Tensor], k: int,
                need_uniform: bool = True):
    """Noise of one step: row ``k`` of the injected tensors, or the numbers the fused kernels would draw from Philox
    (``nfmc_rng_fill`` with the same seed / step / chain keys)."""
    if normals is not None:
        return normals[k].reshape(ses.n, ses.d), (None if uniforms is None else uniforms[k].reshape(ses.n))
    nz = torch.empty(ses.n, ses.d, device=ses.device, dtype=torch.float32)
    un = torch.empty(ses.n, device=ses.device, dtype=torch.float32) if need_uniform else None
    rng = N.rng_desc(ses.seed, step)
    N.check(N.lib().nfmc_rng_fill(C.byref(rng), stream_id, ses.chain0, ses.d, ses.n, 1, N.ptr(nz), N.ptr(un), ses.stream))
    return nz, un


def _sink_ref(sink):
    return None if sink is None else C.byref(sink)


def langevin_steps(target, ses, n_steps: int, step_size: float, imd, adjusted: bool, random_walk: bool, sink,
                   normals=None, uniforms=None):
    """K Langevin (``random_walk=False``; langevin.py:61-122) or random-walk Metropolis (mh.py:44-73) steps."""
    lib = N.lib()
    n, d, s = ses.n, ses.d, ses.stream
    x = ses.x
    need_grad = not random_walk
    u, g = target.value_and_grad(x, need_grad=need_grad)
    xp = torch.empty_like(x)
    lr = torch.empty(n, device=ses.device, dtype=torch.float32)
    st = ses.stats()
    for k in range(n_steps):
        nz, un = _step_noise(ses, 0, ses.local_step + k, normals, uniforms, k, need_uniform=adjusted)
        N.check(lib.nfmc_ext_langevin_propose(N.ptr(x), N.ptr(g), N.ptr(nz), N.ptr(imd), float(step_size), int(random_walk),
                                              n, d, N.ptr(xp), s))
        if adjusted or need_grad:
            up, gp = target.value_and_grad(xp, need_grad=need_grad)
        else:
            up, gp = u, None                                           # plain random walk: nothing to evaluate
        if adjusted:
            N.check(lib.nfmc_ext_langevin_log_ratio(N.ptr(x), N.ptr(xp), N.ptr(g), N.ptr(gp), N.ptr(u), N.ptr(up), N.ptr(imd),
                                                    float(step_size), int(random_walk), n, d, N.ptr(lr), s))
        N.check(lib.nfmc_ext_accept(N.ptr(x), N.ptr(xp), N.ptr(lr), N.ptr(un), int(adjusted), n, d,
                                    N.ptr(u) if adjusted or need_grad else None, N.ptr(up) if adjusted or need_grad else None,
                                    None, None, N.ptr(g), N.ptr(gp), C.byref(st), _sink_ref(sink), k, s))


def hmc_steps(target, ses, n_steps: int, step_size: float, n_leapfrog: int, imd, adjusted: bool, sink, normals=None,
              uniforms=None):
    """K HMC steps (hmc.py:96-126).  grad U is evaluated once per trajectory point (L + 1 per step; the reference's 2L
    evaluations repeat every interior point) and both half-kicks are still applied as separate roundings."""
    lib = N.lib()
    n, d, s = ses.n, ses.d, ses.stream
    x = ses.x
    L = int(n_leapfrog)
    u0, g0 = target.value_and_grad(x)
    p = torch.empty_like(x)
    xt = torch.empty_like(x)
    kin0 = torch.empty(n, device=ses.device, dtype=torch.float32)
    lr = torch.empty(n, device=ses.device, dtype=torch.float32)
    st = ses.stats()
    for k in range(n_steps):
        nz, un = _step_noise(ses, 0, ses.local_step + k, normals, uniforms, k, need_uniform=adjusted)
        N.check(lib.nfmc_ext_hmc_momentum(N.ptr(nz), N.ptr(imd), n, d, N.ptr(p), N.ptr(kin0), s))
        xt.copy_(x)
        u1, g1 = u0, g0
        if L > 0:
            N.check(lib.nfmc_ext_hmc_leapfrog(N.ptr(xt), N.ptr(p), N.ptr(g0), N.ptr(imd), float(step_size), 1, 1, n, d, s))
            for l in range(L):
                u1, g1 = target.value_and_grad(xt)
                more = l + 1 < L
                N.check(lib.nfmc_ext_hmc_leapfrog(N.ptr(xt), N.ptr(p), N.ptr(g1), N.ptr(imd), float(step_size),
                                                  2 if more else 1, int(more), n, d, s))
        if adjusted:
            N.check(lib.nfmc_ext_hmc_log_ratio(N.ptr(p), N.ptr(imd), N.ptr(u0), N.ptr(kin0), N.ptr(u1), n, d, N.ptr(lr), s))
        if L > 0:
            N.check(lib.nfmc_ext_accept(N.ptr(x), N.ptr(xt), N.ptr(lr), N.ptr(un), int(adjusted), n, d, N.ptr(u0), N.ptr(u1),
                                        None, None, N.ptr(g0), N.ptr(g1), C.byref(st), _sink_ref(sink), k, s))
        else:
            N.check(lib.nfmc_ext_accept(N.ptr(x), N.ptr(xt), N.ptr(lr), N.ptr(un), int(adjusted), n, d, None, None, None, None,
                                        None, None, C.byref(st), _sink_ref(sink), k, s))


def ess_steps(nll, ses, n_steps: int, max_iterations: int, sink, normals=None, uniforms=None):
    """K elliptical-slice steps with prior N(0, I) around a callable negative log-likelihood (mcmc/ess.py:12-64,97-116):
    per step the ellipse direction nu (Philox stream 0 or injected), the 2 + M scalar uniforms (Philox stream 2 or injected
    ``[steps, n, 2 + M]``), then at most M bracket rounds of [rotate, evaluate the callable, shrink].  Every chain counts as
    accepted (ess.py:107); counts[3] collects the chain-steps whose bracket produced a point."""
    lib = N.lib()
    n, d, s, dev = ses.n, ses.d, ses.stream, ses.device
    M = int(max_iterations)
    n_uni = 2 + M
    f = ses.x
    u_cur = nll.value(f)
    fp = torch.empty_like(f)
    state = torch.empty(n, 4, device=dev, dtype=torch.float32)
    found = torch.empty(n, device=dev, dtype=torch.int32)
    st = ses.stats()
    for k in range(n_steps):
        step = ses.local_step + k
        nu, _ = _step_noise(ses, 0, step, normals, None, k, need_uniform=False)
        if uniforms is not None:
            un = uniforms[k].reshape(n, n_uni).contiguous()
        else:
            un = torch.empty(n, n_uni, device=dev, dtype=torch.float32)
            N.check(lib.nfmc_ext_ess_uniforms(ses.seed & 0xFFFFFFFFFFFFFFFF, step, ses.chain0, n, n_uni, N.ptr(un), s))
        N.check(lib.nfmc_ext_ess_begin(N.ptr(u_cur), N.ptr(un), n_uni, n, N.ptr(state), N.ptr(found), s))
        for it in range(M):
            if it > 0 and bool(found.all()):                             # later rounds cannot change a found chain (ess.py:50)
                break
            N.check(lib.nfmc_ext_ess_rotate(N.ptr(f), N.ptr(nu), N.ptr(state), n, d, N.ptr(fp), s))
            u_p = nll.value(fp)
            N.check(lib.nfmc_ext_ess_update(N.ptr(f), N.ptr(fp), N.ptr(u_cur), N.ptr(u_p), N.ptr(state), N.ptr(found), N.ptr(un),
                                            n_uni, it, n, d, s))
        ses.counts[3] += found.sum()
        # the mask is all ones (ess.py:107): moments, counters and the sample row of the post-step state
        N.check(lib.nfmc_ext_accept(N.ptr(f), N.ptr(f), None, None, 0, n, d, None, None, None, None, None, None, C.byref(st),
                                    _sink_ref(sink), k, s))


def flow_proposal(flow, ses, z: Optional[torch.Tensor], uniforms: Optional[torch.Tensor]):
    """x' = T^-1(z) with log q(x') for every chain (jump.py:205, imh.py:221): base draw from Philox stream 1 at the
    session's flow step, or injected.  Returns (x' [n,d], log q(x') [n], uniforms [n])."""
    if z is None:
        zz, un = _step_noise(ses, 1, ses.flow_step, None, None, 0)
    else:
        zz = z.reshape(ses.n, ses.d)
        un = None if uniforms is None else uniforms.reshape(ses.n)
    xp = torch.empty(ses.n, ses.d, device=ses.device, dtype=torch.float32)
    lqp = torch.empty(ses.n, device=ses.device, dtype=torch.float32)
    rng = N.rng_desc(ses.seed, ses.flow_step, zz, None)
    bij = flow.bijection
    if bij.row_tile_supported():            # wide / deep conditioner: row-tile fp32 kernel (csrc/train_wide.cu)
        fd, keep = bij.theta_descriptor(ses.device)
        N.check(N.lib().nfmc_flow_wide_sample(fd.d, fd.n_coupling, fd.n_linear, fd.hidden, N.ptr(keep), 1, C.byref(rng), ses.chain0,
                                              N.ptr(xp), N.ptr(lqp), ses.n, ses.stream))
        return xp, lqp, un
    fd, keep = bij.descriptor(ses.device)
    N.check(N.lib().nfmc_flow_sample(C.byref(fd), C.byref(rng), ses.chain0, N.ptr(xp), N.ptr(lqp), ses.n, ses.stream))
    return xp, lqp, un


def jump_step(target, flow, ses, adjusted: bool, sink, z=None, uniforms=None, logq: Optional[torch.Tensor] = None,
              recompute_logq: bool = True, jump_stats: bool = True, u_cache: Optional[torch.Tensor] = None):
    """One flow-proposal MH step for every chain: the NF jump (jump.py:203-243; ``logq=None``) or one IMH iteration
    (imh.py:214-249; ``logq`` = the cache that travels with the state, refreshed by a forward pass when
    ``recompute_logq``).  ``u_cache``: U at the current state (kept up to date here) or None."""
    lib = N.lib()
    n, d, s = ses.n, ses.d, ses.stream
    xp, lqp, un = flow_proposal(flow, ses, z, uniforms)
    st = ses.stats(jump=jump_stats)
    lr = None
    up = None
    if adjusted:
        if logq is None or recompute_logq:
            lq_now = flow.log_prob(ses.x.reshape(n, *flow.event_shape)).reshape(n).contiguous()   # jump.py:218, imh.py:133
            if logq is not None:
                logq.copy_(lq_now)
            else:
                logq = lq_now
        u = u_cache if u_cache is not None else target.value(ses.x)
        up = target.value(xp)
        lr = torch.empty(n, device=ses.device, dtype=torch.float32)
        N.check(lib.nfmc_ext_jump_log_ratio(N.ptr(u), N.ptr(up), N.ptr(logq), N.ptr(lqp), n, N.ptr(lr), s))
    N.check(lib.nfmc_ext_accept(N.ptr(ses.x), N.ptr(xp), N.ptr(lr), N.ptr(un), int(adjusted), n, d,
                                N.ptr(u_cache) if (u_cache is not None and up is not None) else None,
                                N.ptr(up) if (u_cache is not None and up is not None) else None,
                                N.ptr(logq) if adjusted else None, N.ptr(lqp) if adjusted else None,
                                None, None, C.byref(st), _sink_ref(sink), 0, s))


class LatentTarget:
    """``NeuTra.adjusted_target`` (neutra.py:58-68) for a callable target: U~(z) = U(T^-1 z) - log|det dT^-1/dz| with its
    gradient assembled from the flow's inverse pass (``nfmc_realnvp_inverse``), the callable's autograd gradient at
    x = T^-1 z, and the flow's reversible backward sweep seeded with it (``nfmc_neutra_pullback``).  All flow passes use the
    fp32 kernels so that value and gradient belong to the same function.  As in the reference (quirk Q1) the rows a NeuTra
    run records are the latent states."""
    external = True

    def __init__(self, target, flow, row_tile: Optional[bool] = None):
        self.target = target
        self.flow = flow
        # flow passes / backward sweep by the row-tile fp32 kernels of csrc/train_wide.cu (deep or odd-sized conditioners; also a
        # tensor-core flow whose shape the tensor-core NeuTra kernels refuse) instead of the per-chain kernels of flow.cuh
        self.row_tile = flow.bijection.uses_row_tile_pass() if row_tile is None else bool(row_tile)

    def to_data(self, z: torch.Tensor):
        """(x, log|det dT^-1/dz|) for latent rows z [n, d]."""
        bij = self.flow.bijection
        dev = z.device
        n = z.shape[0]
        x = torch.empty_like(z)
        ld = torch.empty(n, device=dev, dtype=torch.float32)
        if self.row_tile:            # deep / odd-sized conditioner: row-tile fp32 pass (csrc/train_wide.cu)
            fd, theta = bij.theta_descriptor(dev)
            N.check(N.lib().nfmc_flow_wide_pass(fd.d, fd.n_coupling, fd.n_linear, fd.hidden, N.ptr(theta), 3, N.ptr(z), N.ptr(x),
                                                N.ptr(ld), n, N.stream_ptr(dev)))
            return x, ld
        fd, keep = bij.descriptor(dev)
        N.check(N.lib().nfmc_realnvp_inverse(C.byref(fd), N.ptr(z), N.ptr(x), N.ptr(ld), n, N.stream_ptr(dev)))
        return x, ld

    def _u(self, x: torch.Tensor) -> torch.Tensor:
        return self.target.value(x) if hasattr(self.target, "value") else self.target.value_and_grad(x, need_grad=False)[0]

    def value(self, z: torch.Tensor) -> torch.Tensor:
        x, ld = self.to_data(z)
        return -((-self._u(x)) + ld)                                            # neutra.py:62-64

    def value_and_grad(self, z: torch.Tensor, need_grad: bool = True):
        if not need_grad:
            return self.value(z), None
        x, ld = self.to_data(z)
        u, gx = self.target.value_and_grad(x)
        gz = torch.empty_like(z)
        bij = self.flow.bijection
        gx = gx.reshape(z.shape).contiguous()
        if self.row_tile:
            # backward sweep of the z -> x pass seeded with grad U(x): grad_in = d/dz [U(x(z)) - log|det dx/dz|] (train_wide.cu,
            # mode SWEEP with the parameter-gradient emitters off)
            fd, theta = bij.theta_descriptor(z.device, transposed=False)
            fdt, theta_t = bij.theta_descriptor(z.device, transposed=True)
            N.check(N.lib().nfmc_flow_wide_pullback(fd.d, fd.n_coupling, fd.n_linear, fd.hidden, N.ptr(theta), N.ptr(theta_t), 1, N.ptr(x),
                                                    N.ptr(gx), z.shape[0], N.ptr(gz), N.stream_ptr(z.device)))
            return -((-u) + ld), gz
        fd, keep = bij.descriptor(z.device)
        N.check(N.lib().nfmc_neutra_pullback(C.byref(fd), N.ptr(z), N.ptr(gx), N.ptr(gz), None,
                                             z.shape[0], N.stream_ptr(z.device)))
        return -((-u) + ld), gz
