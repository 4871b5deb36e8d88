"""Oracle (test infrastructure): the reference's sampler arithmetic restated op-for-op in torch fp32 on CPU.

PINNED: ``tests/test_oracle_golden.py`` replays the fixtures in ``tests/golden/`` (made by importing the
unmodified reference, ``tests/golden/make_golden.py``) through these functions.

Each function cites the reference lines it follows (paths under /root/reference/nfmc/algorithms/sampling/).
Random numbers come from a *draw source* so the same numbers can be injected into the CUDA kernels
(the reference has no injection hook; it draws inline from the global generator in this order:
``mcmc/langevin.py:63,106``, ``mcmc/hmc.py:100,112``, ``mcmc/ess.py:32,35,39,58,126``, ``nfmc/jump.py:205,225``,
``nfmc/imh.py:221,229``).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import torch


# ---------------------------------------------------------------------------------------------------------
# draw sources
# ---------------------------------------------------------------------------------------------------------
class GlobalDraws:
    """Draw from torch's global CPU generator exactly as the reference does, optionally recording."""

    def __init__(self, record: bool = False):
        self.record = record
        self.normals: List[torch.Tensor] = []
        self.uniforms: List[torch.Tensor] = []

    def normal(self, n, event_shape):
        v = torch.randn(size=(n, *event_shape))
        if self.record:
            self.normals.append(v.clone())
        return v

    def uniform(self, n):
        v = torch.rand(n)
        if self.record:
            self.uniforms.append(v.clone())
        return v

    def uniform_scalar(self):
        v = torch.rand(size=())
        if self.record:
            self.uniforms.append(v.clone())
        return v


class TapeDraws:
    """Replay pre-drawn numbers (``normals``: list of [n,*event]; ``uniforms``: list of [n])."""

    def __init__(self, normals, uniforms):
        self.normals = list(normals)
        self.uniforms = list(uniforms)
        self.i_n = 0
        self.i_u = 0

    def normal(self, n, event_shape):
        v = self.normals[self.i_n]
        self.i_n += 1
        assert v.shape == (n, *event_shape)
        return v.clone()

    def uniform(self, n):
        v = self.uniforms[self.i_u]
        self.i_u += 1
        assert v.shape == (n,)
        return v.clone()

    def uniform_scalar(self):
        v = self.uniforms[self.i_u]
        self.i_u += 1
        assert v.numel() == 1
        return v.reshape(()).clone()


# ---------------------------------------------------------------------------------------------------------
# bookkeeping (sampling/base.py)
# ---------------------------------------------------------------------------------------------------------
class StreamingMean:
    """``MCMCExpectation.update`` (sampling/base.py:75-95): running mean of f(x) over (iteration, chain)."""

    def __init__(self, f: Callable):
        self.f = f
        self.n_seen = 0
        self.value = 0.0

    def update(self, x: torch.Tensor, event_ndim: int):
        if x.ndim == event_ndim + 1:
            x = x[None]
        n_new = x.shape[0] * x.shape[1]
        self.value = torch.add(
            self.n_seen / (self.n_seen + n_new) * self.value,
            n_new / (self.n_seen + n_new) * torch.mean(self.f(x.detach()), dim=(0, 1)),
        )
        self.n_seen += n_new


@dataclass
class RunRef:
    """What one oracle run returns (the fields of MCMCOutput / MCMCStatistics that carry numbers)."""
    event_shape: tuple
    x: torch.Tensor = None                      # last state
    samples: Optional[torch.Tensor] = None      # [rows, n, *event]
    n_accepted: int = 0
    n_attempted: int = 0
    n_divergences: int = 0
    n_target_calls: int = 0
    n_grad_calls: int = 0
    n_accepted_jumps: int = 0
    n_attempted_jumps: int = 0
    first: StreamingMean = field(default_factory=lambda: StreamingMean(lambda v: v))
    second: StreamingMean = field(default_factory=lambda: StreamingMean(lambda v: v ** 2))
    rows: list = field(default_factory=list)
    trace: dict = field(default_factory=dict)   # per-step diagnostics for parity tests

    def observe(self, x: torch.Tensor, store: bool):
        k = len(self.event_shape)
        self.first.update(x, k)
        self.second.update(x, k)
        if store:
            blk = x if x.ndim == k + 2 else x[None]
            self.rows.extend(blk.detach().clone())

    def finish(self, store: bool):
        if store and self.rows:
            self.samples = torch.stack(self.rows, dim=0)
        return self

    @property
    def mean(self):
        return self.first.value

    @property
    def second_moment(self):
        return self.second.value

    @property
    def variance(self):
        return self.second.value - self.first.value ** 2


# ---------------------------------------------------------------------------------------------------------
# util.py:382-392
# ---------------------------------------------------------------------------------------------------------
def mh_log_ratio(logp_curr, logp_prime, logq_curr, logq_prime):
    return logp_prime - logp_curr + logq_curr - logq_prime


def value_and_grad(target, x):
    """``u = target(x)`` and ``grad u.sum()`` by autograd (langevin.py:66-70, hmc.py:40-48)."""
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        u = target(xr)
        g, = torch.autograd.grad(u.sum(), xr)
    return u.detach(), g.detach()


# ---------------------------------------------------------------------------------------------------------
# local kernels
# ---------------------------------------------------------------------------------------------------------
def langevin_q(x_to, x_from, grad_from, a_diag, tau):
    """``proposal_potential`` (mcmc/langevin.py:31-42)."""
    term = x_to - x_from + tau * a_diag.view(1, -1) * grad_from
    return (term * (1 / a_diag.view(1, -1)) * term).sum(dim=-1) / (4 * tau)


def mala_propose(x, target, tau: float, imd: torch.Tensor, draws, adjusted: bool = True, trace=None):
    """One Langevin proposal (mcmc/langevin.py:61-122).  ``x`` is ``[n, d]`` (Langevin is 1-D-event only, Q8)."""
    n = x.shape[0]
    noise = draws.normal(n, x.shape[1:])                                       # :63
    u_x, g_x = value_and_grad(target, x)                                       # :66-68
    grad_term = -tau / imd[None].square() * g_x                                # :74
    noise_term = math.sqrt(2 * tau) / imd[None] * noise                        # :75
    x_prime = x + grad_term + noise_term                                       # :76
    if adjusted:
        u_p, g_p = value_and_grad(target, x_prime)                             # :80-82
        a = 1 / imd ** 2
        log_ratio = mh_log_ratio(-u_x, -u_p, -langevin_q(x, x_prime, g_p, a, tau),
                                 -langevin_q(x_prime, x, g_x, a, tau))         # :88-105
        u = draws.uniform(n)
        mask = torch.log(u) < log_ratio                                        # :106
        calls = grads = 2 * n                                                  # :116-120
        if trace is not None:
            trace.setdefault("log_ratio", []).append(log_ratio.clone())
            trace.setdefault("u_prime", []).append(u_p.clone())
    else:
        mask = torch.ones(n, dtype=torch.bool)                                 # :109
        calls = grads = n
    if trace is not None:
        trace.setdefault("x_prime", []).append(x_prime.clone())
        trace.setdefault("mask", []).append(mask.clone())
    return x_prime, mask, calls, grads


def hmc_propose(x, target, tau: float, imd: torch.Tensor, n_leapfrog: int, draws, adjusted: bool = True,
                trace=None):
    """One HMC proposal (mcmc/hmc.py:96-126 with the b-a-b trajectory of :51-77)."""
    n = x.shape[0]
    event_shape = x.shape[1:]
    d = int(math.prod(event_shape))

    def mass_mul(v, diag):                                                     # :26-37
        return torch.einsum('...i,i->...i', v.reshape(n, d), diag.to(v)).view_as(v)

    p0 = mass_mul(draws.normal(n, event_shape), 1 / imd.sqrt())                # :100
    xs, p = x, p0
    for _ in range(n_leapfrog):                                                # :68-71
        p = p - tau / 2 * value_and_grad(target, xs)[1]                        # :51-53
        xs = xs + tau * mass_mul(p, imd)                                       # :56-58
        p = p - tau / 2 * value_and_grad(target, xs)[1]
    x_prime, p_prime = xs, p
    if adjusted:
        h0 = target(x) + 0.5 * mass_mul(p0 ** 2, imd).reshape(n, -1).sum(dim=-1)          # :103-106
        h1 = target(x_prime) + 0.5 * mass_mul(p_prime ** 2, imd).reshape(n, -1).sum(dim=-1)  # :107-110
        log_accept = -h1 - (-h0)                                               # :111
        u = draws.uniform(n)
        mask = torch.log(u) < log_accept                                       # :112-113
        if trace is not None:
            trace.setdefault("log_ratio", []).append(log_accept.detach().clone())
    else:
        mask = torch.ones(n, dtype=torch.bool)
    calls = 2 * n_leapfrog * n + (2 * n if adjusted else 0)                    # :122-125
    grads = 2 * n_leapfrog * n
    if trace is not None:
        trace.setdefault("x_prime", []).append(x_prime.detach().clone())
        trace.setdefault("mask", []).append(mask.clone())
    return x_prime.detach(), mask, calls, grads


def mh_propose(x, target, imd: torch.Tensor, draws, adjusted: bool = True, trace=None):
    """One random-walk Metropolis proposal (mcmc/mh.py:44-73)."""
    n = x.shape[0]
    noise = torch.multiply(draws.normal(n, (int(math.prod(x.shape[1:])),)), imd[None]).view_as(x)   # :52-55
    x_prime = x + noise                                                        # :56
    if adjusted:
        with torch.no_grad():
            log_ratio = mh_log_ratio(-target(x), -target(x_prime), 0, 0)       # :59
        u = draws.uniform(n)
        mask = torch.log(u) < log_ratio                                        # :60
        calls = 2 * n
        if trace is not None:
            trace.setdefault("log_ratio", []).append(log_ratio.clone())
    else:
        mask = torch.ones(n, dtype=torch.bool)
        calls = 0
    if trace is not None:
        trace.setdefault("x_prime", []).append(x_prime.clone())
        trace.setdefault("mask", []).append(mask.clone())
    return x_prime.detach(), mask, calls, 0


def run_mh(x0, target, imd, n_steps, draws, adjusted=True, store=True, trace=False) -> "RunRef":
    return run_local(x0, lambda x, tr: mh_propose(x, target, imd, draws, adjusted, tr), n_steps, store, trace)


def ess_propose(f, nll, max_iterations: int, draws, trace=None):
    """One elliptical-slice step with identity prior covariance (mcmc/ess.py:12-64) wrapped as ``ESS.propose``
    (mcmc/ess.py:97-116): the returned mask is all ones ("technical hack", :107), chains whose bracket never
    produced an acceptable point within ``max_iterations`` keep their state.

    The reference updates ``f`` in place through the alias ``f_proposed = f`` (:43,50); a chain accepted in an earlier
    round therefore feeds its *new* state into later ``f_prime`` evaluations, whose result is discarded (:50 only
    assigns to chains not accepted before) -- so the outcome is "first acceptable point wins", restated here without
    the aliasing.
    """
    n = f.shape[0]
    ev = tuple(f.shape[1:])
    ones = [1] * len(ev)
    nu = draws.normal(n, ev)                                                   # :32 (cov = None -> util.py:411)
    u = draws.uniform(n)                                                       # :35
    with torch.no_grad():
        log_y = -nll(f) + torch.log(u)                                         # :36
    theta = draws.uniform(n).view(n, *ones) * 2 * torch.pi                     # :39
    theta_min = theta - 2 * torch.pi                                           # :40
    theta_max = theta.clone()                                                  # :41 (alias of theta; rebound at :59)
    accepted = torch.zeros(n, dtype=torch.bool)
    out = f.clone()
    for _ in range(max_iterations):
        f_prime = out * torch.cos(theta) + nu * torch.sin(theta)               # :46
        with torch.no_grad():
            upd = -nll(f_prime) > log_y                                        # :47
        take = upd & (~accepted)
        out[take] = f_prime[take]                                              # :50
        neg = theta < 0                                                        # :53
        theta_min = torch.where(neg, theta, theta_min)                         # :54
        theta_max = torch.where(~neg, theta, theta_max)                        # :55
        theta = draws.uniform(n).view(n, *ones) * (theta_max - theta_min) + theta_min   # :58-59
        accepted = accepted | upd                                              # :62
    if trace is not None:
        trace.setdefault("ess_accepted", []).append(accepted.clone())
    mask = torch.ones(n, dtype=torch.bool)                                     # :107
    return out.detach(), mask, (max_iterations + 1) * n, 0                     # :114-116


def run_ess(x0, nll, n_steps, draws, max_iterations: int = 5, store=True, trace=False) -> "RunRef":
    """``ESS.sample`` (mcmc/ess.py:121-127): the given ``x0`` only supplies the number of chains -- the run starts
    from a fresh prior draw."""
    x0 = draws.normal(x0.shape[0], tuple(x0.shape[1:]))                        # :126
    return run_local(x0, lambda x, tr: ess_propose(x, nll, max_iterations, draws, tr), n_steps, store, trace)


def run_local(x0, propose: Callable, n_steps: int, store: bool = True, trace: bool = False) -> RunRef:
    """``MCMCSampler.sample`` (mcmc/base.py:56-102) without tuning; ``propose(x, trace) -> (x', mask, calls, grads)``."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    x = x0.clone().detach()
    tr = out.trace if trace else None
    for _ in range(n_steps):
        x_prime, mask, calls, grads = propose(x, tr)
        x = x.detach()
        x[mask] = x_prime[mask]                                                # :77
        out.n_target_calls += calls
        out.n_grad_calls += grads
        out.n_accepted += int(torch.sum(mask))
        out.n_attempted += n                                                   # :79-85
        out.observe(x, store)                                                  # :86,90
    out.x = x
    return out.finish(store)


def run_mala(x0, target, tau, imd, n_steps, draws, adjusted=True, store=True, trace=False) -> RunRef:
    return run_local(x0, lambda x, tr: mala_propose(x, target, tau, imd, draws, adjusted, tr), n_steps, store, trace)


def run_hmc(x0, target, tau, imd, n_leapfrog, n_steps, draws, adjusted=True, store=True, trace=False) -> RunRef:
    return run_local(x0, lambda x, tr: hmc_propose(x, target, tau, imd, n_leapfrog, draws, adjusted, tr),
                     n_steps, store, trace)


# ---------------------------------------------------------------------------------------------------------
# flow-side helpers (draws routed through the draw source instead of ``flow.sample``'s inline randn)
# ---------------------------------------------------------------------------------------------------------
@torch.no_grad()
def flow_sample_with_logq(flow, n, draws):
    """``flow.sample(n, return_log_prob=True)`` (jump.py:205, imh.py:221) with the base draw injected."""
    z = draws.normal(n, flow.event_shape)
    x, log_det = flow.bijection.inverse(z)
    return x, flow.base_log_prob(z) - log_det


# ---------------------------------------------------------------------------------------------------------
# NF outer loops
# ---------------------------------------------------------------------------------------------------------
def run_jump(x0, target, flow, inner: str, n_outer: int, n_inner: int, draws, tau: float, imd: torch.Tensor,
             n_leapfrog: int = 20, adjusted_jumps: bool = True, inner_adjusted: bool = True,
             store: bool = True, trace: bool = False, nll=None, max_ess_iterations: int = 5) -> RunRef:
    """``JumpNFMC.sample`` (nfmc/jump.py:156-246), frozen flow (``fit_nf=False``)."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    x = x0.clone()
    for _ in range(n_outer):
        if inner == "mala":
            loc = run_mala(x, target, tau, imd, n_inner, draws, inner_adjusted, store=True, trace=trace)
        elif inner == "hmc":
            loc = run_hmc(x, target, tau, imd, n_leapfrog, n_inner, draws, inner_adjusted, store=True, trace=trace)
        elif inner == "ess":      # JumpESS (jump.py:309-319): every local stage restarts from the prior (ess.py:126)
            loc = run_ess(x, nll, n_inner, draws, max_ess_iterations, store=True, trace=trace)
        else:
            raise ValueError(inner)
        out.n_accepted += loc.n_accepted
        out.n_attempted += loc.n_attempted
        out.n_target_calls += loc.n_target_calls
        out.n_grad_calls += loc.n_grad_calls                                   # :180-186
        out.observe(loc.samples, store)                                        # :188-189
        if trace:
            for k, v in loc.trace.items():
                out.trace.setdefault(k, []).extend(v)

        x_prime, f_prime = flow_sample_with_logq(flow, n, draws)               # :205-207
        x = loc.x.clone()                                                      # :209
        if adjusted_jumps:
            with torch.no_grad():
                u_x = target(x)
                u_p = target(x_prime)                                          # :212-213
                out.n_target_calls += 2 * n                                    # :214-216
                f_x = flow.log_prob(x)                                         # :218
            log_alpha = mh_log_ratio(-u_x, -u_p, f_x, f_prime)                 # :219-224
            u = draws.uniform(n)
            mask = torch.log(u) < log_alpha                                    # :225
            if trace:
                out.trace.setdefault("jump_log_alpha", []).append(log_alpha.clone())
                out.trace.setdefault("jump_x_prime", []).append(x_prime.clone())
                out.trace.setdefault("jump_mask", []).append(mask.clone())
        else:
            mask = torch.ones(n, dtype=torch.bool)                             # :229
        x[mask] = x_prime[mask]                                                # :231
        out.n_attempted_jumps += n
        out.n_accepted_jumps += int(torch.sum(mask))                           # :236-239
        out.observe(x, store)                                                  # :240,243
    out.x = x
    return out.finish(store)


def run_fixed_imh(x0, target, flow, n_iterations: int, draws, store: bool = True, trace: bool = False) -> RunRef:
    """``FixedIMH.sample`` (nfmc/imh.py:200-255)."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    x = x0.clone()
    with torch.no_grad():
        f_x = flow.log_prob(x)                                                 # :214
        for _ in range(n_iterations):
            x_prime, f_prime = flow_sample_with_logq(flow, n, draws)           # :221
            log_alpha = mh_log_ratio(-target(x), -target(x_prime), f_x, f_prime)   # :223-228
            u = draws.uniform(n)
            mask = torch.less(torch.log(u), log_alpha)                         # :229-230
            x[mask] = x_prime[mask]
            f_x[mask] = f_prime[mask]                                          # :232-233
            out.n_target_calls += 2 * n
            out.n_accepted += int(torch.sum(mask))
            out.n_attempted += n                                               # :243-247
            out.observe(x, store)                                              # :242,249
            if trace:
                out.trace.setdefault("log_alpha", []).append(log_alpha.clone())
                out.trace.setdefault("x_prime", []).append(x_prime.clone())
                out.trace.setdefault("mask", []).append(mask.clone())
    out.x = x
    return out.finish(store)


def neutra_potential(flow, target):
    """``NeuTra.adjusted_target`` (nfmc/neutra.py:58-68): U~(z) = U(T^-1 z) - log|det dT^-1/dz|."""

    def adjusted(z):
        x, log_det_inverse = flow.bijection.inverse(z)
        log_prob = -target(x)
        return -(log_prob + log_det_inverse.to(log_prob))

    return adjusted


def run_neutra_hmc(z0, target, flow, n_iterations: int, draws, tau: float, imd: torch.Tensor,
                   n_leapfrog: int = 20, store: bool = True, trace: bool = False) -> RunRef:
    """``NeuTra.sample`` (nfmc/neutra.py:109-129): HMC on the latent potential; outputs stay in z-space (Q1)."""
    return run_hmc(z0, neutra_potential(flow, target), tau, imd, n_leapfrog, n_iterations, draws,
                   adjusted=True, store=store, trace=trace)


def run_neutra_mh(z0, target, flow, n_iterations: int, draws, imd: torch.Tensor, adjusted: bool = True,
                  store: bool = True, trace: bool = False) -> RunRef:
    """``NeuTraMH`` (nfmc/neutra.py:147-159): random-walk Metropolis (mcmc/mh.py:44-73) on the latent potential."""
    return run_mh(z0, neutra_potential(flow, target), imd, n_iterations, draws, adjusted=adjusted, store=store, trace=trace)


# ---------------------------------------------------------------------------------------------------------
# transport elliptical slice sampling (nfmc/tess.py)
# ---------------------------------------------------------------------------------------------------------
@torch.no_grad()
def tess_step(u, flow, potential, max_iterations: int, draws, trace=None):
    """``transport_elliptical_slice_sampling_step`` (nfmc/tess.py:15-75), identity covariance.  Quirks kept: the initial
    angle is a NORMAL draw times 2 pi (:44), the log-det enters ``log_pi_hat`` with the sign of :31, and ``theta_max``
    aliases the initial angle (:48), so a negative initial angle collapses the bracket onto itself."""
    n = u.shape[0]
    ev = tuple(u.shape[1:])
    ones = [1] * len(ev)

    def log_pi_hat(inp):
        x, log_det = flow.bijection.inverse(inp)
        return -potential(x) - log_det                                          # :29-32

    v = draws.normal(n, ev)                                                      # :38
    w = draws.uniform(n)                                                         # :41
    log_s = log_pi_hat(u) + flow.base_log_prob(v) + w.log()                      # :42
    theta = (draws.normal(n, ()) * (2 * torch.pi)).view(n, *ones)                # :45-47
    theta_min, theta_max = theta - 2 * torch.pi, theta.clone()                   # :48
    accepted = torch.zeros(n, dtype=torch.bool)
    u_prop = u.clone()
    x_prop = flow.bijection.inverse(u_prop)[0]                                   # :51-52
    for _ in range(max_iterations):
        u_prime = u * torch.cos(theta) + v * torch.sin(theta)                    # :54
        v_prime = v * torch.cos(theta) - u * torch.sin(theta)                    # :55
        x_prime = flow.bijection.inverse(u_prime)[0]                             # :56
        upd = (log_pi_hat(u_prime) + flow.base_log_prob(v_prime)) > log_s        # :57
        take = upd & (~accepted)
        x_prop[take] = x_prime[take]                                             # :60
        u_prop[take] = u_prime[take]                                             # :61
        neg = theta < 0                                                          # :64
        theta_min = torch.where(neg, theta, theta_min)                           # :65
        theta_max = torch.where(~neg, theta, theta_max)                          # :66
        theta = draws.uniform(n).view(n, *ones) * (theta_max - theta_min) + theta_min   # :69-70
        accepted = accepted | upd                                                # :73
    return x_prop.detach(), u_prop.detach(), accepted


def run_tess(x0, potential, flow, n_iterations: int, draws, max_iterations: int = 5, store: bool = True) -> RunRef:
    """``TESS.sample`` (nfmc/tess.py:151-188): the chain state is the latent ``u`` (initialised with ``x0``), the recorded
    samples / moments are the data-space points ``x``; every iteration books ``(max_iterations + 1) n`` target calls."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    u = x0.clone()
    x = None
    for _ in range(n_iterations):
        x, u, mask = tess_step(u, flow, potential, max_iterations, draws)
        out.n_target_calls += (max_iterations + 1) * n                            # :174-178
        out.n_accepted += int(torch.sum(mask))
        out.n_attempted += n
        out.observe(x, store)                                                    # :179-180
    out.x = x
    out.trace["u"] = u
    return out.finish(store)


# ---------------------------------------------------------------------------------------------------------
# deterministic Langevin Monte Carlo (nfmc/dlmc.py)
# ---------------------------------------------------------------------------------------------------------
def run_dlmc(x0, target, nll, flow, n_iterations: int, draws, step_size: float = 0.05, latent_updates: bool = False,
             fit: Optional[Callable] = None, store: bool = True) -> RunRef:
    """``DLMC.sample`` (nfmc/dlmc.py:44-119).  ``fit(flow, x)`` stands for the per-iteration flow refit (:73-79); ``None``
    keeps the flow frozen (what the golden fixture pins -- the refit itself is torchflows' optimiser)."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    _, g = value_and_grad(nll, x0)
    x = (x0 - step_size * g).detach()                                             # :60-62
    out.n_target_calls += n
    out.n_grad_calls += n                                                          # :64-67
    for _ in range(n_iterations):
        if fit is not None:
            fit(flow, x)
        if latent_updates:                                                         # :81-85
            with torch.no_grad():
                z, _ = flow.bijection.forward(x)
            _, g = value_and_grad(target, x)
            z = z - step_size * (g - z)
            with torch.no_grad():
                x, _ = flow.bijection.inverse(z)
        else:                                                                      # :86-88
            _, g = value_and_grad(lambda v: target(v) + flow.log_prob(v), x)
            x = x - step_size * g
        out.n_target_calls += n
        out.n_grad_calls += n                                                      # :90-93
        x_tilde, _ = flow_sample_with_logq(flow, n, draws)                         # :94
        with torch.no_grad():
            log_alpha = mh_log_ratio(-target(x), -target(x_tilde), flow.log_prob(x), flow.log_prob(x_tilde))   # :95-100
        u = draws.uniform(n)
        mask = torch.log(u) < log_alpha                                            # :101-102
        x = x.detach().clone()
        x[mask] = x_tilde[mask]                                                    # :103
        out.observe(x, store)                                                      # :107-108
        out.n_target_calls += 2 * n
        out.n_accepted += int(torch.sum(mask))
        out.n_attempted += n                                                       # :109-113
    out.x = x
    return out.finish(store)


# ---------------------------------------------------------------------------------------------------------
# warm-up tuning (mcmc/base.py:142-161, tuning.py:15-41)
# ---------------------------------------------------------------------------------------------------------
class DualAveragingRef:
    def __init__(self, initial_step, target_rate=0.651, kappa=0.75, gamma=0.05, t0=10):
        self.t = t0
        self.err = 0.0
        self.log_avg = math.log(initial_step)
        self.mu = math.log(10 * initial_step)
        self.kappa, self.gamma, self.target_rate = kappa, gamma, target_rate

    def step_error(self, error) -> float:
        """``DualAveraging.step`` (tuning.py:25-38) with the acceptance-rate error the caller computed."""
        self.err += float(error)
        log_raw = self.mu - self.err / (math.sqrt(self.t) * self.gamma)
        eta = self.t ** -self.kappa
        self.log_avg = eta * log_raw + (1 - eta) * self.log_avg
        self.t += 1
        return math.exp(self.log_avg)

    def step(self, acc_rate: float) -> float:
        self.err += float(self.target_rate - acc_rate)
        log_raw = self.mu - self.err / (math.sqrt(self.t) * self.gamma)
        eta = self.t ** -self.kappa
        self.log_avg = eta * log_raw + (1 - eta) * self.log_avg
        self.t += 1
        return math.exp(self.log_avg)


def tune_inv_mass(imd, x, c=1e-3):
    return c * torch.var(x.flatten(1, -1), dim=0) + (1 - c) * imd


def run_tuned(x0, target, kind: str, tau: float, imd: torch.Tensor, n_steps: int, draws, n_leapfrog: int = 20,
              imd_adjustment: float = 1e-3, store: bool = True) -> RunRef:
    """``MCMCSampler.sample`` with ``params.tuning`` (mcmc/base.py:56-102): after every iteration
    ``MetropolisSampler.update_kernel`` (mcmc/base.py:142-161) moves the inverse-mass diagonal towards the across-chain
    variance and the step size by dual averaging (tuning.py:15-41).  ``kind``: ``'mala'`` or ``'hmc'``.  The run carries
    ``step_traj`` / ``imd_traj``: the kernel after each iteration."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    out.step_traj, out.imd_traj = [], []
    x = x0.clone().detach()
    imd = imd.clone()
    da = DualAveragingRef(tau)
    for _ in range(n_steps):
        if kind == "mala":
            x_prime, mask, calls, grads = mala_propose(x, target, tau, imd, draws, True, None)
        else:
            x_prime, mask, calls, grads = hmc_propose(x, target, tau, imd, n_leapfrog, draws, True, None)
        x = x.detach()
        x[mask] = x_prime[mask]                                                # :77
        out.n_target_calls += calls
        out.n_grad_calls += grads
        out.n_accepted += int(torch.sum(mask))
        out.n_attempted += n
        out.observe(x, store)
        if n > 1:
            imd = tune_inv_mass(imd, x, imd_adjustment)                        # :150-155
        error = da.target_rate - torch.mean(mask.float())                      # :157-158 (a float32 tensor, as there)
        tau = da.step_error(error)                                             # :159-160
        out.step_traj.append(tau)
        out.imd_traj.append(imd.clone())
    out.x = x
    return out.finish(store)


@torch.no_grad()
def run_adaptive_imh(x0, target, flow, n_iterations: int, draws, store: bool = True) -> RunRef:
    """``AdaptiveIMH.sample`` (nfmc/imh.py:102-181) with the refit switched off: both proposal densities are recomputed
    every iteration (:133-134), the counters book ``2 n`` GRADIENT calls (:146, the reference's own bookkeeping), and one
    extra scalar uniform decides whether a refit would happen (:151-153)."""
    n = x0.shape[0]
    out = RunRef(event_shape=tuple(x0.shape[1:]))
    x = x0.clone()
    for i in range(n_iterations):
        z = draws.normal(n, flow.event_shape)                                  # flow.sample(n, no_grad=True) :128
        x_prime, _ = flow.bijection.inverse(z)
        log_alpha = mh_log_ratio(-target(x), -target(x_prime), flow.log_prob(x), flow.log_prob(x_prime))   # :130-135
        log_u = draws.uniform(n).log().to(log_alpha)                           # :136
        mask = torch.less(log_u, log_alpha)
        x[mask] = x_prime[mask]                                                # :138
        x = x.detach()
        out.observe(x, store)                                                  # :144,150
        out.n_grad_calls += 2 * n                                              # :146
        out.n_accepted += int(torch.sum(mask))
        out.n_attempted += n
        draws.uniform_scalar()                                                 # u_prime :152
    out.x = x
    return out.finish(store)


# ---------------------------------------------------------------------------------------------------------
# CPU timing helper for bench.py's cpu_baseline / --impl reference legs
# ---------------------------------------------------------------------------------------------------------
def timed(fn, *args, **kwargs):
    t0 = time.perf_counter()
    out = fn(*args, **kwargs)
    return out, time.perf_counter() - t0
