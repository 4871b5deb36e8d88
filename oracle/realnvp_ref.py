"""Oracle (test infrastructure): RealNVP + Flow restated in plain PyTorch, fp32, CPU.

PARITY UNPINNED for this file: the real arithmetic lives in ``torchflows`` (unpinned, un-vendored, absent;
see ``oracle/__init__.py``).  What is restated here is the published RealNVP construction, shaped by the
surface the reference actually touches:

* ``Flow(RealNVP(event_shape))``                       -- /root/reference/nfmc/algorithms/sampling/base.py:26
* ``RealNVP(event_shape, **kwargs)`` + ``Flow(bij)``    -- /root/reference/nfmc/util.py:280-281,379
* ``flow.sample(n, no_grad=, return_log_prob=)``        -- jump.py:205, imh.py:74,128,221
* ``flow.log_prob(x)``                                  -- jump.py:218, imh.py:133-134,214
* ``flow.bijection.inverse(z) -> (x, log_det)``         -- neutra.py:60,122
* ``flow.bijection.layers`` (sized, grows with n_layers)-- test/test_flow_kwargs.py:18,28,30
* ``conditioner_kwargs={'n_layers','n_hidden'}``        -- test/test_flow_kwargs.py:49
* ``flow.fit`` / ``flow.variational_fit`` / ``state_dict`` -- jump.py:130-151,201; imh.py:67,173; neutra.py:84

Conventions (this file IS the specification the CUDA kernels are tested against):

* ``forward``: data x -> latent z, returns ``(z, log|det dz/dx|)``; ``inverse``: z -> x, returns
  ``(x, log|det dx/dz|)``.
* affine map with unconstrained pair ``(u_a, u_b)``: ``alpha = exp(log(1-m) + u_a/2) + m`` with ``m = 1e-3``,
  ``beta = u_b/2``; forward ``alpha*x + beta`` with log-det ``sum log alpha``.  Zero parameters = identity.
* layer list: ``[Affine] + n_layers x [Reverse, Coupling, ActNorm] + [Affine, ActNorm]``
  (``len(layers) == 3*n_layers + 3``).
* coupling: source = first ``d//2`` flattened dims, target = the rest; conditioner = MLP with ``n_layers``
  linear layers (default 2), hidden width ``max(int(3*log10(d//2)), 4)`` unless given, tanh between;
  output viewed as ``[n, d_target, 2]`` -> ``(u_a, u_b)`` per target dim.
"""
from __future__ import annotations

import math
import time
from copy import deepcopy
from typing import Optional, Tuple, Union

import torch
import torch.nn as nn

MIN_SCALE = 1e-3
_LOG_ONE_MINUS_M = math.log(1.0 - MIN_SCALE)


def get_batch_shape(x: torch.Tensor, event_shape) -> torch.Size:
    return x.shape[: x.ndim - len(event_shape)]


def sum_except_batch(x: torch.Tensor, event_shape) -> torch.Tensor:
    k = len(event_shape)
    if k == 0:
        return x
    return x.sum(dim=tuple(range(x.ndim - k, x.ndim)))


def affine_coefficients(u_a: torch.Tensor, u_b: torch.Tensor):
    """(u_a, u_b) -> (alpha, log_alpha, beta)."""
    alpha = torch.exp(_LOG_ONE_MINUS_M + u_a / 2) + MIN_SCALE
    return alpha, torch.log(alpha), u_b / 2


class BijectionRef(nn.Module):
    def __init__(self, event_shape):
        super().__init__()
        self.event_shape = tuple(int(s) for s in event_shape)
        self.n_dim = int(math.prod(self.event_shape))

    def _flat(self, x):
        batch = get_batch_shape(x, self.event_shape)
        return x.reshape(*batch, self.n_dim), batch

    def forward(self, x, context=None) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def inverse(self, z, context=None) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError


class AffineElementwiseRef(BijectionRef):
    """Per-dimension affine map; ``value[:, 0] = u_a``, ``value[:, 1] = u_b``."""

    def __init__(self, event_shape):
        super().__init__(event_shape)
        self.value = nn.Parameter(torch.zeros(self.n_dim, 2))

    def forward(self, x, context=None):
        xf, batch = self._flat(x)
        alpha, log_alpha, beta = affine_coefficients(self.value[:, 0], self.value[:, 1])
        z = alpha * xf + beta
        return z.reshape(x.shape), log_alpha.sum().expand(batch)

    def inverse(self, z, context=None):
        zf, batch = self._flat(z)
        alpha, log_alpha, beta = affine_coefficients(self.value[:, 0], self.value[:, 1])
        x = (zf - beta) / alpha
        return x.reshape(z.shape), (-log_alpha.sum()).expand(batch)


class ActNormRef(AffineElementwiseRef):
    """Same map; parameters are data-initialised on the first forward pass in training mode."""

    def __init__(self, event_shape):
        super().__init__(event_shape)
        self.register_buffer("initialised", torch.tensor(False))

    @torch.no_grad()
    def _data_init(self, xf):
        flat = xf.reshape(-1, self.n_dim)
        if flat.shape[0] < 2:
            return
        std = flat.std(dim=0).clamp_min(1e-2)
        mean = flat.mean(dim=0)
        alpha = (1.0 / std).clamp_min(2 * MIN_SCALE)
        self.value[:, 0] = 2 * (torch.log(alpha - MIN_SCALE) - _LOG_ONE_MINUS_M)
        self.value[:, 1] = 2 * (-mean * alpha)
        self.initialised.fill_(True)

    def forward(self, x, context=None):
        if self.training and not bool(self.initialised):
            self._data_init(self._flat(x)[0])
        return super().forward(x, context)


class ReverseRef(BijectionRef):
    def forward(self, x, context=None):
        xf, batch = self._flat(x)
        return xf.flip(-1).reshape(x.shape), torch.zeros(batch, dtype=x.dtype, device=x.device)

    inverse = forward


def default_hidden(n_source: int) -> int:
    return max(int(3 * math.log10(n_source)), 4)


class AffineCouplingRef(BijectionRef):
    def __init__(self, event_shape, conditioner_kwargs: Optional[dict] = None, **_ignored):
        super().__init__(event_shape)
        ck = dict(conditioner_kwargs or {})
        self.n_source = self.n_dim // 2
        self.n_target = self.n_dim - self.n_source
        n_lin = int(ck.get("n_layers", 2))
        hidden = ck.get("n_hidden", None)
        if hidden is None:
            hidden = default_hidden(self.n_source)
        hidden = int(hidden)
        mods = []
        if n_lin == 1:
            mods.append(nn.Linear(self.n_source, 2 * self.n_target))
        else:
            mods += [nn.Linear(self.n_source, hidden), nn.Tanh()]
            for _ in range(n_lin - 2):
                mods += [nn.Linear(hidden, hidden), nn.Tanh()]
            mods.append(nn.Linear(hidden, 2 * self.n_target))
        self.net = nn.Sequential(*mods)
        self.n_linear = n_lin
        self.n_hidden = hidden
        # start at the identity map: last layer zero
        with torch.no_grad():
            self.net[-1].weight.zero_()
            self.net[-1].bias.zero_()

    def _coefficients(self, a):
        out = self.net(a).reshape(*a.shape[:-1], self.n_target, 2)
        return affine_coefficients(out[..., 0], out[..., 1])

    def forward(self, x, context=None):
        xf, batch = self._flat(x)
        a, b = xf[..., : self.n_source], xf[..., self.n_source:]
        alpha, log_alpha, beta = self._coefficients(a)
        z = torch.cat([a, alpha * b + beta], dim=-1)
        return z.reshape(x.shape), log_alpha.sum(dim=-1)

    def inverse(self, z, context=None):
        zf, batch = self._flat(z)
        a, b = zf[..., : self.n_source], zf[..., self.n_source:]
        alpha, log_alpha, beta = self._coefficients(a)
        x = torch.cat([a, (b - beta) / alpha], dim=-1)
        return x.reshape(z.shape), -log_alpha.sum(dim=-1)


class CompositionRef(BijectionRef):
    def __init__(self, event_shape, layers):
        super().__init__(event_shape)
        self.layers = nn.ModuleList(layers)

    def forward(self, x, context=None):
        log_det = torch.zeros(get_batch_shape(x, self.event_shape), dtype=x.dtype, device=x.device)
        for layer in self.layers:
            x, ld = layer.forward(x, context)
            log_det = log_det + ld
        return x, log_det

    def inverse(self, z, context=None):
        log_det = torch.zeros(get_batch_shape(z, self.event_shape), dtype=z.dtype, device=z.device)
        for layer in reversed(self.layers):
            z, ld = layer.inverse(z, context)
            log_det = log_det + ld
        return z, log_det


class RealNVPRef(CompositionRef):
    def __init__(self, event_shape, n_layers: int = 2, edge_list=None, **kwargs):
        if isinstance(event_shape, int):
            event_shape = (event_shape,)
        if edge_list is not None:
            raise NotImplementedError("edge_list couplings are outside the hot path")
        layers = [AffineElementwiseRef(event_shape)]
        for _ in range(int(n_layers)):
            layers += [ReverseRef(event_shape), AffineCouplingRef(event_shape, **kwargs), ActNormRef(event_shape)]
        layers += [AffineElementwiseRef(event_shape), ActNormRef(event_shape)]
        super().__init__(event_shape, layers)
        self.n_coupling = int(n_layers)


class FlowRef(nn.Module):
    """Standard-normal base + bijection (the torchflows ``Flow`` surface nfmc uses)."""

    def __init__(self, bijection: BijectionRef):
        super().__init__()
        self.bijection = bijection
        self.register_buffer("_device_probe", torch.zeros(()))

    @property
    def event_shape(self):
        return self.bijection.event_shape

    def get_device(self):
        return self._device_probe.device

    def base_log_prob(self, z):
        zf = z.reshape(*get_batch_shape(z, self.event_shape), -1)
        return (-0.5 * zf.square()).sum(dim=-1) - 0.5 * zf.shape[-1] * math.log(2 * math.pi)

    def base_sample(self, shape):
        return torch.randn(size=(*shape, *self.event_shape)).to(self.get_device())

    def log_prob(self, x, context=None):
        z, log_det = self.bijection.forward(x.to(self.get_device()), context)
        return self.base_log_prob(z) + log_det

    def sample(self, sample_shape, context=None, no_grad: bool = False, return_log_prob: bool = False):
        if isinstance(sample_shape, int):
            sample_shape = (sample_shape,)
        z = self.base_sample(tuple(sample_shape))
        if no_grad:
            with torch.no_grad():
                x, log_det = self.bijection.inverse(z, context)
        else:
            x, log_det = self.bijection.inverse(z, context)
        if return_log_prob:
            return x, self.base_log_prob(z) - log_det
        return x

    # -- training (only reached through warm-up / fit_nf / adaptive IMH; kept simple) ---------------------
    def fit(self, x_train, n_epochs: int = 500, lr: float = 0.05, batch_size=None, shuffle: bool = True,
            show_progress: bool = False, w_train=None, context_train=None, x_val=None, w_val=None,
            context_val=None, keep_best_weights: bool = True, early_stopping: bool = False,
            early_stopping_threshold: int = 50, time_limit_seconds=None, **_ignored):
        x_train = x_train.detach().to(self.get_device())
        n = len(x_train)
        if batch_size == "adaptive":
            batch_size = max(32, min(1024, n // 10 if n >= 320 else n))
        if batch_size is None:
            batch_size = n
        opt = torch.optim.AdamW(self.parameters(), lr=lr)
        best, best_state, since_best = math.inf, None, 0
        t0 = time.time()
        self.train()
        for _ in range(n_epochs):
            if time_limit_seconds is not None and time.time() - t0 > time_limit_seconds:
                break
            perm = torch.randperm(n) if shuffle else torch.arange(n)
            for i in range(0, n, batch_size):
                xb = x_train[perm[i:i + batch_size]]
                opt.zero_grad()
                loss = -self.log_prob(xb).mean()
                if not torch.isfinite(loss):
                    raise ValueError("Flow training diverged")
                loss.backward()
                opt.step()
            with torch.no_grad():
                ref = x_val if x_val is not None else x_train
                score = float(-self.log_prob(ref.to(self.get_device())).mean())
            if score < best:
                best, since_best = score, 0
                if keep_best_weights:
                    best_state = deepcopy(self.state_dict())
            else:
                since_best += 1
                if early_stopping and since_best >= early_stopping_threshold:
                    break
        if keep_best_weights and best_state is not None:
            self.load_state_dict(best_state)
        self.eval()

    def variational_fit(self, target_log_prob, n_epochs: int = 500, lr: float = 0.05, n_samples: int = 1,
                        early_stopping: bool = False, early_stopping_threshold: int = 50,
                        keep_best_weights: bool = True, show_progress: bool = False,
                        check_for_divergences: bool = False, time_limit_seconds=None, **_ignored):
        opt = torch.optim.AdamW(self.parameters(), lr=lr)
        best, best_state, since_best = math.inf, None, 0
        t0 = time.time()
        self.train()
        for _ in range(n_epochs):
            if time_limit_seconds is not None and time.time() - t0 > time_limit_seconds:
                break
            opt.zero_grad()
            x, log_q = self.sample(n_samples, return_log_prob=True)
            loss = (log_q - target_log_prob(x)).mean()
            if not torch.isfinite(loss):
                if check_for_divergences:
                    break
                raise ValueError("Flow training diverged")
            loss.backward()
            opt.step()
            if float(loss) < best:
                best, since_best = float(loss), 0
                if keep_best_weights:
                    best_state = deepcopy(self.state_dict())
            else:
                since_best += 1
                if early_stopping and since_best >= early_stopping_threshold:
                    break
        if keep_best_weights and best_state is not None:
            self.load_state_dict(best_state)
        self.eval()


def make_flow(event_shape, n_layers: int = 2, conditioner_kwargs: Optional[dict] = None,
              perturb: float = 0.0, seed: Optional[int] = None) -> FlowRef:
    """Frozen ``eval()`` flow; ``perturb`` adds ``perturb*randn`` to every parameter so it is not the identity
    (the synthetic-input recipe of SURVEY.md section 8d)."""
    if seed is not None:
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
    kwargs = {} if conditioner_kwargs is None else {"conditioner_kwargs": conditioner_kwargs}
    flow = FlowRef(RealNVPRef(event_shape, n_layers=n_layers, **kwargs))
    if perturb:
        with torch.no_grad():
            for p in flow.parameters():
                p.add_(perturb * torch.randn_like(p))
        for m in flow.modules():
            if isinstance(m, ActNormRef):
                m.initialised.fill_(True)
    if seed is not None:
        torch.random.set_rng_state(gen_state)
    return flow.eval()
