"""Oracle (test infrastructure): the synthetic target potentials as differentiable torch callables.

The reference ships only ``sum(x**2)`` (/root/reference/test/util.py:4-5, README.md:45-46) and
``sum(x**2 / (2*100**2))`` (/root/reference/test/util.py:8-9); the other targets named by BASELINE.json's
configs are defined by this build and restated here so that the reference's
autograd-based samplers can run on exactly the functions the CUDA kernels hard-code.

Every class is ``[n, *event] -> [n]`` (negative log density up to a constant) with ``.event_shape``.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class PotentialRef(nn.Module):
    def __init__(self, event_shape):
        super().__init__()
        self.event_shape = tuple(int(s) for s in event_shape)
        self.n_dim = int(math.prod(self.event_shape))

    def _flat(self, x):
        return x.reshape(*x.shape[: x.ndim - len(self.event_shape)], self.n_dim)

    def forward(self, x):
        raise NotImplementedError


class DiagGaussianRef(PotentialRef):
    """U = 1/2 sum w_i (x_i - mu_i)^2   (``w = 2`` gives the README's ``sum(x**2)``)."""

    def __init__(self, event_shape, precision, mean=None):
        super().__init__(event_shape)
        w = torch.as_tensor(precision, dtype=torch.float32).reshape(-1)
        if w.numel() == 1:
            w = w.expand(self.n_dim).clone()
        self.register_buffer("w", w)
        mu = torch.zeros(self.n_dim) if mean is None else torch.as_tensor(mean, dtype=torch.float32).reshape(-1)
        self.register_buffer("mu", mu)

    def forward(self, x):
        c = self._flat(x) - self.mu
        return 0.5 * (self.w * c * c).sum(dim=-1)


def standard_gaussian_ref(event_shape):
    """G0: ``sum(x**2)`` (README.md:45-46)."""
    return DiagGaussianRef(event_shape, 2.0)


def ill_conditioned_sigmas(d: int) -> torch.Tensor:
    i = torch.arange(d, dtype=torch.float64)
    return (10.0 ** (-1.0 + 3.0 * (i / max(d - 1, 1)))).to(torch.float32)


def ill_conditioned_gaussian_ref(event_shape):
    """G1: sigma_i = 10^(-1 + 3 i/(d-1)), condition number 1e6."""
    d = int(math.prod(event_shape))
    s = ill_conditioned_sigmas(d)
    return DiagGaussianRef(event_shape, 1.0 / (s * s))


class FunnelRef(PotentialRef):
    """FN: U = x0^2/18 + (d-1)/2 x0 + 1/2 exp(-x0) sum_{i>=1} x_i^2."""

    def forward(self, x):
        xf = self._flat(x)
        x0 = xf[..., 0]
        s = (xf[..., 1:] ** 2).sum(dim=-1)
        return x0 * x0 / 18.0 + 0.5 * (self.n_dim - 1) * x0 + 0.5 * torch.exp(-x0) * s


class RosenbrockRef(PotentialRef):
    """RB: U = sum_{k<d/2} (x_k - 1)^2 + 10 (x_{k+d/2} - x_k^2)^2   (d even).

    Independent Rosenbrock pairs; pair k couples coordinates (k, k + d/2) -- the RealNVP half split -- which is
    the interleaved-pair form of SURVEY.md section 8d up to a fixed permutation of the coordinates."""

    def __init__(self, event_shape, scale: float = 10.0):
        super().__init__(event_shape)
        assert self.n_dim % 2 == 0
        self.scale = float(scale)

    def forward(self, x):
        xf = self._flat(x)
        h = self.n_dim // 2
        a, b = xf[..., :h], xf[..., h:]
        return ((a - 1.0) ** 2 + self.scale * (b - a * a) ** 2).sum(dim=-1)


class MixtureRef(PotentialRef):
    """GM: 4 equal-weight isotropic unit Gaussians at (+-a, +-a) on the first two axes,
    U = -logsumexp_k(-1/2 |x - mu_k|^2)."""

    def __init__(self, event_shape, offset: float = 3.0):
        super().__init__(event_shape)
        assert self.n_dim >= 2
        self.offset = float(offset)
        mu = torch.zeros(4, self.n_dim)
        mu[:, 0] = torch.tensor([1.0, 1.0, -1.0, -1.0]) * self.offset
        mu[:, 1] = torch.tensor([1.0, -1.0, 1.0, -1.0]) * self.offset
        self.register_buffer("mu", mu)

    def forward(self, x):
        xf = self._flat(x)
        sq = ((xf[..., None, :] - self.mu) ** 2).sum(dim=-1)
        return -torch.logsumexp(-0.5 * sq, dim=-1)


def make_potential_ref(name: str, event_shape):
    name = name.lower()
    if name in ("g0", "standard_gaussian", "gaussian"):
        return standard_gaussian_ref(event_shape)
    if name in ("g1", "ill_conditioned_gaussian"):
        return ill_conditioned_gaussian_ref(event_shape)
    if name in ("fn", "funnel"):
        return FunnelRef(event_shape)
    if name in ("rb", "rosenbrock"):
        return RosenbrockRef(event_shape)
    if name in ("gm", "mixture"):
        return MixtureRef(event_shape)
    raise ValueError(name)
