"""Built-in analytic target potentials (negative log densities) with closed-form gradients on the device.

The reference accepts any Python callable and differentiates it with autograd
(/root/reference/nfmc/algorithms/sampling/mcmc/langevin.py:66-68, mcmc/hmc.py:40-48).  The fast B200 path fuses
U and grad U into the sampler kernels, which needs one of the potentials below (they mirror
``potentials.base.Potential``: an object with ``event_shape`` that maps ``[n, *event] -> [n]``).  Calling a
potential evaluates it with the CUDA kernel ``nfmc_potential_eval``; there is no CPU path.  Any other callable becomes a
``CallablePotential`` (bottom of this file): autograd supplies U / grad U on the device and the ``nfmc_ext_*`` kernels do
the sampler arithmetic around it.

========================  =====================================================================================
``StandardGaussian``      ``sum(x**2)`` -- the README / test target (README.md:45-46, test/util.py:4-5)
``DiagonalGaussian``      ``1/2 sum w_i (x_i - mu_i)^2``  (``test/util.py:8-9`` is ``w = 1e-4``)
``IllConditionedGaussian````sigma_i = 10^(-1 + 3 i/(d-1))`` (condition number 1e6; config C2)
``Funnel``                Neal's funnel, ``x0 ~ N(0, s^2)``, ``x_i | x0 ~ N(0, e^{x0})`` (config C3)
``Rosenbrock``            ``sum_{k<d/2} (x_k - 1)^2 + c (x_{k+d/2} - x_k^2)^2`` (config C4; pairs are (k, k+d/2))
``GaussianMixture4``      4 unit Gaussians at ``(+-a, +-a, 0, ...)`` (config C5)
========================  =====================================================================================
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import torch

from . import _native as N


class Potential:
    kind: int = -1

    def __init__(self, event_shape):
        if isinstance(event_shape, int):
            event_shape = (event_shape,)
        self.event_shape = tuple(int(s) for s in event_shape)
        self.n_dim = int(math.prod(self.event_shape))
        if not 1 <= self.n_dim <= N.MAX_DIM:
            raise ValueError(f"event size {self.n_dim} outside [1, {N.MAX_DIM}]")
        self._dev_params = {}

    # -- per-dimension parameters (host float tensor, flattened) and scalars --------------------------------
    def host_params(self) -> Optional[torch.Tensor]:
        return None

    def scalars(self) -> Sequence[float]:
        return (0.0, 0.0, 0.0, 0.0)

    def descriptor(self, device: torch.device) -> Tuple[N.PotentialDesc, Optional[torch.Tensor]]:
        """C descriptor + the device tensor that keeps its parameters alive."""
        hp = self.host_params()
        dp = None
        if hp is not None:
            key = str(device)
            if key not in self._dev_params:
                self._dev_params[key] = N.dev_f32(hp, device)
            dp = self._dev_params[key]
        s = list(self.scalars()) + [0.0] * 4
        desc = N.PotentialDesc(self.kind, self.n_dim, None if dp is None else dp.data_ptr(), (C.c_float * 4)(*s[:4]))
        return desc, dp

    def value_and_grad(self, x: torch.Tensor, need_grad: bool = True):
        """U(x) [n] and grad U(x) [n, *event] on x's CUDA device."""
        dev = N.require_cuda(x.device if x.is_cuda else None)
        xd = N.dev_f32(x, dev).reshape(-1, self.n_dim)
        n = xd.shape[0]
        u = torch.empty(n, device=dev, dtype=torch.float32)
        g = torch.empty_like(xd) if need_grad else None
        desc, keep = self.descriptor(dev)
        with torch.cuda.device(dev):           # the C entry point launches on the current device
            N.check(N.lib().nfmc_potential_eval(C.byref(desc), N.ptr(xd), N.ptr(u), N.ptr(g), n, N.stream_ptr(dev)))
        return u, (None if g is None else g.reshape(x.shape))

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.value_and_grad(x, need_grad=False)[0]

    def log_prob_fn(self):
        """``lambda v: -U(v)``, differentiable in torch with the value / gradient supplied by the CUDA kernel
        (what the reference passes to ``flow.variational_fit``: imh.py:68, neutra.py:85)."""
        from .flow_train import target_log_prob_fn
        return target_log_prob_fn(self)


class IsotropicGaussian(Potential):
    """U = 1/2 w sum x_i^2."""
    kind = N.POT_ISO_GAUSSIAN

    def __init__(self, event_shape, precision: float = 1.0):
        super().__init__(event_shape)
        self.precision = float(precision)

    def scalars(self):
        return (self.precision, 0.0, 0.0, 0.0)


class StandardGaussian(IsotropicGaussian):
    """``sum(x**2)`` -- N(0, I/2), the reference's README and test target."""

    def __init__(self, event_shape):
        super().__init__(event_shape, precision=2.0)


class DiagonalGaussian(Potential):
    """U = 1/2 sum w_i (x_i - mu_i)^2 with per-dimension precision ``w`` and mean ``mu``."""
    kind = N.POT_DIAG_GAUSSIAN

    def __init__(self, event_shape, precision, mean=None):
        super().__init__(event_shape)
        w = torch.as_tensor(precision, dtype=torch.float32).reshape(-1)
        if w.numel() == 1:
            w = w.expand(self.n_dim).clone()
        if w.numel() != self.n_dim:
            raise ValueError("precision must have one entry per dimension")
        mu = torch.zeros(self.n_dim) if mean is None else torch.as_tensor(mean, dtype=torch.float32).reshape(-1)
        self.precision, self.mean = w, mu

    def host_params(self):
        return torch.stack([self.precision, self.mean], dim=1).reshape(-1)  # float2[d] {w_i, mu_i}


class IllConditionedGaussian(DiagonalGaussian):
    def __init__(self, event_shape, log10_min: float = -1.0, log10_range: float = 3.0):
        d = int(math.prod((event_shape,) if isinstance(event_shape, int) else event_shape))
        i = torch.arange(d, dtype=torch.float64)
        sigma = (10.0 ** (log10_min + log10_range * (i / max(d - 1, 1)))).to(torch.float32)
        super().__init__(event_shape, 1.0 / (sigma * sigma))
        self.sigma = sigma


class Funnel(Potential):
    kind = N.POT_FUNNEL

    def __init__(self, event_shape, scale: float = 3.0):
        super().__init__(event_shape)
        if self.n_dim < 2:
            raise ValueError("Funnel needs at least 2 dimensions")
        self.scale = float(scale)

    def scalars(self):
        return (self.scale, 0.0, 0.0, 0.0)


class Rosenbrock(Potential):
    kind = N.POT_ROSENBROCK

    def __init__(self, event_shape, scale: float = 10.0):
        super().__init__(event_shape)
        if self.n_dim % 2:
            raise ValueError("Rosenbrock needs an even number of dimensions")
        self.scale = float(scale)

    def scalars(self):
        return (self.scale, 0.0, 0.0, 0.0)


class GaussianMixture4(Potential):
    kind = N.POT_MIXTURE4

    def __init__(self, event_shape, offset: float = 3.0):
        super().__init__(event_shape)
        if self.n_dim < 2:
            raise ValueError("GaussianMixture4 needs at least 2 dimensions")
        self.offset = float(offset)

    def scalars(self):
        return (self.offset, 0.0, 0.0, 0.0)


def make_potential(name: str, event_shape) -> Potential:
    name = name.lower()
    table = {
        "g0": StandardGaussian, "standard_gaussian": StandardGaussian, "gaussian": StandardGaussian,
        "g1": IllConditionedGaussian, "ill_conditioned_gaussian": IllConditionedGaussian,
        "fn": Funnel, "funnel": Funnel,
        "rb": Rosenbrock, "rosenbrock": Rosenbrock,
        "gm": GaussianMixture4, "mixture": GaussianMixture4,
    }
    if name not in table:
        raise ValueError(f"unknown potential {name!r}")
    return table[name](event_shape)


class CallablePotential(Potential):
    """Any Python callable ``[n, *event] -> [n]`` -- the reference's target contract (sample.py:34-36; the README's
    ``lambda x: torch.sum(x ** 2, dim=1)``).  It cannot be fused into the step kernels, so the samplers take their
    *external-target* path for it: U and grad U are evaluated on the device by calling the function under torch autograd
    (exactly what the reference does: mcmc/langevin.py:66-68, mcmc/hmc.py:40-48), and the ``nfmc_ext_*`` kernels
    (csrc/ext_kernels.cu) do the rest of every step.  The function must be written in torch operations that run on CUDA
    tensors."""
    kind = -1
    external = True

    def __init__(self, fn, event_shape):
        super().__init__(event_shape)
        if not callable(fn):
            raise TypeError("target must be callable")
        self.fn = fn

    def descriptor(self, device):
        raise N.NativeError("a callable target has no fused-kernel descriptor; samplers use the external-target path for it")

    def _rows(self, x: torch.Tensor) -> torch.Tensor:
        return x.reshape(-1, *self.event_shape)

    def _check(self, u: torch.Tensor, n: int) -> torch.Tensor:
        if not torch.is_tensor(u) or u.numel() != n:
            raise ValueError(f"target must map [n, *event_shape] to [n]; got {tuple(getattr(u, 'shape', ()))} for n = {n}")
        return u.reshape(n)

    def value(self, x: torch.Tensor) -> torch.Tensor:
        """U(x) as a contiguous fp32 [n] tensor on x's device (no graph)."""
        xr = self._rows(x.detach())
        with torch.no_grad():
            u = self._check(self.fn(xr), xr.shape[0])
        return u.to(torch.float32).contiguous()

    def value_and_grad(self, x: torch.Tensor, need_grad: bool = True):
        if not need_grad:
            return self.value(x), None
        with torch.enable_grad():
            xr = self._rows(x.detach()).clone().requires_grad_(True)
            u = self._check(self.fn(xr), xr.shape[0])
            (g,) = torch.autograd.grad(u.sum(), xr)                      # langevin.py:68, hmc.py:43
        return u.detach().to(torch.float32).contiguous(), g.detach().to(torch.float32).reshape(x.shape).contiguous()

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.fn(x)

    def log_prob_fn(self):
        fn = lambda v: -self.fn(v)     # noqa: E731
        fn.potential = self            # variational_fit: U / grad U by autograd, flow sweep by the native wide trainer
        return fn


Potential.external = False


def resolve_target(target, event_shape) -> Potential:
    """A built-in analytic potential (fused into the kernels), its name, or any callable (external-target path)."""
    if isinstance(target, Potential):
        return target
    if isinstance(target, str):
        return make_potential(target, event_shape)
    if callable(target):
        return CallablePotential(target, event_shape)
    raise TypeError(f"target must be a Potential, the name of one, or a callable [n, *event] -> [n]; got {type(target)!r}")
