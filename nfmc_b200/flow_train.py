"""Flow training on the device: ``Flow.fit`` (maximum likelihood) and ``Flow.variational_fit`` (reverse KL).

Status ("next" row): functional, **library-backed**.  The sampling hot path is the
hand-written kernels; training needs gradients with respect to the *parameters* (wgrad), which the kernels do not
provide yet, so the optimisation step here runs a differentiable torch restatement of the same RealNVP arithmetic on
the GPU (autograd + AdamW).  What is native already: the target's value / gradient inside ``variational_fit`` come
from ``nfmc_potential_eval`` through a custom autograd function, and after every fit the packed blobs are rebuilt so
the samplers keep using the CUDA kernels.  Multi-GPU: gradients are all-reduced (NCCL) once per optimiser step.

Reference call sites: ``flow.fit`` -- /root/reference/nfmc/algorithms/sampling/nfmc/jump.py:139-151,201 and
nfmc/imh.py:171-175; ``flow.variational_fit`` -- nfmc/imh.py:67-72 and nfmc/neutra.py:84-91;
``train_val_split`` -- sampling/tuning.py:44-65.  The optimiser settings follow the kwargs the reference passes
(``lr=0.05``, ``early_stopping``, ``early_stopping_threshold``, ``keep_best_weights``, ``batch_size='adaptive'``,
``time_limit_seconds``); torchflows' own defaults are unpinned (the package is absent).
"""
from __future__ import annotations

import math
import time
from copy import deepcopy
from typing import Callable, Tuple

import torch
import torch.distributed as dist

MIN_SCALE = 1e-3
_LOG_ONE_MINUS_M = math.log(1.0 - MIN_SCALE)


def _affine(u_a, u_b):
    alpha = torch.exp(_LOG_ONE_MINUS_M + u_a / 2) + MIN_SCALE
    return alpha, torch.log(alpha), u_b / 2


def _actnorm_init(layer, h):
    """Data-dependent initialisation on the first training pass (zero mean / unit scale of that batch)."""
    with torch.no_grad():
        if h.shape[0] < 2:
            return
        std = h.std(dim=0).clamp_min(1e-2)
        mean = h.mean(dim=0)
        alpha = (1.0 / std).clamp_min(2 * MIN_SCALE)
        layer.value[:, 0] = 2 * (torch.log(alpha - MIN_SCALE) - _LOG_ONE_MINUS_M)
        layer.value[:, 1] = 2 * (-mean * alpha)
        layer.initialised.fill_(True)


def forward_autograd(bij, x: torch.Tensor, training: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """x -> z and log|det dz/dx|, differentiable with respect to the parameters and x (same arithmetic as the kernels)."""
    from .flow import ActNorm, AffineCoupling, ElementwiseAffine, ReversePermutation
    h = x.reshape(x.shape[0], -1)
    ld = torch.zeros(h.shape[0], device=h.device, dtype=h.dtype)
    for layer in bij.layers:
        if isinstance(layer, ReversePermutation):
            h = h.flip(-1)
        elif isinstance(layer, AffineCoupling):
            a, b = h[:, : layer.n_source], h[:, layer.n_source:]
            out = layer.net(a).reshape(h.shape[0], layer.n_target, 2)
            alpha, log_alpha, beta = _affine(out[..., 0], out[..., 1])
            h = torch.cat([a, alpha * b + beta], dim=1)
            ld = ld + log_alpha.sum(dim=1)
        elif isinstance(layer, ElementwiseAffine):
            if isinstance(layer, ActNorm) and training and not bool(layer.initialised):
                _actnorm_init(layer, h.detach())
            alpha, log_alpha, beta = _affine(layer.value[:, 0], layer.value[:, 1])
            h = alpha * h + beta
            ld = ld + log_alpha.sum()
    return h, ld


def inverse_autograd(bij, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    from .flow import AffineCoupling, ElementwiseAffine, ReversePermutation
    h = z.reshape(z.shape[0], -1)
    ld = torch.zeros(h.shape[0], device=h.device, dtype=h.dtype)
    for layer in reversed(list(bij.layers)):
        if isinstance(layer, ReversePermutation):
            h = h.flip(-1)
        elif isinstance(layer, AffineCoupling):
            a, b = h[:, : layer.n_source], h[:, layer.n_source:]
            out = layer.net(a).reshape(h.shape[0], layer.n_target, 2)
            alpha, log_alpha, beta = _affine(out[..., 0], out[..., 1])
            h = torch.cat([a, (b - beta) / alpha], dim=1)
            ld = ld - log_alpha.sum(dim=1)
        elif isinstance(layer, ElementwiseAffine):
            alpha, log_alpha, beta = _affine(layer.value[:, 0], layer.value[:, 1])
            h = (h - beta) / alpha
            ld = ld - log_alpha.sum()
    return h, ld


def log_prob_autograd(flow, x, training=False):
    z, ld = forward_autograd(flow.bijection, x, training)
    return (-0.5 * z.square()).sum(dim=1) - 0.5 * z.shape[1] * math.log(2 * math.pi) + ld


class _PotentialFn(torch.autograd.Function):
    """-U(x) with the value and gradient taken from the CUDA kernel (nfmc_potential_eval)."""

    @staticmethod
    def forward(ctx, x, potential):
        u, g = potential.value_and_grad(x.detach(), need_grad=True)
        ctx.save_for_backward(g)
        return -u

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return -grad_out[:, None] * g.reshape(g.shape[0], -1), None


def target_log_prob_fn(potential) -> Callable:
    """``lambda v: -target(v)`` of the reference (imh.py:68, neutra.py:85), differentiable through the native kernel."""
    return lambda v: _PotentialFn.apply(v.reshape(v.shape[0], -1), potential)


def _sync_grads(params):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
        dist.all_reduce(flat)
        flat /= dist.get_world_size()
        off = 0
        for p in params:
            if p.grad is not None:
                n = p.grad.numel()
                p.grad.copy_(flat[off:off + n].view_as(p.grad))
                off += n


def train_val_split(x: torch.Tensor, train_pct: float, max_train_size: int, max_val_size: int, shuffle: bool = True):
    """Reference: sampling/tuning.py:44-65 -- flatten (iteration, chain), shuffle, split, cap."""
    flat = x.flatten(0, 1) if x.ndim >= 3 else x
    if shuffle:
        flat = flat[torch.randperm(len(flat), device=flat.device)]
    n_train = int(train_pct * len(flat))
    return flat[:n_train][:max_train_size], flat[n_train:][:max_val_size]


def fit(flow, x_train, n_epochs: int = 500, lr: float = 0.05, batch_size=None, shuffle: bool = True,
        show_progress: bool = False, x_val=None, keep_best_weights: bool = True, early_stopping: bool = False,
        early_stopping_threshold: int = 50, time_limit_seconds=None, **_ignored):
    dev = flow._compute_device()
    flow.to(dev)                       # parameters live where the optimiser runs; the packed blobs are rebuilt lazily
    x_train = x_train.detach().to(dev, torch.float32).reshape(len(x_train), -1)
    if x_val is not None:
        x_val = x_val.detach().to(dev, torch.float32).reshape(len(x_val), -1)
    n = len(x_train)
    if n == 0:
        return
    if batch_size == "adaptive":
        batch_size = max(32, min(1024, n // 10 if n >= 320 else n))
    if batch_size is None:
        batch_size = n
    params = [p for p in flow.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=lr)
    best, best_state, since_best = math.inf, None, 0
    t0 = time.time()
    flow.train()
    try:
        with torch.enable_grad():
            for _ in range(n_epochs):
                if time_limit_seconds is not None and time.time() - t0 > time_limit_seconds:
                    break
                perm = torch.randperm(n, device=dev) if shuffle else torch.arange(n, device=dev)
                for i in range(0, n, batch_size):
                    opt.zero_grad(set_to_none=True)
                    loss = -log_prob_autograd(flow, x_train[perm[i:i + batch_size]], training=True).mean()
                    if not torch.isfinite(loss):
                        raise ValueError("Flow training diverged")          # the reference rolls back on ValueError
                    loss.backward()
                    _sync_grads(params)
                    opt.step()
                with torch.no_grad():
                    ref = x_val if x_val is not None and len(x_val) else x_train
                    score = float(-log_prob_autograd(flow, ref).mean())
                if score < best:
                    best, since_best = score, 0
                    if keep_best_weights:
                        best_state = deepcopy(flow.state_dict())
                else:
                    since_best += 1
                    if early_stopping and since_best >= early_stopping_threshold:
                        break
        if keep_best_weights and best_state is not None:
            flow.load_state_dict(best_state)
    finally:
        flow.eval()


def variational_fit(flow, target_log_prob: Callable, n_epochs: int = 500, lr: float = 0.05, n_samples: int = 1,
                    early_stopping: bool = False, early_stopping_threshold: int = 50, keep_best_weights: bool = True,
                    show_progress: bool = False, check_for_divergences: bool = False, time_limit_seconds=None, **_ignored):
    dev = flow._compute_device()
    flow.to(dev)
    d = flow.bijection.n_dim
    params = [p for p in flow.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=lr)
    best, best_state, since_best = math.inf, None, 0
    t0 = time.time()
    flow.train()
    try:
        with torch.enable_grad():
            for _ in range(n_epochs):
                if time_limit_seconds is not None and time.time() - t0 > time_limit_seconds:
                    break
                opt.zero_grad(set_to_none=True)
                z = torch.randn(n_samples, d, device=dev)
                x, ld = inverse_autograd(flow.bijection, z)
                log_q = (-0.5 * z.square()).sum(dim=1) - 0.5 * d * math.log(2 * math.pi) - ld
                loss = (log_q - target_log_prob(x)).mean()
                if not torch.isfinite(loss):
                    if check_for_divergences:
                        break
                    raise ValueError("Flow training diverged")
                loss.backward()
                _sync_grads(params)
                opt.step()
                val = float(loss.detach())
                if val < best:
                    best, since_best = val, 0
                    if keep_best_weights:
                        best_state = deepcopy(flow.state_dict())
                else:
                    since_best += 1
                    if early_stopping and since_best >= early_stopping_threshold:
                        break
        if keep_best_weights and best_state is not None:
            flow.load_state_dict(best_state)
    finally:
        flow.eval()
