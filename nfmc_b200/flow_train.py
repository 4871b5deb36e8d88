"""Flow training on the device: ``Flow.fit`` (maximum likelihood) and ``Flow.variational_fit`` (reverse KL).

Status: training runs **natively** for every conditioner shape.
* default conditioners (2 linear layers, <= 8 hidden units -- the register-resident kernel path):
  ``csrc/train_kernels.cu`` computes the loss and the gradient with respect to every flow parameter in one launch
  (reversible backward sweep, no stored activations), ``csrc/train_api.cu`` packs the parameters, maps the gradient back
  to module order and applies AdamW; one C call per epoch on a single GPU, per-step calls with an NCCL all-reduce of the
  gradient in between on several.
* wide / deep conditioners (any number of linear layers, any hidden width): ``csrc/train_wide.cu`` -- row tiles in shared
  memory, forward / dgrad / wgrad contractions as register-blocked fp32 loops, parameters and gradients in module order
  (no pack / unpack); reverse KL = fp32 pass kernel + ``nfmc_potential_eval`` + backward-sweep kernel.
There is no torch-autograd training loop in the product.  ``forward_autograd`` / ``inverse_autograd`` below are the torch
restatement used for the one-off data-dependent ActNorm initialisation and by the tests.  After every fit the packed blobs
are rebuilt so the samplers keep using the CUDA kernels.

Reference call sites: ``flow.fit`` -- /root/reference/nfmc/algorithms/sampling/nfmc/jump.py:139-151,201 and
nfmc/imh.py:171-175; ``flow.variational_fit`` -- nfmc/imh.py:67-72 and nfmc/neutra.py:84-91;
``train_val_split`` -- sampling/tuning.py:44-65.  The optimiser settings follow the kwargs the reference passes
(``lr=0.05``, ``early_stopping``, ``early_stopping_threshold``, ``keep_best_weights``, ``batch_size='adaptive'``,
``time_limit_seconds``); torchflows' own defaults are unpinned (the package is absent).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _native as N

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
    fn = lambda v: _PotentialFn.apply(v.reshape(v.shape[0], -1), potential)   # noqa: E731
    fn.potential = potential          # lets variational_fit take the native reverse-KL kernel
    return fn


# ---------------------------------------------------------------------------------------------------------------
# native training (csrc/train_kernels.cu, csrc/train_api.cu)
# ---------------------------------------------------------------------------------------------------------------
ADAMW = dict(beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01)     # torch.optim.AdamW defaults


def _same_shape(flow):
    bij = flow.bijection
    M, H = bij.conditioner_shape()
    return all((c.n_linear, c.n_hidden) == (M, H) for c in bij.couplings()), M, H


def native_supported(flow) -> bool:
    """The register-resident kernels cover conditioners with 2 linear layers and <= 8 hidden units (every default one)."""
    if os.environ.get("NFMC_B200_WIDE_TRAINING") == "1":          # tests: force the wide kernel on a default flow
        return False
    same, M, H = _same_shape(flow)
    bij = flow.bijection
    return same and M == 2 and H <= 8 and N.lib().nfmc_flow_param_count(bij.n_dim, bij.n_coupling, M, H) > 0


def wide_supported(flow) -> bool:
    """Every other shape goes to csrc/train_wide.cu (all couplings must share one conditioner shape)."""
    same, M, H = _same_shape(flow)
    bij = flow.bijection
    return same and N.lib().nfmc_flow_wide_param_count(bij.n_dim, bij.n_coupling, M, H) > 0


def _unsupported(flow):
    same, M, H = _same_shape(flow)
    return NotImplementedError(f"flow training: conditioner shape (n_layers={M}, n_hidden={H}, uniform={same}) is outside the "
                               "native kernels; there is no eager fallback")


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _out_of_time(t0: float, limit, dev) -> bool:
    """Time-limit test that every rank answers the same way (a rank leaving the loop alone would hang the others)."""
    if limit is None:
        return False
    over = time.time() - t0 > limit
    if _world() > 1:
        flag = torch.tensor([1.0 if over else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        over = bool(flag.item() > 0)
    return over


class NativeTrainer:
    """Flat parameter vector + AdamW state on the device, driven through the C ABI."""

    def __init__(self, flow, dev, lr: float):
        self.flow, self.dev, self.lr = flow, dev, float(lr)
        bij = flow.bijection
        self.d, self.Lc = bij.n_dim, bij.n_coupling
        self.M, self.H = bij.conditioner_shape()
        self.params = [p for p in bij.parameters()]
        self.theta = torch.cat([p.detach().reshape(-1) for p in self.params]).to(dev, torch.float32).contiguous()
        if _world() > 1:                      # data parallel: every rank starts from rank 0's parameters (this also makes
            dist.broadcast(self.theta, src=0)  # the data-dependent ActNorm initialisation rank 0's)
        P = N.lib().nfmc_flow_param_count(self.d, self.Lc, self.M, self.H)
        if P != self.theta.numel():
            raise RuntimeError(f"parameter count mismatch: module {self.theta.numel()} vs native {P}")
        from .flow import blob_floats
        nb = blob_floats(self.d, self.Lc, self.M, self.H)
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.blob = torch.empty(nb, device=dev, dtype=torch.float32)
        self.gblob = torch.empty(nb, device=dev, dtype=torch.float32)
        self.gtheta = torch.empty_like(self.theta)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float64)
        self.step = 0
        self.stream = N.stream_ptr(dev)

    def desc(self) -> N.RealNVPDesc:
        return N.RealNVPDesc(self.d, self.Lc, self.M, self.H, self.blob.data_ptr(), self.blob.numel())

    def pack(self):
        N.check(N.lib().nfmc_flow_pack(self.d, self.Lc, self.M, self.H, N.ptr(self.theta), N.ptr(self.blob), self.stream))

    def unpack(self, scale: float):
        N.check(N.lib().nfmc_flow_grad_unpack(self.d, self.Lc, self.M, self.H, N.ptr(self.theta), N.ptr(self.gblob),
                                              float(scale), N.ptr(self.gtheta), self.stream))

    def sync_grad(self):
        if _world() > 1:
            dist.all_reduce(self.gtheta)
            self.gtheta /= _world()

    def adamw(self):
        self.step += 1
        N.check(N.lib().nfmc_adamw_step(N.ptr(self.theta), N.ptr(self.gtheta), N.ptr(self.m), N.ptr(self.v),
                                        self.theta.numel(), self.lr, ADAMW["beta1"], ADAMW["beta2"], ADAMW["eps"],
                                        ADAMW["weight_decay"], self.step, self.stream))

    def nll_step(self, x: torch.Tensor, rows: Optional[torch.Tensor], n: int):
        """One optimiser step on ``x[rows]`` (maximum likelihood); ``self.loss`` holds the summed loss afterwards."""
        self.pack()
        d = self.desc()
        N.check(N.lib().nfmc_flow_nll_grad(C.byref(d), N.ptr(x), None if rows is None else rows.data_ptr(), n,
                                           N.ptr(self.gblob), N.ptr(self.loss), 0, self.stream))
        self.unpack(1.0 / n)
        self.sync_grad()
        self.adamw()

    def nll_epoch(self, x: torch.Tensor, perm: torch.Tensor, batch_size: int, losses: torch.Tensor):
        n = perm.numel()
        if _world() > 1:
            for b, i in enumerate(range(0, n, batch_size)):
                m = min(batch_size, n - i)
                self.nll_step(x, perm[i:i + m], m)
                losses[b] = self.loss[0]
            return
        N.check(N.lib().nfmc_flow_fit_epoch(self.d, self.Lc, self.M, self.H, N.ptr(self.theta), N.ptr(self.m), N.ptr(self.v),
                                            N.ptr(self.blob), N.ptr(self.gblob), N.ptr(self.gtheta), N.ptr(losses),
                                            N.ptr(x), perm.data_ptr(), n, batch_size, self.lr, ADAMW["beta1"],
                                            ADAMW["beta2"], ADAMW["eps"], ADAMW["weight_decay"], self.step, self.stream))
        self.step += (n + batch_size - 1) // batch_size

    def kl_step(self, potential, n_samples: int, seed: int, step0: int, z: Optional[torch.Tensor] = None):
        self.pack()
        d = self.desc()
        pot, keep = potential.descriptor(self.dev)
        rng = N.rng_desc(seed, step0, z, None)
        chain0 = (dist.get_rank() * n_samples) if _world() > 1 else 0
        N.check(N.lib().nfmc_flow_kl_grad(C.byref(pot), C.byref(d), C.byref(rng), chain0, n_samples, N.ptr(self.gblob),
                                          N.ptr(self.loss), 0, self.stream))
        self.unpack(1.0 / n_samples)
        self.sync_grad()
        self.adamw()

    def mean_nll(self, x: torch.Tensor) -> torch.Tensor:
        """-mean log q(x) under the current theta (device scalar)."""
        self.pack()
        d = self.desc()
        lq = torch.empty(x.shape[0], device=self.dev, dtype=torch.float32)
        N.check(N.lib().nfmc_flow_log_prob(C.byref(d), N.ptr(x), N.ptr(lq), x.shape[0], self.stream))
        return -lq.double().mean()

    @torch.no_grad()
    def write_back(self, theta: Optional[torch.Tensor] = None):
        theta = self.theta if theta is None else theta
        off = 0
        for p in self.params:
            k = p.numel()
            p.copy_(theta[off:off + k].view_as(p))
            off += k


class WideTrainer(NativeTrainer):
    """Same driver for wide / deep conditioners (csrc/train_wide.cu): theta and its gradient stay in module order."""

    def __init__(self, flow, dev, lr: float):
        self.flow, self.dev, self.lr = flow, dev, float(lr)
        bij = flow.bijection
        self.d, self.Lc = bij.n_dim, bij.n_coupling
        self.M, self.H = bij.conditioner_shape()
        self.params = [p for p in bij.parameters()]
        self.theta = torch.cat([p.detach().reshape(-1) for p in self.params]).to(dev, torch.float32).contiguous()
        if _world() > 1:
            dist.broadcast(self.theta, src=0)
        P = N.lib().nfmc_flow_wide_param_count(self.d, self.Lc, self.M, self.H)
        if P != self.theta.numel():
            raise RuntimeError(f"parameter count mismatch: module {self.theta.numel()} vs native {P}")
        self.m = torch.zeros_like(self.theta)
        self.v = torch.zeros_like(self.theta)
        self.gtheta = torch.empty_like(self.theta)
        self.loss = torch.zeros(1, device=dev, dtype=torch.float64)
        self.step = 0
        self.stream = N.stream_ptr(dev)

    def _shape(self):
        return self.d, self.Lc, self.M, self.H

    def adamw_scaled(self, scale: float):
        self.step += 1
        N.check(N.lib().nfmc_adamw_step_scaled(N.ptr(self.theta), N.ptr(self.gtheta), float(scale), N.ptr(self.m), N.ptr(self.v),
                                               self.theta.numel(), self.lr, ADAMW["beta1"], ADAMW["beta2"], ADAMW["eps"],
                                               ADAMW["weight_decay"], self.step, self.stream))

    def nll_grad(self, x: torch.Tensor, rows: Optional[torch.Tensor], n: int, grad_x: Optional[torch.Tensor] = None):
        N.check(N.lib().nfmc_flow_wide_nll_grad(*self._shape(), N.ptr(self.theta), N.ptr(x), None if rows is None else rows.data_ptr(),
                                                n, N.ptr(self.gtheta), N.ptr(self.loss), N.ptr(grad_x), 0, self.stream))

    def nll_step(self, x: torch.Tensor, rows: Optional[torch.Tensor], n: int):
        self.nll_grad(x, rows, n)
        self.sync_grad()
        self.adamw_scaled(1.0 / n)

    def nll_epoch(self, x: torch.Tensor, perm: torch.Tensor, batch_size: int, losses: torch.Tensor):
        n = perm.numel()
        if _world() > 1:
            for b, i in enumerate(range(0, n, batch_size)):
                m = min(batch_size, n - i)
                self.nll_step(x, perm[i:i + m], m)
                losses[b] = self.loss[0]
            return
        N.check(N.lib().nfmc_flow_wide_fit_epoch(*self._shape(), N.ptr(self.theta), N.ptr(self.m), N.ptr(self.v), N.ptr(self.gtheta),
                                                 N.ptr(losses), N.ptr(x), perm.data_ptr(), n, batch_size, self.lr, ADAMW["beta1"],
                                                 ADAMW["beta2"], ADAMW["eps"], ADAMW["weight_decay"], self.step, self.stream))
        self.step += (n + batch_size - 1) // batch_size

    def run_pass(self, v: torch.Tensor, inverse: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """fp32 pass under the current theta: (output [n, d], log|det| [n])."""
        n = v.shape[0]
        out = torch.empty_like(v)
        ld = torch.empty(n, device=self.dev, dtype=torch.float32)
        N.check(N.lib().nfmc_flow_wide_pass(*self._shape(), N.ptr(self.theta), 1 if inverse else 0, N.ptr(v), N.ptr(out), N.ptr(ld),
                                            n, self.stream))
        return out, ld

    def kl_step(self, potential, n_samples: int, seed: int, step0: int, z: Optional[torch.Tensor] = None):
        """One reverse-KL step: z ~ N(0, I) (Philox stream 1, as the register-resident kernel draws it), x = T^-1(z) by the
        pass kernel, U and grad U by the potential kernel, then the backward sweep and AdamW.  ``self.loss`` = sum_i
        [log q(x_i) + U(x_i)] under the parameters before the step."""
        d = self.d
        if z is None:
            z = torch.empty(n_samples, d, device=self.dev, dtype=torch.float32)
            rng = N.rng_desc(seed, step0, None, None)
            chain0 = (dist.get_rank() * n_samples) if _world() > 1 else 0
            N.check(N.lib().nfmc_rng_fill(C.byref(rng), 1, chain0, d, n_samples, 1, N.ptr(z), None, self.stream))
        x, ld_inv = self.run_pass(z, inverse=True)
        u, gu = potential.value_and_grad(x, need_grad=True)
        gu = gu.reshape(n_samples, d).contiguous()
        log_q = (-0.5 * z.square()).sum(dim=1) - 0.5 * d * math.log(2 * math.pi) - ld_inv
        self.loss[0] = (log_q + u).double().sum()
        N.check(N.lib().nfmc_flow_wide_sweep(*self._shape(), N.ptr(self.theta), 1, N.ptr(x), N.ptr(gu), n_samples, N.ptr(self.gtheta),
                                             None, 0, self.stream))
        self.sync_grad()
        self.adamw_scaled(1.0 / n_samples)

    def mean_nll(self, x: torch.Tensor) -> torch.Tensor:
        z, ld = self.run_pass(x, inverse=False)
        lq = (-0.5 * z.square()).sum(dim=1) - 0.5 * self.d * math.log(2 * math.pi) + ld
        return -lq.double().mean()


def make_trainer(flow, dev, lr: float, external_target: bool = False) -> NativeTrainer:
    """``external_target``: the reverse-KL target is a Python callable -- only the wide trainer takes U / grad U from outside
    (its sweep kernel is seeded with grad U(x)); the register-resident kernel evaluates the potential itself."""
    if native_supported(flow) and not external_target:
        return NativeTrainer(flow, dev, lr)
    if wide_supported(flow):
        return WideTrainer(flow, dev, lr)
    raise _unsupported(flow)


def _init_actnorms(flow, x_first: torch.Tensor):
    """One-off data-dependent ActNorm initialisation (torch restatement; runs once per flow lifetime)."""
    from .flow import ActNorm
    if any(isinstance(l, ActNorm) and not bool(l.initialised) for l in flow.bijection.layers):
        with torch.no_grad():
            forward_autograd(flow.bijection, x_first, training=True)


def _fit_native(flow, dev, x_train, x_val, n_epochs, lr, batch_size, shuffle, keep_best_weights, early_stopping,
                early_stopping_threshold, time_limit_seconds):
    n = len(x_train)
    if _world() > 1:
        # every rank issues one gradient all-reduce per minibatch: agree on the training-set size first (shard remainders
        # and the train/val split can leave the ranks a few rows apart), otherwise the collectives would not pair up
        t = torch.tensor([n], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t)
        if n < 1:
            raise ValueError("flow.fit: a rank has no training rows")
        x_train = x_train[:n]
    _init_actnorms(flow, x_train[:batch_size])
    tr = make_trainer(flow, dev, lr)
    n_batches = (n + batch_size - 1) // batch_size
    losses = torch.zeros(n_batches, device=dev, dtype=torch.float64)
    ref = x_val if x_val is not None and len(x_val) else x_train
    best, since_best = math.inf, 0
    best_theta = tr.theta.clone() if keep_best_weights else None
    t0 = time.time()
    try:
        for _ in range(n_epochs):
            if _out_of_time(t0, time_limit_seconds, dev):
                break
            perm = torch.randperm(n, device=dev) if shuffle else torch.arange(n, device=dev)
            tr.nll_epoch(x_train, perm, batch_size, losses)
            pair = torch.stack([tr.mean_nll(ref), losses.sum()])
            if _world() > 1:                   # every rank must take the same keep-best / early-stopping decisions
                dist.all_reduce(pair)
                pair /= _world()
            score, train_sum = (float(v) for v in pair.cpu())                                  # the epoch's one sync
            if not math.isfinite(train_sum):
                raise ValueError("Flow training diverged")                 # the reference rolls back on ValueError
            if score < best:
                best, since_best = score, 0
                if keep_best_weights:
                    best_theta.copy_(tr.theta)
            else:
                since_best += 1
                if early_stopping and since_best >= early_stopping_threshold:
                    break
        tr.write_back(best_theta if keep_best_weights and math.isfinite(best) else None)
    finally:
        flow.eval()


def _variational_fit_native(flow, dev, potential, n_epochs, lr, n_samples, early_stopping, early_stopping_threshold,
                            keep_best_weights, check_for_divergences, time_limit_seconds):
    tr = make_trainer(flow, dev, lr, external_target=bool(getattr(potential, "external", False)))
    seed = int(torch.randint(0, 2 ** 62, (), dtype=torch.int64))
    best, since_best = math.inf, 0
    best_theta = tr.theta.clone() if keep_best_weights else None
    prev = tr.theta.clone()
    t0 = time.time()
    try:
        for epoch in range(n_epochs):
            if _out_of_time(t0, time_limit_seconds, dev):
                break
            prev.copy_(tr.theta)
            tr.kl_step(potential, n_samples, seed, epoch)
            if _world() > 1:
                dist.all_reduce(tr.loss)
                tr.loss /= _world()
            val = float(tr.loss[0]) / n_samples            # loss of the parameters BEFORE this step (as autograd reports it)
            if not math.isfinite(val):
                tr.theta.copy_(prev)
                if check_for_divergences:
                    break
                raise ValueError("Flow training diverged")
            if val < best:
                best, since_best = val, 0
                if keep_best_weights:
                    best_theta.copy_(prev)
            else:
                since_best += 1
                if early_stopping and since_best >= early_stopping_threshold:
                    break
        tr.write_back(best_theta if keep_best_weights and math.isfinite(best) else None)
    finally:
        flow.eval()


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
    batch_size = int(batch_size)
    return _fit_native(flow, dev, x_train.contiguous(), None if x_val is None else x_val.contiguous(), n_epochs, lr,
                       batch_size, shuffle, keep_best_weights, early_stopping, early_stopping_threshold,
                       time_limit_seconds)


def variational_fit(flow, target_log_prob: Callable, n_epochs: int = 500, lr: float = 0.05, n_samples: int = 1,
                    early_stopping: bool = False, early_stopping_threshold: int = 50, keep_best_weights: bool = True,
                    show_progress: bool = False, check_for_divergences: bool = False, time_limit_seconds=None, **_ignored):
    dev = flow._compute_device()
    flow.to(dev)
    d = flow.bijection.n_dim
    potential = getattr(target_log_prob, "potential", None)
    if potential is None:
        raise NotImplementedError("variational_fit needs `potential.log_prob_fn()` of a nfmc_b200.potentials object (a built-in "
                                  "potential or CallablePotential(fn, event_shape)), which carries U / grad U for the native sweep")
    return _variational_fit_native(flow, dev, potential, n_epochs, lr, int(n_samples), early_stopping,
                                   early_stopping_threshold, keep_best_weights, check_for_divergences,
                                   time_limit_seconds)
