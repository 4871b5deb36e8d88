"""Result and configuration records -- the reference's contract, re-implemented on top of device-side sums.

Same class names, fields, properties and meaning as ``/root/reference/nfmc/algorithms/sampling/base.py``
(``MCMCStatistics`` :126-212, ``MCMCSamples`` :215-271, ``MCMCOutput`` :274-314, kernels / parameters :9-61),
``mcmc/base.py:105-131``, ``mcmc/langevin.py:10-28``, ``mcmc/hmc.py:10-23``, ``nfmc/jump.py:21-81``,
``nfmc/neutra.py:19-33`` and ``nfmc/imh.py:13-36``.  The difference is in how the running moments are kept:
the reference streams a mean over host tensors (``MCMCExpectation.update``, base.py:75-95); here the kernels
accumulate ``sum x`` / ``sum x^2`` in fp64 on the device and the records hold those sums plus a count, which
is what makes pooling over GPUs a plain all-reduce.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from .flow import Flow, RealNVP


# ---------------------------------------------------------------------------------------------------------------
# kernels (adaptive state) and parameters (hyper-parameters)
# ---------------------------------------------------------------------------------------------------------------
@dataclass
class MCMCKernel:
    def __post_init__(self):
        pass


@dataclass
class NFMCKernel(MCMCKernel):
    event_shape: Union[Tuple[int, ...], torch.Size]
    flow: Flow = None

    def __post_init__(self):
        super().__post_init__()
        if self.flow is None:
            self.flow = Flow(RealNVP(self.event_shape))


@dataclass
class IMHKernel(NFMCKernel):
    pass


@dataclass
class NeuTraKernel(NFMCKernel):
    pass


@dataclass
class DualAveragingParams:
    target_acceptance_rate: float = 0.651
    kappa: float = 0.75
    gamma: float = 0.05
    t0: int = 10


class DualAveraging:
    """Nesterov dual averaging of the log step size (reference: sampling/tuning.py:15-41)."""

    def __init__(self, initial_step_size: float, params: DualAveragingParams):
        self.p = params
        self.t = params.t0
        self.error_sum = 0.0
        self.log_step_averaged = math.log(initial_step_size)
        self.log_step = math.inf
        self.mu = math.log(10 * initial_step_size)

    def step(self, acceptance_rate_error: float):
        self.error_sum += float(acceptance_rate_error)
        self.log_step = self.mu - self.error_sum / (math.sqrt(self.t) * self.p.gamma)
        eta = self.t ** -self.p.kappa
        self.log_step_averaged = eta * self.log_step + (1 - eta) * self.log_step_averaged
        self.t += 1

    @property
    def value(self) -> float:
        return math.exp(self.log_step_averaged)

    def __repr__(self):
        return f'DA error: {self.error_sum:.2f}'


@dataclass
class MetropolisKernel(MCMCKernel):
    event_size: int
    inv_mass_diag: torch.Tensor = None
    step_size: float = 0.01
    da: DualAveraging = None
    da_params: DualAveragingParams = None

    def __post_init__(self):
        super().__post_init__()
        if self.inv_mass_diag is None:
            self.inv_mass_diag = torch.ones(self.event_size)
        elif tuple(self.inv_mass_diag.shape) != (self.event_size,):
            raise ValueError("inv_mass_diag must have shape (event_size,)")
        if self.da_params is None:
            self.da_params = DualAveragingParams()
        if self.da is None:
            self.da = DualAveraging(self.step_size, self.da_params)

    def has_unit_mass(self) -> bool:
        if self.inv_mass_diag.is_cuda:           # device-resident during warm-up: never the untouched default, and no sync
            return False
        return bool(torch.all(self.inv_mass_diag == 1.0))


@dataclass
class LangevinKernel(MetropolisKernel):
    event_size: int = None
    step_size: Optional[float] = None

    def __post_init__(self):
        if self.step_size is None:
            self.step_size = self.event_size ** (-1 / 3)   # reference: mcmc/langevin.py:17-18
        super().__post_init__()

    def __repr__(self):
        return (f'log step: {math.log(self.step_size):.2f}, '
                f'mass norm: {torch.max(torch.abs(self.inv_mass_diag)):.2f}')


@dataclass
class HMCKernel(MetropolisKernel):
    event_size: int = None
    n_leapfrog_steps: int = 20

    def __repr__(self):
        return (f'log step: {math.log(self.step_size):.2f}, leapfrogs: {self.n_leapfrog_steps}, '
                f'mass norm: {torch.max(torch.abs(self.inv_mass_diag)):.2f}')


@dataclass
class MHKernel(MetropolisKernel):
    event_size: int = None

    def __repr__(self):
        return (f'log step: {math.log(self.step_size):.2f}, '
                f'mass norm: {torch.max(torch.abs(self.inv_mass_diag)):.2f}')


@dataclass
class ESSKernel(MCMCKernel):
    """Reference: mcmc/ess.py:67-70.  Only the identity prior covariance (``cov=None``) runs on the device path."""
    event_shape: tuple = None
    cov: torch.Tensor = None

    def __post_init__(self):
        super().__post_init__()
        if self.cov is not None:
            raise NotImplementedError("ESSKernel: only cov=None (identity prior covariance) is supported on the device path")

    def __repr__(self):
        return 'ESS kernel (identity prior covariance)'


@dataclass
class MCMCParameters:
    n_iterations: int = 100
    n_warmup_iterations: int = 100
    tuning: bool = False
    store_samples: bool = True

    def __post_init__(self):
        pass

    def tuning_mode(self):
        self.tuning = True

    def sampling_mode(self):
        self.tuning = False


@dataclass
class ESSParameters(MCMCParameters):
    max_ess_step_iterations: int = 5          # reference: mcmc/ess.py:73-75


@dataclass
class MetropolisParameters(MCMCParameters):
    tune_inv_mass_diag: bool = True
    tune_step_size: bool = True
    adjustment: bool = True
    imd_adjustment: float = 1e-3


@dataclass
class LangevinParameters(MetropolisParameters):
    pass


@dataclass
class HMCParameters(MetropolisParameters):
    pass


@dataclass
class MHParameters(MetropolisParameters):
    imd_adjustment: float = 1e-5

    def __post_init__(self):
        self.tune_step_size = False          # reference: mcmc/mh.py:22-24
        self.tune_inv_mass_diag = True


@dataclass
class NFMCParameters(MCMCParameters):
    train_pct: float = 0.7
    max_train_size: int = 4096
    max_val_size: int = 4096
    flow_fit_kwargs: Dict[str, Any] = None

    def __post_init__(self):
        super().__post_init__()
        if self.flow_fit_kwargs is None:
            self.flow_fit_kwargs = dict(early_stopping=True, early_stopping_threshold=50, batch_size='adaptive',
                                        show_progress=False)


@dataclass
class JumpNFMCParameters(NFMCParameters):
    adjusted_jumps: bool = True
    fit_nf: bool = False
    warmup_fit_kwargs: dict = None
    n_jumps_before_training: int = 10

    def __post_init__(self):
        super().__post_init__()
        if self.warmup_fit_kwargs is None:
            self.warmup_fit_kwargs = dict(early_stopping=True, early_stopping_threshold=50, keep_best_weights=True,
                                          n_samples=1, n_epochs=500, lr=0.05)


@dataclass
class DLMCKernel(NFMCKernel):
    step_size: float = 0.05                      # reference: nfmc/dlmc.py:12-14


@dataclass
class DLMCParameters(NFMCParameters):
    latent_updates: bool = False                 # reference: nfmc/dlmc.py:17-19


@dataclass
class TESSKernel(NFMCKernel):
    """Reference: nfmc/tess.py:78-80 (ESSKernel + NFMCKernel).  Only the identity covariance runs on the device path."""
    cov: torch.Tensor = None

    def __post_init__(self):
        super().__post_init__()
        if self.cov is not None:
            raise NotImplementedError("TESSKernel: only cov=None (identity covariance) is supported on the device path")


@dataclass
class TESSParameters(NFMCParameters):
    """Reference: nfmc/tess.py:83-85 (ESSParameters + NFMCParameters)."""
    max_ess_step_iterations: int = 5
    n_warmup_iterations: int = 20


@dataclass
class NeuTraParameters(NFMCParameters):
    batch_inverse_size: int = 128
    warmup_fit_kwargs: dict = None

    def __post_init__(self):
        # as in the reference (nfmc/neutra.py:24-33) the parent hook is not called: flow_fit_kwargs stays None (Q4)
        if self.warmup_fit_kwargs is None:
            self.warmup_fit_kwargs = dict(early_stopping=True, early_stopping_threshold=5000, keep_best_weights=True,
                                          n_samples=1, n_epochs=50000, lr=0.05)


@dataclass
class IMHParameters(NFMCParameters):
    train_distribution: str = 'uniform'
    adaptation_dropoff: float = 0.9999
    warmup_fit_kwargs: dict = None

    def __post_init__(self):
        if self.train_distribution not in ('bounded_geom_approx', 'bounded_geom', 'uniform'):
            raise ValueError(self.train_distribution)
        if self.warmup_fit_kwargs is None:
            self.warmup_fit_kwargs = dict(early_stopping=True, early_stopping_threshold=50, keep_best_weights=True,
                                          n_samples=1, n_epochs=500, lr=0.05, check_for_divergences=True)


# ---------------------------------------------------------------------------------------------------------------
# statistics
# ---------------------------------------------------------------------------------------------------------------
class MomentSums:
    """sum f(x) over every (iteration, chain) pair seen so far, in fp64; E[f(x)] = sum / n_seen."""

    def __init__(self, event_shape):
        self.event_shape = tuple(event_shape)
        d = int(math.prod(self.event_shape))
        self.sum_x = torch.zeros(d, dtype=torch.float64)
        self.sum_x2 = torch.zeros(d, dtype=torch.float64)
        self.n_seen = 0

    def add_sums(self, sum_x: torch.Tensor, sum_x2: torch.Tensor, count: int):
        self.sum_x += sum_x.detach().to("cpu", torch.float64).reshape(-1)
        self.sum_x2 += sum_x2.detach().to("cpu", torch.float64).reshape(-1)
        self.n_seen += int(count)

    def update(self, x: torch.Tensor):
        """Accumulate a block ``[k, n, *event]`` or ``[n, *event]`` (API compatibility with
        ``MCMCExpectationDict.update``, base.py:110-113)."""
        k = len(self.event_shape)
        flat = x.detach().reshape(-1, *x.shape[x.ndim - k:]).reshape(-1, self.sum_x.numel()).to(torch.float64)
        self.add_sums(flat.sum(0), flat.square().sum(0), flat.shape[0])

    def reset(self):
        self.sum_x.zero_()
        self.sum_x2.zero_()
        self.n_seen = 0

    def first(self):
        if self.n_seen == 0:
            return 0.0
        return (self.sum_x / self.n_seen).to(torch.float32).reshape(self.event_shape)

    def second(self):
        if self.n_seen == 0:
            return 0.0
        return (self.sum_x2 / self.n_seen).to(torch.float32).reshape(self.event_shape)

    # dict-style access used by callers of the reference API: expectations['first_moment'].as_tensor()
    class _View:
        def __init__(self, fn):
            self._fn = fn

        def as_tensor(self):
            return self._fn()

    def __getitem__(self, key):
        if key == 'first_moment':
            return MomentSums._View(self.first)
        if key == 'second_moment':
            return MomentSums._View(self.second)
        raise KeyError(key)

    def as_tensor(self):
        return {'first_moment': self.first(), 'second_moment': self.second()}


@dataclass
class MCMCStatistics:
    event_shape: Union[Tuple[int, ...], torch.Size]
    n_accepted_trajectories: Optional[int] = 0
    n_attempted_trajectories: Optional[int] = 0
    n_divergences: Optional[int] = 0
    n_target_gradient_calls: Optional[int] = 0
    n_target_calls: Optional[int] = 0
    elapsed_time_seconds: Optional[float] = 0.0
    data_transform: callable = lambda v: v
    expectations: MomentSums = None

    def __post_init__(self):
        self.expectations = MomentSums(self.event_shape)

    def update_counters(self, n_accepted_trajectories: int = 0, n_attempted_trajectories: int = 0,
                        n_divergences: int = 0, n_target_gradient_calls: int = 0, n_target_calls: int = 0):
        self.n_accepted_trajectories = int(self.n_accepted_trajectories + n_accepted_trajectories)
        self.n_attempted_trajectories = int(self.n_attempted_trajectories + n_attempted_trajectories)
        self.n_divergences = int(self.n_divergences + n_divergences)
        self.n_target_gradient_calls = int(self.n_target_gradient_calls + n_target_gradient_calls)
        self.n_target_calls = int(self.n_target_calls + n_target_calls)

    def update_elapsed_time(self, delta_time_seconds: float):
        self.elapsed_time_seconds = float(self.elapsed_time_seconds + delta_time_seconds)

    @property
    def running_first_moment(self):
        return self.expectations.first()

    @property
    def running_second_moment(self):
        return self.expectations.second()

    @property
    def running_variance(self):
        return self.running_second_moment - self.running_first_moment ** 2

    @property
    def acceptance_rate(self):
        if self.n_attempted_trajectories == 0:
            return torch.nan
        return self.n_accepted_trajectories / self.n_attempted_trajectories

    @property
    def calls_per_second(self):
        return self.n_target_calls / self.elapsed_time_seconds if self.elapsed_time_seconds > 0 else torch.nan

    @property
    def grads_per_second(self):
        return self.n_target_gradient_calls / self.elapsed_time_seconds if self.elapsed_time_seconds > 0 else torch.nan

    def __repr__(self):
        return (f"acc-rate: {self.acceptance_rate:.2f}, kcalls/s: {self.calls_per_second / 1000:.2f}, "
                f"kgrads/s: {self.grads_per_second / 1000:.2f}, divergences: {self.n_divergences}")

    def __dict__(self):
        return dict(n_accepted_trajectories=self.n_accepted_trajectories,
                    n_attempted_trajectories=self.n_attempted_trajectories, n_divergences=self.n_divergences,
                    n_target_gradient_calls=self.n_target_gradient_calls, n_target_calls=self.n_target_calls,
                    elapsed_time_seconds=self.elapsed_time_seconds, grads_per_second=self.grads_per_second,
                    acceptance_rate=self.acceptance_rate, calls_per_second=self.calls_per_second)


@dataclass
class JumpNFMCStatistics(MCMCStatistics):
    n_accepted_jumps: int = 0
    n_attempted_jumps: int = 0

    @property
    def jump_acceptance_rate(self):
        if self.n_attempted_jumps == 0:
            return torch.nan
        return self.n_accepted_jumps / self.n_attempted_jumps

    def update_counters(self, n_accepted_jumps: int = 0, n_attempted_jumps: int = 0, **kwargs):
        super().update_counters(**kwargs)
        self.n_accepted_jumps = int(self.n_accepted_jumps + n_accepted_jumps)
        self.n_attempted_jumps = int(self.n_attempted_jumps + n_attempted_jumps)

    def __repr__(self):
        return (f"MCMC acc-rate: {self.acceptance_rate:.2f}, Jump acc-rate: {self.jump_acceptance_rate:.2f}, "
                f"kcalls/s: {self.calls_per_second / 1000:.2f}, kgrads/s: {self.grads_per_second / 1000:.2f}, "
                f"divergences: {self.n_divergences}")

    def __dict__(self):
        return {**super().__dict__(), 'jump_acceptance_rate': self.jump_acceptance_rate}


# ---------------------------------------------------------------------------------------------------------------
# samples and output
# ---------------------------------------------------------------------------------------------------------------
DEVICE_SAMPLE_BUDGET_BYTES = int(os.environ.get("NFMC_B200_DEVICE_SAMPLE_BYTES", 48 << 30))
PINNED_HOST_LIMIT_BYTES = 8 << 30


class MCMCSamples:
    """Sample store (reference: sampling/base.py:215-271).  Same attributes and semantics (thinning, ``max_samples``
    window, ``last_sample``), different residency: blocks written by the kernels' sample sink **stay on the GPU** (up
    to ``NFMC_B200_DEVICE_SAMPLE_BYTES``, default 48 GiB of the 180 GB of HBM; beyond that they spill to the host as they
    arrive) and the host tensor the reference would have built with one ``.cpu()`` per iteration is materialised once,
    on demand, by ``as_tensor()`` -- through a pinned buffer, all copies in flight together.  Consumers on the device
    (flow refits ``imh.py:152-175``, warm-up hand-off ``sample.py:307-313``) use ``row()`` / ``device_tensor()`` and
    never touch the host.  ``last_sample`` is materialised lazily in the same way."""

    def __init__(self, event_shape, store_samples: bool = True, n_samples: int = 0, last_sample: torch.Tensor = None,
                 thinning: int = 1, seen_samples: int = 0, max_samples: int = None):
        self.event_shape = tuple(event_shape)
        self.store_samples = store_samples
        self.n_samples = n_samples
        self.thinning = thinning
        self.seen_samples = seen_samples
        self.max_samples = max_samples
        self._blocks: List[torch.Tensor] = []   # each [k, n, *event]; on the device while the budget allows, else host
        self._device_bytes = 0
        self._last = last_sample
        self.last_sample_device: Optional[torch.Tensor] = None

    @property
    def last_sample(self):
        if self._last is None and self.last_sample_device is not None:
            self._last = self.last_sample_device.detach().cpu()
        return self._last

    @last_sample.setter
    def last_sample(self, value):
        self._last = value
        self.last_sample_device = None

    def set_last_device(self, x_dev: torch.Tensor):
        """Record the final state without copying it to the host."""
        self._last = None
        self.last_sample_device = x_dev

    def row(self, index: int) -> torch.Tensor:
        """Stored iteration ``index`` as ``[n, *event]`` wherever it lives (device or host), without materialising the
        whole store."""
        if index < 0:
            index += self.n_samples
        for blk in self._blocks:
            if index < len(blk):
                return blk[index]
            index -= len(blk)
        raise IndexError("sample index out of range")

    def __getitem__(self, index):
        if index == -1 or index == self.n_samples - 1:      # reference: base.py:229-232 (the last SEEN state)
            return self.last_sample_device if self.last_sample_device is not None else self.last_sample
        return self.row(index)

    def add(self, x: torch.Tensor, already_thinned: bool = False, n_seen: Optional[int] = None):
        """Append a block ``[k, n, *event]`` or one state ``[n, *event]`` (reference: base.py:234-263).
        ``already_thinned``: ``x`` is a buffer the device sink filled under the thinning rule (owned by the store from
        here on); ``n_seen`` = steps the block stands for."""
        k = len(self.event_shape)
        if x.ndim == k + 1 and tuple(x.shape[1:]) == tuple(self.event_shape):
            x = x[None]
        elif not (x.ndim == k + 2 and tuple(x.shape[2:]) == tuple(self.event_shape)):
            raise ValueError(f"Expected x.shape[1:] or x.shape[2:] to be {self.event_shape}, got {x.shape = }")
        if len(x):
            if x.is_cuda:
                self.set_last_device(x[-1].detach().clone())
            else:
                self.last_sample = x[-1].detach().clone()
        if not self.store_samples:
            return
        if already_thinned:
            kept = x.detach()
            self.seen_samples += int(n_seen if n_seen is not None else len(x))
        else:
            idx = torch.arange(self.seen_samples, self.seen_samples + len(x))
            kept = x.detach()[(idx % self.thinning) == 0]      # advanced indexing: a copy, never a view of live state
            self.seen_samples += len(x)
        if len(kept):
            nbytes = kept.numel() * kept.element_size()
            if kept.is_cuda and self._device_bytes + nbytes <= DEVICE_SAMPLE_BUDGET_BYTES:
                self._device_bytes += nbytes
            else:
                kept = kept.cpu()
            self._blocks.append(kept)
            self.n_samples += len(kept)
        if self.max_samples is not None and self.n_samples > self.max_samples:   # keep the last max_samples rows
            drop = self.n_samples - self.max_samples
            while drop > 0:
                blk = self._blocks[0]
                if len(blk) <= drop:
                    self._blocks.pop(0)
                    drop -= len(blk)
                    if blk.is_cuda:
                        self._device_bytes -= blk.numel() * blk.element_size()
                else:
                    self._blocks[0] = blk[drop:]
                    drop = 0
            self.n_samples = self.max_samples

    def device_tensor(self) -> Optional[torch.Tensor]:
        """All stored iterations ``[n_samples, n, *event]`` on the GPU, or ``None`` if some block spilled to the host."""
        if not self._blocks or not all(b.is_cuda for b in self._blocks):
            return None
        if len(self._blocks) > 1:
            self._blocks = [torch.cat(self._blocks, dim=0)]
        return self._blocks[0]

    def as_tensor(self) -> torch.Tensor:
        """The host tensor ``[n_samples, n, *event]`` (what the reference's ``torch.stack(self._running)`` returns)."""
        if not self._blocks:
            raise RuntimeError("no samples stored")
        if len(self._blocks) == 1 and not self._blocks[0].is_cuda:
            return self._blocks[0]
        shape = (sum(len(b) for b in self._blocks), *self._blocks[0].shape[1:])
        nbytes = 4 * int(torch.Size(shape).numel())
        any_cuda = any(b.is_cuda for b in self._blocks)
        host = torch.empty(shape, dtype=torch.float32, pin_memory=bool(any_cuda and nbytes <= PINNED_HOST_LIMIT_BYTES))
        off = 0
        for b in self._blocks:
            host[off:off + len(b)].copy_(b, non_blocking=True)
            off += len(b)
        if any_cuda:
            torch.cuda.synchronize()
        self._blocks = [host]
        self._device_bytes = 0
        return host

    def reset(self):
        self._blocks = []
        self._device_bytes = 0
        self.n_samples = 0


@dataclass
class MCMCOutput:
    event_shape: Union[Tuple[int, ...], torch.Size]
    running_samples: MCMCSamples = None
    statistics: Optional[MCMCStatistics] = None
    kernel: Optional[MCMCKernel] = None
    store_samples: bool = True
    max_samples: int = None

    def __post_init__(self):
        if self.running_samples is None:
            self.running_samples = MCMCSamples(self.event_shape, store_samples=self.store_samples,
                                               max_samples=self.max_samples)
        if self.statistics is None:
            self.statistics = MCMCStatistics(self.event_shape)

    @property
    def samples(self) -> Union[torch.Tensor, None]:
        if not self.store_samples:
            return None
        return self.running_samples.as_tensor()

    def resample(self, n: int) -> torch.Tensor:
        flat = self.samples.flatten(0, 1)
        return flat[torch.randint(low=0, high=len(flat), size=(n,))]

    @property
    def mean(self):
        return self.statistics.running_first_moment

    @property
    def variance(self):
        return self.statistics.running_second_moment - self.statistics.running_first_moment ** 2

    @property
    def second_moment(self):
        return self.statistics.running_second_moment


class JumpNFMCOutput(MCMCOutput):
    def __init__(self, event_shape, *args, **kwargs):
        kwargs['statistics'] = JumpNFMCStatistics(event_shape)
        super().__init__(event_shape, *args, **kwargs)
