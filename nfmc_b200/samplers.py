"""Sampler objects with the reference's names and operator contract, driving the sm_100a kernels.

``Sampler.sample(x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput`` and ``.warmup(...)`` are the
contract of ``/root/reference/nfmc/algorithms/sampling/base.py:317-348``.  The loop bodies these classes
replace are ``mcmc/base.py:69-99`` (local loop), ``nfmc/jump.py:173-243`` (Jump NFMC), ``nfmc/imh.py:216-252``
(fixed IMH), ``nfmc/imh.py:122-178`` (adaptive IMH) and ``nfmc/neutra.py:116-129`` (NeuTra).  Chain state
lives on the GPU for the whole call; results come back as the reference's own record types with host tensors.

Counter formulas (checked against the reference in tests): MALA ``calls = grads = 2n`` per step
(langevin.py:116-120, ``n`` when unadjusted); HMC ``calls = 2Ln + 2n``, ``grads = 2Ln`` (hmc.py:122-125);
jump ``calls += 2n``, ``attempted_jumps += n`` (jump.py:214-216,236-239); fixed IMH ``calls += 2n``
(imh.py:243-247); adaptive IMH books them as gradient calls (imh.py:146, quirk Q3).
"""
from __future__ import annotations

import ctypes as C
import math
import time
from copy import deepcopy
from typing import Optional, Tuple

import torch

from . import _native as N
from . import external
from .flow import Flow
from .potentials import resolve_target
from .records import (DLMCKernel, DLMCParameters, TESSKernel, TESSParameters, ESSKernel, ESSParameters, MHKernel, MHParameters, HMCKernel, HMCParameters, IMHKernel, IMHParameters, JumpNFMCOutput, JumpNFMCParameters,
                      LangevinKernel, LangevinParameters, MCMCKernel, MCMCOutput, MCMCParameters, MetropolisKernel,
                      MetropolisParameters, NeuTraKernel, NeuTraParameters, NFMCKernel)

try:  # progress bars are optional
    from tqdm import tqdm
except Exception:  # pragma: no cover
    tqdm = None

MAX_DEVICE_SAMPLE_BYTES = 8 << 30  # device-side sample buffer per launch group


def _progress(it, desc, show):
    if show and tqdm is not None:
        return tqdm(it, desc=desc)
    return it


def draw_seed() -> int:
    """Philox seed taken from torch's global generator, so ``torch.manual_seed`` makes runs reproducible."""
    return int(torch.randint(0, 2 ** 62, (), dtype=torch.int64))


class DeviceSession:
    """Device-resident state of one ``sample()`` call: chains, statistics accumulators, RNG counters, timing."""

    def __init__(self, x0: torch.Tensor, event_shape, device=None, seed: Optional[int] = None, chain0: int = 0):
        self.device = N.require_cuda(device if device is not None else (x0.device if x0.is_cuda else None))
        self.event_shape = tuple(event_shape)
        self.d = int(math.prod(self.event_shape))
        self.n = int(x0.shape[0])
        self.x = N.dev_f32(x0, self.device).reshape(self.n, self.d).clone()
        self.moments = torch.zeros(2 * self.d, device=self.device, dtype=torch.float64)
        self.counts = torch.zeros(8, device=self.device, dtype=torch.int64)   # [0:4] local, [4:8] jump
        self.seed = draw_seed() if seed is None else int(seed)
        self.chain0 = int(chain0)
        self.local_step = 0   # global index of the next local step (Philox stream 0)
        self.flow_step = 0    # global index of the next flow draw (Philox stream 1)
        self.stream = N.stream_ptr(self.device)
        self._t0 = torch.cuda.Event(enable_timing=True)
        self._t1 = torch.cuda.Event(enable_timing=True)
        self._keep = []
        self._workspace = None

    def logq_scratch(self) -> torch.Tensor:
        """n floats of device scratch for the two-kernel NF jump (log q(x) between its kernels)."""
        if getattr(self, "_logq", None) is None:
            self._logq = torch.empty(self.n, device=self.device, dtype=torch.float32)
        return self._logq

    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._workspace is None or self._workspace.numel() < nbytes:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._workspace

    # -- descriptors ------------------------------------------------------------------------------------------
    def stats(self, jump: bool = False) -> N.StatsDesc:
        base = self.counts.data_ptr() + (32 if jump else 0)
        return N.StatsDesc(self.moments.data_ptr(), self.moments.data_ptr() + 8 * self.d, base)

    def sink(self, buf: Optional[torch.Tensor], seen0: int, thinning: int) -> Optional[N.SinkDesc]:
        if buf is None:
            return None
        return N.SinkDesc(buf.data_ptr(), seen0, thinning)

    def tic(self):
        self._t0.record(torch.cuda.current_stream(self.device))

    def toc(self) -> float:
        self._t1.record(torch.cuda.current_stream(self.device))
        self._t1.synchronize()
        return self._t0.elapsed_time(self._t1) * 1e-3

    def read_back(self):
        """(sum_x, sum_x2, counts) on the host -- one synchronising copy."""
        m = self.moments.cpu()
        c = self.counts.cpu()
        return m[: self.d], m[self.d:], [int(v) for v in c]


class _DeviceTuner:
    """Warm-up adaptation with the state on the device (reference: MetropolisSampler.update_kernel, mcmc/base.py:142-161;
    DualAveraging, tuning.py:15-41).  Per warm-up iteration: ``nfmc_chain_sums`` (sum x, sum x^2 per coordinate over this
    rank's chains, the accepted count, n), one all-reduce of those 2d+2 doubles when several ranks sample together
    (every rank ends up with the same inverse mass and the same step size), ``nfmc_tune_inv_mass`` (unbiased
    variance -> EMA into the device-resident inverse-mass diagonal) and one 16-byte read for the scalar dual-averaging
    update, which stays on the host as in the reference."""

    def __init__(self, ses: "DeviceSession", kernel: MetropolisKernel):
        self.sums = torch.empty(2 * ses.d + 2, dtype=torch.float64, device=ses.device)
        self.imd = N.dev_f32(kernel.inv_mass_diag, ses.device).clone()
        self.prev_acc = 0.0

    def update(self, ses: "DeviceSession", kernel: MetropolisKernel, params: MetropolisParameters, n_steps: int):
        import numpy as np
        import torch.distributed as dist
        d = ses.d
        N.check(N.lib().nfmc_chain_sums(N.ptr(ses.x), ses.n, d, N.ptr(self.sums), C.c_void_p(ses.counts.data_ptr()), ses.stream))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.sums)
        if params.tune_inv_mass_diag:
            N.check(N.lib().nfmc_tune_inv_mass(N.ptr(self.sums), d, float(params.imd_adjustment), N.ptr(self.imd), ses.stream))
            kernel.inv_mass_diag = self.imd                  # stays on the device until the warm-up ends
        if params.tune_step_size and params.adjustment:
            acc, n_total = (float(v) for v in self.sums[2 * d:].cpu())
            # the reference forms mean(mask.float()) and the error in fp32 (base.py:157-158)
            rate = np.float32(acc - self.prev_acc) / np.float32(n_total * n_steps)
            self.prev_acc = acc
            kernel.da.step(float(np.float32(kernel.da_params.target_acceptance_rate) - rate))
            kernel.step_size = kernel.da.value

    def finish(self, kernel: MetropolisKernel):
        if kernel.inv_mass_diag is self.imd:
            kernel.inv_mass_diag = self.imd.cpu()             # the records keep the reference's host tensor


def _require_analytic(target, who: str):
    """Samplers whose kernels differentiate through the flow need the target's gradient inside the kernel."""
    if getattr(target, "external", False):
        raise NotImplementedError(
            f"{who} needs a built-in analytic potential (nfmc_b200.potentials.*): its kernels evaluate the target inside the "
            f"flow sweep.  Callable targets are supported by mala / ula / hmc / uhmc / mh / random walk, their jump_* "
            f"variants, ess / jump_ess, imh, adaptive_imh, neutra_hmc and neutra_mh.")


def _imd_device(kernel: MetropolisKernel, device) -> Optional[torch.Tensor]:
    return None if kernel.has_unit_mass() else N.dev_f32(kernel.inv_mass_diag, device)


def _rows_kept(seen0: int, k: int, thinning: int) -> int:
    first = (seen0 + thinning - 1) // thinning
    last = (seen0 + k + thinning - 1) // thinning
    return last - first


def _device_scoped(fn):
    """Run a public sampler method with its device current: the C entry points size grids, create their side streams
    and launch on the CURRENT device, while tensors and the torch stream live on the sampler's device."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, x0, *args, **kwargs):
        dev = N.require_cuda(self.device if self.device is not None else (x0.device if torch.is_tensor(x0) and x0.is_cuda else None))
        if torch.is_tensor(x0) and x0.ndim >= 1 and x0.shape[0] == 0:
            # the reference divides by zero in its acceptance-rate bookkeeping on an empty batch (sampling/base.py:90-97)
            raise ValueError("x0 holds no chains (shape %s): nothing to sample" % (tuple(x0.shape),))
        with torch.cuda.device(dev):
            return fn(self, x0, *args, **kwargs)

    wrapper._device_scoped = True
    return wrapper


class Sampler:
    def __init__(self, event_shape, target, kernel: MCMCKernel, params: MCMCParameters):
        self.event_shape = tuple(event_shape)
        self.event_size = int(math.prod(self.event_shape))
        self.target = resolve_target(target, self.event_shape)
        self.kernel = kernel
        self.params = params
        self.seed: Optional[int] = None      # fixed Philox seed (None: drawn from torch's global generator per call)
        self.chain0: int = 0                 # global index of this process's first chain (multi-GPU sharding)
        self.device = None
        self._n_sessions = 0

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        for name in ("sample", "warmup"):
            fn = cls.__dict__.get(name)
            if fn is not None and not getattr(fn, "_device_scoped", False):
                setattr(cls, name, _device_scoped(fn))

    def session_seed(self) -> Optional[int]:
        """Philox seed of the next ``sample()`` / ``warmup()`` call.  Every call starts its step counters at 0, so a fixed
        ``self.seed`` is advanced by the call index (a run of warm-up then sampling, or a loop continuing from
        ``last_sample``, must not replay the same noise); ``None`` lets the session draw a fresh seed."""
        if self.seed is None:
            return None
        k = self._n_sessions
        self._n_sessions += 1
        return (int(self.seed) + k * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF

    @property
    def name(self):
        return "Generic sampler"

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        raise NotImplementedError

    def sample(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        raise NotImplementedError


# ---------------------------------------------------------------------------------------------------------------
# local samplers
# ---------------------------------------------------------------------------------------------------------------
class MetropolisSampler(Sampler):
    """Generic local loop (reference: MCMCSampler.sample, mcmc/base.py:56-102) over a fused K-step kernel."""

    #: how many steps one launch may fuse when nothing has to be observed in between
    max_fused_steps = 1 << 20

    def _launch(self, ses: DeviceSession, n_steps: int, sink, normals=None, uniforms=None):
        raise NotImplementedError

    def _calls_grads(self, n: int) -> Tuple[int, int]:
        raise NotImplementedError

    def begin_stage(self, ses: DeviceSession, init_normals=None):
        """Called once before a run of local steps (a ``sample()`` call, or each local stage of Jump NFMC).  ESS
        restarts from the prior here (mcmc/ess.py:126); every other kernel continues from the session state."""

    def run_steps(self, ses: DeviceSession, out: MCMCOutput, n_steps: int, store: bool, normals=None, uniforms=None):
        """Advance all chains ``n_steps`` local steps; book rows / counters into ``out``.  Returns device rows or None."""
        rs = out.running_samples
        buf = None
        sink = None
        if store:
            rows = _rows_kept(rs.seen_samples, n_steps, rs.thinning)
            nbytes = rows * ses.n * ses.d * 4
            if nbytes > MAX_DEVICE_SAMPLE_BYTES:
                raise MemoryError(f"storing {rows} x {ses.n} x {ses.d} samples needs {nbytes / 2**30:.1f} GiB on the "
                                  f"device; use params.store_samples=False (moments and last_sample are still returned)")
            buf = torch.empty(rows, ses.n, ses.d, device=ses.device, dtype=torch.float32)
            sink = ses.sink(buf, rs.seen_samples, rs.thinning)
        self._launch(ses, n_steps, sink, normals, uniforms)
        ses.local_step += n_steps
        calls, grads = self._calls_grads(ses.n)
        out.statistics.update_counters(n_target_calls=calls * n_steps, n_target_gradient_calls=grads * n_steps)
        return buf

    def sample(self, x0: torch.Tensor, show_progress: bool = True, time_limit_seconds=None, normals=None, uniforms=None) -> MCMCOutput:
        """``normals [T,n,d]`` / ``uniforms [T,n]`` optionally inject the random numbers (parity tests)."""
        event_shape = tuple(x0.shape[1:])
        out = MCMCOutput(event_shape, store_samples=self.params.store_samples)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        self.begin_stage(ses)
        T = int(self.params.n_iterations)
        tuning = bool(self.params.tuning) and self.tunable
        chunk = 1 if (tuning or time_limit_seconds is not None or show_progress) else min(T, self.max_fused_steps)
        done = 0
        label = f'{self.name} (tuning)' if self.params.tuning else self.name
        bar = _progress(range(0, T, max(chunk, 1)), label, show_progress)
        tuner = _DeviceTuner(ses, self.kernel) if tuning else None
        for start in bar:
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds > time_limit_seconds:
                break
            k = min(chunk, T - start)
            nz = None if normals is None else N.dev_f32(normals[start:start + k], ses.device)
            un = None if uniforms is None else N.dev_f32(uniforms[start:start + k], ses.device)
            ses.tic()
            buf = self.run_steps(ses, out, k, self.params.store_samples, nz, un)
            dt = ses.toc()
            out.statistics.update_elapsed_time(dt)
            if buf is not None:
                out.running_samples.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=k)
            done += k
            if tuner is not None:                                # mcmc/base.py:92-96
                tuner.update(ses, self.kernel, self.params, k)
        if tuner is not None:
            tuner.finish(self.kernel)
        self._finish(ses, out, done)
        out.kernel = self.kernel
        return out

    #: ESS has nothing to tune (mcmc/ess.py:118-119)
    tunable = True

    def _finish(self, ses: DeviceSession, out: MCMCOutput, steps_done: int):
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * steps_done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1])
        out.statistics.n_nonfinite = cnt[2]
        out.running_samples.set_last_device(ses.x.reshape(ses.n, *out.event_shape))

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        """Reference: MCMCSampler.warmup (mcmc/base.py:39-54): tune on a copy, adopt its kernel."""
        cp = deepcopy(self)
        cp.params.tuning_mode()
        cp.params.n_iterations = self.params.n_warmup_iterations
        out = cp.sample(x0, show_progress=show_progress, time_limit_seconds=time_limit_seconds)
        self.kernel = cp.kernel
        new_params = cp.params
        new_params.n_iterations = self.params.n_iterations
        self.params = new_params
        self.params.sampling_mode()
        return out


class Langevin(MetropolisSampler):
    def __init__(self, event_shape, target, kernel: Optional[LangevinKernel] = None,
                 params: Optional[LangevinParameters] = None):
        es = int(math.prod(tuple(event_shape)))
        super().__init__(event_shape, target, kernel or LangevinKernel(event_size=es), params or LangevinParameters())

    @property
    def name(self):
        return 'LMC'

    def _calls_grads(self, n):
        return (2 * n, 2 * n) if self.params.adjustment else (n, n)

    def _launch(self, ses, n_steps, sink, normals=None, uniforms=None):
        if self.target.external:                                  # callable target: autograd + the nfmc_ext_* kernels
            return external.langevin_steps(self.target, ses, n_steps, float(self.kernel.step_size),
                                           _imd_device(self.kernel, ses.device), bool(self.params.adjustment), False, sink,
                                           normals, uniforms)
        pot, keep = self.target.descriptor(ses.device)
        imd = _imd_device(self.kernel, ses.device)
        rng = N.rng_desc(ses.seed, ses.local_step, normals, uniforms)
        st = ses.stats()
        N.check(N.lib().nfmc_mala_steps(C.byref(pot), N.ptr(ses.x), ses.n, n_steps, float(self.kernel.step_size),
                                        N.ptr(imd), int(bool(self.params.adjustment)), C.byref(rng), ses.chain0,
                                        C.byref(st), None if sink is None else C.byref(sink), ses.stream))


class MALA(Langevin):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params.adjustment = True


class ULA(Langevin):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params.adjustment = False


class HMC(MetropolisSampler):
    def __init__(self, event_shape, target, kernel: Optional[HMCKernel] = None, params: Optional[HMCParameters] = None):
        es = int(math.prod(tuple(event_shape)))
        super().__init__(event_shape, target, kernel or HMCKernel(event_size=es), params or HMCParameters())

    @property
    def name(self):
        return 'HMC'

    def _calls_grads(self, n):
        L = int(self.kernel.n_leapfrog_steps)
        return (2 * L * n + (2 * n if self.params.adjustment else 0), 2 * L * n)

    def _launch(self, ses, n_steps, sink, normals=None, uniforms=None):
        if self.target.external:
            return external.hmc_steps(self.target, ses, n_steps, float(self.kernel.step_size),
                                      int(self.kernel.n_leapfrog_steps), _imd_device(self.kernel, ses.device),
                                      bool(self.params.adjustment), sink, normals, uniforms)
        pot, keep = self.target.descriptor(ses.device)
        imd = _imd_device(self.kernel, ses.device)
        rng = N.rng_desc(ses.seed, ses.local_step, normals, uniforms)
        st = ses.stats()
        N.check(N.lib().nfmc_hmc_steps(C.byref(pot), N.ptr(ses.x), ses.n, n_steps, float(self.kernel.step_size),
                                       int(self.kernel.n_leapfrog_steps), N.ptr(imd), int(bool(self.params.adjustment)),
                                       C.byref(rng), ses.chain0, C.byref(st),
                                       None if sink is None else C.byref(sink), ses.stream))


class MH(MetropolisSampler):
    """Random-walk Metropolis (reference: mcmc/mh.py:27-73): x' = x + inv_mass_diag * xi, accept iff log u < U(x) - U(x')."""

    def __init__(self, event_shape, target, kernel: Optional[MHKernel] = None, params: Optional[MHParameters] = None):
        es = int(math.prod(tuple(event_shape)))
        super().__init__(event_shape, target, kernel or MHKernel(event_size=es), params or MHParameters())

    @property
    def name(self):
        return 'MH'

    def _calls_grads(self, n):
        return ((2 * n) if self.params.adjustment else 0, 0)                        # mh.py:68-71

    def _launch(self, ses, n_steps, sink, normals=None, uniforms=None):
        if self.target.external:
            return external.langevin_steps(self.target, ses, n_steps, 1.0, _imd_device(self.kernel, ses.device),
                                           bool(self.params.adjustment), True, sink, normals, uniforms)
        pot, keep = self.target.descriptor(ses.device)
        imd = _imd_device(self.kernel, ses.device)
        rng = N.rng_desc(ses.seed, ses.local_step, normals, uniforms)
        st = ses.stats()
        N.check(N.lib().nfmc_mh_steps(C.byref(pot), N.ptr(ses.x), ses.n, n_steps, N.ptr(imd),
                                      int(bool(self.params.adjustment)), C.byref(rng), ses.chain0, C.byref(st),
                                      None if sink is None else C.byref(sink), ses.stream))


class ESS(MetropolisSampler):
    """Elliptical slice sampling with prior N(0, I) (reference: mcmc/ess.py:78-127).  ``negative_log_likelihood`` is the
    potential the slice is taken on; ``target`` is carried for the API only (the reference never evaluates it either).
    As in the reference, ``x0`` only supplies the number of chains: the run starts from a fresh prior draw
    (ess.py:126), and every step counts as accepted (ess.py:107)."""

    def __init__(self, event_shape, target, negative_log_likelihood, kernel: Optional[ESSKernel] = None,
                 params: Optional[ESSParameters] = None):
        super().__init__(event_shape, target, kernel or ESSKernel(tuple(event_shape)), params or ESSParameters())
        self.negative_log_likelihood = resolve_target(negative_log_likelihood, self.event_shape)

    @property
    def name(self):
        return 'ESS'

    def _calls_grads(self, n):
        return ((int(self.params.max_ess_step_iterations) + 1) * n, 0)            # ess.py:114-115

    tunable = False                                                                  # ess.py:118-119: nothing to tune

    def begin_stage(self, ses: DeviceSession, init_normals=None):
        if init_normals is not None:
            ses.x.copy_(N.dev_f32(init_normals, ses.device).reshape(ses.n, ses.d))
            return
        # Philox stream 3 = prior restarts, keyed by the index of the next local step
        rng = N.rng_desc(ses.seed, ses.local_step)
        N.check(N.lib().nfmc_rng_fill(C.byref(rng), 3, ses.chain0, ses.d, ses.n, 1, N.ptr(ses.x), None, ses.stream))

    def _launch(self, ses, n_steps, sink, normals=None, uniforms=None):
        if self.negative_log_likelihood.external:                 # callable likelihood: bracket rounds around its evaluation
            return external.ess_steps(self.negative_log_likelihood, ses, n_steps, int(self.params.max_ess_step_iterations), sink,
                                      normals, uniforms)
        pot, keep = self.negative_log_likelihood.descriptor(ses.device)
        rng = N.rng_desc(ses.seed, ses.local_step, normals, uniforms)
        st = ses.stats()
        N.check(N.lib().nfmc_ess_steps(C.byref(pot), N.ptr(ses.x), ses.n, n_steps,
                                       int(self.params.max_ess_step_iterations), C.byref(rng), ses.chain0, C.byref(st),
                                       None if sink is None else C.byref(sink), ses.stream))


class RandomWalk(MH):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params.adjustment = False


class UHMC(HMC):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.params.adjustment = False


# ---------------------------------------------------------------------------------------------------------------
# Jump NFMC
# ---------------------------------------------------------------------------------------------------------------
class JumpNFMC(Sampler):
    """K local steps, then one flow-proposal MH jump for every chain, T times (reference: nfmc/jump.py:156-246)."""

    def __init__(self, event_shape, target, inner_sampler: MetropolisSampler, kernel: NFMCKernel = None,
                 params: JumpNFMCParameters = None):
        super().__init__(event_shape, target, kernel or NFMCKernel(tuple(event_shape)), params or JumpNFMCParameters())
        self.inner_sampler = inner_sampler

    @property
    def name(self):
        return 'Jump MCMC'

    def _fused_inner_kind(self):
        """inner_kind of nfmc_jump_sample_device for this inner sampler, or None if it has no fused whole-run path."""
        inner = self.inner_sampler
        if isinstance(inner, Langevin):
            return 0
        if isinstance(inner, HMC):
            return 1
        if isinstance(inner, MH):
            return 2
        return None

    def jump(self, ses: DeviceSession, sink=None, z=None, uniforms=None):
        flow: Flow = self.kernel.flow
        if self.target.external:                                  # callable target: U(x), U(x') by calling it
            external.jump_step(self.target, flow, ses, bool(self.params.adjusted_jumps), sink, z, uniforms)
            ses.flow_step += 1
            return
        pot, keep = self.target.descriptor(ses.device)
        rng = N.rng_desc(ses.seed, ses.flow_step, z, uniforms)
        st = ses.stats(jump=True)
        sk = None if sink is None else C.byref(sink)
        if flow.bijection.uses_tensor_cores():                    # wide conditioner: flow passes on tcgen05
            fd, keep2 = flow.bijection.tc_descriptor(ses.device)
            nb = N.lib().nfmc_jump_tc_workspace_bytes(ses.d, ses.n)
            ws = ses.workspace(nb)
            N.check(N.lib().nfmc_jump_step_tc(C.byref(pot), C.byref(fd), N.ptr(ses.x), None, 1, ses.n,
                                              int(bool(self.params.adjusted_jumps)), C.byref(rng), ses.chain0,
                                              C.byref(st), sk, N.ptr(ws), nb, ses.stream))
        elif flow.bijection.uses_row_tile_pass():                 # deep / odd-sized conditioner: row-tile fp32 passes
            fd, keep2 = flow.bijection.theta_descriptor(ses.device)
            nb = N.lib().nfmc_jump_tc_workspace_bytes(ses.d, ses.n)
            ws = ses.workspace(nb)
            N.check(N.lib().nfmc_jump_step_wide(C.byref(pot), C.byref(fd), 1, N.ptr(ses.x), None, 1, ses.n,
                                                int(bool(self.params.adjusted_jumps)), C.byref(rng), ses.chain0,
                                                C.byref(st), sk, N.ptr(ws), nb, ses.stream))
        else:
            fd, keep2 = flow.bijection.descriptor(ses.device)
            # two kernels (forward pass for log q(x), then proposal + accept): faster than the fused jump kernel
            N.check(N.lib().nfmc_jump_step2(C.byref(pot), C.byref(fd), N.ptr(ses.x), N.ptr(ses.logq_scratch()), ses.n,
                                            int(bool(self.params.adjusted_jumps)), C.byref(rng), ses.chain0, C.byref(st),
                                            sk, ses.stream))
        ses.flow_step += 1

    def sample(self, x0: torch.Tensor, show_progress: bool = True, time_limit_seconds=None,
               normals=None, uniforms=None, jump_z=None, jump_uniforms=None, stage_normals=None) -> MCMCOutput:
        """``normals [T,K,n,d]`` / ``uniforms [T,K,n]`` / ``jump_z [T,n,d]`` / ``jump_uniforms [T,n]`` optionally inject
        the random numbers (parity tests); otherwise they come from the Philox generator.  ``stage_normals [T,n,d]``:
        the prior draw each ESS local stage restarts from (jump_ess only)."""
        p: JumpNFMCParameters = self.params
        inner = self.inner_sampler
        if not inner.params.store_samples:
            raise ValueError("Inner sampler in jump HMC must store samples")     # reference: jump.py:163-164
        event_shape = tuple(x0.shape[1:])
        out = JumpNFMCOutput(event_shape=event_shape, store_samples=p.store_samples)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        K = int(inner.params.n_iterations)
        T = int(p.n_iterations)
        store = bool(p.store_samples)
        rs = out.running_samples
        done = 0
        dev = ses.device
        kind = self._fused_inner_kind()
        if (kind is not None and T > 0 and not store and not p.fit_nf and time_limit_seconds is None and not show_progress
                and not self.target.external and not inner.target.external
                and all(v is None for v in (normals, uniforms, jump_z, jump_uniforms, stage_normals))
                and not self.kernel.flow.bijection.uses_tensor_cores() and not self.kernel.flow.bijection.uses_row_tile_pass()):
            # whole run in one call (nfmc_jump_sample_device): slabs of chains pipelined over several streams so that the
            # jump kernel of one slab overlaps the local kernel of another; same Philox steps as the loop below
            pot, keep = self.target.descriptor(dev)
            fd, keep2 = self.kernel.flow.bijection.descriptor(dev)
            imd = _imd_device(inner.kernel, dev)
            st_l, st_j = ses.stats(), ses.stats(jump=True)
            ses.tic()
            N.check(N.lib().nfmc_jump_sample_device(
                C.byref(pot), C.byref(fd), N.ptr(ses.x), ses.n, kind, T, K, float(inner.kernel.step_size),
                int(getattr(inner.kernel, 'n_leapfrog_steps', 0)), N.ptr(imd), int(bool(inner.params.adjustment)),
                int(bool(p.adjusted_jumps)), ses.seed & 0xFFFFFFFFFFFFFFFF, ses.local_step, ses.flow_step, ses.chain0,
                C.byref(st_l), C.byref(st_j), N.ptr(ses.logq_scratch()), ses.stream))
            out.statistics.update_elapsed_time(ses.toc())
            ses.local_step += T * K
            ses.flow_step += T
            calls, grads = inner._calls_grads(ses.n)
            out.statistics.update_counters(n_target_calls=calls * K * T + (2 * ses.n * T if p.adjusted_jumps else 0),
                                           n_target_gradient_calls=grads * K * T)
            done = T
            T = 0                                                                   # nothing left for the loop below
        for i in _progress(range(T), 'Jump MCMC', show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds >= time_limit_seconds:
                break
            ses.tic()
            nz = None if normals is None else N.dev_f32(normals[i], dev)
            un = None if uniforms is None else N.dev_f32(uniforms[i], dev)
            need_block = store or (p.fit_nf and i >= p.n_jumps_before_training and K * ses.n * ses.d * 4 <= (2 << 30))
            inner.begin_stage(ses, None if stage_normals is None else stage_normals[i])
            buf = inner.run_steps(ses, out, K, need_block, nz, un)                  # jump.py:178-189
            if p.fit_nf and i >= p.n_jumps_before_training:                         # jump.py:193-201
                from .flow_train import train_val_split
                # training rows: the inner (step, chain) block when it fits on the device, else the current states
                pool = buf if buf is not None else ses.x[None]
                x_train, x_val = train_val_split(pool, p.train_pct, p.max_train_size, p.max_val_size)
                self.kernel.flow.fit(x_train=x_train, x_val=x_val, **p.flow_fit_kwargs)
            if not store:
                buf = None
            jbuf = None
            jsink = None
            if store:
                seen = rs.seen_samples + K
                if _rows_kept(seen, 1, rs.thinning):
                    jbuf = torch.empty(1, ses.n, ses.d, device=dev, dtype=torch.float32)
                    jsink = ses.sink(jbuf, seen, rs.thinning)
            jz = None if jump_z is None else N.dev_f32(jump_z[i], dev).reshape(1, ses.n, ses.d)
            ju = None if jump_uniforms is None else N.dev_f32(jump_uniforms[i], dev).reshape(1, ses.n)
            self.jump(ses, jsink, jz, ju)                                            # jump.py:203-243
            if p.adjusted_jumps:
                out.statistics.update_counters(n_target_calls=2 * ses.n)             # jump.py:214-216
            out.statistics.update_elapsed_time(ses.toc())
            if store:
                rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=K)
                if jbuf is not None:
                    rs.add(jbuf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=1)
                else:
                    rs.seen_samples += 1
            done += 1
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done * (K + 1))
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1],
                                       n_accepted_jumps=cnt[4], n_attempted_jumps=cnt[5])
        out.statistics.n_nonfinite = cnt[2] + cnt[6]
        rs.set_last_device(ses.x.reshape(ses.n, *event_shape))
        out.kernel = self.kernel
        return out

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        """Reference: JumpNFMC.warmup (jump.py:104-154): tune the inner sampler (70 % of the budget), then fit the flow
        to the warm-up samples with weight rollback if the fit diverges."""
        from .flow_train import train_val_split
        limit = None if time_limit_seconds is None else 0.7 * time_limit_seconds
        t0 = time.time()
        self.inner_sampler.params.store_samples = True
        out = self.inner_sampler.warmup(x0, show_progress=show_progress, time_limit_seconds=limit)
        p: JumpNFMCParameters = self.params
        pool = out.running_samples.device_tensor()          # warm-up samples stay on the GPU when they fit
        x_train, x_val = train_val_split(pool if pool is not None else out.samples, p.train_pct, p.max_train_size,
                                         p.max_val_size)
        backup = deepcopy(self.kernel.flow.state_dict())
        fit_limit = None if time_limit_seconds is None else max(time_limit_seconds - (time.time() - t0), 0.0)
        try:
            self.kernel.flow.fit(x_train=x_train, x_val=x_val,
                                 **{**p.flow_fit_kwargs, "show_progress": show_progress, "time_limit_seconds": fit_limit})
        except ValueError:
            self.kernel.flow.load_state_dict(backup)
        return out


def _make_inner(cls, event_shape, target, kernel, params):
    return cls(event_shape, target, kernel, params)


class JumpMALA(JumpNFMC):
    def __init__(self, event_shape, target, kernel=None, params=None, inner_kernel=None, inner_params=None):
        super().__init__(event_shape, target, MALA(event_shape, target, inner_kernel, inner_params), kernel, params)


class JumpULA(JumpNFMC):
    def __init__(self, event_shape, target, kernel=None, params=None, inner_kernel=None, inner_params=None):
        super().__init__(event_shape, target, ULA(event_shape, target, inner_kernel, inner_params), kernel, params)


class JumpMH(JumpNFMC):
    def __init__(self, event_shape, target, kernel=None, params=None, inner_kernel=None, inner_params=None):
        super().__init__(event_shape, target, MH(event_shape, target, inner_kernel, inner_params), kernel, params)


class JumpESS(JumpNFMC):
    """Reference: jump.py:309-319.  Note the reference quirk kept here: the ESS local stage ignores the state left by
    the jump and restarts from the prior every outer iteration (ess.py:126)."""

    def __init__(self, event_shape, target, negative_log_likelihood, kernel=None, params=None, inner_kernel=None,
                 inner_params=None):
        super().__init__(event_shape, target, ESS(event_shape, target, negative_log_likelihood, inner_kernel, inner_params),
                         kernel, params)


class JumpHMC(JumpNFMC):
    def __init__(self, event_shape, target, kernel=None, params=None, inner_kernel=None, inner_params=None):
        super().__init__(event_shape, target, HMC(event_shape, target, inner_kernel, inner_params), kernel, params)


class JumpUHMC(JumpNFMC):
    def __init__(self, event_shape, target, kernel=None, params=None, inner_kernel=None, inner_params=None):
        super().__init__(event_shape, target, UHMC(event_shape, target, inner_kernel, inner_params), kernel, params)


# ---------------------------------------------------------------------------------------------------------------
# independence Metropolis-Hastings
# ---------------------------------------------------------------------------------------------------------------
class AbstractIMH(Sampler):
    recompute_logq = False

    def __init__(self, event_shape, target, kernel: Optional[IMHKernel] = None, params: Optional[IMHParameters] = None):
        super().__init__(event_shape, target, kernel or IMHKernel(tuple(event_shape)), params or IMHParameters())

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        """Reference: AbstractIMH.warmup (imh.py:60-75): variational fit of the flow to the target, then the initial
        state is a draw from the fitted flow."""
        self.kernel.flow.variational_fit(self.target.log_prob_fn(), **self.params.warmup_fit_kwargs,
                                         show_progress=show_progress, time_limit_seconds=time_limit_seconds)
        out = MCMCOutput(event_shape=tuple(x0.shape[1:]), store_samples=self.params.store_samples)
        out.running_samples.add(self.kernel.flow.sample(x0.shape[0]))
        return out

    def _run(self, x0, show_progress, time_limit_seconds, store, z=None, uniforms=None, after_iteration=None) -> MCMCOutput:
        event_shape = tuple(x0.shape[1:])
        out = MCMCOutput(event_shape=event_shape, store_samples=store)
        flow: Flow = self.kernel.flow
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        dev = ses.device
        T = int(self.params.n_iterations)
        if self.target.external:
            return self._run_external(ses, out, flow, T, show_progress, time_limit_seconds, store, z, uniforms, after_iteration)
        pot, keep = self.target.descriptor(dev)
        def describe():
            """(tensor cores?, row-tile fp32 pass?, descriptor, tensor it points into) under the flow's current parameters"""
            bij = flow.bijection
            if bij.uses_tensor_cores():
                return (True, False) + bij.tc_descriptor(dev)
            if bij.uses_row_tile_pass():
                return (False, True) + bij.theta_descriptor(dev)
            return (False, False) + bij.descriptor(dev)
        tc, wide, fd, keep2 = describe()
        logq = torch.empty(ses.n, device=dev, dtype=torch.float32)
        ses.tic()
        if not self.recompute_logq:                                                  # imh.py:214
            if tc:
                N.check(N.lib().nfmc_flow_tc_pass(C.byref(fd), 2, N.ptr(ses.x), None, N.ptr(logq), ses.n, ses.stream))
            elif wide:
                N.check(N.lib().nfmc_flow_wide_log_prob(fd.d, fd.n_coupling, fd.n_linear, fd.hidden, N.ptr(keep2), 1, N.ptr(ses.x),
                                                        N.ptr(logq), ses.n, ses.stream))
            else:
                N.check(N.lib().nfmc_flow_log_prob(C.byref(fd), N.ptr(ses.x), N.ptr(logq), ses.n, ses.stream))
        out.statistics.update_elapsed_time(ses.toc())
        chunk = 1 if (tc or wide or after_iteration is not None or time_limit_seconds is not None or show_progress) else T
        rs = out.running_samples
        done = 0
        for start in _progress(range(0, T, max(chunk, 1)), self.name, show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds >= time_limit_seconds:
                break
            k = min(chunk, T - start)
            buf, sink = None, None
            if store:
                rows = _rows_kept(rs.seen_samples, k, rs.thinning)
                if rows * ses.n * ses.d * 4 > MAX_DEVICE_SAMPLE_BYTES:
                    raise MemoryError("sample buffer too large for the device; use store_samples=False")
                buf = torch.empty(rows, ses.n, ses.d, device=dev, dtype=torch.float32)
                sink = ses.sink(buf, rs.seen_samples, rs.thinning)
            zz = None if z is None else N.dev_f32(z[start:start + k], dev)
            uu = None if uniforms is None else N.dev_f32(uniforms[start:start + k], dev)
            rng = N.rng_desc(ses.seed, ses.flow_step, zz, uu)
            st = ses.stats()
            ses.tic()
            if tc or wide:
                nb = N.lib().nfmc_jump_tc_workspace_bytes(ses.d, ses.n)
                ws = ses.workspace(nb)
                tail = (N.ptr(ses.x), N.ptr(logq), int(self.recompute_logq), ses.n, 1, C.byref(rng), ses.chain0, C.byref(st),
                        None if sink is None else C.byref(sink), N.ptr(ws), nb, ses.stream)
                if tc:
                    N.check(N.lib().nfmc_jump_step_tc(C.byref(pot), C.byref(fd), *tail))
                else:
                    N.check(N.lib().nfmc_jump_step_wide(C.byref(pot), C.byref(fd), 1, *tail))      # 1: theta packed transposed
            else:
                N.check(N.lib().nfmc_imh_steps(C.byref(pot), C.byref(fd), N.ptr(ses.x), N.ptr(logq), ses.n, k,
                                               int(self.recompute_logq), C.byref(rng), ses.chain0, C.byref(st),
                                               None if sink is None else C.byref(sink), ses.stream))
            out.statistics.update_elapsed_time(ses.toc())
            ses.flow_step += k
            done += k
            if buf is not None:
                rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=k)
            if after_iteration is not None:
                after_iteration(start, out)
                tc, wide, fd, keep2 = describe()                                      # parameters changed: re-pack
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1])
        if self.recompute_logq:
            out.statistics.update_counters(n_target_gradient_calls=2 * ses.n * done)  # imh.py:146 (quirk Q3)
        else:
            out.statistics.update_counters(n_target_calls=2 * ses.n * done)           # imh.py:243-247
        rs.set_last_device(ses.x.reshape(ses.n, *event_shape))
        out.kernel = self.kernel
        return out


    def _run_external(self, ses, out, flow, T, show_progress, time_limit_seconds, store, z, uniforms, after_iteration):
        """The same loop for a callable target: one ``external.jump_step`` per iteration; log q and U travel with the state."""
        dev = ses.device
        event_shape = out.event_shape
        rs = out.running_samples
        ses.tic()
        logq = flow.log_prob(ses.x.reshape(ses.n, *event_shape)).reshape(ses.n).contiguous()          # imh.py:214
        u_cache = self.target.value(ses.x)
        out.statistics.update_elapsed_time(ses.toc())
        done = 0
        for i in _progress(range(T), self.name, show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds >= time_limit_seconds:
                break
            buf, sink = None, None
            if store and _rows_kept(rs.seen_samples, 1, rs.thinning):
                buf = torch.empty(1, ses.n, ses.d, device=dev, dtype=torch.float32)
                sink = ses.sink(buf, rs.seen_samples, rs.thinning)
            zz = None if z is None else N.dev_f32(z[i], dev)
            uu = None if uniforms is None else N.dev_f32(uniforms[i], dev)
            ses.tic()
            external.jump_step(self.target, flow, ses, True, sink, zz, uu, logq=logq, recompute_logq=self.recompute_logq,
                               jump_stats=False, u_cache=u_cache)
            out.statistics.update_elapsed_time(ses.toc())
            ses.flow_step += 1
            done += 1
            if buf is not None:
                rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=1)
            elif store:
                rs.seen_samples += 1
            if after_iteration is not None:
                after_iteration(i, out)
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1])
        if self.recompute_logq:
            out.statistics.update_counters(n_target_gradient_calls=2 * ses.n * done)  # imh.py:146 (quirk Q3)
        else:
            out.statistics.update_counters(n_target_calls=2 * ses.n * done)           # imh.py:243-247
        rs.set_last_device(ses.x.reshape(ses.n, *event_shape))
        out.kernel = self.kernel
        return out


class FixedIMH(AbstractIMH):
    @property
    def name(self):
        return "Fixed IMH"

    def sample(self, x0, show_progress=True, time_limit_seconds=None, z=None, uniforms=None) -> MCMCOutput:
        return self._run(x0, show_progress, time_limit_seconds, self.params.store_samples, z, uniforms)


class AdaptiveIMH(AbstractIMH):
    """Adaptive IMH (reference: imh.py:78-181).  The MH part follows imh.py:122-150: log q(x) is recomputed every
    iteration and samples are always stored (quirk Q2).  With ``adapt=True`` (default) the flow is refitted for one
    epoch on a randomly chosen stored iteration with probability ``adaptation_dropoff**i`` (imh.py:152-175, rollback on
    ``ValueError``); that needs one launch per iteration.  ``adapt=False`` fuses all iterations into one launch."""
    recompute_logq = True
    adapt = True

    @property
    def name(self):
        return "Adaptive IMH"

    def sample(self, x0, show_progress=True, time_limit_seconds=None, z=None, uniforms=None) -> MCMCOutput:
        if not self.adapt:
            return self._run(x0, show_progress, time_limit_seconds, True, z, uniforms)
        return self._run(x0, show_progress, time_limit_seconds, True, z, uniforms, after_iteration=self._maybe_refit)

    def _maybe_refit(self, i: int, out: MCMCOutput):
        p: IMHParameters = self.params
        if float(torch.rand(())) >= p.adaptation_dropoff ** i:                       # imh.py:152-154
            return
        n_samples = out.running_samples.n_samples
        if n_samples == 0:
            return
        if p.train_distribution == 'uniform':
            k = int(torch.randint(0, n_samples, ()))
        elif p.train_distribution == 'bounded_geom_approx':
            k = int(torch.randint(max(0, n_samples - 100), n_samples, ()))
        else:  # bounded_geom (imh.py:39-45)
            v = torch.arange(0, n_samples)
            pdf = 0.025 * (1 - 0.025) ** (n_samples - 1 - v) / (1 - (1 - 0.025) ** n_samples)
            k = int(torch.searchsorted(torch.cumsum(pdf, 0), float(torch.rand(())), right=True).clamp(max=n_samples - 1))
        x_train = out.running_samples[k]
        backup = deepcopy(self.kernel.flow.state_dict())
        try:
            self.kernel.flow.fit(x_train, n_epochs=1, show_progress=False)            # imh.py:171-175
        except ValueError:
            self.kernel.flow.load_state_dict(backup)


# ---------------------------------------------------------------------------------------------------------------
# NeuTra
# ---------------------------------------------------------------------------------------------------------------
class NeuTraHMC(Sampler):
    """HMC in the flow's latent space on U~(z) = U(T^-1 z) - log|det dT^-1/dz| (reference: nfmc/neutra.py:36-144).
    Samples and moments are latent-space quantities, exactly as the reference returns them (quirk Q1)."""

    def __init__(self, event_shape, target, inner_kernel: HMCKernel = None, inner_params: HMCParameters = None,
                 kernel: NeuTraKernel = None, params: NeuTraParameters = None):
        es = int(math.prod(tuple(event_shape)))
        super().__init__(event_shape, target, kernel or NeuTraKernel(tuple(event_shape)), params or NeuTraParameters())
        self.inner_kernel = inner_kernel or HMCKernel(event_size=es)
        self.inner_params = inner_params or HMCParameters()
        self.inner_params.n_iterations = self.params.n_iterations

    @property
    def name(self):
        return "NeuTra HMC"

    def sample(self, x0, show_progress=True, time_limit_seconds=None, normals=None, uniforms=None) -> MCMCOutput:
        return self._run(x0, int(self.params.n_iterations), False, show_progress, time_limit_seconds, normals, uniforms)

    def _run(self, x0, T, tuning, show_progress=True, time_limit_seconds=None, normals=None, uniforms=None) -> MCMCOutput:
        event_shape = tuple(x0.shape[1:])
        store = bool(self.params.store_samples)
        out = MCMCOutput(event_shape, store_samples=store)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        dev = ses.device
        bij = self.kernel.flow.bijection
        # callable target, or a conditioner shape only the row-tile fp32 kernels cover (deep, odd d, ...): the latent step is
        # composed from the inverse pass, U / grad U at x = T^-1 z, the backward sweep and the nfmc_ext_* kernels
        row_tile = bij.row_tile_supported() and not bij.uses_tensor_cores_for_neutra(ses.n)
        ext = self.target.external or row_tile
        latent = external.LatentTarget(self.target, self.kernel.flow, row_tile=row_tile if row_tile else None) if ext else None
        pot, keep = (None, None) if ext else self.target.descriptor(dev)
        fd, keep2 = self.kernel.flow.bijection.descriptor(dev)
        imd = _imd_device(self.inner_kernel, dev)
        chunk = 1 if (tuning or time_limit_seconds is not None or show_progress) else T
        tuner = None
        rs = out.running_samples
        done = 0
        for start in _progress(range(0, T, max(chunk, 1)), self.name, show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds > time_limit_seconds:
                break
            k = min(chunk, T - start)
            buf, sink = None, None
            if store:
                rows = _rows_kept(rs.seen_samples, k, rs.thinning)
                if rows * ses.n * ses.d * 4 > MAX_DEVICE_SAMPLE_BYTES:
                    raise MemoryError("sample buffer too large for the device; use store_samples=False")
                buf = torch.empty(rows, ses.n, ses.d, device=dev, dtype=torch.float32)
                sink = ses.sink(buf, rs.seen_samples, rs.thinning)
            nz = None if normals is None else N.dev_f32(normals[start:start + k], dev)
            un = None if uniforms is None else N.dev_f32(uniforms[start:start + k], dev)
            rng = N.rng_desc(ses.seed, ses.local_step, nz, un)
            st = ses.stats()
            ses.tic()
            if ext:
                self._launch_latent_external(latent, ses, k, imd, sink, nz, un)
            else:
                self._launch_latent(ses, pot, fd, k, imd, rng, st, sink)
            out.statistics.update_elapsed_time(ses.toc())
            ses.local_step += k
            done += k
            if buf is not None:
                rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=k)
            if tuning:                                                                # mcmc/base.py:142-161 on the latent chain
                if tuner is None:
                    tuner = _DeviceTuner(ses, self.inner_kernel)
                tuner.update(ses, self.inner_kernel, self.inner_params, k)
                imd = _imd_device(self.inner_kernel, dev)
        if tuner is not None:
            tuner.finish(self.inner_kernel)
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done)
        calls, grads = self._calls_grads(ses.n)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1],
                                       n_target_calls=calls * done, n_target_gradient_calls=grads * done)
        out.statistics.n_nonfinite = cnt[2]
        rs.set_last_device(ses.x.reshape(ses.n, *event_shape))
        out.kernel = self.inner_kernel
        out.kernel.flow = self.kernel.flow                                              # neutra.py:128
        return out

    def _launch_latent(self, ses, pot, fd, k, imd, rng, st, sink):
        bij = self.kernel.flow.bijection
        if bij.uses_tensor_cores_for_neutra(ses.n):
            # wide flow: conditioner forward and input-VJP on tcgen05 (csrc/tc_neutra.cu)
            dev = ses.device
            td, keep = bij.tc_descriptor(dev)
            bt = bij.tc_transposed(dev)
            nb = N.lib().nfmc_neutra_tc_workspace_bytes(ses.d, ses.n)
            ws = ses.workspace(nb)
            N.check(N.lib().nfmc_neutra_hmc_steps_tc(C.byref(pot), C.byref(td), N.ptr(bt), bt.numel(), N.ptr(ses.x), ses.n, k,
                                                     float(self.inner_kernel.step_size), int(self.inner_kernel.n_leapfrog_steps),
                                                     N.ptr(imd), 1, C.byref(rng), ses.chain0, C.byref(st),
                                                     None if sink is None else C.byref(sink), N.ptr(ws), nb, ses.stream))
            return
        N.check(N.lib().nfmc_neutra_hmc_steps(C.byref(pot), C.byref(fd), N.ptr(ses.x), ses.n, k,
                                              float(self.inner_kernel.step_size), int(self.inner_kernel.n_leapfrog_steps),
                                              N.ptr(imd), C.byref(rng), ses.chain0, C.byref(st),
                                              None if sink is None else C.byref(sink), ses.stream))

    def _launch_latent_external(self, latent, ses, k, imd, sink, normals, uniforms):
        external.hmc_steps(latent, ses, k, float(self.inner_kernel.step_size), int(self.inner_kernel.n_leapfrog_steps), imd,
                           True, sink, normals, uniforms)           # recorded rows stay in z-space (reference quirk Q1)

    def _calls_grads(self, n):
        L = int(self.inner_kernel.n_leapfrog_steps)
        return (2 * L + 2) * n, 2 * L * n                                             # hmc.py:122-125

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        """Reference: NeuTra.warmup (neutra.py:70-107): variational fit of the flow (30 % of the budget), then tune the
        latent HMC (step size by dual averaging, inverse mass by the across-chain variance EMA)."""
        flow_limit = None if time_limit_seconds is None else 0.3 * time_limit_seconds
        t0 = time.time()
        self.kernel.flow.variational_fit(self.target.log_prob_fn(),
                                         **{"time_limit_seconds": flow_limit, **self.params.warmup_fit_kwargs},
                                         show_progress=show_progress)
        left = None if time_limit_seconds is None else max(time_limit_seconds - (time.time() - t0), 0.0)
        self.inner_params.n_warmup_iterations = self.params.n_warmup_iterations
        return self._run(x0, int(self.params.n_warmup_iterations), True, show_progress, left)


class NeuTraMH(NeuTraHMC):
    """Random-walk Metropolis in the flow's latent space (reference: nfmc/neutra.py:147-159 = MH.propose, mcmc/mh.py:44-73,
    on ``NeuTra.adjusted_target``).  Shares the outer loop, warm-up and output conventions of :class:`NeuTraHMC`."""

    def __init__(self, event_shape, target, inner_kernel: MHKernel = None, inner_params: MHParameters = None,
                 kernel: NeuTraKernel = None, params: NeuTraParameters = None):
        es = int(math.prod(tuple(event_shape)))
        super().__init__(event_shape, target, inner_kernel or MHKernel(event_size=es), inner_params or MHParameters(), kernel, params)

    @property
    def name(self):
        return "NeuTra MH"

    def _launch_latent(self, ses, pot, fd, k, imd, rng, st, sink):
        N.check(N.lib().nfmc_neutra_mh_steps(C.byref(pot), C.byref(fd), N.ptr(ses.x), ses.n, k, N.ptr(imd),
                                             int(bool(self.inner_params.adjustment)), C.byref(rng), ses.chain0, C.byref(st),
                                             None if sink is None else C.byref(sink), ses.stream))

    def _launch_latent_external(self, latent, ses, k, imd, sink, normals, uniforms):
        external.langevin_steps(latent, ses, k, 1.0, imd, bool(self.inner_params.adjustment), True, sink, normals, uniforms)

    def _calls_grads(self, n):
        return ((2 * n) if self.inner_params.adjustment else 0), 0                     # mh.py:68-71


# ---------------------------------------------------------------------------------------------------------------
# transport elliptical slice sampling
# ---------------------------------------------------------------------------------------------------------------
class TESS(Sampler):
    """Transport elliptical slice sampling (reference: nfmc/tess.py:88-188).  The chain state is the latent ``u``
    (``sample`` starts it at ``x0``, as the reference does); recorded samples and moments are the data-space points
    ``x = T^-1(u)``.  ``negative_log_likelihood`` is the potential the slice is taken on (tess.py passes it as
    ``potential``); ``target`` is carried for the API only."""

    def __init__(self, event_shape, target, negative_log_likelihood, kernel: Optional[TESSKernel] = None,
                 params: Optional[TESSParameters] = None):
        super().__init__(event_shape, target, kernel or TESSKernel(tuple(event_shape)), params or TESSParameters())
        self.negative_log_likelihood = resolve_target(negative_log_likelihood, self.event_shape)
        _require_analytic(self.negative_log_likelihood, "TESS (negative_log_likelihood)")

    @property
    def name(self):
        return "TESS"

    def _steps(self, ses: DeviceSession, k: int, sink, normals=None, uniforms=None):
        pot, keep = self.negative_log_likelihood.descriptor(ses.device)
        fd, keep2 = self.kernel.flow.bijection.descriptor(ses.device)
        rng = N.rng_desc(ses.seed, ses.local_step, normals, uniforms)
        st = ses.stats()
        N.check(N.lib().nfmc_tess_steps(C.byref(pot), C.byref(fd), N.ptr(ses.x), ses.n, k,
                                        int(self.params.max_ess_step_iterations), C.byref(rng), ses.chain0, C.byref(st),
                                        None if sink is None else C.byref(sink), ses.stream))
        ses.local_step += k

    def sample(self, x0, show_progress=True, time_limit_seconds=None, normals=None, uniforms=None) -> MCMCOutput:
        """``normals [T,n,d]`` / ``uniforms [T,n,2+M]`` optionally inject the random numbers (parity tests)."""
        event_shape = tuple(x0.shape[1:])
        store = bool(self.params.store_samples)
        out = MCMCOutput(event_shape, store_samples=store)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)      # u = x0 (tess.py:163)
        dev = ses.device
        T = int(self.params.n_iterations)
        M = int(self.params.max_ess_step_iterations)
        chunk = 1 if (time_limit_seconds is not None or show_progress) else T
        rs = out.running_samples
        done = 0
        for start in _progress(range(0, T, max(chunk, 1)), 'TESS sampling', show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds >= time_limit_seconds:
                break
            k = min(chunk, T - start)
            buf, sink = None, None
            if store:
                rows = _rows_kept(rs.seen_samples, k, rs.thinning)
                if rows * ses.n * ses.d * 4 > MAX_DEVICE_SAMPLE_BYTES:
                    raise MemoryError("sample buffer too large for the device; use store_samples=False")
                buf = torch.empty(rows, ses.n, ses.d, device=dev, dtype=torch.float32)
                sink = ses.sink(buf, rs.seen_samples, rs.thinning)
            nz = None if normals is None else N.dev_f32(normals[start:start + k], dev)
            un = None if uniforms is None else N.dev_f32(uniforms[start:start + k], dev)
            ses.tic()
            self._steps(ses, k, sink, nz, un)
            out.statistics.update_elapsed_time(ses.toc())
            done += k
            if buf is not None:
                rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=k)
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1],
                                       n_target_calls=(M + 1) * ses.n * done)          # tess.py:174-178
        self.latent_state = ses.x.reshape(ses.n, *event_shape)                         # the chains' final u (device)
        if not store or rs.n_samples == 0:
            # last_sample = the last recorded x: one more inverse pass on the final latent state
            rs.set_last_device(self.kernel.flow.bijection.inverse(self.latent_state)[0])
        out.kernel = self.kernel
        return out

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        """Reference: TESS.warmup (tess.py:101-149): start from a prior draw, alternate one TESS step with a flow fit on
        the step's data-space points (shuffled, split by ``train_pct``); the warm-up output records the LATENT states."""
        from .flow_train import train_val_split
        event_shape = tuple(x0.shape[1:])
        p: TESSParameters = self.params
        out = MCMCOutput(event_shape, store_samples=p.store_samples)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        rng = N.rng_desc(ses.seed, 0)                                                  # u ~ N(0, I) (tess.py:112), stream 3
        N.check(N.lib().nfmc_rng_fill(C.byref(rng), 3, ses.chain0, ses.d, ses.n, 1, N.ptr(ses.x), None, ses.stream))
        M = int(p.max_ess_step_iterations)
        done = 0
        x_buf = torch.empty(1, ses.n, ses.d, device=ses.device, dtype=torch.float32)
        u_sum = torch.zeros(2, ses.d, dtype=torch.float64)
        t_all = 0.0
        for _ in _progress(range(int(p.n_warmup_iterations)), '[Warmup] TESS', show_progress):
            if time_limit_seconds is not None and t_all >= time_limit_seconds:
                break
            t0 = time.time()
            ses.tic()
            self._steps(ses, 1, ses.sink(x_buf, 0, 1))
            ses.toc()
            u = ses.x.reshape(ses.n, *event_shape)
            out.running_samples.add(u.clone())                                       # tess.py:128-129 (latent states)
            ud = ses.x.double()
            u_sum += torch.stack([ud.sum(0), ud.square().sum(0)]).cpu()
            done += 1
            x_train, x_val = train_val_split(x_buf, p.train_pct, ses.n, ses.n)         # tess.py:138-142 (no size cap)
            self.kernel.flow.fit(x_train, x_val=x_val, **p.flow_fit_kwargs)
            t_all += time.time() - t0
        _, _, cnt = ses.read_back()
        out.statistics.expectations.add_sums(u_sum[0], u_sum[1], ses.n * done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1],
                                       n_target_calls=(M + 1) * ses.n * done)
        out.statistics.update_elapsed_time(t_all)
        out.kernel = self.kernel
        return out


# ---------------------------------------------------------------------------------------------------------------
# deterministic Langevin Monte Carlo
# ---------------------------------------------------------------------------------------------------------------
class DLMC(Sampler):
    """Deterministic Langevin Monte Carlo (reference: nfmc/dlmc.py:22-119): particles take one gradient step on
    ``U + log q`` (or, with ``latent_updates``, the step of dlmc.py:81-85 between ``T`` and ``T^-1``) with the flow ``q``
    refitted to the particles every iteration, followed by a flow-proposal MH correction.  ``negative_log_likelihood``
    only drives the initial update (dlmc.py:60-62), as in the reference."""

    def __init__(self, event_shape, target, negative_log_likelihood, kernel: Optional[DLMCKernel] = None,
                 params: Optional[DLMCParameters] = None):
        super().__init__(event_shape, target, kernel or DLMCKernel(tuple(event_shape)), params or DLMCParameters())
        self.negative_log_likelihood = resolve_target(negative_log_likelihood, self.event_shape)
        _require_analytic(self.target, "DLMC")
        _require_analytic(self.negative_log_likelihood, "DLMC (negative_log_likelihood)")

    @property
    def name(self):
        return "DLMC"

    def warmup(self, x0, show_progress=True, time_limit_seconds=None) -> MCMCOutput:
        out = MCMCOutput(event_shape=tuple(x0.shape[1:]), store_samples=self.params.store_samples)   # dlmc.py:37-42
        out.running_samples.add(x0)
        return out

    def sample(self, x0, show_progress=True, time_limit_seconds=None, z=None, uniforms=None, refit: bool = True) -> MCMCOutput:
        """``z [T,n,d]`` / ``uniforms [T,n]`` optionally inject the proposal draws; ``refit=False`` freezes the flow
        (parity tests: the refit is an optimiser run, everything else is pinned to the reference)."""
        from .flow_train import train_val_split
        p: DLMCParameters = self.params
        flow: Flow = self.kernel.flow
        if flow.bijection.uses_tensor_cores() or N.lib().nfmc_flow_param_count(
                flow.bijection.n_dim, flow.bijection.n_coupling, *flow.bijection.conditioner_shape()) < 0:
            raise NotImplementedError("dlmc needs the gradient of log q: conditioners with 2 linear layers and <= 8 hidden units")
        event_shape = tuple(x0.shape[1:])
        store = bool(p.store_samples)
        out = MCMCOutput(event_shape, store_samples=store)
        ses = DeviceSession(x0, event_shape, self.device, self.session_seed(), self.chain0)
        dev = ses.device
        eps = float(self.kernel.step_size)
        tgt, keep = self.target.descriptor(dev)
        nll, keep_n = self.negative_log_likelihood.descriptor(dev)
        logq = torch.empty(ses.n, device=dev, dtype=torch.float32)
        rs = out.running_samples
        ses.tic()
        N.check(N.lib().nfmc_potential_step(C.byref(nll), N.ptr(ses.x), ses.n, eps, ses.stream))      # dlmc.py:60-62
        out.statistics.update_counters(n_target_calls=ses.n, n_target_gradient_calls=ses.n)
        out.statistics.update_elapsed_time(ses.toc())
        done = 0
        for i in _progress(range(int(p.n_iterations)), 'DLMC sampling', show_progress):
            if time_limit_seconds is not None and out.statistics.elapsed_time_seconds >= time_limit_seconds:
                break
            t0 = time.time()
            if refit:                                                                              # dlmc.py:73-79
                x_train, x_val = train_val_split(ses.x, p.train_pct, p.max_train_size, p.max_val_size)
                flow.fit(x_train, x_val=x_val, **p.flow_fit_kwargs)
            fd, keep_f = flow.bijection.descriptor(dev)
            if p.latent_updates:                                                                   # dlmc.py:81-85
                zz, _ = flow.bijection.forward(ses.x)
                _, grad = self.target.value_and_grad(ses.x)
                N.check(N.lib().nfmc_dlmc_latent_update(N.ptr(zz), N.ptr(grad), eps, zz.numel(), ses.stream))
                ses.x.copy_(flow.bijection.inverse(zz)[0])
            else:                                                                                  # dlmc.py:86-88
                N.check(N.lib().nfmc_dlmc_update(C.byref(tgt), C.byref(fd), N.ptr(ses.x), ses.n, eps, ses.stream))
            out.statistics.update_counters(n_target_calls=ses.n, n_target_gradient_calls=ses.n)
            buf, sink = None, None
            if store:
                if _rows_kept(rs.seen_samples, 1, rs.thinning):
                    buf = torch.empty(1, ses.n, ses.d, device=dev, dtype=torch.float32)
                    sink = ses.sink(buf, rs.seen_samples, rs.thinning)
            zi = None if z is None else N.dev_f32(z[i], dev).reshape(1, ses.n, ses.d)
            ui = None if uniforms is None else N.dev_f32(uniforms[i], dev).reshape(1, ses.n)
            rng = N.rng_desc(ses.seed, ses.flow_step, zi, ui)
            st = ses.stats()
            N.check(N.lib().nfmc_imh_steps(C.byref(tgt), C.byref(fd), N.ptr(ses.x), N.ptr(logq), ses.n, 1, 1, C.byref(rng),
                                           ses.chain0, C.byref(st), None if sink is None else C.byref(sink), ses.stream))
            ses.flow_step += 1
            out.statistics.update_counters(n_target_calls=2 * ses.n)                               # dlmc.py:109-113
            done += 1
            if store:
                if buf is not None:
                    rs.add(buf.reshape(-1, ses.n, *event_shape), already_thinned=True, n_seen=1)
                else:
                    rs.seen_samples += 1
            torch.cuda.current_stream(dev).synchronize()
            out.statistics.update_elapsed_time(time.time() - t0)
        sx, sx2, cnt = ses.read_back()
        out.statistics.expectations.add_sums(sx, sx2, ses.n * done)
        out.statistics.update_counters(n_accepted_trajectories=cnt[0], n_attempted_trajectories=cnt[1])
        out.statistics.n_nonfinite = cnt[2]
        rs.set_last_device(ses.x.reshape(ses.n, *event_shape))
        out.kernel = self.kernel
        return out
