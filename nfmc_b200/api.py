"""``sample`` / ``create_sampler`` with the reference's signature (``/root/reference/nfmc/sample.py:20-30,243-255``).

Strategies on the accelerated hot path: ``jump_mala``, ``jump_ula``, ``jump_hmc``, ``jump_uhmc``, ``neutra_hmc``,
``imh`` / ``fixed_imh``, ``adaptive_imh`` and the local kernels they are built from (``mala``, ``ula``, ``hmc``,
``uhmc``, ``mh`` / ``jump_mh`` / ``neutra_mh``, ``ess`` / ``jump_ess``).  ``tess``, ``dlmc`` -- every strategy of the reference's ``get_supported_samplers()``.  Anything else (``nuts`` ...) is outside the scope table and raises ``NotImplementedError`` -- there is no eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import torch

from . import _native as N
from .flow import Flow, create_flow_object
from .potentials import Potential, resolve_target
from .records import (DLMCKernel, DLMCParameters, TESSKernel, TESSParameters, ESSKernel, ESSParameters, MHKernel, MHParameters, HMCKernel, HMCParameters, IMHKernel, IMHParameters, JumpNFMCParameters, LangevinKernel,
                      LangevinParameters, MCMCOutput, NeuTraKernel, NeuTraParameters, NFMCKernel)
from .samplers import (DLMC, TESS, ESS, JumpESS, MH, JumpMH, NeuTraMH, HMC, MALA, UHMC, ULA, AdaptiveIMH, FixedIMH, JumpHMC, JumpMALA, JumpUHMC, JumpULA, NeuTraHMC,
                       Sampler)

LOCAL_STRATEGIES = ('hmc', 'uhmc', 'ula', 'mala', 'mh', 'ess')
NF_STRATEGIES = ('imh', 'fixed_imh', 'adaptive_imh', 'jump_mala', 'jump_ula', 'jump_hmc', 'jump_uhmc', 'jump_mh', 'jump_ess',
                 'neutra_hmc', 'neutra_mh', 'tess', 'dlmc')


def get_supported_samplers():
    return list(LOCAL_STRATEGIES) + list(NF_STRATEGIES)


def create_sampler(target, event_shape: Optional[Tuple[int, ...]] = None, flow: Optional[Union[str, Flow]] = 'realnvp',
                   strategy: str = "imh", negative_log_likelihood=None, kernel_kwargs: Optional[dict] = None,
                   param_kwargs: Optional[dict] = None, inner_kernel_kwargs: Optional[dict] = None,
                   inner_param_kwargs: Optional[dict] = None, device: torch.device = None,
                   flow_kwargs: Optional[dict] = None) -> Sampler:
    flow_kwargs = flow_kwargs or {}
    kernel_kwargs = kernel_kwargs or {}
    param_kwargs = param_kwargs or {'n_iterations': 100}
    inner_kernel_kwargs = inner_kernel_kwargs or {}
    inner_param_kwargs = dict(inner_param_kwargs or {})

    if flow is not None and not isinstance(flow, str):
        event_shape = flow.event_shape
    elif isinstance(target, Potential) or (callable(target) and hasattr(target, 'event_shape')):
        event_shape = tuple(target.event_shape)   # our potentials, or any `potentials.base.Potential`-like callable (sample.py:65-66)
    if event_shape is None:
        raise ValueError("event_shape is required")
    event_shape = tuple(event_shape)
    event_size = int(math.prod(event_shape))
    target = resolve_target(target, event_shape)
    dev = N.require_cuda(None if device is None or torch.device(device).type != 'cuda' else device)

    def finish(s: Sampler) -> Sampler:
        s.device = dev
        if hasattr(s, 'inner_sampler'):
            s.inner_sampler.device = dev
        return s

    if strategy in LOCAL_STRATEGIES:
        if strategy in ('hmc', 'uhmc'):
            cls = HMC if strategy == 'hmc' else UHMC
            return finish(cls(event_shape, target, HMCKernel(event_size=event_size, **kernel_kwargs),
                              HMCParameters(**param_kwargs)))
        if strategy == 'ess':
            if negative_log_likelihood is None:
                raise ValueError("Negative log likelihood must be provided")          # reference: sample.py:93-94
            return finish(ESS(event_shape, target, negative_log_likelihood, ESSKernel(event_shape=event_shape, **kernel_kwargs),
                              ESSParameters(**param_kwargs)))
        if strategy == 'mh':
            return finish(MH(event_shape, target, MHKernel(event_size=event_size, **kernel_kwargs), MHParameters(**param_kwargs)))
        cls = MALA if strategy == 'mala' else ULA
        return finish(cls(event_shape, target, LangevinKernel(event_size=event_size, **kernel_kwargs),
                          LangevinParameters(**param_kwargs)))
    if strategy not in NF_STRATEGIES:
        raise NotImplementedError(f"strategy {strategy!r} is outside the accelerated hot path "
                                  f"(supported: {get_supported_samplers()})")
    if flow is None:
        raise ValueError("Flow object must be provided")
    if isinstance(flow, str):
        flow_object = create_flow_object(flow_string=flow, event_shape=event_shape, **flow_kwargs).to(dev)
    elif isinstance(flow, Flow):
        flow_object = flow.to(dev)
    else:
        raise ValueError(f"Unknown type for normalizing flow: {type(flow)}")

    if strategy in ('imh', 'fixed_imh'):
        return finish(FixedIMH(event_shape, target, IMHKernel(event_shape, flow=flow_object), IMHParameters(**param_kwargs)))
    if strategy == 'adaptive_imh':
        # the reference drops param_kwargs here (sample.py:129, quirk Q2): always 100 iterations, samples stored
        return finish(AdaptiveIMH(event_shape, target, IMHKernel(event_shape, flow=flow_object), IMHParameters()))
    if strategy in ('jump_mala', 'jump_ula'):
        cls = JumpMALA if strategy == 'jump_mala' else JumpULA
        return finish(cls(event_shape, target, kernel=NFMCKernel(event_shape, flow=flow_object),
                          params=JumpNFMCParameters(**param_kwargs),
                          inner_kernel=LangevinKernel(event_size=event_size, **inner_kernel_kwargs),
                          inner_params=LangevinParameters(**inner_param_kwargs)))
    if strategy == 'jump_mh':
        return finish(JumpMH(event_shape, target, kernel=NFMCKernel(event_shape, flow=flow_object),
                             params=JumpNFMCParameters(**param_kwargs),
                             inner_kernel=MHKernel(event_size=event_size, **inner_kernel_kwargs),
                             inner_params=MHParameters(**inner_param_kwargs)))
    if strategy == 'jump_ess':
        if negative_log_likelihood is None:
            raise ValueError("Negative log likelihood must be provided")              # reference: sample.py:199-200
        return finish(JumpESS(event_shape, target, negative_log_likelihood, kernel=NFMCKernel(event_shape, flow=flow_object),
                              params=JumpNFMCParameters(**param_kwargs),
                              inner_kernel=ESSKernel(event_shape=event_shape, **inner_kernel_kwargs),
                              inner_params=ESSParameters(**inner_param_kwargs)))
    if strategy in ('jump_hmc', 'jump_uhmc'):
        if strategy == 'jump_hmc' and 'n_iterations' not in inner_param_kwargs:
            inner_param_kwargs['n_iterations'] = 5                                   # reference: sample.py:161-162
        cls = JumpHMC if strategy == 'jump_hmc' else JumpUHMC
        return finish(cls(event_shape, target, kernel=NFMCKernel(event_shape, flow=flow_object),
                          params=JumpNFMCParameters(**param_kwargs),
                          inner_kernel=HMCKernel(event_size=event_size, **inner_kernel_kwargs),
                          inner_params=HMCParameters(**inner_param_kwargs)))
    if strategy == 'tess':                                                               # reference: sample.py:214-219
        if negative_log_likelihood is None:
            raise ValueError("Negative log likelihood must be provided")
        return finish(TESS(event_shape, target, negative_log_likelihood, TESSKernel(event_shape, flow=flow_object),
                           TESSParameters(**param_kwargs)))
    if strategy == 'dlmc':                                                               # reference: sample.py:220-225
        if negative_log_likelihood is None:
            raise ValueError("Negative log likelihood must be provided")
        return finish(DLMC(event_shape, target, negative_log_likelihood, DLMCKernel(event_shape, flow=flow_object),
                           DLMCParameters(**param_kwargs)))
    if strategy == 'neutra_mh':                                                          # reference: sample.py:232-237
        return finish(NeuTraMH(event_shape, target, MHKernel(event_size=event_size, **inner_kernel_kwargs),
                               MHParameters(**inner_param_kwargs), NeuTraKernel(event_shape, flow=flow_object),
                               NeuTraParameters(**param_kwargs)))
    # neutra_hmc
    return finish(NeuTraHMC(event_shape, target, HMCKernel(event_size=event_size, **inner_kernel_kwargs),
                            HMCParameters(**inner_param_kwargs), NeuTraKernel(event_shape, flow=flow_object),
                            NeuTraParameters(**param_kwargs)))


def sample(target, event_shape: Optional[Tuple[int, ...]] = None, flow: Optional[Union[str, Flow]] = 'realnvp',
           strategy: str = "imh", n_iterations: int = 100, n_warmup_iterations: int = 100, n_chains: int = 100,
           x0: torch.Tensor = None, warmup: bool = False, show_progress: bool = True,
           sampling_time_limit_seconds=None, warmup_time_limit_seconds=None, **kwargs) -> MCMCOutput:
    """Sample from ``target``: a ``nfmc_b200.potentials.Potential`` or its name (fused kernels), or any callable
    ``[n, *event] -> [n]`` in torch operations (external-target path, as the reference's own contract).  Same arguments and return
    type as the reference's ``nfmc.sample`` (sample.py:243-314); ``x0`` may be a host tensor -- it is copied to
    the GPU once and results come back as host tensors."""
    if flow == 'None':
        flow = None
    if flow is not None and not isinstance(flow, str):
        event_shape = flow.event_shape
    elif isinstance(target, Potential) or (callable(target) and hasattr(target, 'event_shape')):
        event_shape = tuple(target.event_shape)   # our potentials, or any `potentials.base.Potential`-like callable (sample.py:65-66)
    kwargs['param_kwargs'] = {**kwargs.get('param_kwargs', {}),
                              'n_iterations': n_iterations, 'n_warmup_iterations': n_warmup_iterations}
    sampler = create_sampler(target=target, event_shape=event_shape, flow=flow, strategy=strategy, **kwargs)
    if x0 is None:
        x0 = torch.randn(size=(n_chains, *tuple(event_shape)))                      # reference: sample.py:305
    if warmup:
        w = sampler.warmup(x0=x0, show_progress=show_progress, time_limit_seconds=warmup_time_limit_seconds)
        rs = w.running_samples
        if w.store_samples:
            pool = rs.device_tensor()                   # shuffle on the GPU when the warm-up samples live there
            flat = (pool if pool is not None else w.samples).flatten(0, 1)
            x0 = flat[torch.randperm(len(flat), device=flat.device)[:n_chains]]   # reference: sample.py:309-311
        else:                                           # stay on the device when the warm-up left its state there
            x0 = rs.last_sample_device if rs.last_sample_device is not None else rs.last_sample
    return sampler.sample(x0=x0, show_progress=show_progress, time_limit_seconds=sampling_time_limit_seconds)
