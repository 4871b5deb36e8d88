"""Helpers for the GPU parity tests: product objects that mirror an oracle / golden case."""
import math

import torch

import nfmc_b200
from nfmc_b200 import potentials as P
from nfmc_b200.flow import Flow, RealNVP
from nfmc_b200.records import MCMCOutput
from nfmc_b200.samplers import DeviceSession


def product_flow_from_oracle(oflow, conditioner_dtype="fp32") -> Flow:
    bij = oflow.bijection
    cpl = [l for l in bij.layers if hasattr(l, "net")]
    ck = dict(n_layers=cpl[0].n_linear, n_hidden=cpl[0].n_hidden) if cpl else None
    f = Flow(RealNVP(bij.event_shape, n_layers=bij.n_coupling, conditioner_kwargs=ck, conditioner_dtype=conditioner_dtype))
    missing = f.load_state_dict(oflow.state_dict(), strict=True)
    return f.to("cuda").eval()


def product_target(name, d, callable_target: bool = False):
    """The built-in analytic potential, or (``callable_target``) the same function as a plain Python callable in torch
    operations -- what a user of the reference passes as ``target`` -- which sends the samplers down their external-target
    path (autograd for U / grad U, nfmc_ext_* kernels for the rest)."""
    if callable_target:
        from oracle.potentials_ref import make_potential_ref
        ref = make_potential_ref(str(name), (d,)).cuda()
        return lambda x: ref(x)
    return P.make_potential(str(name), (d,))


def run_local_injected(sampler, x0, normals, uniforms, store=True):
    """K steps of a local sampler with injected noise; returns (samples [K,n,d] on host, session, output)."""
    K = normals.shape[0]
    out = MCMCOutput(tuple(x0.shape[1:]), store_samples=store)
    ses = DeviceSession(x0, tuple(x0.shape[1:]), None, seed=0)
    dev = ses.device
    buf = sampler.run_steps(ses, out, K, store, normals.to(dev).contiguous(), None if uniforms is None else uniforms.to(dev).contiguous())
    torch.cuda.synchronize()
    sx, sx2, cnt = ses.read_back()
    return (None if buf is None else buf.cpu()), ses, (sx, sx2, cnt)


def oracle_flow_from_product(flow, device=None):
    """The ORACLE RealNVP (oracle/realnvp_ref.py) carrying the product flow's parameters -- the reference the training
    kernels are differentiated against (torch autograd through the oracle's own forward / inverse), on `device`."""
    from oracle.realnvp_ref import FlowRef, RealNVPRef
    bij = flow.bijection
    M, H = bij.conditioner_shape()
    ck = dict(n_layers=M, n_hidden=H) if bij.n_coupling > 0 else None
    oflow = FlowRef(RealNVPRef(bij.event_shape, n_layers=bij.n_coupling, conditioner_kwargs=ck))
    oflow.load_state_dict({k: v.detach().clone() for k, v in flow.state_dict().items()}, strict=True)
    for l in oflow.bijection.layers:
        if hasattr(l, "initialised"):
            l.initialised.fill_(True)
    oflow = oflow.to(device if device is not None else next(flow.parameters()).device).eval()
    assert [tuple(p.shape) for p in oflow.bijection.parameters()] == [tuple(p.shape) for p in flow.bijection.parameters()]
    return oflow
