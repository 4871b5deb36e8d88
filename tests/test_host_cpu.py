"""CPU: host-side logic and the C-ABI boundary (no compute calls -- there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from nfmc_b200 import _native as N
    lib = N.lib()
    header = open(os.path.join(ROOT, "include", "nfmc_b200.h")).read()
    declared = set(re.findall(r"NFMC_API\s+[\w\s\*]+?\b(nfmc_\w+)\s*\(", header))
    assert len(declared) >= 18
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.nfmc_abi_version() == 1


def test_layout_and_blob_size_match_python_side():
    from nfmc_b200 import _native as N
    from nfmc_b200.flow import create_flow_object, pack_realnvp
    from oracle.philox_ref import layout_for_dim
    for d in [2, 5, 6, 8, 9, 25, 26, 27, 33, 64, 100, 101, 200, 784, 1000, 1024]:
        assert N.layout_for_dim(d) == layout_for_dim(d)
    for d, spec in [(6, "realnvp"), (7, 'realnvp%{"n_layers": 3, "conditioner_kwargs": {"n_layers": 3, "n_hidden": 6}}'),
                    (100, 'realnvp%{"n_layers": 10, "conditioner_kwargs": {"n_layers": 5, "n_hidden": 100}}'),
                    (9, 'realnvp%{"conditioner_kwargs": {"n_layers": 1}}')]:
        f = create_flow_object(spec, (d,))
        M, H = f.bijection.conditioner_shape()
        blob = pack_realnvp(f.bijection)
        assert blob.numel() == N.lib().nfmc_realnvp_blob_floats(d, f.bijection.n_coupling, M, H)
        assert len(f.bijection.layers) == 3 * f.bijection.n_coupling + 3


def test_errors_are_reported_not_swallowed():
    from nfmc_b200 import _native as N
    gs, e = C.c_int32(), C.c_int32()
    assert N.lib().nfmc_layout_for_dim(5000, C.byref(gs), C.byref(e)) != 0
    assert b"out of range" in N.lib().nfmc_last_error()
    with pytest.raises(N.NativeError):
        N.check(N.lib().nfmc_layout_for_dim(0, C.byref(gs), C.byref(e)))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import nfmc_b200
    from nfmc_b200 import _native as N
    with pytest.raises(N.NativeError):
        nfmc_b200.sample("g0", (5,), strategy="jump_mala", n_chains=4, n_iterations=2, show_progress=False)
    # a callable target resolves to the external-target wrapper (no compute here); anything else is a TypeError
    t = nfmc_b200.potentials.resolve_target(lambda x: (x ** 2).sum(-1), (5,))
    assert t.external and t.event_shape == (5,)
    with pytest.raises(N.NativeError):
        t.descriptor(torch.device("cpu"))
    with pytest.raises(TypeError):
        nfmc_b200.potentials.resolve_target(3.0, (5,))

    class ForeignPotential:                      # duck-typed like potentials.base.Potential: callable with .event_shape
        event_shape = (3, 2)

        def __call__(self, x):
            return (x ** 2).flatten(1).sum(1)

    t = nfmc_b200.potentials.resolve_target(ForeignPotential(), (3, 2))
    assert t.external and t.n_dim == 6


def test_flow_state_dict_is_interchangeable_with_oracle():
    from nfmc_b200.flow import Flow, RealNVP
    from oracle.realnvp_ref import make_flow
    o = make_flow((7,), n_layers=3, conditioner_kwargs=dict(n_layers=3, n_hidden=6), perturb=0.1, seed=3)
    f = Flow(RealNVP((7,), n_layers=3, conditioner_kwargs=dict(n_layers=3, n_hidden=6)))
    f.load_state_dict(o.state_dict(), strict=True)
    assert set(f.state_dict()) == set(o.state_dict())


def test_pack_folds_reversals():
    """Packed affine parameters of layers behind an odd number of reversals are stored flipped; the trailing
    ElementwiseAffine + ActNorm are composed into the last act-norm."""
    from nfmc_b200.flow import Flow, RealNVP, pack_realnvp, blob_floats
    f = Flow(RealNVP((6,), n_layers=2))
    with torch.no_grad():
        f.bijection.layers[3].value[:, 1] = torch.arange(6.0)     # act-norm after coupling 0: 1 reversal
        f.bijection.layers[6].value[:, 1] = torch.arange(6.0)     # act-norm after coupling 1: 2 reversals
        f.bijection.layers[8].value[:, 1] = 2.0                   # trailing act-norm (composed into affine 2)
    blob = pack_realnvp(f.bijection)
    d = 6
    assert blob.numel() == blob_floats(6, 2, 2, 4)
    fwd1 = blob[1 * 4 * d: 1 * 4 * d + 2 * d].reshape(d, 2)       # {alpha, beta} of affine 1
    fwd2 = blob[2 * 4 * d: 2 * 4 * d + 2 * d].reshape(d, 2)
    assert torch.allclose(fwd1[:, 1], (torch.arange(6.0) / 2).flip(0))
    assert torch.allclose(fwd2[:, 1], torch.arange(6.0) / 2 + 1.0)
    assert torch.allclose(fwd1[:, 0], torch.ones(6), atol=1e-6)


def test_records_contract():
    from nfmc_b200.records import (MCMCOutput, MCMCStatistics, JumpNFMCOutput, JumpNFMCStatistics, MCMCSamples,
                                   LangevinKernel, HMCKernel, JumpNFMCParameters, IMHParameters, NeuTraParameters)
    st = MCMCStatistics((3,))
    for f in ["n_accepted_trajectories", "n_attempted_trajectories", "n_divergences", "n_target_gradient_calls",
              "n_target_calls", "elapsed_time_seconds"]:
        assert hasattr(st, f)
    assert st.acceptance_rate != st.acceptance_rate  # nan before any step
    x = torch.randn(4, 5, 3)
    st.expectations.update(x)
    assert torch.allclose(st.running_first_moment, x.mean((0, 1)), atol=1e-6)
    assert torch.allclose(st.running_second_moment, (x ** 2).mean((0, 1)), atol=1e-6)
    assert torch.allclose(st.expectations['first_moment'].as_tensor(), x.mean((0, 1)), atol=1e-6)
    st.update_counters(n_accepted_trajectories=3, n_attempted_trajectories=4)
    assert st.acceptance_rate == 0.75 and set(st.__dict__()) >= {"acceptance_rate", "calls_per_second"}
    jo = JumpNFMCOutput((3,))
    assert isinstance(jo, MCMCOutput) and isinstance(jo.statistics, JumpNFMCStatistics)
    jo.statistics.update_counters(n_accepted_jumps=1, n_attempted_jumps=2, n_target_calls=5)
    assert jo.statistics.jump_acceptance_rate == 0.5 and jo.statistics.n_target_calls == 5
    # sample store: thinning / max_samples / store_samples=False (reference base.py:234-263)
    rs = MCMCSamples((3,), thinning=2, max_samples=3)
    rs.add(torch.arange(7 * 2 * 3.0).reshape(7, 2, 3))
    assert rs.n_samples == 3 and rs.as_tensor().shape == (3, 2, 3) and torch.equal(rs.last_sample, torch.arange(36.0, 42).reshape(2, 3))
    out = MCMCOutput((3,), store_samples=False)
    out.running_samples.add(torch.zeros(2, 3))
    assert out.samples is None and out.running_samples.last_sample.shape == (2, 3)
    # defaults that the reference fixes
    assert abs(LangevinKernel(event_size=100).step_size - 100 ** (-1 / 3)) < 1e-12        # langevin.py:17-18
    assert HMCKernel(event_size=5).step_size == 0.01 and HMCKernel(event_size=5).n_leapfrog_steps == 20
    assert JumpNFMCParameters().adjusted_jumps and not JumpNFMCParameters().fit_nf
    assert IMHParameters().flow_fit_kwargs is None and NeuTraParameters().flow_fit_kwargs is None   # quirk Q4
    with pytest.raises(ValueError):
        IMHParameters(train_distribution="nope")


def test_dual_averaging_matches_oracle():
    from nfmc_b200.records import DualAveraging, DualAveragingParams
    from oracle.samplers_ref import DualAveragingRef
    a, b = DualAveraging(0.1, DualAveragingParams()), DualAveragingRef(0.1)
    for acc in [0.9, 0.2, 0.7, 0.65, 0.1]:
        a.step(0.651 - acc)
        v = b.step(acc)
        assert abs(a.value - v) < 1e-12


def test_sample_store_matches_reference_semantics():
    """MCMCSamples.add / thinning / max_samples window / indexing (reference: sampling/base.py:215-271), restated as a
    plain list and compared block by block, for single states and multi-row blocks."""
    import torch
    from nfmc_b200.records import MCMCSamples

    def reference_store(blocks, thinning, max_samples):
        running, seen = [], 0
        for x in blocks:
            x = x[None] if x.ndim == 2 else x
            mask = (torch.arange(seen, seen + len(x)) % thinning) == 0
            seen += len(x)
            running.extend(x[mask])
            if max_samples is not None:
                running = running[-max_samples:]
        return torch.stack(running) if running else None

    torch.manual_seed(0)
    blocks = [torch.randn(4, 3), torch.randn(5, 4, 3), torch.randn(1, 4, 3), torch.randn(4, 3), torch.randn(7, 4, 3)]
    for thinning in (1, 2, 3):
        for max_samples in (None, 4, 100):
            rs = MCMCSamples((3,), thinning=thinning, max_samples=max_samples)
            for b in blocks:
                rs.add(b)
            ref = reference_store(blocks, thinning, max_samples)
            assert rs.n_samples == len(ref)
            for i in range(len(ref)):
                assert torch.equal(rs.row(i), ref[i])
            assert torch.equal(rs.as_tensor(), ref)
            assert torch.equal(rs.last_sample, blocks[-1][-1])
            assert torch.equal(rs[-1], blocks[-1][-1])
    rs = MCMCSamples((3,), store_samples=False)
    rs.add(blocks[0])
    assert rs.n_samples == 0 and torch.equal(rs.last_sample, blocks[0])
