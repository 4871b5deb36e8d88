"""CPU: oracle self-checks that need neither the reference nor a GPU (i, ii, iii) and the
Philox known-answer vectors."""
import numpy as np
import pytest
import torch

from oracle.philox_ref import philox4x32_10, step_noise, layout_for_dim
from oracle.realnvp_ref import make_flow
from oracle.potentials_ref import make_potential_ref


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds: counter(4) key(2) -> expected(4)
    kat = [
        ([0x00000000] * 4, [0x00000000] * 2, [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kat:
        got = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array([key], dtype=np.uint32))[0]
        assert [int(v) for v in got] == want


def test_step_noise_is_standard_normal():
    z, u = step_noise(seed=7, stream=0, step=3, chain0=0, n=4096, d=100)
    assert z.shape == (4096, 100) and u.shape == (4096,)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    # chains keyed by GLOBAL index: shifting chain0 shifts the rows
    z2, u2 = step_noise(seed=7, stream=0, step=3, chain0=16, n=64, d=100)
    np.testing.assert_array_equal(z2, z[16:80])
    np.testing.assert_array_equal(u2, u[16:80])


@pytest.mark.parametrize("d,n_layers,ck", [(6, 2, None), (7, 3, dict(n_layers=3, n_hidden=6)), (8, 1, dict(n_layers=1)),
                                           (25, 2, None)])
def test_flow_bijective_and_logdet(d, n_layers, ck):
    torch.manual_seed(0)
    flow = make_flow((d,), n_layers=n_layers, conditioner_kwargs=ck, perturb=0.2, seed=5)
    x = torch.randn(9, d)
    with torch.no_grad():
        z, ld_f = flow.bijection.forward(x)
        xr, ld_i = flow.bijection.inverse(z)
    assert torch.allclose(xr, x, atol=2e-5)
    assert torch.allclose(ld_f + ld_i, torch.zeros(9), atol=2e-5)
    if d <= 8:
        for i in range(3):
            J = torch.autograd.functional.jacobian(lambda v: flow.bijection.forward(v[None])[0][0], x[i])
            assert abs(float(torch.linalg.slogdet(J)[1]) - float(ld_f[i])) < 1e-4


def test_flow_sample_log_prob_consistent():
    torch.manual_seed(1)
    flow = make_flow((6,), perturb=0.2, seed=6)
    with torch.no_grad():
        x, lq = flow.sample(32, return_log_prob=True)
        assert torch.allclose(flow.log_prob(x), lq, atol=1e-4)


def test_layout_table():
    assert layout_for_dim(100) == (4, 13) and layout_for_dim(25) == (1, 13) and layout_for_dim(1000) == (32, 16)
    assert layout_for_dim(6) == (1, 4) and layout_for_dim(64) == (2, 16)


def test_potentials_match_reference_test_targets():
    x = torch.randn(5, 10)
    assert torch.equal(make_potential_ref("g0", (10,))(x), torch.sum(x ** 2, dim=-1))   # /root/reference/test/util.py:4-5
