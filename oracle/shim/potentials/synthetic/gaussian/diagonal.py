import torch
from oracle.potentials_ref import DiagGaussianRef


class DiagonalGaussian1(DiagGaussianRef):
    """100-dim diagonal Gaussian (/root/reference/test/test_moment_estimation.py:10,16 needs only this surface)."""

    def __init__(self):
        sigma = torch.linspace(1.0, 10.0, 100)
        super().__init__((100,), 1.0 / sigma ** 2)
