"""nfmc_b200 -- B200-native (sm_100a) implementation of nfmc's batched sampler hot path.

Drop-in for the reference's ``nfmc.sample(strategy=..., flow='realnvp')`` on that path: same ``sample`` /
``create_sampler`` signature, same ``MCMCOutput`` / ``MCMCStatistics`` records.  All compute runs in
``libnfmc_b200.so`` (hand-written CUDA, C ABI in ``include/nfmc_b200.h``); importing this package does not need a
GPU, calling it does.
"""
from .api import sample, create_sampler, get_supported_samplers  # noqa: F401
from .records import MCMCOutput, MCMCStatistics, JumpNFMCStatistics  # noqa: F401
from . import potentials, flow, samplers, records  # noqa: F401

__version__ = "0.1.0"
