"""Import shim (test infrastructure): exposes the oracle's RealNVP restatement as ``torchflows``."""
from oracle.realnvp_ref import FlowRef as Flow, RealNVPRef as RealNVP  # noqa: F401
from . import flows, utils, architectures  # noqa: F401
