"""Only RealNVP is on the hot path; every other architecture name resolves to a stub that refuses to build."""
from oracle.realnvp_ref import RealNVPRef as RealNVP  # noqa: F401


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)

    class _Unavailable:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"torchflows.{name} is outside the hot path (oracle shim)")

    _Unavailable.__name__ = name
    return _Unavailable
