"""Import shim (test infrastructure): exposes the oracle's potentials as ``potentials``."""
