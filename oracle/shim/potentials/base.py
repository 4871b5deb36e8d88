from oracle.potentials_ref import PotentialRef as Potential  # noqa: F401
