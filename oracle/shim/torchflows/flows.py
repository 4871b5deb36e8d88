from oracle.realnvp_ref import FlowRef as Flow  # noqa: F401
