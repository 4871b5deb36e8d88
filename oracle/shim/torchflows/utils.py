from oracle.realnvp_ref import get_batch_shape, sum_except_batch  # noqa: F401
