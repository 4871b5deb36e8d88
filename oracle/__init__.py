"""CPU oracle for the nfmc hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Everything under ``oracle/`` is a plain-PyTorch (fp32, CPU) restatement of the algorithm that
davidnabergoj/nfmc runs for the jump_mala / jump_hmc / neutra_hmc / imh + RealNVP path.  It exists only
to *check* the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  ``nfmc_b200`` (the product) never imports it and has no CPU
fallback.

Pinning status
--------------
* Sampler arithmetic (``oracle/samplers_ref.py``): PINNED against the reference itself.  The golden
  fixtures in ``tests/golden/*.npz`` were produced by importing the unmodified reference from
  ``/root/reference`` (``tests/golden/make_golden.py``) and the restatement reproduces them.
* Flow arithmetic (``oracle/realnvp_ref.py``): PARITY UNPINNED.  RealNVP lives in the third-party package
  ``torchflows`` (github davidnabergoj/torchflows), an unpinned, un-vendored dependency of the reference
  (``/root/reference/setup.py:51-56``) that is not installed here and cannot be fetched (no network).
  The restatement follows the published RealNVP construction (affine coupling, half split, MLP
  conditioner, reverse permutation, act-norm) and is anchored on the reference's own call sites
  (``sampling/base.py:26``, ``nfmc/util.py:280-281``, ``jump.py:205,218``, ``imh.py:214,221``,
  ``neutra.py:60``) and on ``test/test_flow_kwargs.py:18,28,30,49``.  It is self-checked for
  bijectivity and against ``torch.autograd.functional.jacobian`` log-determinants.
"""
