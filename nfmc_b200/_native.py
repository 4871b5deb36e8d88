"""ctypes binding of ``libnfmc_b200.so`` (the C ABI declared in ``include/nfmc_b200.h``).

This is the only layer between the Python host mirror and the CUDA kernels.  There is deliberately no
fallback: if the shared library is missing, or a compute entry point is called without a CUDA device, the
call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NFMC_B200_LIB", os.path.join(_HERE, "libnfmc_b200.so"))
CSRC_DIR = os.path.join(_HERE, "csrc")

# ---- enums / structs (mirror include/nfmc_b200.h) ---------------------------------------------------------
POT_ISO_GAUSSIAN, POT_DIAG_GAUSSIAN, POT_FUNNEL, POT_ROSENBROCK, POT_MIXTURE4 = range(5)
MAX_DIM = 1024


class PotentialDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("d", C.c_int32), ("params", C.c_void_p), ("scalar", C.c_float * 4)]


class RealNVPDesc(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_coupling", C.c_int32), ("n_linear", C.c_int32), ("hidden", C.c_int32),
                ("blob", C.c_void_p), ("blob_floats", C.c_int64)]


class RealNVPTcDesc(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_coupling", C.c_int32), ("hidden", C.c_int32), ("reserved", C.c_int32),
                ("blob", C.c_void_p), ("blob_bytes", C.c_int64)]


class RngDesc(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("step0", C.c_uint64), ("normals", C.c_void_p), ("uniforms", C.c_void_p)]


class StatsDesc(C.Structure):
    _fields_ = [("sum_x", C.c_void_p), ("sum_x2", C.c_void_p), ("counts", C.c_void_p)]


class SinkDesc(C.Structure):
    _fields_ = [("samples", C.c_void_p), ("seen0", C.c_int64), ("thinning", C.c_int32)]


P = C.POINTER
_vp, _i32, _i64, _f32, _u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes): every symbol include/nfmc_b200.h declares
SIGNATURES = {
    "nfmc_last_error": (C.c_char_p, []),
    "nfmc_abi_version": (C.c_int, []),
    "nfmc_layout_for_dim": (C.c_int, [_i32, P(_i32), P(_i32)]),
    "nfmc_realnvp_blob_floats": (_i64, [_i32, _i32, _i32, _i32]),
    "nfmc_potential_eval": (C.c_int, [P(PotentialDesc), _vp, _vp, _vp, _i64, _vp]),
    "nfmc_realnvp_forward": (C.c_int, [P(RealNVPDesc), _vp, _vp, _vp, _i64, _vp]),
    "nfmc_realnvp_inverse": (C.c_int, [P(RealNVPDesc), _vp, _vp, _vp, _i64, _vp]),
    "nfmc_flow_log_prob": (C.c_int, [P(RealNVPDesc), _vp, _vp, _i64, _vp]),
    "nfmc_realnvp_tc_blob_bytes": (_i64, [_i32, _i32, _i32]),
    "nfmc_flow_tc_pass": (C.c_int, [P(RealNVPTcDesc), _i32, _vp, _vp, _vp, _i64, _vp]),
    "nfmc_jump_tc_workspace_bytes": (_i64, [_i32, _i64]),
    "nfmc_jump_step_tc": (C.c_int, [P(PotentialDesc), P(RealNVPTcDesc), _vp, _vp, _i32, _i64, _i32, P(RngDesc), _i64,
                                    P(StatsDesc), P(SinkDesc), _vp, _i64, _vp]),
    "nfmc_jump_step_wide": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _i32, _vp, _vp, _i32, _i64, _i32, P(RngDesc), _i64,
                                      P(StatsDesc), P(SinkDesc), _vp, _i64, _vp]),
    "nfmc_flow_sample": (C.c_int, [P(RealNVPDesc), P(RngDesc), _i64, _vp, _vp, _i64, _vp]),
    "nfmc_mala_steps": (C.c_int, [P(PotentialDesc), _vp, _i64, _i32, _f32, _vp, _i32, P(RngDesc), _i64,
                                  P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_mh_steps": (C.c_int, [P(PotentialDesc), _vp, _i64, _i32, _vp, _i32, P(RngDesc), _i64, P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_ess_steps": (C.c_int, [P(PotentialDesc), _vp, _i64, _i32, _i32, P(RngDesc), _i64, P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_hmc_steps": (C.c_int, [P(PotentialDesc), _vp, _i64, _i32, _f32, _i32, _vp, _i32, P(RngDesc), _i64,
                                 P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_chain_sums": (C.c_int, [_vp, _i64, _i32, _vp, _vp, _vp]),
    "nfmc_tune_inv_mass": (C.c_int, [_vp, _i32, _f32, _vp, _vp]),
    "nfmc_jump_step": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _i32, P(RngDesc), _i64,
                                 P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_jump_step2": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _vp, _i64, _i32, P(RngDesc), _i64,
                                  P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_imh_steps": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _vp, _i64, _i32, _i32, P(RngDesc), _i64,
                                 P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_neutra_hmc_steps": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _i32, _f32, _i32, _vp,
                                        P(RngDesc), _i64, P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_tess_steps": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _i32, _i32, P(RngDesc), _i64, P(StatsDesc),
                                  P(SinkDesc), _vp]),
    "nfmc_neutra_mh_steps": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _i32, _vp, _i32, P(RngDesc), _i64,
                                       P(StatsDesc), P(SinkDesc), _vp]),
    "nfmc_neutra_potential": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _vp, _vp, _i64, _vp]),
    "nfmc_neutra_tc_transposed_bytes": (_i64, [_i32, _i32, _i32]),
    "nfmc_neutra_tc_workspace_bytes": (_i64, [_i32, _i64]),
    "nfmc_neutra_potential_tc": (C.c_int, [P(PotentialDesc), P(RealNVPTcDesc), _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "nfmc_neutra_hmc_steps_tc": (C.c_int, [P(PotentialDesc), P(RealNVPTcDesc), _vp, _i64, _vp, _i64, _i32, _f32, _i32, _vp, _i32,
                                           P(RngDesc), _i64, P(StatsDesc), P(SinkDesc), _vp, _i64, _vp]),
    "nfmc_flow_param_count": (_i64, [_i32, _i32, _i32, _i32]),
    "nfmc_flow_pack": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "nfmc_flow_nll_grad": (C.c_int, [P(RealNVPDesc), _vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "nfmc_flow_kl_grad": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), P(RngDesc), _i64, _i64, _vp, _vp, _i32, _vp]),
    "nfmc_flow_grad_unpack": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _f32, _vp, _vp]),
    "nfmc_adamw_step": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
    "nfmc_flow_fit_epoch": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                      _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
    "nfmc_flow_wide_param_count": (_i64, [_i32, _i32, _i32, _i32]),
    "nfmc_flow_wide_nll_grad": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]),
    "nfmc_flow_wide_pass": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _vp, _i64, _vp]),
    "nfmc_flow_wide_log_prob": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _i64, _vp]),
    "nfmc_flow_wide_sample": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, P(RngDesc), _i64, _vp, _vp, _i64, _vp]),
    "nfmc_flow_wide_pullback": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _i32, _vp, _vp, _i64, _vp, _vp]),
    "nfmc_flow_wide_sweep": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _i32, _vp, _vp, _i64, _vp, _vp, _i32, _vp]),
    "nfmc_adamw_step_scaled": (C.c_int, [_vp, _vp, _f32, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
    "nfmc_flow_wide_fit_epoch": (C.c_int, [_i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64,
                                           _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
    "nfmc_rng_fill": (C.c_int, [P(RngDesc), _i32, _i64, _i32, _i64, _i32, _vp, _vp, _vp]),
    "nfmc_ext_langevin_propose": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _i32, _i64, _i32, _vp, _vp]),
    "nfmc_ext_langevin_log_ratio": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32, _i32, _i64, _i32, _vp, _vp]),
    "nfmc_ext_hmc_momentum": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "nfmc_ext_hmc_leapfrog": (C.c_int, [_vp, _vp, _vp, _vp, _f32, _i32, _i32, _i64, _i32, _vp]),
    "nfmc_ext_hmc_log_ratio": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "nfmc_ext_jump_log_ratio": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nfmc_ext_accept": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, P(StatsDesc), P(SinkDesc),
                                  _i32, _vp]),
    "nfmc_ext_ess_uniforms": (C.c_int, [_u64, _u64, _i64, _i64, _i32, _vp, _vp]),
    "nfmc_ext_ess_begin": (C.c_int, [_vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "nfmc_ext_ess_rotate": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp]),
    "nfmc_ext_ess_update": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _i32, _vp]),
    "nfmc_neutra_pullback": (C.c_int, [P(RealNVPDesc), _vp, _vp, _vp, _vp, _i64, _vp]),
    "nfmc_potential_step": (C.c_int, [P(PotentialDesc), _vp, _i64, _f32, _vp]),
    "nfmc_dlmc_update": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _f32, _vp]),
    "nfmc_dlmc_latent_update": (C.c_int, [_vp, _vp, _f32, _i64, _vp]),
    "nfmc_jump_sample_slabs": (_i64, [_i32, _i64, _i32]),
    "nfmc_jump_sample_device": (C.c_int, [P(PotentialDesc), P(RealNVPDesc), _vp, _i64, _i32, _i32, _i32, _f32, _i32, _vp, _i32,
                                          _i32, C.c_uint64, C.c_uint64, C.c_uint64, _i64, P(StatsDesc), P(StatsDesc), _vp, _vp]),
    "nfmc_jump_workspace_bytes": (_i64, [_i32, _i64, _i64]),
    "nfmc_jump_sample_host": (C.c_int, [P(PotentialDesc), _vp, _i64, P(RealNVPDesc), _vp, _vp, _i64, _i32, _i32, _i32,
                                        _f32, _i32, _u64, _i64, _vp, _vp, _vp, _vp, _i64, _vp]),
}

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


def build(verbose: bool = False, jobs: Optional[int] = None) -> str:
    """Compile ``libnfmc_b200.so`` for sm_100a with nvcc (``make -C nfmc_b200/csrc``)."""
    jobs = jobs or max(1, min(16, os.cpu_count() or 1))
    r = subprocess.run(["make", "-C", CSRC_DIR, f"-j{jobs}"], capture_output=True, text=True)
    if r.returncode != 0:
        raise NativeError("building libnfmc_b200.so failed:\n" + r.stdout[-4000:] + "\n" + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NativeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                  f"or `make -C {CSRC_DIR} -j`.  nfmc_b200 has no CPU / eager fallback.")
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            if handle.nfmc_abi_version() != 1:
                raise NativeError("libnfmc_b200.so ABI version mismatch")
            _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise NativeError(lib().nfmc_last_error().decode())


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise NativeError("nfmc_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise NativeError(f"nfmc_b200 runs on CUDA devices only, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dev_f32(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def layout_for_dim(d: int):
    gs, e = C.c_int32(), C.c_int32()
    check(lib().nfmc_layout_for_dim(d, C.byref(gs), C.byref(e)))
    return gs.value, e.value


def rng_desc(seed: int, step0: int, normals: Optional[torch.Tensor] = None, uniforms: Optional[torch.Tensor] = None) -> RngDesc:
    return RngDesc(seed & 0xFFFFFFFFFFFFFFFF, step0, None if normals is None else normals.data_ptr(),
                   None if uniforms is None else uniforms.data_ptr())
