"""ctypes binding of libfmd_b200.so (the C ABI declared in include/fmd_b200.h).

There is NO CPU fallback: every compute entry point raises if the shared library is missing or
a tensor is not CUDA-resident.  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_uint64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(os.path.dirname(_HERE), "csrc")
LIB_PATH = os.path.join(_CSRC, "libfmd_b200.so")

F32, F16 = 0, 1
ACT_NONE, ACT_TANH, ACT_TANH_CLAMPED = 0, 1, 2
PRIOR_BONDS, PRIOR_ANGLES, PRIOR_DIHEDRALS, PRIOR_REPULSION = 0, 1, 2, 3
# kinds evaluated by fmd_priors_csr only (the fused step): polynomial bonds, the other angle forms, impropers
PRIOR_POLY_BONDS, PRIOR_POLY_ANGLES, PRIOR_RESTRICTED_ANGLES, PRIOR_RAW_ANGLES, PRIOR_IMPROPERS, PRIOR_SHIFTED_IMPROPERS = 4, 5, 6, 7, 8, 9
ANGLE_FORM = {1: 0, 5: 1, 6: 2, 7: 3}      # PRIOR_* kind -> FMD_ANGLE_* form code
IMPROPER_FORM = {8: 0, 9: 1}

_lib = None


class DenseStage(ctypes.Structure):
    """fmd_dense_stage of include/fmd_b200.h."""
    _fields_ = [("W", c_void_p), ("bias", c_void_p), ("wdt", c_int), ("N", c_int), ("epi_act", c_int),
                ("aux", c_void_p), ("auxdt", c_int), ("res", c_void_p), ("Y", c_void_p), ("ydt", c_int),
                ("round_f16", c_int)]


MAX_CHAIN = 4

_SIGS = {
    "fmd_version": ([], c_int),
    "fmd_sm_count": ([], c_int),
    "fmd_nl_count": ([c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p], c_int),
    "fmd_exclusive_scan_i32": ([c_void_p, c_void_p, c_int, c_void_p, c_void_p], c_int),
    "fmd_nl_fill": ([c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int, c_void_p, c_void_p,
                     c_int, c_void_p, c_void_p], c_int),
    "fmd_nl_reverse": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "fmd_build_csr": ([c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p], c_int),
    "fmd_dist_rbf_cutoff_fwd": ([c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_float,
                                 c_float, c_void_p, c_void_p, c_void_p], c_int),
    "fmd_rbf_bwd": ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p,
                     c_int, c_void_p], c_int),
    "fmd_edge_grad_to_pos_atomic": ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                     c_void_p, c_void_p], c_int),
    "fmd_edge_grad_to_forces_csr": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                     c_float, c_void_p, c_int, c_int, c_void_p], c_int),
    "fmd_nl_step": ([c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_int, c_void_p,
                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                     c_void_p, c_void_p], c_int),
    "fmd_cfconv_csr": ([c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                        c_int, c_float, c_void_p, c_void_p], c_int),
    "fmd_cfconv_grad_filter": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_int,
                                c_float, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p], c_int),
    "fmd_filter_cfconv_fwd": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_float, c_float, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p], c_int),
    "fmd_filter_cfconv_bwd": ([c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_float, c_float, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p],
                              c_int),
    "fmd_debug_set_trace_fwd": ([c_void_p], c_int),
    "fmd_debug_set_trace_bwd": ([c_void_p], c_int),
    "fmd_linear": ([c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                    c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p], c_int),
    "fmd_linear_tc": ([c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                       c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p], c_int),
    "fmd_linear_x3": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                       c_void_p, c_void_p], c_int),
    "fmd_linear_x3_rbf_bwd": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_float, c_float,
                               c_void_p, c_int, c_void_p], c_int),
    "fmd_linear_chain_tc": ([c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(DenseStage), c_int, c_void_p],
                            c_int),
    "fmd_embedding": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p], c_int),
    "fmd_out_head": ([c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p], c_int),
    "fmd_segment_sum": ([c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p], c_int),
    "fmd_prior_energy_forces": ([c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 c_void_p, c_void_p, c_void_p], c_int),
    "fmd_priors_csr": ([c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                        c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                        c_int, c_void_p], c_int),
    "fmd_baoab_pre": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p,
                       c_uint64, c_int, c_float, c_float, c_float, c_void_p], c_int),
    "fmd_overdamped_step": ([c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_uint64, c_int,
                             c_void_p], c_int),
    "fmd_increment_u64": ([c_void_p, c_void_p], c_int),
    "fmd_baoab_post": ([c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p],
                       c_int),
    "fmd_philox_normal": ([c_uint64, c_uint64, c_int, c_void_p, c_void_p], c_int),
    "fmd_pt_decide": ([c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_uint64, c_uint64, c_void_p,
                       c_void_p], c_int),
    "fmd_pt_swap": ([c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p], c_int),
}

EXPORTED = tuple(_SIGS.keys()) + ("fmd_last_error",)


def load(build_if_missing: bool = False):
    """dlopen the library (once).  Raises RuntimeError when it is absent — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            import importlib.util
            spec = importlib.util.spec_from_file_location("fmd_build", os.path.join(_CSRC, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        else:
            raise RuntimeError(
                f"flashmd: CUDA library {LIB_PATH} not found. Build it with "
                f"`python {os.path.join(_CSRC, 'build.py')}` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU fallback for the kernel path.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    lib.fmd_last_error.argtypes = []
    lib.fmd_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def is_loaded() -> bool:
    return _lib is not None


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("flashmd kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("flashmd kernels need contiguous tensors")
    return t.data_ptr()


def dt_code(t) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise RuntimeError(f"unsupported dtype {t.dtype}")


def idx_bytes(t) -> int:
    if t.dtype == torch.int64:
        return 8
    if t.dtype == torch.int32:
        return 4
    raise RuntimeError(f"index tensors must be int32 or int64, got {t.dtype}")


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = _lib.fmd_last_error().decode() if _lib is not None else ""
        raise RuntimeError(f"libfmd_b200 {what} failed (code {rc}): {msg}")


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
