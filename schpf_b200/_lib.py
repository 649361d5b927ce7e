"""ctypes binding of libschpf_b200.so (the C ABI declared in include/schpf_b200.h).

There is no CPU fallback: if the library has not been built, or no CUDA device
is usable, the product path raises.  Build with ``python -m schpf_b200.build``
(or ``__graft_entry__.build()``).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCHPF_B200_LIB") or os.path.join(_HERE, "_C", "libschpf_b200.so")

c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_dbl = ctypes.c_double
c_u64 = ctypes.c_uint64
c_vp = ctypes.c_void_p
p_dbl = ctypes.POINTER(ctypes.c_double)
p_i32 = ctypes.POINTER(ctypes.c_int32)

# name -> argtypes; every function returns int except where noted
SIGNATURES = {
    "schpf_version": [],
    "schpf_last_error": [],
    "schpf_device_count": [],
    "schpf_release_scratch": [c_int],
    "schpf_psi": [c_int, c_i64, p_dbl, p_dbl],
    "schpf_gammaln": [c_int, c_i64, p_dbl, p_dbl],
    "schpf_compute_Xphi_data": [c_int, c_i64, c_i64, c_i64, c_int, p_i32, p_i32, p_i32,
                                p_dbl, p_dbl, p_dbl, p_dbl, p_dbl],
    "schpf_compute_loading_shape_update": [c_int, c_i64, c_int, p_dbl, p_i32, c_i64, c_dbl, p_dbl],
    "schpf_compute_loading_rate_update": [c_int, c_i64, c_i64, c_int, p_dbl, p_dbl, p_dbl, p_dbl, p_dbl],
    "schpf_compute_capacity_rate_update": [c_int, c_i64, c_int, p_dbl, p_dbl, c_dbl, p_dbl],
    "schpf_compute_pois_llh": [c_int, c_i64, c_i64, c_i64, c_int, p_i32, p_i32, p_i32,
                               p_dbl, p_dbl, p_dbl, p_dbl, p_dbl],
    "schpf_create": [ctypes.POINTER(c_vp), c_int, c_i64, c_i64, c_int, c_vp],
    "schpf_destroy": [c_vp],
    "schpf_set_option": [c_vp, ctypes.c_char_p, c_i64],
    "schpf_set_coo": [c_vp, c_vp, c_vp, c_vp, c_i64],
    "schpf_set_coo_device": [c_vp, c_vp, c_vp, c_vp, c_i64],
    "schpf_set_hyper": [c_vp, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl, c_dbl],
    "schpf_set_state": [c_vp] + [c_vp] * 8,
    "schpf_get_state": [c_vp] + [c_vp] * 8,
    "schpf_copy_gene_state": [c_vp, c_vp],
    "schpf_copy_cell_state": [c_vp, c_i64, c_vp, c_i64, c_i64],
    "schpf_step": [c_vp, c_int, c_int],
    "schpf_step_with_xphi": [c_vp, p_dbl, c_int],
    "schpf_step_random_phi": [c_vp, c_u64, c_int],
    "schpf_step_begin": [c_vp, c_int, c_int, c_u64],
    "schpf_exchange_buffer": [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64)],
    "schpf_step_end": [c_vp, c_int],
    "schpf_comm_unique_id": [ctypes.c_char_p],
    "schpf_comm_create": [ctypes.POINTER(c_vp), c_int, ctypes.c_char_p, c_int, c_int],
    "schpf_comm_destroy": [c_vp],
    "schpf_comm_attach": [c_vp, c_vp],
    "schpf_loss": [c_vp, p_dbl],
    "schpf_loss_parts": [c_vp, p_dbl, ctypes.POINTER(c_i64)],
    "schpf_llh_pointwise": [c_vp, p_dbl],
    "schpf_xphi_debug": [c_vp, p_dbl],
    "schpf_layout_dump": [c_vp, c_int, c_i64, p_i32, p_i32, p_i32, ctypes.POINTER(c_i64)],
    "schpf_count_lines": [c_int, c_vp, c_vp, c_i64, c_i64, ctypes.POINTER(c_i64)],
    "schpf_parse_triples": [c_int, c_vp, c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_i64,
                            ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)],
    "schpf_synchronize": [c_vp],
    "schpf_counter": [c_vp, ctypes.c_char_p, p_dbl],
}

FREEZE_GENES = 1
SIMULTANEOUS = 2
CELLS_FIRST = 4
PHASE_CELLS = 8
PHASE_GENES = 16
MAX_FACTORS = 64

_lib = None


class SchpfError(RuntimeError):
    pass


def load():
    """Load the shared library (once).  Raises if it is missing: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SchpfError(
            "schpf_b200: %s not found -- build it with `python -m schpf_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = ctypes.c_char_p if name == "schpf_last_error" else c_int
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().schpf_last_error()
        raise SchpfError("schpf_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def to_host(a):
    """numpy view of a host array, or a host copy of a CUDA tensor (io.DeviceCOO fields)"""
    if hasattr(a, "is_cuda"):
        return a.cpu().numpy()
    return np.asarray(a)


def as_i32(a, what="index"):
    a = to_host(a)
    if a.dtype != np.int32:
        if a.size and (a.min() < np.iinfo(np.int32).min or a.max() > np.iinfo(np.int32).max):
            raise ValueError("%s values do not fit int32" % what)
        if a.dtype.kind == "f" and a.size and not np.all(a == np.floor(a)):
            raise ValueError("%s values must be integers" % what)
    return np.ascontiguousarray(a, dtype=np.int32)


def dptr(a):
    return a.ctypes.data_as(p_dbl)


def iptr(a):
    return a.ctypes.data_as(p_i32)
