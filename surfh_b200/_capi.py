"""ctypes binding of libsurfh_b200.so (include/surfh_b200.h).

There is deliberately no fallback: if the library is missing or no CUDA device is present,
every entry point fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

LIB_NAME = "libsurfh_b200.so"
# SURFH_B200_LIB points at an experimental build of the same library (kernel A/B runs); there is still
# no fallback: whichever path is named must exist.
LIB_PATH = os.environ.get("SURFH_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

F32, F64 = 0, 1
ADJ_EXACT, ADJ_REFERENCE = 0, 1
CG_NSCALARS = 8
ABI_VERSION = 5
SPECTRAL_LSF, SPECTRAL_BETA_SUM = 0, 1
FFT_BACKENDS = {"auto": 0, "cufft": 1, "own": 2}

ADJOINT_MODES = {"exact": ADJ_EXACT, "reference": ADJ_REFERENCE}

SYMBOLS = [
    "surfh_abi_version", "surfh_create", "surfh_set_otf", "surfh_add_band", "surfh_finalize", "surfh_destroy",
    "surfh_last_error", "surfh_input_size", "surfh_output_size", "surfh_workspace_bytes", "surfh_forward",
    "surfh_adjoint", "surfh_fwadj", "surfh_maps_to_cube", "surfh_forward_host", "surfh_adjoint_host",
    "surfh_cg_regularise_dot", "surfh_laplacian_axpby", "surfh_cg_start", "surfh_cg_update", "surfh_cg_refresh", "surfh_criterion_terms", "surfh_cg_dot_x_b_plus_r", "surfh_axpy_device_scalar", "surfh_precond_build", "surfh_precond_apply", "surfh_pcg_update", "surfh_pcg_direction", "surfh_shepard",
    "surfh_launch_count", "surfh_own_launch_count", "surfh_contraction_info", "surfh_profile_enable", "surfh_profile_read", "surfh_rfft2",
]


class ModelDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("n_templates", C.c_int32), ("n_alpha", C.c_int32), ("n_beta", C.c_int32),
                ("n_lambda", C.c_int32), ("chunk", C.c_int32), ("templates", C.c_void_p),
                ("fft_backend", C.c_int32)]


class CsrDesc(C.Structure):
    _fields_ = [("n_rows", C.c_int32), ("nnz", C.c_int64), ("row_pixel", C.c_void_p), ("row_ptr", C.c_void_p),
                ("col", C.c_void_p), ("val", C.c_void_p)]


class BandDesc(C.Structure):
    _fields_ = [("n_pointing", C.c_int32), ("n_slit", C.c_int32), ("na", C.c_int32), ("nb", C.c_int32),
                ("srf", C.c_int32), ("local_a", C.c_int32), ("local_b", C.c_int32), ("wave_start", C.c_int32),
                ("n_wave", C.c_int32), ("n_det", C.c_int32), ("spectral_mode", C.c_int32), ("det_start", C.c_int32),
                ("out_offset", C.c_int64),
                ("slit_a0", C.c_void_p), ("slit_b0", C.c_void_p), ("slit_w", C.c_void_p), ("lsf", C.c_void_p),
                ("grid_base", C.c_void_p), ("grid_frac", C.c_void_p),
                ("adj_exact", CsrDesc), ("adj_reference", CsrDesc)]


class SurfhError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"surfh_b200 error {code}: {message}")
        self.code = code
        self.message = message


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'` at the repository root). "
            "surfh_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    sig = {
        "surfh_abi_version": (C.c_int, []),
        "surfh_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(vp)]),
        "surfh_set_otf": (C.c_int, [vp, i32, i32, vp]),
        "surfh_add_band": (C.c_int, [vp, C.POINTER(BandDesc)]),
        "surfh_finalize": (C.c_int, [vp]),
        "surfh_destroy": (None, [vp]),
        "surfh_last_error": (C.c_char_p, [vp]),
        "surfh_input_size": (i64, [vp]),
        "surfh_output_size": (i64, [vp]),
        "surfh_workspace_bytes": (i64, [vp]),
        "surfh_forward": (C.c_int, [vp, vp, vp, vp]),
        "surfh_adjoint": (C.c_int, [vp, vp, vp, i32, vp]),
        "surfh_fwadj": (C.c_int, [vp, vp, vp, i32, vp, vp]),
        "surfh_maps_to_cube": (C.c_int, [vp, vp, vp, vp]),
        "surfh_forward_host": (C.c_int, [vp, vp, vp]),
        "surfh_adjoint_host": (C.c_int, [vp, vp, vp, i32]),
        "surfh_cg_regularise_dot": (C.c_int, [vp, vp, vp, dbl, dbl, vp, vp]),
        "surfh_laplacian_axpby": (C.c_int, [vp, vp, vp, dbl, dbl, vp]),
        "surfh_cg_start": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "surfh_cg_update": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
        "surfh_cg_refresh": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "surfh_criterion_terms": (C.c_int, [vp, vp, vp, i64, vp, vp, vp]),
        "surfh_cg_dot_x_b_plus_r": (C.c_int, [vp, vp, vp, vp, vp, vp]),
        "surfh_axpy_device_scalar": (C.c_int, [vp, vp, vp, i64, vp, i32, vp]),
        "surfh_precond_build": (C.c_int, [vp, vp, dbl, dbl, i32]),
        "surfh_precond_apply": (C.c_int, [vp, vp, vp, vp]),
        "surfh_pcg_update": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp]),
        "surfh_pcg_direction": (C.c_int, [vp, vp, vp, vp, vp, i32, vp]),
        "surfh_shepard": (C.c_int, [vp, vp, vp, i32, vp, vp, i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, vp, vp]),
        "surfh_rfft2": (C.c_int, [i32, i32, i32, i32, i32, vp, vp, vp]),
        "surfh_launch_count": (i64, [vp]),
        "surfh_own_launch_count": (i64, [vp]),
        "surfh_contraction_info": (C.c_int, [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]),
        "surfh_profile_enable": (C.c_int, [vp, i32]),
        "surfh_profile_read": (C.c_int, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(C.c_float),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.surfh_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version mismatch")
    _lib = lib
    return lib


def check(handle, code: int) -> None:
    if code != 0:
        msg = load().surfh_last_error(handle)
        raise SurfhError(code, msg.decode() if msg else "unknown error")


def ptr(arr: np.ndarray) -> int:
    return arr.ctypes.data


def csr_desc(csr, keep: list) -> CsrDesc:
    """Build a CsrDesc from geometry.Csr, keeping the contiguous arrays alive in `keep`."""
    if csr is None or csr.n_rows == 0:
        return CsrDesc(0, 0, None, None, None, None)
    pix = np.ascontiguousarray(csr.row_pixel, dtype=np.int32)
    rp = np.ascontiguousarray(csr.row_ptr, dtype=np.int64)
    col = np.ascontiguousarray(csr.col, dtype=np.int32)
    val = np.ascontiguousarray(csr.val, dtype=np.float64)
    keep.extend([pix, rp, col, val])
    return CsrDesc(csr.n_rows, csr.nnz, ptr(pix), ptr(rp), ptr(col), ptr(val))
