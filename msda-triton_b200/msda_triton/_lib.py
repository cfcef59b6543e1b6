"""ctypes binding of libmsda_b200.so (C ABI: include/msda_b200.h).

The library is the ONLY compute path for CUDA tensors: if it is missing or fails to load, the first CUDA call
raises MsdaLibraryError -- there is no Python / torch / CPU fallback for CUDA inputs anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path

_PKG_ROOT = Path(__file__).resolve().parent.parent          # msda-triton_b200/
LIB_PATH = Path(os.environ.get("MSDA_B200_LIB", _PKG_ROOT / "lib" / "libmsda_b200.so"))

ABI_VERSION = 1

DTYPE_F32, DTYPE_F16, DTYPE_BF16, DTYPE_F64 = 0, 1, 2, 3
PAD_ZEROS, PAD_BORDER = 0, 1
BWD_NEED_IMG, BWD_NEED_POINTS, BWD_NEED_WEIGHTS, BWD_DETERMINISTIC, BWD_NEED_REF = 1, 2, 4, 8, 16
BWD_VALUE_COLSUM = 32


class MsdaPeerCtx(ctypes.Structure):
    """struct msda_peer_ctx (include/msda_b200.h): every rank's device pointers, as host arrays."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("peer_pyramids", ctypes.POINTER(ctypes.c_void_p)), ("peer_partials", ctypes.POINTER(ctypes.c_void_p)),
                ("peer_flags", ctypes.POINTER(ctypes.c_void_p)), ("counters", ctypes.c_void_p)]


class MsdaProblem(ctypes.Structure):
    """struct msda_problem (include/msda_b200.h)."""
    _fields_ = [(n, ctypes.c_int64) for n in ("B", "Npix", "H", "D", "Q", "L", "K")] + [
        ("dtype", ctypes.c_int32), ("padding_mode", ctypes.c_int32), ("align_corners", ctypes.c_int32),
        ("reserved", ctypes.c_int32)]


class MsdaLibraryError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not LIB_PATH.exists():
        raise MsdaLibraryError(
            f"{LIB_PATH} not found: build it with `python msda-triton_b200/build.py` (needs nvcc, targets sm_100a). "
            "msda_triton has no fallback for CUDA tensors.")
    lib = ctypes.CDLL(str(LIB_PATH))
    vp, i64, ci, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
    pp = ctypes.POINTER(MsdaProblem)
    lib.msda_abi_version.restype = ci
    lib.msda_abi_version.argtypes = []
    lib.msda_last_error.restype = ctypes.c_char_p
    lib.msda_last_error.argtypes = []
    lib.msda_reload_tuning.restype = None
    lib.msda_reload_tuning.argtypes = []
    lib.msda_forward.restype = ci
    lib.msda_forward.argtypes = [vp, vp, vp, vp, vp, pp, vp]
    lib.msda_backward_workspace_bytes.restype = sz
    lib.msda_backward_workspace_bytes.argtypes = [pp, ci]
    lib.msda_backward.restype = ci
    lib.msda_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, pp, ci, vp, sz, vp]
    lib.msda_module_supported.restype = ci
    lib.msda_module_supported.argtypes = [pp, ci]
    lib.msda_module_colsum_supported.restype = ci
    lib.msda_module_colsum_supported.argtypes = [pp]
    lib.msda_module_colsum_offset.restype = ctypes.c_size_t
    lib.msda_module_colsum_offset.argtypes = [pp]
    lib.msda_module_forward.restype = ci
    lib.msda_module_forward.argtypes = [vp, vp, vp, vp, vp, ci, pp, vp]
    lib.msda_module_backward.restype = ci
    lib.msda_module_backward.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ci, pp, ci, vp, sz, vp]
    lib.msda_level_table.restype = ci
    lib.msda_level_table.argtypes = [vp, vp, i64, i64, vp]
    lib.msda_peer_all_gather.restype = ci
    lib.msda_peer_all_gather.argtypes = [vp, ctypes.POINTER(MsdaPeerCtx), i64, i64, vp]
    lib.msda_peer_reduce_scatter.restype = ci
    lib.msda_peer_reduce_scatter.argtypes = [vp, ctypes.POINTER(MsdaPeerCtx), i64, i64, vp]
    lib.msda_probe_gather.restype = ci
    lib.msda_probe_gather.argtypes = [vp, vp, i64, i64, ctypes.c_uint32, vp]
    lib.msda_probe_scatter.restype = ci
    lib.msda_probe_scatter.argtypes = [vp, i64, i64, ctypes.c_uint32, vp]
    got = lib.msda_abi_version()
    if got != ABI_VERSION:
        raise MsdaLibraryError(f"{LIB_PATH}: ABI version {got}, this package expects {ABI_VERSION}; rebuild the library")
    return lib


_lib_handle = None


def get_lib() -> ctypes.CDLL:
    """Loads libmsda_b200.so on first use; raises MsdaLibraryError (never falls back) if it is unavailable."""
    global _lib_handle
    if _lib_handle is None:
        _lib_handle = _load()
    return _lib_handle


def reload_tuning() -> None:
    """Has the library re-read its MSDA_B200_* knobs from the environment (it reads them once, at first use)."""
    get_lib().msda_reload_tuning()


def check(rc: int, what: str) -> None:
    """0 -> ok; negative -> argument error (ValueError); positive -> CUDA error (RuntimeError)."""
    if rc == 0:
        return
    msg = get_lib().msda_last_error().decode("utf-8", "replace")
    if rc < 0:
        raise ValueError(f"{what}: {msg} (msda error {rc})")
    raise RuntimeError(f"{what}: {msg} (cudaError {rc})")
