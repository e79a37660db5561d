"""ctypes binding of libdm_b200.so (include/dm_b200.h).

There is no CPU path: if the library is missing, or a call fails, this module raises.  The
library is built in-tree by `python image_compression_analysis_b200/csrc/build.py`
(or `__graft_entry__.build()`), never JIT-compiled at import time.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

ABI_VERSION = 9

DM_OK, DM_EARG, DM_ECUDA, DM_EUNSUPPORTED = 0, 1, 2, 3
DM_U8, DM_U16, DM_I16 = 0, 1, 2
DM_BSQ, DM_BIP, DM_BIL = 0, 1, 2
DM_VALID_METRICS, DM_VALID_QUICKLOOK, DM_VALID_SPECTRAL = 1, 2, 4
DM_NSTAT = 8
DM_S_N, DM_S_X, DM_S_Y, DM_S_XX, DM_S_YY, DM_S_XY, DM_S_ABS, DM_S_SSE = range(8)
DM_M_MAXERR, DM_M_ABSXY, DM_M_UMAX, DM_M_UNEGMIN, DM_M_LOW4, DM_M_LOW2 = range(6)
DM_STATS_NO_MOMENTS, DM_STATS_GENERIC = 1, 2
DM_REQ_TRUNC, DM_REQ_ROUND = 0, 1
DM_EM_MEAN, DM_EM_RMS, DM_EM_COUNT3, DM_EM_MAX, DM_EM_P95 = range(5)
DM_DIFF_MODULO, DM_DIFF_SATURATE = 0, 1

LIB_PATH = Path(__file__).resolve().parent / "libdm_b200.so"


class DmPair(C.Structure):
    """dm_pair_t"""
    _fields_ = [
        ("ref", C.c_void_p), ("tst", C.c_void_p),
        ("dtype", C.c_int32), ("layout", C.c_int32),
        ("bands", C.c_int64), ("rows", C.c_int64), ("width", C.c_int64), ("band_stride", C.c_int64),
        ("ref_has_nodata", C.c_int32), ("ref_nodata", C.c_int32),
        ("tst_has_nodata", C.c_int32), ("tst_nodata", C.c_int32),
    ]


class DmCube(C.Structure):
    """dm_cube_t"""
    _fields_ = [
        ("data", C.c_void_p), ("dtype", C.c_int32), ("layout", C.c_int32),
        ("bands", C.c_int64), ("rows", C.c_int64), ("width", C.c_int64), ("band_stride", C.c_int64),
    ]


class DmError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libdm_b200 error {code}: {text}")
        self.code = code


# name -> (restype, argtypes); every symbol include/dm_b200.h declares
_P = C.c_void_p
SYMBOLS = {
    "dm_abi_version": (C.c_int, []),
    "dm_last_error": (C.c_char_p, []),
    "dm_device_sm_count": (C.c_int, []),
    "dm_launch_count": (C.c_int64, []),
    "dm_launch_chaining": (None, [C.c_int32]),
    "dm_fused_bip_variant": (C.c_int, [C.c_int32]),
    "dm_validity": (C.c_int, [C.POINTER(DmPair), _P, _P, _P, _P]),
    "dm_fused_stats": (C.c_int, [C.POINTER(DmPair), _P, C.c_int32, C.c_int32, C.c_uint32, _P, _P, _P, _P]),
    "dm_fused_stats_batch": (C.c_int, [C.POINTER(DmPair), _P, C.c_int32, C.c_uint32, _P]),
    "dm_workspace_bytes": (C.c_int64, []),
    "dm_spectral": (C.c_int, [C.POINTER(DmPair), _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P,
                              C.c_int32, C.c_int32, _P, _P, _P]),
    "dm_fused_bip": (C.c_int, [C.POINTER(DmPair), _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P,
                               C.c_int32, _P, _P, _P]),
    "dm_fused_bip_scan": (C.c_int, [C.POINTER(DmPair), _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P,
                                    C.c_int32, _P, _P, _P]),
    "dm_fused_bsq": (C.c_int, [C.POINTER(DmPair), _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "dm_sobel_mag": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "dm_sobel_nblocks": (C.c_int, []),
    "dm_sobel_lmse": (C.c_int, [C.POINTER(DmPair), C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, _P]),
    "dm_ssim_nblocks": (C.c_int, []),
    "dm_ssim_variant": (C.c_int, [C.c_int32]),
    "dm_spectral_lanes_per_pixel": (C.c_int, [C.c_int32]),
    "dm_ssim_gauss": (C.c_int, [C.POINTER(DmPair), C.c_double, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, _P, _P]),
    "dm_combine_partials": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P, _P]),
    "dm_bip_to_bsq": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int64, C.c_int64, _P]),
    "dm_band_hist": (C.c_int, [C.POINTER(DmCube), C.POINTER(C.c_int32), C.c_int32, _P, C.c_int32, _P, _P]),
    "dm_lut_bands_u8": (C.c_int, [C.POINTER(DmCube), C.POINTER(C.c_int32), C.c_int32, _P, _P, _P]),
    "dm_requantize": (C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "dm_scene_error": (C.c_int, [C.POINTER(DmPair), _P, C.c_int32, C.c_int32, C.c_uint32, _P, _P, _P]),
    "dm_scale_plane_u8": (C.c_int, [_P, C.c_int64, C.c_float, C.c_float, _P, _P]),
    "dm_diff1": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, _P]),
    "dm_interleave": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64, _P]),
    "dm_p2p_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), _P]),
    "dm_p2p_open": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "dm_p2p_close": (C.c_int, [_P]),
    "dm_p2p_free": (C.c_int, [_P]),
    "dm_p2p_zero": (C.c_int, [_P, C.c_int64, _P]),
    "dm_p2p_push": (C.c_int, [_P, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int32, C.c_uint64, _P]),
    "dm_p2p_combine": (C.c_int, [_P, _P, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_int64, _P, _P, C.c_double, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python image_compression_analysis_b200/csrc/build.py` "
                "(the distortion metrics have no CPU implementation)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)          # AttributeError if the export is missing
            fn.restype, fn.argtypes = res, args
        got = handle.dm_abi_version()
        if got != ABI_VERSION:
            raise ImportError(f"libdm_b200.so ABI {got} != binding ABI {ABI_VERSION}; rebuild the library")
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != DM_OK:
        raise DmError(code, (lib().dm_last_error() or b"").decode("utf-8", "replace"))
