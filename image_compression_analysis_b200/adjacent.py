"""Device-level calls of the kernels either side of the distortion path (csrc/adjacent.cu).

Everything takes and returns device tensors (16-bit samples stored as torch.int16, as everywhere in this
package); the mirrors of the reference's functions live in quicklooks.py (RGB quicklook), baseline.py
(make_baseline_A / make_baseline_B) and transforms.py (codec wrappers).  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from ._lib import (DM_BIL, DM_BIP, DM_BSQ, DM_DIFF_MODULO, DM_DIFF_SATURATE, DM_EM_COUNT3, DM_EM_MAX, DM_EM_MEAN,
                   DM_EM_P95, DM_EM_RMS, DM_REQ_ROUND, DM_REQ_TRUNC, DmCube, check, lib)
from .engine import _DTYPE_CODES, DevicePair, _ptr, _stream_ptr

LAYOUT_CODES = {"bsq": DM_BSQ, "bip": DM_BIP, "bil": DM_BIL}
ERR_MODES = {"mean": DM_EM_MEAN, "rms": DM_EM_RMS, "count3": DM_EM_COUNT3, "max": DM_EM_MAX, "p95": DM_EM_P95}


def c_cube(t: torch.Tensor, np_dtype: str, layout: str, bands: int, rows: int, width: int,
           band_stride: Optional[int] = None) -> DmCube:
    c = DmCube()
    c.data, c.dtype, c.layout = t.data_ptr(), _DTYPE_CODES[np_dtype], LAYOUT_CODES[layout]
    c.bands, c.rows, c.width = bands, rows, width
    c.band_stride = rows * width if band_stride is None else band_stride
    return c


def _sel(sel: Sequence[int]):
    if not 1 <= len(sel) <= 4:
        raise ValueError("1..4 selected bands")
    return (C.c_int32 * len(sel))(*[int(s) for s in sel])


def first_value(np_dtype: str) -> int:
    """Sample value of histogram bin 0."""
    return -32768 if np_dtype == "int16" else 0


def band_hist(t: torch.Tensor, np_dtype: str, layout: str, bands: int, rows: int, width: int, sel: Sequence[int],
              plane: Optional[torch.Tensor] = None, plane_bit: int = 0xff) -> torch.Tensor:
    """(len(sel), 65536) int64 value histograms of the 0-based bands `sel` (dm_band_hist)."""
    hist = torch.zeros((len(sel), 65536), dtype=torch.int64, device=t.device)
    cube = c_cube(t, np_dtype, layout, bands, rows, width)
    check(lib().dm_band_hist(C.byref(cube), _sel(sel), len(sel), _ptr(plane), plane_bit, _ptr(hist), _stream_ptr()))
    return hist


def lut_bands_u8(t: torch.Tensor, np_dtype: str, layout: str, bands: int, rows: int, width: int, sel: Sequence[int],
                 luts: np.ndarray) -> torch.Tensor:
    """(len(sel), rows, width) uint8: luts[i][bin(sample)] of the 0-based bands `sel` (dm_lut_bands_u8)."""
    luts = np.ascontiguousarray(luts, dtype=np.uint8).reshape(len(sel), 65536)
    dl = torch.from_numpy(luts).to(t.device)
    out = torch.empty((len(sel), rows, width), dtype=torch.uint8, device=t.device)
    cube = c_cube(t, np_dtype, layout, bands, rows, width)
    check(lib().dm_lut_bands_u8(C.byref(cube), _sel(sel), len(sel), _ptr(dl), _ptr(out), _stream_ptr()))
    return out


def requantize(t: torch.Tensor, np_dtype: str, mode: str, k: int, nodata: Optional[int] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mode "trunc": ((u >> k) << k) with nodata samples kept; mode "round": ((u + 2^(k-1)) >> k) << k (dm_requantize)."""
    if np_dtype not in ("uint16", "int16"):
        raise TypeError("requantize handles 16-bit samples")
    if not t.is_contiguous():
        t = t.contiguous()
    out = torch.empty_like(t) if out is None else out
    check(lib().dm_requantize(_ptr(t), _ptr(out), _DTYPE_CODES[np_dtype], t.numel(),
                              DM_REQ_TRUNC if mode == "trunc" else DM_REQ_ROUND, int(k),
                              0 if nodata is None else 1, 0 if nodata is None else int(nodata), _stream_ptr()))
    return out


def scene_error(pair: DevicePair, valid: Optional[torch.Tensor], mode: str, k_bits: int) -> Tuple[torch.Tensor, float]:
    """Per-pixel error plane of make_scene_error_map (float32 (rows,width)) and its maximum (dm_scene_error)."""
    plane = torch.empty((pair.rows, pair.width), dtype=torch.float32, device=pair.ref.device)
    mx = torch.zeros(1, dtype=torch.int32, device=pair.ref.device)
    # thr = (cdf[..., -1] * 0.95).astype(np.uint32) with cdf[..., -1] == bands for every pixel (make_baseline_B.py:363-364)
    thr = int((np.uint32(pair.bands) * 0.95).astype(np.uint32))
    cp = pair.c_pair()
    check(lib().dm_scene_error(C.byref(cp), _ptr(valid), ERR_MODES[mode], int(k_bits), thr, _ptr(plane), _ptr(mx),
                               _stream_ptr()))
    return plane, float(mx.cpu().numpy().view(np.float32)[0])


def scale_plane_u8(plane: torch.Tensor, emax: int) -> torch.Tensor:
    """(np.clip(v, 0, emax) * (255.0/emax) + 0.5).astype(np.uint8) (make_baseline_B.py:417; dm_scale_plane_u8)."""
    out = torch.empty(plane.shape, dtype=torch.uint8, device=plane.device)
    scale = float(np.float32(255.0 / emax))
    check(lib().dm_scale_plane_u8(_ptr(plane), plane.numel(), float(emax), scale, _ptr(out), _stream_ptr()))
    return out


def diff1(t: torch.Tensor, np_dtype: str, inverse: bool, saturate: bool = False,
          out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Band differencing (inverse: running sum) along axis 0 of a contiguous (B, ...) cube (dm_diff1)."""
    if not t.is_contiguous():
        t = t.contiguous()
    out = torch.empty_like(t) if out is None else out
    bands = t.shape[0]
    npix = t.numel() // bands if bands else 0
    check(lib().dm_diff1(_ptr(t), _ptr(out), _DTYPE_CODES[np_dtype], DM_DIFF_SATURATE if saturate else DM_DIFF_MODULO,
                         1 if inverse else 0, bands, npix, npix, _stream_ptr()))
    return out


def interleave(t: torch.Tensor, src: str, dst: str, bands: int, rows: int, width: int) -> torch.Tensor:
    """Contiguous cube in layout `src` ("bsq" (B,H,W) / "bil" (H,B,W) / "bip" (H,W,B)) -> layout `dst` (dm_interleave)."""
    if not t.is_contiguous():
        t = t.contiguous()
    shape = {"bsq": (bands, rows, width), "bil": (rows, bands, width), "bip": (rows, width, bands)}[dst]
    out = torch.empty(shape, dtype=t.dtype, device=t.device)
    check(lib().dm_interleave(_ptr(t), _ptr(out), t.element_size(), LAYOUT_CODES[src], LAYOUT_CODES[dst], bands, rows,
                              width, _stream_ptr()))
    return out
