"""B200 forms of the codec wrappers' reversible transforms: the steps either side of the
external encoder.  Same names and arguments as the reference's helpers; arrays in, arrays out.

  _diff1_bsq_signed / _int1_bsq_signed / _diff1_bsq_unsigned / _int1_bsq_unsigned
                                      tools/codecs/ccsds121/ccsds121_wrap.py:66-85   dm_diff1 (modulo 2^16)
  _diff1_forward / _diff1_inverse     tools/codecs/jpegls/jpegls_wrap.py:92-120      dm_diff1 (modulo / int16 saturating)
  _write_raw_interleaved / _read_raw_interleaved
                                      tools/codecs/ccsds121/ccsds121_wrap.py:44-64   dm_interleave
"""
from __future__ import annotations

import numpy as np
import torch

from . import adjacent
from .engine import to_device


def _host(t: torch.Tensor, dtype) -> np.ndarray:
    a = t.cpu().numpy()
    return a.view(np.dtype(dtype)) if a.dtype != np.dtype(dtype) else a


def _diff1_bsq_signed(tile_bsq: np.ndarray) -> np.ndarray:
    """R[b] = X[b] - X[b-1] mod 2^16 on the uint16 view, returned as int16 (ccsds121_wrap.py:66-69)."""
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(tile_bsq).view(np.int16)), "int16", inverse=False), np.int16)


def _int1_bsq_signed(R: np.ndarray) -> np.ndarray:
    """Running sum mod 2^16 (ccsds121_wrap.py:71-74)."""
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(R).view(np.int16)), "int16", inverse=True), np.int16)


def _diff1_bsq_unsigned(tile_bsq: np.ndarray) -> np.ndarray:
    """ccsds121_wrap.py:76-79"""
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(tile_bsq, dtype=np.uint16)), "uint16", inverse=False), np.uint16)


def _int1_bsq_unsigned(R: np.ndarray) -> np.ndarray:
    """ccsds121_wrap.py:81-84"""
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(R, dtype=np.uint16)), "uint16", inverse=True), np.uint16)


def _pair_op(cur, prev, dtype_str: str, inverse: bool):
    if dtype_str not in ("uint16", "int16", "uint8"):
        return cur
    dt = np.dtype(dtype_str)
    two = np.stack([np.asarray(prev).astype(dt, copy=False), np.asarray(cur).astype(dt, copy=False)], 0)
    out = adjacent.diff1(to_device(two), dtype_str, inverse=inverse, saturate=(dtype_str == "int16"))
    return _host(out[1], dt)


def _diff1_forward(cur: np.ndarray, prev, dtype_str: str) -> np.ndarray:
    """R = X[b] - X[b-1] (mod 2^N; int16 clipped) for one band (jpegls_wrap.py:92-106)."""
    if prev is None:
        return cur
    return _pair_op(cur, prev, dtype_str, inverse=False)


def _diff1_inverse(R: np.ndarray, prev_recon, dtype_str: str) -> np.ndarray:
    """X[b] = R[b] + X[b-1] (mod 2^N; int16 clipped) for one band (jpegls_wrap.py:108-120)."""
    if prev_recon is None:
        return R
    return _pair_op(R, prev_recon, dtype_str, inverse=True)


def diff1_cube_forward(cube: np.ndarray, dtype_str: str) -> np.ndarray:
    """The JPEG-LS wrapper's whole band loop in lossless mode in one launch: band b against ORIGINAL band b-1."""
    dt = np.dtype(dtype_str)
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(cube, dtype=dt)), dtype_str, inverse=False,
                                saturate=(dtype_str == "int16")), dt)


def diff1_cube_inverse(res: np.ndarray, dtype_str: str) -> np.ndarray:
    """... and its inverse: band b from the RECONSTRUCTED band b-1."""
    dt = np.dtype(dtype_str)
    return _host(adjacent.diff1(to_device(np.ascontiguousarray(res, dtype=dt)), dtype_str, inverse=True,
                                saturate=(dtype_str == "int16")), dt)


def interleave_arrays(cube: np.ndarray, src: str, dst: str, B: int, Ht: int, Wt: int) -> np.ndarray:
    """Contiguous cube in layout src -> layout dst ("bsq" / "bil" / "bip") on the device."""
    a = np.ascontiguousarray(cube)
    return _host(adjacent.interleave(to_device(a), src, dst, B, Ht, Wt), a.dtype)


def _write_raw_interleaved(tile_bsq, interleave, out_path, np_dtype):
    """ccsds121_wrap.py:44-56 / ccsds123_wrap.py:43-56"""
    if interleave not in ("bsq", "bil", "bip"):
        raise ValueError("interleave must be one of: bsq, bil, bip")
    B, Ht, Wt = tile_bsq.shape
    arr = np.ascontiguousarray(tile_bsq).astype(np_dtype, copy=False)
    out = arr if interleave == "bsq" else interleave_arrays(arr, "bsq", interleave, B, Ht, Wt)
    with open(out_path, "wb") as f:
        out.tofile(f)


def _read_raw_interleaved(in_path, interleave, np_dtype, B, Ht, Wt):
    """ccsds121_wrap.py:58-64: returns the (B,Ht,Wt) cube (a contiguous copy, where the reference returns a view)."""
    if interleave not in ("bsq", "bil", "bip"):
        raise ValueError("interleave must be one of: bsq, bil, bip")
    arr = np.fromfile(in_path, dtype=np_dtype)
    if arr.size != B * Ht * Wt:
        raise RuntimeError("Unexpected RAW size")
    if interleave == "bsq":
        return arr.reshape(B, Ht, Wt)
    return interleave_arrays(arr, interleave, "bsq", B, Ht, Wt)
