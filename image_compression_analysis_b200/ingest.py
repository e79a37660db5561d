"""Read an original / decoded GeoTIFF pair once and keep it resident in HBM.

The reference re-opens and re-reads both files in each of its three metric calls per rep
(run_codec.py:518, :522, :526 -- and once more inside effective_data_range, :245), and the original
is the same file for every rate and rep of a sweep.  Here a cube is read once into a reusable PINNED
staging buffer, copied to the device, and cached by (path, mtime, size); the three drop-in functions
of one rep then share the device-resident pair, and the original stays resident across the sweep.

With the built-in GeoTIFF reader a pixel-interleaved file is uploaded as it is stored ((H,W,B), "bip"),
which is the layout the one-pass BIP kernel wants; rasterio always de-interleaves to (B,H,W), and many-band
cubes that arrive that way are transposed once on the device.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .engine import DevicePair, dtype_code, integral_nodata, require_cuda
from .raster_io import explicit_mask, open_raster

MAX_CACHED_CUBES = 3          # original + the decoded cube of the current rep (+ one spare)


@dataclass
class Cube:
    tensor: torch.Tensor      # device, int16 storage for 16-bit samples
    np_dtype: str
    layout: str               # "bsq" (B,H,W) | "bip" (H,W,B)
    bands: int
    rows: int
    width: int
    nodata: Optional[float]
    mask: Optional[np.ndarray]        # explicit dataset mask (alpha / .msk) as bool (H,W), else None
    meta: dict


_CUBES: "OrderedDict[Tuple[str, int, int, int], Cube]" = OrderedDict()
_STAGING: Dict[int, torch.Tensor] = {}
_STAGING_BUSY: Dict[int, "torch.cuda.Event"] = {}      # id(staging buffer) -> "the upload that read it has finished"
STATS = {"reads": 0, "hits": 0}


def clear_cache() -> None:
    _CUBES.clear()
    _STAGING.clear()
    _STAGING_BUSY.clear()


def _key(path) -> Tuple[str, int, int, int]:
    """(path, mtime, size, CUDA device): a cube cached on one GPU is not handed to code running on another."""
    p = Path(path)
    st = os.stat(p)
    return (str(p.resolve()), st.st_mtime_ns, st.st_size, torch.cuda.current_device() if torch.cuda.is_available() else -1)


def _wait_staging_free(stage: torch.Tensor) -> None:
    """Block until the upload that last read `stage` has finished -- whichever stream or device issued it."""
    ev = _STAGING_BUSY.pop(id(stage), None)
    if ev is not None:
        ev.synchronize()


def _staging(nbytes: int) -> torch.Tensor:
    """A reusable pinned host buffer of at least nbytes (pinning memory is slow: do it once per size)."""
    for n, t in _STAGING.items():
        if n >= nbytes:
            return t
    _STAGING.clear()
    t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    _STAGING[nbytes] = t
    return t


def load_cube(path) -> Cube:
    """The cube of a GeoTIFF on the current device (cached)."""
    try:
        key = _key(path)
    except OSError:
        key = None                      # not a real file (e.g. an in-memory stand-in): no caching
    if key is not None and key in _CUBES:
        _CUBES.move_to_end(key)
        STATS["hits"] += 1
        return _CUBES[key]
    dev = require_cuda()
    STATS["reads"] += 1
    with open_raster(path) as ds:
        B, H, W = int(ds.count), int(ds.height), int(ds.width)
        name = np.dtype(ds.dtypes[0]).name
        dtype_code(name)
        tdt = torch.int16 if np.dtype(name).itemsize == 2 else torch.uint8
        if hasattr(ds, "read_native"):
            # the built-in reader decodes straight into the pinned staging buffer
            shape, layout = ds.native_shape()
            nbytes = int(np.prod(shape)) * np.dtype(name).itemsize
            stage = _staging(nbytes)
            # the staging buffer is reused by the next read: wait for the previous upload before overwriting it
            _wait_staging_free(stage)
            host = stage[:nbytes].view(tdt).view(shape)
            ds.read_native(out=host.numpy().view(np.dtype(name)))
        else:
            arr, layout = np.ascontiguousarray(ds.read()), "bsq"
            if arr.dtype == np.uint16:
                arr = arr.view(np.int16)
            stage = _staging(arr.nbytes)
            _wait_staging_free(stage)
            host = stage[:arr.nbytes].view(tdt).view(arr.shape)
            host.numpy()[...] = arr
        nodata, mask, meta = ds.nodata, explicit_mask(ds), ds.meta.copy()
    t = host.to(dev, non_blocking=True)
    busy = torch.cuda.Event()
    busy.record()
    _STAGING_BUSY[id(stage)] = busy
    if layout == "bsq" and B >= 16 and B % 4 == 0 and np.dtype(name).itemsize == 2:
        # rasterio always de-interleaves to (B,H,W).  Many-band cubes are evaluated from (H,W,B): the one-pass
        # kernel, the register-resident SID kernel and the BIP Sobel kernel all want whole spectra together,
        # and one device transpose (0.22 ms for a Case-B cube) is cheaper than what the BSQ kernels lose.
        from . import adjacent
        t = adjacent.interleave(t, "bsq", "bip", B, H, W)
        layout = "bip"
    cube = Cube(t, name, layout, B, H, W, nodata, mask, meta)
    if key is not None:
        _CUBES[key] = cube
        while len(_CUBES) > MAX_CACHED_CUBES:
            _CUBES.popitem(last=False)
    return cube


def load_pair(ref_path, tst_path) -> Tuple[DevicePair, dict]:
    """Both cubes as one DevicePair + what the callers need besides (masks, shape)."""
    a, b = load_cube(ref_path), load_cube(tst_path)
    assert a.bands == b.bands and a.width == b.width and a.rows == b.rows, \
        "Reference and test must match in size and band count."
    assert a.np_dtype == b.np_dtype, "Reference and test must have the same sample type."
    tb = b.tensor
    if b.layout != a.layout:            # one file chunky, the other planar: bring the decoded cube to the original's layout
        tb = (tb.permute(1, 2, 0) if a.layout == "bip" else tb.permute(2, 0, 1)).contiguous()
    pair = DevicePair(a.tensor, tb, a.np_dtype, a.layout, a.bands, a.rows, a.width,
                      integral_nodata(a.nodata, a.np_dtype), integral_nodata(b.nodata, b.np_dtype))
    info = dict(ref_nodata=a.nodata, tst_nodata=b.nodata, ref_mask=a.mask, tst_mask=b.mask, H=a.rows, W=a.width,
                B=a.bands, ref_meta=a.meta)
    return pair, info
