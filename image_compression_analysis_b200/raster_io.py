"""Raster access for the path-level entry points (the reference reads GeoTIFFs through rasterio:
run_codec.py:242, quicklooks.py:122).  rasterio is used when importable; otherwise the minimal
reader/writer in geotiff.py (uncompressed / DEFLATE baseline GeoTIFF) takes over."""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import numpy as np


def _rasterio():
    try:
        import rasterio  # noqa: F401
        return rasterio
    except ImportError:
        return None


def open_raster(path, mode: str = "r", **meta):
    rio = _rasterio()
    if rio is not None:
        return rio.open(Path(path).as_posix() if mode != "r" else path, mode, **meta) if mode != "r" else rio.open(path)
    from . import geotiff
    return geotiff.open(path, mode, **meta)


def uint8_dtype():
    rio = _rasterio()
    return rio.uint8 if rio is not None else "uint8"


def explicit_mask(ds) -> Optional[np.ndarray]:
    """dataset_mask() as bool when it carries information beyond nodata (alpha band / .msk);
    None when everything is valid.  With a nodata value the mask is derived on the GPU instead."""
    nd = ds.nodata
    if nd is not None and np.isfinite(nd):
        return None
    m = ds.dataset_mask()
    if m is None:
        return None
    m = np.asarray(m) > 0
    return None if bool(m.all()) else m
