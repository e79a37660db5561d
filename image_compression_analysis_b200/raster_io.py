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


def has_explicit_mask(ds) -> bool:
    """Does the dataset carry a per-dataset or alpha mask?  rasterio's dataset_mask() gives such a mask priority
    over the nodata value (run_codec.py:249, quicklooks.py:37), so it must be honoured whether or not the file
    also has nodata.  rasterio: mask_flag_enums; built-in reader: .msk sidecar / internal mask directory."""
    if hasattr(ds, "has_explicit_mask"):
        return bool(ds.has_explicit_mask())
    flags = getattr(ds, "mask_flag_enums", None)
    if flags is not None:
        try:
            names = {getattr(f, "name", str(f)) for band in flags for f in band}
            return bool(names & {"per_dataset", "alpha"})
        except TypeError:
            return False
    return False


def explicit_mask(ds) -> Optional[np.ndarray]:
    """dataset_mask() as bool when it carries information beyond nodata (alpha band / .msk sidecar / internal
    mask), else None.  A nodata-derived dataset mask is NOT read here: the GPU derives it from the samples.
    A file with BOTH an explicit mask and a nodata value returns the explicit mask (it has priority in
    rasterio); the nodata value still goes to the GPU, where the per-band nodata tests of compute_metrics
    (run_codec.py:250-259) and of _valid_mask_from_ds (quicklooks.py:41-42) imply the nodata-derived
    dataset mask, so ANDing the explicit mask on top reproduces the reference's masks."""
    nd = ds.nodata
    has_nd = nd is not None and np.isfinite(nd)
    if has_nd and not has_explicit_mask(ds):
        return None
    m = ds.dataset_mask()
    if m is None:
        return None
    m = np.asarray(m) > 0
    return None if bool(m.all()) else m
