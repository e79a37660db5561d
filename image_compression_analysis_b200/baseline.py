"""B200 forms of the baseline builders' sample arithmetic: the step BEFORE the distortion
path.  Same function names and arguments as the reference where it has a function; `*_arrays` are the
in-memory forms.

  trunc_uint16, write_truncated_copy   tools/make_baseline_B.py:279-316   dm_requantize (DM_REQ_TRUNC)
  to_12in16                            tools/make_baseline_A.py:137-170   dm_requantize (DM_REQ_ROUND, k = 4)
  make_scene_error_map                 tools/make_baseline_B.py:324-419   dm_scene_error + dm_scale_plane_u8
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional, Tuple

import numpy as np
import torch

from . import adjacent
from .engine import DevicePair, integral_nodata, to_device
from .raster_io import open_raster


def _np_from_device(t: torch.Tensor, np_dtype: str) -> np.ndarray:
    a = t.cpu().numpy()
    return a.view(np.uint16) if np_dtype == "uint16" else a


def trunc_uint16(u16, k: int):
    """((u16 >> k) << k) (make_baseline_B.py:279-282)."""
    if k <= 0:
        return u16
    u16 = np.asarray(u16)
    if u16.dtype != np.uint16:
        raise TypeError("trunc_uint16 takes uint16 samples")
    return _np_from_device(adjacent.requantize(to_device(u16), "uint16", "trunc", k), "uint16")


def truncated_copy_arrays(arr, k: int, nodata=None) -> np.ndarray:
    """Sample content of write_truncated_copy (make_baseline_B.py:298-311): k LSBs of the uint16 view cleared,
    samples equal to nodata untouched; int16 in, int16 out."""
    arr = np.asarray(arr)
    name = arr.dtype.name
    if name not in ("uint16", "int16"):
        raise TypeError("truncated_copy handles uint16 / int16 cubes")
    t = adjacent.requantize(to_device(arr), name, "trunc", max(int(k), 0), integral_nodata(nodata, name))
    return _np_from_device(t, name)


def to_12in16_arrays(arr) -> np.ndarray:
    """(((x + 8) >> 4) << 4) in uint16 arithmetic (make_baseline_A.py:163-167)."""
    arr = np.asarray(arr).astype(np.uint16, copy=False)
    return _np_from_device(adjacent.requantize(to_device(arr), "uint16", "round", 4), "uint16")


def write_truncated_copy(input_path: Path, output_path: Path, k: int, tile: int = 512):
    """make_baseline_B.py:284-316: whole-file form (the reference walks 512x512 windows band by band)."""
    with open_raster(input_path) as src:
        data = src.read()
        nd = src.nodata
        meta = src.meta.copy()
    out = truncated_copy_arrays(data, k, nd)
    meta.update(dtype=str(data.dtype), count=data.shape[0], tiled=True, blockxsize=tile, blockysize=tile, compress=None,
                BIGTIFF="YES", nodata=nd)
    with open_raster(Path(output_path).as_posix(), "w", **meta) as dst:
        dst.write(out)
    print(f"wrote 14-in-16: {output_path} (k={k})")


def to_12in16(in_path, out_path) -> Path:
    """make_baseline_A.py:137-170."""
    in_path, out_path = Path(in_path), Path(out_path)
    with open_raster(in_path) as src:
        data = src.read()
        meta = src.meta.copy()
    meta.update(driver="GTiff", dtype="uint16")
    meta.pop("compress", None)
    out_path.parent.mkdir(parents=True, exist_ok=True)
    with open_raster(out_path.as_posix(), "w", **meta) as dst:
        dst.write(to_12in16_arrays(data))
    return out_path


def scene_error_map_pair(pair: DevicePair, valid, err_scale: str, k_bits: int, err_mode: str = "mean") -> Tuple[np.ndarray, int]:
    """make_scene_error_map's image for a device-resident pair: (uint8 (H,W), emax)."""
    if err_mode not in adjacent.ERR_MODES:
        raise ValueError(f"err_mode must be one of {sorted(adjacent.ERR_MODES)}")
    vdev = None
    if valid is not None:
        valid = np.asarray(valid)
        assert valid.shape == (pair.rows, pair.width), "mask shape must be (H, W)"
        vdev = to_device(valid.astype(bool))
    plane, global_max = adjacent.scene_error(pair, vdev, err_mode, k_bits)
    kmax = (1 << k_bits) - 1
    if err_mode == "count3":
        emax = max(1, pair.bands) if err_scale == "fixed" else max(1, int(global_max))
    else:
        emax = kmax if err_scale == "fixed" else max(1, int(np.ceil(global_max)))
    return adjacent.scale_plane_u8(plane, emax).cpu().numpy(), emax


def scene_error_map_arrays(ref, cmp, valid, err_scale: str, k_bits: int, err_mode: str = "mean",
                           layout: str = "bsq") -> Tuple[np.ndarray, int]:
    assert tuple(ref.shape) == tuple(cmp.shape), "ref16 and 14-in-16 must match in size and band count"
    return scene_error_map_pair(DevicePair.from_arrays(ref, cmp, layout), valid, err_scale, k_bits, err_mode)


def read_mask(mask_path: Optional[Path]):
    """make_baseline_B.py:318-322"""
    if not mask_path or not Path(mask_path).exists():
        return None
    with open_raster(mask_path) as m:
        return m.read(1) > 0


def make_scene_error_map(ref16_path: Path, scene14_path: Path, mask_path: Optional[Path], err_scale: str, k_bits: int,
                         out_png: Path, err_mode: str = "mean"):
    """make_baseline_B.py:324-419: 8-bit scene error map (mode max / mean / rms / p95 / count3) written as PNG."""
    from PIL import Image
    from .ingest import load_pair
    try:
        pair, _ = load_pair(ref16_path, scene14_path)
    except AssertionError:
        raise AssertionError("ref16 and 14-in-16 must match in size and band count")
    img, emax = scene_error_map_pair(pair, read_mask(mask_path), err_scale, k_bits, err_mode)
    Image.fromarray(img, mode="L").save(out_png)
    print(f"SCENE error ({err_mode}) scale=0..{emax} DN -> {out_png}")
