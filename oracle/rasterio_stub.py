"""In-memory stand-in for the `rasterio` module -- TEST INFRASTRUCTURE ONLY.

The reference (`/root/reference/tools/run_codec.py:32`, `tools/quicklooks.py:27`)
imports rasterio at module top; rasterio is not installed in this image.  This
stub implements only the API surface the reference's hot path touches
 on top of numpy arrays held in a process-local registry,
so the reference's metric functions can be executed UNMODIFIED to pin the
oracle and to generate the golden fixtures under tests/golden/.

Nothing in the product package imports this file.

dataset_mask() follows rasterio's documented semantics (third-party, not in
/root/reference): an explicit per-dataset mask wins; otherwise, with a nodata
value, a pixel is valid (255) where ANY band differs from nodata; otherwise
everything is valid.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np

_REGISTRY: dict[str, "MemRaster"] = {}


def _key(path) -> str:
    return Path(str(path)).as_posix()


class MemRaster:
    """One in-memory raster: (B,H,W) samples + optional nodata + optional mask."""

    def __init__(self, data, nodata=None, mask=None, descriptions=None):
        data = np.asarray(data)
        if data.ndim == 2:
            data = data[None]
        assert data.ndim == 3
        self.data = data
        self.nodata = None if nodata is None else float(nodata)
        self.mask = None if mask is None else (np.asarray(mask) > 0)
        self.descriptions = descriptions
        self.tags: dict[str, str] = {}
        self.meta_extra: dict = {}


def register(path, data, nodata=None, mask=None, descriptions=None) -> str:
    k = _key(path)
    _REGISTRY[k] = MemRaster(data, nodata=nodata, mask=mask, descriptions=descriptions)
    return k


def fetch(path) -> MemRaster:
    return _REGISTRY[_key(path)]


def clear() -> None:
    _REGISTRY.clear()


class _Reader:
    def __init__(self, key: str, r: MemRaster, writable: bool = False):
        self._key, self._r, self._writable = key, r, writable

    # context manager -----------------------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass

    # attributes ----------------------------------------------------------
    @property
    def count(self):
        return int(self._r.data.shape[0])

    @property
    def height(self):
        return int(self._r.data.shape[1])

    @property
    def width(self):
        return int(self._r.data.shape[2])

    @property
    def dtypes(self):
        return tuple(str(self._r.data.dtype) for _ in range(self.count))

    @property
    def nodata(self):
        return self._r.nodata

    @nodata.setter
    def nodata(self, v):
        assert self._writable
        self._r.nodata = None if v is None else float(v)

    @property
    def descriptions(self):
        d = self._r.descriptions
        return tuple(d) if d else tuple(None for _ in range(self.count))

    @property
    def meta(self):
        m = {
            "driver": "GTiff",
            "dtype": str(self._r.data.dtype),
            "nodata": self._r.nodata,
            "width": self.width,
            "height": self.height,
            "count": self.count,
            "crs": None,
            "transform": None,
        }
        return m

    @property
    def nodatavals(self):
        return tuple(self._r.nodata for _ in range(self.count))

    @property
    def profile(self):
        m = self.meta
        m.update(tiled=False)
        return m

    @property
    def block_shapes(self):
        return [(min(512, self.height), min(512, self.width)) for _ in range(self.count)]

    # pixel access --------------------------------------------------------
    def read(self, indexes=None, out_dtype=None, window=None):
        d = self._r.data
        if window is not None:
            d = d[:, window.row_off:window.row_off + window.height, window.col_off:window.col_off + window.width]
        if indexes is None:
            out = d.copy()
        elif isinstance(indexes, (list, tuple)):
            out = np.stack([d[int(i) - 1] for i in indexes], 0)
        else:
            out = d[int(indexes) - 1].copy()
        if out_dtype is not None:
            out = out.astype(out_dtype)
        return out

    def dataset_mask(self):
        r = self._r
        if r.mask is not None:
            m = r.mask
        elif r.nodata is not None and np.isfinite(r.nodata):
            m = np.any(r.data != r.nodata, axis=0)
        else:
            m = np.ones(r.data.shape[1:], bool)
        return m.astype(np.uint8) * 255


class _Writer:
    def __init__(self, key: str, meta: dict):
        self._key, self._meta = key, dict(meta)
        self._data = None
        self._mask = None
        self._tags: dict[str, str] = {}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        r = MemRaster(self._data if self._data is not None else
                      np.zeros((self._meta.get("count", 1), self._meta.get("height", 0),
                                self._meta.get("width", 0)), self._meta.get("dtype", "uint8")),
                      nodata=self._meta.get("nodata"), mask=self._mask)
        r.tags = dict(self._tags)
        r.meta_extra = dict(self._meta)
        _REGISTRY[self._key] = r
        return False

    def write(self, arr, indexes=None, window=None):
        arr = np.asarray(arr)
        if window is None and indexes is None:
            if arr.ndim == 2:
                arr = arr[None]
            self._data = arr.copy()
            return
        # band-wise / windowed writes (the baseline builders): fill a full-size array piece by piece
        if self._data is None:
            self._data = np.zeros((self._meta.get("count", 1), self._meta["height"], self._meta["width"]),
                                  self._meta.get("dtype", arr.dtype))
        r0, c0 = (window.row_off, window.col_off) if window is not None else (0, 0)
        if indexes is None:
            if arr.ndim == 2:
                arr = arr[None]
            self._data[:, r0:r0 + arr.shape[1], c0:c0 + arr.shape[2]] = arr
        else:
            self._data[int(indexes) - 1, r0:r0 + arr.shape[0], c0:c0 + arr.shape[1]] = arr

    def write_mask(self, m):
        self._mask = np.asarray(m) > 0

    def update_tags(self, **kw):
        self._tags.update({k: str(v) for k, v in kw.items()})


def _open(path, mode="r", **meta):
    k = _key(path)
    if mode == "w":
        return _Writer(k, meta)
    if k not in _REGISTRY:
        raise FileNotFoundError(k)
    return _Reader(k, _REGISTRY[k], writable=(mode == "r+"))


class Window:
    """rasterio.windows.Window(col_off, row_off, width, height)"""

    def __init__(self, col_off=0, row_off=0, width=0, height=0):
        self.col_off, self.row_off, self.width, self.height = int(col_off), int(row_off), int(width), int(height)


def install() -> types.ModuleType:
    """Register the stub as `rasterio` in sys.modules (idempotent)."""
    mod = sys.modules.get("rasterio")
    if mod is not None and getattr(mod, "__dm_stub__", False):
        return mod
    mod = types.ModuleType("rasterio")
    mod.__dm_stub__ = True
    mod.open = _open
    mod.uint8 = "uint8"
    mod.uint16 = "uint16"
    mod.int16 = "int16"
    mod.DatasetReader = _Reader
    sys.modules["rasterio"] = mod
    # submodules the baseline builders and codec wrappers import at module top
    win = types.ModuleType("rasterio.windows")
    win.Window = Window
    enums = types.ModuleType("rasterio.enums")
    enums.Resampling = types.SimpleNamespace(nearest=0, bilinear=1, average=5)
    mod.windows, mod.enums = win, enums
    sys.modules["rasterio.windows"] = win
    sys.modules["rasterio.enums"] = enums
    if "matplotlib" not in sys.modules:           # make_baseline_B.py:34 imports pyplot for plots we never draw
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.pyplot = types.ModuleType("matplotlib.pyplot")
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = mpl.pyplot
    return mod
