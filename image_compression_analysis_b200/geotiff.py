"""Minimal GeoTIFF reader / writer with the slice of the rasterio dataset API the reference's metric
path touches (run_codec.py:242-259, 312-319; quicklooks.py:35-45, 122-205).  Used by raster_io when
rasterio is not importable.

Reader: classic TIFF and BigTIFF, little or big endian, strips or tiles, 8/16-bit unsigned or signed
samples, chunky (pixel-interleaved) or planar layout, compression NONE or DEFLATE (zlib), predictor 1
or 2.  That covers what the reference's tools write (tiled 512x512, BIGTIFF, uncompressed or DEFLATE:
make_baseline_A.py:76-78, make_baseline_B.py:258-293, quicklooks.py:152-163, the codec wrappers).
Anything else raises NotImplementedError naming the feature -- never a silent mis-read.

Beyond rasterio: `read_native()` returns a chunky file's samples as (H,W,B) without de-interleaving, so
a pixel-interleaved EnMAP cube goes to the GPU in the layout the one-pass BIP kernel wants.

Writer: single IFD, tiled, uint8/uint16/int16, DEFLATE or NONE, GDAL_NODATA / GDAL_METADATA tags, the
source's georeferencing tags passed through, `write_mask` as a GDAL-style `<name>.msk` sidecar.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# tags
_W, _H, _BPS, _COMP, _PHOTO, _STRIPOFF, _SPP, _RPS, _STRIPCNT = 256, 257, 258, 259, 262, 273, 277, 278, 279
_PLANAR, _PRED, _TW, _TL, _TOFF, _TCNT, _EXTRA, _SFMT = 284, 317, 322, 323, 324, 325, 338, 339
_SUBFILE = 254
_GDAL_META, _GDAL_NODATA = 42112, 42113
_GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737)   # pixel scale, tiepoints, transform, geo keys

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}


class _Ifd:
    def __init__(self):
        self.tags: Dict[int, Tuple[int, tuple]] = {}     # tag -> (type, values)

    def get(self, tag, default=None):
        v = self.tags.get(tag)
        return default if v is None else v[1]

    def one(self, tag, default=None):
        v = self.get(tag)
        return default if v is None or len(v) == 0 else v[0]


def _parse(buf: memoryview):
    bo = bytes(buf[:2])
    if bo == b"II":
        e = "<"
    elif bo == b"MM":
        e = ">"
    else:
        raise ValueError("not a TIFF file")
    magic = struct.unpack(e + "H", buf[2:4])[0]
    if magic == 42:
        big, first = False, struct.unpack(e + "I", buf[4:8])[0]
    elif magic == 43:
        big, first = True, struct.unpack(e + "Q", buf[8:16])[0]
    else:
        raise ValueError("not a TIFF file")
    ifds: List[_Ifd] = []
    off = first
    while off and len(ifds) < 16:
        ifd = _Ifd()
        if big:
            n = struct.unpack(e + "Q", buf[off:off + 8])[0]
            p, esz, cnt_fmt, inl = off + 8, 20, "Q", 8
        else:
            n = struct.unpack(e + "H", buf[off:off + 2])[0]
            p, esz, cnt_fmt, inl = off + 2, 12, "I", 4
        for i in range(n):
            ent = buf[p + i * esz:p + (i + 1) * esz]
            tag, typ = struct.unpack(e + "HH", ent[:4])
            cnt = struct.unpack(e + cnt_fmt, ent[4:4 + inl])[0]
            if typ not in _TYPE_SIZE:
                continue
            nbytes = cnt * _TYPE_SIZE[typ]
            if nbytes <= inl:
                data = ent[4 + inl:4 + inl + nbytes]
            else:
                o = struct.unpack(e + cnt_fmt, ent[4 + inl:4 + 2 * inl])[0]
                data = buf[o:o + nbytes]
            if typ == 2:
                vals = (bytes(data).split(b"\0")[0].decode("latin-1"),)
            elif typ in (5, 10):
                raw = struct.unpack(e + ("II" if typ == 5 else "ii") * cnt, data)
                vals = tuple(raw[2 * k] / raw[2 * k + 1] if raw[2 * k + 1] else 0.0 for k in range(cnt))
            elif typ in (3, 4, 16) and cnt > 64:
                vals = tuple(np.frombuffer(data, dtype=np.dtype({3: "u2", 4: "u4", 16: "u8"}[typ]).newbyteorder(e)).tolist())
            else:
                vals = struct.unpack(e + _TYPE_FMT[typ] * cnt, data)
            ifd.tags[tag] = (typ, vals)
        ifds.append(ifd)
        nxt = buf[p + n * esz:p + n * esz + (8 if big else 4)]
        off = struct.unpack(e + ("Q" if big else "I"), nxt)[0]
    return e, ifds


def _nodata_from(text: Optional[str]):
    if text is None:
        return None
    try:
        v = float(text.strip())
    except ValueError:
        return None
    return v


class Reader:
    """Read-only dataset: count/width/height/dtypes/nodata/descriptions/meta, read(), dataset_mask()."""

    def __init__(self, path):
        self.name = str(path)
        self._path = Path(path)
        self._mm = np.memmap(self._path, dtype=np.uint8, mode="r")
        self._buf = memoryview(self._mm)
        self._e, ifds = _parse(self._buf)
        main = [i for i in ifds if not (i.one(_SUBFILE, 0) & 4)]
        if not main:
            raise ValueError(f"{path}: no image directory")
        self._ifd = main[0]
        self._mask_ifd = next((i for i in ifds if i.one(_SUBFILE, 0) & 4), None)
        d = self._ifd
        self.width, self.height = int(d.one(_W)), int(d.one(_H))
        self.count = int(d.one(_SPP, 1))
        bps = d.get(_BPS, (1,))
        if len(set(bps)) != 1 or bps[0] not in (8, 16):
            raise NotImplementedError(f"{path}: BitsPerSample {bps} (8 and 16 are supported)")
        sf = d.get(_SFMT, (1,))
        if len(set(sf)) != 1 or sf[0] not in (1, 2):
            raise NotImplementedError(f"{path}: SampleFormat {sf} (integer samples only)")
        self._dtype = np.dtype({(8, 1): "u1", (8, 2): "i1", (16, 1): "u2", (16, 2): "i2"}[(bps[0], sf[0])])
        if self._dtype == np.dtype("i1"):
            raise NotImplementedError(f"{path}: int8 samples")
        self.dtypes = tuple([self._dtype.name] * self.count)
        self._planar = int(d.one(_PLANAR, 1))
        self._comp = int(d.one(_COMP, 1))
        if self._comp not in (1, 8, 32946):
            raise NotImplementedError(f"{path}: TIFF compression {self._comp} (NONE and DEFLATE are supported)")
        self._pred = int(d.one(_PRED, 1))
        if self._pred not in (1, 2):
            raise NotImplementedError(f"{path}: predictor {self._pred}")
        self.nodata = _nodata_from(d.one(_GDAL_NODATA))
        self.descriptions = tuple([None] * self.count)
        self.tiled = d.get(_TW) is not None
        if self.tiled:
            self._bw, self._bh = int(d.one(_TW)), int(d.one(_TL))
            self._offs, self._cnts = d.get(_TOFF), d.get(_TCNT)
        else:
            self._bw, self._bh = self.width, int(d.one(_RPS, self.height))
            self._bh = min(self._bh, self.height)
            self._offs, self._cnts = d.get(_STRIPOFF), d.get(_STRIPCNT)
        self._geotags = {t: d.tags[t] for t in _GEO_TAGS if t in d.tags}
        self.meta = {"driver": "GTiff", "dtype": self._dtype.name, "nodata": self.nodata, "width": self.width,
                     "height": self.height, "count": self.count, "crs": None, "transform": None,
                     "_geotags": self._geotags}
        self.interleave = "pixel" if (self._planar == 1 and self.count > 1) else "band"

    # context manager ---------------------------------------------------------------------------
    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self._buf = None
        self._mm = None

    # blocks ------------------------------------------------------------------------------------
    def _block(self, idx: int, rows: int, cols: int, spp: int) -> np.ndarray:
        """Decoded block `idx` as (rows, cols, spp) in native dtype."""
        o, n = int(self._offs[idx]), int(self._cnts[idx])
        raw = self._buf[o:o + n]
        want = rows * cols * spp * self._dtype.itemsize
        if self._comp != 1:
            raw = zlib.decompress(raw)
        if len(raw) < want:
            raise ValueError(f"{self.name}: block {idx} is short ({len(raw)} < {want} bytes)")
        a = np.frombuffer(raw, dtype=self._dtype.newbyteorder(self._e), count=rows * cols * spp).reshape(rows, cols, spp)
        if self._pred == 2 and self._comp != 1:
            a = np.cumsum(a.astype(self._dtype.newbyteorder("=")), axis=1, dtype=self._dtype)   # wraps modulo 2^bits
        return a

    def _grid(self):
        return (self.width + self._bw - 1) // self._bw, (self.height + self._bh - 1) // self._bh

    def native_shape(self):
        """(shape, layout) of what read_native() returns."""
        B, H, W = self.count, self.height, self.width
        if self._planar == 1 and B > 1:
            return (H, W, B), "bip"
        return (B, H, W), "bsq"

    def read_native(self, out: Optional[np.ndarray] = None, threads: int = 8):
        """All samples without de-interleaving: ((H,W,B) array, "bip") for chunky multi-band files,
        ((B,H,W) array, "bsq") otherwise.  `out`: a caller buffer of native_shape() and the file's
        dtype to decode into (e.g. pinned memory); blocks are decoded by a small thread pool
        (zlib and the numpy copies release the GIL)."""
        nx, ny = self._grid()
        B, H, W = self.count, self.height, self.width
        shape, layout = self.native_shape()
        if out is None:
            out = np.empty(shape, self._dtype)
        elif tuple(out.shape) != shape or out.dtype != self._dtype:
            raise ValueError(f"read_native: out must be {shape} {self._dtype}, not {out.shape} {out.dtype}")
        chunky = self._planar == 1
        view = out.reshape(H, W, B) if (chunky and B == 1) else out       # (1,H,W) == (H,W,1) in memory
        per_band = nx * ny

        def job(k: int) -> None:
            b, rest = (0, k) if chunky else divmod(k, per_band)
            by, bx = divmod(rest, nx)
            y0, x0 = by * self._bh, bx * self._bw
            rows_in_file = self._bh if self.tiled else min(self._bh, H - y0)
            h, w = min(self._bh, H - y0), min(self._bw, W - x0)
            if chunky:
                view[y0:y0 + h, x0:x0 + w, :] = self._block(k, rows_in_file, self._bw, B)[:h, :w, :]
            else:
                view[b, y0:y0 + h, x0:x0 + w] = self._block(k, rows_in_file, self._bw, 1)[:h, :w, 0]

        nblocks = per_band if chunky else per_band * B
        if threads > 1 and nblocks > 1:
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=min(threads, nblocks)) as ex:
                list(ex.map(job, range(nblocks)))
        else:
            for k in range(nblocks):
                job(k)
        return out, layout

    def read(self, indexes=None, out_dtype=None):
        """rasterio semantics: read() -> (B,H,W); read(i) -> (H,W); read([i,j]) -> (k,H,W); 1-based."""
        arr, layout = self.read_native()
        if layout == "bip":
            arr = np.ascontiguousarray(np.moveaxis(arr, -1, 0))
        if indexes is None:
            out = arr
        elif isinstance(indexes, (int, np.integer)):
            out = arr[int(indexes) - 1]
        else:
            out = arr[[int(i) - 1 for i in indexes]]
        return out.astype(out_dtype) if out_dtype is not None else out

    def has_explicit_mask(self) -> bool:
        """A per-dataset mask exists (.msk sidecar or internal mask directory): rasterio gives it priority over nodata."""
        return Path(str(self._path) + ".msk").exists() or self._mask_ifd is not None

    def dataset_mask(self):
        """uint8 (H,W), 0/255: .msk sidecar or internal mask if present, else any band != nodata,
        else all valid (rasterio's rule)."""
        msk = Path(str(self._path) + ".msk")
        if msk.exists():
            with Reader(msk) as m:
                return np.where(m.read(1) > 0, 255, 0).astype(np.uint8)
        if self._mask_ifd is not None:
            raise NotImplementedError(f"{self.name}: internal 1-bit mask directories are not supported")
        if self.nodata is not None and np.isfinite(self.nodata):
            a = self.read()
            return np.where((a != self.nodata).any(axis=0), 255, 0).astype(np.uint8)
        return np.full((self.height, self.width), 255, np.uint8)


# -------------------------------------------------------------------------------------------------
# writer
# -------------------------------------------------------------------------------------------------
class Writer:
    """Write-only dataset: write(arr (B,H,W) or (H,W)), write_mask(m), update_tags(**kw)."""

    def __init__(self, path, **meta):
        self._path = Path(path)
        self.count = int(meta.get("count", 1))
        self.width, self.height = int(meta["width"]), int(meta["height"])
        self._dtype = np.dtype(str(meta.get("dtype", "uint8")))
        if self._dtype not in (np.dtype("u1"), np.dtype("u2"), np.dtype("i2")):
            raise NotImplementedError(f"GeoTIFF writer: dtype {self._dtype}")
        comp = str(meta.get("compress") or "NONE").upper()
        if comp not in ("NONE", "DEFLATE"):
            raise NotImplementedError(f"GeoTIFF writer: compress={comp}")
        self._deflate = comp == "DEFLATE"
        self._bw = int(meta.get("blockxsize", 512)) if meta.get("tiled", True) else self.width
        self._bh = int(meta.get("blockysize", 512)) if meta.get("tiled", True) else self.height
        self._bw, self._bh = max(16, (self._bw + 15) // 16 * 16), max(16, (self._bh + 15) // 16 * 16)
        self.nodata = meta.get("nodata")
        self._force_big = str(meta.get("BIGTIFF", "")).upper() == "YES"
        self._geotags = dict(meta.get("_geotags") or {})
        self._tags: Dict[str, str] = {}
        self._data: Optional[np.ndarray] = None
        self._mask: Optional[np.ndarray] = None
        self.name = str(path)

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close()

    def write(self, arr, indexes=None):
        a = np.asarray(arr)
        if a.ndim == 2:
            a = a[None]
        if a.shape != (self.count, self.height, self.width):
            raise ValueError(f"write: array shape {a.shape} != {(self.count, self.height, self.width)}")
        self._data = np.ascontiguousarray(a.astype(self._dtype, copy=False))

    def write_mask(self, m):
        self._mask = np.asarray(m)

    def update_tags(self, **kw):
        self._tags.update({k: str(v) for k, v in kw.items()})

    def close(self):
        if self._data is None:
            raise ValueError("GeoTIFF writer: nothing written")
        _write_file(self._path, self._data, self._bw, self._bh, self._deflate, self.nodata, self._tags, self._geotags,
                    self._force_big)
        if self._mask is not None:
            m = np.where(np.asarray(self._mask) > 0, 255, 0).astype(np.uint8)[None]
            _write_file(Path(str(self._path) + ".msk"), m, self._bw, self._bh, True, None, {}, {})


def _write_file(path: Path, data: np.ndarray, bw: int, bh: int, deflate: bool, nodata, tags: Dict[str, str], geotags,
                force_big: bool = False):
    B, H, W = data.shape
    nx, ny = (W + bw - 1) // bw, (H + bh - 1) // bh
    dt = data.dtype
    blocks: List[bytes] = []
    # chunky (pixel-interleaved) tiles, like GDAL's default INTERLEAVE=PIXEL
    hwb = np.moveaxis(data, 0, -1)
    for by in range(ny):
        for bx in range(nx):
            tile = np.zeros((bh, bw, B), dt)
            y0, x0 = by * bh, bx * bw
            h, w = min(bh, H - y0), min(bw, W - x0)
            tile[:h, :w] = hwb[y0:y0 + h, x0:x0 + w]
            raw = tile.astype(dt.newbyteorder("<"), copy=False).tobytes()
            blocks.append(zlib.compress(raw, 6) if deflate else raw)
    payload = sum(len(b) for b in blocks)
    big = force_big or payload + 65536 > 0xFFFFFFF0
    entries: List[Tuple[int, int, int, bytes]] = []      # tag, type, count, data bytes

    def add(tag, typ, vals):
        if typ == 2:
            bts = vals.encode("latin-1") + b"\0"
            entries.append((tag, 2, len(bts), bts))
        else:
            vals = list(vals)
            entries.append((tag, typ, len(vals), struct.pack("<" + _TYPE_FMT[typ] * len(vals), *vals)))

    otype = 16 if big else 4
    add(_W, 4, [W]); add(_H, 4, [H]); add(_BPS, 3, [dt.itemsize * 8] * B)
    add(_COMP, 3, [8 if deflate else 1]); add(_PHOTO, 3, [1]); add(_SPP, 3, [B]); add(_PLANAR, 3, [1])
    add(_TW, 3, [bw]); add(_TL, 3, [bh])
    add(_SFMT, 3, [2 if dt.kind == "i" else 1] * B)
    if B > 1:
        add(_EXTRA, 3, [0] * (B - 1))
    for t, (typ, vals) in sorted(geotags.items()):
        if typ == 2:
            add(t, 2, vals[0])
        elif typ in (5, 10):
            continue
        else:
            add(t, typ, vals)
    if tags:
        items = "".join(f'  <Item name="{k}">{_xml(v)}</Item>\n' for k, v in tags.items())
        add(_GDAL_META, 2, f"<GDALMetadata>\n{items}</GDALMetadata>\n")
    if nodata is not None:
        add(_GDAL_NODATA, 2, repr(float(nodata)) if float(nodata) != int(float(nodata)) else str(int(float(nodata))))
    # block offsets / counts are filled below
    head = 16 if big else 8
    n_ent = len(entries) + 2
    esz, inl = (20, 8) if big else (12, 4)
    ifd_size = (8 if big else 2) + n_ent * esz + (8 if big else 4)
    ext_pos = head + ifd_size
    counts = [len(b) for b in blocks]
    ext = bytearray()
    offs_pos = ext_pos + len(ext); ext += b"\0" * (len(blocks) * _TYPE_SIZE[otype])
    cnts_pos = ext_pos + len(ext); ext += struct.pack("<" + _TYPE_FMT[otype] * len(blocks), *counts)
    packed = []
    for tag, typ, cnt, bts in entries:
        if len(bts) <= inl:
            packed.append((tag, typ, cnt, bts.ljust(inl, b"\0"), None))
        else:
            if len(ext) % 2:
                ext += b"\0"
            packed.append((tag, typ, cnt, None, ext_pos + len(ext)))
            ext += bts
    if len(ext) % 16:
        ext += b"\0" * (16 - len(ext) % 16)
    data_pos = ext_pos + len(ext)
    offsets, p = [], data_pos
    for b in blocks:
        offsets.append(p); p += len(b)
    ext[offs_pos - ext_pos:offs_pos - ext_pos + len(blocks) * _TYPE_SIZE[otype]] = struct.pack("<" + _TYPE_FMT[otype] * len(blocks), *offsets)
    if len(blocks) * _TYPE_SIZE[otype] <= inl:
        packed.append((_TOFF, otype, len(blocks), struct.pack("<" + _TYPE_FMT[otype] * len(blocks), *offsets).ljust(inl, b"\0"), None))
        packed.append((_TCNT, otype, len(blocks), struct.pack("<" + _TYPE_FMT[otype] * len(blocks), *counts).ljust(inl, b"\0"), None))
    else:
        packed.append((_TOFF, otype, len(blocks), None, offs_pos))
        packed.append((_TCNT, otype, len(blocks), None, cnts_pos))
    packed.sort(key=lambda t: t[0])
    out = bytearray()
    if big:
        out += b"II" + struct.pack("<HHHQ", 43, 8, 0, head)
        out += struct.pack("<Q", n_ent)
    else:
        out += b"II" + struct.pack("<HI", 42, head)
        out += struct.pack("<H", n_ent)
    for tag, typ, cnt, inline, off in packed:
        out += struct.pack("<HH", tag, typ) + struct.pack("<Q" if big else "<I", cnt)
        out += inline if inline is not None else struct.pack("<Q" if big else "<I", off)
    out += struct.pack("<Q" if big else "<I", 0)
    assert len(out) == ext_pos, (len(out), ext_pos)
    path.parent.mkdir(parents=True, exist_ok=True)
    with path.open("wb") as f:
        f.write(out); f.write(ext)
        for b in blocks:
            f.write(b)


def _xml(s: str) -> str:
    return s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;")


def read_tags(path) -> Dict[str, str]:
    """The GDAL_METADATA items of a file written by `Writer` (or GDAL)."""
    import re
    with Reader(path) as r:
        txt = r._ifd.one(_GDAL_META)
    if not txt:
        return {}
    return {k: v.replace("&lt;", "<").replace("&gt;", ">").replace("&amp;", "&")
            for k, v in re.findall(r'<Item name="([^"]+)"[^>]*>([^<]*)</Item>', txt)}


def open(path, mode: str = "r", **meta):      # noqa: A001  (mirrors rasterio.open)
    if mode == "r":
        return Reader(path)
    if mode == "w":
        return Writer(path, **meta)
    raise NotImplementedError(f"geotiff.open mode {mode!r} (the metric path only reads and writes whole files)")
