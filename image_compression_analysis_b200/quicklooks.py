"""Drop-in replacement for the reference's tools/quicklooks.py (run_codec.py loads it through
`--quicklooks <path>`, run_codec.py:419-430): same function names, arguments, file names, tags.

  write_error_max8               quicklooks.py:115-207   B200 kernels (dm_spectral)
  stretch_params_from_baseline   quicklooks.py:51-72     dm_band_hist (exact value histograms) -> percentiles on the host
  write_rgb_8bit                 quicklooks.py:78-109    dm_lut_bands_u8 (stretch tabulated with the reference's expression)

The error map is max over bands of |A-B| per pixel, zero where either input is invalid, scaled to
8 bit with the reference's float32 expression.  That expression is elementwise on an integer-valued
plane, so the kernel indexes a table built on the host with the reference's own formula
(finish.err8_lut); planes are bit-exact, STATISTICS_MEAN is bit-exact, STATISTICS_STDDEV agrees to
~1e-16 relative (it is derived from the plane's 256-bin histogram instead of a two-pass float sum).
"""
from __future__ import annotations

import argparse
from pathlib import Path
from typing import Dict

import numpy as np

from . import finish
from ._lib import DM_VALID_QUICKLOOK
from .engine import DevicePair, Want, evaluate, to_device
from .raster_io import open_raster, uint8_dtype

RGB_ORDER = [3, 2, 1]  # 1-based band indices, as in the reference


def _valid_mask_from_ds(ds):
    """quicklooks.py:35-45 (host form, used by the RGB helpers)."""
    m = ds.dataset_mask() > 0
    nd = ds.nodata
    if nd is not None and np.isfinite(nd):
        try:
            m &= (ds.read(1) != nd)
        except Exception:
            pass
    return m


def _quicklook_plane(cube_t, np_dtype, layout, B, H, W, nodata):
    """Device plane with DM_VALID_QUICKLOOK = dataset mask AND band 1 != nodata (quicklooks.py:35-45), or None
    when the file has no nodata value (everything valid)."""
    import ctypes as C
    import torch
    from . import engine
    from ._lib import check, lib
    nd = engine.integral_nodata(nodata, np_dtype)
    if nd is None:
        return None
    # dm_validity works on pairs: the cube stands on both sides, which leaves every bit unchanged
    pair = DevicePair(cube_t, cube_t, np_dtype, layout, B, H, W, nd, nd)
    plane = torch.empty(H * W, dtype=torch.uint8, device=cube_t.device)
    cnt = torch.zeros(3, dtype=torch.int64, device=cube_t.device)
    cp = pair.c_pair()
    check(lib().dm_validity(C.byref(cp), None, engine._ptr(plane), engine._ptr(cnt), engine._stream_ptr()))
    return plane


def stretch_params_cube(cube_t, np_dtype, layout, B, H, W, nodata=None, extra_mask=None, rgb_order=RGB_ORDER, pct=(2, 98)):
    """stretch_params_from_baseline for a device-resident cube: exact value histograms of the three bands
    (dm_band_hist) -> numpy's interpolated percentiles on the host (finish.percentiles_from_hist)."""
    import torch
    from . import adjacent
    plane = _quicklook_plane(cube_t, np_dtype, layout, B, H, W, nodata)
    bit = DM_VALID_QUICKLOOK
    if extra_mask is not None:              # explicit alpha / .msk mask: fold it into the plane
        m = to_device(np.asarray(extra_mask) > 0).reshape(-1)
        plane = (m * bit) if plane is None else (plane & (m * bit))
        plane = plane.to(torch.uint8)
    sel = [int(i) - 1 for i in rgb_order]
    hist = adjacent.band_hist(cube_t, np_dtype, layout, B, H, W, sel, plane, bit).cpu().numpy()
    params = []
    for h in hist:
        got = finish.percentiles_from_hist(h, adjacent.first_value(np_dtype), pct)
        if got is None:
            lo, hi = 0.0, 1.0
        else:
            lo, hi = got
            if not np.isfinite(lo):
                lo = 0.0
            if (not np.isfinite(hi)) or hi <= lo:
                hi = lo + 1.0
        params.append((float(lo), float(hi)))
    return params


def rgb_8bit_cube(cube_t, np_dtype, layout, B, H, W, params, rgb_order=RGB_ORDER):
    """write_rgb_8bit's pixels for a device-resident cube: (3,H,W) uint8 device tensor (dm_lut_bands_u8)."""
    from . import adjacent
    luts = np.stack([finish.stretch8_lut(lo, hi, np_dtype) for lo, hi in params], 0)
    return adjacent.lut_bands_u8(cube_t, np_dtype, layout, B, H, W, [int(i) - 1 for i in rgb_order], luts)


def stretch_params_arrays(cube, nodata=None, rgb_order=RGB_ORDER, pct=(2, 98), layout: str = "bsq"):
    """stretch_params_from_baseline on an in-memory cube ((B,H,W) for "bsq", (H,W,B) for "bip")."""
    cube = np.asarray(cube)
    B, H, W = cube.shape if layout == "bsq" else (cube.shape[2], cube.shape[0], cube.shape[1])
    return stretch_params_cube(to_device(cube), cube.dtype.name, layout, B, H, W, nodata, None, rgb_order, pct)


def rgb_8bit_arrays(cube, params, rgb_order=RGB_ORDER, layout: str = "bsq") -> np.ndarray:
    cube = np.asarray(cube)
    B, H, W = cube.shape if layout == "bsq" else (cube.shape[2], cube.shape[0], cube.shape[1])
    return rgb_8bit_cube(to_device(cube), cube.dtype.name, layout, B, H, W, params, rgb_order).cpu().numpy()


def stretch_params_from_baseline(path, rgb_order=RGB_ORDER, pct=(2, 98)):
    """Per-channel (lo, hi) percentile stretch ignoring invalid pixels (quicklooks.py:51-72)."""
    from .ingest import load_cube
    c = load_cube(path)
    return stretch_params_cube(c.tensor, c.np_dtype, c.layout, c.bands, c.rows, c.width, c.nodata, c.mask, rgb_order, pct)


def write_rgb_8bit(src_path, out_path, params, rgb_order=RGB_ORDER):
    """8-bit RGB quicklook with the source's valid mask, no nodata carried over (quicklooks.py:78-109)."""
    from .ingest import load_cube
    c = load_cube(src_path)
    assert c.bands >= 3, f"Need ≥3 bands for RGB in {src_path}"
    rgb = rgb_8bit_cube(c.tensor, c.np_dtype, c.layout, c.bands, c.rows, c.width, params, rgb_order).cpu().numpy()
    meta = dict(c.meta)
    meta.update(driver="GTiff", dtype=uint8_dtype(), count=3, photometric="RGB", tiled=True,
                blockxsize=512, blockysize=512, compress="DEFLATE")
    meta.pop("nodata", None)
    out_path = Path(out_path)
    out_path.parent.mkdir(parents=True, exist_ok=True)
    with open_raster(src_path) as ds:
        mask = ds.dataset_mask()
    with open_raster(out_path.as_posix(), "w", **meta) as dst:
        dst.write(rgb)
        try:
            dst.write_mask(mask)
        except Exception:
            pass


def error_max8_arrays(A, B, err_max_global=255, err_max_zoom=None, pct=(2, 98), *, a_nodata=None, b_nodata=None,
                      layout: str = "bsq", a_mask=None, b_mask=None) -> Dict[str, object]:
    """Pixel content of write_error_max8 for in-memory cubes.

    Returns err8_g / err8_z (uint8 (H,W) numpy; err8_z None without a zoom cap), valid (bool (H,W)),
    cap_g / cap_z (the integers in the file names) and mean_* / std_* (the statistics tags).
    """
    assert tuple(A.shape) == tuple(B.shape), "Dims/band count must match"
    pair = DevicePair.from_arrays(A, B, layout, a_nodata, b_nodata)
    return error_max8_pair(pair, err_max_global, err_max_zoom, pct, a_mask=a_mask, b_mask=b_mask)


def error_max8_pair(pair: DevicePair, err_max_global=255, err_max_zoom=None, pct=(2, 98), *, a_mask=None,
                    b_mask=None) -> Dict[str, object]:
    """error_max8_arrays on a device-resident pair."""
    H, W = pair.rows, pair.width
    extra = None
    for m in (a_mask, b_mask):
        if m is not None:
            extra = np.asarray(m) > 0 if extra is None else (extra & (np.asarray(m) > 0))

    def lut_for(cap, errmax_dev):
        if cap is not None:
            return finish.err8_lut(cap), int(round(float(cap)))
        # percentile branch (quicklooks.py:137-146): unreachable from run_codec, kept for the CLI.  The 2nd /
        # 98th percentile of the non-zero errors comes from the exact value histogram of the uint16 error
        # plane (dm_band_hist), like the RGB stretch; bin 0 is dropped (nz = e[e > 0]).
        from . import adjacent
        e = errmax_dev()
        hist = adjacent.band_hist(e, "uint16", "bsq", 1, H, W, [0]).cpu().numpy()[0]
        got = finish.percentiles_from_hist(hist[1:], 1, pct)
        if got is None:
            lo, hi = 0.0, 1.0
        else:
            # np.percentile hands the reference numpy float64 scalars, which make its float32 expression
            # evaluate in float64 (NumPy 2 promotion); Python floats would keep it in float32
            lo, hi = np.float64(got[0]), np.float64(got[1])
            if not np.isfinite(lo):
                lo = 0.0
            if (not np.isfinite(hi)) or hi <= lo:
                hi = lo + 1.0
        nzb = np.nonzero(hist)[0]
        top = int(nzb[-1]) if nzb.size else 0
        grid = np.arange(top + 1, dtype=np.int64).astype(np.float32)
        lut = (np.clip((grid - lo) / (hi - lo + 1e-9), 0, 1) * 255.0).astype(np.uint8)
        return lut, int(round(hi))

    _cache = {}

    def errmax_host():
        """uint16 plane of max_b |A - B| on the device (0 where invalid), explicit masks folded in."""
        if "e" not in _cache:
            P0 = evaluate(pair, Want(stats=False, errmax=True))
            e = P0.planes["errmax"]
            if extra is not None:
                e = e * to_device(extra.reshape(-1)).to(e.dtype)
            _cache["e"] = e
        return _cache["e"]

    from . import engine
    lut_g, cap_g = lut_for(err_max_global, errmax_host)
    caps = [(lut_g, cap_g)]
    if err_max_zoom is not None:
        caps.append(lut_for(err_max_zoom, errmax_host))
    # one spectral pass writes both planes; LUTs go in explicitly (they may come from percentiles)
    import ctypes as C
    import torch
    from ._lib import check, lib
    dev = pair.ref.device
    P = engine.Partials.allocate(pair.bands, 0, dev, pair.np_dtype)
    plane = None
    cp = pair.c_pair()
    st = engine._stream_ptr()
    if engine.needs_plane(pair, None):
        plane = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
        check(lib().dm_validity(C.byref(cp), None, engine._ptr(plane), engine._ptr(P.counts), st))
    lg = to_device(caps[0][0])
    og = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
    lz = oz = None
    if len(caps) > 1:
        lz = to_device(caps[1][0])
        oz = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
    from . import _lib
    plane_args = (engine._ptr(lg), lg.numel() - 1, engine._ptr(og), engine._ptr(P.hist8_g),
                  engine._ptr(lz), 0 if lz is None else lz.numel() - 1, engine._ptr(oz), engine._ptr(P.hist8_z))
    # the one-pass kernels write the planes at HBM speed where they apply (16-bit BIP cubes with a multiple of four
    # bands -- EnMAP through ingest's transposition --, BSQ cubes of up to four bands -- Case A); their per-band
    # statistics land in P and are not used here.  Everything else takes the per-pixel spectral pass.
    rc = _lib.DM_EUNSUPPORTED
    if pair.layout == "bip":
        rc = lib().dm_fused_bip(C.byref(cp), engine._ptr(plane), engine._ptr(P.sums), engine._ptr(P.imax), None, *plane_args,
                                0, None, None, st)
    elif pair.bands <= 4:
        rc = lib().dm_fused_bsq(C.byref(cp), engine._ptr(plane), engine._ptr(P.sums), engine._ptr(P.imax), None, *plane_args, st)
    if rc == _lib.DM_EUNSUPPORTED:
        rc = lib().dm_spectral(C.byref(cp), engine._ptr(plane), None, *plane_args, 0, 0, None, None, st)
    check(rc)
    h = P.to_host()
    err8_g = og.cpu().numpy().reshape(H, W)
    err8_z = None if oz is None else oz.cpu().numpy().reshape(H, W)
    if plane is not None:
        valid = (plane.cpu().numpy().reshape(H, W) & engine.DM_VALID_QUICKLOOK) != 0
    else:
        valid = np.ones((H, W), bool)
    hg, hz = h.hist8_g.copy(), h.hist8_z.copy()
    if extra is not None:      # explicit alpha/.msk masks (rare): zero the invalid pixels on the host
        for arr, hh, lut in ((err8_g, hg, caps[0][0]), (err8_z, hz, caps[-1][0])):
            if arr is None:
                continue
            bad = valid & ~extra
            np.subtract.at(hh, arr[bad], 1)
            arr[bad] = lut[0]
            hh[int(lut[0])] += int(bad.sum())
        valid &= extra
    out: Dict[str, object] = {"err8_g": err8_g, "err8_z": err8_z, "valid": valid, "cap_g": caps[0][1],
                              "cap_z": caps[1][1] if len(caps) > 1 else None}
    sg = finish.err8_stats_from_hist(hg)
    out.update(mean_g=sg["mean"], std_g=sg["std"])
    if err8_z is not None:
        sz = finish.err8_stats_from_hist(hz)
        out.update(mean_z=sz["mean"], std_z=sz["std"])
    return out


def write_error_max8(a_path, b_path, out_path_base, err_max_global=255, err_max_zoom=None, pct=(2, 98)):
    """
    Generate 8-bit error map(s) from per-pixel max abs diff across bands (quicklooks.py:115-207):
      - <base>_ERR8_0_<err_max_global>.tif
      - <base>_ERR8_0_<err_max_zoom>.tif (optional)
    Returns: (global_path, zoom_path or None)
    """
    from .ingest import load_pair
    try:
        pair, info = load_pair(a_path, b_path)      # read once per rep, shared with compute_metrics
    except AssertionError:
        raise AssertionError("Dims/band count must match")
    res = error_max8_pair(pair, err_max_global, err_max_zoom, pct, a_mask=info["ref_mask"], b_mask=info["tst_mask"])
    meta = dict(info["ref_meta"])
    meta.update(driver="GTiff", count=1, dtype=uint8_dtype(), photometric="MINISBLACK", tiled=True,
                blockxsize=512, blockysize=512, compress="DEFLATE")
    meta.pop("nodata", None)
    out_base = Path(out_path_base)
    out_base.parent.mkdir(parents=True, exist_ok=True)
    mask255 = res["valid"].astype(np.uint8) * 255

    def emit(plane, cap, mean, std):
        out = out_base.with_name(out_base.stem + f"_ERR8_0_{cap}.tif")
        with open_raster(out.as_posix(), "w", **meta) as dst:
            dst.write(plane[None, ...])
            try:
                dst.write_mask(mask255)
            except Exception:
                pass
            dst.update_tags(STATISTICS_MINIMUM="0", STATISTICS_MAXIMUM="255", STATISTICS_MEAN=str(float(mean)),
                            STATISTICS_STDDEV=str(float(std)), PIXEL_MINIMUM="0", PIXEL_MAXIMUM="255")
        return out

    out_g = emit(res["err8_g"], res["cap_g"], res["mean_g"], res["std_g"])
    out_z = None
    if err_max_zoom is not None:
        out_z = emit(res["err8_z"], res["cap_z"], res["mean_z"], res["std_z"])
    return out_g, out_z


def main(argv=None) -> int:
    """The reference module's command line (quicklooks.py:213-239), used by make_baseline_A.py:189-198."""
    ap = argparse.ArgumentParser(description="RGB quicklook and 8-bit error maps (B200)")
    ap.add_argument("--baseline", required=True, help="Reference multiband image")
    ap.add_argument("--out", help="Output 8-bit RGB from baseline (optional)")
    ap.add_argument("--error-against", help="Image to compare with baseline (same shape)")
    ap.add_argument("--err-out-base", help="Output prefix for error maps (no suffix)")
    ap.add_argument("--err-max-global", type=int, default=255)
    ap.add_argument("--err-max-zoom", type=int, default=None)
    ap.add_argument("--rgb-order", nargs=3, type=int, default=RGB_ORDER)
    ap.add_argument("--rgb-pct", nargs=2, type=float, default=(2, 98))
    args = ap.parse_args(argv)
    p = Path(args.baseline)
    if args.out:
        prm = stretch_params_from_baseline(p, rgb_order=args.rgb_order, pct=tuple(args.rgb_pct))
        write_rgb_8bit(p, Path(args.out), prm, rgb_order=args.rgb_order)
    if args.error_against:
        base = Path(args.err_out_base) if args.err_out_base else Path(args.baseline).with_suffix("")
        write_error_max8(a_path=args.baseline, b_path=args.error_against, out_path_base=base.as_posix(),
                         err_max_global=args.err_max_global, err_max_zoom=args.err_max_zoom)
    return 0


if __name__ == "__main__":          # python -m image_compression_analysis_b200.quicklooks ...
    raise SystemExit(main())
