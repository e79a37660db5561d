"""Drop-in mirror of the reference's distortion-metric functions, running on B200 kernels.

Same names, argument meaning, return keys/types and error behaviour as
/root/reference/tools/run_codec.py:

    mse(a, b)                                   :55-57
    psnr(a, b, data_range)                      :60-64
    ssim_global(a, b, data_range)               :67-80
    effective_data_range(ds)                    :86-117
    sobel_mag(img)                              :123-137   (float64 magnitude map)
    compute_metrics(ref_path, tst_path, valid)  :240-304
    compute_sam_sid_lmse_caseB(ref_path, ...)   :308-347

plus array-level twins (`*_arrays`) that take (B,H,W) / (H,W,B) arrays instead of GeoTIFF paths,
and `all_metrics_arrays`, which uploads a pair once and evaluates everything.  Integer quantities
are bit-exact with the reference; PSNR is bit-exact (host math.log10 on exact integers); SSIM, SAM,
SID, LMSE agree to ~1e-15 relative (gate: 1e-6).  There is no CPU path.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch

from . import finish
from .engine import DevicePair, Partials, Want, dtype_code, evaluate, to_device

# ---------------------------------------------------------------------------------------------
# scalar metrics (run_codec.py:55-80)
# ---------------------------------------------------------------------------------------------


def _flat_pair(a, b) -> DevicePair:
    a = np.asarray(a)
    b = np.asarray(b)
    if a.shape != b.shape:
        raise ValueError(f"operands could not be broadcast together with shapes {a.shape} {b.shape}")
    if a.dtype != b.dtype:
        raise TypeError("mse/psnr/ssim_global need two arrays of the same integer sample type")
    dtype_code(a.dtype)
    n = int(a.size)
    return DevicePair.from_arrays(a.reshape(1, 1, n), b.reshape(1, 1, n), "bsq")


def _scalar_partials(a, b):
    pair = _flat_pair(a, b)
    return evaluate(pair, Want(stats=True)).to_host()


def mse(a, b) -> float:
    """run_codec.py:55-57.  float(sum d^2)/N on exact integers == np.mean(d*d) in float64."""
    h = _scalar_partials(a, b)
    n = int(h.sums[0, 0])
    if n == 0:
        return float("nan")
    return float(int(h.sums[0, 7])) / n


def psnr(a, b, data_range: float) -> float:
    """run_codec.py:60-64."""
    h = _scalar_partials(a, b)
    return finish.psnr_from_sse(int(h.sums[0, 7]), int(h.sums[0, 0]), data_range)


def ssim_global(a, b, data_range: float) -> float:
    """run_codec.py:67-80 (window-less SSIM from global moments)."""
    h = _scalar_partials(a, b)
    s = h.sums[0]
    return finish.ssim_global_from_moments(int(s[0]), int(s[1]), int(s[2]), int(s[3]), int(s[4]), int(s[5]), data_range)


def effective_data_range_arrays(ref, layout: str = "bsq") -> int:
    """run_codec.py:86-117 on an array: one unmasked range scan of the reference cube."""
    ref = np.asarray(ref)
    if ref.dtype == np.uint8:
        return 255
    dtype_code(ref.dtype)
    t = to_device(ref)
    shp = ref.shape
    B, H, W = (shp if layout == "bsq" else (shp[2], shp[0], shp[1]))
    pair = DevicePair(t, t, ref.dtype.name, layout, B, H, W)
    h = evaluate(pair, Want(stats=True, moments=False)).to_host()
    return finish.data_range_from_maxs(dtype_code(ref.dtype), h.maxs)


def effective_data_range(ds) -> int:
    """run_codec.py:86-117 on a rasterio-like dataset."""
    dtype = ds.dtypes[0]
    if dtype == "uint8":
        return 255
    if dtype in ("uint16", "int16"):
        return effective_data_range_arrays(ds.read())
    try:
        return int(np.iinfo(np.dtype(dtype)).max)
    except Exception:
        return 65535


# ---------------------------------------------------------------------------------------------
# compute_metrics (run_codec.py:240-304)
# ---------------------------------------------------------------------------------------------


def _valid_to_device(valid, H: int, W: int, what: str):
    if valid is None:
        return None
    if tuple(valid.shape) != (H, W):
        raise ValueError(what)
    return to_device(np.ascontiguousarray(np.asarray(valid).astype(bool)).reshape(-1))


def metrics_partials(pair: DevicePair, valid_dev, want: Want) -> Partials:
    """Fused stats with the reference's mask rule: an all-False mask means "use every pixel"
    (run_codec.py:264), which is only known after the validity counts come back."""
    P = evaluate(pair, want, valid_dev)
    if P.used_mask:
        n_valid = int(P.counts[0].item())
        if n_valid == 0:
            P.sums.zero_()
            P.imax.zero_()
            if want.hist_bins:
                P.hist.zero_()
            w2 = Want(stats=True, moments=want.moments, hist_bins=want.hist_bins, generic_stats=want.generic_stats)
            evaluate(pair, w2, valid_dev, out=P, metrics_mask=False, plane=P.planes.get("valid"))
            P.used_mask = False
    return P


def compute_metrics_arrays(ref, tst, valid=None, *, ref_nodata=None, tst_nodata=None, layout: str = "bsq",
                           hist_bins: int = 0, extras: bool = False) -> Dict[str, float]:
    """compute_metrics on in-memory cubes.  Returns the reference's keys; with extras=True also the
    MAE / SSE / histogram additions (new keys only)."""
    pair = DevicePair.from_arrays(ref, tst, layout, ref_nodata, tst_nodata)
    return compute_metrics_pair(pair, valid, hist_bins=hist_bins, extras=extras)


def compute_metrics_pair(pair: DevicePair, valid=None, *, hist_bins: int = 0, extras: bool = False) -> Dict[str, float]:
    """compute_metrics on a device-resident pair."""
    vdev = _valid_to_device(valid, pair.rows, pair.width, f"Mask shape {None if valid is None else valid.shape} != {(pair.rows, pair.width)}")
    P = metrics_partials(pair, vdev, Want(stats=True, hist_bins=hist_bins))
    h = P.to_host()
    return finish.finish_compute_metrics(dtype_code(pair.np_dtype), h.sums, h.maxs, h.hist, extras=extras)


def _fold_masks(valid, info):
    """AND the explicit dataset masks (alpha / .msk, rare) into the caller's mask."""
    m = valid
    for k in ("ref_mask", "tst_mask"):
        if info[k] is not None:
            m = info[k] if m is None else (np.asarray(m).astype(bool) & info[k])
    return m


def compute_metrics(ref_path: Path, tst_path: Path, valid: Optional[np.ndarray] = None) -> Dict[str, float]:
    """Compute PSNR/SSIM per band + global PSNR/SSIM and per-band max|d| (run_codec.py:240-304)."""
    from .ingest import load_pair
    pair, info = load_pair(ref_path, tst_path)          # read once per rep, shared with the other two calls
    if valid is not None and tuple(valid.shape) != (info["H"], info["W"]):
        raise ValueError(f"Mask shape {valid.shape} != {(info['H'], info['W'])}")
    return compute_metrics_pair(pair, _fold_masks(valid, info))


# ---------------------------------------------------------------------------------------------
# Case-B spectral metrics (run_codec.py:308-347)
# ---------------------------------------------------------------------------------------------


def compute_sam_sid_lmse_caseB_arrays(ref, tst, valid=None, *, ref_nodata=None, tst_nodata=None,
                                      layout: str = "bsq") -> Dict[str, float]:
    pair = DevicePair.from_arrays(ref, tst, layout, ref_nodata, tst_nodata)
    return compute_sam_sid_lmse_caseB_pair(pair, valid)


def compute_sam_sid_lmse_caseB_pair(pair: DevicePair, valid=None) -> Dict[str, float]:
    """SAM / SID / LMSE on a device-resident pair."""
    vdev = _valid_to_device(valid, pair.rows, pair.width, "Mask shape mismatch for Case B metrics")
    P = evaluate(pair, Want(stats=False, sam=True, sid=True, lmse=True), vdev)
    h = P.to_host()
    return finish.finish_spectral(float(h.spec[0]), float(h.spec[1]), float(h.spec[2]), h.lmse, pair.npix)


def compute_sam_sid_lmse_caseB(ref_path: Path, tst_path: Path, valid: Optional[np.ndarray] = None) -> Dict[str, float]:
    """Compute SAM (deg), SID, and LMSE for Case B (run_codec.py:308-347)."""
    from .ingest import load_pair
    pair, info = load_pair(ref_path, tst_path)
    if valid is not None:
        if tuple(valid.shape) != (info["H"], info["W"]):
            raise ValueError("Mask shape mismatch for Case B metrics")
        m = valid                      # the dataset masks are NOT applied when `valid` is given (:314-319)
    else:
        m = _fold_masks(None, info)
    return compute_sam_sid_lmse_caseB_pair(pair, m)


def sobel_mag(img: np.ndarray) -> np.ndarray:
    """3x3 Sobel gradient magnitude of one (H,W) band as float64 (run_codec.py:123-137), on the GPU.
    Integer samples of up to 16 bits (what the path sees); other sample types are an error, not a CPU detour."""
    from ._lib import check, lib
    from .engine import _ptr, _stream_ptr
    img = np.asarray(img)
    if img.ndim != 2:
        raise ValueError("sobel_mag takes one (H,W) band")
    code = dtype_code(img.dtype)
    H, W = img.shape
    dev_img = to_device(img)
    out = torch.empty((H, W), dtype=torch.float64, device=dev_img.device)
    check(lib().dm_sobel_mag(_ptr(dev_img), code, H, W, _ptr(out), _stream_ptr()))
    return out.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# additions: Gaussian-window SSIM, everything-at-once
# ---------------------------------------------------------------------------------------------


def ssim_gaussian_arrays(ref, tst, data_range: Optional[float] = None, layout: str = "bsq") -> Dict[str, float]:
    """Per-band Gaussian-window SSIM (sigma 1.5, 11 taps, skimage semantics) under NEW keys
    ssimw_b{i} / ssimw_band_avg; `ssim_b{i}` keeps the reference's window-less definition."""
    pair = DevicePair.from_arrays(ref, tst, layout)
    if data_range is None:
        h0 = evaluate(pair, Want(stats=True, moments=False)).to_host()
        data_range = finish.data_range_from_maxs(dtype_code(pair.np_dtype), h0.maxs)
    h = evaluate(pair, Want(stats=False, ssim_gauss=True), data_range=data_range).to_host()
    return finish.finish_ssim_gauss(h.ssimw_sum, h.ssimw_cnt)


def all_metrics_arrays(ref, tst, valid=None, *, ref_nodata=None, tst_nodata=None, layout: str = "bsq",
                       case_b: bool = False, hist_bins: int = 0, ssim_window: bool = False,
                       extras: bool = True) -> Dict[str, object]:
    """Upload the pair once and evaluate compute_metrics (+ Case-B metrics, + Gaussian SSIM)."""
    pair = DevicePair.from_arrays(ref, tst, layout, ref_nodata, tst_nodata)
    vdev = _valid_to_device(valid, pair.rows, pair.width, f"Mask shape != {(pair.rows, pair.width)}")
    P = metrics_partials(pair, vdev, Want(stats=True, hist_bins=hist_bins))
    if case_b:
        evaluate(pair, Want(stats=False, sam=True, sid=True, lmse=True), vdev, out=P, plane=P.planes.get("valid"))
    h = P.to_host()
    code = dtype_code(pair.np_dtype)
    out = finish.finish_compute_metrics(code, h.sums, h.maxs, h.hist, extras=extras)
    if case_b:
        out.update(finish.finish_spectral(float(h.spec[0]), float(h.spec[1]), float(h.spec[2]), h.lmse, pair.npix))
    if ssim_window:
        L = finish.data_range_from_maxs(code, h.maxs)
        h2 = evaluate(pair, Want(stats=False, ssim_gauss=True), data_range=L).to_host()
        out.update(finish.finish_ssim_gauss(h2.ssimw_sum, h2.ssimw_cnt))
    return out
