"""CPU oracle for the reconstruction-distortion path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement, at ARRAY level (no rasterio, no files), of what the
reference computes after every codec decode:

  * compute_metrics            /root/reference/tools/run_codec.py:240-304
  * compute_sam_sid_lmse_caseB /root/reference/tools/run_codec.py:308-347
  * mse / psnr / ssim_global   /root/reference/tools/run_codec.py:55-80
  * effective_data_range       /root/reference/tools/run_codec.py:86-117
  * sobel_mag                  /root/reference/tools/run_codec.py:123-137
  * write_error_max8 (pixels)  /root/reference/tools/quicklooks.py:115-207
  * _valid_mask_from_ds        /root/reference/tools/quicklooks.py:35-45

plus the three additions BASELINE.json's north_star names that the reference
does not contain: MAE, per-band |d| histograms
and a Gaussian-window SSIM.

Pinning status
--------------
* In-reference metrics: PINNED.  tests/test_oracle_vs_reference.py runs the
  real reference (under oracle/rasterio_stub.py) beside this file whenever
  /root/reference is mounted, and tests/golden/*.json holds outputs generated
  by the real reference with oracle/make_golden.py for the GPU box.
* MAE / histograms: defined on the reference's own `diff_i32`
  (run_codec.py:275); trivially pinned by that definition.
* Gaussian SSIM: PARITY UNPINNED AGAINST SCIKIT-IMAGE (it is not installed
  here, and the reference has no windowed SSIM of its own); pinned instead on
  a third-party computation: tests/golden/ssimw_cv2.npz holds values obtained
  with OpenCV's GaussianBlur(11 x 11, sigma 1.5) -- the formulation of OpenCV's
  own SSIM sample -- plus skimage's 5-px crop (oracle/make_golden_ssim_cv2.py),
  and both statements below agree with them to 1e-15.  The definition restates
  skimage.metrics.structural_similarity(gaussian_weights=True, sigma=1.5,
  use_sample_covariance=False) on top of scipy.ndimage.gaussian_filter;
  ssim_gaussian_band_direct() states the same quantity a second time from the
  SSIM paper's definition (direct 11 x 11 window sums, no scipy).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  It is the checker, never the product: the product
package (image_compression_analysis_b200) does not import it and has no CPU
path of its own.

The numpy operations mirror the reference's (same temporaries, same order) so
that (a) floating-point results agree to the last bit wherever the reference is
deterministic and (b) timing this file is a fair stand-in for timing the
reference's CPU path on a box where /root/reference is not mounted.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------
# scalar metrics (run_codec.py:55-80)
# --------------------------------------------------------------------------


def mse(a: np.ndarray, b: np.ndarray) -> float:
    """run_codec.py:55-57 -- float64 mean of squared differences."""
    e = a.astype(np.float64) - b.astype(np.float64)
    return float(np.mean(e * e))


def psnr(a: np.ndarray, b: np.ndarray, data_range: float) -> float:
    """run_codec.py:60-64 -- inf for identical inputs, math.log10 otherwise."""
    m = mse(a, b)
    if m == 0:
        return float("inf")
    return 20.0 * math.log10(data_range) - 10.0 * math.log10(m)


def ssim_global(a: np.ndarray, b: np.ndarray, data_range: float) -> float:
    """run_codec.py:67-80 -- window-less SSIM from global moments, clamped to [0,1]."""
    x = a.astype(np.float64)
    y = b.astype(np.float64)
    mx = float(np.mean(x))
    my = float(np.mean(y))
    vx = float(np.var(x))
    vy = float(np.var(y))
    cxy = float(np.mean((x - mx) * (y - my)))
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    num = (2 * mx * my + c1) * (2 * cxy + c2)
    den = (mx ** 2 + my ** 2 + c1) * (vx + vy + c2)
    if den == 0:
        return 1.0
    return max(0.0, min(1.0, num / den))


# --------------------------------------------------------------------------
# data range heuristic (run_codec.py:86-117)
# --------------------------------------------------------------------------


def effective_data_range(ref: np.ndarray) -> int:
    """run_codec.py:86-117 on a (B,H,W) array: unmasked scan of the whole cube."""
    dt = str(ref.dtype)
    if dt == "uint8":
        return 255
    if dt == "uint16":
        packed12 = True
        top = 0
        for band in ref:
            top = max(top, int(band.max()))
            if packed12 and np.any((band & 0xF) != 0):
                packed12 = False
        return 4095 if (packed12 and top <= 4095 * 16) else 65535
    if dt == "int16":
        packed14 = True
        lo, hi = 0, 0
        for band in ref:
            lo = min(lo, int(band.min()))
            hi = max(hi, int(band.max()))
            if packed14 and np.any((band & 0x3) != 0):
                packed14 = False
        if packed14 and lo >= -8192 and hi <= 8191:
            return 8191
        return int(max(abs(lo), abs(hi)))
    try:
        return int(np.iinfo(ref.dtype).max)
    except Exception:
        return 65535


# --------------------------------------------------------------------------
# masks
# --------------------------------------------------------------------------


def dataset_mask(cube: np.ndarray, nodata=None, mask=None) -> np.ndarray:
    """rasterio DatasetReader.dataset_mask() semantics as bool (H,W) (third-party;
    see oracle/rasterio_stub.py): explicit mask, else any-band != nodata, else all valid."""
    if mask is not None:
        return np.asarray(mask) > 0
    if nodata is not None and np.isfinite(nodata):
        return np.any(cube != nodata, axis=0)
    return np.ones(cube.shape[1:], bool)


def metrics_valid_mask(ref, tst, valid=None, ref_nodata=None, tst_nodata=None,
                       ref_mask=None, tst_mask=None) -> np.ndarray:
    """run_codec.py:249-263 -- dataset masks AND every band != nodata (both cubes) AND `valid`."""
    B, H, W = ref.shape
    vm = dataset_mask(ref, ref_nodata, ref_mask) & dataset_mask(tst, tst_nodata, tst_mask)
    if ref_nodata is not None and np.isfinite(ref_nodata):
        for band in ref:
            vm &= (band != ref_nodata)
    if tst_nodata is not None and np.isfinite(tst_nodata):
        for band in tst:
            vm &= (band != tst_nodata)
    if valid is not None:
        if valid.shape != (H, W):
            raise ValueError(f"Mask shape {valid.shape} != {(H, W)}")
        vm &= valid.astype(bool)
    return vm


def quicklook_valid_mask(cube, nodata=None, mask=None) -> np.ndarray:
    """quicklooks.py:35-45 -- dataset mask AND band 1 != nodata (first band only)."""
    m = dataset_mask(cube, nodata, mask).copy()
    if nodata is not None and np.isfinite(nodata):
        m &= (cube[0] != nodata)
    return m


# --------------------------------------------------------------------------
# compute_metrics (run_codec.py:240-304) + MAE / histogram additions
# --------------------------------------------------------------------------


def compute_metrics(ref: np.ndarray, tst: np.ndarray, valid: Optional[np.ndarray] = None, *,
                    ref_nodata=None, tst_nodata=None, ref_mask=None, tst_mask=None,
                    hist_bins: int = 0, extras: bool = True) -> Dict[str, object]:
    """Array-level compute_metrics.  ref/tst are (B,H,W).

    Keys psnr_b{i}, ssim_b{i}, maxerr_b{i}, psnr_band_avg, ssim_band_avg,
    psnr_global, ssim_global, max_abs_err, lossless follow run_codec.py:293-303.
    With extras=True also: mae_b{i}, mae_global, sse_b{i} (int), n_valid (int)
    and, when hist_bins=K>0, hist_b{i} = bincount(min(|d|,K-1)).
    """
    assert ref.shape == tst.shape and ref.ndim == 3, "Reference and test must match in size and band count."
    B, H, W = ref.shape
    rng = effective_data_range(ref)
    vm = metrics_valid_mask(ref, tst, valid, ref_nodata, tst_nodata, ref_mask, tst_mask)
    use_mask = bool(np.any(vm))                      # :264 all-False mask => evaluate everything

    psnrs, ssims, maxerrs = [], [], []
    maes, sses, hists = [], [], []
    sse_total = 0.0
    n_total = 0
    rng_obs = 0.0
    abs_total = 0
    for i in range(B):
        A = ref[i]
        R = tst[i]
        if use_mask:
            a = A[vm]
            r = R[vm]
        else:
            a = A
            r = R
        dabs = np.abs(a.astype(np.int32) - r.astype(np.int32))          # :275
        maxerrs.append(int(dabs.max()) if dabs.size else 0)              # :276
        psnrs.append(psnr(a, r, rng) if a.size else float("nan"))        # :278
        ssims.append(ssim_global(a, r, rng) if a.size else float("nan")) # :279
        e = a.astype(np.float64) - r.astype(np.float64)                  # :281
        sse_total += float(np.sum(e * e))                                # :282
        n_total += int(a.size)                                           # :283
        if a.size:
            rng_obs = max(rng_obs, float(np.max(np.abs(a))), float(np.max(np.abs(r))))  # :285
        if extras:
            s_abs = int(dabs.sum(dtype=np.int64))
            abs_total += s_abs
            maes.append(s_abs / a.size if a.size else float("nan"))
            sses.append(int((dabs.astype(np.int64) ** 2).sum()))
            if hist_bins:
                hists.append(np.bincount(np.minimum(dabs, hist_bins - 1).ravel(),
                                         minlength=hist_bins).astype(np.int64))
    if n_total > 0:
        rng_use = float(max(rng, rng_obs)) if np.isfinite(rng) else float(rng_obs)  # :287
        if sse_total == 0.0:
            psnr_total = float("inf")
        else:
            psnr_total = 20.0 * math.log10(rng_use) - 10.0 * math.log10(sse_total / n_total)
    else:
        psnr_total = float("nan")
    with np.errstate(all="ignore"):
        out: Dict[str, object] = {
            "psnr_band_avg": float(np.nanmean(psnrs)) if psnrs else float("nan"),
            "ssim_band_avg": float(np.nanmean(ssims)) if ssims else float("nan"),
            "psnr_global": psnr_total,
            "ssim_global": float(np.nanmean(ssims)) if ssims else float("nan"),
            "max_abs_err": int(max(maxerrs)) if maxerrs else 0,
            "lossless": 1 if max(maxerrs) == 0 else 0,
        }
    for i in range(B):
        out[f"psnr_b{i+1}"] = psnrs[i]
        out[f"ssim_b{i+1}"] = ssims[i]
        out[f"maxerr_b{i+1}"] = maxerrs[i]
    if extras:
        out["n_valid"] = n_total // B if B else 0
        out["data_range"] = rng
        out["mae_global"] = (abs_total / n_total) if n_total else float("nan")
        for i in range(B):
            out[f"mae_b{i+1}"] = maes[i]
            out[f"sse_b{i+1}"] = sses[i]
            if hist_bins:
                out[f"hist_b{i+1}"] = hists[i]
    return out


# --------------------------------------------------------------------------
# Case-B spectral metrics (run_codec.py:308-347) and Sobel (run_codec.py:123-137)
# --------------------------------------------------------------------------


def sobel_mag(img: np.ndarray) -> np.ndarray:
    """run_codec.py:123-137 -- 3x3 Sobel magnitude, edge-replicated border, float64."""
    f = img.astype(np.float64)
    kx = np.array([[1, 0, -1], [2, 0, -2], [1, 0, -1]], dtype=np.float64)
    ky = np.array([[1, 2, 1], [0, 0, 0], [-1, -2, -1]], dtype=np.float64)
    p = np.pad(f, ((1, 1), (1, 1)), mode="edge")
    H, W = f.shape
    gx = np.zeros_like(f)
    gy = np.zeros_like(f)
    for di in range(3):
        for dj in range(3):
            win = p[di:di + H, dj:dj + W]
            gx += kx[di, dj] * win
            gy += ky[di, dj] * win
    return np.sqrt(gx * gx + gy * gy)


def compute_sam_sid_lmse_caseB(ref: np.ndarray, tst: np.ndarray, valid: Optional[np.ndarray] = None, *,
                               ref_nodata=None, tst_nodata=None, ref_mask=None, tst_mask=None
                               ) -> Dict[str, float]:
    """Array-level compute_sam_sid_lmse_caseB.  ref/tst are (B,H,W).

    Mask = `valid` if given else the two dataset masks (NO per-band nodata test,
    NO all-False fallback; run_codec.py:314-319).  LMSE ignores the mask (:341-346).
    """
    B, H, W = ref.shape
    A = ref.astype(np.float64)
    R = tst.astype(np.float64)
    if valid is not None:
        if valid.shape != (H, W):
            raise ValueError("Mask shape mismatch for Case B metrics")
        vm = valid.astype(bool)
    else:
        vm = dataset_mask(ref, ref_nodata, ref_mask) & dataset_mask(tst, tst_nodata, tst_mask)
    sel = vm.ravel()
    A2 = A.reshape(B, -1)[:, sel]          # F-ordered like the reference's fancy index (:322)
    R2 = R.reshape(B, -1)[:, sel]
    n = A2.shape[1]
    if n == 0:
        return {"sam_deg": float("nan"), "sid": float("nan"), "lmse": float("nan")}
    dot = np.sum(A2 * R2, axis=0)                                        # :328
    na = np.sqrt(np.sum(A2 * A2, axis=0)) + 1e-12                        # :329
    nr = np.sqrt(np.sum(R2 * R2, axis=0)) + 1e-12                        # :330
    cosang = np.clip(dot / (na * nr), -1.0, 1.0)                         # :331
    sam_deg = float(np.degrees(np.mean(np.arccos(cosang))))              # :332
    Ap = A2 - A2.min(axis=0) + 1e-12                                     # :334-335
    Rp = R2 - R2.min(axis=0) + 1e-12
    Ap /= np.sum(Ap, axis=0, keepdims=True)                              # :336
    Rp /= np.sum(Rp, axis=0, keepdims=True)                              # :337
    sid = float(np.mean(np.sum(Ap * np.log((Ap + 1e-15) / (Rp + 1e-15)), axis=0) +
                        np.sum(Rp * np.log((Rp + 1e-15) / (Ap + 1e-15)), axis=0)))  # :338-339
    acc = 0.0
    for b in range(B):                                                   # :342-345
        acc += mse(sobel_mag(A[b]), sobel_mag(R[b]))
    return {"sam_deg": sam_deg, "sid": sid, "lmse": float(acc / B)}


def sam_caseB(ref: np.ndarray, tst: np.ndarray, valid: Optional[np.ndarray] = None) -> float:
    """SAM part alone of compute_sam_sid_lmse_caseB (run_codec.py:312-332): the float64 casts,
    the mask gather and the arccos mean exactly as the reference performs them, without the SID
    and LMSE that the reference function always computes alongside.  bench.py's CPU arm uses it
    when the GPU step it is compared with evaluates SAM only."""
    B, H, W = ref.shape
    A = ref.astype(np.float64)
    R = tst.astype(np.float64)
    vm = valid.astype(bool) if valid is not None else np.ones((H, W), bool)
    sel = vm.ravel()
    A2 = A.reshape(B, -1)[:, sel]
    R2 = R.reshape(B, -1)[:, sel]
    if A2.shape[1] == 0:
        return float("nan")
    dot = np.sum(A2 * R2, axis=0)
    na = np.sqrt(np.sum(A2 * A2, axis=0)) + 1e-12
    nr = np.sqrt(np.sum(R2 * R2, axis=0)) + 1e-12
    cosang = np.clip(dot / (na * nr), -1.0, 1.0)
    return float(np.degrees(np.mean(np.arccos(cosang))))


# --------------------------------------------------------------------------
# error quicklooks (quicklooks.py:115-207) -- pixel content only, no file I/O
# --------------------------------------------------------------------------


def err8_lut(cap: int) -> np.ndarray:
    """quicklooks.py:136-150 `to_err8` with cap given, tabulated for err = 0..cap.

    The reference evaluates clip((err-0.0)/(cap-0.0+1e-9),0,1)*255.0 on a float32
    array and truncates to uint8; err >= cap saturates at 255, so a (cap+1)-entry
    table indexed by min(err,cap) reproduces it exactly for every integer err.
    """
    cap = int(cap)
    e = np.arange(cap + 1, dtype=np.int64).astype(np.float32)
    lo, hi = 0.0, float(cap)
    e8 = np.clip((e - lo) / (hi - lo + 1e-9), 0, 1) * 255.0
    return e8.astype(np.uint8)


def error_max8(ref: np.ndarray, tst: np.ndarray, err_max_global: Optional[int] = 255,
               err_max_zoom: Optional[int] = None, *, ref_nodata=None, tst_nodata=None,
               ref_mask=None, tst_mask=None) -> Dict[str, object]:
    """Pixel content of write_error_max8 (quicklooks.py:123-205) for fixed caps.

    Returns err (float32 (H,W)), valid (bool), err8_g / err8_z (uint8 (H,W)),
    cap_g / cap_z (int, the value in the file name) and mean/std tags.
    """
    A = ref.astype(np.int32)
    Bc = tst.astype(np.int32)
    assert A.shape == Bc.shape, "Dims/band count must match"
    valid = quicklook_valid_mask(ref, ref_nodata, ref_mask) & quicklook_valid_mask(tst, tst_nodata, tst_mask)
    err = np.max(np.abs(A - Bc), axis=0).astype(np.float32)              # :133
    err[~valid] = 0.0                                                    # :134

    def scale(cap):
        if cap is None:                                                  # :137-146 (pct = (2, 98), the CLI's branch)
            nz = err[err > 0]
            if nz.size:
                lo, hi = np.percentile(nz, (2, 98))
                if not np.isfinite(lo):
                    lo = 0.0
                if (not np.isfinite(hi)) or hi <= lo:
                    hi = lo + 1.0
            else:
                lo, hi = 0.0, 1.0
        else:
            lo, hi = 0.0, float(cap)
        e8 = np.clip((err - lo) / (hi - lo + 1e-9), 0, 1) * 255.0        # :149
        return e8.astype(np.uint8), int(round(hi))

    out: Dict[str, object] = {"err": err, "valid": valid}
    g, cap_g = scale(err_max_global)
    out.update(err8_g=g, cap_g=cap_g, mean_g=float(g.mean()), std_g=float(g.std()))
    if err_max_zoom is not None:
        z, cap_z = scale(err_max_zoom)
        out.update(err8_z=z, cap_z=cap_z, mean_z=float(z.mean()), std_z=float(z.std()))
    else:
        out.update(err8_z=None, cap_z=None)
    return out


# --------------------------------------------------------------------------
# Gaussian-window SSIM (addition x1; pinned on OpenCV-computed fixtures, not on skimage itself: see header)
# --------------------------------------------------------------------------

SSIMW_SIGMA = 1.5
SSIMW_TRUNCATE = 3.5
SSIMW_RADIUS = int(SSIMW_TRUNCATE * SSIMW_SIGMA + 0.5)   # 5 -> 11 taps


def gaussian_taps() -> np.ndarray:
    """The 11 normalised taps scipy.ndimage.gaussian_filter(sigma=1.5, truncate=3.5) uses."""
    x = np.arange(-SSIMW_RADIUS, SSIMW_RADIUS + 1, dtype=np.float64)
    w = np.exp(-0.5 / (SSIMW_SIGMA * SSIMW_SIGMA) * x ** 2)
    return w / w.sum()


def ssim_gaussian_band(a: np.ndarray, b: np.ndarray, data_range: float) -> float:
    """skimage structural_similarity(gaussian_weights=True, sigma=1.5,
    use_sample_covariance=False, data_range=L) on one (H,W) band, float64."""
    from scipy.ndimage import gaussian_filter

    x = a.astype(np.float64)
    y = b.astype(np.float64)
    kw = dict(sigma=SSIMW_SIGMA, truncate=SSIMW_TRUNCATE, mode="reflect")
    ux = gaussian_filter(x, **kw)
    uy = gaussian_filter(y, **kw)
    uxx = gaussian_filter(x * x, **kw)
    uyy = gaussian_filter(y * y, **kw)
    uxy = gaussian_filter(x * y, **kw)
    vx = uxx - ux * ux
    vy = uyy - uy * uy
    vxy = uxy - ux * uy
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    p = SSIMW_RADIUS
    return float(s[p:-p, p:-p].mean(dtype=np.float64))


def ssim_gaussian_band_direct(a: np.ndarray, b: np.ndarray, data_range: float) -> float:
    """SECOND, independent statement of the same definition, straight from the SSIM paper (Wang, Bovik, Sheikh,
    Simoncelli 2004, eq. 13 with an 11 x 11 circular-symmetric Gaussian window, sigma 1.5, unit sum) with no
    scipy.ndimage: the 121-term 2-D weighted sums are written out for every window that lies inside the image
    -- which are exactly the windows skimage keeps after its 5-px crop -- so no boundary rule, no separable
    filtering and no truncation logic is shared with ssim_gaussian_band().  The two must agree to rounding
    (tests/test_oracle_golden.py); the CUDA kernel is checked against both.  O(121 N): small images only."""
    x = a.astype(np.float64)
    y = b.astype(np.float64)
    H, W = x.shape
    r = SSIMW_RADIUS
    if H <= 2 * r or W <= 2 * r:
        return float("nan")
    k = np.arange(-r, r + 1, dtype=np.float64)
    w2 = np.exp(-(k[:, None] ** 2 + k[None, :] ** 2) / (2.0 * SSIMW_SIGMA ** 2))
    w2 /= w2.sum()                                  # normalised as a 2-D window
    oh, ow = H - 2 * r, W - 2 * r
    ux = np.zeros((oh, ow)); uy = np.zeros((oh, ow)); uxx = np.zeros((oh, ow)); uyy = np.zeros((oh, ow)); uxy = np.zeros((oh, ow))
    for i in range(2 * r + 1):
        for j in range(2 * r + 1):
            xs, ys, w = x[i:i + oh, j:j + ow], y[i:i + oh, j:j + ow], w2[i, j]
            ux += w * xs; uy += w * ys; uxx += w * (xs * xs); uyy += w * (ys * ys); uxy += w * (xs * ys)
    c1 = (0.01 * data_range) ** 2
    c2 = (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * (uxy - ux * uy) + c2)) / ((ux * ux + uy * uy + c1) * ((uxx - ux * ux) + (uyy - uy * uy) + c2))
    return float(s.mean(dtype=np.float64))


def ssim_gaussian(ref: np.ndarray, tst: np.ndarray, data_range: Optional[float] = None) -> Dict[str, float]:
    """Per-band Gaussian SSIM + band average under new keys ssimw_b{i} / ssimw_band_avg."""
    L = effective_data_range(ref) if data_range is None else data_range
    vals = [ssim_gaussian_band(ref[i], tst[i], L) for i in range(ref.shape[0])]
    out = {f"ssimw_b{i+1}": v for i, v in enumerate(vals)}
    out["ssimw_band_avg"] = float(np.mean(vals))
    return out
