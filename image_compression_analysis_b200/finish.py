"""Host-side finish: exact integer partials -> the reference's metric values.

Pure functions of small integer / float64 vectors (no CUDA, no torch): whatever produced the
partials -- one GPU, row strips on eight GPUs combined by allreduce -- the finish is identical,
which is what makes the integer-derived metrics bit-stable across GPU counts.

Reference arithmetic mirrored here (file:line under /root/reference/tools):
  psnr / mse                run_codec.py:55-64     float(SSE)/N and math.log10: bit-exact
  ssim_global               run_codec.py:67-80     from {N,Sx,Sy,Sxx,Syy,Sxy} in exact rationals
  effective_data_range      run_codec.py:86-117
  compute_metrics tail      run_codec.py:286-303
  SAM / SID / LMSE means    run_codec.py:332, 338-346
  to_err8                   quicklooks.py:136-150  (tabulated: err8_lut)
"""
from __future__ import annotations

import math
import warnings
from typing import Dict, Optional, Sequence

import numpy as np

from ._lib import (DM_I16, DM_M_ABSXY, DM_M_LOW2, DM_M_LOW4, DM_M_MAXERR, DM_M_UMAX, DM_M_UNEGMIN, DM_NSTAT,
                   DM_S_ABS, DM_S_N, DM_S_SSE, DM_S_X, DM_S_XX, DM_S_XY, DM_S_Y, DM_S_YY, DM_U8, DM_U16)


def data_range_from_maxs(dtype_code: int, maxs: np.ndarray) -> int:
    """effective_data_range (run_codec.py:86-117) from the cube-wide maxima of the fused pass."""
    m = np.asarray(maxs, dtype=np.int64).reshape(-1, DM_NSTAT).max(axis=0)
    if dtype_code == DM_U8:
        return 255
    if dtype_code == DM_U16:
        top = int(m[DM_M_UMAX])
        return 4095 if (int(m[DM_M_LOW4]) == 0 and top <= 4095 * 16) else 65535
    if dtype_code == DM_I16:
        mn, mx = -int(m[DM_M_UNEGMIN]), int(m[DM_M_UMAX])
        if int(m[DM_M_LOW2]) == 0 and mn >= -8192 and mx <= 8191:
            return 8191
        return int(max(abs(mn), abs(mx)))
    raise ValueError(f"unsupported dtype code {dtype_code}")


def psnr_from_sse(sse: int, n: int, data_range: float) -> float:
    """psnr() (run_codec.py:60-64): mse = float(sum d^2)/N is exact, logs are math.log10."""
    if n == 0:
        return float("nan")
    m = float(sse) / n
    if m == 0:
        return float("inf")
    return 20.0 * math.log10(data_range) - 10.0 * math.log10(m)


def ssim_global_from_moments(n: int, sx: int, sy: int, sxx: int, syy: int, sxy: int, data_range: float) -> float:
    """ssim_global() (run_codec.py:67-80) from exact integer moments.

    Means, (population) variances and the covariance are formed as exact rationals and rounded
    once; the reference's two-pass float64 evaluation agrees to ~1e-15 relative.
    """
    if n == 0:
        return float("nan")
    mu_x = sx / n
    mu_y = sy / n
    n2 = n * n
    sigma_x2 = (n * sxx - sx * sx) / n2
    sigma_y2 = (n * syy - sy * sy) / n2
    sigma_xy = (n * sxy - sx * sy) / n2
    L = data_range
    C1 = (0.01 * L) ** 2
    C2 = (0.03 * L) ** 2
    num = (2 * mu_x * mu_y + C1) * (2 * sigma_xy + C2)
    den = (mu_x ** 2 + mu_y ** 2 + C1) * (sigma_x2 + sigma_y2 + C2)
    if den == 0:
        return 1.0
    return max(0.0, min(1.0, num / den))


def _nanmean(vals: Sequence[float]) -> float:
    if not len(vals):
        return float("nan")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            return float(np.nanmean(vals))


def finish_compute_metrics(dtype_code: int, sums: np.ndarray, maxs: np.ndarray,
                           hist: Optional[np.ndarray] = None, extras: bool = False) -> Dict[str, object]:
    """The dict compute_metrics returns (run_codec.py:286-303) from per-band integer partials.

    sums / maxs: int64 (B, DM_NSTAT).  With extras=True the additions of SURVEY.md 8a (x2, x3) are
    returned under new keys: mae_b{i}, mae_global, sse_b{i}, n_valid, data_range, hist_b{i}.
    """
    S = np.asarray(sums, dtype=np.int64).reshape(-1, DM_NSTAT)
    M = np.asarray(maxs, dtype=np.int64).reshape(-1, DM_NSTAT)
    B = S.shape[0]
    rng = data_range_from_maxs(dtype_code, M)
    psnrs, ssims, maxerrs = [], [], []
    sse_total = 0.0
    n_total = 0
    abs_total = 0
    for b in range(B):
        n, sse = int(S[b, DM_S_N]), int(S[b, DM_S_SSE])
        maxerrs.append(int(M[b, DM_M_MAXERR]))
        psnrs.append(psnr_from_sse(sse, n, rng))
        ssims.append(ssim_global_from_moments(n, int(S[b, DM_S_X]), int(S[b, DM_S_Y]), int(S[b, DM_S_XX]),
                                              int(S[b, DM_S_YY]), int(S[b, DM_S_XY]), rng))
        sse_total += float(sse)          # run_codec.py:282 accumulates a Python float
        n_total += n
        abs_total += int(S[b, DM_S_ABS])
    rng_obs = float(max(0, int(M[:, DM_M_ABSXY].max()))) if B else 0.0
    if n_total > 0:
        rng_use = float(max(rng, rng_obs))
        psnr_total = float("inf") if sse_total == 0.0 else (
            20.0 * math.log10(rng_use) - 10.0 * math.log10(sse_total / n_total))
    else:
        psnr_total = float("nan")
    out: Dict[str, object] = {
        "psnr_band_avg": _nanmean(psnrs),
        "ssim_band_avg": _nanmean(ssims),
        "psnr_global": psnr_total,
        "ssim_global": _nanmean(ssims),
        "max_abs_err": int(max(maxerrs)) if maxerrs else 0,
        "lossless": 1 if (maxerrs and max(maxerrs) == 0) else 0,
    }
    for i, (p, s, me) in enumerate(zip(psnrs, ssims, maxerrs), start=1):
        out[f"psnr_b{i}"] = p
        out[f"ssim_b{i}"] = s
        out[f"maxerr_b{i}"] = me
    if extras:
        out["n_valid"] = int(S[0, DM_S_N]) if B else 0
        out["data_range"] = rng
        out["mae_global"] = (abs_total / n_total) if n_total else float("nan")
        for b in range(B):
            n = int(S[b, DM_S_N])
            out[f"mae_b{b+1}"] = (int(S[b, DM_S_ABS]) / n) if n else float("nan")
            out[f"sse_b{b+1}"] = int(S[b, DM_S_SSE])
            if hist is not None:
                out[f"hist_b{b+1}"] = np.asarray(hist[b], dtype=np.int64).copy()
    return out


def finish_spectral(sum_acos: float, sum_sid: float, n: float, lmse_band_sums: Optional[np.ndarray],
                    npix_image: int) -> Dict[str, float]:
    """sam_deg / sid / lmse (run_codec.py:325-346) from float64 partial sums.

    n == 0 -> all three NaN (run_codec.py:325-326).  LMSE is unmasked: per band
    sum / (H*W), then the mean over bands.
    """
    if n == 0:
        return {"sam_deg": float("nan"), "sid": float("nan"), "lmse": float("nan")}
    out = {"sam_deg": float(np.degrees(sum_acos / n)), "sid": float(sum_sid / n)}
    if lmse_band_sums is not None:
        acc = 0.0
        for v in np.asarray(lmse_band_sums, dtype=np.float64):
            acc += float(v) / npix_image
        out["lmse"] = float(acc / len(lmse_band_sums))
    else:
        out["lmse"] = float("nan")
    return out


def finish_ssim_gauss(band_sums: np.ndarray, band_counts: np.ndarray) -> Dict[str, float]:
    """ssimw_b{i} / ssimw_band_avg (addition x1) from per-band {sum S, count}."""
    vals = []
    for s, c in zip(np.asarray(band_sums, dtype=np.float64), np.asarray(band_counts, dtype=np.float64)):
        vals.append(float(s / c) if c > 0 else float("nan"))
    out = {f"ssimw_b{i+1}": v for i, v in enumerate(vals)}
    with np.errstate(all="ignore"):
        out["ssimw_band_avg"] = float(np.mean(vals)) if vals else float("nan")
    return out


def err8_lut(cap: float) -> np.ndarray:
    """to_err8 with a fixed cap (quicklooks.py:136-150), tabulated for integer err = 0..ceil(cap).

    The reference evaluates clip((err - 0.0) / (cap - 0.0 + 1e-9), 0, 1) * 255.0 on a float32 array
    and truncates to uint8; every err >= cap saturates at 255, so indexing this table with
    min(err, len-1) reproduces the plane exactly.  The expression below is the reference's own.
    """
    hi = float(cap)
    top = int(math.ceil(hi)) if hi > 0 else 0
    top = max(1, min(top, 65535))      # at least {0, 1}: err >= 1 must still saturate when cap <= 0
    err = np.arange(top + 1, dtype=np.int64).astype(np.float32)
    lo = 0.0
    e8 = np.clip((err - lo) / (hi - lo + 1e-9), 0, 1) * 255.0
    return e8.astype(np.uint8)


def err8_stats_from_hist(hist256: np.ndarray) -> Dict[str, float]:
    """STATISTICS_MEAN / STATISTICS_STDDEV tags (quicklooks.py:175-184) from the plane's histogram.

    numpy evaluates uint8.mean()/std() in float64 over the plane; the histogram gives the same
    moments as exact integers (differences are at the 1e-16 level).
    """
    h = np.asarray(hist256, dtype=np.int64)
    n = int(h.sum())
    if n == 0:
        return {"mean": float("nan"), "std": float("nan")}
    k = np.arange(256, dtype=np.int64)
    s1 = int((h * k).sum())
    s2 = int((h * k * k).sum())
    mean = s1 / n
    var = (n * s2 - s1 * s1) / (n * n)
    return {"mean": mean, "std": math.sqrt(var)}


# ---- RGB quicklook ------------------------------------------------------------------
def percentiles_from_hist(hist: np.ndarray, first_value: int, pct) -> Optional[tuple]:
    """np.percentile(v, pct) of the float32 samples behind a value histogram (quicklooks.py:63), without
    the samples: hist[k] = number of samples equal to first_value + k.

    numpy's "linear" method needs only n and the two order statistics around (n-1)*q; both come
    straight from the cumulative histogram.  The arithmetic below is numpy's own for a float32 array
    and float64 quantiles (np.lib._function_base_impl._quantile/_lerp): the difference of the two
    neighbours in float32, everything else in float64, and the mirrored form for weights >= 0.5.
    Returns None for an empty selection."""
    h = np.asarray(hist, dtype=np.int64)
    n = int(h.sum())
    if n == 0:
        return None
    cum = np.cumsum(h)
    q = np.true_divide(np.asanyarray(pct), np.float32(100))          # as np.percentile prepares it
    out = []
    for qi in np.atleast_1d(q):
        vi = (n - 1) * qi
        if vi >= n - 1:
            ip = inn = n - 1
        elif vi < 0:
            ip = inn = 0
        else:
            ip = int(np.floor(vi))
            inn = ip + 1
        a = np.float32(first_value + int(np.searchsorted(cum, ip, side="right")))
        b = np.float32(first_value + int(np.searchsorted(cum, inn, side="right")))
        prev = np.floor(vi) if 0 <= vi < n - 1 else (np.float64(-1) if vi >= n - 1 else np.float64(0))
        t = np.float64(vi - prev)
        diff = np.subtract(b, a)                                      # float32
        r = np.float64(a) + np.float64(diff) * t
        if t >= 0.5:
            r = np.float64(b) - np.float64(diff) * (1 - t)
        out.append(float(r))
    return tuple(out)


def stretch8_lut(lo: float, hi: float, np_dtype) -> np.ndarray:
    """stretch8 of write_rgb_8bit (quicklooks.py:81-83) tabulated over every value of the sample type,
    in bin order (int16: -32768 first) -- the expression is the reference's own, on a float32 array."""
    info = np.iinfo(np.dtype(np_dtype))
    x = np.arange(info.min, info.max + 1, dtype=np.int64).astype(np.dtype(np_dtype))
    y = np.clip((x.astype(np.float32) - lo) / (hi - lo + 1e-9), 0, 1)
    lut = (y * 255.0).astype(np.uint8)
    if lut.size < 65536:
        lut = np.concatenate([lut, np.zeros(65536 - lut.size, np.uint8)])
    return lut
