"""numpy restatement of the rows either side of the distortion path -- TEST
INFRASTRUCTURE ONLY: imported by tests/, never by the product package.

Array-level forms of reference functions that take file paths; every function cites the reference
lines it follows (paths under /root/reference).  Pinned against the UNMODIFIED reference executed in
the build container (tests/test_adjacent_oracle.py, oracle/make_golden_adjacent.py -> tests/golden/adj_*.npz).
"""
from __future__ import annotations

import numpy as np


# ---- RGB quicklook: tools/quicklooks.py:51-109 ---------------------------------------------------
def stretch_params(bands, mvalid, pct=(2, 98)):
    """stretch_params_from_baseline (quicklooks.py:51-72) on the selected bands (3,H,W) and the valid mask."""
    bands = np.asarray(bands).astype(np.float32)
    params = []
    for i in range(bands.shape[0]):
        vals = bands[i]
        v = vals[mvalid & np.isfinite(vals)]
        if v.size == 0:
            lo, hi = 0.0, 1.0
        else:
            lo, hi = np.percentile(v, pct)
            if not np.isfinite(lo):
                lo = 0.0
            if (not np.isfinite(hi)) or hi <= lo:
                hi = lo + 1.0
        params.append((float(lo), float(hi)))
    return params


def stretch8(x, lo, hi):
    """quicklooks.py:81-83"""
    y = np.clip((x.astype(np.float32) - lo) / (hi - lo + 1e-9), 0, 1)
    return (y * 255.0).astype(np.uint8)


def rgb_8bit(b, params):
    """write_rgb_8bit's pixel content (quicklooks.py:88-89): b = ds.read(rgb_order)."""
    return np.stack([stretch8(b[i], *params[i]) for i in range(len(params))], 0)


# ---- baseline builders ---------------------------------------------------------------------------
def trunc_uint16(u16, k):
    """make_baseline_B.py:279-282"""
    if k <= 0:
        return u16
    return ((u16 >> k) << k).astype(np.uint16, copy=False)


def truncated_copy(arr, k, nodata=None):
    """write_truncated_copy's sample arithmetic (make_baseline_B.py:298-311) on a whole array."""
    ref = np.asarray(arr)
    u = ref.view(np.uint16) if ref.dtype == np.int16 else ref.astype(np.uint16, copy=False)
    ut = trunc_uint16(u, k)
    out = ut.view(np.int16).copy() if ref.dtype == np.int16 else ut.astype(ref.dtype, copy=True)
    if nodata is not None:
        out[ref == nodata] = nodata
    return out


def to_12in16(arr):
    """make_baseline_A.py:163-167: round to the nearest multiple of 16 in uint16 arithmetic."""
    arr = np.asarray(arr).astype(np.uint16, copy=False)
    return (((arr.astype(np.uint16) + 8) >> 4) << 4).astype(np.uint16, copy=False)


def scene_error_map(ref, cmp, valid, err_scale, k_bits, err_mode="mean", tile=512):
    """make_scene_error_map (make_baseline_B.py:324-419) on (B,H,W) arrays: returns (uint8 (H,W), emax)."""
    ref, cmp = np.asarray(ref), np.asarray(cmp)
    B, H, W = ref.shape
    kmax = (1 << k_bits) - 1
    nbins = kmax + 1

    def strip(r0, r1):
        h = r1 - r0
        acc = np.zeros((h, W), np.float32)
        ssq = np.zeros((h, W), np.float32)
        cnt3 = np.zeros((h, W), np.uint16)
        accmax = np.zeros((h, W), np.uint16)
        hist = np.zeros((h, W, nbins), np.uint32) if err_mode == "p95" else None
        for b in range(B):
            a = ref[b, r0:r1].astype(np.int32)
            c = cmp[b, r0:r1].astype(np.int32)
            d = np.abs(a - c)
            if valid is not None:
                d[~valid[r0:r1, 0:W]] = 0
            if err_mode == "mean":
                acc += d
            elif err_mode == "rms":
                ssq += (d * d)
            elif err_mode == "count3":
                cnt3 += (d == kmax)
            elif err_mode == "max":
                accmax = np.maximum(accmax, d.astype(np.uint16))
            elif err_mode == "p95":
                d_clip = np.clip(d, 0, kmax)
                for k in range(nbins):
                    hist[..., k] += (d_clip == k)
        if err_mode == "mean":
            return acc / B
        if err_mode == "rms":
            return np.sqrt(ssq / B)
        if err_mode == "count3":
            return cnt3.astype(np.float32)
        if err_mode == "max":
            return accmax.astype(np.float32)
        cdf = np.cumsum(hist, axis=2)
        thr = (cdf[..., -1] * 0.95).astype(np.uint32)
        out_tile = np.zeros((h, W), np.float32)
        for k in range(nbins):
            m = (cdf[..., k] >= thr) & (out_tile == 0)
            out_tile[m] = k
        return out_tile

    tiles = [(r0, min(H, r0 + tile)) for r0 in range(0, H, tile)]
    outs = [strip(r0, r1) for r0, r1 in tiles]
    global_max = 0
    for o in outs:
        global_max = max(global_max, float(o.max()))
    if err_mode == "count3":
        emax = max(1, B) if err_scale == "fixed" else max(1, int(global_max))
    else:
        emax = kmax if err_scale == "fixed" else max(1, int(np.ceil(global_max)))
    img = np.zeros((H, W), np.uint8)
    for (r0, r1), o in zip(tiles, outs):
        img[r0:r1] = (np.clip(o, 0, emax) * (255.0 / emax) + 0.5).astype(np.uint8)
    return img, emax


# ---- codec wrappers: reversible band differencing and raw interleave ------------------------------
def diff1_bsq_signed(tile_bsq):
    """ccsds121_wrap.py:66-69"""
    X = tile_bsq.view(np.uint16).astype(np.uint32, copy=False)
    R = X.copy()
    R[1:] = (X[1:] - X[:-1]) & 0xFFFF
    return R.astype(np.uint16, copy=False).view(np.int16)


def int1_bsq_signed(R):
    """ccsds121_wrap.py:71-74"""
    X = R.view(np.uint16).astype(np.uint32, copy=True)
    for b in range(1, X.shape[0]):
        X[b] = (X[b] + X[b - 1]) & 0xFFFF
    return X.astype(np.uint16, copy=False).view(np.int16)


def diff1_bsq_unsigned(tile_bsq):
    """ccsds121_wrap.py:76-79"""
    R = tile_bsq.astype(np.uint32, copy=True)
    R[1:] = (R[1:] - tile_bsq.astype(np.uint32)[:-1]) & 0xFFFF
    return R.astype(np.uint16, copy=False)


def int1_bsq_unsigned(R):
    """ccsds121_wrap.py:81-84"""
    X = R.astype(np.uint32, copy=True)
    for b in range(1, X.shape[0]):
        X[b] = (X[b] + X[b - 1]) & 0xFFFF
    return X.astype(np.uint16, copy=False)


def diff1_forward(cur, prev, dtype_str):
    """jpegls_wrap.py:92-106"""
    if prev is None:
        return cur
    if dtype_str == "uint16":
        return ((cur.astype(np.uint32) - prev.astype(np.uint32)) & 0xFFFF).astype(np.uint16)
    if dtype_str == "int16":
        return np.clip(cur.astype(np.int32) - prev.astype(np.int32), -32768, 32767).astype(np.int16)
    if dtype_str == "uint8":
        return ((cur.astype(np.uint16) - prev.astype(np.uint16)) & 0xFF).astype(np.uint8)
    return cur


def diff1_inverse(R, prev_recon, dtype_str):
    """jpegls_wrap.py:108-120"""
    if prev_recon is None:
        return R
    if dtype_str == "uint16":
        return ((R.astype(np.uint32) + prev_recon.astype(np.uint32)) & 0xFFFF).astype(np.uint16)
    if dtype_str == "int16":
        return np.clip(R.astype(np.int32) + prev_recon.astype(np.int32), -32768, 32767).astype(np.int16)
    if dtype_str == "uint8":
        return ((R.astype(np.uint16) + prev_recon.astype(np.uint16)) & 0xFF).astype(np.uint8)
    return R


def diff1_cube_forward(cube, dtype_str):
    """The JPEG-LS wrapper's band loop in lossless mode: band b is differenced against ORIGINAL band b-1."""
    out = np.empty_like(cube)
    for b in range(cube.shape[0]):
        out[b] = diff1_forward(cube[b], cube[b - 1] if b else None, dtype_str)
    return out


def diff1_cube_inverse(res, dtype_str):
    """... and decoded against the RECONSTRUCTED band b-1."""
    out = np.empty_like(res)
    for b in range(res.shape[0]):
        out[b] = diff1_inverse(res[b], out[b - 1] if b else None, dtype_str)
    return out


def to_interleave(tile_bsq, interleave):
    """Sample order _write_raw_interleaved puts in the RAW file (ccsds121_wrap.py:44-56), as a flat array."""
    if interleave == "bsq":
        return tile_bsq.reshape(-1).copy()
    if interleave == "bil":
        return np.concatenate([tile_bsq[:, r, :].reshape(-1) for r in range(tile_bsq.shape[1])])
    if interleave == "bip":
        return np.ascontiguousarray(np.moveaxis(tile_bsq, 0, -1)).reshape(-1)
    raise ValueError("interleave must be one of: bsq, bil, bip")


def from_interleave(flat, interleave, B, Ht, Wt):
    """_read_raw_interleaved (ccsds121_wrap.py:58-64)"""
    if interleave == "bsq":
        return flat.reshape(B, Ht, Wt)
    if interleave == "bil":
        return np.moveaxis(flat.reshape(Ht, B, Wt), 1, 0)
    if interleave == "bip":
        return np.moveaxis(flat.reshape(Ht, Wt, B), -1, 0)
    raise ValueError("interleave must be one of: bsq, bil, bip")
