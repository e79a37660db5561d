"""Seeded synthetic original/decoded pairs of the BASELINE.json shapes (host side, numpy).

These are the inputs SURVEY.md section 8d prescribes: Sentinel-2-like "12-in-16"
uint16 tiles (values are multiples of 16, as tools/make_baseline_A.py:137-170
produces) and EnMAP-like spectrally smooth cubes with the two LSBs zeroed
("14-in-16", tools/make_baseline_B.py:281-316).  Used by tests, bench.py and
oracle/make_golden.py; nothing here is a metric.
"""
from __future__ import annotations

import numpy as np


def _smooth_field(rng, shape, passes=3):
    f = rng.standard_normal(shape)
    for _ in range(passes):                       # cheap separable low-pass
        f = (np.roll(f, 1, -1) + f + np.roll(f, -1, -1)) / 3.0
        f = (np.roll(f, 1, -2) + f + np.roll(f, -1, -2)) / 3.0
    f -= f.min()
    f /= max(float(f.max()), 1e-12)
    return f


def case_a_pair(seed=1, bands=4, height=1024, width=1024, sigma=2.0, mode="gauss"):
    """(ref, dec) uint16 (B,H,W) BSQ, 12-in-16.  mode: gauss | near3 | identical."""
    rng = np.random.default_rng(seed)
    base = (_smooth_field(rng, (bands, height, width)) * 4095.0).astype(np.int64)
    ref = (base << 4).astype(np.uint16)
    if mode == "identical":
        return ref, ref.copy()
    if mode == "near3":
        noise = rng.integers(-3, 4, size=ref.shape) * 16
    else:
        noise = np.rint(rng.standard_normal(ref.shape) * sigma).astype(np.int64) * 16
    dec = np.clip(ref.astype(np.int64) + noise, 0, 65520).astype(np.uint16)
    return ref, dec


def case_b_pair(seed=2, bands=180, height=1024, width=1024, amp=3, dtype="uint16", layout="bsq"):
    """(ref, dec) EnMAP-like cube.  ref has its 2 LSBs zeroed; dec = ref + U{-amp..amp}.

    layout "bsq" returns (B,H,W); "bip" returns (H,W,B) (C-contiguous, same samples).
    dtype "int16" keeps values in +-8191 so the 14-in-16 branch of
    effective_data_range applies.
    """
    rng = np.random.default_rng(seed)
    start = rng.integers(500, 4000, size=(height, width, 1))
    steps = rng.integers(-60, 64, size=(height, width, bands))
    steps[..., 0] = 0
    walk = start + np.cumsum(steps, axis=-1)
    if dtype == "int16":
        walk = np.clip(walk, -8000, 8000)
    else:
        walk = np.clip(walk, 0, 10000)
    ref = (walk >> 2) << 2
    noise = rng.integers(-amp, amp + 1, size=ref.shape) if amp > 0 else 0
    if dtype == "int16":
        dec = np.clip(ref + noise, -8191, 8191)
    else:
        dec = np.clip(ref + noise, 0, 65535)
    ref = ref.astype(dtype)
    dec = dec.astype(dtype)
    if layout == "bip":
        return np.ascontiguousarray(ref), np.ascontiguousarray(dec)
    return (np.ascontiguousarray(np.moveaxis(ref, -1, 0)),
            np.ascontiguousarray(np.moveaxis(dec, -1, 0)))


def random_valid_mask(seed, height, width, frac_invalid=0.05):
    rng = np.random.default_rng(seed)
    return rng.random((height, width)) >= frac_invalid


def plant_nodata(cube_bsq, nodata, mask_invalid, extra_hits=0, seed=0):
    """Write `nodata` into every band where mask_invalid, plus `extra_hits` single-band hits."""
    out = cube_bsq.copy()
    out[:, mask_invalid] = nodata
    if extra_hits:
        rng = np.random.default_rng(seed)
        B, H, W = out.shape
        for _ in range(extra_hits):
            out[rng.integers(B), rng.integers(H), rng.integers(W)] = nodata
    return out
