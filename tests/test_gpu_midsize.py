"""GPU parity at BASELINE sizes the numpy oracle still finishes in seconds.

The goldens are <= 48 x 64 pixels; at that size every grid-stride / software-pipelined loop of the FP64 kernels
runs at most ONE iteration per block or warp (the Gaussian SSIM kernel has 296 blocks per band, the SID kernel
strides 9 472 warps, the BIP Sobel kernel marches along x with a one-column prefetch).  The cases here are sized so
that every block / warp sees several tiles / pixels, and they compare with the oracle, not with another kernel:

  * BASELINE configs[0] and configs[2] at their real size, 1024 x 1024 x 4 (Case A): compute_metrics, the ERR8
    planes at caps 255 + 32, per-band Gaussian SSIM (608 tiles per band > 296 blocks)
  * Case B at 256 x 256 x 180, BIP and BSQ, uint16 and int16 + nodata: SAM / SID / LMSE (>= 6 pixels per warp)
  * the Gaussian SSIM against the SECOND, scipy-free oracle (direct 11 x 11 window sums)
  * the NVLink peer-memory exchange against NCCL, run -- not skipped -- whenever two GPUs are visible

Bars as everywhere: integers bit-exact, PSNR bit-exact, SSIM / SAM / SID / LMSE / Gaussian SSIM 1e-6 relative."""
import math
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from tests import goldenio

pytestmark = pytest.mark.gpu

REL = 1e-6
ROOT = Path(__file__).resolve().parent.parent


def _close(g, w, rel=REL):
    return goldenio.close(g, w, rel=rel)


def _check(got, want):
    for k, w in want.items():
        if isinstance(w, np.ndarray):
            assert np.array_equal(got[k], w), k
            continue
        g = got[k]
        if isinstance(w, (int, np.integer)):
            assert g == int(w), (k, g, w)
        elif k.startswith("psnr"):
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)
        else:
            assert _close(g, w), (k, g, w)


def _bip(x):
    return np.ascontiguousarray(np.moveaxis(x, 0, -1))


@pytest.mark.parametrize("mode", ["gauss", "near3", "identical"])
def test_config1_case_a_tile_full_size_vs_oracle(mode):
    """configs[0]: Sentinel-2 1024 x 1024 x 4 uint16 12-in-16 tile: every compute_metrics key + MAE + histograms."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import synth
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_a_pair(seed=1, bands=4, height=1024, width=1024, sigma=2.0, mode=mode)
    want = orc.compute_metrics(ref, dec, hist_bins=256)
    got = dm.compute_metrics_arrays(ref, dec, hist_bins=256, extras=True)
    _check(got, want)
    assert got["lossless"] == (1 if mode == "identical" else 0)
    # masked (5 % invalid) through the same kernels
    valid = synth.random_valid_mask(9, 1024, 1024)
    _check(dm.compute_metrics_arrays(ref, dec, valid), orc.compute_metrics(ref, dec, valid, extras=False))


def test_config3_gaussian_ssim_and_err8_full_size_vs_oracle():
    """configs[2]: per-band Gaussian SSIM of the 1024 x 1024 x 4 tile (19 x 32 = 608 tiles of 54 x 32 per band on
    296 blocks: every block runs the prefetching tile loop 2-3 times) + the ERR8 quicklooks at caps 255 and 32."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import quicklooks as ql, synth
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_a_pair(seed=3, bands=4, height=1024, width=1024, sigma=2.0)
    want = orc.ssim_gaussian(ref, dec)               # data range from the cube: 4095 (12-in-16)
    got = dm.ssim_gaussian_arrays(ref, dec)
    assert set(got) == set(want)
    for k in want:
        assert _close(got[k], want[k]), (k, got[k], want[k])
    # a heavier distortion (SSIM well below 1) with the full 16-bit constants
    rng = np.random.default_rng(33)
    dec2 = np.clip(ref.astype(np.int64) + rng.integers(-3000, 3001, ref.shape), 0, 65535).astype(np.uint16)
    want = orc.ssim_gaussian(ref, dec2, 65535.0)
    got = dm.ssim_gaussian_arrays(ref, dec2, 65535.0)
    for k in want:
        assert _close(got[k], want[k]), (k, got[k], want[k])
    assert want["ssimw_band_avg"] < 0.95
    e = ql.error_max8_arrays(ref, dec, 255, 32)
    o = orc.error_max8(ref, dec, 255, 32)
    for k in ("err8_g", "err8_z", "valid"):
        assert np.array_equal(e[k], o[k]), k
    assert e["mean_g"] == o["mean_g"] and e["mean_z"] == o["mean_z"]
    assert _close(e["std_g"], o["std_g"], 1e-12) and _close(e["std_z"], o["std_z"], 1e-12)
    # all Case-A metrics from one upload (the path bench.py's C3/C4 lines time)
    allm = dm.all_metrics_arrays(ref, dec, ssim_window=True, hist_bins=256)
    _check(allm, orc.compute_metrics(ref, dec, hist_bins=256))
    for k, w in orc.ssim_gaussian(ref, dec).items():
        assert _close(allm[k], w), (k, allm[k], w)


@pytest.mark.parametrize("dtype,L,amp,shape", [("uint16", 4095.0, 48, (2, 64, 96)), ("uint16", 65535.0, 20000, (1, 33, 47)),
                                               ("int16", 8191.0, 300, (2, 40, 70)), ("uint8", 255.0, 7, (3, 29, 61))])
def test_gaussian_ssim_vs_second_oracle(dtype, L, amp, shape):
    """The kernel against the scipy-free statement of the definition (direct 11 x 11 window sums)."""
    import image_compression_analysis_b200 as dm
    from oracle import distortion_oracle as orc
    rng = np.random.default_rng(71)
    info = np.iinfo(dtype)
    a = rng.integers(info.min // 2, info.max // 2, size=shape).astype(np.int64)
    b = np.clip(a + rng.integers(-amp, amp + 1, size=shape), info.min, info.max)
    a, b = a.astype(dtype), b.astype(dtype)
    got = dm.ssim_gaussian_arrays(a, b, L)
    for i in range(shape[0]):
        w = orc.ssim_gaussian_band_direct(a[i], b[i], L)
        assert _close(got[f"ssimw_b{i+1}"], w), (i, got[f"ssimw_b{i+1}"], w)


_CASEB_CACHE = {}


def _caseb_256(dtype):
    """256 x 256 x 180 EnMAP-like pair + what the oracle says about it (computed once per dtype, ~5 s)."""
    if dtype in _CASEB_CACHE:
        return _CASEB_CACHE[dtype]
    from image_compression_analysis_b200 import synth
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_b_pair(seed=22, bands=180, height=256, width=256, amp=3, dtype=dtype, layout="bsq")
    kw = {}
    valid = None
    if dtype == "int16":
        nd = -32768
        rng = np.random.default_rng(23)
        whole = rng.random((256, 256)) < 0.05
        ref = synth.plant_nodata(ref, nd, whole, extra_hits=200, seed=1)
        dec = synth.plant_nodata(dec, nd, whole, extra_hits=200, seed=2)
        kw = dict(ref_nodata=nd, tst_nodata=nd)
    # a few pixels with a relative error above the SID kernel's series guard (5 %), zero spectra, and a block of
    # large errors: the slow branches must be reached from inside the strided loops too
    rng = np.random.default_rng(24)
    ys, xs = rng.integers(0, 256, 300), rng.integers(0, 256, 300)
    lo, hi = (0, 10000) if dtype == "uint16" else (-8000, 8000)
    dec[:, ys[:150], xs[:150]] = np.clip(dec[:, ys[:150], xs[:150]].astype(np.int64)
                                         + rng.integers(-900, 901, (180, 150)), lo, hi).astype(dtype)
    ref[:, ys[150:170], xs[150:170]] = 0
    dec[:, ys[170:190], xs[170:190]] = 0
    want = orc.compute_metrics(ref, dec, valid, extras=False, **kw)
    want.update(orc.compute_sam_sid_lmse_caseB(ref, dec, valid, **kw))
    err = orc.error_max8(ref, dec, 255, 32, **kw)
    _CASEB_CACHE.clear()
    _CASEB_CACHE[dtype] = (ref, dec, kw, want, err)
    return _CASEB_CACHE[dtype]


@pytest.mark.parametrize("layout", ["bip", "bsq"])
@pytest.mark.parametrize("dtype", ["uint16", "int16"])
def test_case_b_256x256x180_all_metrics_vs_oracle(dtype, layout):
    """65 536 pixels x 180 bands: ~7 pixels per warp of the SID kernel, 4 strides of the Sobel blocks, 1 024 tiles
    of the one-pass kernel on 148 persistent CTAs.  compute_metrics + SAM / SID / LMSE + ERR8 against the oracle,
    BIP (one-pass kernels) and BSQ (two-pass kernels), uint16 and the real EnMAP int16 + nodata -32768."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import quicklooks as ql
    ref, dec, kw, want, err = _caseb_256(dtype)
    r, d = (ref, dec) if layout == "bsq" else (_bip(ref), _bip(dec))
    got = dm.compute_metrics_arrays(r, d, layout=layout, **kw)
    got.update(dm.compute_sam_sid_lmse_caseB_arrays(r, d, layout=layout, **kw))
    _check(got, want)
    assert want["sid"] > 0 and want["lmse"] > 0 and want["sam_deg"] > 0
    e = ql.error_max8_arrays(r, d, 255, 32, layout=layout, a_nodata=kw.get("ref_nodata"), b_nodata=kw.get("tst_nodata"))
    for k in ("err8_g", "err8_z", "valid"):
        assert np.array_equal(e[k], err[k]), k
    # everything from one upload
    allm = dm.all_metrics_arrays(r, d, layout=layout, case_b=True, extras=False, **kw)
    _check(allm, want)


def test_case_b_256_with_caller_mask_vs_oracle():
    """The caller's `valid` mask (run_codec.py:260-263, :314-319) on the strided kernels: SAM / SID take the mask,
    LMSE ignores it."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import synth
    from oracle import distortion_oracle as orc
    ref, dec, kw, _, _ = _caseb_256("uint16")
    valid = synth.random_valid_mask(5, 256, 256, 0.3)
    want = orc.compute_metrics(ref, dec, valid, extras=False)
    want.update(orc.compute_sam_sid_lmse_caseB(ref, dec, valid))
    for layout in ("bip", "bsq"):
        r, d = (ref, dec) if layout == "bsq" else (_bip(ref), _bip(dec))
        got = dm.compute_metrics_arrays(r, d, valid, layout=layout)
        got.update(dm.compute_sam_sid_lmse_caseB_arrays(r, d, valid, layout=layout))
        _check(got, want)


def test_p2p_exchange_is_bit_identical_to_nccl_when_two_gpus_are_visible():
    """tools/check_p2p.py under torchrun: dm_p2p_push / dm_p2p_combine == NCCL all-gather + dm_combine_partials on
    every rank.  Runs whenever the box shows >= 2 GPUs (the 1-GPU box of the round-end suite cannot)."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip(f"{n} GPU visible: the peer-memory exchange needs two (bench.py asserts the same identity "
                    "on its own records at every N > 1)")
    world = min(n, 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29571", str(ROOT / "tools" / "check_p2p.py")],
                       capture_output=True, text=True, timeout=600, env={**os.environ, "MASTER_ADDR": "127.0.0.1"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("p2p == nccl: True; same on all ranks: True") == world


def test_strip_path_all_false_mask_means_unmasked():
    """run_codec.py:264 on the row-strip path: an all-False mask (globally) falls back to every pixel; a strip that
    merely has no valid pixel of its own does not."""
    from image_compression_analysis_b200 import finish, sharding, synth
    from image_compression_analysis_b200.engine import Want
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_a_pair(seed=8, bands=3, height=96, width=80)
    H = 96
    s = sharding.strips(H, 1)[0]
    none_valid = np.zeros((H, 80), bool)
    P = sharding.evaluate_strip(ref, dec, s, H, "bsq", Want(stats=True), none_valid)
    h = P.to_host()
    _check(finish.finish_compute_metrics(1, h.sums, h.maxs), orc.compute_metrics(ref, dec, none_valid, extras=False))
    assert int(h.sums[0, 0]) == H * 80
    # two strips combined by hand: the upper strip has no valid pixel, the lower one has -> the mask stands
    lower = none_valid.copy(); lower[60:, :] = True
    tot = None
    parts = []
    for st in sharding.strips(H, 2):
        Pp = sharding.evaluate_strip(sharding.cut_bsq(ref, st), sharding.cut_bsq(dec, st), st, H, "bsq", Want(stats=True),
                                     lower[st.buf0:st.buf1], reduce=False)
        parts.append(Pp.to_host())
    assert int(parts[0].counts[0]) == 0 and int(parts[1].counts[0]) > 0
    tot = parts[0]
    tot.isum += parts[1].isum
    tot.imax = np.maximum(tot.imax, parts[1].imax)
    _check(finish.finish_compute_metrics(1, tot.sums, tot.maxs), orc.compute_metrics(ref, dec, lower, extras=False))


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_both_gaussian_ssim_kernels_vs_oracle(variant):
    """dm_ssim_variant: 0 = auto (ring kernel for 16-bit cubes of even width, else tiled), 1 = warp-streaming kernel
    (integer horizontal pass in a 31-bit fixed-point window), 2 = tiled all-FP64 kernel, 3 = ring kernel (row-streaming,
    all FP64, register-resident vertical scatter).  Both against the scipy oracle at a size where every warp of the
    streaming kernel runs several tasks and both trip copies of its unrolled row loop (segments > 22 rows), with odd
    sizes (partial last strip / segment), int16 (offset-binary path) and uint8."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200._lib import check, lib
    from oracle import distortion_oracle as orc
    check(lib().dm_ssim_variant(variant))
    try:
        rng = np.random.default_rng(90 + variant)
        for dtype, L, amp, shape in (("uint16", 65535.0, 2500, (2, 701, 333)), ("int16", 8191.0, 120, (1, 300, 130)),
                                     ("uint8", 255.0, 6, (3, 97, 75)), ("uint16", 4095.0, 40, (1, 11, 11)),
                                     ("uint16", 4095.0, 40, (1, 12, 43)), ("uint16", 65535.0, 900, (2, 533, 398)),
                                     ("int16", 32767.0, 3000, (1, 97, 268)), ("uint16", 4095.0, 40, (1, 12, 12)),
                                     ("uint16", 4095.0, 40, (3, 11, 140))):
            info = np.iinfo(dtype)
            a = rng.integers(info.min, info.max + 1, size=shape).astype(np.int64)
            b = np.clip(a + rng.integers(-amp, amp + 1, size=shape), info.min, info.max)
            a, b = a.astype(dtype), b.astype(dtype)
            want = orc.ssim_gaussian(a, b, L)
            got = dm.ssim_gaussian_arrays(a, b, L)
            for k in want:
                assert _close(got[k], want[k]), (dtype, shape, k, got[k], want[k])
        # smooth 12-in-16 data with a small error: SSIM close to 1, variances far below the means (cancellation)
        from image_compression_analysis_b200 import synth
        ref, dec = synth.case_a_pair(seed=12, bands=2, height=400, width=300, sigma=1.0)
        want = orc.ssim_gaussian(ref, dec)
        got = dm.ssim_gaussian_arrays(ref, dec)
        for k in want:
            assert _close(got[k], want[k]), (k, got[k], want[k])
        # images no larger than the crop have no window: NaN like the oracle's empty mean
        tiny = np.zeros((1, 10, 30), np.uint16)
        assert math.isnan(dm.ssim_gaussian_arrays(tiny, tiny, 255.0)["ssimw_b1"])
    finally:
        lib().dm_ssim_variant(0)


@pytest.mark.parametrize("dtype", ["uint16", "int16", "uint8"])
def test_batched_stats_equal_pair_by_pair(dtype):
    """dm_fused_stats_batch (one launch, grid.y = pair) == dm_fused_stats per pair, bit for bit, and == the oracle."""
    import torch
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200._lib import DM_U8, DM_U16, DM_I16
    from image_compression_analysis_b200.engine import DevicePair, Partials, PreparedStatsBatch, Want, evaluate
    from oracle import distortion_oracle as orc
    rng = np.random.default_rng(7)
    B, H, W, n = 4, 120, 136, 9                  # 16 320 pixels per band: band starts stay 16-byte aligned
    info = np.iinfo(dtype)
    host = []
    for i in range(n):
        a = rng.integers(info.min, info.max + 1, size=(B, H, W)).astype(np.int64)
        b = np.clip(a + rng.integers(-(i + 1) * 40, (i + 1) * 40 + 1, size=a.shape), info.min, info.max)
        host.append((a.astype(dtype), b.astype(dtype)))
    host[3] = (host[3][0], host[3][0].copy())        # a lossless pair in the middle
    pairs = [DevicePair.from_arrays(a, b, "bsq") for a, b in host]
    run, outs = Partials.allocate_run(n, B, 0, pairs[0].ref.device, dtype)
    PreparedStatsBatch(pairs, outs).launch()
    torch.cuda.synchronize()
    code = {"uint8": DM_U8, "uint16": DM_U16, "int16": DM_I16}[dtype]
    for (a, b), pair, P in zip(host, pairs, outs):
        single = evaluate(pair, Want(stats=True)).to_host()
        h = P.to_host()
        assert np.array_equal(h.isum, single.isum) and np.array_equal(h.imax, single.imax)
        _check(finish.finish_compute_metrics(code, h.sums, h.maxs), orc.compute_metrics(a, b, extras=False))
    with pytest.raises(ValueError):
        PreparedStatsBatch(pairs[:2], outs[:1])


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_sid_sam_every_lane_grouping_vs_oracle(lanes):
    """dm_spectral_lanes_per_pixel: the register-resident SAM / SID kernel with 8, 16 or 32 lanes per pixel (four,
    two, one pixel per warp), at band counts that fill the last register word partially, with int16 + nodata,
    a caller mask (groups of a warp differ in validity) and an image whose pixel count is not a multiple of four."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import synth
    from image_compression_analysis_b200._lib import check, lib
    from oracle import distortion_oracle as orc
    check(lib().dm_spectral_lanes_per_pixel(lanes))
    try:
        for bands, dtype in ((180, "uint16"), (180, "int16"), (46, "uint16"), (242, "int16"), (256, "uint16")):
            ref, dec = synth.case_b_pair(seed=bands, bands=bands, height=37, width=53, amp=4, dtype=dtype, layout="bsq")
            rng = np.random.default_rng(bands)
            dec[:, rng.integers(0, 37, 40), rng.integers(0, 53, 40)] += 700          # large relative errors: the log() list
            ref[:, 3, 5] = 0                                                             # a zero spectrum
            valid = synth.random_valid_mask(bands, 37, 53, 0.3)
            kw = dict(ref_nodata=-32768, tst_nodata=-32768) if dtype == "int16" else {}
            for v in (None, valid):
                want = orc.compute_sam_sid_lmse_caseB(ref, dec, v, **kw)
                got = dm.compute_sam_sid_lmse_caseB_arrays(_bip(ref), _bip(dec), v, layout="bip", **kw)
                for k in ("sam_deg", "sid", "lmse"):
                    assert _close(got[k], want[k]), (lanes, bands, dtype, v is not None, k, got[k], want[k])
    finally:
        lib().dm_spectral_lanes_per_pixel(0)


def _scan_pair(dtype, H, W, ref_nd, tst_nd, seed):
    """A 180-band BIP pair whose files carry nodata values: whole invalid pixels (shared and one-sided), single-band
    hits in either cube (METRICS drops the pixel, the dataset mask keeps it), hits on band 1 (QUICKLOOK rule)."""
    from image_compression_analysis_b200 import synth
    ref, dec = synth.case_b_pair(seed=seed, bands=180, height=H, width=W, amp=3, dtype=dtype, layout="bsq")
    rng = np.random.default_rng(seed + 1)
    both = rng.random((H, W)) < 0.06
    if ref_nd is not None:
        ref = synth.plant_nodata(ref, ref_nd, both | (rng.random((H, W)) < 0.01), extra_hits=150, seed=seed + 2)
        ref[0, rng.integers(0, H, 40), rng.integers(0, W, 40)] = ref_nd
    if tst_nd is not None:
        dec = synth.plant_nodata(dec, tst_nd, both | (rng.random((H, W)) < 0.01), extra_hits=150, seed=seed + 3)
        dec[0, rng.integers(0, H, 40), rng.integers(0, W, 40)] = tst_nd
    return _bip(ref), _bip(dec)


@pytest.mark.parametrize("variant", [0, 12, 23])
@pytest.mark.parametrize("case", ["i16_both", "u16_ref_only", "u16_tst_only_mask", "i16_tail_planes", "u16_mixed_planes"])
def test_validity_folding_one_pass_kernel_equals_validity_plus_masked_kernel(case, variant):
    """dm_fused_bip_scan (validity computed by the pixel warps inside the one-pass kernel: ONE read of the pair)
    against dm_validity + dm_fused_bip (two reads): the validity plane, the three counts, every integer partial, the
    SAM sum and the ERR8 planes / histograms must be IDENTICAL -- same rule (run_codec.py:249-263, quicklooks.py:35-45,
    run_codec.py:314-319), same arithmetic, same reduction order.  1 184+ tiles on 148 CTAs, so every stage of the
    ring is reused behind the mask barrier; both builds of the kernel; a partial last tile; a caller mask."""
    import torch
    from image_compression_analysis_b200 import synth
    from image_compression_analysis_b200._lib import lib
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate, to_device
    dtype, H, W, rnd, tnd, use_valid, planes = {
        "i16_both": ("int16", 296, 256, -32768, -32768, False, False),
        "u16_ref_only": ("uint16", 296, 256, 0, None, False, False),
        "u16_tst_only_mask": ("uint16", 300, 256, None, 65535, True, False),
        "i16_tail_planes": ("int16", 211, 173, -32768, -32768, True, True),      # 36 503 pixels = 570 tiles + 23
        "u16_mixed_planes": ("uint16", 296, 256, 0, 7, False, True),
    }[case]
    r, d = _scan_pair(dtype, H, W, rnd, tnd, seed=31)
    if tnd == 7:
        d[d == 0] = 1                       # 0 is only the ORIGINAL's nodata here
    pair = DevicePair.from_arrays(r, d, "bip", rnd, tnd)
    valid = to_device(synth.random_valid_mask(9, H, W, 0.25).reshape(-1)) if use_valid else None
    want = Want(stats=True, sam=True, err8_caps=(255, 32) if planes else (None, None), errmax=planes)
    L = lib()
    assert L.dm_fused_bip_variant(variant) == 0
    try:
        one = evaluate(pair, want, valid)
        two = evaluate(pair, Want(**{**want.__dict__, "fused_scan": False}), valid)
        torch.cuda.synchronize()
    finally:
        L.dm_fused_bip_variant(0)
    assert one.used_mask and two.used_mask
    a, b = one.to_host(), two.to_host()
    assert np.array_equal(one.planes["valid"].cpu().numpy(), two.planes["valid"].cpu().numpy())
    assert np.array_equal(a.counts, b.counts) and 0 < int(a.counts[0]) < int(a.counts[2]) <= H * W
    assert np.array_equal(a.isum, b.isum)
    assert np.array_equal(a.imax, b.imax)
    assert np.array_equal(a.fsum, b.fsum)                       # SAM: same pixels, same order, same bits
    for k in ("errmax", "err8_g", "err8_z") if planes else ():
        assert np.array_equal(one.planes[k].cpu().numpy(), two.planes[k].cpu().numpy()), k


def test_validity_folding_kernel_vs_oracle_and_all_false_rule():
    """The one-read route end to end through the public mirror (compute_metrics_arrays takes it by itself for a
    180-band BIP pair with nodata) against the ORACLE, stats-only (the pixel warps then only scan) and with SAM; and
    run_codec.py:264: when NO pixel is valid the statistics are those of every pixel (counts come back 0, the host
    reruns unmasked)."""
    import image_compression_analysis_b200 as dm
    from oracle import distortion_oracle as orc
    r, d = _scan_pair("int16", 148, 128, -32768, -32768, seed=41)
    rb, db = np.ascontiguousarray(np.moveaxis(r, -1, 0)), np.ascontiguousarray(np.moveaxis(d, -1, 0))
    kw = dict(ref_nodata=-32768, tst_nodata=-32768)
    want = orc.compute_metrics(rb, db, None, extras=False, **kw)
    _check(dm.compute_metrics_arrays(r, d, layout="bip", **kw), want)
    want.update(orc.compute_sam_sid_lmse_caseB(rb, db, None, **kw))
    _check(dm.all_metrics_arrays(r, d, layout="bip", case_b=True, extras=False, **kw), want)
    # every pixel carries one nodata sample somewhere: METRICS selects nothing -> everything counts
    r2 = r.copy()
    rng = np.random.default_rng(5)
    r2.reshape(-1, 180)[np.arange(148 * 128), rng.integers(1, 180, 148 * 128)] = -32768
    rb2 = np.ascontiguousarray(np.moveaxis(r2, -1, 0))
    want = orc.compute_metrics(rb2, db, None, extras=False, **kw)
    got = dm.compute_metrics_arrays(r2, d, layout="bip", **kw)
    _check(got, want)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_gaussian_ssim_kernels_vs_opencv_fixtures(variant):
    """Every build of the Gaussian-SSIM kernel against the OpenCV-computed fixtures (tests/golden/ssimw_cv2.npz,
    oracle/make_golden_ssim_cv2.py): a filter implementation that shares nothing with scipy or this repository."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200._lib import lib
    z = np.load(ROOT / "tests" / "golden" / "ssimw_cv2.npz")
    L_ = lib()
    if L_.dm_ssim_variant(variant) != 0:
        pytest.skip(f"dm_ssim_variant({variant}) not offered by this build")
    try:
        for n in sorted({k.split("__")[0] for k in z.files}):
            a, b, L, want = z[n + "__a"], z[n + "__b"], float(z[n + "__L"]), float(z[n + "__ssim"])
            got = dm.ssim_gaussian_arrays(a[None], b[None], L)["ssimw_b1"]
            assert _close(got, want), (n, got, want)
    finally:
        L_.dm_ssim_variant(0)


def test_sid_gain_errors_dark_spectra_and_mixtures_vs_oracle():
    """SID on error patterns the synthetic codec noise does not produce: pure gain errors of 1 % and 20 % (the SID of
    such a pair is rounding noise, ~1e-8, under per-pixel sums of 1e-2: any formulation that lets sum d^2/s meet
    (sum d)^2 cancels here -- the kernel's one-exact-numerator form does not; an r02 experiment that did, 5 FP64
    operations per sample instead of 18, was no faster because the kernel is issue bound, and was dropped), dark noisy
    spectra (most samples on the log() list), an image that mixes them pixel by pixel, identical cubes (0 up to the eps terms: ~1e-33)."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200._lib import check, lib
    from oracle import distortion_oracle as orc
    rng = np.random.default_rng(77)
    B, H, W = 180, 40, 64
    base = (rng.integers(0, 2500, (B, H, W)) * 4).astype(np.int64)
    noise = base + rng.integers(-3, 4, base.shape)
    gain1 = np.rint(base * 0.99).astype(np.int64)
    gain20 = np.rint(base * 1.2).astype(np.int64)
    dark = rng.integers(0, 30, (B, H, W)).astype(np.int64)
    dark_n = dark + rng.integers(-3, 4, dark.shape)
    mixed_r = base.copy()
    mixed_d = noise.copy()
    odd = (np.arange(H * W).reshape(H, W) % 3) == 1
    mixed_d[:, odd] = gain1[:, odd]
    mixed_r[:, 5:9, :] = dark[:, 5:9, :]; mixed_d[:, 5:9, :] = dark_n[:, 5:9, :]
    cases = {"noise": (base, noise, 1e-9), "gain 0.99": (base, gain1, REL), "gain 1.2": (base, gain20, REL),
             "dark": (dark, dark_n, 1e-9), "mixed": (mixed_r, mixed_d, 1e-8), "identical": (base, base, 0.0)}
    for lanes in (16, 8, 32):
        check(lib().dm_spectral_lanes_per_pixel(lanes))
        try:
            for name, (r, d, rel) in cases.items():
                r16, d16 = np.clip(r, 0, 65535).astype(np.uint16), np.clip(d, 0, 65535).astype(np.uint16)
                want = orc.compute_sam_sid_lmse_caseB(r16, d16)["sid"]
                got = dm.compute_sam_sid_lmse_caseB_arrays(_bip(r16), _bip(d16), layout="bip")["sid"]
                if name == "identical":
                    assert want == 0.0 and abs(got) <= 1e-30, got      # the eps terms of the exact numerator leave ~1e-33
                else:
                    assert abs(got - want) <= rel * abs(want), (lanes, name, got, want, abs(got - want) / abs(want))
        finally:
            lib().dm_spectral_lanes_per_pixel(0)


@pytest.mark.parametrize("side", [False, True])
def test_prepared_case_a_all_vs_oracle(side):
    """engine.PreparedCaseAAll (what bench.py's configs C3 / C4 launch): statistics + both ERR8 planes in one pass,
    256-bin histograms and the Gaussian SSIM of a BSQ image from prepared arguments, with and without the side
    stream that runs the two HBM-bound passes in the SSIM kernel's shadow -- against the oracle, launched twice into
    the same vector (accumulation) and read back after a plain stream synchronisation."""
    import torch
    from image_compression_analysis_b200 import finish, synth, _lib
    from image_compression_analysis_b200.engine import DevicePair, Partials, PreparedCaseAAll
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_a_pair(seed=12, bands=4, height=300, width=408)
    pair = DevicePair.from_arrays(ref, dec, "bsq")
    P = Partials.allocate(4, 256, pair.ref.device, "uint16")
    st = torch.cuda.Stream() if side else None
    prep = PreparedCaseAAll(pair, pair, (0, 300), P, 4095.0, side_stream=st)
    prep.launch()
    torch.cuda.current_stream().synchronize()
    h = P.to_host()
    got = finish.finish_compute_metrics(_lib.DM_U16, h.sums, h.maxs)
    _check(got, orc.compute_metrics(ref, dec, extras=False))
    want = orc.compute_metrics(ref, dec, hist_bins=256)
    for b in range(4):
        assert np.array_equal(h.hist[b], want[f"hist_b{b+1}"]), b
    sw = finish.finish_ssim_gauss(h.ssimw_sum, h.ssimw_cnt)
    for k, w in orc.ssim_gaussian(ref, dec, 4095.0).items():
        assert _close(sw[k], w), (k, sw[k], w)
    e = orc.error_max8(ref, dec, 255, 32)
    assert np.array_equal(prep.planes["err8_g"].cpu().numpy().reshape(300, 408), e["err8_g"])
    assert np.array_equal(prep.planes["err8_z"].cpu().numpy().reshape(300, 408), e["err8_z"])
    prep.launch()                                   # accumulates: every integer doubles
    torch.cuda.current_stream().synchronize()
    h2 = P.to_host()
    assert np.array_equal(h2.sums, 2 * h.sums) and np.array_equal(h2.hist, 2 * h.hist)


def test_validity_folding_kernel_is_repeatable_under_load():
    """The in-kernel scan hands tiles from the pixel warps to the band warps through a third mbarrier per stage and
    the band warps wait on that barrier alone: 40 launches on a full-size EnMAP-like int16 + nodata cube (16 384
    tiles each, scattered single-band nodata hits so that both sweeps of the scan run) must give the same plane, counts
    and integer partials every time, and the same as dm_validity + dm_fused_bip."""
    import torch
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    B, H, W = 180, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
    tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(-32768, 32767)
    bad = torch.rand((H, W), device="cuda", generator=g) < 0.03
    ref[bad] = -32768
    tst[bad] = -32768
    hits = torch.randint(0, H * W * B, (5000,), device="cuda", generator=g)
    ref.view(-1)[hits[:2500]] = -32768
    tst.view(-1)[hits[2500:]] = -32768
    pair = DevicePair(ref, tst, "int16", "bip", B, H, W, -32768, -32768)
    two = evaluate(pair, Want(stats=True, sam=True, fused_scan=False))
    want_plane = two.planes["valid"].clone()
    t = two.to_host()
    for i in range(40):
        one = evaluate(pair, Want(stats=True, sam=True))
        assert torch.equal(one.planes["valid"], want_plane), i
        h = one.to_host()
        assert np.array_equal(h.counts, t.counts) and np.array_equal(h.isum, t.isum) and np.array_equal(h.imax, t.imax) \
            and np.array_equal(h.fsum, t.fsum), i
    assert 0 < int(t.counts[0]) < int(t.counts[2]) < H * W
