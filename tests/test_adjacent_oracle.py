"""Oracle for the rows either side of the path: oracle/adjacent_oracle.py against the
fixtures the UNMODIFIED reference produced (tests/golden/adj_*.npz, oracle/make_golden_adjacent.py), and
the host-side pieces of the product (percentiles from a histogram, stretch tables) against numpy."""
from pathlib import Path

import numpy as np
import pytest

from oracle import adjacent_oracle as adj

GOLDEN = Path(__file__).resolve().parent / "golden"


def g(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def _nodata(rec):
    v = float(rec["nodata"][0])
    return None if np.isnan(v) else v


def _valid_mask(cube, nodata):
    """quicklooks.py:35-45 on an array: dataset mask (any band != nodata) AND band 1 != nodata."""
    if nodata is None:
        return np.ones(cube.shape[1:], bool)
    return np.any(cube != nodata, axis=0) & (cube[0] != nodata)


@pytest.mark.parametrize("name", ["adj_rgb_caseA", "adj_rgb_i16_nodata", "adj_rgb_flat"])
def test_rgb_oracle_matches_reference(name):
    r = g(name)
    nd = _nodata(r)
    sel = [int(i) - 1 for i in r["order"]]
    params = adj.stretch_params(r["cube"][sel], _valid_mask(r["cube"], nd), tuple(r["pct"]))
    assert np.array_equal(np.array(params), r["params"])
    assert np.array_equal(adj.rgb_8bit(r["cube"][sel], params), r["rgb"])


@pytest.mark.parametrize("name", ["adj_rgb_caseA", "adj_rgb_i16_nodata", "adj_rgb_flat"])
def test_percentiles_from_hist_and_lut_match_reference(name):
    """The product's host finish: the same parameters from value histograms, the same pixels from tables."""
    from image_compression_analysis_b200 import finish
    r = g(name)
    nd = _nodata(r)
    cube = r["cube"]
    first = -32768 if cube.dtype == np.int16 else 0
    ok = _valid_mask(cube, nd)
    for ch, band in enumerate(int(i) - 1 for i in r["order"]):
        h = np.bincount(cube[band][ok].astype(np.int64) - first, minlength=65536)
        got = finish.percentiles_from_hist(h, first, tuple(r["pct"]))
        lo, hi = got
        if hi <= lo:
            hi = lo + 1.0
        assert (lo, hi) == tuple(r["params"][ch])
        lut = finish.stretch8_lut(lo, hi, cube.dtype)
        assert np.array_equal(lut[cube[band].astype(np.int64) - first], r["rgb"][ch])


def test_percentiles_from_hist_random():
    from image_compression_analysis_b200 import finish
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 3000))
        lo = int(rng.integers(0, 65000))
        v = rng.integers(lo, min(lo + int(rng.integers(1, 500)), 65535) + 1, n).astype(np.uint16)
        pct = [(2, 98), (0, 100), (50.0, 50.0), (0.5, 99.5)][trial % 4]
        want = np.percentile(v.astype(np.float32), pct)
        got = finish.percentiles_from_hist(np.bincount(v, minlength=65536), 0, pct)
        assert (float(want[0]), float(want[1])) == got
    assert finish.percentiles_from_hist(np.zeros(65536, np.int64), 0, (2, 98)) is None


@pytest.mark.parametrize("name", ["adj_trunc_i16_k2", "adj_trunc_u16_k3", "adj_trunc_u16_nd"])
def test_truncation_oracle(name):
    r = g(name)
    assert np.array_equal(adj.truncated_copy(r["cube"], int(r["k"][0]), _nodata(r)), r["out"])


def test_to_12in16_oracle():
    r = g("adj_to12in16")
    assert np.array_equal(adj.to_12in16(r["cube"]), r["out"])


@pytest.mark.parametrize("name", ["adj_scene_i16_k2", "adj_scene_i16_big", "adj_scene_u16_k3"])
def test_scene_error_oracle(name):
    r = g(name)
    for mode in ("mean", "rms", "count3", "max", "p95"):
        for scale in ("fixed", "auto"):
            img, _ = adj.scene_error_map(r["ref"], r["cmp"], r.get("mask"), scale, int(r["k_bits"][0]), mode)
            assert np.array_equal(img, r[f"img_{mode}_{scale}"]), (mode, scale)


def test_diff1_oracle():
    r = g("adj_diff1_ccsds")
    assert np.array_equal(adj.diff1_bsq_signed(r["s"]), r["d_s"])
    assert np.array_equal(adj.diff1_bsq_unsigned(r["u"]), r["d_u"])
    assert np.array_equal(adj.int1_bsq_signed(r["d_s"]), r["i_s"]) and np.array_equal(r["i_s"], r["s"])
    assert np.array_equal(adj.int1_bsq_unsigned(r["d_u"]), r["i_u"]) and np.array_equal(r["i_u"], r["u"])
    j = g("adj_diff1_jpegls")
    for dt in ("uint16", "int16", "uint8"):
        assert np.array_equal(adj.diff1_cube_forward(j[f"x_{dt}"], dt), j[f"fwd_{dt}"])
        assert np.array_equal(adj.diff1_cube_inverse(j[f"fwd_{dt}"], dt), j[f"inv_{dt}"])


def test_interleave_oracle():
    r = g("adj_interleave")
    t = r["tile"]
    for mode in ("bsq", "bil", "bip"):
        assert np.array_equal(adj.to_interleave(t, mode), r[f"raw_{mode}"])
        assert np.array_equal(adj.from_interleave(r[f"raw_{mode}"], mode, *t.shape), t)


def test_oracle_vs_live_reference_random():
    """Where the reference tree is mounted (the build container), pin the restatement on fresh random cubes."""
    from oracle import reference_loader as rl
    if not rl.available():
        pytest.skip("reference tree not mounted")
    cw, jw, mb = rl.ccsds121_wrap(), rl.jpegls_wrap(), rl.make_baseline_B()
    rng = np.random.default_rng(77)
    for _ in range(5):
        B, H, W = int(rng.integers(1, 9)), int(rng.integers(1, 40)), int(rng.integers(1, 40))
        s = rng.integers(-32768, 32768, (B, H, W)).astype(np.int16)
        u = s.view(np.uint16)
        assert np.array_equal(adj.diff1_bsq_signed(s), cw._diff1_bsq_signed(s))
        assert np.array_equal(adj.diff1_bsq_unsigned(u), cw._diff1_bsq_unsigned(u))
        k = int(rng.integers(0, 6))
        assert np.array_equal(adj.trunc_uint16(u, k), mb.trunc_uint16(u, k))
        if B > 1:
            assert np.array_equal(adj.diff1_forward(s[1], s[0], "int16"), jw._diff1_forward(s[1], s[0], "int16"))
            assert np.array_equal(adj.diff1_inverse(s[1], s[0], "int16"), jw._diff1_inverse(s[1], s[0], "int16"))


def test_scene_error_oracle_vs_live_reference_random(tmp_path):
    """make_scene_error_map of the mounted reference on fresh random cubes, all modes and both scales."""
    from oracle import rasterio_stub, reference_loader as rl
    if not rl.available():
        pytest.skip("reference tree not mounted")
    from PIL import Image
    mb = rl.make_baseline_B()
    rng = np.random.default_rng(91)
    for trial, (dt, kb) in enumerate((("uint16", 2), ("int16", 3), ("uint8", 1))):
        info = np.iinfo(dt)
        B, H, W = int(rng.integers(1, 7)), int(rng.integers(3, 30)), int(rng.integers(3, 30))
        ref = rng.integers(max(info.min, -3000), min(info.max, 3000) + 1, (B, H, W)).astype(dt)
        cmp_ = np.clip(ref.astype(np.int64) + rng.integers(-(1 << kb), (1 << kb) + 1, ref.shape), info.min, info.max).astype(dt)
        mask = rng.random((H, W)) > 0.2
        rasterio_stub.clear()
        rasterio_stub.register("/mem/r.tif", ref)
        rasterio_stub.register("/mem/c.tif", cmp_)
        mp = tmp_path / f"m{trial}.tif"
        mp.write_bytes(b"x")
        rasterio_stub.register(mp, mask.astype(np.uint8))
        for mode in ("mean", "rms", "count3", "max", "p95"):
            for scale in ("fixed", "auto"):
                png = tmp_path / f"{trial}_{mode}_{scale}.png"
                mb.make_scene_error_map("/mem/r.tif", "/mem/c.tif", mp, scale, kb, png, err_mode=mode)
                want = np.array(Image.open(png))
                got, _ = adj.scene_error_map(ref, cmp_, mask, scale, kb, mode)
                assert np.array_equal(got, want), (dt, mode, scale)
