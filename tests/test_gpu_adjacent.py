"""GPU parity for the rows either side of the path: the kernels of csrc/adjacent.cu,
through the Python mirrors (quicklooks / baseline / transforms) and the C ABI, against (a) the fixtures the
unmodified reference produced (tests/golden/adj_*.npz) and (b) the numpy oracle on seeded inputs, plus
size-independent properties at full size.  Everything here is integer / float32-chain work: bit-exact."""
from pathlib import Path

import numpy as np
import pytest

from oracle import adjacent_oracle as adj

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"


def g(name):
    z = np.load(GOLDEN / f"{name}.npz")
    return {k: z[k] for k in z.files}


def _nodata(rec):
    v = float(rec["nodata"][0])
    return None if np.isnan(v) else v


def _bip(c):
    return np.ascontiguousarray(np.moveaxis(c, 0, -1))


# ---- RGB quicklook ------------------------------------------------------------------------------------
@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("name", ["adj_rgb_caseA", "adj_rgb_i16_nodata", "adj_rgb_flat"])
def test_rgb_quicklook_golden(name, layout):
    from image_compression_analysis_b200 import quicklooks as ql
    r = g(name)
    cube = r["cube"] if layout == "bsq" else _bip(r["cube"])
    order, pct = [int(i) for i in r["order"]], tuple(r["pct"])
    params = ql.stretch_params_arrays(cube, _nodata(r), order, pct, layout)
    assert np.array_equal(np.array(params), r["params"])
    assert np.array_equal(ql.rgb_8bit_arrays(cube, params, order, layout), r["rgb"])


def test_band_hist_random_and_masked():
    import torch
    from image_compression_analysis_b200 import adjacent
    from image_compression_analysis_b200.engine import to_device
    rng = np.random.default_rng(3)
    for dt, lo, hi in (("uint16", 0, 65536), ("int16", -32768, 32768), ("uint8", 0, 256)):
        B, H, W = 5, 301, 263                       # > one flush period, odd sizes
        cube = rng.integers(lo, hi, (B, H, W)).astype(dt)
        cube[2] = (cube[2].astype(np.int64) // 97 * 97).astype(dt)      # few distinct values: counters run high
        plane = (rng.random(H * W) < 0.7).astype(np.uint8) * 2
        for layout, arr in (("bsq", cube), ("bip", _bip(cube))):
            for pl in (None, plane):
                h = adjacent.band_hist(to_device(arr), dt, layout, B, H, W, [4, 2, 0],
                                       None if pl is None else torch.from_numpy(pl).cuda(), 2).cpu().numpy()
                for i, b in enumerate((4, 2, 0)):
                    v = cube[b].reshape(-1)
                    if pl is not None:
                        v = v[pl != 0]
                    want = np.bincount(v.astype(np.int64) - lo, minlength=65536)
                    assert np.array_equal(h[i], want), (dt, layout, b)


def test_band_hist_full_scene_counts():
    """10980 x 10980 band (BASELINE config 4 geometry): every pixel lands in exactly one bin; one value repeated
    across a whole flush period drives a packed 16-bit counter to its bound."""
    import torch
    from image_compression_analysis_b200 import adjacent
    H = W = 10980
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.randint(0, 4096, (1, H, W), device="cuda", dtype=torch.int16, generator=gen) * 16)
    x[0, :200] = 1234                                    # 2.2 M identical samples
    h = adjacent.band_hist(x, "uint16", "bsq", 1, H, W, [0])
    assert int(h.sum()) == H * W
    want = torch.bincount(x.reshape(-1).to(torch.int64) & 0xffff, minlength=65536)
    assert torch.equal(h[0], want)


# ---- baseline builders ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["adj_trunc_i16_k2", "adj_trunc_u16_k3", "adj_trunc_u16_nd"])
def test_truncated_copy_golden(name):
    from image_compression_analysis_b200 import baseline
    r = g(name)
    assert np.array_equal(baseline.truncated_copy_arrays(r["cube"], int(r["k"][0]), _nodata(r)), r["out"])


def test_to_12in16_golden_and_unaligned():
    from image_compression_analysis_b200 import adjacent, baseline
    from image_compression_analysis_b200.engine import to_device
    r = g("adj_to12in16")
    assert np.array_equal(baseline.to_12in16_arrays(r["cube"]), r["out"])
    x = np.arange(65536, dtype=np.uint16)                # every value, and a misaligned view of it
    assert np.array_equal(baseline.to_12in16_arrays(x), adj.to_12in16(x))
    t = to_device(np.concatenate([x, x]))[3:65536 + 40]
    got = adjacent.requantize(t, "uint16", "round", 4).cpu().numpy().view(np.uint16)
    assert np.array_equal(got, adj.to_12in16(np.concatenate([x, x])[3:65536 + 40]))
    for k in range(0, 9):
        assert np.array_equal(baseline.trunc_uint16(x, k), adj.trunc_uint16(x, k))


@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("name", ["adj_scene_i16_k2", "adj_scene_i16_big", "adj_scene_u16_k3"])
def test_scene_error_map_golden(name, layout):
    from image_compression_analysis_b200 import baseline
    r = g(name)
    ref, cmp_ = (r["ref"], r["cmp"]) if layout == "bsq" else (_bip(r["ref"]), _bip(r["cmp"]))
    for mode in ("mean", "rms", "count3", "max", "p95"):
        for scale in ("fixed", "auto"):
            img, _ = baseline.scene_error_map_arrays(ref, cmp_, r.get("mask"), scale, int(r["k_bits"][0]), mode, layout)
            assert np.array_equal(img, r[f"img_{mode}_{scale}"]), (mode, scale)


def test_scene_error_map_random_vs_oracle():
    from image_compression_analysis_b200 import baseline
    rng = np.random.default_rng(9)
    for B, H, W, kb in ((180, 17, 23, 2), (13, 65, 31, 4), (1, 5, 7, 1)):
        ref = rng.integers(0, 16000, (B, H, W)).astype(np.uint16)
        cmp_ = np.clip(ref.astype(np.int64) + rng.integers(-(1 << kb), (1 << kb) + 1, ref.shape), 0, 65535).astype(np.uint16)
        mask = rng.random((H, W)) > 0.3
        for mode in ("mean", "rms", "count3", "max", "p95"):
            want, emax = adj.scene_error_map(ref, cmp_, mask, "auto", kb, mode)
            for layout, a, c in (("bsq", ref, cmp_), ("bip", _bip(ref), _bip(cmp_))):
                got, e2 = baseline.scene_error_map_arrays(a, c, mask, "auto", kb, mode, layout)
                assert e2 == emax and np.array_equal(got, want), (B, mode, layout)


# ---- codec wrappers' transforms ------------------------------------------------------------------------
def test_diff1_golden():
    from image_compression_analysis_b200 import transforms as tf
    r = g("adj_diff1_ccsds")
    assert np.array_equal(tf._diff1_bsq_signed(r["s"]), r["d_s"])
    assert np.array_equal(tf._diff1_bsq_unsigned(r["u"]), r["d_u"])
    assert np.array_equal(tf._int1_bsq_signed(r["d_s"]), r["i_s"])
    assert np.array_equal(tf._int1_bsq_unsigned(r["d_u"]), r["i_u"])
    j = g("adj_diff1_jpegls")
    for dt in ("uint16", "int16", "uint8"):
        x, fwd, inv = j[f"x_{dt}"], j[f"fwd_{dt}"], j[f"inv_{dt}"]
        assert np.array_equal(tf.diff1_cube_forward(x, dt), fwd)
        assert np.array_equal(tf.diff1_cube_inverse(fwd, dt), inv)
        assert np.array_equal(tf._diff1_forward(x[2], x[1], dt), fwd[2])
        assert np.array_equal(tf._diff1_inverse(fwd[2], inv[1], dt), inv[2])
        first = x[0]
        assert tf._diff1_forward(first, None, dt) is first


def test_diff1_aligned_vector_path_vs_oracle():
    from image_compression_analysis_b200 import transforms as tf
    rng = np.random.default_rng(21)
    s = rng.integers(-32768, 32768, (11, 64, 128)).astype(np.int16)      # npix % 8 == 0: 16-byte vectors
    assert np.array_equal(tf._diff1_bsq_signed(s), adj.diff1_bsq_signed(s))
    assert np.array_equal(tf._int1_bsq_signed(adj.diff1_bsq_signed(s)), s)
    assert np.array_equal(tf.diff1_cube_forward(s, "int16"), adj.diff1_cube_forward(s, "int16"))
    assert np.array_equal(tf.diff1_cube_inverse(s, "int16"), adj.diff1_cube_inverse(s, "int16"))
    u8 = rng.integers(0, 256, (6, 32, 64)).astype(np.uint8)
    assert np.array_equal(tf.diff1_cube_forward(u8, "uint8"), adj.diff1_cube_forward(u8, "uint8"))
    assert np.array_equal(tf.diff1_cube_inverse(u8, "uint8"), adj.diff1_cube_inverse(u8, "uint8"))


def test_diff1_roundtrip_case_b_cube():
    """Full Case-B geometry (180 x 1024 x 1024): inverse(forward(x)) == x, and the forward residual of a
    constant-slope spectrum is that slope."""
    import torch
    from image_compression_analysis_b200 import adjacent
    gen = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randint(-32768, 32768, (180, 1024, 1024), device="cuda", dtype=torch.int16, generator=gen)
    d = adjacent.diff1(x, "int16", inverse=False)
    assert torch.equal(adjacent.diff1(d, "int16", inverse=True), x)
    assert torch.equal(d[1:], x[1:] - x[:-1])            # torch int16 arithmetic wraps modulo 2^16 too
    assert torch.equal(d[0], x[0])


def test_interleave_golden_and_roundtrip(tmp_path):
    from image_compression_analysis_b200 import transforms as tf
    r = g("adj_interleave")
    t = r["tile"]
    for mode in ("bsq", "bil", "bip"):
        p = tmp_path / f"{mode}.raw"
        tf._write_raw_interleaved(t, mode, p, np.dtype("<u2"))
        assert np.array_equal(np.fromfile(p, dtype="<u2"), r[f"raw_{mode}"])
        assert np.array_equal(tf._read_raw_interleaved(p, mode, np.dtype("<u2"), *t.shape), t)
    with pytest.raises(ValueError):
        tf._write_raw_interleaved(t, "xyz", tmp_path / "x.raw", np.dtype("<u2"))
    rng = np.random.default_rng(4)
    for dt in (np.uint16, np.uint8):
        c = rng.integers(0, 256, (7, 70, 131)).astype(dt)
        B, H, W = c.shape
        forms = {"bsq": c, "bil": np.ascontiguousarray(np.moveaxis(c, 0, 1)), "bip": _bip(c)}
        for a in forms:
            for b in forms:
                assert np.array_equal(tf.interleave_arrays(forms[a], a, b, B, H, W), forms[b]), (a, b)


def test_interleave_case_b_cube_roundtrip():
    import torch
    from image_compression_analysis_b200 import adjacent
    gen = torch.Generator(device="cuda").manual_seed(6)
    x = torch.randint(-32768, 32768, (1024, 1024, 180), device="cuda", dtype=torch.int16, generator=gen)
    bsq = adjacent.interleave(x, "bip", "bsq", 180, 1024, 1024)
    assert torch.equal(bsq, x.permute(2, 0, 1))
    bil = adjacent.interleave(bsq, "bsq", "bil", 180, 1024, 1024)
    assert torch.equal(bil, x.permute(0, 2, 1))
    assert torch.equal(adjacent.interleave(bil, "bil", "bip", 180, 1024, 1024), x)


def test_c_abi_argument_errors():
    import ctypes as C
    from image_compression_analysis_b200 import _lib
    L = _lib.lib()
    assert L.dm_requantize(None, None, _lib.DM_U16, 8, 0, 2, 0, 0, None) == _lib.DM_EARG
    assert L.dm_diff1(None, None, _lib.DM_I16, 0, 0, 4, 16, 16, None) == _lib.DM_EARG
    assert L.dm_interleave(None, None, 2, 0, 1, 4, 4, 4, None) == _lib.DM_EARG
    assert b"null" in L.dm_last_error()


def test_path_level_functions_on_real_files(tmp_path, monkeypatch):
    """The reference-named entry points on REAL GeoTIFFs (built-in reader/writer, no rasterio):
    stretch_params_from_baseline / write_rgb_8bit (quicklooks.py:51-109), write_truncated_copy /
    make_scene_error_map (make_baseline_B.py:284-419), to_12in16 (make_baseline_A.py:137-170)."""
    import sys
    monkeypatch.delitem(sys.modules, "rasterio", raising=False)
    from PIL import Image
    from image_compression_analysis_b200 import baseline, geotiff, ingest, quicklooks as ql
    ingest.clear_cache()
    r = g("adj_rgb_i16_nodata")
    cube, nd = r["cube"], _nodata(r)
    B, H, W = cube.shape
    src = tmp_path / "scene.tif"
    with geotiff.open(src, "w", dtype="int16", count=B, width=W, height=H, tiled=True, blockxsize=64, blockysize=64,
                      nodata=nd) as dst:
        dst.write(cube)
    order, pct = [int(i) for i in r["order"]], tuple(r["pct"])
    params = ql.stretch_params_from_baseline(src, rgb_order=order, pct=pct)
    assert np.array_equal(np.array(params), r["params"])
    ql.write_rgb_8bit(src, tmp_path / "ql" / "rgb.tif", params, rgb_order=order)
    with geotiff.open(tmp_path / "ql" / "rgb.tif") as d:
        assert d.count == 3 and d.dtypes[0] == "uint8" and d.nodata is None
        assert np.array_equal(d.read(), r["rgb"])
    # 14-in-16 copy with nodata kept, then the scene error map against the original
    t = g("adj_trunc_i16_k2")
    with geotiff.open(tmp_path / "in16.tif", "w", dtype="int16", count=t["cube"].shape[0], width=t["cube"].shape[2],
                      height=t["cube"].shape[1], nodata=_nodata(t)) as dst:
        dst.write(t["cube"])
    baseline.write_truncated_copy(tmp_path / "in16.tif", tmp_path / "out14.tif", int(t["k"][0]))
    with geotiff.open(tmp_path / "out14.tif") as d:
        assert np.array_equal(d.read(), t["out"]) and d.nodata == _nodata(t)
    sc = g("adj_scene_i16_k2")
    for name, arr in (("ref.tif", sc["ref"]), ("cmp.tif", sc["cmp"])):
        with geotiff.open(tmp_path / name, "w", dtype="int16", count=arr.shape[0], width=arr.shape[2], height=arr.shape[1]) as dst:
            dst.write(arr)
    with geotiff.open(tmp_path / "mask.tif", "w", dtype="uint8", count=1, width=sc["mask"].shape[1], height=sc["mask"].shape[0]) as dst:
        dst.write(sc["mask"].astype(np.uint8)[None])
    for mode in ("mean", "p95"):
        baseline.make_scene_error_map(tmp_path / "ref.tif", tmp_path / "cmp.tif", tmp_path / "mask.tif", "auto",
                                      int(sc["k_bits"][0]), tmp_path / f"{mode}.png", err_mode=mode)
        assert np.array_equal(np.array(Image.open(tmp_path / f"{mode}.png")), sc[f"img_{mode}_auto"])
    a = g("adj_to12in16")
    with geotiff.open(tmp_path / "a.tif", "w", dtype="uint16", count=4, width=a["cube"].shape[2], height=a["cube"].shape[1]) as dst:
        dst.write(a["cube"])
    baseline.to_12in16(tmp_path / "a.tif", tmp_path / "o" / "a12.tif")
    with geotiff.open(tmp_path / "o" / "a12.tif") as d:
        assert np.array_equal(d.read(), a["out"])
