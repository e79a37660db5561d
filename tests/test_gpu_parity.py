"""GPU parity: the CUDA path (through the public Python mirror and the C-ABI library) against
(a) the reference-generated golden fixtures and (b) the numpy oracle on seeded inputs.

Bars: integers bit-exact; PSNR bit-exact (host math.log10 on exact integers);
SSIM / SAM / SID / LMSE / Gaussian SSIM within 1e-6 relative (abs floor 1e-12)."""
import math

import numpy as np
import pytest

from tests import goldenio

pytestmark = pytest.mark.gpu

REL = 1e-6


def _close(g, w, rel=REL):
    return goldenio.close(g, w, rel=rel)


def _check_metrics(got, want, exact_psnr=True):
    assert set(want) <= set(got)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, (int, np.integer)):
            assert isinstance(g, int) and g == int(w), (k, g, w)
        elif k.startswith("psnr") and exact_psnr:
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)
        else:
            assert isinstance(g, float) and _close(g, w), (k, g, w)


@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("name", goldenio.names())
def test_compute_metrics_golden(name, layout):
    import image_compression_analysis_b200 as dm
    c = goldenio.load(name)
    ref, tst = c["ref"], c["tst"]
    if layout == "bip":
        ref, tst = np.ascontiguousarray(np.moveaxis(ref, 0, -1)), np.ascontiguousarray(np.moveaxis(tst, 0, -1))
    got = dm.compute_metrics_arrays(ref, tst, c["valid"], ref_nodata=c["ref_nodata"], tst_nodata=c["tst_nodata"],
                                    layout=layout)
    assert set(got) == set(c["compute_metrics"])
    _check_metrics(got, c["compute_metrics"])


@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("name", goldenio.names())
def test_sam_sid_lmse_golden(name, layout):
    import image_compression_analysis_b200 as dm
    c = goldenio.load(name)
    ref, tst = c["ref"], c["tst"]
    if layout == "bip":
        ref, tst = np.ascontiguousarray(np.moveaxis(ref, 0, -1)), np.ascontiguousarray(np.moveaxis(tst, 0, -1))
    got = dm.compute_sam_sid_lmse_caseB_arrays(ref, tst, c["valid"], ref_nodata=c["ref_nodata"],
                                               tst_nodata=c["tst_nodata"], layout=layout)
    assert set(got) == {"sam_deg", "sid", "lmse"}
    for k, w in c["sam_sid_lmse"].items():
        assert _close(got[k], w), (k, got[k], w)


@pytest.mark.parametrize("name", goldenio.names())
def test_error_max8_golden(name):
    from image_compression_analysis_b200 import quicklooks as ql
    c = goldenio.load(name)
    for n, e in enumerate(c["err8"]):
        got = ql.error_max8_arrays(c["ref"], c["tst"], e["cap_g"], e["cap_z"], a_nodata=c["ref_nodata"],
                                   b_nodata=c["tst_nodata"])
        assert np.array_equal(got["err8_g"], c["planes"][f"err8_{n}_g"])
        assert np.array_equal(got["valid"].astype(np.uint8), c["planes"][f"err8_{n}_mask"])
        assert e["name_g"] == f"recon_ERR8_0_{got['cap_g']}.tif"
        assert float(e["tags_g"]["STATISTICS_MEAN"]) == got["mean_g"]
        assert _close(got["std_g"], float(e["tags_g"]["STATISTICS_STDDEV"]), rel=1e-12)
        if e["cap_z"] is not None:
            assert np.array_equal(got["err8_z"], c["planes"][f"err8_{n}_z"])
            assert e["name_z"] == f"recon_ERR8_0_{got['cap_z']}.tif"
            assert float(e["tags_z"]["STATISTICS_MEAN"]) == got["mean_z"]


def _rand_pair(seed, dtype, B, H, W, amp, lowbits=0):
    rng = np.random.default_rng(seed)
    info = np.iinfo(dtype)
    ref = rng.integers(info.min, int(info.max) + 1, size=(B, H, W)).astype(dtype)
    if lowbits:
        ref = ((ref >> lowbits) << lowbits).astype(dtype)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-amp, amp + 1, size=ref.shape), info.min, info.max).astype(dtype)
    return ref, dec


@pytest.mark.parametrize("generic", [False, True])
@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("dtype,B,H,W,amp,masked", [
    ("uint16", 4, 257, 515, 40, False),      # odd sizes: scalar head/tail of the packed kernel
    ("uint16", 4, 256, 512, 65535, True),    # full-range errors: every 32-bit partial at its bound
    ("int16", 6, 130, 259, 3000, True),
    ("int16", 12, 64, 100, 65535, False),
    ("uint8", 3, 99, 131, 9, True),
    ("uint16", 7, 67, 93, 500, True),        # odd band count (BIP -> generic kernel)
    ("uint16", 180, 32, 48, 5, False),       # EnMAP band count
])
def test_fused_stats_vs_oracle(dtype, B, H, W, amp, masked, layout, generic):
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import DevicePair, Want, dtype_code
    from image_compression_analysis_b200.metrics import _valid_to_device, metrics_partials
    from oracle import distortion_oracle as orc
    ref, dec = _rand_pair(7, dtype, B, H, W, amp)
    valid = (np.random.default_rng(8).random((H, W)) < 0.7) if masked else None
    want = orc.compute_metrics(ref, dec, valid, hist_bins=256, extras=True)
    r, d = (ref, dec) if layout == "bsq" else (np.ascontiguousarray(np.moveaxis(ref, 0, -1)),
                                                np.ascontiguousarray(np.moveaxis(dec, 0, -1)))
    pair = DevicePair.from_arrays(r, d, layout)
    P = metrics_partials(pair, _valid_to_device(valid, H, W, "shape"), Want(stats=True, hist_bins=256, generic_stats=generic))
    h = P.to_host()
    got = finish.finish_compute_metrics(dtype_code(dtype), h.sums, h.maxs, h.hist, extras=True)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, np.ndarray):
            assert np.array_equal(g, w), k
        elif isinstance(w, (int, np.integer)):
            assert g == int(w), (k, g, w)
        elif k.startswith("psnr") or k.startswith("mae"):
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)
        else:
            assert _close(g, w, rel=1e-9), (k, g, w)


@pytest.mark.parametrize("dtype", ["uint16", "int16"])
def test_no_moments_variant_matches(dtype):
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    ref, dec = _rand_pair(21, dtype, 5, 200, 333, 65535)
    pair = DevicePair.from_arrays(ref, dec)
    a = evaluate(pair, Want(stats=True)).to_host()
    b = evaluate(pair, Want(stats=True, moments=False)).to_host()
    for k in (0, 6, 7):       # N, S|d|, SSE
        assert np.array_equal(a.sums[:, k], b.sums[:, k])
    assert np.array_equal(a.maxs, b.maxs)
    d = dec.astype(np.int64) - ref.astype(np.int64)
    assert np.array_equal(a.sums[:, 7], (d * d).reshape(5, -1).sum(1))


def test_misaligned_views():
    """Row strips of a cube whose row pitch is not a multiple of 16 bytes: the packed kernel must
    peel a scalar head per band and still agree with numpy."""
    import torch
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    ref, dec = _rand_pair(31, "uint16", 3, 41, 37, 300)
    full = DevicePair.from_arrays(ref, dec)
    for r0, r1 in [(1, 40), (3, 4), (0, 41), (7, 30)]:
        # same storage, offset by r0 rows, band stride of the full cube
        sub = DevicePair(full.ref.view(-1)[r0 * 37:], full.tst.view(-1)[r0 * 37:], "uint16", "bsq", 3, r1 - r0, 37,
                         band_stride=41 * 37)
        h = evaluate(sub, Want(stats=True)).to_host()
        a = ref[:, r0:r1].astype(np.int64)
        b = dec[:, r0:r1].astype(np.int64)
        assert np.array_equal(h.sums[:, 7], ((a - b) ** 2).reshape(3, -1).sum(1))
        assert np.array_equal(h.sums[:, 5], (a * b).reshape(3, -1).sum(1))
        assert np.array_equal(h.maxs[:, 0], np.abs(a - b).reshape(3, -1).max(1))
        assert int(h.sums[0, 0]) == (r1 - r0) * 37


def test_ssim_gaussian_vs_oracle():
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import synth
    from oracle import distortion_oracle as orc
    ref, dec = synth.case_a_pair(seed=3, bands=3, height=150, width=203, sigma=2.0)
    want = orc.ssim_gaussian(ref, dec)
    got = dm.ssim_gaussian_arrays(ref, dec)
    assert set(got) == set(want)
    for k in want:
        assert _close(got[k], want[k]), (k, got[k], want[k])
    # full-range noise (SSIM far from 1) and int16
    ref, dec = _rand_pair(5, "int16", 2, 64, 80, 20000)
    want = orc.ssim_gaussian(ref, dec, data_range=65535.0)
    got = dm.ssim_gaussian_arrays(ref, dec, data_range=65535.0)
    for k in want:
        assert _close(got[k], want[k]), (k, got[k], want[k])


def test_scalar_helpers_match_oracle():
    import image_compression_analysis_b200 as dm
    from oracle import distortion_oracle as orc
    rng = np.random.default_rng(5)
    a = rng.integers(0, 65536, size=(33, 41)).astype(np.uint16)
    b = rng.integers(0, 65536, size=(33, 41)).astype(np.uint16)
    assert dm.mse(a, b) == orc.mse(a, b)
    assert dm.psnr(a, b, 65535) == orc.psnr(a, b, 65535)
    assert dm.psnr(a, a, 65535) == float("inf")
    assert _close(dm.ssim_global(a, b, 4095), orc.ssim_global(a, b, 4095), rel=1e-12)
    cube = np.stack([a, b])
    assert dm.effective_data_range_arrays(cube) == orc.effective_data_range(cube) == 65535
    assert dm.effective_data_range_arrays((cube >> 4) << 4) == 4095
    with pytest.raises(TypeError):
        dm.mse(a.astype(np.float64), b.astype(np.float64))


def test_path_level_functions_with_rasterio_stub(tmp_path):
    """compute_metrics / compute_sam_sid_lmse_caseB / write_error_max8 with the reference's own
    signatures, reading through `rasterio` (here: the in-memory stub from oracle/)."""
    from pathlib import Path
    from oracle import rasterio_stub
    rasterio_stub.install()
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import quicklooks as ql
    c = goldenio.load("b_i16_nodata_masked")
    rasterio_stub.clear()
    rasterio_stub.register("/mem/ref.tif", c["ref"], nodata=c["ref_nodata"])
    rasterio_stub.register("/mem/tst.tif", c["tst"], nodata=c["tst_nodata"])
    got = dm.compute_metrics(Path("/mem/ref.tif"), Path("/mem/tst.tif"), valid=c["valid"])
    _check_metrics(got, c["compute_metrics"])
    got = dm.compute_sam_sid_lmse_caseB(Path("/mem/ref.tif"), Path("/mem/tst.tif"), valid=c["valid"])
    for k, w in c["sam_sid_lmse"].items():
        assert _close(got[k], w), k
    with pytest.raises(ValueError):
        dm.compute_metrics(Path("/mem/ref.tif"), Path("/mem/tst.tif"), valid=np.ones((3, 3), bool))
    e = c["err8"][0]
    og, oz = ql.write_error_max8("/mem/ref.tif", "/mem/tst.tif", "/mem/out/recon", err_max_global=e["cap_g"],
                                 err_max_zoom=e["cap_z"])
    assert Path(og).name == e["name_g"] and oz is None
    w = rasterio_stub.fetch(og)
    assert np.array_equal(w.data[0], c["planes"]["err8_0_g"])
    assert np.array_equal(w.mask.astype(np.uint8), c["planes"]["err8_0_mask"])
    assert w.tags["STATISTICS_MEAN"] == e["tags_g"]["STATISTICS_MEAN"]


def _fused_vs_separate(ref_bsq, dec_bsq, valid, ref_nodata=None, tst_nodata=None, caps=(255, 32)):
    """dm_fused_bip (one pass) against dm_fused_stats + dm_spectral (two passes) on the BIP view."""
    import torch
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    from image_compression_analysis_b200.metrics import _valid_to_device
    r = np.ascontiguousarray(np.moveaxis(ref_bsq, 0, -1))
    d = np.ascontiguousarray(np.moveaxis(dec_bsq, 0, -1))
    H, W, B = r.shape
    pair = DevicePair.from_arrays(r, d, "bip", ref_nodata, tst_nodata)
    vdev = _valid_to_device(valid, H, W, "shape")
    outs = []
    for fused in (True, False):
        P = evaluate(pair, Want(stats=True, sam=True, errmax=True, err8_caps=caps, fused=fused), vdev)
        torch.cuda.synchronize()
        outs.append((P.to_host(), {k: v.cpu().numpy() for k, v in P.planes.items()}))
    (hf, pf), (hs, ps) = outs
    assert np.array_equal(hf.isum, hs.isum)
    assert np.array_equal(hf.imax.reshape(-1, 8).max(0)[1:], hs.imax.reshape(-1, 8).max(0)[1:])   # cube-wide maxima
    assert np.array_equal(hf.maxs[:, 0], hs.maxs[:, 0])                                           # per-band max|d|
    for k in ("errmax", "err8_g", "err8_z"):
        assert np.array_equal(pf[k], ps[k]), k
    assert hf.spec[2] == hs.spec[2]
    assert goldenio.close(hf.spec[0], hs.spec[0], rel=1e-12)
    return hf


@pytest.mark.parametrize("name", ["a_gauss", "a_identical", "a_near3_masked", "a_mask_all_false", "b_u16_masked",
                                  "b_i16_nodata", "b_i16_nodata_masked", "b_zero_spectra"])
def test_fused_bip_equals_two_pass_on_goldens(name):
    from image_compression_analysis_b200 import _lib
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    c = goldenio.load(name)
    assert c["ref"].shape[0] % 4 == 0
    hf = _fused_vs_separate(c["ref"], c["tst"], c["valid"], c["ref_nodata"], c["tst_nodata"])
    # and the fused kernel really ran (a second call must not return DM_EUNSUPPORTED)
    import ctypes as C
    import torch
    r = np.ascontiguousarray(np.moveaxis(c["ref"], 0, -1))
    pair = DevicePair.from_arrays(r, r, "bip")
    sums = torch.zeros(pair.bands * 8, dtype=torch.int64, device="cuda")
    maxs = torch.zeros_like(sums)
    rc = _lib.lib().dm_fused_bip(C.byref(pair.c_pair()), None, sums.data_ptr(), maxs.data_ptr(), None, None, 0, None,
                                 None, None, 0, None, None, 0, None, None, None)
    assert rc == _lib.DM_OK, _lib.lib().dm_last_error()
    assert int(sums[0].item()) == pair.npix


@pytest.mark.parametrize("dtype,B,H,W,amp,masked", [
    ("uint16", 180, 37, 53, 5, False),       # EnMAP bands, odd pixel count: generic tail tile + single pixel
    ("uint16", 180, 64, 64, 65535, True),    # full-range errors, every partial at its bound
    ("int16", 180, 40, 41, 30000, True),
    ("uint16", 256, 16, 40, 65535, False),   # the largest band count the lo/hi partials allow
    ("int16", 4, 129, 257, 700, False),
    ("uint16", 8, 300, 301, 9, True),
])
def test_fused_bip_vs_oracle(dtype, B, H, W, amp, masked):
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import dtype_code
    from oracle import distortion_oracle as orc
    ref, dec = _rand_pair(11, dtype, B, H, W, amp)
    if amp == 65535 and dtype == "uint16":
        ref[:, :8] = 65535; dec[:, :8] = 0          # adversarial block: maximal products everywhere
    valid = (np.random.default_rng(12).random((H, W)) < 0.6) if masked else None
    hf = _fused_vs_separate(ref, dec, valid)
    got = finish.finish_compute_metrics(dtype_code(dtype), hf.sums, hf.maxs)
    want = orc.compute_metrics(ref, dec, valid, extras=False)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, int):
            assert g == w, (k, g, w)
        elif k.startswith("psnr"):
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)
        else:
            assert _close(g, w, rel=1e-9), (k, g, w)
    sam_want = orc.sam_caseB(ref, dec, valid)
    sam_got = finish.finish_spectral(float(hf.spec[0]), 0.0, float(hf.spec[2]), None, H * W)["sam_deg"]
    assert _close(sam_got, sam_want), (sam_got, sam_want)


def _fused_bsq_vs_separate(ref, dec, valid, ref_nodata=None, tst_nodata=None, caps=(255, 32)):
    """dm_fused_bsq (one pass) against dm_fused_stats + dm_spectral (two passes), BSQ cubes of <= 4 bands."""
    import ctypes as C
    import torch
    from image_compression_analysis_b200 import _lib
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    from image_compression_analysis_b200.metrics import _valid_to_device
    B, H, W = ref.shape
    pair = DevicePair.from_arrays(ref, dec, "bsq", ref_nodata, tst_nodata)
    vdev = _valid_to_device(valid, H, W, "shape")
    outs = []
    l0 = _lib.lib().dm_launch_count()
    for fused in (True, False):
        P = evaluate(pair, Want(stats=True, errmax=True, err8_caps=caps, fused=fused), vdev)
        torch.cuda.synchronize()
        outs.append((P.to_host(), {k: v.cpu().numpy() for k, v in P.planes.items()}))
    (hf, pf), (hs, ps) = outs
    assert np.array_equal(hf.isum, hs.isum)
    assert np.array_equal(hf.imax.reshape(-1, 8).max(0)[1:], hs.imax.reshape(-1, 8).max(0)[1:])   # cube-wide maxima
    assert np.array_equal(hf.maxs[:, 0], hs.maxs[:, 0])                                           # per-band max|d|
    for k in ("errmax", "err8_g", "err8_z"):
        if k in ps:
            assert np.array_equal(pf[k], ps[k]), k
    # and the one-pass kernel really ran: a direct call must not answer DM_EUNSUPPORTED
    sums = torch.zeros(B * 8, dtype=torch.int64, device="cuda")
    maxs = torch.zeros_like(sums)
    rc = _lib.lib().dm_fused_bsq(C.byref(pair.c_pair()), None, sums.data_ptr(), maxs.data_ptr(), None, None, 0, None,
                                 None, None, 0, None, None, None)
    if B == 1 or (H * W) % 8 == 0:            # bands of a packed (B,H,W) array start on 16-byte boundaries
        assert rc == _lib.DM_OK, _lib.lib().dm_last_error()
        assert int(sums[0].item()) == H * W
    else:
        assert rc == _lib.DM_EUNSUPPORTED
    return hf, pf


@pytest.mark.parametrize("name", ["a_gauss", "a_identical", "a_near3_masked", "a_mask_all_false", "single_band_odd"])
def test_fused_bsq_equals_two_pass_on_goldens(name):
    c = goldenio.load(name)
    if c["ref"].shape[0] > 4 or c["ref"].dtype == np.uint8:
        pytest.skip("dm_fused_bsq handles 16-bit cubes of up to 4 bands")
    _fused_bsq_vs_separate(c["ref"], c["tst"], c["valid"], c["ref_nodata"], c["tst_nodata"])


@pytest.mark.parametrize("dtype,B,H,W,amp,masked,caps", [
    ("uint16", 4, 256, 512, 40, False, (255, 32)),
    ("uint16", 4, 257, 515, 300, True, (255, 32)),      # odd pixel count: scalar tail; errors above both caps
    ("uint16", 4, 128, 256, 65535, False, (255, None)),  # full-range errors, every partial at its bound
    ("int16", 3, 130, 264, 3000, True, (100, 7)),
    ("int16", 1, 64, 1000, 65535, False, (255, 32)),
    ("uint16", 2, 3, 2, 9, False, (255, 32)),            # fewer than eight pixels
])
def test_fused_bsq_vs_oracle(dtype, B, H, W, amp, masked, caps):
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import dtype_code
    from oracle import distortion_oracle as orc
    ref, dec = _rand_pair(21, dtype, B, H, W, amp)
    if amp == 65535 and dtype == "uint16":
        ref[:, :8] = 65535; dec[:, :8] = 0
    valid = (np.random.default_rng(22).random((H, W)) < 0.6) if masked else None
    hf, pf = _fused_bsq_vs_separate(ref, dec, valid, caps=caps)
    got = finish.finish_compute_metrics(dtype_code(dtype), hf.sums, hf.maxs)
    want = orc.compute_metrics(ref, dec, valid, extras=False)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, int):
            assert g == w, (k, g, w)
        elif k.startswith("psnr"):
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)
        else:
            assert _close(g, w, rel=1e-9), (k, g, w)
    o = orc.error_max8(ref, dec, caps[0], caps[1])
    assert np.array_equal(pf["err8_g"].reshape(H, W), o["err8_g"])
    if caps[1] is not None:
        assert np.array_equal(pf["err8_z"].reshape(H, W), o["err8_z"])
        assert np.array_equal(np.bincount(o["err8_z"].ravel(), minlength=256), hf.hist8_z)
    assert np.array_equal(np.bincount(o["err8_g"].ravel(), minlength=256), hf.hist8_g)


@pytest.mark.parametrize("records", [1, 5])
def test_combine_partials_kernel(records):
    """dm_combine_partials (the reduction after the multi-GPU all-gather) against numpy."""
    import ctypes as C
    import torch
    from image_compression_analysis_b200 import _lib
    rng = np.random.default_rng(3)
    world, ns, nm, nf = 8, 1443, 1440, 543
    L = ns + nm + nf
    G = np.zeros((world, records, L), np.int64)
    G[:, :, :ns + nm] = rng.integers(-2**40, 2**40, size=(world, records, ns + nm))
    F = rng.standard_normal((world, records, nf)) * 1e6
    G[:, :, ns + nm:] = F.view(np.int64)
    g = torch.from_numpy(G).cuda()
    out = torch.zeros(records * L, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().dm_combine_partials(C.c_void_p(g.data_ptr()), world, records, ns, nm, nf,
                                              C.c_void_p(out.data_ptr()), None))
    o = out.cpu().numpy().reshape(records, L)
    assert np.array_equal(o[:, :ns], G[:, :, :ns].sum(0))
    assert np.array_equal(o[:, ns:ns + nm], G[:, :, ns:ns + nm].max(0))
    want = np.zeros((records, nf))
    for r in range(world):
        want = want + F[r]                      # rank order, like the kernel
    assert np.array_equal(o[:, ns + nm:].view(np.float64), want)


def test_launcher_rebinds_the_reference_module(tmp_path):
    """launcher.bind on a stand-in `run_codec` module: the reference's call sites
    (tools/run_codec.py:518-526) reach the GPU functions with the reference's signatures."""
    import types
    from pathlib import Path
    from oracle import rasterio_stub
    rasterio_stub.install()
    from image_compression_analysis_b200 import launcher
    mod = types.ModuleType("run_codec")
    launcher.bind(mod)
    c = goldenio.load("a_gauss")
    rasterio_stub.clear()
    rasterio_stub.register("/mem/src.tif", c["ref"])
    rasterio_stub.register("/mem/recon.tif", c["tst"])
    row = mod.compute_metrics(Path("/mem/src.tif"), Path("/mem/recon.tif"), valid=None)
    _check_metrics(row, c["compute_metrics"])
    extra = mod.compute_sam_sid_lmse_caseB(Path("/mem/src.tif"), Path("/mem/recon.tif"), valid=None)
    for k, w in c["sam_sid_lmse"].items():
        assert _close(extra[k], w), k


@pytest.mark.gpu
def test_real_geotiff_files_read_once_per_rep(tmp_path, monkeypatch):
    """The three drop-in calls of one rep (run_codec.py:518, :522, :526) on REAL files: built-in GeoTIFF
    reader (no rasterio), pixel-interleaved cubes uploaded as stored (BIP), each file read ONCE."""
    import sys
    monkeypatch.delitem(sys.modules, "rasterio", raising=False)      # a stub another test may have installed
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import geotiff, ingest, quicklooks as ql
    from oracle import distortion_oracle as orc
    ref, dec = _rand_pair(31, "uint16", 8, 96, 80, 6, lowbits=2)
    valid = np.random.default_rng(32).random((96, 80)) < 0.7
    for name, cube in (("src.tif", ref), ("recon.tif", dec)):
        with geotiff.open(tmp_path / name, "w", dtype="uint16", count=8, width=80, height=96, tiled=True,
                          blockxsize=64, blockysize=64, compress="NONE" if name == "src.tif" else "DEFLATE") as dst:
            dst.write(cube)
    ingest.clear_cache()
    r0 = ingest.STATS["reads"]
    src, rec = tmp_path / "src.tif", tmp_path / "recon.tif"
    og, oz = ql.write_error_max8(src, rec, tmp_path / "recon", err_max_global=255, err_max_zoom=32)
    got = dm.compute_metrics(src, rec, valid=valid)
    spec = dm.compute_sam_sid_lmse_caseB(src, rec, valid=valid)
    assert ingest.STATS["reads"] - r0 == 2, "every file must be read exactly once for the three calls"
    pair, _ = ingest.load_pair(src, rec)
    assert pair.layout == "bip"
    _check_metrics(got, orc.compute_metrics(ref, dec, valid, extras=False))
    want = orc.compute_sam_sid_lmse_caseB(ref, dec, valid)
    for k, w in want.items():
        assert _close(spec[k], w), (k, spec[k], w)
    o = orc.error_max8(ref, dec, 255, 32)
    assert og.name == "recon_ERR8_0_255.tif" and oz.name == "recon_ERR8_0_32.tif"
    with geotiff.open(og) as d:
        assert np.array_equal(d.read(1), o["err8_g"]) and d.dtypes[0] == "uint8" and d.tiled
        assert np.array_equal(d.dataset_mask(), np.full((96, 80), 255, np.uint8))
    with geotiff.open(oz) as d:
        assert np.array_equal(d.read(1), o["err8_z"])
    tags = geotiff.read_tags(og)
    assert tags["STATISTICS_MEAN"] == str(float(o["err8_g"].mean())) and tags["STATISTICS_MAXIMUM"] == "255"
    # a new decode at the same path (next rep) must not be served from the cache
    import os, time
    dec2 = dec.copy(); dec2[0, 0, 0] ^= 0x100
    with geotiff.open(rec, "w", dtype="uint16", count=8, width=80, height=96, compress="DEFLATE") as dst:
        dst.write(dec2)
    os.utime(rec, ns=(time.time_ns(), time.time_ns() + 1_000_000))
    got2 = dm.compute_metrics(src, rec)
    assert got2["max_abs_err"] == orc.compute_metrics(ref, dec2, extras=False)["max_abs_err"]
    assert ingest.STATS["reads"] - r0 == 3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,nd,with_valid,ref_only", [("int16", -32768, False, False), ("int16", -32768, True, False),
                                                          ("uint16", 0, False, True), ("uint16", 65535, True, False)])
def test_nodata_180_bands_vs_oracle(dtype, nd, with_valid, ref_only):
    """The real EnMAP configuration: 180-band BIP cubes with a nodata value (validity pre-pass through the
    TMA-staged kernel, masked one-pass kernel with the mask bytes staged next to the tile), all three
    drop-in results against the oracle.  Whole-pixel nodata, single-band nodata in either cube and
    band-1-only nodata exercise the three different mask rules."""
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import quicklooks as ql
    from oracle import distortion_oracle as orc
    B, H, W = 180, 45, 67                                   # 3015 pixels: 47 full tiles + a partial one
    rng = np.random.default_rng(41)
    ref, dec = _rand_pair(40, dtype, B, H, W, 25)
    ref[ref == nd] += 1; dec[dec == nd] += 1                # nodata only where planted below
    whole = rng.random((H, W)) < 0.08
    ref[:, whole] = nd
    if not ref_only:
        dec[:, whole] = nd
    one = rng.random((H, W)) < 0.03; ref[rng.integers(1, B), one] = nd          # one band of the original
    if not ref_only:
        two = rng.random((H, W)) < 0.03; dec[rng.integers(1, B), two] = nd      # one band of the decoded cube
    b1 = rng.random((H, W)) < 0.02; ref[0, b1] = nd                             # band 1 only (quicklook rule)
    valid = (rng.random((H, W)) < 0.8) if with_valid else None
    kw = dict(ref_nodata=nd, tst_nodata=None if ref_only else nd)
    r_bip = np.ascontiguousarray(np.moveaxis(ref, 0, -1)); d_bip = np.ascontiguousarray(np.moveaxis(dec, 0, -1))
    got = dm.compute_metrics_arrays(r_bip, d_bip, valid, layout="bip", **kw)
    _check_metrics(got, orc.compute_metrics(ref, dec, valid, extras=False, **kw))
    spec = dm.compute_sam_sid_lmse_caseB_arrays(r_bip, d_bip, valid, layout="bip", **kw)
    for k, w in orc.compute_sam_sid_lmse_caseB(ref, dec, valid, **kw).items():
        assert _close(spec[k], w), (k, spec[k], w)
    e = ql.error_max8_arrays(r_bip, d_bip, 255, 32, layout="bip", a_nodata=kw["ref_nodata"], b_nodata=kw["tst_nodata"])
    o = orc.error_max8(ref, dec, 255, 32, **kw)
    assert np.array_equal(e["err8_g"], o["err8_g"]) and np.array_equal(e["err8_z"], o["err8_z"])
    assert np.array_equal(e["valid"], o["valid"])
    # everything at once through the one-pass kernel (stats + SAM + both planes share the staged mask bytes)
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
    from image_compression_analysis_b200.metrics import _valid_to_device
    import torch
    pair = DevicePair.from_arrays(r_bip, d_bip, "bip", kw["ref_nodata"], kw["tst_nodata"])
    P = evaluate(pair, Want(stats=True, sam=True, err8_caps=(255, 32)), _valid_to_device(valid, H, W, "shape"))
    torch.cuda.synchronize()
    assert np.array_equal(P.planes["err8_g"].cpu().numpy().reshape(H, W), o["err8_g"])
    assert np.array_equal(P.planes["err8_z"].cpu().numpy().reshape(H, W), o["err8_z"])


def test_evaluate_host_pairs_matches_pair_by_pair():
    """The pipelined sweep of pinned host pairs yields, in order, what evaluate() + to_host() yields per pair."""
    import torch
    from image_compression_analysis_b200 import synth
    from image_compression_analysis_b200.engine import DevicePair, Want, evaluate, evaluate_host_pairs
    want = Want(stats=True, sam=True)
    host, ref_out = [], []
    for seed in range(5):
        ref, dec = synth.case_b_pair(seed=40 + seed, bands=180, height=16, width=40, amp=1 + seed, layout="bsq")
        r = torch.from_numpy(np.ascontiguousarray(np.moveaxis(ref, 0, -1)).view(np.int16)).pin_memory()
        d = torch.from_numpy(np.ascontiguousarray(np.moveaxis(dec, 0, -1)).view(np.int16)).pin_memory()
        host.append((r.view(torch.uint16), d.view(torch.uint16)))
        ref_out.append(evaluate(DevicePair.from_arrays(r.view(torch.uint16), d.view(torch.uint16), "bip"), want).to_host())
    got = list(evaluate_host_pairs(iter(host), want, layout="bip"))
    assert len(got) == len(ref_out)
    for g, w in zip(got, ref_out):
        assert np.array_equal(g.isum, w.isum) and np.array_equal(g.imax, w.imax)
        assert np.array_equal(g.fsum.view(np.int64), w.fsum.view(np.int64))       # same kernels, same order: bit-identical
    assert list(evaluate_host_pairs(iter([]), want)) == []


@pytest.mark.parametrize("layout", ["bsq", "bip"])
@pytest.mark.parametrize("world", [2, 3, 5])
def test_row_strips_with_halos_match_single_shot(layout, world):
    """the image cut into row strips (1 halo row for Sobel, 5 for the Gaussian SSIM, replicated
    at the cuts) and evaluated strip by strip gives the single-shot result: integers exactly, float sums to
    rounding -- and both agree with the oracle.  Covers the BIP Sobel kernel's strip arguments."""
    from image_compression_analysis_b200 import finish, sharding, synth
    from image_compression_analysis_b200.engine import Want
    from oracle import distortion_oracle as orc
    B, H, W = 6, 67, 45
    ref, dec = synth.case_b_pair(seed=77, bands=B, height=H, width=W, amp=5, layout="bsq")
    if layout == "bip":
        cube_r, cube_d = np.ascontiguousarray(np.moveaxis(ref, 0, -1)), np.ascontiguousarray(np.moveaxis(dec, 0, -1))
        cut = sharding.cut_bip
    else:
        cube_r, cube_d, cut = ref, dec, sharding.cut_bsq
    want = Want(stats=True, sam=True, sid=True, lmse=True, ssim_gauss=True)
    L = 4095.0
    tot = None
    for s in sharding.strips(H, world, halo=sharding.HALO_SSIM):
        P = sharding.evaluate_strip(cut(cube_r, s), cut(cube_d, s), s, H, layout, want, data_range=L, reduce=False)
        h = P.to_host()
        if tot is None:
            tot = h
        else:
            tot.isum += h.isum
            tot.imax = np.maximum(tot.imax, h.imax)
            tot.fsum += h.fsum
    got = finish.finish_compute_metrics(1, tot.sums, tot.maxs)
    got.update(finish.finish_spectral(float(tot.spec[0]), float(tot.spec[1]), float(tot.spec[2]), tot.lmse, H * W))
    got.update(finish.finish_ssim_gauss(tot.ssimw_sum, tot.ssimw_cnt))
    want_d = orc.compute_metrics(ref, dec, extras=False)
    want_d.update(orc.compute_sam_sid_lmse_caseB(ref, dec))
    want_d.update(orc.ssim_gaussian(ref, dec, L))
    for k, w in want_d.items():
        if isinstance(w, np.ndarray):
            continue
        if isinstance(w, (int, np.integer)):
            assert got[k] == int(w), (k, got[k], w)
        else:
            assert _close(got[k], w), (k, got[k], w)


@pytest.mark.parametrize("layout", ["bsq", "bip"])
def test_lmse_flat_and_zero_images(layout):
    """Zero gradients everywhere (the branch-free square root must return 0, not NaN) and a pair whose
    gradients differ in one pixel only."""
    import image_compression_analysis_b200 as dm
    from oracle import distortion_oracle as orc
    B, H, W = 4, 12, 20
    flat = np.full((B, H, W), 1234, np.uint16)
    zero = np.zeros((B, H, W), np.uint16)
    spike = flat.copy()
    spike[1, 5, 7] = 60000
    for a, b in ((flat, flat), (zero, zero), (flat, zero), (flat, spike)):
        ra, rb = (a, b) if layout == "bsq" else (np.ascontiguousarray(np.moveaxis(a, 0, -1)), np.ascontiguousarray(np.moveaxis(b, 0, -1)))
        got = dm.compute_sam_sid_lmse_caseB_arrays(ra, rb, layout=layout)
        want = orc.compute_sam_sid_lmse_caseB(a, b)
        assert not math.isnan(got["lmse"]) and _close(got["lmse"], want["lmse"]), (got, want)
        assert _close(got["sid"], want["sid"]) and _close(got["sam_deg"], want["sam_deg"]), (got, want)


@pytest.fixture
def bip_variant():
    """Pins the build of the one-pass BIP kernel (dm_fused_bip_variant) for one test, restores the default after."""
    from image_compression_analysis_b200._lib import check, lib

    def pin(v):
        check(lib().dm_fused_bip_variant(v))
    yield pin
    lib().dm_fused_bip_variant(0)


@pytest.mark.parametrize("variant", [12, 23, 1])
@pytest.mark.parametrize("args", [("uint16", 180, 37, 53, 5, False), ("uint16", 180, 64, 64, 65535, True),
                                  ("int16", 180, 40, 41, 30000, True)])
def test_fused_ct_both_band_warp_layouts(args, variant, bip_variant):
    """The 180-band kernel exists with 12 band warps (ldmatrix.x4, 96 registers) and with 23 (ldmatrix.x2, 64
    registers), next to the run-time-geometry kernel (1); dm_fused_bip_variant pins one.  All must match the oracle."""
    bip_variant(variant)
    test_fused_bip_vs_oracle(*args)


@pytest.mark.parametrize("variant", [12, 23, 1])
def test_nodata_180_bands_both_band_warp_layouts(variant, bip_variant):
    bip_variant(variant)
    test_nodata_180_bands_vs_oracle("int16", -32768, True, False)
    test_nodata_180_bands_vs_oracle("uint16", 65535, True, False)


def test_fused_bip_variant_rejects_unknown_values():
    from image_compression_analysis_b200 import _lib
    assert _lib.lib().dm_fused_bip_variant(7) == _lib.DM_EARG
    assert _lib.lib().dm_fused_bip_variant(0) == _lib.DM_OK


@pytest.mark.parametrize("layout", ["bsq", "bip"])
def test_error_max8_percentile_scaling(layout):
    """err_max_global=None (quicklooks.py:137-146): the 2nd / 98th percentile of the non-zero errors from the
    device histogram of the error plane, the reference's float64 evaluation of the stretch."""
    from image_compression_analysis_b200 import quicklooks as ql
    from oracle import distortion_oracle as orc
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 4000, (6, 30, 41)).astype(np.uint16)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-40, 41, ref.shape), 0, 65535).astype(np.uint16)
    dec[:, :4] = ref[:, :4]
    for a, b in ((ref, dec), (ref, ref)):                      # identical pair: no non-zero error at all
        want = orc.error_max8(a, b, None, 16)
        ra, rb = (a, b) if layout == "bsq" else (np.ascontiguousarray(np.moveaxis(a, 0, -1)), np.ascontiguousarray(np.moveaxis(b, 0, -1)))
        got = ql.error_max8_arrays(ra, rb, None, 16, layout=layout)
        assert got["cap_g"] == want["cap_g"]
        assert np.array_equal(got["err8_g"], want["err8_g"]) and np.array_equal(got["err8_z"], want["err8_z"])


def test_many_band_cube_read_through_rasterio_is_evaluated_as_bip():
    """rasterio hands (B,H,W); a 180-band cube is transposed once on the device so that the three drop-in
    calls take the BIP kernels (one-pass stats + SAM, register-resident SID, BIP Sobel).  Results vs oracle."""
    from pathlib import Path
    from oracle import distortion_oracle as orc, rasterio_stub
    rasterio_stub.install()
    import image_compression_analysis_b200 as dm
    from image_compression_analysis_b200 import ingest, quicklooks as ql
    B, H, W, nd = 180, 21, 35, -32768
    ref, dec = _rand_pair(61, "int16", B, H, W, 12)
    ref[ref == nd] += 1; dec[dec == nd] += 1
    bad = np.random.default_rng(62).random((H, W)) < 0.1
    ref[:, bad] = nd; dec[:, bad] = nd
    rasterio_stub.clear()
    rasterio_stub.register("/mem/src.tif", ref, nodata=nd)
    rasterio_stub.register("/mem/recon.tif", dec, nodata=nd)
    pair, _ = ingest.load_pair("/mem/src.tif", "/mem/recon.tif")
    assert pair.layout == "bip" and tuple(pair.ref.shape) == (H, W, B)
    got = dm.compute_metrics(Path("/mem/src.tif"), Path("/mem/recon.tif"))
    _check_metrics(got, orc.compute_metrics(ref, dec, None, extras=False, ref_nodata=nd, tst_nodata=nd))
    spec = dm.compute_sam_sid_lmse_caseB(Path("/mem/src.tif"), Path("/mem/recon.tif"))
    for k, w in orc.compute_sam_sid_lmse_caseB(ref, dec, None, ref_nodata=nd, tst_nodata=nd).items():
        assert _close(spec[k], w), (k, spec[k], w)
    og, oz = ql.write_error_max8("/mem/src.tif", "/mem/recon.tif", "/mem/out/recon", err_max_global=255, err_max_zoom=32)
    o = orc.error_max8(ref, dec, 255, 32, ref_nodata=nd, tst_nodata=nd)
    assert np.array_equal(rasterio_stub.fetch(og).data[0], o["err8_g"])
    assert np.array_equal(rasterio_stub.fetch(oz).data[0], o["err8_z"])


def test_p2p_exchange_matches_nccl_on_two_gpus():
    """NVLink peer-memory exchange (dm_p2p_push / dm_p2p_combine) == NCCL all-gather + dm_combine_partials, bit
    for bit, on every rank.  Needs two GPUs: skipped on a one-GPU box (run with `gpurun --gpus 2`)."""
    import subprocess, sys, torch
    from pathlib import Path
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29561", str(root / "tools" / "check_p2p.py")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("p2p == nccl: True; same on all ranks: True") == 2


@pytest.mark.parametrize("dtype", ["uint16", "int16", "uint8"])
def test_sobel_mag_map_is_bit_identical(dtype):
    """sobel_mag(img) as a drop-in of its own (run_codec.py:123-137): exact integer gradients and a correctly
    rounded square root, so the float64 map equals the reference's bit for bit (odd sizes, 1-pixel-wide image)."""
    import image_compression_analysis_b200 as dm
    from oracle import distortion_oracle as orc
    info = np.iinfo(dtype)
    rng = np.random.default_rng(3)
    for H, W in ((37, 53), (1, 9), (8, 1), (64, 64)):
        img = rng.integers(info.min, info.max + 1, (H, W)).astype(dtype)
        got = dm.sobel_mag(img)
        want = orc.sobel_mag(img)
        assert got.dtype == np.float64 and got.shape == (H, W)
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (dtype, H, W)
    with pytest.raises(TypeError):
        dm.sobel_mag(np.zeros((4, 4), np.float32))
