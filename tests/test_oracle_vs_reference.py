"""CPU, build container only: the numpy oracle against the LIVE unmodified reference
(loaded under oracle/rasterio_stub.py) on fresh random inputs.  Skipped where
/root/reference is not mounted (the GPU box) -- tests/golden covers that case."""
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import distortion_oracle as orc
from oracle import rasterio_stub, reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference tree not mounted")


def _same(g, w):
    if isinstance(w, (int, np.integer)) and not isinstance(w, bool):
        return int(g) == int(w)
    return (math.isnan(g) and math.isnan(w)) or g == w


def _pair(rng, dtype, B, H, W, amp):
    info = np.iinfo(dtype)
    ref = rng.integers(info.min, int(info.max) + 1, size=(B, H, W)).astype(dtype)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-amp, amp + 1, size=ref.shape), info.min, info.max).astype(dtype)
    return ref, dec


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("dtype", ["uint8", "uint16", "int16"])
def test_random_pairs(seed, dtype):
    rng = np.random.default_rng(1000 + seed)
    B, H, W = int(rng.integers(1, 9)), int(rng.integers(3, 30)), int(rng.integers(3, 30))
    ref, dec = _pair(rng, dtype, B, H, W, amp=int(rng.integers(0, 50)))
    nodata = None
    if dtype == "int16" and seed % 2:
        nodata = -32768
        hit = rng.random((H, W)) < 0.1
        ref[:, hit] = nodata
        dec[:, hit] = nodata
        ref[0, 0, 0] = nodata
    valid = (rng.random((H, W)) < 0.8) if seed % 3 == 0 else None
    rc, ql = reference_loader.run_codec(), reference_loader.quicklooks()
    rasterio_stub.clear()
    rasterio_stub.register("/m/a.tif", ref, nodata=nodata)
    rasterio_stub.register("/m/b.tif", dec, nodata=nodata)
    want = rc.compute_metrics(Path("/m/a.tif"), Path("/m/b.tif"), valid=valid)
    got = orc.compute_metrics(ref, dec, valid, ref_nodata=nodata, tst_nodata=nodata, extras=False)
    assert set(got) == set(want)
    for k in want:
        assert _same(got[k], want[k]), k
    want = rc.compute_sam_sid_lmse_caseB(Path("/m/a.tif"), Path("/m/b.tif"), valid=valid)
    got = orc.compute_sam_sid_lmse_caseB(ref, dec, valid, ref_nodata=nodata, tst_nodata=nodata)
    for k in want:
        assert _same(got[k], want[k]), k
    og, oz = ql.write_error_max8("/m/a.tif", "/m/b.tif", "/m/o/recon", err_max_global=200, err_max_zoom=17)
    got = orc.error_max8(ref, dec, 200, 17, ref_nodata=nodata, tst_nodata=nodata)
    assert np.array_equal(rasterio_stub.fetch(og).data[0], got["err8_g"])
    assert np.array_equal(rasterio_stub.fetch(oz).data[0], got["err8_z"])
    assert np.array_equal(rasterio_stub.fetch(og).mask, got["valid"])


def test_scalar_helpers():
    rc = reference_loader.run_codec()
    rng = np.random.default_rng(5)
    a = rng.integers(0, 65536, size=(33, 41)).astype(np.uint16)
    b = rng.integers(0, 65536, size=(33, 41)).astype(np.uint16)
    assert orc.mse(a, b) == rc.mse(a, b)
    assert orc.psnr(a, b, 65535) == rc.psnr(a, b, 65535)
    assert orc.psnr(a, a, 65535) == rc.psnr(a, a, 65535) == float("inf")
    assert orc.ssim_global(a, b, 4095) == rc.ssim_global(a, b, 4095)
    assert np.array_equal(orc.sobel_mag(a), rc.sobel_mag(a))


def test_error_max8_percentile_branch_vs_reference(tmp_path):
    """write_error_max8 with err_max_global=None (quicklooks.py:137-146, the CLI's percentile scaling)."""
    from oracle import distortion_oracle as orc, rasterio_stub, reference_loader as rl
    if not rl.available():
        pytest.skip("reference tree not mounted")
    ql = rl.quicklooks()
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 4000, (5, 30, 41)).astype(np.uint16)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-40, 41, ref.shape), 0, 65535).astype(np.uint16)
    dec[:, :4] = ref[:, :4]                                     # zero-error pixels are excluded from the percentiles
    rasterio_stub.clear()
    rasterio_stub.register("/mem/a.tif", ref)
    rasterio_stub.register("/mem/b.tif", dec)
    og, oz = ql.write_error_max8("/mem/a.tif", "/mem/b.tif", "/mem/out", err_max_global=None, err_max_zoom=16)
    want = orc.error_max8(ref, dec, None, 16)
    assert str(og).endswith(f"_ERR8_0_{want['cap_g']}.tif")
    assert np.array_equal(rasterio_stub.fetch(og).data[0], want["err8_g"])
    assert np.array_equal(rasterio_stub.fetch(oz).data[0], want["err8_z"])
