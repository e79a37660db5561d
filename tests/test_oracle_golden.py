"""CPU: the numpy oracle (oracle/distortion_oracle.py) against the reference-generated goldens."""
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import distortion_oracle as orc
from tests import goldenio


@pytest.mark.parametrize("name", goldenio.names())
def test_compute_metrics_matches_reference_golden(name):
    c = goldenio.load(name)
    got = orc.compute_metrics(c["ref"], c["tst"], c["valid"], ref_nodata=c["ref_nodata"],
                              tst_nodata=c["tst_nodata"], extras=False)
    want = c["compute_metrics"]
    assert set(got) == set(want)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, int):
            assert isinstance(g, int) and g == w, k
        else:
            # same numpy ops in the same order: expect bit equality, not just 1e-6
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)


@pytest.mark.parametrize("name", goldenio.names())
def test_sam_sid_lmse_matches_reference_golden(name):
    c = goldenio.load(name)
    got = orc.compute_sam_sid_lmse_caseB(c["ref"], c["tst"], c["valid"], ref_nodata=c["ref_nodata"],
                                         tst_nodata=c["tst_nodata"])
    for k, w in c["sam_sid_lmse"].items():
        g = got[k]
        assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)


@pytest.mark.parametrize("name", goldenio.names())
def test_error_max8_matches_reference_golden(name):
    c = goldenio.load(name)
    for n, e in enumerate(c["err8"]):
        got = orc.error_max8(c["ref"], c["tst"], e["cap_g"], e["cap_z"], ref_nodata=c["ref_nodata"],
                             tst_nodata=c["tst_nodata"])
        assert np.array_equal(got["err8_g"], c["planes"][f"err8_{n}_g"])
        assert np.array_equal(got["valid"].astype(np.uint8), c["planes"][f"err8_{n}_mask"])
        assert e["name_g"] == f"recon_ERR8_0_{got['cap_g']}.tif"
        assert float(e["tags_g"]["STATISTICS_MEAN"]) == got["mean_g"]
        assert float(e["tags_g"]["STATISTICS_STDDEV"]) == got["std_g"]
        if e["cap_z"] is not None:
            assert np.array_equal(got["err8_z"], c["planes"][f"err8_{n}_z"])
            assert e["name_z"] == f"recon_ERR8_0_{got['cap_z']}.tif"
        # the LUT restatement reproduces the float32 chain for every integer error
        lut = orc.err8_lut(e["cap_g"])
        err_i = got["err"].astype(np.int64)
        assert np.array_equal(lut[np.minimum(err_i, e["cap_g"])], got["err8_g"])


def test_extras_self_consistency():
    c = goldenio.load("a_gauss")
    got = orc.compute_metrics(c["ref"], c["tst"], hist_bins=256)
    for i in range(1, 5):
        h = got[f"hist_b{i}"]
        k = np.arange(256)
        assert h.sum() == got["n_valid"]
        assert (h * k * k).sum() == got[f"sse_b{i}"]
        assert (h * k).sum() / got["n_valid"] == got[f"mae_b{i}"]
        assert np.nonzero(h)[0].max() == got[f"maxerr_b{i}"]


def test_gaussian_taps_are_scipy_taps():
    from scipy.ndimage import gaussian_filter
    imp = np.zeros(41); imp[20] = 1.0
    resp = gaussian_filter(imp, sigma=1.5, truncate=3.5, mode="reflect")
    taps = orc.gaussian_taps()
    assert taps.size == 11
    assert np.allclose(resp[15:26], taps, rtol=0, atol=1e-17)
    assert abs(taps[5] - 0.266011724862) < 1e-11 and abs(taps[0] - 0.001028380084) < 1e-11


@pytest.mark.parametrize("dtype,L,amp", [("uint16", 4095.0, 40), ("uint16", 65535.0, 30000), ("int16", 8191.0, 500), ("uint8", 255.0, 9)])
def test_two_independent_gaussian_ssim_statements_agree(dtype, L, amp):
    """scipy.ndimage (separable, reflect + crop) against the written-out 11 x 11 window sums of the SSIM paper."""
    rng = np.random.default_rng(17)
    info = np.iinfo(dtype)
    a = rng.integers(info.min // 2, info.max // 2, size=(37, 53)).astype(np.int64)
    b = np.clip(a + rng.integers(-amp, amp + 1, size=a.shape), info.min, info.max)
    a, b = a.astype(dtype), b.astype(dtype)
    s1 = orc.ssim_gaussian_band(a, b, L)
    s2 = orc.ssim_gaussian_band_direct(a, b, L)
    assert abs(s1 - s2) <= 1e-11 * max(1.0, abs(s1)), (s1, s2)
    assert orc.ssim_gaussian_band_direct(a, a, L) == pytest.approx(1.0, abs=1e-12)


def _cv2_fixture():
    z = np.load(Path(__file__).resolve().parent / "golden" / "ssimw_cv2.npz")
    for n in sorted({k.split("__")[0] for k in z.files}):
        yield n, z[n + "__a"], z[n + "__b"], float(z[n + "__L"]), float(z[n + "__ssim"])


def test_gaussian_ssim_oracle_pinned_on_opencv_fixtures():
    """Third-party pin of the one metric the reference does not have: both oracle statements of the Gaussian-window
    SSIM against values computed with OpenCV's GaussianBlur(11 x 11, sigma 1.5) (oracle/make_golden_ssim_cv2.py,
    the formulation of OpenCV's own SSIM sample + skimage's 5-px crop).  Observed agreement 1e-15; bar 1e-12."""
    from oracle import distortion_oracle as orc
    n_cases = 0
    for name, a, b, L, want in _cv2_fixture():
        assert orc.ssim_gaussian_band(a, b, L) == pytest.approx(want, rel=1e-12), name
        assert orc.ssim_gaussian_band_direct(a, b, L) == pytest.approx(want, rel=1e-12), name
        n_cases += 1
    assert n_cases == 5


def test_opencv_fixtures_reproduce_when_opencv_is_installed():
    """The committed fixture is what OpenCV says today (skipped where cv2 is missing)."""
    pytest.importorskip("cv2")
    from oracle import make_golden_ssim_cv2 as mk
    for name, a, b, L, want in _cv2_fixture():
        assert mk.ssim_cv2(a, b, L) == pytest.approx(want, rel=1e-13), name
    for name, dtype, shape, L, amp, seed in mk.CASES:
        a, b = mk.make_pair(dtype, shape, amp, seed)
        fa = [x for x in _cv2_fixture() if x[0] == name][0]
        assert np.array_equal(a, fa[1]) and np.array_equal(b, fa[2]), name
