"""CPU: the minimal GeoTIFF reader/writer (raster_io's stand-in for rasterio on the metric path)."""
import numpy as np
import pytest

from image_compression_analysis_b200 import geotiff


def _cube(dtype, B, H, W, seed=0):
    info = np.iinfo(dtype)
    return np.random.default_rng(seed).integers(info.min, int(info.max) + 1, size=(B, H, W)).astype(dtype)


@pytest.mark.parametrize("dtype,B,H,W,opts", [
    ("uint16", 4, 700, 530, dict(tiled=True, blockxsize=512, blockysize=512, compress="DEFLATE")),
    ("int16", 7, 65, 33, dict(tiled=True, blockxsize=16, blockysize=32, compress=None)),
    ("uint8", 1, 100, 257, dict(tiled=True, blockxsize=64, blockysize=64, compress="DEFLATE")),
    ("uint16", 3, 40, 50, dict(tiled=True, blockxsize=32, blockysize=32, BIGTIFF="YES")),
])
def test_round_trip(tmp_path, dtype, B, H, W, opts):
    a = _cube(dtype, B, H, W)
    path = tmp_path / "x.tif"
    with geotiff.open(path, "w", driver="GTiff", dtype=dtype, count=B, width=W, height=H, nodata=None, **opts) as dst:
        dst.write(a)
        dst.update_tags(STATISTICS_MEAN="1.5", NOTE="a<b&c")
    with geotiff.open(path) as src:
        assert (src.count, src.height, src.width, src.dtypes[0]) == (B, H, W, dtype)
        assert src.nodata is None
        assert np.array_equal(src.read(), a)
        assert np.array_equal(src.read(1), a[0])
        if B >= 3:
            assert np.array_equal(src.read([3, 1]), a[[2, 0]])
        assert src.read(1, out_dtype="int32").dtype == np.int32
        nat, layout = src.read_native()
        assert layout == ("bip" if B > 1 else "bsq")
        assert np.array_equal(nat if B == 1 else np.moveaxis(nat, -1, 0), a)
        assert np.array_equal(src.dataset_mask(), np.full((H, W), 255, np.uint8))
    assert geotiff.read_tags(path) == {"STATISTICS_MEAN": "1.5", "NOTE": "a<b&c"}


def test_nodata_and_mask_sidecar(tmp_path):
    a = _cube("int16", 2, 48, 40, seed=3)
    a[:, :5, :7] = -32768
    a[0, 10, 10] = -32768                      # one band only: still valid for dataset_mask
    path = tmp_path / "n.tif"
    with geotiff.open(path, "w", dtype="int16", count=2, width=40, height=48, nodata=-32768, compress="DEFLATE") as dst:
        dst.write(a)
    with geotiff.open(path) as src:
        assert src.nodata == -32768
        m = src.dataset_mask()
        want = np.full((48, 40), 255, np.uint8); want[:5, :7] = 0
        assert np.array_equal(m, want)
    msk = np.zeros((48, 40), np.uint8); msk[20:, :] = 255
    p2 = tmp_path / "m.tif"
    with geotiff.open(p2, "w", dtype="uint8", count=1, width=40, height=48, compress="DEFLATE") as dst:
        dst.write(np.zeros((1, 48, 40), np.uint8))
        dst.write_mask(msk)
    with geotiff.open(p2) as src:
        assert np.array_equal(src.dataset_mask(), msk)


def test_against_pillow(tmp_path):
    """Files written by libtiff (through Pillow): strips, DEFLATE; and our output read back by Pillow."""
    Image = pytest.importorskip("PIL.Image")
    a = _cube("uint16", 1, 301, 211, seed=5)[0]
    p = tmp_path / "pil.tif"
    Image.fromarray(a).save(p, compression="tiff_deflate")
    with geotiff.open(p) as src:
        assert (src.count, src.dtypes[0], src.tiled) == (1, "uint16", False)
        assert np.array_equal(src.read(1), a)
    p0 = tmp_path / "pil_raw.tif"
    Image.fromarray(a).save(p0)
    with geotiff.open(p0) as src:
        assert np.array_equal(src.read(1), a)
    b = _cube("uint8", 1, 130, 70, seed=6)
    q = tmp_path / "ours.tif"
    with geotiff.open(q, "w", dtype="uint8", count=1, width=70, height=130, tiled=True, blockxsize=64, blockysize=64,
                      compress="DEFLATE") as dst:
        dst.write(b)
    assert np.array_equal(np.asarray(Image.open(q)), b[0])


def test_unsupported_features_raise(tmp_path):
    Image = pytest.importorskip("PIL.Image")
    a = _cube("uint16", 1, 20, 20)[0]
    p = tmp_path / "lzw.tif"
    Image.fromarray(a).save(p, compression="tiff_lzw")
    with pytest.raises(NotImplementedError, match="compression"):
        geotiff.open(p)
