"""Generate tests/golden/adj_*.npz by running the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_adjacent       (from the repo root, in the build container)

Rows either side of the distortion path: the reference's own functions
  tools/quicklooks.py           stretch_params_from_baseline, write_rgb_8bit            (:51-109)
  tools/make_baseline_A.py      to_12in16                                               (:137-170)
  tools/make_baseline_B.py      write_truncated_copy, make_scene_error_map              (:284-316, :324-419)
  tools/codecs/ccsds121/...     _diff1_bsq_*, _int1_bsq_*, _write/_read_raw_interleaved  (:44-85)
  tools/codecs/jpegls/...       _diff1_forward, _diff1_inverse                          (:92-120)
are executed under oracle/rasterio_stub.py on small seeded inputs; inputs and outputs are frozen in one
.npz per case.  The GPU box has no /root/reference, so the tests read these files instead.
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import rasterio_stub, reference_loader  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def smooth_cube(rng, bands, h, w, top, dtype):
    base = rng.normal(size=(bands, h // 4 + 2, w // 4 + 2))
    up = np.kron(base, np.ones((1, 4, 4)))[:, :h, :w]
    up = (up - up.min()) / (up.max() - up.min() + 1e-12)
    return (up * top).astype(dtype)


def rgb_cases():
    ql = reference_loader.quicklooks()
    rng = np.random.default_rng(101)
    out = {}
    # Case A: 4-band uint16 12-in-16, no nodata, default RGB order [3, 2, 1]
    a = (smooth_cube(rng, 4, 96, 130, 4095, np.uint16) << 4).astype(np.uint16)
    # Case B-like: int16 with nodata -32768 where band 1 is blanked, RGB order [5, 3, 1], other percentiles
    b = smooth_cube(rng, 6, 64, 81, 9000, np.int16) - 500
    bad = rng.random((64, 81)) < 0.1
    b[:, bad] = -32768
    b[0, 5:9, 7:30] = -32768                       # band 1 == nodata only: masked by _valid_mask_from_ds
    # flat band -> hi <= lo branch
    c = np.full((3, 20, 24), 1234, np.uint16)
    c[1] = rng.integers(0, 65536, (20, 24))
    for name, cube, nodata, order, pct in (("adj_rgb_caseA", a, None, [3, 2, 1], (2, 98)),
                                           ("adj_rgb_i16_nodata", b, -32768, [5, 3, 1], (1.5, 99.0)),
                                           ("adj_rgb_flat", c, None, [1, 2, 3], (2, 98))):
        rasterio_stub.clear()
        rasterio_stub.register("/mem/base.tif", cube, nodata=nodata)
        params = ql.stretch_params_from_baseline("/mem/base.tif", rgb_order=order, pct=pct)
        ql.write_rgb_8bit("/mem/base.tif", "/mem/rgb.tif", params, rgb_order=order)
        got = rasterio_stub.fetch("/mem/rgb.tif")
        out[name] = dict(cube=cube, nodata=np.array([np.nan if nodata is None else nodata]), order=np.array(order),
                         pct=np.array(pct, dtype=np.float64), params=np.array(params, dtype=np.float64), rgb=got.data)
    return out


def requant_cases():
    mb, ma = reference_loader.make_baseline_B(), reference_loader.make_baseline_A()
    rng = np.random.default_rng(102)
    out = {}
    x = rng.integers(-32768, 32768, (3, 70, 90)).astype(np.int16)
    x[:, rng.random((70, 90)) < 0.07] = -32768
    for name, cube, nodata, k in (("adj_trunc_i16_k2", x, -32768, 2),
                                  ("adj_trunc_u16_k3", rng.integers(0, 65536, (2, 33, 47)).astype(np.uint16), None, 3),
                                  ("adj_trunc_u16_nd", rng.integers(0, 40, (2, 33, 47)).astype(np.uint16), 7, 2)):
        rasterio_stub.clear()
        rasterio_stub.register("/mem/in.tif", cube, nodata=nodata)
        mb.write_truncated_copy("/mem/in.tif", "/mem/out.tif", k, tile=32)
        out[name] = dict(cube=cube, nodata=np.array([np.nan if nodata is None else nodata]), k=np.array([k]),
                         out=rasterio_stub.fetch("/mem/out.tif").data)
    y = rng.integers(0, 65536, (4, 50, 70)).astype(np.uint16)
    y[0, 0, :8] = [65527, 65528, 65535, 0, 7, 8, 15, 16]          # the uint16 sum wraps above 65527
    rasterio_stub.clear()
    rasterio_stub.register("/mem/in.tif", y)
    with tempfile.TemporaryDirectory() as td:
        src = Path(td) / "in.tif"
        src.write_bytes(b"x")                                      # _assert_inputs_exist wants a real file
        rasterio_stub.register(src, y)
        ma.to_12in16(src, Path(td) / "o" / "out.tif")
        out["adj_to12in16"] = dict(cube=y, out=rasterio_stub.fetch(Path(td) / "o" / "out.tif").data)
    return out


def scene_cases():
    from PIL import Image
    mb = reference_loader.make_baseline_B()
    rng = np.random.default_rng(103)
    out = {}
    B, H, W = 9, 530, 37                                           # two 512-row strips
    ref = smooth_cube(rng, B, H, W, 12000, np.int16)
    cmp_ = mb.trunc_uint16(ref.view(np.uint16), 2).view(np.int16)
    big = (ref + rng.integers(-3000, 3001, ref.shape)).astype(np.int16)      # rms/mean with float32 rounding at work
    mask = rng.random((H, W)) > 0.15
    u_ref = rng.integers(0, 60000, (5, 40, 52)).astype(np.uint16)
    u_cmp = (u_ref & 0xFFF8).astype(np.uint16)
    with tempfile.TemporaryDirectory() as td:
        for tag, r, c, m, kb in (("i16_k2", ref, cmp_, mask, 2), ("i16_big", ref, big, None, 2), ("u16_k3", u_ref, u_cmp, None, 3)):
            for mode in ("mean", "rms", "count3", "max", "p95"):
                for scale in ("fixed", "auto"):
                    rasterio_stub.clear()
                    rasterio_stub.register("/mem/ref.tif", r)
                    rasterio_stub.register("/mem/cmp.tif", c)
                    mpath = None
                    if m is not None:
                        mp = Path(td) / "mask.tif"
                        mp.write_bytes(b"x")                       # read_mask() tests Path.exists()
                        rasterio_stub.register(mp, m.astype(np.uint8))
                        mpath = mp
                    png = Path(td) / f"{tag}_{mode}_{scale}.png"
                    mb.make_scene_error_map("/mem/ref.tif", "/mem/cmp.tif", mpath, scale, kb, png, err_mode=mode)
                    out.setdefault(f"adj_scene_{tag}", dict(ref=r, cmp=c, k_bits=np.array([kb]),
                                                            **({"mask": m} if m is not None else {})))
                    out[f"adj_scene_{tag}"][f"img_{mode}_{scale}"] = np.array(Image.open(png))
    return out


def transform_cases():
    cw, jw = reference_loader.ccsds121_wrap(), reference_loader.jpegls_wrap()
    rng = np.random.default_rng(104)
    out = {}
    s = rng.integers(-32768, 32768, (7, 33, 50)).astype(np.int16)
    u = rng.integers(0, 65536, (5, 24, 40)).astype(np.uint16)
    d_s, d_u = cw._diff1_bsq_signed(s), cw._diff1_bsq_unsigned(u)
    out["adj_diff1_ccsds"] = dict(s=s, u=u, d_s=d_s, d_u=d_u, i_s=cw._int1_bsq_signed(d_s.copy()),
                                  i_u=cw._int1_bsq_unsigned(d_u.copy()))
    rec = {}
    for dt, cube in (("uint16", u), ("int16", s), ("uint8", rng.integers(0, 256, (4, 19, 23)).astype(np.uint8))):
        fwd = np.stack([jw._diff1_forward(cube[b], cube[b - 1] if b else None, dt) for b in range(cube.shape[0])], 0)
        inv = np.empty_like(fwd)
        for b in range(cube.shape[0]):
            inv[b] = jw._diff1_inverse(fwd[b], inv[b - 1] if b else None, dt)
        rec.update({f"x_{dt}": cube, f"fwd_{dt}": fwd, f"inv_{dt}": inv})
    out["adj_diff1_jpegls"] = rec
    il = {}
    t = rng.integers(0, 65536, (6, 11, 17)).astype(np.uint16)
    il["tile"] = t
    with tempfile.TemporaryDirectory() as td:
        for mode in ("bsq", "bil", "bip"):
            p = Path(td) / f"{mode}.raw"
            cw._write_raw_interleaved(t, mode, p, np.dtype("<u2"))
            flat = np.fromfile(p, dtype="<u2")
            il[f"raw_{mode}"] = flat
            back = cw._read_raw_interleaved(p, mode, np.dtype("<u2"), *t.shape)
            assert np.array_equal(back, t)
    out["adj_interleave"] = il
    return out


def main():
    if not reference_loader.available():
        raise SystemExit("reference tree not mounted")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    allc = {}
    for fn in (rgb_cases, requant_cases, scene_cases, transform_cases):
        allc.update(fn())
    for name, rec in allc.items():
        np.savez_compressed(GOLDEN / f"{name}.npz", **rec)
        print(name, {k: getattr(v, "shape", None) for k, v in rec.items()})


if __name__ == "__main__":
    main()
