"""Generate tests/golden/ fixtures by running the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden            (from the repo root, in the build container)

For every case below the real reference functions
(`/root/reference/tools/run_codec.py:240-347`, `tools/quicklooks.py:115-207`)
are executed under oracle/rasterio_stub.py on small seeded inputs; inputs go to
tests/golden/<case>.npz, reference outputs to tests/golden/<case>.json (+ the
ERR8 planes inside the npz).  The GPU box has no /root/reference, so the `-m gpu`
parity tests and the CPU oracle tests read these files instead.
"""
from __future__ import annotations

import json
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import rasterio_stub, reference_loader  # noqa: E402
from image_compression_analysis_b200 import synth   # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def _jsonable(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        v = float(v)
        if math.isnan(v):
            return "nan"
        if math.isinf(v):
            return "inf" if v > 0 else "-inf"
        return float.hex(v)            # exact round trip
    return v


def cases():
    """name -> dict(ref, tst, valid, ref_nodata, tst_nodata, caps=[(g,z),...], case_b=bool)."""
    out = {}
    # Case A, 12-in-16 uint16, Gaussian noise decode (scaled down)
    ref, dec = synth.case_a_pair(seed=11, bands=4, height=48, width=64, sigma=2.0)
    out["a_gauss"] = dict(ref=ref, tst=dec, caps=[(255, 32), (255, None)], case_b=True)
    # identical pair -> inf / 1.0 / 0 / lossless
    ref, dec = synth.case_a_pair(seed=12, bands=4, height=32, width=40, mode="identical")
    out["a_identical"] = dict(ref=ref, tst=dec, caps=[(255, 32)], case_b=True)
    # NEAR=3 pair with a valid mask
    ref, dec = synth.case_a_pair(seed=13, bands=4, height=40, width=56, mode="near3")
    out["a_near3_masked"] = dict(ref=ref, tst=dec, valid=synth.random_valid_mask(13, 40, 56, 0.2),
                                 caps=[(48, 16)], case_b=True)
    # all-False mask => compute_metrics evaluates everything, Case-B metrics give NaN
    out["a_mask_all_false"] = dict(ref=ref, tst=dec, valid=np.zeros((40, 56), bool), caps=[(255, None)], case_b=True)
    # full-range uint16 (not 12-in-16), one band lossless
    rng = np.random.default_rng(14)
    ref = rng.integers(0, 65536, size=(3, 24, 36)).astype(np.uint16)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-700, 701, size=ref.shape), 0, 65535).astype(np.uint16)
    dec[1] = ref[1]
    out["u16_fullrange"] = dict(ref=ref, tst=dec, caps=[(255, 32), (65535, 1000)], case_b=True)
    # uint8
    ref = rng.integers(0, 256, size=(3, 20, 28)).astype(np.uint8)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-9, 10, size=ref.shape), 0, 255).astype(np.uint8)
    out["u8"] = dict(ref=ref, tst=dec, caps=[(255, 8)], case_b=True)
    # Case B uint16 14-in-16, B=20, with valid mask
    ref, dec = synth.case_b_pair(seed=15, bands=20, height=24, width=28, amp=3)
    out["b_u16_masked"] = dict(ref=ref, tst=dec, valid=synth.random_valid_mask(15, 24, 28, 0.1),
                               caps=[(255, 32)], case_b=True)
    # Case B int16 14-in-16 with nodata -32768 (whole-pixel + stray single-band hits), no valid mask
    ref, dec = synth.case_b_pair(seed=16, bands=12, height=32, width=40, amp=5, dtype="int16")
    inval = ~synth.random_valid_mask(16, 32, 40, 0.07)
    ref = synth.plant_nodata(ref, -32768, inval, extra_hits=6, seed=16)
    dec = synth.plant_nodata(dec, -32768, inval, extra_hits=3, seed=17)
    out["b_i16_nodata"] = dict(ref=ref, tst=dec, ref_nodata=-32768, tst_nodata=-32768,
                               caps=[(255, 32)], case_b=True)
    # same data, nodata + valid mask together
    out["b_i16_nodata_masked"] = dict(ref=ref, tst=dec, ref_nodata=-32768, tst_nodata=-32768,
                                      valid=synth.random_valid_mask(18, 32, 40, 0.15),
                                      caps=[(255, None)], case_b=True)
    # int16, no nodata declared, values reaching -32768 (np.abs wraps, run_codec.py:285)
    ref = rng.integers(-32768, 32768, size=(2, 16, 24)).astype(np.int16)
    ref[0, 0, 0] = -32768
    dec = np.clip(ref.astype(np.int64) + rng.integers(-40, 41, size=ref.shape), -32768, 32767).astype(np.int16)
    dec[0, 0, 0] = -32768
    out["i16_fullrange_wrap"] = dict(ref=ref, tst=dec, caps=[(255, 32)], case_b=True)
    # zero spectra for SAM (arccos(0)) and flat spectra for SID
    ref, dec = synth.case_b_pair(seed=19, bands=8, height=12, width=16, amp=2)
    ref[:, 0, :4] = 0
    dec[:, 1, :4] = 0
    ref[:, 2, :4] = 7
    dec[:, 2, :4] = 7
    out["b_zero_spectra"] = dict(ref=ref, tst=dec, caps=[(255, 2)], case_b=True)
    # odd sizes / single band
    ref = rng.integers(0, 4096, size=(1, 17, 23)).astype(np.uint16) << 4
    dec = (np.clip(ref.astype(np.int64) + 16 * rng.integers(-2, 3, size=ref.shape), 0, 65520)).astype(np.uint16)
    out["single_band_odd"] = dict(ref=ref, tst=dec, caps=[(255, 32)], case_b=True)
    # 7 bands (odd band count), odd width
    ref = rng.integers(0, 3000, size=(7, 19, 21)).astype(np.uint16)
    dec = np.clip(ref.astype(np.int64) + rng.integers(-20, 21, size=ref.shape), 0, 65535).astype(np.uint16)
    out["seven_bands_odd"] = dict(ref=ref, tst=dec, valid=synth.random_valid_mask(20, 19, 21, 0.3),
                                  caps=[(255, 32)], case_b=True)
    return out


def run_reference(case: dict) -> dict:
    rc = reference_loader.run_codec()
    ql = reference_loader.quicklooks()
    rasterio_stub.clear()
    rasterio_stub.register("/mem/ref.tif", case["ref"], nodata=case.get("ref_nodata"))
    rasterio_stub.register("/mem/tst.tif", case["tst"], nodata=case.get("tst_nodata"))
    valid = case.get("valid")
    res = {"compute_metrics": {k: _jsonable(v) for k, v in
                               rc.compute_metrics(Path("/mem/ref.tif"), Path("/mem/tst.tif"), valid=valid).items()}}
    if case.get("case_b"):
        res["sam_sid_lmse"] = {k: _jsonable(v) for k, v in
                               rc.compute_sam_sid_lmse_caseB(Path("/mem/ref.tif"), Path("/mem/tst.tif"),
                                                             valid=valid).items()}
    planes = {}
    res["err8"] = []
    for n, (g, z) in enumerate(case.get("caps", [])):
        og, oz = ql.write_error_max8("/mem/ref.tif", "/mem/tst.tif", f"/mem/out{n}/recon",
                                     err_max_global=g, err_max_zoom=z)
        rg = rasterio_stub.fetch(og)
        entry = {"cap_g": g, "cap_z": z, "name_g": Path(og).name, "tags_g": dict(rg.tags)}
        planes[f"err8_{n}_g"] = rg.data[0]
        planes[f"err8_{n}_mask"] = rg.mask.astype(np.uint8)
        if oz is not None:
            rz = rasterio_stub.fetch(oz)
            entry.update(name_z=Path(oz).name, tags_z=dict(rz.tags))
            planes[f"err8_{n}_z"] = rz.data[0]
        res["err8"].append(entry)
    return res, planes


def main():
    if not reference_loader.available():
        raise SystemExit("reference tree not mounted; golden fixtures can only be generated in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    for name, case in cases().items():
        res, planes = run_reference(case)
        arrays = {"ref": case["ref"], "tst": case["tst"]}
        if case.get("valid") is not None:
            arrays["valid"] = case["valid"]
        arrays.update(planes)
        np.savez_compressed(GOLDEN / f"{name}.npz", **arrays)
        res["meta"] = {"ref_nodata": case.get("ref_nodata"), "tst_nodata": case.get("tst_nodata"),
                       "dtype": str(case["ref"].dtype), "shape": list(case["ref"].shape),
                       "generator": "oracle/make_golden.py (unmodified reference under oracle/rasterio_stub.py)",
                       "numpy": np.__version__}
        (GOLDEN / f"{name}.json").write_text(json.dumps(res, indent=1, sort_keys=True))
        print(f"[golden] {name}: {case['ref'].dtype} {case['ref'].shape}")


if __name__ == "__main__":
    main()
