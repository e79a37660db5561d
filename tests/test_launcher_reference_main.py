"""CPU: the zero-edit switch of INTEGRATION.md, exercised on the REAL reference driver.

1. `--quicklooks <plugin>` exactly as run_codec.py loads it (tools/run_codec.py:419-430: the file's directory
   goes first on sys.path, then `import quicklooks`) must yield THIS repo's module, not the silent fall-back
   to the reference's CPU quicklooks.py -- in a fresh interpreter, with and without the repo on PYTHONPATH.
2. `launcher.main()` runs the reference's unmodified `run_codec.main()` (:374-670) over a two-rep Case-B
   manifest under the in-memory rasterio stand-in.  There is no GPU in this container and no reference on
   the GPU box, so the three rebound entry points are replaced by recorders that answer from the numpy
   oracle: what is checked is the plumbing -- main() calls the rebound functions and the plugin's
   write_error_max8 with the reference's arguments, and their values land in metrics.csv.
"""
import csv
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PLUGIN = ROOT / "image_compression_analysis_b200" / "plugin" / "quicklooks.py"


@pytest.mark.parametrize("with_pythonpath", [False, True])
def test_plugin_imports_the_way_run_codec_imports_it(with_pythonpath, tmp_path):
    code = (
        "import sys\n"
        "from pathlib import Path\n"
        f"ql_path = Path({str(PLUGIN)!r})\n"
        "assert ql_path.exists()\n"
        "sys.path.insert(0, ql_path.parent.as_posix())      # run_codec.py:422\n"
        "import quicklooks as ql_mod                         # run_codec.py:424\n"
        "assert getattr(ql_mod, 'B200_NATIVE', False), ql_mod.__file__\n"
        "for f in ('stretch_params_from_baseline', 'write_rgb_8bit', 'write_error_max8'):\n"
        "    assert callable(getattr(ql_mod, f)), f\n"
        "import image_compression_analysis_b200.quicklooks as ours\n"
        "assert ql_mod.write_error_max8 is ours.write_error_max8\n"
        "# nothing else of the package became importable by a bare name\n"
        "import importlib.util\n"
        "for bare in ('metrics', 'geotiff', 'transforms', 'engine'):\n"
        "    assert importlib.util.find_spec(bare) is None, bare\n"
        "print('ok')\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    if with_pythonpath:
        env["PYTHONPATH"] = str(ROOT)
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_plugin_directory_holds_only_the_plugin():
    names = sorted(p.name for p in PLUGIN.parent.iterdir() if p.name != "__pycache__")
    assert names == ["quicklooks.py"], names


def test_launcher_runs_the_unmodified_reference_main(tmp_path, monkeypatch):
    from oracle import distortion_oracle as orc
    from oracle import rasterio_stub, reference_loader
    root = reference_loader.reference_root()
    if root is None:
        pytest.skip("reference tree not mounted")
    rasterio_stub.install()
    import image_compression_analysis_b200 as dm
    import image_compression_analysis_b200.quicklooks as ours_ql
    from image_compression_analysis_b200 import launcher

    rng = np.random.default_rng(5)
    B, H, W = 6, 24, 40
    ref = (rng.integers(0, 2500, (B, H, W)) * 4).astype(np.uint16)
    src = tmp_path / "data" / "tileB.tif"
    src.parent.mkdir()
    src.write_bytes(b"")                                   # main() asserts the path exists (:450)
    rasterio_stub.register(src, ref, nodata=None)
    outdir = tmp_path / "runs"
    recs = []
    for rep in (1, 2):                                     # an existing recon.tif means "codec already ran" (:489-491)
        d = outdir / "t0" / "norate" / f"rep_{rep:02d}"
        d.mkdir(parents=True)
        dec = (ref.astype(np.int32) + rng.integers(-rep, rep + 1, ref.shape)).clip(0, 65535).astype(np.uint16)
        (d / "recon.tif").write_bytes(b"")
        rasterio_stub.register(d / "recon.tif", dec)
        recs.append(dec)
    idx = tmp_path / "index.json"
    idx.write_text(json.dumps({"case": "caseB", "asset": "tile_512", "items": [{"tile_id": "t0", "path": str(src)}]}))

    calls = {"cm": [], "sp": [], "ql": []}

    def fake_cm(ref_path, tst_path, valid=None):
        calls["cm"].append((str(ref_path), str(tst_path)))
        return orc.compute_metrics(rasterio_stub.fetch(ref_path).data, rasterio_stub.fetch(tst_path).data, valid, extras=False)

    def fake_sp(ref_path, tst_path, valid=None):
        calls["sp"].append((str(ref_path), str(tst_path)))
        return orc.compute_sam_sid_lmse_caseB(rasterio_stub.fetch(ref_path).data, rasterio_stub.fetch(tst_path).data, valid)

    def fake_ql(a_path, b_path, out_path_base, err_max_global=255, err_max_zoom=None, pct=(2, 98)):
        calls["ql"].append((a_path, b_path, out_path_base, err_max_global, err_max_zoom))
        return Path(out_path_base + "_ERR8_0_255.tif"), None

    monkeypatch.setattr(dm, "compute_metrics", fake_cm)
    monkeypatch.setattr(dm, "compute_sam_sid_lmse_caseB", fake_sp)
    monkeypatch.setattr(ours_ql, "write_error_max8", fake_ql)
    for m in ("quicklooks", "run_codec"):
        monkeypatch.delitem(sys.modules, m, raising=False)
    monkeypatch.setattr(sys, "path", list(sys.path))
    monkeypatch.setattr(sys, "argv", list(sys.argv))

    rc = launcher.main([str(root / "tools"), "--indices", str(idx), "--codec", "stub", "--outdir", str(outdir),
                        "--compressor-cmd", "false", "--reps", "2", "--ql-err-zoom", "32"])
    assert rc == 0
    # the module run_codec picked up through --quicklooks is the plugin, not tools/quicklooks.py
    assert getattr(sys.modules["quicklooks"], "B200_NATIVE", False)
    assert Path(sys.modules["run_codec"].__file__).resolve() == (root / "tools" / "run_codec.py").resolve()
    assert len(calls["cm"]) == 2 and len(calls["sp"]) == 2 and len(calls["ql"]) == 2
    assert calls["ql"][0][3:] == (255, 32) and calls["ql"][0][2].endswith("rep_01/recon")
    assert calls["cm"][1] == (src.resolve().as_posix(), (outdir / "t0/norate/rep_02/recon.tif").resolve().as_posix())
    # the values the rebound functions returned are what the reference wrote (';' + decimal comma, 6 decimals)
    with open(outdir / "metrics.csv", newline="") as f:
        rows = list(csv.DictReader(f, delimiter=";"))
    assert len(rows) == 2
    for row, dec in zip(rows, recs):
        want = orc.compute_metrics(ref, dec, None, extras=False)
        want.update(orc.compute_sam_sid_lmse_caseB(ref, dec, None))
        for k in ("psnr_global", "ssim_global", "sam_deg", "psnr_b1", "ssim_b6"):
            assert abs(float(row[k].replace(",", ".")) - want[k]) <= 1e-6 * max(1.0, abs(want[k])), (k, row[k], want[k])
        assert int(float(row["max_abs_err"].replace(",", "."))) == want["max_abs_err"]
    assert (outdir / "metrics_mean.csv").exists()
