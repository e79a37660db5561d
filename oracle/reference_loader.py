"""Load the UNMODIFIED reference modules under the rasterio stub -- TEST INFRASTRUCTURE ONLY.

Works only where `/root/reference` (or $DM_REFERENCE_ROOT) is mounted, i.e. in
the build container.  It never runs on the GPU box: the `-m gpu` tests, smoke()
and bench.py use the committed fixtures in tests/golden/ and the numpy
restatement in oracle/distortion_oracle.py instead.

Used by oracle/make_golden.py (fixture generation) and by the `not gpu` tests
that pin the restatement against the real reference when it is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from pathlib import Path

from . import rasterio_stub

_CACHE: dict[str, object] = {}


def reference_root() -> Path | None:
    for cand in (os.environ.get("DM_REFERENCE_ROOT"), "/root/reference"):
        if cand and (Path(cand) / "tools" / "run_codec.py").exists():
            return Path(cand)
    return None


def available() -> bool:
    return reference_root() is not None


def _load(name: str, rel: str):
    if name in _CACHE:
        return _CACHE[name]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not mounted (set DM_REFERENCE_ROOT)")
    rasterio_stub.install()
    spec = importlib.util.spec_from_file_location(f"_dm_ref_{name}", root / rel)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    _CACHE[name] = mod
    return mod


def run_codec():
    """The reference's tools/run_codec.py as a module (metric functions at :55-137, :240-347)."""
    return _load("run_codec", "tools/run_codec.py")


def quicklooks():
    """The reference's tools/quicklooks.py as a module (write_error_max8 at :115-207)."""
    return _load("quicklooks", "tools/quicklooks.py")


def make_baseline_A():
    """tools/make_baseline_A.py (to_12in16 at :137-170)."""
    return _load("make_baseline_A", "tools/make_baseline_A.py")


def make_baseline_B():
    """tools/make_baseline_B.py (trunc_uint16 / write_truncated_copy :279-316, make_scene_error_map :324-419)."""
    return _load("make_baseline_B", "tools/make_baseline_B.py")


def ccsds121_wrap():
    """tools/codecs/ccsds121/ccsds121_wrap.py (raw interleave :44-64, diff1 :66-85)."""
    return _load("ccsds121_wrap", "tools/codecs/ccsds121/ccsds121_wrap.py")


def jpegls_wrap():
    """tools/codecs/jpegls/jpegls_wrap.py (_diff1_forward / _diff1_inverse :92-120).

    The wrapper does `from tools.common.proc_metrics import ...` at module top; this repo has its own
    `tools` package, so the reference's module is bound under that name for the duration of the import."""
    if "jpegls_wrap" in _CACHE:
        return _CACHE["jpegls_wrap"]
    import types
    saved = {k: sys.modules.get(k) for k in ("tools", "tools.common", "tools.common.proc_metrics")}
    path_before = list(sys.path)
    try:
        pm = _load("proc_metrics", "tools/common/proc_metrics.py")
        pkg, sub = types.ModuleType("tools"), types.ModuleType("tools.common")
        pkg.common, sub.proc_metrics = sub, pm
        sys.modules.update({"tools": pkg, "tools.common": sub, "tools.common.proc_metrics": pm})
        return _load("jpegls_wrap", "tools/codecs/jpegls/jpegls_wrap.py")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        sys.path[:] = path_before
