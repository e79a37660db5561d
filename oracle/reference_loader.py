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
