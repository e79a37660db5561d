"""The file to hand to `run_codec.py --quicklooks <path>`.

run_codec.py puts this file's DIRECTORY first on sys.path and runs `import quicklooks`
(tools/run_codec.py:419-430), i.e. it loads the plugin as a TOP-LEVEL module; any failure makes it fall
back silently to the reference's own CPU quicklooks.py.  This directory therefore holds nothing but this
file (nothing else becomes importable by a bare name), and the file imports the package by its absolute
name after making the repository root importable.  It also runs as a script with the reference module's
command line (tools/quicklooks.py:213-239, used by tools/make_baseline_A.py:189-198).
"""
import sys
from pathlib import Path

_ROOT = str(Path(__file__).resolve().parents[2])
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

from image_compression_analysis_b200.quicklooks import (  # noqa: E402,F401
    RGB_ORDER, _valid_mask_from_ds, error_max8_arrays, main, stretch_params_from_baseline, write_error_max8,
    write_rgb_8bit)

B200_NATIVE = True      # lets a caller (and the tests) tell this module from the reference's quicklooks.py

if __name__ == "__main__":
    raise SystemExit(main())
