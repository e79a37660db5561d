"""Run the reference's experiment driver unedited on the B200 metric path.

    python -m image_compression_analysis_b200.launcher /path/to/reference/tools  <run_codec.py args...>

Imports the reference's `run_codec` module, rebinds its metric functions
(`compute_metrics` tools/run_codec.py:240, `compute_sam_sid_lmse_caseB` :308, `mse/psnr/ssim_global`
:55-80) to the GPU implementations and makes `--quicklooks` default to this package's drop-in module
(the plugin hook at tools/run_codec.py:419-430), then calls `run_codec.main()`.  The reference needs
rasterio for its own file handling; this launcher does not change that.
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path


QUICKLOOKS_PLUGIN = Path(__file__).resolve().parent / "plugin" / "quicklooks.py"


def bind(run_codec_module) -> None:
    """Rebind the reference module's metric functions to the GPU implementations."""
    import image_compression_analysis_b200 as dm
    run_codec_module.compute_metrics = dm.compute_metrics
    run_codec_module.compute_sam_sid_lmse_caseB = dm.compute_sam_sid_lmse_caseB
    run_codec_module.mse = dm.mse
    run_codec_module.psnr = dm.psnr
    run_codec_module.ssim_global = dm.ssim_global
    run_codec_module.sobel_mag = dm.sobel_mag


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__, file=sys.stderr)
        return 2
    tools_dir, rest = Path(argv[0]), argv[1:]
    if not (tools_dir / "run_codec.py").exists():
        print(f"{tools_dir}/run_codec.py not found", file=sys.stderr)
        return 2
    sys.path.insert(0, str(tools_dir))
    run_codec = importlib.import_module("run_codec")
    bind(run_codec)
    if "--quicklooks" not in rest:
        rest = ["--quicklooks", str(QUICKLOOKS_PLUGIN), *rest]
    sys.argv = ["run_codec.py", *rest]
    run_codec.main()
    return 0


if __name__ == "__main__":
    sys.exit(main())
