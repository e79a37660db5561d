"""B200-native reconstruction-distortion metrics (drop-in for the metric path of
Angela0110/Image-compression-analysis' tools/run_codec.py).  See DESIGN.md."""

from .metrics import (all_metrics_arrays, compute_metrics, compute_metrics_arrays,  # noqa: F401
                      compute_sam_sid_lmse_caseB, compute_sam_sid_lmse_caseB_arrays, effective_data_range,
                      effective_data_range_arrays, mse, psnr, sobel_mag, ssim_gaussian_arrays, ssim_global)

__all__ = ["all_metrics_arrays", "compute_metrics", "compute_metrics_arrays", "compute_sam_sid_lmse_caseB",
           "compute_sam_sid_lmse_caseB_arrays", "effective_data_range", "effective_data_range_arrays", "mse",
           "psnr", "sobel_mag", "ssim_gaussian_arrays", "ssim_global"]
