"""Third-party pin of the Gaussian-window SSIM: fixtures computed with OpenCV.

scikit-image is not installed in this image, so oracle/distortion_oracle.py:ssim_gaussian_band restates
skimage.metrics.structural_similarity(gaussian_weights=True, sigma=1.5, use_sample_covariance=False) on
scipy.ndimage.  This script computes the SAME quantity with a filter implementation that shares nothing with scipy or
with this repository: cv2.GaussianBlur(img, (11, 11), 1.5) on float64 -- the formulation of OpenCV's own SSIM sample
("Similarity check (PSNR and SSIM)", getMSSIM: GaussianBlur(I, Size(11, 11), 1.5), C1 = (0.01 L)^2, C2 = (0.03 L)^2) --
followed by skimage's crop of (win_size - 1) / 2 = 5 border pixels before the mean.  Inside that crop every 11 x 11
window lies inside the image, so the border modes of the two libraries (OpenCV reflect-101, skimage/scipy 'reflect')
never enter, and getGaussianKernel(11, 1.5) is the same normalised exp(-k^2 / 4.5) as scipy's truncate-3.5 kernel.

    python oracle/make_golden_ssim_cv2.py        ->  tests/golden/ssimw_cv2.npz  (inputs + cv2 results)

Test infrastructure only (tests/test_oracle_golden.py pins the oracle on it, tests/test_gpu_midsize.py the kernels)."""
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
CASES = [  # name, dtype, (H, W), data range L, noise amplitude, seed
    ("u8_64x80", "uint8", (64, 80), 255.0, 6, 1),
    ("u16_150x203", "uint16", (150, 203), 4095.0, 40, 2),
    ("u16_12bit_97x131", "uint16", (97, 131), 65535.0, 900, 3),
    ("i16_75x77", "int16", (75, 77), 32767.0, 300, 4),
    ("u16_flat_40x40", "uint16", (40, 40), 4095.0, 0, 5),
]


def ssim_cv2(a: np.ndarray, b: np.ndarray, L: float) -> float:
    import cv2
    x, y = a.astype(np.float64), b.astype(np.float64)
    blur = lambda z: cv2.GaussianBlur(z, (11, 11), 1.5)      # noqa: E731  (sigmaY = sigmaX; default border never read below)
    ux, uy = blur(x), blur(y)
    vx, vy, vxy = blur(x * x) - ux * ux, blur(y * y) - uy * uy, blur(x * y) - ux * uy
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(s[5:-5, 5:-5].mean(dtype=np.float64))


def make_pair(dtype, shape, amp, seed):
    rng = np.random.default_rng(seed)
    H, W = shape
    info = np.iinfo(dtype)
    yy, xx = np.mgrid[0:H, 0:W]
    base = 0.45 * (np.sin(xx / 7.0) * np.cos(yy / 11.0) + 1.0) + 0.1 * rng.random((H, W))      # smooth structure + texture
    lo, hi = (0, min(info.max, 4095 if amp < 100 else info.max)) if info.min == 0 else (-6000, 6000)
    a = (lo + base / 1.1 * (hi - lo)).astype(np.int64)
    b = a + (rng.integers(-amp, amp + 1, (H, W)) if amp else 0)
    if amp == 0:
        a[:] = a[0, 0]; b = a.copy(); b[7, 9] += 3          # flat image: variances are exactly zero nearly everywhere
    return np.clip(a, info.min, info.max).astype(dtype), np.clip(b, info.min, info.max).astype(dtype)


def main():
    out = {}
    for name, dtype, shape, L, amp, seed in CASES:
        a, b = make_pair(dtype, shape, amp, seed)
        out[f"{name}__a"], out[f"{name}__b"] = a, b
        out[f"{name}__L"] = np.float64(L)
        out[f"{name}__ssim"] = np.float64(ssim_cv2(a, b, L))
        print(name, out[f"{name}__ssim"])
    np.savez_compressed(ROOT / "tests" / "golden" / "ssimw_cv2.npz", **out)


if __name__ == "__main__":
    main()
