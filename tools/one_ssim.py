"""One Gaussian-SSIM / Sobel-LMSE launch on a Case-A scene strip (for ncu captures)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 4, 4096, 4096
ref = torch.randint(0, 4096, (B, H, W), device="cuda", dtype=torch.int16, generator=g)
tst = (ref + torch.randint(-3, 4, (B, H, W), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
pair = DevicePair(ref, tst, "uint16", "bsq", B, H, W)
what = sys.argv[1] if len(sys.argv) > 1 else "ssim"
for _ in range(3):
    P = evaluate(pair, Want(stats=False, ssim_gauss=(what == "ssim"), lmse=(what == "lmse")), data_range=4095.0)
torch.cuda.synchronize()
print("ok", float(P.ssimw_sum[0]), float(P.lmse[0]))
