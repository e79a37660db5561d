"""One dm_fused_bsq launch sequence on a Case-A scene strip (for ncu captures)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 4, 10980, 10980
ref = torch.randint(0, 4096, (B, H, W), device="cuda", dtype=torch.int16, generator=g) * 16
tst = (ref + 16 * torch.randint(-3, 4, (B, H, W), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
pair = DevicePair(ref, tst, "uint16", "bsq", B, H, W)
for _ in range(3):
    P = evaluate(pair, Want(stats=True, err8_caps=(255, 32)))
torch.cuda.synchronize()
print("ok", int(P.sums[0]))
