"""A few launches of dm_fused_bip_scan on an EnMAP-like int16 + nodata cube pair (for ncu captures)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
B, H, W = 180, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(3)
ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
yy = torch.arange(H, device="cuda").view(H, 1); xx = torch.arange(W, device="cuda").view(1, W)
corner = (yy + xx) < int((2 * 0.05 * H * W) ** 0.5)
ref[corner] = -32768
tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(-32768, 32767)
tst[corner] = -32768
pair = DevicePair(ref, tst, "int16", "bip", B, H, W, -32768, -32768)
outs = [Partials.allocate(B, 0, ref.device, "int16") for _ in range(4)]
for P in outs:
    evaluate(pair, Want(stats=True, sam=True), out=P)
torch.cuda.synchronize()
print("ok", outs[0].counts.tolist(), outs[0].spec.tolist())
