"""One launch of a Case-B BIP kernel (for ncu captures): python tools/one_caseb.py lmse|sid"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Want, evaluate
g = torch.Generator(device="cuda").manual_seed(1)
B, H, W = 180, 512, 1024
ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
pair = DevicePair(ref, tst, "uint16", "bip", B, H, W)
what = sys.argv[1] if len(sys.argv) > 1 else "lmse"
for _ in range(2):
    P = evaluate(pair, Want(stats=False, lmse=(what == "lmse"), sid=(what == "sid")))
torch.cuda.synchronize()
print("ok", float(P.lmse[0]), float(P.spec[1]))
