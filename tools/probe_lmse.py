"""LMSE / SID timing on the Case-B cube (BIP and BSQ) + parity of the BIP Sobel path against the BSQ one."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
B, H, W = 180, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(3)
ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
bip = DevicePair(ref, tst, "uint16", "bip", B, H, W)
bsq = bip.as_bsq()
def t(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
from image_compression_analysis_b200._lib import lib
for lpp in (8, 16, 32):
    lib().dm_spectral_lanes_per_pixel(lpp)
    for wn, want in (("sid", Want(stats=False, sid=True)), ("sam+sid", Want(stats=False, sam=True, sid=True))):
        P = Partials.allocate(B, 0, ref.device, "uint16")
        us = t(lambda: evaluate(bip, want, out=P))
        P.zero_(); evaluate(bip, want, out=P); torch.cuda.synchronize()
        print(f"bip {wn:8s} lanes/pixel {lpp:2d} {us:9.1f} us  {4*B*H*W/us/1e3:8.1f} GB/s   spec {P.spec.cpu().numpy().tolist()}", flush=True)
lib().dm_spectral_lanes_per_pixel(0)
for name, pair in (("bip", bip), ("bsq", bsq)):
    for wn, want in (("lmse", Want(stats=False, lmse=True)), ("sid", Want(stats=False, sid=True)), ("sam+sid", Want(stats=False, sam=True, sid=True))):
        P = Partials.allocate(B, 0, ref.device, "uint16")
        us = t(lambda: evaluate(pair, want, out=P))
        print(f"{name} {wn:8s} {us:9.1f} us  {4*B*H*W/us/1e3:8.1f} GB/s", flush=True)
Pa = evaluate(bip, Want(stats=False, lmse=True, sid=True, sam=True)); Pb = evaluate(bsq, Want(stats=False, lmse=True, sid=True, sam=True))
torch.cuda.synchronize()
a, b = Pa.lmse.cpu().numpy(), Pb.lmse.cpu().numpy()
print("lmse bip vs bsq max rel diff", float(abs(a - b).max() / abs(b).max()), "sum", float(a.sum()), float(b.sum()))
print("spec bip", Pa.spec.cpu().numpy().tolist(), "bsq", Pb.spec.cpu().numpy().tolist())
