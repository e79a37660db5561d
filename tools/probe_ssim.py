"""Gaussian-SSIM kernel timing: Case-A tile (rotated) and the 10980^2 scene."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
def mk(B, H, W, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ref = torch.randint(0, 2040, (B, H, W), device="cuda", dtype=torch.int16, generator=g) * 16
    tst = (ref + 16 * torch.randint(-3, 4, (B, H, W), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
    return DevicePair(ref, tst, "uint16", "bsq", B, H, W)
def t(fn, n=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
from image_compression_analysis_b200._lib import lib
for name, B, H, W in (("caseA tile", 4, 1024, 1024), ("scene", 4, 10980, 10980)):
    pair = mk(B, H, W, 1)
    res = {}
    for variant, label in ((3, "ring"), (2, "tiled")):
        lib().dm_ssim_variant(variant)
        P = Partials.allocate(B, 0, pair.ref.device, "uint16")
        us = t(lambda: evaluate(pair, Want(stats=False, ssim_gauss=True), out=P, data_range=4095.0))
        P.zero_()
        evaluate(pair, Want(stats=False, ssim_gauss=True), out=P, data_range=4095.0)
        h = P.to_host()
        res[variant] = h.ssimw_sum / h.ssimw_cnt
        print(f"{name:12s} ssim_gauss[{label:9s}] {us:9.1f} us  {4*B*H*W/us/1e3:8.1f} GB/s   ssimw = {res[variant]}", flush=True)
    lib().dm_ssim_variant(0)
    print(f"{name:12s} max |ring - tiled| / tiled = {float(abs(res[3] - res[2]).max() / abs(res[2]).max()):.3e}")
    del pair
    torch.cuda.empty_cache()
