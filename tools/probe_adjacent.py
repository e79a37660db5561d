"""Timing of the kernels either side of the path (csrc/adjacent.cu) at BASELINE geometries, CUDA events,
inputs larger than L2 or rotated.  Development tool; numbers go to profiles/ and DESIGN.md."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200 import adjacent, finish
from image_compression_analysis_b200.engine import DevicePair


def t(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def line(name, secs, nbytes):
    print(f"{name:58s} {secs*1e6:9.1f} us  {nbytes/secs/1e9:8.1f} GB/s", flush=True)


g = torch.Generator(device="cuda").manual_seed(1)
# Case B cube, BSQ and BIP
B, H, W = 180, 1024, 1024
bsq = torch.randint(0, 2500, (B, H, W), device="cuda", dtype=torch.int16, generator=g) * 4
dec = bsq + torch.randint(-3, 4, (B, H, W), device="cuda", dtype=torch.int16, generator=g)
cube_bytes = bsq.numel() * 2
out = torch.empty_like(bsq)
line("requantize trunc k=2 (Case-B cube, read+write)", t(lambda: adjacent.requantize(bsq, "int16", "trunc", 2, -32768, out)), 2 * cube_bytes)
line("requantize round k=4", t(lambda: adjacent.requantize(bsq, "uint16", "round", 4, None, out)), 2 * cube_bytes)
line("diff1 forward modulo (Case-B cube BSQ, read+write)", t(lambda: adjacent.diff1(bsq, "int16", False, False, out)), 2 * cube_bytes)
line("diff1 inverse modulo", t(lambda: adjacent.diff1(bsq, "int16", True, False, out)), 2 * cube_bytes)
line("diff1 forward saturating", t(lambda: adjacent.diff1(bsq, "int16", False, True, out)), 2 * cube_bytes)
line("diff1 inverse saturating", t(lambda: adjacent.diff1(bsq, "int16", True, True, out)), 2 * cube_bytes)
pair = DevicePair(bsq, dec, "int16", "bsq", B, H, W)
for mode in ("mean", "rms", "count3", "max", "p95"):
    line(f"scene_error {mode} BSQ (pair bytes)", t(lambda: adjacent.scene_error(pair, None, mode, 2), n=5), 2 * cube_bytes)
bip = adjacent.interleave(bsq, "bsq", "bip", B, H, W)
dbip = adjacent.interleave(dec, "bsq", "bip", B, H, W)
pairb = DevicePair(bip, dbip, "int16", "bip", B, H, W)
for mode in ("mean", "rms", "max", "p95"):
    line(f"scene_error {mode} BIP (pair bytes)", t(lambda: adjacent.scene_error(pairb, None, mode, 2), n=5), 2 * cube_bytes)
for a, b_, src in (("bsq", "bip", bsq), ("bip", "bsq", bip), ("bsq", "bil", bsq), ("bip", "bil", bip)):
    line(f"interleave {a}->{b_} (read+write)", t(lambda: adjacent.interleave(src, a, b_, B, H, W), n=5), 2 * cube_bytes)
line("band_hist 3 bands of the BIP cube (3 bands' bytes)", t(lambda: adjacent.band_hist(bip, "int16", "bip", B, H, W, [47, 29, 10])), 3 * H * W * 2)
del bsq, dec, out, bip, dbip, pair, pairb
torch.cuda.empty_cache()
# full Sentinel-2 scene, 4 bands BSQ
Bs, Hs, Ws = 4, 10980, 10980
scene = torch.randint(0, 4096, (Bs, Hs, Ws), device="cuda", dtype=torch.int16, generator=g) * 16
line("band_hist 3 bands of the 10980^2 scene (12-in-16 data)", t(lambda: adjacent.band_hist(scene, "uint16", "bsq", Bs, Hs, Ws, [2, 1, 0]), n=5), 3 * Hs * Ws * 2)
rnd = torch.randint(-32768, 32768, (Bs, Hs, Ws), device="cuda", dtype=torch.int16, generator=g)
line("band_hist 3 bands, uniformly random 16-bit data (worst case)", t(lambda: adjacent.band_hist(rnd, "uint16", "bsq", Bs, Hs, Ws, [2, 1, 0]), n=3), 3 * Hs * Ws * 2)
luts = np.stack([finish.stretch8_lut(300.0, 9000.0, "uint16")] * 3)
line("lut_bands_u8 3 bands of the scene (2 B read + 1 B written / px)", t(lambda: adjacent.lut_bands_u8(scene, "uint16", "bsq", Bs, Hs, Ws, [2, 1, 0], luts), n=5), 3 * Hs * Ws * 3)
