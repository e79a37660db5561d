"""Timing of the validity-folding one-pass kernel (dm_fused_bip_scan: ONE read of an EnMAP int16 + nodata pair)
against dm_validity + dm_fused_bip (two reads), through the C ABI with fixed buffers.  Development tool."""
import ctypes as C
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200 import _lib
from image_compression_analysis_b200.engine import DevicePair, Partials, _lut_on_device, _ptr, _stream_ptr, workspace
B, H, W = 180, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(3)
ref = (torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4)
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
if len(sys.argv) > 2 and sys.argv[2] == "corner":      # a no-data corner (EnMAP tiles are cut from a rotated swath)
    yy = torch.arange(H, device="cuda").view(H, 1); xx = torch.arange(W, device="cuda").view(1, W)
    bad = (yy + xx) < int((2 * frac * H * W) ** 0.5)
else:                                                  # scattered single pixels (the adverse case: every warp meets one)
    bad = torch.rand((H, W), device="cuda", generator=g) < frac
ref[bad] = -32768
tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(-32768, 32767)
tst[bad] = -32768
pair = DevicePair(ref, tst, "int16", "bip", B, H, W, -32768, -32768)
L = _lib.lib()
plane = torch.empty(H * W, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
cp = pair.c_pair()
P = Partials.allocate(B, 0, ref.device, "int16")
ws = workspace(ref.device)
pl_g = torch.empty(H * W, dtype=torch.uint8, device="cuda"); pl_z = torch.empty_like(pl_g)
lg, lz = _lut_on_device(255, ref.device), _lut_on_device(32, ref.device)


def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def args(sam, err):
    if err:
        return (_ptr(P.sums), _ptr(P.imax), None, _ptr(lg), 255, _ptr(pl_g), _ptr(P.hist8_g), _ptr(lz), 32, _ptr(pl_z),
                _ptr(P.hist8_z), sam, _ptr(P.spec), _ptr(ws), _stream_ptr())
    return (_ptr(P.sums), _ptr(P.imax), None, None, 0, None, None, None, 0, None, None, sam, _ptr(P.spec), _ptr(ws), _stream_ptr())


print(f"invalid pixels: {frac:.0%} ({sys.argv[2] if len(sys.argv) > 2 else 'scattered'})")
for variant in (12, 23):
    L.dm_fused_bip_variant(variant)
    for wn, sam, err in (("stats", 0, False), ("stats+sam", 1, False), ("stats+sam+err8", 1, True)):
        a = args(sam, err)
        one = t(lambda: L.dm_fused_bip_scan(C.byref(cp), None, _ptr(plane), _ptr(cnt), *a))

        def two():
            L.dm_validity(C.byref(cp), None, _ptr(plane), _ptr(cnt), _stream_ptr())
            L.dm_fused_bip(C.byref(cp), _ptr(plane), *a)
        tw = t(two)
        msk = t(lambda: L.dm_fused_bip(C.byref(cp), _ptr(plane), *a))
        print(f"{variant} band warps  {wn:15s} one read {one:7.1f} us ({4*B*H*W/one/1e3:6.0f} GB/s)   validity + masked {tw:7.1f} us"
              f"   masked alone {msk:7.1f} us", flush=True)
L.dm_fused_bip_variant(0)
