"""Timing of the REAL Case-B configuration: EnMAP int16 BIP cube with nodata (-32768) in both files
(validity plane + masked one-pass kernel).  Development tool."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
B, H, W = 180, 1024, 1024
g = torch.Generator(device="cuda").manual_seed(3)
ref = (torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4)
bad = torch.rand((H, W), device="cuda", generator=g) < 0.05
ref[bad] = -32768
tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(-32768, 32767)
tst[bad] = -32768
for name, pair in (("int16 + nodata (plane + masked kernel)", DevicePair(ref, tst, "int16", "bip", B, H, W, -32768, -32768)),
                   ("int16, no nodata", DevicePair(ref, tst, "int16", "bip", B, H, W)),
                   ("uint16 view, no nodata", DevicePair(ref, tst, "uint16", "bip", B, H, W))):
    for want, wn in ((Want(stats=True, sam=True), "stats+sam"), (Want(stats=True, sam=True, err8_caps=(255, 32)), "stats+sam+err8")):
        outs = [Partials.allocate(B, 0, ref.device, pair.np_dtype) for _ in range(12)]
        for P in outs[:3]:
            evaluate(pair, want, out=P)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for P in outs[3:]:
            evaluate(pair, want, out=P)
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 9
        print(f"{name:40s} {wn:16s} {ms*1e3:8.1f} us/pair  {4*B*H*W/ms/1e6:8.1f} GB/s", flush=True)

# breakdown of the nodata path: validity pre-pass alone, masked kernel alone (plane given)
import ctypes as C
from image_compression_analysis_b200 import _lib
from image_compression_analysis_b200.engine import _ptr, _stream_ptr
pair = DevicePair(ref, tst, "int16", "bip", B, H, W, -32768, -32768)
L = _lib.lib()
plane = torch.empty(H * W, dtype=torch.uint8, device="cuda")
cnt = torch.zeros(3, dtype=torch.int64, device="cuda")
cp = pair.c_pair()
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print(f"dm_validity alone                         {t(lambda: L.dm_validity(C.byref(cp), None, _ptr(plane), _ptr(cnt), _stream_ptr())):8.1f} us")
P = Partials.allocate(B, 0, ref.device, "int16")
print(f"masked one-pass kernel (plane given) sam  {t(lambda: evaluate(pair, Want(stats=True, sam=True), out=P, plane=plane)):8.1f} us")
print(f"masked one-pass kernel (plane given) err8 {t(lambda: evaluate(pair, Want(stats=True, sam=True, err8_caps=(255, 32)), out=P, plane=plane)):8.1f} us")
l0 = L.dm_launch_count()
Pn = Partials.allocate(B, 0, ref.device, "int16")
evaluate(pair, Want(stats=True, sam=True, err8_caps=(255, 32)), out=Pn)
torch.cuda.synchronize()
print("launches in one evaluate (nodata, err8):", L.dm_launch_count() - l0, "used_mask", Pn.used_mask)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        evaluate(pair, Want(stats=True, sam=True, err8_caps=(255, 32)), out=Pn)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=70))
