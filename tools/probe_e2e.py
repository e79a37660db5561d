"""End-to-end sweep from pinned host cubes: per-buffer H2D rates, then engine.evaluate_host_pairs with and without a
shared original (development probe; bench.py's e2e arm is the contract number)."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import Want, evaluate_host_pairs, bind_host_to_gpu_numa
print("numa node", bind_host_to_gpu_numa(0))
H, W, B = 1024, 1024, 180
g = torch.Generator(device="cuda").manual_seed(1)
host = []
for i in range(2):
    ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
    tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
    r = torch.empty((H, W, B), dtype=torch.int16).pin_memory(); d = torch.empty((H, W, B), dtype=torch.int16).pin_memory()
    r.copy_(ref); d.copy_(tst)
    host.append((r.view(torch.uint16), d.view(torch.uint16)))
torch.cuda.synchronize()
dst = torch.empty((H, W, B), dtype=torch.int16, device="cuda")
for i, pair in enumerate(host):
    for j, t in enumerate(pair):
        src = t.view(torch.int16)
        dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
        print(f"host[{i}][{j}] pinned={t.is_pinned()}  H2D {src.numel()*2/dt/1e9:6.1f} GB/s", flush=True)
want = Want(stats=True, sam=True)
for share in (False, True, False):
    for n in (2, 10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        src = (host[0] for _ in range(n)) if share else (host[i % 2] for i in range(n))
        k = sum(1 for _ in evaluate_host_pairs(src, want, layout="bip", share_ref=share))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"share_ref={share} n={n}: {dt*1e3:7.2f} ms per pair   {4*H*W*B/dt/1e9:6.1f} GB/s of pair bytes", flush=True)
