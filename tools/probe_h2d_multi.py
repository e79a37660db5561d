"""Bare host-to-device ceiling with N ranks copying at once (torchrun --nproc-per-node N tools/probe_h2d_multi.py):
every rank copies its own pinned 377 MB cube to its GPU `reps` times, all ranks between the same two barriers.
Explains the end-to-end scaling of bench.py's e2e arm: that arm is the PCIe copy, so it cannot beat this."""
import os, sys, time
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import bind_host_to_gpu_numa
local = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
node = None if os.environ.get("DM_NO_NUMA_BIND") else bind_host_to_gpu_numa(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1024 * 1024 * 180
src = [torch.empty(n, dtype=torch.int16).pin_memory() for _ in range(2)]
for s in src:
    s.zero_()
dst = [torch.empty(n, dtype=torch.int16, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
def run(two_streams, reps=10):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        for k in range(2):
            with torch.cuda.stream(streams[k if two_streams else 0]):
                dst[k].copy_(src[k], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([2 * reps * n * 2 / dt / 1e9], dtype=torch.float64, device="cuda")
    lo = gbs.clone()
    if world > 1:
        dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return float(gbs.item()), float(lo.item())
run(False, 2)
for two in (False, True):
    tot, lo = run(two)
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"N={world} {'two streams' if two else 'one stream '} per rank: aggregate {tot:7.1f} GB/s, slowest rank {lo:6.1f} GB/s, numa node of rank 0: {node}", flush=True)
if world > 1:
    dist.destroy_process_group()
