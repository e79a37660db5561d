"""torchrun --nproc-per-node N tools/check_p2p.py: the NVLink peer-memory exchange (P2PRunCombiner) against the
NCCL one (RunCombiner) on rank-dependent partial vectors -- results must be bit-identical on every rank."""
import os, sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import Partials
from image_compression_analysis_b200.sharding import P2PRunCombiner, RunCombiner

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
B, n = 180, 37
ni, nm, nf = Partials.sizes(B, 0)
def make():
    run, outs = Partials.allocate_run(n, B, 0, torch.device("cuda", local), "uint16")
    g = torch.Generator(device="cuda").manual_seed(100 + rank)
    run[:, :ni + nm] = torch.randint(-2**40, 2**40, (n, ni + nm), device="cuda", generator=g)
    run[:, ni + nm:] = torch.rand((n, nf), device="cuda", dtype=torch.float64, generator=g).view(torch.int64)
    return run
a, b = make(), make()
assert torch.equal(a, b)
ca, cb = RunCombiner(a, B, 0, batch=8), P2PRunCombiner(b, B, 0, batch=8)
for i in range(n):
    ca.done(i); cb.done(i)
ca.finish(n); cb.finish(n)
torch.cuda.synchronize()
cb.check_status()
same = torch.equal(a, b)
# every rank must hold the same combined run
ref = a.clone(); dist.broadcast(ref, 0)
print(f"rank {rank}/{world}: p2p == nccl: {same}; same on all ranks: {torch.equal(ref, a)}", flush=True)
cb.close()
dist.destroy_process_group()
sys.exit(0 if same else 1)
