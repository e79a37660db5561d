"""One launch sequence of dm_fused_bip on the bench workload (for ncu captures)."""
import os, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import Partials, Want, evaluate
from tools.probe_fused import make_pair
B, H, W = 180, 1024, 1024
pairs = [make_pair(B, H, W, seed=s) for s in range(2)]
want = Want(stats=True, sam=True, err8_caps=(255, 32) if os.environ.get("DM_PROBE_ERR8") else (None, None))
outs = [Partials.allocate(B, 0, pairs[0].ref.device, "uint16") for _ in pairs]
for k in range(4):
    evaluate(pairs[k % 2], want, out=outs[k % 2])
torch.cuda.synchronize()
print("ok", outs[0].spec.tolist())
