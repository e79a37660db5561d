"""BASELINE.json configs 3 and 4 on N GPUs (development benchmark; bench.py is the contract benchmark).

    python tools/bench_configs.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 ... tools/bench_configs.py

config 3  full Sentinel-2 scene 10980 x 10980 x 4 uint16, ALL Case-A metrics (per-band and global statistics,
          both ERR8 planes, 256-bin error histograms, Gaussian-window SSIM), sharded by row strips with 8 halo
          rows (5 needed; 8 keeps the counted rows 16-byte aligned), one exchange of the partial vectors
config 4  Case-B rate sweep: 42 decoded 1024 x 1024 x 180 cubes against one original, ALL Case-B metrics
          (statistics + SAM in one pass, SID, Sobel-LMSE), sharded by PAIR (whole pairs per rank: no halo, no
          exchange until the single gather of the results at the end)
Synthetic cubes generated on the device; CUDA events; max over ranks; one JSON line on rank 0."""
import json, os, sys
from pathlib import Path
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
from image_compression_analysis_b200 import sharding

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- config 3 -------------------------------------------------------------------------------------------
B, H, W = 4, 10980, 10980
# 8 halo rows instead of the 5 the Gaussian window needs: the counted rows then start 8 rows = a multiple of 16
# bytes into the strip buffer for ANY row pitch, which the one-pass BSQ kernel wants (10980 * 2 B * 5 rows is not)
s = sharding.strips(H, world, halo=8)[rank]
g = torch.Generator(device=dev).manual_seed(7 + rank)
rows = s.buf1 - s.buf0
ref = torch.randint(0, 4096, (B, rows, W), device=dev, dtype=torch.int16, generator=g) * 16
tst = (ref + 16 * torch.randint(-3, 4, (B, rows, W), device=dev, dtype=torch.int16, generator=g)).clamp_(0, 32767)
full = DevicePair(ref, tst, "uint16", "bsq", B, rows, W, img_row0=s.buf0, img_rows=H)
c0, c1 = s.count_range
core = DevicePair(ref.view(-1)[c0 * W:], tst.view(-1)[c0 * W:], "uint16", "bsq", B, c1 - c0, W, None, None, s.row0, H,
                  band_stride=rows * W)
P = Partials.allocate(B, 256, dev, "uint16")


def scene_all():
    P.zero_()
    evaluate(core, Want(stats=True, err8_caps=(255, 32)), out=P)                     # one pass: statistics + both planes
    evaluate(core, Want(stats=True, hist_bins=256), out=Partials.allocate(B, 256, dev, "uint16"))   # per-band histograms
    evaluate(full, Want(stats=False, ssim_gauss=True), out=P, rows=(c0, c1), data_range=4095.0)
    P.allreduce_()


ms3 = timed(scene_all, 3)
scene_bytes = 4 * B * H * W
del ref, tst, full, core
torch.cuda.empty_cache()

# ---- config 4 -------------------------------------------------------------------------------------------
Bb, Hb, Wb, NPAIRS = 180, 1024, 1024, 42
g = torch.Generator(device=dev).manual_seed(11)
orig = torch.randint(0, 2500, (Hb, Wb, Bb), device=dev, dtype=torch.int16, generator=g) * 4
decs = [(orig + torch.randint(-a, a + 1, (Hb, Wb, Bb), device=dev, dtype=torch.int16, generator=g)).clamp_(0, 32767) for a in (1, 3, 9)]
mine = list(range(rank, NPAIRS, world))
run, outs = Partials.allocate_run(NPAIRS, Bb, 0, dev, "uint16")


def sweep():
    run.zero_()
    for i in mine:
        pair = DevicePair(orig, decs[i % 3], "uint16", "bip", Bb, Hb, Wb)
        evaluate(pair, Want(stats=True, sam=True), out=outs[i])
        evaluate(pair, Want(stats=False, sid=True, lmse=True), out=outs[i])
    if world > 1:
        # every pair was evaluated by exactly one rank and the other ranks' vectors are all-zero bits, so an int64
        # SUM over the whole run gathers sums, maxima and (bit patterns of) float sums alike
        dist.all_reduce(run, op=dist.ReduceOp.SUM)


ms4 = timed(sweep, 2)
if rank == 0:
    print(json.dumps({"n_gpus": world,
                      "config3_scene_all_metrics": {"ms": ms3, "GBps": scene_bytes / ms3 / 1e6, "pair_bytes": scene_bytes},
                      "config4_caseB_sweep_42_pairs_all_metrics": {"ms": ms4, "GBps": NPAIRS * 4 * Bb * Hb * Wb / ms4 / 1e6,
                                                                    "ms_per_pair_per_gpu": ms4 / max(1, len(mine))}}), flush=True)
if world > 1:
    dist.destroy_process_group()
