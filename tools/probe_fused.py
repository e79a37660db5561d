"""Timing / A-B probe of dm_fused_bip on the bench workload (Case B cube, 1024x1024x180 BIP).

Needs the EXPERIMENT build of the library (the shipping build reads no environment variable):
    DM_DEBUG_HOOKS=1 python image_compression_analysis_b200/csrc/build.py --force
and afterwards `python image_compression_analysis_b200/csrc/build.py --force` to restore the shipping build.
In that build DM_FUSED_DEBUG is read by the library at every launch:
    0 normal   1 producer skips the copies (compute only)   2 band group idle   4 pixel group idle
    8 force the generic (runtime-geometry) kernel
Development tool; bench.py is the contract benchmark."""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate  # noqa: E402


def make_pair(B, H, W, seed=0, amp=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
    noise = torch.randint(-amp, amp + 1, (H, W, B), device="cuda", dtype=torch.int16, generator=g)
    tst = (ref + noise).clamp_(0, 32767)
    return DevicePair(ref, tst, "uint16", "bip", B, H, W)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="0,8,1,2,4,3,5,6")
    ap.add_argument("--err8", action="store_true")
    args = ap.parse_args()
    B, H, W = 180, 1024, 1024
    pairs = [make_pair(B, H, W, seed=s) for s in range(3)]
    nbytes = 4 * B * H * W
    want = Want(stats=True, sam=True, err8_caps=(255, 32) if args.err8 else (None, None))
    outs = [Partials.allocate(B, 0, pairs[0].ref.device, "uint16") for _ in pairs]
    # cross-check: specialised vs generic kernel on the same pair
    res = {}
    for mode in ("0", "8"):
        os.environ["DM_FUSED_DEBUG"] = mode
        P = Partials.allocate(B, 0, pairs[0].ref.device, "uint16")
        evaluate(pairs[0], want, out=P)
        torch.cuda.synchronize()
        res[mode] = P.flat.clone()
    same_int = torch.equal(res["0"][: -(3 + 3 * B)], res["8"][: -(3 + 3 * B)])
    f0, f8 = res["0"][-(3 + 3 * B):].view(torch.float64)[:3], res["8"][-(3 + 3 * B):].view(torch.float64)[:3]
    print("specialised == generic (integers):", same_int, " spec:", f0.tolist(), f8.tolist(), flush=True)
    K = 12
    for mode in args.modes.split(","):
        os.environ["DM_FUSED_DEBUG"] = mode
        # K back-to-back steps captured in one CUDA graph: replay time / K is pure device time
        # (no host launch gaps); the eager loop next to it shows what the host adds
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in range(K):
                evaluate(pairs[k % len(pairs)], want, out=outs[k % len(pairs)])
        med, best = timeit(graph.replay, iters=10)
        state = {"i": 0}

        def fn():
            for k in range(K):
                evaluate(pairs[k % len(pairs)], want, out=outs[k % len(pairs)])
        emed, ebest = timeit(fn, iters=5)
        print(f"DM_FUSED_DEBUG={mode:3s} graph: median {med/K*1e3:8.1f} us  best {best/K*1e3:8.1f} us   "
              f"{nbytes*K/med/1e6:8.1f} GB/s (median) {nbytes*K/best/1e6:8.1f} GB/s (best)   eager loop: {emed/K*1e3:8.1f} us/step", flush=True)
    os.environ["DM_FUSED_DEBUG"] = "0"


if __name__ == "__main__":
    main()
