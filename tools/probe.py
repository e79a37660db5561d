"""Quick device-side timing probe of individual kernels (CUDA events, inputs > L2 or rotated).
Development tool; bench.py is the contract benchmark."""
import argparse
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200 import _lib  # noqa: E402
from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate  # noqa: E402


def make_pair(B, H, W, layout, dtype="uint16", amp=3, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = (B, H, W) if layout == "bsq" else (H, W, B)
    ref = torch.randint(0, 2500, shape, device="cuda", dtype=torch.int16, generator=g) * 4
    noise = torch.randint(-amp, amp + 1, shape, device="cuda", dtype=torch.int16, generator=g)
    tst = (ref + noise).clamp_(0, 32767)
    return DevicePair(ref, tst, dtype, layout, B, H, W)


def timeit(fn, iters=10, warm=3):
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
    ap.add_argument("--what", default="stats")
    args = ap.parse_args()
    cases = [("caseA tile bsq", 4, 1024, 1024, "bsq"), ("caseB cube bsq", 180, 1024, 1024, "bsq"),
             ("caseB cube bip", 180, 1024, 1024, "bip"), ("scene bsq", 4, 10980, 10980, "bsq")]
    for name, B, H, W, layout in cases:
        pairs = [make_pair(B, H, W, layout, seed=s) for s in range(2 if B * H * W > 5e7 else 16)]
        nbytes = 4 * B * H * W
        variants = {
            "moments": Want(stats=True),
            "no_moments": Want(stats=True, moments=False),
            "moments+hist256": Want(stats=True, hist_bins=256),
            "generic": Want(stats=True, generic_stats=True),
            "spectral sam": Want(stats=False, sam=True),
            "stats+sam (fused if bip)": Want(stats=True, sam=True),
            "stats+sam+err8 (fused if bip)": Want(stats=True, sam=True, err8_caps=(255, 32)),
            "spectral errmax+err8": Want(stats=False, err8_caps=(255, 32)),
            "stats+err8 (fused if <=4 bands bsq)": Want(stats=True, err8_caps=(255, 32)),
            "stats+err8 two passes": Want(stats=True, err8_caps=(255, 32), fused=False),
        }
        if args.what == "all":
            variants["spectral sam+sid"] = Want(stats=False, sam=True, sid=True)
            variants["lmse"] = Want(stats=False, lmse=True)
            if B <= 8:
                variants["ssim_gauss"] = Want(stats=False, ssim_gauss=True)
        for vn, want in variants.items():
            if vn == "generic" and B * H * W > 3e8:
                continue
            outs = [Partials.allocate(B, want.hist_bins, pairs[0].ref.device, "uint16") for _ in pairs]
            state = {"i": 0}

            def fn():
                i = state["i"] % len(pairs)
                state["i"] += 1
                evaluate(pairs[i], want, out=outs[i], data_range=4095.0)
            med, best = timeit(fn, warm=max(3, len(pairs)))      # every output vector allocates its planes once: keep cudaMalloc out of the timed calls
            print(f"{name:16s} {vn:22s} median {med*1e3:9.1f} us  best {best*1e3:9.1f} us   "
                  f"{nbytes/med/1e6:8.1f} GB/s (median)  {nbytes/best/1e6:8.1f} GB/s (best)", flush=True)
        del pairs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
