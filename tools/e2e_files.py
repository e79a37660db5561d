"""End-to-end time of ONE rep of the reference's metric calls on real GeoTIFF files (the
"[SKIP] Reusing reconstruction" re-run path of run_codec.py:489-531): write_error_max8 +
compute_metrics + compute_sam_sid_lmse_caseB on src.tif / recon.tif, through the drop-in functions
(built-in GeoTIFF reader, read-once ingest cache, B200 kernels).  Development tool."""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import image_compression_analysis_b200 as dm  # noqa: E402
from image_compression_analysis_b200 import geotiff, ingest, quicklooks as ql, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--bands", type=int, default=180)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--dir", default="/tmp/dm_e2e")
    ap.add_argument("--compress", default="NONE")
    args = ap.parse_args()
    d = Path(args.dir); d.mkdir(parents=True, exist_ok=True)
    ref, dec = synth.case_b_pair(seed=2, bands=args.bands, height=args.size, width=args.size, amp=3, layout="bsq")
    meta = dict(dtype="uint16", count=args.bands, width=args.size, height=args.size, tiled=True, blockxsize=512,
                blockysize=512, compress=args.compress, BIGTIFF="YES")
    with geotiff.open(d / "src.tif", "w", **meta) as dst:
        dst.write(ref)
    recs = []
    for r in range(args.reps):
        p = d / f"recon_{r}.tif"
        with geotiff.open(p, "w", **meta) as dst:
            dst.write(np.roll(dec, r, axis=1))
        recs.append(p)
    pair_mb = 2 * ref.nbytes / 1e6
    ingest.clear_cache()
    torch.cuda.synchronize()
    for r, p in enumerate(recs):
        t0 = time.perf_counter()
        ql.write_error_max8(d / "src.tif", p, d / f"recon_{r}", err_max_global=255, err_max_zoom=32)
        t1 = time.perf_counter()
        m = dm.compute_metrics(d / "src.tif", p)
        t2 = time.perf_counter()
        s = dm.compute_sam_sid_lmse_caseB(d / "src.tif", p)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        print(f"rep {r}: quicklooks {1e3*(t1-t0):7.1f} ms (incl. reading {'both files' if r == 0 else 'recon'}), "
              f"compute_metrics {1e3*(t2-t1):6.1f} ms, sam/sid/lmse {1e3*(t3-t2):6.1f} ms, total {1e3*(t3-t0):7.1f} ms "
              f"= {pair_mb/1e3/(t3-t0):6.2f} GB/s of pair bytes; psnr_global={m['psnr_global']:.4f} sam={s['sam_deg']:.5f}", flush=True)
    print("ingest:", ingest.STATS)


if __name__ == "__main__":
    main()
