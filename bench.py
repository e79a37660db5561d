#!/usr/bin/env python3
"""bench.py -- distortion-metric throughput (GB/s of image-pair bytes) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): one Case-B EnMAP
cube pair, 1024 x 1024 x 180 uint16, BIP, per GPU; a step evaluates everything compute_metrics
returns (per-band and global PSNR / SSIM / MAXAE, data-range scan) plus SAM.  With N GPUs the image
is N x 1024 rows sharded by row strips (weak scaling: per-GPU work fixed) and the step ends with the
allreduce of the integer / float64 partials, the path's only exchange.

One JSON line on stdout (rank 0):
  value      whole-job GB/s with the cubes resident in HBM (CUDA events, max over ranks); launches are
             prepared once (engine.PreparedFused), every step writes its own partial vector of a run
  e2e        the same metric through the public API from pinned HOST buffers (H2D + D2H inside);
             e2e.shared_original: the same sweep when the decoded cubes share ONE original (uploaded once)
  roofline   the dominant kernel against the measured HBM copy peak (MEASURED_PEAKS.json); traffic = DRAM bytes
             of the committed ncu capture of the same kernel instance (profiles/traffic.json says which)
  configs    the other BASELINE.json configurations in the same run: C1 Case-A tile statistics (batched launch +
             single-pair latency), C3 Case-A tile Gaussian SSIM + ERR8 planes, C4 the 10980 x 10980 x 4 scene with
             every Case-A metric STRONG-scaled over the ranks, C5 the 42-cube Case-B sweep with every Case-B metric
             sharded by pair -- each with time, GB/s of algorithmic bytes and its roofline fraction
  cpu_baseline  the reference's CPU path (the unmodified reference when its tree is mounted, else the numpy port
             of oracle/) on a bounded sample, one core, rank 0, N=1 only
`--impl reference` times that CPU path with all host cores instead (rank 0 only; median step).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "distortion-metric throughput (original+decoded image-pair bytes)"
UNIT = "GB/s"
BANDS, ROWS, WIDTH = 180, 1024, 1024
PAIR_BYTES = 2 * 2 * BANDS * ROWS * WIDTH            # 754 974 720: SURVEY.md 8d algorithmic bytes
COMBINE_BATCH = 64                                   # pairs per multi-GPU exchange (latency bound, 31.5 KB per pair; a sweep needs its results at its end)
WORKLOAD = "Case B EnMAP 1024x1024x180 uint16 BIP cube pair per GPU: compute_metrics (per-band+global PSNR/SSIM/MAXAE) + SAM"


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    try:
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampling (pynvml, in a thread, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.001)

    def start(self):
        if self._nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the numpy port of the reference (oracle/), the ONE place the bench executes oracle code
# ------------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_sample(rows: int, seed: int):
    key = (rows, seed)
    if key not in _CPU_CACHE:
        from image_compression_analysis_b200 import synth
        _CPU_CACHE.clear()
        _CPU_CACHE[key] = synth.case_b_pair(seed=seed, bands=BANDS, height=rows, width=WIDTH, amp=3, layout="bsq")
    return _CPU_CACHE[key]


def _reference_modules():
    """The UNMODIFIED reference (tools/run_codec.py) under the in-memory rasterio stand-in when its tree is mounted:
    $DM_REFERENCE_ROOT -> baseline/_ref -> /root/reference, in that order.  None on the GPU
    box, where only the repository travels: the numpy port (pinned bit-for-bit against the reference by
    tests/test_oracle_vs_reference.py) is timed instead and the line says kind = "port"."""
    for cand in (os.environ.get("DM_REFERENCE_ROOT"), str(ROOT / "baseline" / "_ref"), "/root/reference"):
        if cand and (Path(cand) / "tools" / "run_codec.py").exists():
            os.environ["DM_REFERENCE_ROOT"] = cand
            try:
                from oracle import rasterio_stub, reference_loader
                return reference_loader.run_codec(), rasterio_stub, cand
            except Exception as e:      # noqa: BLE001
                print(f"[bench] reference at {cand} did not load ({e}); timing the port", file=sys.stderr)
                return None
    return None


def _cpu_work(args):
    """One worker: compute_metrics + SAM on its own strip.  compute_metrics is the reference's own function
    (tools/run_codec.py:240-304, unmodified, through the rasterio stand-in) when the reference tree is mounted, else
    the numpy port; SAM is always the port's restatement of run_codec.py:312-332 (the reference only offers SAM
    together with SID and LMSE, which the GPU step it is compared with does not compute).  The strip is generated
    once per process (first call) and is not part of the timed calls after that."""
    rows, seed, reps = args
    from oracle import distortion_oracle as orc
    ref, dec = _cpu_sample(rows, seed)
    mods = _CPU_CACHE.get("mods", False)
    if mods is False:
        mods = _CPU_CACHE["mods"] = _reference_modules()
        if mods is not None:
            mods[1].register("mem_ref.tif", ref)
            mods[1].register("mem_dec.tif", dec)
    t0 = time.perf_counter()
    for _ in range(reps):
        if mods is not None:
            mods[0].compute_metrics("mem_ref.tif", "mem_dec.tif", None)
        else:
            orc.compute_metrics(ref, dec, extras=False)
        orc.sam_caseB(ref, dec)
    return time.perf_counter() - t0, reps * 2 * ref.nbytes, "reference" if mods is not None else "port"


def cpu_baseline_single(rows: int = 512, reps: int = 3):
    """Scalar (1 core) run on a `rows` x 1024 x 180 strip: median of `reps` passes."""
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(k, "1")
    _cpu_sample(rows, 2)
    passes = [_cpu_work((rows, 2, 1)) for _ in range(max(1, reps))]
    dt = sorted(p[0] for p in passes)[len(passes) // 2]
    nbytes, kind = passes[0][1], passes[0][2]
    what = ("the reference's own compute_metrics (tools/run_codec.py:240-304, unmodified, rasterio stand-in) + the port's sam_caseB"
            if kind == "reference" else "oracle/distortion_oracle.py compute_metrics + sam_caseB (numpy port, pinned bit-for-bit)")
    return {"value": nbytes / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"median of {len(passes)} passes over a ({rows} rows x {WIDTH} x {BANDS} bands) strip of the workload, {what}, "
                      f"single thread, {dt:.1f} s per pass"}


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation of the path with every host core, one strip per
    process (the unmodified reference when its tree is mounted, else the numpy port: see _reference_modules)."""
    import multiprocessing as mp
    rank = _env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"                     # numpy elementwise code is single-threaded; one process per core
    workers = max(1, min(cores, 64))
    rows = 16                                   # per worker per step: 16 x 1024 x 180 -> 11.8 MB per pair
    ctx = mp.get_context("fork")
    times = []
    kind = "port"
    steps = max(3, args.steps)                  # at least three timed steps: the value is their MEDIAN
    with ctx.Pool(workers) as pool:
        for it in range(max(1, args.warmup) + steps):
            res = pool.map(_cpu_work, [(rows, 100, 1) for _ in range(workers)], chunksize=1)
            kind = res[0][2]
            if it >= max(1, args.warmup):
                # step time = the slowest worker's own metric time (its strip is cached per process,
                # so synthetic-input generation never enters the number)
                times.append((max(r[0] for r in res), sum(r[1] for r in res)))
    times.sort(key=lambda tb: tb[0])
    med_t, med_b = times[len(times) // 2]
    value = med_b / med_t / 1e9
    what = ("the reference's own compute_metrics (tools/run_codec.py:240-304, unmodified, under oracle/rasterio_stub.py) + the port's "
            "sam_caseB (run_codec.py:312-332)" if kind == "reference" else
            "oracle/distortion_oracle.py compute_metrics + sam_caseB (numpy port of run_codec.py:240-332, pinned bit-for-bit; the reference "
            "tree is not mounted on this box)")
    sample = (f"per step {workers} processes x one ({rows} rows x {WIDTH} x {BANDS} bands) strip each of the workload; {what}; "
              f"step time = slowest worker; value = median of {len(times)} steps (min {times[0][0]*1e3:.0f} ms, max {times[-1][0]*1e3:.0f} ms)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
        "warmup": args.warmup, "ms_per_step": 1e3 * med_t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_rows_per_worker": rows, "workers": workers},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def make_device_pairs(torch, n_pairs: int, seed: int):
    """Synthetic EnMAP-like BIP pairs generated on the device (14-in-16 reference, +-3 DN decode)."""
    from image_compression_analysis_b200.engine import DevicePair
    pairs = []
    for i in range(n_pairs):
        g = torch.Generator(device="cuda").manual_seed(seed * 1000 + i)
        ref = torch.randint(0, 2500, (ROWS, WIDTH, BANDS), device="cuda", dtype=torch.int16, generator=g) * 4
        noise = torch.randint(-3, 4, (ROWS, WIDTH, BANDS), device="cuda", dtype=torch.int16, generator=g)
        tst = (ref + noise).clamp_(0, 32767)
        pairs.append(DevicePair(ref, tst, "uint16", "bip", BANDS, ROWS, WIDTH))
    return pairs



# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (configs[0], [2], [3], [4]) -- same run, same JSON line, key "configs"
# ------------------------------------------------------------------------------------------------
FP64_LANES_PER_CLK_SM = 64          # DFMA lanes per clock and SM (measured 61.5-62.7: profiles/r02_ubench_fp64.txt)


def run_configs(torch, dist, world, rank, local, peak, clocks_mhz):
    """C1 Case-A tile statistics (batched + single-pair latency), C3 Case-A tile Gaussian SSIM + ERR8 planes,
    C4 the full 10980 x 10980 x 4 scene with every Case-A metric STRONG-scaled over the ranks by row strips,
    C5 the 42-pair Case-B rate sweep with every Case-B metric sharded by pair.  CUDA events, max over ranks.
    Algorithmic bytes as SURVEY.md 8d: 4 B per sample pair, the pair counted ONCE however many kernels read it."""
    import numpy as np
    from image_compression_analysis_b200 import _lib, finish, sharding
    from image_compression_analysis_b200.engine import (DevicePair, Partials, PreparedCaseAAll, PreparedFused, PreparedStats,
                                                        PreparedStatsBatch, Want, evaluate)
    dev = torch.device("cuda", local)
    L = _lib.lib()
    out = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.dm_launch_count()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / reps), (L.dm_launch_count() - l0) // reps

    fp64_peak = 148 * FP64_LANES_PER_CLK_SM * (clocks_mhz or 1965.0) * 1e6      # DFMA lane-ops per second at this clock

    def roof(bytes_, ms, bound, fp64_ops=None):
        gbps = bytes_ / (ms * 1e-3) / 1e9
        r = {"bound": bound, "achieved": gbps, "peak": peak, "unit": UNIT, "frac": gbps / (world * peak)}
        if fp64_ops is not None:
            r["fp64"] = {"ops_per_launch_set": fp64_ops, "achieved_gops": fp64_ops / (ms * 1e-3) / 1e9,
                         "peak_gops": world * fp64_peak / 1e9, "frac": fp64_ops / (ms * 1e-3) / (world * fp64_peak),
                         "note": "FP64-pipe operations this formulation needs (DFMA = 1) against 148 SMs x 64 lanes x the sampled SM clock"}
        return r

    # ---- C1 / C3: Case-A tiles, 32 distinct pairs per GPU (537 MB > L2), weak scaling -----------------------
    B, H, W, NT = 4, 1024, 1024, 32
    tile_bytes = 4 * B * H * W
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    tiles = []
    for i in range(NT):
        ref = torch.randint(0, 2040, (B, H, W), device=dev, dtype=torch.int16, generator=g) * 16      # 12-in-16; + 48 stays below 2^15
        tst = (ref + 16 * torch.randint(-3, 4, (B, H, W), device=dev, dtype=torch.int16, generator=g)).clamp_(0, 32767)
        tiles.append(DevicePair(ref, tst, "uint16", "bsq", B, H, W))
    run, outs = Partials.allocate_run(NT, B, 0, dev, "uint16")
    batch = PreparedStatsBatch(tiles, outs)
    batch.launch()
    torch.cuda.synchronize()
    h0 = outs[5].to_host()
    chk = finish.finish_compute_metrics(_lib.DM_U16, h0.sums, h0.maxs)
    assert int(h0.sums[0, 0]) == H * W and chk["max_abs_err"] == 48 and chk["lossless"] == 0, chk
    ms, nl = timed(batch.launch, 20)
    # pair by pair (one launch per tile, prepared arguments) over the same rotation
    singles = [PreparedStats(t, o) for t, o in zip(tiles, outs)]

    def one_by_one():
        for sp in singles:
            sp.launch()
    ms_1, _ = timed(one_by_one, 5)
    # latency of ONE pair: device time of an isolated launch, and host wall time call -> results on the host
    lat_dev = []
    for i in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200_000)
        a.record(); singles[i % NT].launch(); b.record(); b.synchronize()
        lat_dev.append(a.elapsed_time(b) * 1e3)
    # the same measurement for the same kernel on a 64 x 64 x 4 tile pair (64 KB): the floor of the method (launch, ramp-up,
    # flush and the two event records), which the 16.8 MB tile's figure has to be read against
    tiny_r = torch.randint(0, 2040, (B, 64, 64), device=dev, dtype=torch.int16, generator=g) * 16
    tiny = PreparedStats(DevicePair(tiny_r, tiny_r.clone(), "uint16", "bsq", B, 64, 64), Partials.allocate(B, 0, dev, "uint16"))
    lat_floor = []
    for i in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200_000)
        a.record(); tiny.launch(); b.record(); b.synchronize()
        lat_floor.append(a.elapsed_time(b) * 1e3)
    lat_floor.sort()
    lat_host = []
    pin = torch.empty(outs[0].flat.numel(), dtype=torch.int64).pin_memory()
    for i in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        singles[i % NT].launch()
        pin.copy_(outs[i % NT].flat, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        lat_host.append((time.perf_counter() - t0) * 1e6)
    lat_dev.sort(); lat_host.sort()
    out["C1_caseA_tile_stats"] = {
        "workload": "configs[0]: Case A Sentinel-2 1024x1024x4 uint16 BSQ tile pairs, compute_metrics statistics (PSNR/MSE/MAXAE/SSIM moments/data range)",
        "pairs_per_gpu": NT, "l2_policy": f"{NT} distinct 16.8 MB pairs rotated ({NT * tile_bytes / 1e6:.0f} MB per GPU)", "scaling": "weak",
        "batched": {"api": "engine.PreparedStatsBatch (dm_fused_stats_batch: one launch for all pairs)", "ms_per_batch": ms,
                    "us_per_pair": ms * 1e3 / NT, "GBps": world * NT * tile_bytes / ms / 1e6, "launches_per_batch": int(nl)},
        "pair_by_pair": {"api": "engine.PreparedStats (one dm_fused_stats launch per pair)", "us_per_pair": ms_1 * 1e3 / NT,
                         "GBps": world * NT * tile_bytes / ms_1 / 1e6},
        "single_pair_latency_us": {"device_isolated_launch_median": lat_dev[len(lat_dev) // 2],
                                   "device_isolated_launch_floor_64x64_tile": lat_floor[len(lat_floor) // 2],
                                   "host_call_to_result_median": lat_host[len(lat_host) // 2],
                                   "note": "device: CUDA events around one launch behind a spin kernel; host: perf_counter around "
                                           "launch + read-back of the partial vector into pinned memory + stream sync"},
        "roofline": roof(world * NT * tile_bytes, ms, "hbm"),
    }
    # C3: Gaussian SSIM + both ERR8 planes per tile
    Pc3 = [Partials.allocate(B, 256, dev, "uint16") for _ in range(NT)]
    # pairs rotate over four CUDA streams (a prepared launch carries the stream it was built on, with that stream's
    # workspace and scratch): the tail of one pair's SSIM kernel -- 2 432 tiles on 296 blocks -- overlaps the next pairs' start.
    # Measured per tile pair with 1 / 2 / 3 / 4 streams: 71.1 / 56.9 / 53.3 / 52.6 us (DM_ALT_STREAMS=n for A/B runs)
    alt = [torch.cuda.Stream(device=dev) for _ in range(max(1, int(os.environ.get("DM_ALT_STREAMS", "4"))))]
    c3 = []
    for k, (t, P) in enumerate(zip(tiles, Pc3)):
        with torch.cuda.stream(alt[k % len(alt)]):
            c3.append(PreparedCaseAAll(t, t, (0, H), P, 4095.0, hist_bins=0))

    def fork_join(body):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for st in alt:
            st.wait_event(ev)
        body()
        for st in alt:
            e = torch.cuda.Event()
            e.record(st)
            cur.wait_event(e)

    def c3_all():
        fork_join(lambda: [c.launch() for c in c3])
    c3_all()
    torch.cuda.synchronize()
    hs = Pc3[3].to_host()
    sw = finish.finish_ssim_gauss(hs.ssimw_sum, hs.ssimw_cnt)
    assert int(hs.ssimw_cnt[0]) == (H - 10) * (W - 10) and -1.0 < sw["ssimw_b1"] < 1.0 and int(hs.hist8_g.sum()) == H * W, sw
    for P in Pc3:
        P.zero_()
    ms3, nl3 = timed(c3_all, 5)
    ssim_ops = 104.0 * B * H * W * NT                   # 88 filter DFMA + 16 formula per band pixel
    out["C3_caseA_tile_ssim_err8"] = {
        "workload": "configs[2]: Case A 1024x1024x4 tile pairs: per-band statistics + ERR8 quicklook planes at caps 255 and 32 (one pass, dm_fused_bsq) "
                    "+ per-band Gaussian-window SSIM (dm_ssim_gauss)",
        "pairs_per_gpu": NT, "scaling": "weak", "us_per_pair": ms3 * 1e3 / NT, "GBps": world * NT * tile_bytes / ms3 / 1e6,
        "launches_per_pair": int(nl3) // NT, "streams": len(alt), "roofline": roof(world * NT * tile_bytes, ms3, "fp64", world * ssim_ops),
    }
    del tiles, run, outs, batch, singles, Pc3, c3
    torch.cuda.empty_cache()

    # ---- C4: one 10980 x 10980 x 4 scene, ALL Case-A metrics, strong-scaled by row strips --------------------
    B, H, W = 4, 10980, 10980
    scene_bytes = 4 * B * H * W
    # 8 halo rows (5 needed): the counted rows then start a multiple of 16 bytes into the strip for any row pitch
    # strips start on even rows: with the scene's 21 960-byte row pitch every band of a strip buffer then starts on a
    # 16-byte boundary, which the one-pass BSQ kernel needs
    s = sharding.strips(H, world, halo=8, align=2)[rank]
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    rows = s.buf1 - s.buf0
    ref = torch.randint(0, 2040, (B, rows, W), device=dev, dtype=torch.int16, generator=g) * 16
    tst = (ref + 16 * torch.randint(-3, 4, (B, rows, W), device=dev, dtype=torch.int16, generator=g)).clamp_(0, 32767)
    full = DevicePair(ref, tst, "uint16", "bsq", B, rows, W, img_row0=s.buf0, img_rows=H)
    c0, c1 = s.count_range
    core = DevicePair(ref.view(-1)[c0 * W:], tst.view(-1)[c0 * W:], "uint16", "bsq", B, c1 - c0, W, None, None, s.row0, H,
                      band_stride=rows * W)
    NREC = 16                                             # every repetition writes its own zeroed partial vector
    run4, outs4 = Partials.allocate_run(NREC, B, 256, dev, "uint16")
    # the two HBM-bound passes run on a side stream in the shadow of the FP64-bound SSIM kernel (5.72 -> 5.45 ms on one GPU)
    side4 = torch.cuda.Stream(device=dev)
    prep4 = [PreparedCaseAAll(core, full, (c0, c1), P, 4095.0, side_stream=side4) for P in outs4]
    comb = None
    exchange4 = "none"
    if world > 1:
        try:
            comb = sharding.P2PRunCombiner(run4, B, 256, batch=1, timeout_s=10.0)
            exchange4 = "nvlink peer memory (dm_p2p_push / dm_p2p_combine), one exchange per scene"
        except Exception as e:      # noqa: BLE001
            print(f"[bench] rank {rank}: C4 P2P exchange unavailable ({e}); NCCL", file=sys.stderr)
            comb = sharding.RunCombiner(run4, B, 256, batch=1)
            exchange4 = "NCCL all-gather + dm_combine_partials, one exchange per scene"
    state = {"i": 0}

    def scene():
        i = state["i"]
        prep4[i].launch()
        if comb is not None:
            comb.done(i)
        state["i"] = i + 1

    # warm-up 3 + timed 10 = 13 records of the 16
    for _ in range(3):
        scene()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = L.dm_launch_count()
    e0.record()
    for _ in range(10):
        scene()
    if comb is not None:
        comb.finish(state["i"])
    e1.record()
    barrier()
    ms4 = max_over_ranks(e0.elapsed_time(e1) / 10)
    nl4 = (L.dm_launch_count() - l0) // 10
    if isinstance(comb, sharding.P2PRunCombiner):
        comb.check_status()
    # one scene at a time (exchange included, nothing overlapped): the latency a single call sees
    lat4 = []
    for _ in range(3):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        scene()
        if comb is not None:
            comb.finish(state["i"])
        b.record()
        barrier()
        lat4.append(max_over_ranks(a.elapsed_time(b)))
    h4 = outs4[5].to_host()
    torch.cuda.synchronize()
    m4 = finish.finish_compute_metrics(_lib.DM_U16, h4.sums, h4.maxs, h4.hist, extras=True)
    sw4 = finish.finish_ssim_gauss(h4.ssimw_sum, h4.ssimw_cnt)
    assert int(h4.sums[0, 0]) == H * W and m4["max_abs_err"] == 48 and int(h4.hist8_g.sum()) == H * W \
        and int(h4.ssimw_cnt[0]) == (H - 10) * (W - 10) and int(h4.hist[0].sum()) == H * W, (m4["max_abs_err"], sw4)
    out["C4_scene_all_caseA_metrics"] = {
        "workload": "configs[3]: ONE Sentinel-2 scene 10980x10980x4 uint16 BSQ, all Case-A metrics: per-band+global statistics and both ERR8 planes "
                    "(one pass), per-band 256-bin |d| histograms, per-band Gaussian-window SSIM",
        "scaling": "strong", "sharding": f"row strips over {world} GPU(s), 8 halo rows, partial vectors combined once per scene",
        "exchange": exchange4, "ms_per_scene": ms4, "ms_single_scene_latency": sorted(lat4)[1], "GBps": scene_bytes / ms4 / 1e6,
        "launches_per_scene_per_gpu": int(nl4), "pair_bytes": scene_bytes,
        "streams": "SSIM kernel on the main stream, the two HBM-bound passes on a side stream in its shadow (engine.PreparedCaseAAll(side_stream=...))",
        "roofline": roof(scene_bytes, ms4, "fp64", 104.0 * B * H * W),
        "note": "time is dominated by the FP64-bound Gaussian SSIM kernel; the pair is read three times (statistics+planes, histograms, SSIM)",
    }
    if isinstance(comb, sharding.P2PRunCombiner):
        comb.close()
    del ref, tst, full, core, run4, outs4, prep4, comb
    torch.cuda.empty_cache()

    # ---- C5: Case-B rate sweep, 14 rates x 3 reps = 42 decoded cubes against ONE original, all Case-B metrics ---
    Bb, Hb, Wb, NP = 180, 1024, 1024, 42
    pair_bytes = 4 * Bb * Hb * Wb
    mine = list(range(rank, NP, world))
    g = torch.Generator(device=dev).manual_seed(11)      # the same original on every rank
    orig = torch.randint(0, 2500, (Hb, Wb, Bb), device=dev, dtype=torch.int16, generator=g) * 4
    decs = {}
    for i in mine:                                        # noise amplitude grows with the "rate" index, reps differ by seed
        gi = torch.Generator(device=dev).manual_seed(1000 + i)
        a = 1 + (i // 3)
        decs[i] = (orig + torch.randint(-a, a + 1, (Hb, Wb, Bb), device=dev, dtype=torch.int16, generator=gi)).clamp_(0, 32767)
    run5, outs5 = Partials.allocate_run(NP, Bb, 0, dev, "uint16")
    pairs5 = {i: DevicePair(orig, decs[i], "uint16", "bip", Bb, Hb, Wb) for i in mine}
    fused5 = {}
    for k, i in enumerate(mine):                          # pairs rotate over the streams (see C3)
        with torch.cuda.stream(alt[k % len(alt)]):
            fused5[i] = PreparedFused(pairs5[i], Want(stats=True, sam=True), outs5[i])
    rest = Want(stats=False, sid=True, lmse=True)
    comb5 = None
    exchange5 = "none"
    if world > 1:
        try:
            comb5 = sharding.P2PRunCombiner(run5, Bb, 0, batch=NP, timeout_s=20.0)
            exchange5 = "nvlink peer memory, ONE exchange of the 42 partial vectors at the end of the sweep"
        except Exception as e:      # noqa: BLE001
            print(f"[bench] rank {rank}: C5 P2P exchange unavailable ({e}); NCCL", file=sys.stderr)
            exchange5 = "NCCL all-reduce of the run at the end of the sweep"

    def sweep():
        run5.zero_()

        def body():
            for k, i in enumerate(mine):
                with torch.cuda.stream(alt[k % len(alt)]):
                    fused5[i].launch(chain=False)
                    evaluate(pairs5[i], rest, out=outs5[i])
        fork_join(body)

    def sweep_and_exchange():
        sweep()
        if comb5 is not None:
            comb5.finish(NP)
        elif world > 1:
            # every pair was evaluated by exactly one rank and the other ranks' vectors are all-zero bits, so an int64
            # SUM over the whole run gathers sums, maxima and (bit patterns of) float sums alike
            dist.all_reduce(run5, op=dist.ReduceOp.SUM)

    sweep_and_exchange()
    barrier()
    if comb5 is not None:
        comb5.check_status()
    h5 = outs5[NP - 1].to_host()
    m5 = finish.finish_compute_metrics(_lib.DM_U16, h5.sums, h5.maxs)
    s5 = finish.finish_spectral(float(h5.spec[0]), float(h5.spec[1]), float(h5.spec[2]), h5.lmse, Hb * Wb)
    assert int(h5.sums[0, 0]) == Hb * Wb and m5["max_abs_err"] == 1 + (NP - 1) // 3 and s5["sid"] > 0 and s5["lmse"] > 0 \
        and 0 < s5["sam_deg"] < 5, (m5["max_abs_err"], s5)
    times5 = []
    nl5 = 0
    for _ in range(3):
        if comb5 is not None:
            comb5.reset()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.dm_launch_count()
        a.record()
        sweep_and_exchange()
        b.record()
        barrier()
        times5.append(max_over_ranks(a.elapsed_time(b)))
        nl5 = L.dm_launch_count() - l0
    ms5 = sorted(times5)[1]
    sid_lmse_ops = (15.0 + 20.0) * Bb * Hb * Wb * NP
    out["C5_caseB_sweep_all_metrics"] = {
        "workload": "configs[4]: Case B rate sweep, 14 rates x 3 reps = 42 decoded 1024x1024x180 uint16 BIP cubes against one original, all Case-B metrics: "
                    "compute_metrics + SAM (one pass, dm_fused_bip), SID (dm_spectral), Sobel-LMSE (dm_sobel_lmse)",
        "scaling": "strong", "sharding": f"by pair over {world} GPU(s) ({len(mine)} pairs on rank 0), results gathered once", "exchange": exchange5,
        "ms_per_sweep": ms5, "ms_per_pair_per_gpu": ms5 / max(1, len(mine)), "GBps": NP * pair_bytes / ms5 / 1e6,
        "launches_per_sweep_rank0": int(nl5), "pair_bytes": pair_bytes, "streams": f"pairs rotate over {len(alt)} CUDA streams (kernel tails overlap the next pairs)",
        "roofline": roof(NP * pair_bytes, ms5, "fp64", sid_lmse_ops),
        "note": "each pair is read three times (statistics+SAM at HBM speed, then the issue/FP64-bound SID and Sobel-LMSE kernels); "
                "42 pairs on 8 ranks cannot scale past 42/6 = 7x",
    }
    if comb5 is not None:
        comb5.close()

    # ---- C2r: the REAL EnMAP configuration of configs[1]: int16 samples, nodata -32768 in both files ------------
    # (tools/make_baseline_B.py:302-312 writes that; mask rule run_codec.py:249-263).  A no-data corner (a triangle,
    # 5 % of the pixels: EnMAP tiles are cut from a rotated swath) in the original and the decoded cubes; per GPU three
    # decoded cubes of this rank's share of the sweep against the one original, rotated (> L2).  One launch per pair:
    # the validity rule is evaluated inside the one-pass kernel (dm_fused_bip_scan), the pair is read ONCE.
    yy = torch.arange(Hb, device=dev).view(Hb, 1)
    xx = torch.arange(Wb, device=dev).view(1, Wb)
    corner = (yy + xx) < int((2 * 0.05 * Hb * Wb) ** 0.5)
    n_corner = int(corner.sum().item())
    orig_r = orig.clone()
    orig_r[corner] = -32768
    ids_r = mine[:3]
    for i in ids_r:
        decs[i][corner] = -32768
    run_r, outs_r = Partials.allocate_run(64, Bb, 0, dev, "int16")
    pairs_r = [DevicePair(orig_r, decs[i], "int16", "bip", Bb, Hb, Wb, -32768, -32768) for i in ids_r]
    one_r = [PreparedFused(pairs_r[k % len(pairs_r)], Want(stats=True, sam=True), outs_r[k], scan=True) for k in range(64)]
    plane_r = torch.empty(Hb * Wb, dtype=torch.uint8, device=dev)
    two_out = Partials.allocate(Bb, 0, dev, "int16")

    def two_reads(k):
        evaluate(pairs_r[k % len(pairs_r)], Want(stats=True, sam=True, fused_scan=False), out=two_out)

    def sweep_r():
        run_r.zero_()
        for pf in one_r:
            pf.launch(chain=True)

    sweep_r()
    two_reads(0)
    torch.cuda.synchronize()
    hr, h2 = outs_r[0].to_host(), two_out.to_host()
    assert int(hr.counts[0]) == Hb * Wb - n_corner and int(hr.sums[0, 0]) == Hb * Wb - n_corner, (hr.counts, n_corner)
    assert np.array_equal(hr.isum, h2.isum) and np.array_equal(hr.imax, h2.imax) and np.array_equal(hr.fsum, h2.fsum), \
        "one-read route != dm_validity + dm_fused_bip"
    ms_r, _ = timed(sweep_r, 3, warm=1)
    ms_r /= len(one_r)
    ms_2, nl_2 = timed(lambda: [two_reads(k) for k in range(6)], 3, warm=1)
    ms_2 /= 6
    out["C2r_caseB_int16_nodata_one_read"] = {
        "workload": "configs[1] as the real EnMAP product: 1024x1024x180 int16 BIP cube pair per GPU with nodata -32768 in both "
                    "files (a no-data corner, 5 % of the pixels): compute_metrics (run_codec.py:249-304 mask rule included) + SAM",
        "scaling": "weak", "pairs_per_gpu_rotated": len(pairs_r), "launches_per_pair": 1,
        "api": "engine.PreparedFused(scan=True) -> dm_fused_bip_scan: validity evaluated by the kernel's pixel warps, pair read once",
        "us_per_pair": ms_r * 1e3, "GBps": world * pair_bytes / ms_r / 1e6,
        "two_reads_us_per_pair": ms_2 * 1e3, "two_reads_launches_per_pair": int(nl_2) // 6,
        "two_reads_api": "evaluate(fused_scan=False): dm_validity + dm_fused_bip (r01 path)",
        "roofline": roof(world * pair_bytes, ms_r, "hbm"),
        "checked": "counts, every integer partial and the SAM sum bit-identical to the two-read route on this rank's first pair",
    }
    del one_r, pairs_r, run_r, outs_r, orig_r, plane_r, two_out
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configurations (key \"configs\")")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line.  Libraries write banners to file descriptor 1 behind Python's back
    # (NCCL prints "NCCL version ..." there on this image), so descriptor 1 is pointed at stderr for the whole
    # run and the JSON line goes to the saved original descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj) -> None:
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist
    from image_compression_analysis_b200 import _lib, finish
    from image_compression_analysis_b200.engine import DevicePair, Partials, PreparedFused, Want, evaluate

    world, rank, local = _env_int("WORLD_SIZE", 1), _env_int("RANK", 0), _env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    from image_compression_analysis_b200.engine import bind_host_to_gpu_numa
    numa_node = None if os.environ.get("DM_NO_NUMA_BIND") else bind_host_to_gpu_numa(local)   # pinned staging local to the GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world and rank == 0:
        print(f"[bench] --gpus {args.gpus} but WORLD_SIZE={world}: reporting n_gpus={world}", file=sys.stderr)
    L = _lib.lib()
    want = Want(stats=True, sam=True)

    # ---- device-resident arm -------------------------------------------------------------------
    n_pairs = 3                                   # 2.26 GB rotating, every pair 6x the 126 MB L2
    pairs = make_device_pairs(torch, n_pairs, seed=rank + 1)
    # every step writes its own zeroed partial vector (31.5 KB), like the pairs of a rate sweep: no
    # memset inside the timed region, and with N GPUs no wait for the previous combine of a buffer
    # (the vectors of the run are contiguous, so with N GPUs a batch of them is exchanged with ONE
    # all-gather + dm_combine_partials on a side stream: the exchange is latency bound)
    run, outs = Partials.allocate_run(args.warmup + args.steps, BANDS, 0, pairs[0].ref.device, "uint16")

    from image_compression_analysis_b200.sharding import P2PRunCombiner, RunCombiner
    combiner, exchange = None, "none"
    NCCL_EXCHANGE = "NCCL all-gather + dm_combine_partials"
    if world > 1:
        # Default: partial vectors pushed over NVLink peer memory (no collective library in the loop).  It is
        # proven during the warm-up below and replaced by the NCCL exchange -- on ALL ranks together -- if the
        # IPC set-up fails or a peer's data does not arrive.  DM_EXCHANGE=nccl forces the NCCL path.
        ok = 0
        if os.environ.get("DM_EXCHANGE", "p2p") == "p2p":
            try:
                combiner = P2PRunCombiner(run, BANDS, 0, batch=COMBINE_BATCH, timeout_s=10.0)
                ok = 1
            except Exception as e:      # noqa: BLE001  (IPC not available in this container, ...)
                print(f"[bench] rank {rank}: P2P exchange unavailable ({e})", file=sys.stderr)
        t_ok = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        if int(t_ok.item()) == 1:
            exchange = "nvlink peer memory (dm_p2p_push / dm_p2p_combine)"
        else:
            combiner, exchange = RunCombiner(run, BANDS, 0, batch=COMBINE_BATCH), NCCL_EXCHANGE

    # the launches of the sweep are prepared once (all ctypes arguments built ahead): a step is one foreign call
    prepared = [PreparedFused(pairs[i % n_pairs], want, outs[i]) for i in range(args.warmup + args.steps)]

    def step(i):
        P = outs[i]
        prepared[i].launch()
        if combiner is not None:
            combiner.done(i)                  # every COMBINE_BATCH pairs: one exchange, overlapped
        return P

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    if combiner is not None:
        combiner.finish(args.warmup)
    barrier()
    if isinstance(combiner, P2PRunCombiner):
        # the warm-up pushed and combined real records: did every peer's data arrive, on every rank?
        bad = 0
        try:
            combiner.check_status()
        except RuntimeError as e:
            print(f"[bench] rank {rank}: {e}", file=sys.stderr)
            bad = 1
        if os.environ.get("DM_EXCHANGE_FORCE_FALLBACK"):      # exercises the switch below (tests)
            bad = 1
        t_bad = torch.tensor([bad], dtype=torch.int32, device="cuda")
        dist.all_reduce(t_bad, op=dist.ReduceOp.MAX)
        if int(t_bad.item()):
            combiner, exchange = RunCombiner(run, BANDS, 0, batch=COMBINE_BATCH), NCCL_EXCHANGE
            combiner._next = args.warmup            # the warm-up records are not part of the result
        barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                       # one sampling thread per box is enough (and NVML serialises)
    launches0 = L.dm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    if combiner is not None:
        combiner.finish(args.warmup + args.steps)   # the timed region ends when the last combine has finished
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if isinstance(combiner, P2PRunCombiner):
        combiner.check_status()
    ms_local = ms_total                      # this rank's own timed region (the max over ranks is taken below)
    launches = L.dm_launch_count() - launches0
    clocks = sampler.stop()

    # per-launch timing of the step's kernel: the same launches with CUDA events around each one on the
    # launching stream.  A short spin kernel is queued first so that both events and the launch are
    # already in the stream when the GPU reaches them (the event delta is then the kernel, not the
    # host's launch path).
    kern_ms = {"dm_fused_bip": []}
    scratch = [Partials.allocate(BANDS, 0, pairs[0].ref.device, "uint16") for _ in range(min(args.steps, 50))]
    torch.cuda.synchronize()
    for i, P in enumerate(scratch):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.dm_launch_count()
        torch.cuda._sleep(400_000)
        a.record()
        evaluate(pairs[i % n_pairs], want, out=P)
        b.record()
        b.synchronize()
        assert L.dm_launch_count() - l0 == 1, "the step is expected to be ONE launch of dm_fused_bip"
        kern_ms["dm_fused_bip"].append(a.elapsed_time(b))
    barrier()

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * PAIR_BYTES * args.steps / (ms_total * 1e-3) / 1e9

    # sanity: the step's result must be a real metric dict (guards against timing a no-op)
    Pchk = outs[-1]
    h = Pchk.to_host()
    torch.cuda.synchronize()
    res = finish.finish_compute_metrics(_lib.DM_U16, h.sums, h.maxs)
    sam = finish.finish_spectral(float(h.spec[0]), float(h.spec[1]), float(h.spec[2]), None, 1)
    assert int(h.sums[0, 0]) == world * ROWS * WIDTH and res["max_abs_err"] == 3 and 0 < sam["sam_deg"] < 1, (res, sam)

    # ---- end-to-end arm: pinned host cubes -> public API -> metrics dict -------------------------
    e2e = None
    if not args.no_e2e:
        host = []
        for i in range(2):
            r = torch.empty((ROWS, WIDTH, BANDS), dtype=torch.int16).pin_memory()
            d = torch.empty((ROWS, WIDTH, BANDS), dtype=torch.int16).pin_memory()
            r.copy_(pairs[i].ref); d.copy_(pairs[i].tst)
            host.append((r.view(torch.uint16), d.view(torch.uint16)))
        torch.cuda.synchronize()

        from image_compression_analysis_b200.engine import evaluate_host_pairs

        def e2e_run(n, share_ref=False):
            """n pairs from pinned host memory through the public sweep API: every pair is uploaded, evaluated
            (+ exchanged across ranks), read back and finished on the host; uploads of the next pair overlap.
            share_ref: the pairs of a rate sweep share ONE original (BASELINE configs[4]); it is uploaded once."""
            outs_, hp_ = [], None
            src = (host[0] for i in range(n)) if share_ref else (host[i % 2] for i in range(n))
            for hp_ in evaluate_host_pairs(src, want, layout="bip", share_ref=share_ref):
                o = finish.finish_compute_metrics(_lib.DM_U16, hp_.sums, hp_.maxs)
                o.update(finish.finish_spectral(float(hp_.spec[0]), float(hp_.spec[1]), float(hp_.spec[2]), None, 1))
                outs_.append(o)
            return outs_, hp_

        e2e_steps = max(3, min(args.steps, 10))
        e2e_run(2)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        results, hp = e2e_run(e2e_steps)
        b.record()
        barrier()
        assert len(results) == e2e_steps and results[-1]["max_abs_err"] == 3
        ms_e2e = a.elapsed_time(b)
        te = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        ms_e2e = float(te.item())
        d2h = (hp.isum.nbytes + hp.imax.nbytes + hp.fsum.nbytes)
        e2e = {"value": world * PAIR_BYTES * e2e_steps / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": PAIR_BYTES, "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "ms_per_step": ms_e2e / e2e_steps,
               "api": "engine.evaluate_host_pairs(pinned host cubes) [upload of pair i+1 overlaps the kernels, exchange, "
                      "read-back and host finish of pair i] -> finish.*", "host_numa_node": numa_node}
        # the same sweep when the decoded cubes share ONE original, as the pairs of a rate sweep do (run_codec.py:472-475):
        # the original is uploaded once, only the decoded cube crosses the link per step
        e2e_run(2, share_ref=True)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        results, hp = e2e_run(e2e_steps, share_ref=True)
        b.record()
        barrier()
        assert len(results) == e2e_steps and results[-1]["max_abs_err"] == 3
        ts = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ms_sh = float(ts.item())
        e2e["shared_original"] = {"value": world * PAIR_BYTES * e2e_steps / (ms_sh * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_sh / e2e_steps,
                                  "h2d_bytes_per_step": PAIR_BYTES // 2, "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                                  "note": "engine.evaluate_host_pairs(share_ref=True): GB/s still counts both cubes of every pair (SURVEY 8d); "
                                          "the original is uploaded once per sweep (its one-off upload is inside the timed region)"}

    peak, peak_src = hbm_peak()
    configs = None
    if not args.no_configs:
        # free the headline arm's cubes first (the sweep of config 5 alone holds 16 GB at one GPU)
        del pairs, prepared, scratch
        torch.cuda.empty_cache()
        configs = run_configs(torch, dist, world, rank, local, peak, clocks.get("sm_mhz") if rank == 0 else None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    dominant = "dm_fused_bip"
    # the step IS one launch of this kernel, back to back on one stream: its average duration over the
    # timed region is the region's CUDA-event time / launches (this rank's own clock); the isolated
    # launches measured above (each behind a spin kernel: pipeline empty at start and end) are reported next to it
    dom_ms = ms_local / args.steps
    iso_ms = sum(kern_ms[dominant]) / len(kern_ms[dominant])
    achieved = PAIR_BYTES / (dom_ms * 1e-3) / 1e9
    # DRAM traffic per launch can only come from a profiler: the figure is the committed ncu capture of THIS kernel
    # instance (profiles/traffic.json names the capture), not something measured in this run
    traffic, traffic_src = None, None
    tr = ROOT / "profiles" / "traffic.json"
    if tr.exists():
        try:
            tj = json.loads(tr.read_text())
            traffic, traffic_src = tj.get(dominant), tj.get("source")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "layout": "bip", "bands": BANDS, "rows_per_gpu": ROWS, "width": WIDTH,
                   "pair_bytes_per_gpu": PAIR_BYTES, "exchange": exchange, "l2_policy": f"inputs larger than L2: {n_pairs} distinct 755 MB pairs rotated",
                   "sharding": (f"row strips, one per GPU; the integer/float64 partials of every {COMBINE_BATCH} pairs are combined "
                                "with one NCCL all-gather + dm_combine_partials on a side stream, overlapped with the next "
                                "pairs' kernels; the timed region ends after the last combine") if world > 1 else "single GPU",
                   "kernels_per_step": ["dm_fused_bip (fused_ct_kernel<180>, 23 band + 8 pixel warps: per-band stats + per-pixel SAM from one read, "
                                        "SAM partials reduced in-kernel), launched through engine.PreparedFused; consecutive launches overlap "
                                        "tail and ramp-up through programmatic dependent launch"]},
        "frac_of_hbm_peak": value / (world * peak),
        "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": UNIT,
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel_instance": "fused_ct_kernel<180, u16, unmasked, no planes, 23 band warps> (dm_fused_bip_variant 0 = auto)",
                     "algorithmic_bytes_per_launch": PAIR_BYTES,
                     "launch_ms": {dominant: dom_ms}, "isolated_launch_ms": {dominant: iso_ms},
                     "note": "launch_ms: CUDA events over the timed region on the launching stream / launches (the step is "
                             "one launch; consecutive launches are chained by programmatic dependent launch, so the "
                             "ramp-up of one overlaps the tail of the previous and this is the steady-state duration); "
                             "isolated_launch_ms: events around single launches queued behind a spin kernel (includes "
                             "the kernel's ramp-up and tail on an otherwise idle GPU)"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "configs": configs,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_single(rows=512, reps=3)
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
