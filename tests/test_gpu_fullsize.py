"""GPU: the BASELINE.json configurations at their FULL sizes, checked through size-independent
properties (the oracle would need minutes to hours there): exact integer identities against
independent torch integer reductions, histogram identities, shard-and-combine == single shot,
one-pass == two-pass, lossless pairs, monotonic rate sweeps."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _torch_band_sums(ref, tst, band_axis):
    """Independent per-band integer reductions with torch int64 ops (chunked to bound memory)."""
    import torch
    B = ref.shape[band_axis]
    out = {k: torch.zeros(B, dtype=torch.int64, device=ref.device) for k in ("x", "y", "xx", "yy", "xy", "abs", "sse")}
    mx = torch.zeros(B, dtype=torch.int64, device=ref.device)
    other = 0 if band_axis != 0 else 1
    n = ref.shape[other]
    step = max(1, n // 16)
    dims = tuple(d for d in range(ref.dim()) if d != band_axis)
    for i in range(0, n, step):
        sl = [slice(None)] * ref.dim()
        sl[other] = slice(i, min(n, i + step))
        x = ref[tuple(sl)].to(torch.int64) & 0xFFFF
        y = tst[tuple(sl)].to(torch.int64) & 0xFFFF
        d = x - y
        out["x"] += x.sum(dims); out["y"] += y.sum(dims)
        out["xx"] += (x * x).sum(dims); out["yy"] += (y * y).sum(dims); out["xy"] += (x * y).sum(dims)
        out["abs"] += d.abs().sum(dims); out["sse"] += (d * d).sum(dims)
        mx = torch.maximum(mx, d.abs().amax(dims))
    return {k: v.cpu().numpy() for k, v in out.items()}, mx.cpu().numpy()


def _pair(B, H, W, layout, seed, amp=3, top=2500, mul=4):
    import torch
    from image_compression_analysis_b200.engine import DevicePair
    g = torch.Generator(device="cuda").manual_seed(seed)
    shape = (B, H, W) if layout == "bsq" else (H, W, B)
    if mul == 16:
        # 12-in-16 data up to 65520: uint16 bit patterns in int16 storage (wrapping arithmetic), kept
        # amp*16 away from both ends so that no clamp is needed
        ref = torch.randint(amp, top - amp, shape, device="cuda", dtype=torch.int16, generator=g) * 16
        tst = ref + torch.randint(-amp, amp + 1, shape, device="cuda", dtype=torch.int16, generator=g) * 16
        return DevicePair(ref, tst, "uint16", layout, B, H, W)
    ref = torch.randint(0, top, shape, device="cuda", dtype=torch.int16, generator=g) * mul
    tst = ref.clone()
    if amp:
        tst += torch.randint(-amp, amp + 1, shape, device="cuda", dtype=torch.int16, generator=g)
        tst.clamp_(0, 32767)
    return DevicePair(ref, tst, "uint16", layout, B, H, W)


def test_case_b_cube_full_size():
    """configs[1]: EnMAP 1024x1024x180 BIP -- the bench workload."""
    import torch
    from image_compression_analysis_b200 import _lib, finish
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    B, H, W = 180, 1024, 1024
    pair = _pair(B, H, W, "bip", seed=2)
    P = evaluate(pair, Want(stats=True, sam=True))
    torch.cuda.synchronize()
    h = P.to_host()
    sums, mx = _torch_band_sums(pair.ref, pair.tst, band_axis=2)
    assert np.array_equal(h.sums[:, _lib.DM_S_N], np.full(B, H * W))
    for k, col in (("x", _lib.DM_S_X), ("y", _lib.DM_S_Y), ("xx", _lib.DM_S_XX), ("yy", _lib.DM_S_YY),
                   ("xy", _lib.DM_S_XY), ("abs", _lib.DM_S_ABS), ("sse", _lib.DM_S_SSE)):
        assert np.array_equal(h.sums[:, col], sums[k]), k
    assert np.array_equal(h.maxs[:, _lib.DM_M_MAXERR], mx)
    # SAM against a float64 torch evaluation of run_codec.py:328-332
    x = (pair.ref.to(torch.float64)); y = pair.tst.to(torch.float64)
    dot = (x * y).sum(-1); na = x.pow(2).sum(-1).sqrt() + 1e-12; nr = y.pow(2).sum(-1).sqrt() + 1e-12
    want_sam = float(torch.rad2deg(torch.arccos(torch.clamp(dot / (na * nr), -1, 1)).mean()))
    del x, y, dot, na, nr
    got_sam = finish.finish_spectral(float(h.spec[0]), 0.0, float(h.spec[2]), None, H * W)["sam_deg"]
    assert h.spec[2] == H * W and abs(got_sam - want_sam) <= 1e-9 * want_sam, (got_sam, want_sam)
    # one pass == two passes; 4 row strips accumulated == single shot (integers exact, SAM to rounding)
    P2 = evaluate(pair, Want(stats=True, sam=True, fused=False))
    Ps = Partials.allocate(B, 0, pair.ref.device, "uint16")
    for s in range(4):
        r0, r1 = s * H // 4, (s + 1) * H // 4
        strip = DevicePair(pair.ref[r0:r1], pair.tst[r0:r1], "uint16", "bip", B, r1 - r0, W)
        evaluate(strip, Want(stats=True, sam=True), out=Ps)
    torch.cuda.synchronize()
    for Q in (P2, Ps):
        q = Q.to_host()
        assert np.array_equal(q.isum, h.isum)
        assert np.array_equal(q.maxs.max(0), h.maxs.max(0)) and np.array_equal(q.maxs[:, 0], h.maxs[:, 0])
        assert q.spec[2] == h.spec[2] and abs(q.spec[0] - h.spec[0]) <= 1e-11 * h.spec[0]
    res = finish.finish_compute_metrics(_lib.DM_U16, h.sums, h.maxs)
    assert res["max_abs_err"] == 3 and res["lossless"] == 0 and 0 < res["ssim_global"] <= 1


def test_full_scene_all_integer_metrics():
    """configs[3]: Sentinel-2 scene 10980x10980x4 BSQ, 12-in-16 data: stats + ERR8 quicklooks in one pass."""
    import torch
    from image_compression_analysis_b200 import _lib, finish
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    B, H, W = 4, 10980, 10980
    pair = _pair(B, H, W, "bsq", seed=4, amp=3, top=4096, mul=16)
    P = evaluate(pair, Want(stats=True, err8_caps=(255, 32)))
    Ph = evaluate(pair, Want(stats=True, hist_bins=256))
    torch.cuda.synchronize()
    h, hh = P.to_host(), Ph.to_host()
    N = H * W
    sums, mx = _torch_band_sums(pair.ref, pair.tst, band_axis=0)
    for k, col in (("x", _lib.DM_S_X), ("y", _lib.DM_S_Y), ("xx", _lib.DM_S_XX), ("yy", _lib.DM_S_YY),
                   ("xy", _lib.DM_S_XY), ("abs", _lib.DM_S_ABS), ("sse", _lib.DM_S_SSE)):
        assert np.array_equal(h.sums[:, col], sums[k]), k
    assert np.array_equal(h.isum[:B * 8], hh.isum[:B * 8])                   # one-pass kernel == plain stats kernel
    # histogram identities: N, sum k H = S|d|, sum k^2 H = SSE (max < K), top bin = max
    k = np.arange(256, dtype=np.int64)
    for b in range(B):
        Hb = hh.hist[b]
        assert Hb.sum() == N and (k * Hb).sum() == h.sums[b, _lib.DM_S_ABS] and (k * k * Hb).sum() == h.sums[b, _lib.DM_S_SSE]
        assert np.nonzero(Hb)[0].max() == h.maxs[b, _lib.DM_M_MAXERR] == mx[b] == 48
    # 12-in-16 detection (run_codec.py:100-101) and the observed range
    assert finish.data_range_from_maxs(_lib.DM_U16, h.maxs) == 4095
    # ERR8 planes: LUT of the per-pixel max over bands, and their histograms
    e = ((pair.ref.to(torch.int32) & 0xFFFF) - (pair.tst.to(torch.int32) & 0xFFFF)).abs().amax(0).reshape(-1)
    for cap, key, hist in ((255, "err8_g", h.hist8_g), (32, "err8_z", h.hist8_z)):
        lut = torch.from_numpy(finish.err8_lut(cap)).cuda()
        want = lut[e.clamp(max=cap).long()]
        assert torch.equal(P.planes[key], want), key
        assert np.array_equal(hist, torch.bincount(want.long(), minlength=256).cpu().numpy())
    del e
    # 8 row strips accumulated == single shot
    Ps = Partials.allocate(B, 0, pair.ref.device, "uint16")
    rows = [(s * H // 8, (s + 1) * H // 8) for s in range(8)]
    for r0, r1 in rows:
        strip = DevicePair(pair.ref.view(-1)[r0 * W:], pair.tst.view(-1)[r0 * W:], "uint16", "bsq", B, r1 - r0, W,
                           band_stride=H * W)
        evaluate(strip, Want(stats=True), out=Ps)
    torch.cuda.synchronize()
    q = Ps.to_host()
    assert np.array_equal(q.isum[:B * 8], h.isum[:B * 8]) and np.array_equal(q.maxs[:, 0], h.maxs[:, 0])
    # identical pair: lossless
    same = DevicePair(pair.ref, pair.ref, "uint16", "bsq", B, H, W)
    z = evaluate(same, Want(stats=True, err8_caps=(255, None))).to_host()
    r = finish.finish_compute_metrics(_lib.DM_U16, z.sums, z.maxs)
    assert r["lossless"] == 1 and r["max_abs_err"] == 0 and math.isinf(r["psnr_global"]) and r["ssim_global"] == 1.0
    assert z.hist8_g[0] == N and z.hist8_g[1:].sum() == 0


def test_rate_sweep_is_monotonic_and_runs_in_one_run():
    """configs[4]: one original, 14 rates (error amplitude falling with rate) x 3 reps, partial vectors of
    the sweep in ONE contiguous run: PSNR rises and SAM falls with the rate, reps agree closely."""
    import torch
    from image_compression_analysis_b200 import _lib, finish
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    B, H, W = 180, 1024, 1024
    base = _pair(B, H, W, "bip", seed=5, amp=0)
    rates = list(range(14))
    run, outs = Partials.allocate_run(len(rates) * 3, B, 0, base.ref.device, "uint16")
    g = torch.Generator(device="cuda").manual_seed(55)
    dec = torch.empty_like(base.ref)
    i = 0
    for r in rates:
        amp = 2 ** (13 - r) if r < 13 else 0
        for rep in range(3):
            dec.copy_(base.ref)
            if amp:
                dec += torch.randint(-amp, amp + 1, dec.shape, device="cuda", dtype=torch.int16, generator=g)
                dec.clamp_(0, 32767)
            evaluate(DevicePair(base.ref, dec, "uint16", "bip", B, H, W), Want(stats=True, sam=True), out=outs[i])
            i += 1
    torch.cuda.synchronize()
    host = run.cpu().numpy()
    ni, nm, nf = Partials.sizes(B, 0)
    psnr, sam = [], []
    for j in range(len(rates) * 3):
        isum, imax, fsum = host[j, :ni], host[j, ni:ni + nm], host[j, ni + nm:].view(np.float64)
        res = finish.finish_compute_metrics(_lib.DM_U16, isum[:B * 8].reshape(B, 8), imax.reshape(B, 8))
        psnr.append(res["psnr_global"])
        sam.append(finish.finish_spectral(float(fsum[0]), 0.0, float(fsum[2]), None, H * W)["sam_deg"])
    psnr, sam = np.array(psnr).reshape(14, 3), np.array(sam).reshape(14, 3)
    assert np.all(np.isinf(psnr[13])) and np.all(sam[13] < 1e-5)              # last rate: lossless
    assert np.all(np.diff(psnr[:13].mean(1)) > 0) and np.all(np.diff(sam[:13].mean(1)) < 0)
    assert np.all(np.ptp(psnr[:13], axis=1) < 0.01) and np.all(np.ptp(sam[:13], axis=1) < 1e-3 * sam[:13].mean(1))


def test_chained_launches_accumulate_exactly():
    """Programmatic dependent launch lets launch N+1 start while launch N drains.  Fifty launches on the
    Case-B cube accumulating into ONE partial vector (64-bit RED for the integers, the last block's ordered
    `+=` for the SAM sums, the shared workspace counter) must equal fifty times the single result -- any race
    between a draining launch and its successor would show here."""
    import torch
    from image_compression_analysis_b200.engine import DevicePair, Partials, PreparedFused, Want, evaluate
    B, H, W = 180, 1024, 1024
    g = torch.Generator(device="cuda").manual_seed(9)
    ref = torch.randint(0, 2500, (H, W, B), device="cuda", dtype=torch.int16, generator=g) * 4
    tst = (ref + torch.randint(-3, 4, (H, W, B), device="cuda", dtype=torch.int16, generator=g)).clamp_(0, 32767)
    pair = DevicePair(ref, tst, "uint16", "bip", B, H, W)
    want = Want(stats=True, sam=True)
    one = evaluate(pair, want).to_host()
    n = 50
    P = Partials.allocate(B, 0, ref.device, "uint16")
    launch = PreparedFused(pair, want, P)
    for _ in range(n):
        launch.launch()
    many = P.to_host()
    assert np.array_equal(many.isum, n * one.isum)
    assert np.array_equal(many.imax, one.imax)
    acc = 0.0
    for _ in range(n):
        acc += float(one.spec[0])                       # the same left-to-right float64 additions
    assert float(many.spec[0]) == acc and float(many.spec[2]) == n * float(one.spec[2])
    # and interleaved with a launch that writes planes (no early trigger) and a masked one
    P2 = Partials.allocate(B, 0, ref.device, "uint16")
    for k in range(6):
        evaluate(pair, Want(stats=True, sam=True, err8_caps=(255, 32) if k % 2 else (None, None)), out=P2)
    h2 = P2.to_host()
    assert np.array_equal(h2.isum[:B * 8], 6 * one.isum[:B * 8])


def test_case_b_cube_fp64_metrics_strips_sum_to_whole_and_one_strip_matches_oracle():
    """configs[1] / [4] at FULL size for the float64 kernels (SAM, SID, Sobel-LMSE): the oracle needs about a minute
    for a 1024 x 1024 x 180 cube, so (a) the whole cube in one launch must equal the sum over eight 128-row strips
    evaluated one by one (halo row for the stencil), i.e. the grid-stride / prefetching paths at 9 472 warps x 110
    pixels agree with launches of the size the oracle checks elsewhere, and (b) ONE of those strips, taken as an image
    of its own, is compared with the oracle directly (131 072 pixels of the full-size data)."""
    import torch
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    from oracle import distortion_oracle as orc
    B, H, W = 180, 1024, 1024
    pair = _pair(B, H, W, "bip", seed=21)
    want = Want(stats=False, sam=True, sid=True, lmse=True)
    whole = evaluate(pair, want).to_host()
    acc = Partials.allocate(B, 0, pair.ref.device, "uint16")
    n = 8
    for k in range(n):
        r0, r1 = H * k // n, H * (k + 1) // n
        b0, b1 = max(0, r0 - 1), min(H, r1 + 1)
        core = DevicePair(pair.ref[r0:r1], pair.tst[r0:r1], "uint16", "bip", B, r1 - r0, W)
        evaluate(core, Want(stats=False, sam=True, sid=True), out=acc)
        buf = DevicePair(pair.ref[b0:b1], pair.tst[b0:b1], "uint16", "bip", B, b1 - b0, W, None, None, b0, H)
        evaluate(buf, Want(stats=False, lmse=True), out=acc, rows=(r0 - b0, r1 - b0))
    torch.cuda.synchronize()
    parts = acc.to_host()
    assert parts.spec[2] == whole.spec[2] == H * W
    assert abs(parts.spec[0] - whole.spec[0]) <= 1e-11 * abs(whole.spec[0])          # sum of arccos
    assert abs(parts.spec[1] - whole.spec[1]) <= 1e-11 * abs(whole.spec[1])          # sum of SID
    assert np.allclose(parts.lmse, whole.lmse, rtol=1e-11, atol=0)
    assert whole.spec[1] > 0 and whole.lmse.min() > 0
    # (b) strip 3 as an image of its own against the oracle
    r0, r1 = 3 * H // n, 4 * H // n
    sub = DevicePair(pair.ref[r0:r1].contiguous(), pair.tst[r0:r1].contiguous(), "uint16", "bip", B, r1 - r0, W)
    h = evaluate(sub, want).to_host()
    got = finish.finish_spectral(float(h.spec[0]), float(h.spec[1]), float(h.spec[2]), h.lmse, (r1 - r0) * W)
    a = np.ascontiguousarray(np.moveaxis(sub.ref.cpu().numpy().view(np.uint16), -1, 0))
    b = np.ascontiguousarray(np.moveaxis(sub.tst.cpu().numpy().view(np.uint16), -1, 0))
    ref_vals = orc.compute_sam_sid_lmse_caseB(a, b)
    for k_ in ("sam_deg", "sid", "lmse"):
        assert abs(got[k_] - ref_vals[k_]) <= 1e-6 * abs(ref_vals[k_]), (k_, got[k_], ref_vals[k_])


def test_scene_gaussian_ssim_strips_sum_to_whole():
    """configs[3] at FULL size for the Gaussian-window SSIM: the 10980 x 10980 x 4 scene in one launch against the sum
    over five row strips with their 5-row halos (each strip is of the size range the oracle checks in
    tests/test_gpu_midsize.py); counts exact, sums to 1e-11."""
    import torch
    from image_compression_analysis_b200.engine import DevicePair, Partials, Want, evaluate
    B, H, W = 4, 10980, 10980
    pair = _pair(B, H, W, "bsq", seed=5, amp=3, top=4090, mul=16)
    want = Want(stats=False, ssim_gauss=True)
    whole = evaluate(pair, want, data_range=65535.0).to_host()
    acc = Partials.allocate(B, 0, pair.ref.device, "uint16")
    n = 5
    for k in range(n):
        r0, r1 = H * k // n, H * (k + 1) // n
        b0, b1 = max(0, r0 - 5), min(H, r1 + 5)
        buf = DevicePair(pair.ref[:, b0:b1].contiguous(), pair.tst[:, b0:b1].contiguous(), "uint16", "bsq", B, b1 - b0, W, None, None, b0, H)
        evaluate(buf, want, out=acc, rows=(r0 - b0, r1 - b0), data_range=65535.0)
    torch.cuda.synchronize()
    parts = acc.to_host()
    assert np.array_equal(parts.ssimw_cnt, whole.ssimw_cnt) and int(whole.ssimw_cnt[0]) == (H - 10) * (W - 10)
    assert np.allclose(parts.ssimw_sum, whole.ssimw_sum, rtol=1e-11, atol=0)
    assert np.all(whole.ssimw_sum / whole.ssimw_cnt < 1.0) and np.all(whole.ssimw_sum / whole.ssimw_cnt > 0.5)
