"""CPU: the C-ABI library's exports, the host-side finish, sharding and the gloo combine."""
import math
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from image_compression_analysis_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    header = (ROOT / "include" / "dm_b200.h").read_text()
    declared = set(re.findall(r"^\s*(?:int|int64_t|void|const char\*)\s+(dm_\w+)\s*\(", header, flags=re.M))
    assert declared, "no declarations parsed from include/dm_b200.h"
    assert declared == set(built_lib.SYMBOLS), (declared ^ set(built_lib.SYMBOLS))
    lib = built_lib.lib()               # binds every symbol, raises if one is missing
    for name in declared:
        assert hasattr(lib, name)
    assert lib.dm_abi_version() == built_lib.ABI_VERSION
    m = re.search(r"#define DM_ABI_VERSION (\d+)", header)
    assert int(m.group(1)) == built_lib.ABI_VERSION


def test_no_gpu_is_an_error_not_a_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import image_compression_analysis_b200 as dm
    a = np.zeros((1, 4, 4), np.uint16)
    with pytest.raises(RuntimeError):
        dm.compute_metrics_arrays(a, a)
    assert built_lib.lib().dm_device_sm_count() < 0
    assert b"CUDA" in built_lib.lib().dm_last_error()


def test_product_does_not_import_the_oracle():
    pkg = ROOT / "image_compression_analysis_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def _partials_numpy(ref, tst, sel=None):
    """Independent numpy statement of the per-band partials (test helper)."""
    from image_compression_analysis_b200 import _lib as L
    B = ref.shape[0]
    S = np.zeros((B, 8), np.int64)
    M = np.zeros((B, 8), np.int64)
    for b in range(B):
        x = ref[b].astype(np.int64)
        y = tst[b].astype(np.int64)
        M[b, L.DM_M_UMAX] = max(0, int(x.max()))
        M[b, L.DM_M_UNEGMIN] = max(0, -int(x.min()))
        M[b, L.DM_M_LOW4] = int(np.any(ref[b] & 0xF))
        M[b, L.DM_M_LOW2] = int(np.any(ref[b] & 0x3))
        if sel is not None:
            x, y = x[sel], y[sel]
        d = np.abs(x - y)
        S[b] = [x.size, x.sum(), y.sum(), (x * x).sum(), (y * y).sum(), (x * y).sum(), d.sum(), (d * d).sum()]
        if x.size:
            M[b, L.DM_M_MAXERR] = d.max()
            ax = np.abs(ref[b][sel] if sel is not None else ref[b]).astype(np.int64)    # native-dtype abs (wraps)
            ay = np.abs(tst[b][sel] if sel is not None else tst[b]).astype(np.int64)
            M[b, L.DM_M_ABSXY] = max(0, int(ax.max()), int(ay.max()))
    return S, M


@pytest.mark.parametrize("name", ["a_gauss", "a_identical", "a_near3_masked", "u16_fullrange", "u8", "b_u16_masked",
                                  "i16_fullrange_wrap", "seven_bands_odd", "single_band_odd"])
def test_finish_reproduces_reference_goldens(built_lib, name):
    """finish_compute_metrics on exact integer partials == the reference's own outputs."""
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import dtype_code
    from tests import goldenio
    c = goldenio.load(name)
    sel = c["valid"] if (c["valid"] is not None and c["valid"].any()) else None
    S, M = _partials_numpy(c["ref"], c["tst"], sel)
    got = finish.finish_compute_metrics(dtype_code(c["ref"].dtype), S, M)
    want = c["compute_metrics"]
    assert set(got) == set(want)
    for k, w in want.items():
        g = got[k]
        if isinstance(w, int):
            assert isinstance(g, int) and g == w, k
        elif k.startswith("psnr"):
            assert (math.isnan(g) and math.isnan(w)) or g == w, (k, g, w)       # bit-exact
        else:
            assert goldenio.close(g, w, rel=1e-12), (k, g, w)


def test_err8_lut_matches_reference_expression(built_lib):
    from image_compression_analysis_b200 import finish
    from oracle import distortion_oracle as orc
    for cap in (1, 2, 7, 16, 32, 48, 200, 255, 1000, 4095, 65535):
        assert np.array_equal(finish.err8_lut(cap), orc.err8_lut(cap)), cap
    lut0 = finish.err8_lut(0)
    assert lut0[0] == 0 and lut0[-1] == 255
    h = np.zeros(256, np.int64); h[3] = 10; h[200] = 5
    st = finish.err8_stats_from_hist(h)
    plane = np.array([3] * 10 + [200] * 5, np.uint8)
    assert st["mean"] == float(plane.mean()) and abs(st["std"] - float(plane.std())) < 1e-12


def test_strips_cover_the_image():
    from image_compression_analysis_b200.sharding import strips
    for H in (1, 7, 1024, 10980):
        for world in (1, 2, 3, 8):
            for halo in (0, 1, 5):
                for align in (1, 2, 8):
                    ss = strips(H, world, halo, align)
                    assert len(ss) == world
                    assert ss[0].row0 == 0 and ss[-1].row1 == H
                    for a, b in zip(ss, ss[1:]):
                        assert a.row1 == b.row0
                    for s in ss:
                        assert 0 <= s.buf0 <= s.row0 <= s.row1 <= s.buf1 <= H
                        if s.rows:
                            assert s.buf0 == max(0, s.row0 - halo) and s.buf1 == min(H, s.row1 + halo)
                            assert s.row0 % align == 0


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DM_ROOT"])
from tests.test_host_logic import _partials_numpy
from tests import goldenio
from image_compression_analysis_b200 import finish
from image_compression_analysis_b200.engine import Partials, dtype_code
from image_compression_analysis_b200.sharding import strips
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
c = goldenio.load("a_gauss")
ref, tst = c["ref"], c["tst"]
B, H, W = ref.shape
s = strips(H, world)[rank]
S, M = _partials_numpy(ref[:, s.row0:s.row1], tst[:, s.row0:s.row1])
P = Partials.allocate(B, 0, torch.device("cpu"), "uint16")
P.sums.copy_(torch.from_numpy(S.reshape(-1)))
P.imax.copy_(torch.from_numpy(M.reshape(-1)))
P.spec.copy_(torch.tensor([0.25 * (rank + 1), 0.0, float(s.rows * W)], dtype=torch.float64))
P.allreduce_()
h = P.to_host()
got = finish.finish_compute_metrics(dtype_code(ref.dtype), h.sums, h.maxs)
for k, w in c["compute_metrics"].items():
    g = got[k]
    ok = (g == w) if (isinstance(w, int) or k.startswith("psnr")) else goldenio.close(g, w, rel=1e-12)
    assert ok, (k, g, w)
assert h.spec[2] == H * W and abs(h.spec[0] - 0.25 * sum(range(1, world + 1))) < 1e-15
# the same through a RUN of partial vectors (the pairs of a sweep, one exchange for several of them)
run, Ps = Partials.allocate_run(3, B, 0, torch.device("cpu"), "uint16")
for i, Q in enumerate(Ps):
    Q.sums.copy_(torch.from_numpy(S.reshape(-1)) * (i + 1))
    Q.imax.copy_(torch.from_numpy(M.reshape(-1)) + i)
    Q.spec.copy_(torch.tensor([0.5 * (rank + 1) + i, 0.0, float(s.rows * W)], dtype=torch.float64))
Partials.allreduce_run_(run, 1, 3, B, 0)
h0, h2 = Ps[0].to_host(), Ps[2].to_host()
assert np.array_equal(h0.sums, S)                                   # outside the exchanged range: untouched
assert np.array_equal(h2.sums[:, 0], np.full(B, 3 * H * W))          # N of record 2: 3 x the image's pixels
assert np.array_equal(h2.maxs[:, 0], h.maxs[:, 0] + 2)
assert h2.spec[2] == H * W and abs(h2.spec[0] - (0.5 * sum(range(1, world + 1)) + 2 * world)) < 1e-12
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_gloo_world2_combine_equals_single_shot(built_lib, tmp_path):
    """world_size-2 gloo: per-strip partials + Partials.allreduce_ + finish == the reference golden."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, DM_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                       capture_output=True, text=True, env=env, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2, r.stdout + r.stderr


_P2P_FAIL_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["DM_ROOT"])
import torch, torch.distributed as dist
dist.init_process_group("gloo")
from image_compression_analysis_b200.engine import Partials
from image_compression_analysis_b200.sharding import P2PRunCombiner
run, _ = Partials.allocate_run(4, 6, 0, torch.device("cpu"), "uint16")
try:
    P2PRunCombiner(run, 6, 0, batch=2)
    print("unexpected success")
except RuntimeError as e:
    print("raised together:", str(e)[:60])
dist.barrier()
print("ok")
'''


def test_p2p_exchange_setup_fails_collectively(built_lib, tmp_path):
    """Without a usable GPU the peer-memory set-up cannot allocate: every rank must leave the constructor with
    the same RuntimeError (no rank left waiting in a collective), so that the caller can fall back to NCCL."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the set-up would succeed")
    script = tmp_path / "worker_p2p.py"
    script.write_text(_P2P_FAIL_WORKER)
    env = dict(os.environ, DM_ROOT=str(ROOT), MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29519", str(script)],
                       capture_output=True, text=True, env=env, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("raised together") == 2 and r.stdout.count("ok") == 2, r.stdout + r.stderr


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/dm_b200.h must compile as C99 (and as C++) on its own."""
    hdr = str(ROOT / "include" / "dm_b200.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_small_host_helpers():
    """Best-effort NUMA binding never raises; stretch tables cover the whole sample type in bin order."""
    from image_compression_analysis_b200 import finish
    from image_compression_analysis_b200.engine import bind_host_to_gpu_numa, integral_nodata
    assert bind_host_to_gpu_numa(0) in (None, 0, 1, 2, 3, 4, 5, 6, 7)
    lut = finish.stretch8_lut(100.0, 200.0, "uint16")
    assert lut.shape == (65536,) and lut[0] == 0 and lut[100] == 0 and lut[200] == 255 and lut[65535] == 255
    x = np.arange(65536, dtype=np.int64).astype(np.uint16)
    want = (np.clip((x.astype(np.float32) - 100.0) / (200.0 - 100.0 + 1e-9), 0, 1) * 255.0).astype(np.uint8)
    assert np.array_equal(lut, want)
    li = finish.stretch8_lut(-50.0, 50.0, "int16")          # bin 0 is -32768
    assert li[0] == 0 and li[32768 - 50] == 0 and li[32768 + 50] >= 254 and li[65535] == 255
    lb = finish.stretch8_lut(0.0, 255.0, "uint8")
    assert lb.shape == (65536,) and lb[255] >= 254 and not lb[256:].any()
    assert integral_nodata(-32768.0, "int16") == -32768 and integral_nodata(float("nan"), "int16") is None
    assert integral_nodata(70000, "uint16") is None and integral_nodata(1.5, "uint16") is None


def test_c_consumer_links_and_runs(built_lib, tmp_path):
    """examples/c_abi_example.c: a plain C program compiles against include/dm_b200.h, links libdm_b200.so and
    runs -- without a GPU it must report the library's error text and exit 0 (no CPU path, no crash)."""
    exe = tmp_path / "dm_example"
    pkg = ROOT / "image_compression_analysis_b200"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                        str(ROOT / "examples" / "c_abi_example.c"), f"-L{pkg}", "-ldm_b200", f"-Wl,-rpath,{pkg}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"ABI {built_lib.ABI_VERSION}" in r.stdout
