"""Device-side evaluation of one original/decoded pair: buffers, kernel launches, combine.

torch is plumbing here (device memory, streams, torch.distributed); every number comes from
libdm_b200.so.  One `Partials` object holds all outputs of a pair in THREE flat device vectors so
that multi-GPU combination is three allreduces and the host read-back is one copy:

    isum  int64   [ sums B*8 | hist B*K | counts 3 | hist8_g 256 | hist8_z 256 ]      (SUM)
    imax  int64   [ maxs B*8 ]                                                         (MAX)
    fsum  float64 [ sum_acos, sum_sid, n_spec | lmse B | ssimw_sum B | ssimw_cnt B ]   (SUM)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import (DM_BIP, DM_BSQ, DM_I16, DM_NSTAT, DM_U8, DM_U16, DM_VALID_METRICS, DM_VALID_QUICKLOOK,
                   DM_VALID_SPECTRAL, DmPair, check, lib)

_DTYPE_CODES = {"uint8": DM_U8, "uint16": DM_U16, "int16": DM_I16}
_TORCH_STORE = {"uint8": torch.uint8, "uint16": torch.int16, "int16": torch.int16}   # bytes only


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("image_compression_analysis_b200 needs a CUDA device: the distortion metrics "
                           "have no CPU implementation")
    return torch.device("cuda", torch.cuda.current_device())


def dtype_code(np_dtype) -> int:
    name = np.dtype(np_dtype).name
    if name not in _DTYPE_CODES:
        raise TypeError(f"unsupported sample type {name}: the GPU path handles uint8, uint16 and int16")
    return _DTYPE_CODES[name]


def integral_nodata(nodata, np_dtype) -> Optional[int]:
    """ds.nodata as an integer the samples can equal, else None (a value no sample can take
    never matches `band != nodata`, run_codec.py:250-259)."""
    if nodata is None:
        return None
    try:
        f = float(nodata)
    except (TypeError, ValueError):
        return None
    if not np.isfinite(f) or f != int(f):
        return None
    info = np.iinfo(np.dtype(np_dtype))
    v = int(f)
    return v if info.min <= v <= info.max else None


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def to_device(arr, device=None, non_blocking: bool = True) -> torch.Tensor:
    """numpy array / torch CPU tensor -> device tensor of the same bytes (uint16 stored as int16)."""
    device = device or require_cuda()
    if isinstance(arr, torch.Tensor):
        t = arr
        if t.dtype == torch.uint16:
            t = t.view(torch.int16)
        return t.to(device, non_blocking=non_blocking) if t.device.type == "cpu" else t
    a = np.ascontiguousarray(arr)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    elif a.dtype == np.bool_:
        a = a.view(np.uint8)
    t = torch.from_numpy(a)
    return t.to(device, non_blocking=non_blocking)


def bind_host_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned staging buffers
    allocated afterwards are local to the GPU's PCIe root (host-to-device copies of several ranks otherwise
    cross the socket interconnect).  Returns the node, or None when the topology cannot be read or the
    affinity cannot be set (containers with a restricted cpuset): best effort, never an error."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


_UPLOAD_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def upload_pair(ref, tst, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Both cubes of a pair on the device.  Two PINNED host tensors go up on two streams at once: one
    host-to-device copy does not fill the link (measured on this pool: 51.7 GB/s for one stream, 55.6 GB/s
    for two concurrent copies, tools/probe_h2d.py); the current stream waits for the second copy."""
    device = device or require_cuda()
    if not (isinstance(ref, torch.Tensor) and isinstance(tst, torch.Tensor) and ref.device.type == "cpu"
            and tst.device.type == "cpu" and ref.is_pinned() and tst.is_pinned()):
        return to_device(ref, device), to_device(tst, device)
    hr = ref.view(torch.int16) if ref.dtype == torch.uint16 else ref
    ht = tst.view(torch.int16) if tst.dtype == torch.uint16 else tst
    main = torch.cuda.current_stream(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    side = _UPLOAD_STREAMS.get(idx)
    if side is None:
        side = _UPLOAD_STREAMS[idx] = torch.cuda.Stream(device)
    dr = torch.empty(hr.shape, dtype=hr.dtype, device=device)
    dt = torch.empty(ht.shape, dtype=ht.dtype, device=device)
    side.wait_stream(main)              # dt's memory may have earlier users on the current stream
    dr.copy_(hr, non_blocking=True)
    with torch.cuda.stream(side):
        dt.copy_(ht, non_blocking=True)
    main.wait_stream(side)              # (and dt is only ever used on the current stream afterwards)
    return dr, dt


@dataclass
class DevicePair:
    """An original/decoded pair resident in HBM (or a row strip of one)."""
    ref: torch.Tensor
    tst: torch.Tensor
    np_dtype: str                 # "uint8" | "uint16" | "int16"
    layout: str                   # "bsq" (B,H,W) | "bip" (H,W,B)
    bands: int
    rows: int
    width: int
    ref_nodata: Optional[int] = None
    tst_nodata: Optional[int] = None
    img_row0: int = 0             # image row of buffer row 0 (row strips)
    img_rows: Optional[int] = None
    band_stride: Optional[int] = None

    def __post_init__(self):
        if self.img_rows is None:
            self.img_rows = self.rows
        if self.band_stride is None:
            self.band_stride = self.rows * self.width

    @property
    def npix(self) -> int:
        return self.rows * self.width

    def c_pair(self) -> DmPair:
        p = DmPair()
        p.ref, p.tst = self.ref.data_ptr(), self.tst.data_ptr()
        p.dtype = _DTYPE_CODES[self.np_dtype]
        p.layout = DM_BSQ if self.layout == "bsq" else DM_BIP
        p.bands, p.rows, p.width, p.band_stride = self.bands, self.rows, self.width, self.band_stride
        p.ref_has_nodata = 0 if self.ref_nodata is None else 1
        p.ref_nodata = 0 if self.ref_nodata is None else int(self.ref_nodata)
        p.tst_has_nodata = 0 if self.tst_nodata is None else 1
        p.tst_nodata = 0 if self.tst_nodata is None else int(self.tst_nodata)
        return p

    @staticmethod
    def from_arrays(ref, tst, layout: str = "bsq", ref_nodata=None, tst_nodata=None, device=None) -> "DevicePair":
        """Upload a host pair.  ref/tst: numpy arrays or torch CPU tensors, (B,H,W) for "bsq",
        (H,W,B) for "bip"; shapes and dtypes must match (run_codec.py:243-244)."""
        shape_r, shape_t = tuple(ref.shape), tuple(tst.shape)
        assert shape_r == shape_t and len(shape_r) == 3, "Reference and test must match in size and band count."
        name = _np_name(ref)
        assert name == _np_name(tst), "Reference and test must have the same sample type."
        dtype_code(name)
        if layout == "bsq":
            B, H, W = shape_r
        elif layout == "bip":
            H, W, B = shape_r
        else:
            raise ValueError(f"layout must be 'bsq' or 'bip', not {layout!r}")
        dr, dt = upload_pair(ref, tst, device)
        return DevicePair(dr, dt, name, layout, B, H, W,
                          integral_nodata(ref_nodata, name), integral_nodata(tst_nodata, name))

    def as_bsq(self) -> "DevicePair":
        """Same pair in (B,H,W) layout (device transpose through dm_bip_to_bsq when needed)."""
        if self.layout == "bsq":
            return self
        eb = 1 if self.np_dtype == "uint8" else 2
        outs = []
        for t in (self.ref, self.tst):
            o = torch.empty((self.bands, self.rows, self.width), dtype=t.dtype, device=t.device)
            check(lib().dm_bip_to_bsq(_ptr(t), _ptr(o), eb, self.bands, self.rows, self.width, _stream_ptr()))
            outs.append(o)
        return DevicePair(outs[0], outs[1], self.np_dtype, "bsq", self.bands, self.rows, self.width,
                          self.ref_nodata, self.tst_nodata, self.img_row0, self.img_rows)


def _np_name(a) -> str:
    if isinstance(a, torch.Tensor):
        return {torch.uint8: "uint8", torch.int16: "int16", torch.uint16: "uint16"}.get(a.dtype, str(a.dtype))
    return np.dtype(a.dtype).name


@dataclass
class Want:
    """Which parts of the path to evaluate for a pair."""
    stats: bool = True            # compute_metrics partials (+ data range scan)
    moments: bool = True          # False: PSNR-only variant (no SSIM moments)
    hist_bins: int = 0            # per-band |d| histogram bins (0 = off, power of two <= 1024)
    errmax: bool = False          # uint16 plane of max_b |d|
    err8_caps: Tuple[Optional[float], Optional[float]] = (None, None)   # (global cap, zoom cap)
    sam: bool = False
    sid: bool = False
    lmse: bool = False
    ssim_gauss: bool = False
    generic_stats: bool = False   # force the scalar cross-check kernel
    fused: bool = True            # allow the one-pass BIP kernel (dm_fused_bip) when it applies
    fused_scan: bool = True       # allow its validity-folding build (dm_fused_bip_scan: nodata / caller mask, ONE read)


@dataclass
class Partials:
    """Device-resident outputs of one pair (or one strip); see the module docstring."""
    bands: int
    hist_bins: int
    isum: torch.Tensor
    imax: torch.Tensor
    fsum: torch.Tensor
    planes: Dict[str, torch.Tensor] = field(default_factory=dict)   # errmax / err8_g / err8_z / valid
    np_dtype: str = "uint16"
    used_mask: bool = False
    flat: Optional[torch.Tensor] = None     # the buffer isum / imax / fsum are views of

    # layout helpers ------------------------------------------------------------------------
    @staticmethod
    def sizes(bands: int, hist_bins: int) -> Tuple[int, int, int]:
        return bands * DM_NSTAT + bands * hist_bins + 3 + 512, bands * DM_NSTAT, 3 + 3 * bands

    @staticmethod
    def allocate(bands: int, hist_bins: int, device, np_dtype: str) -> "Partials":
        """One flat 8-byte-word buffer [isum | imax | fsum]; the three vectors are views of it, so the
        multi-GPU exchange is a single all-gather and the host read-back a single copy."""
        ni, nm, nf = Partials.sizes(bands, hist_bins)
        flat = torch.zeros(ni + nm + nf, dtype=torch.int64, device=device)
        P = Partials(bands, hist_bins, flat[:ni], flat[ni:ni + nm], flat[ni + nm:].view(torch.float64), {}, np_dtype)
        P.flat = flat
        return P

    @staticmethod
    def allocate_run(n: int, bands: int, hist_bins: int, device, np_dtype: str):
        """`n` zeroed partial vectors in ONE contiguous buffer (the pairs of a sweep): returns
        (run, [Partials...]) where every Partials is a view of run[i].  A contiguous range of the run
        combines across GPUs with one all-gather (`allreduce_run_`)."""
        ni, nm, nf = Partials.sizes(bands, hist_bins)
        L = ni + nm + nf
        run = torch.zeros((n, L), dtype=torch.int64, device=device)
        out = []
        for i in range(n):
            flat = run[i]
            P = Partials(bands, hist_bins, flat[:ni], flat[ni:ni + nm], flat[ni + nm:].view(torch.float64), {}, np_dtype)
            P.flat = flat
            out.append(P)
        return run, out

    @staticmethod
    def allreduce_run_(run: torch.Tensor, i0: int, i1: int, bands: int, hist_bins: int, group=None) -> None:
        """Combine the partial vectors run[i0:i1] of all ranks in place with ONE all-gather +
        dm_combine_partials: the exchange is latency bound, so a sweep batches several pairs per call."""
        import torch.distributed as dist
        if i1 <= i0 or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        world = dist.get_world_size(group)
        ni, nm, nf = Partials.sizes(bands, hist_bins)
        part = run[i0:i1]
        if not run.is_cuda:                     # gloo (the CPU tests): three all-reduces on the column blocks
            a, b, c = part[:, :ni].contiguous(), part[:, ni:ni + nm].contiguous(), part[:, ni + nm:].contiguous().view(torch.float64)
            dist.all_reduce(a, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(b, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
            part[:, :ni], part[:, ni:ni + nm], part[:, ni + nm:] = a, b, c.view(torch.int64)
            return
        gathered = torch.empty((world,) + tuple(part.shape), dtype=torch.int64, device=run.device)
        dist.all_gather_into_tensor(gathered, part, group=group)
        check(lib().dm_combine_partials(_ptr(gathered), world, i1 - i0, ni, nm, nf, _ptr(part), _stream_ptr()))

    def zero_(self) -> "Partials":
        self.flat.zero_()
        return self

    def _o(self):
        B, K = self.bands, self.hist_bins
        o_hist = B * DM_NSTAT
        o_cnt = o_hist + B * K
        return o_hist, o_cnt, o_cnt + 3, o_cnt + 3 + 256

    @property
    def sums(self): return self.isum[: self.bands * DM_NSTAT]
    @property
    def hist(self):
        o_hist, o_cnt, _, _ = self._o()
        return self.isum[o_hist:o_cnt]
    @property
    def counts(self):
        _, o_cnt, o_g, _ = self._o()
        return self.isum[o_cnt:o_g]
    @property
    def hist8_g(self):
        _, _, o_g, o_z = self._o()
        return self.isum[o_g:o_z]
    @property
    def hist8_z(self):
        _, _, _, o_z = self._o()
        return self.isum[o_z:o_z + 256]
    @property
    def spec(self): return self.fsum[0:3]
    @property
    def lmse(self): return self.fsum[3:3 + self.bands]
    @property
    def ssimw_sum(self): return self.fsum[3 + self.bands:3 + 2 * self.bands]
    @property
    def ssimw_cnt(self): return self.fsum[3 + 2 * self.bands:3 + 3 * self.bands]

    def allreduce_(self, group=None) -> "Partials":
        """Combine the partials of all ranks in place: int64 SUM, int64 MAX, float64 SUM.

        The only exchange step of the path: payload B*(8+K+8+3)*8 bytes, latency
        bound.  On CUDA tensors it is ONE NCCL all-gather of the flat buffer followed by
        dm_combine_partials (float64 sums in rank order: bit-identical on every rank); on CPU
        tensors (gloo, the CPU tests) three all-reduces."""
        import torch.distributed as dist
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return self
        world = dist.get_world_size(group)
        if self.flat is not None and self.flat.is_cuda:
            gathered = torch.empty(world * self.flat.numel(), dtype=torch.int64, device=self.flat.device)
            dist.all_gather_into_tensor(gathered, self.flat, group=group)
            check(lib().dm_combine_partials(_ptr(gathered), world, 1, self.isum.numel(), self.imax.numel(),
                                            self.fsum.numel(), _ptr(self.flat), _stream_ptr()))
            return self
        works = [dist.all_reduce(self.isum, op=dist.ReduceOp.SUM, group=group, async_op=True),
                 dist.all_reduce(self.imax, op=dist.ReduceOp.MAX, group=group, async_op=True),
                 dist.all_reduce(self.fsum, op=dist.ReduceOp.SUM, group=group, async_op=True)]
        for w in works:
            w.wait()
        return self

    def to_host(self) -> "HostPartials":
        flat = (self.flat if self.flat is not None
                else torch.cat([self.isum, self.imax, self.fsum.view(torch.int64)])).cpu().numpy()
        ni, nm = self.isum.numel(), self.imax.numel()
        return HostPartials(self.bands, self.hist_bins, flat[:ni].copy(), flat[ni:ni + nm].copy(),
                            flat[ni + nm:].view(np.float64).copy(), self.np_dtype, self.used_mask)


@dataclass
class HostPartials:
    bands: int
    hist_bins: int
    isum: np.ndarray
    imax: np.ndarray
    fsum: np.ndarray
    np_dtype: str = "uint16"
    used_mask: bool = False

    @property
    def sums(self): return self.isum[: self.bands * DM_NSTAT].reshape(self.bands, DM_NSTAT)
    @property
    def maxs(self): return self.imax.reshape(self.bands, DM_NSTAT)
    @property
    def hist(self):
        o = self.bands * DM_NSTAT
        return self.isum[o:o + self.bands * self.hist_bins].reshape(self.bands, max(self.hist_bins, 0)) \
            if self.hist_bins else None
    @property
    def counts(self):
        o = self.bands * DM_NSTAT + self.bands * self.hist_bins
        return self.isum[o:o + 3]
    @property
    def hist8_g(self):
        o = self.bands * DM_NSTAT + self.bands * self.hist_bins + 3
        return self.isum[o:o + 256]
    @property
    def hist8_z(self):
        o = self.bands * DM_NSTAT + self.bands * self.hist_bins + 3 + 256
        return self.isum[o:o + 256]
    @property
    def spec(self): return self.fsum[0:3]
    @property
    def lmse(self): return self.fsum[3:3 + self.bands]
    @property
    def ssimw_sum(self): return self.fsum[3 + self.bands:3 + 2 * self.bands]
    @property
    def ssimw_cnt(self): return self.fsum[3 + 2 * self.bands:3 + 3 * self.bands]


_LUT_CACHE: Dict[Tuple[float, int], torch.Tensor] = {}


def _lut_on_device(cap: float, device) -> torch.Tensor:
    from .finish import err8_lut
    key = (float(cap), device.index if device.index is not None else 0)
    t = _LUT_CACHE.get(key)
    if t is None:
        t = torch.from_numpy(err8_lut(cap)).to(device)
        _LUT_CACHE[key] = t
    return t


_WORKSPACES: Dict[Tuple[int, int], torch.Tensor] = {}


def workspace(device) -> torch.Tensor:
    """The zeroed device scratch dm_spectral / dm_fused_bip need (dm_workspace_bytes()), one per
    (device, stream): launches on one stream share it, the kernels reset it themselves."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream().cuda_stream)
    w = _WORKSPACES.get(key)
    if w is None:
        w = torch.zeros(int(lib().dm_workspace_bytes()) // 8 + 1, dtype=torch.int64, device=device)
        _WORKSPACES[key] = w
    return w


_SCRATCH: Dict[Tuple[int, int, str], torch.Tensor] = {}


def _scratch(device, what: str, n_f64: int) -> torch.Tensor:
    """Per-(device, stream) float64 scratch of the stencil kernels' block partials (contents irrelevant between
    launches; launches on one stream are ordered): allocated once, grown on demand -- no allocation in a sweep."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream().cuda_stream, what)
    t = _SCRATCH.get(key)
    if t is None or t.numel() < n_f64:
        t = _SCRATCH[key] = torch.empty(max(n_f64, 1), dtype=torch.float64, device=device)
    return t


def needs_plane(pair: DevicePair, valid) -> bool:
    return valid is not None or pair.ref_nodata is not None or pair.tst_nodata is not None


class PreparedFused:
    """A prepared launch of the one-pass BIP kernel (dm_fused_bip: per-band stats [+ SAM]) for one pair and
    one output vector: every ctypes argument is built once, `launch()` is a single foreign call.
    For sweeps that evaluate thousands of pairs, where the per-call Python work of `evaluate` (a few
    tens of microseconds) would otherwise rival the 150 us kernel when several ranks share a host."""

    def __init__(self, pair: DevicePair, want: Want, out: "Partials", plane: Optional[torch.Tensor] = None,
                 scan: bool = False, valid: Optional[torch.Tensor] = None, plane_out: Optional[torch.Tensor] = None):
        """plane: a validity plane computed earlier (dm_validity).  scan=True instead: the pair carries nodata values
        and / or `valid` is the caller's mask, and the kernel derives the validity itself while it reads the pair
        (dm_fused_bip_scan, 180-band cubes; counts land in out.counts, the plane in plane_out when given).  The
        reference's all-False-mask rule (run_codec.py:264) is the caller's to apply on out.counts[0]."""
        if pair.layout != "bip" or want.hist_bins or want.sid or want.lmse or want.ssim_gauss or want.errmax \
                or want.err8_caps != (None, None) or not want.stats:
            raise ValueError("PreparedFused covers stats (+ SAM) on BIP cubes; use evaluate() for the rest")
        if scan and (plane is not None or pair.bands != 180 or (plane_out is None and pair.npix % 64)):
            raise ValueError("scan=True: 180-band cubes, no precomputed plane, plane_out unless rows*width % 64 == 0")
        self._keep = (pair, out, plane, workspace(pair.ref.device) if want.sam else None, valid, plane_out)
        self._cp = pair.c_pair()
        self._chain = lib().dm_launch_chaining
        ws = self._keep[3]
        tail = (_ptr(out.sums), _ptr(out.imax), None, None, 0, None, None,
                None, 0, None, None, 1 if want.sam else 0, _ptr(out.spec), _ptr(ws), _stream_ptr())
        if scan:
            self._fn = lib().dm_fused_bip_scan
            self._args = (C.byref(self._cp), _ptr(valid), _ptr(plane_out), _ptr(out.counts), *tail)
        else:
            self._fn = lib().dm_fused_bip
            self._args = (C.byref(self._cp), _ptr(plane), *tail)
        out.used_mask = plane is not None or scan

    def launch(self, chain: bool = True) -> None:
        """chain=True: this launch may start while the previous kernel on the stream drains (programmatic
        dependent launch; see dm_launch_chaining in include/dm_b200.h).  The cubes of the pair must not be the
        output of the kernel issued immediately before this launch on the same stream -- true for a sweep over
        resident cubes, which is what this class is for; pass chain=False otherwise."""
        if chain:
            self._chain(1)
            try:
                check(self._fn(*self._args))
            finally:
                self._chain(0)
        else:
            check(self._fn(*self._args))


class PreparedStats:
    """A prepared launch of dm_fused_stats for one unmasked pair and one output vector (every ctypes argument built
    once; `launch()` is a single foreign call).  For latency-critical single tiles: a Case-A tile's kernel is a few
    microseconds, the per-call Python work of `evaluate` several times that."""

    def __init__(self, pair: DevicePair, out: "Partials", moments: bool = True, hist_bins: int = 0):
        if needs_plane(pair, None):
            raise ValueError("PreparedStats covers unmasked pairs; use evaluate() for the rest")
        if hist_bins != out.hist_bins and hist_bins:
            raise ValueError("hist_bins must match the partial vector's")
        self._keep = (pair, out)
        self._cp = pair.c_pair()
        self._fn = lib().dm_fused_stats
        flags = 0 if moments else _lib.DM_STATS_NO_MOMENTS
        self._args = (C.byref(self._cp), None, DM_VALID_METRICS, hist_bins, flags, _ptr(out.sums), _ptr(out.imax),
                      _ptr(out.hist) if hist_bins else None, _stream_ptr())

    def launch(self) -> None:
        check(self._fn(*self._args))


class PreparedCaseAAll:
    """Every Case-A metric of one (strip of a) BSQ image, prepared once and launched without any allocation:
    per-band statistics + both ERR8 planes in ONE pass (dm_fused_bsq), per-band |d| histograms (dm_fused_stats,
    statistics-light variant; only its histogram is kept), Gaussian-window SSIM (dm_ssim_gauss).  Three foreign
    calls per launch; planes, scratch and the histogram pass's throw-away statistics are owned by the object.

    core   the COUNTED rows (a view of `full`'s buffers: same storage, band_stride of the buffer)
    full   the resident rows including the halo the SSIM window needs; rows = counted rows in buffer coordinates"""

    def __init__(self, core: DevicePair, full: DevicePair, rows: Tuple[int, int], out: "Partials", data_range: float,
                 err8_caps=(255, 32), hist_bins: int = 256, side_stream: Optional["torch.cuda.Stream"] = None):
        """side_stream: run the two HBM-bound passes (statistics + planes, histograms) there, BEHIND the FP64-bound SSIM
        kernel that is issued first on the current stream: the three kernels write disjoint outputs, so the short
        passes run in the SSIM kernel's shadow (its blocks leave registers and shared memory for one more block per SM)
        instead of in front of it; launch() forks and joins the side stream with events."""
        dev = core.ref.device
        L = lib()
        self._keep = (core, full, out)
        self._cc, self._cf = core.c_pair(), full.c_pair()
        self.planes = {"err8_g": torch.empty(core.npix, dtype=torch.uint8, device=dev),
                       "err8_z": torch.empty(core.npix, dtype=torch.uint8, device=dev)}
        lut_g, lut_z = _lut_on_device(err8_caps[0], dev), _lut_on_device(err8_caps[1], dev)
        self._luts = (lut_g, lut_z)
        st = _stream_ptr()
        self._side = side_stream
        st_hbm = C.c_void_p(side_stream.cuda_stream) if side_stream is not None else st
        self._ws = workspace(dev)
        self._a_bsq = (C.byref(self._cc), None, _ptr(out.sums), _ptr(out.imax), None,
                       _ptr(lut_g), lut_g.numel() - 1, _ptr(self.planes["err8_g"]), _ptr(out.hist8_g),
                       _ptr(lut_z), lut_z.numel() - 1, _ptr(self.planes["err8_z"]), _ptr(out.hist8_z), st_hbm)
        self._junk = torch.zeros(2 * core.bands * DM_NSTAT, dtype=torch.int64, device=dev)
        self._a_hist = (C.byref(self._cc), None, DM_VALID_METRICS, hist_bins, _lib.DM_STATS_NO_MOMENTS, _ptr(self._junk),
                        _ptr(self._junk[core.bands * DM_NSTAT:]), _ptr(out.hist), st_hbm)
        self._scr = _scratch(dev, "ssim", full.bands * L.dm_ssim_nblocks() * 2)
        self._a_ssim = (C.byref(self._cf), float(data_range), rows[0], rows[1], full.img_row0, full.img_rows,
                        _ptr(self._scr), _ptr(out.ssimw_sum), _ptr(out.ssimw_cnt), _ptr(self._ws), st)
        self._hist_bins = hist_bins
        self._L = L

    def launch(self) -> None:
        L = self._L
        if self._side is None:
            check(L.dm_fused_bsq(*self._a_bsq))
            if self._hist_bins:
                check(L.dm_fused_stats(*self._a_hist))
            check(L.dm_ssim_gauss(*self._a_ssim))
            return
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        self._side.wait_event(fork)
        check(L.dm_ssim_gauss(*self._a_ssim))
        check(L.dm_fused_bsq(*self._a_bsq))
        if self._hist_bins:
            check(L.dm_fused_stats(*self._a_hist))
        join = torch.cuda.Event()
        join.record(self._side)
        cur.wait_event(join)


class PreparedStatsBatch:
    """A prepared ONE-LAUNCH evaluation of the compute_metrics statistics of many pairs of one geometry (the tiles
    of a manifest / the decoded tiles of a rate sweep; dm_fused_stats_batch).  A Case-A tile pair is 16.8 MB: 2.6 us
    at HBM speed, less than a kernel launch, so pair-by-pair evaluation is launch bound; here all pairs share one
    launch and the host makes one foreign call.  BSQ cubes, no mask, no histogram (evaluate() serves the rest)."""

    def __init__(self, pairs, outs, moments: bool = True):
        pairs, outs = list(pairs), list(outs)
        if not pairs or len(pairs) != len(outs):
            raise ValueError("PreparedStatsBatch needs as many partial vectors as pairs (at least one)")
        p0 = pairs[0]
        for q in pairs:
            if (q.layout, q.np_dtype, q.bands, q.rows, q.width, q.band_stride) != (p0.layout, p0.np_dtype, p0.bands, p0.rows,
                                                                                   p0.width, p0.band_stride):
                raise ValueError("the pairs of a batch must share one geometry and sample type")
            if q.layout != "bsq" or needs_plane(q, None):
                raise ValueError("PreparedStatsBatch covers unmasked BSQ pairs; use evaluate() for the rest")
            if (q.ref.data_ptr() | q.tst.data_ptr()) & 15:
                raise ValueError("the cubes of a batch must be 16-byte aligned")
        items = np.empty((len(pairs), 4), dtype=np.int64)
        for i, (q, o) in enumerate(zip(pairs, outs)):
            items[i] = (q.ref.data_ptr(), q.tst.data_ptr(), o.sums.data_ptr(), o.imax.data_ptr())
        self._items = torch.from_numpy(items).to(p0.ref.device)        # dm_batch_item_t[n] on the device
        self._cp = p0.c_pair()
        self._keep = (pairs, outs)
        self._n = len(pairs)
        self._flags = 0 if moments else _lib.DM_STATS_NO_MOMENTS
        self._fn = lib().dm_fused_stats_batch
        torch.cuda.current_stream().synchronize()                       # the item array is resident before the first launch

    def launch(self) -> None:
        check(self._fn(C.byref(self._cp), _ptr(self._items), self._n, self._flags, _stream_ptr()))


def evaluate(pair: DevicePair, want: Want, valid: Optional[torch.Tensor] = None,
             out: Optional[Partials] = None, rows: Optional[Tuple[int, int]] = None,
             data_range: Optional[float] = None, metrics_mask: bool = True,
             plane: Optional[torch.Tensor] = None) -> Partials:
    """Launch the kernels `want` asks for on the current stream; nothing is synchronised.

    valid     optional device uint8 plane (rows*width, nonzero = valid): the caller's `valid`
    out       accumulate into existing partials (row strips of one image on one GPU)
    rows      buffer rows [begin,end) this call COUNTS for the stencil kernels; the buffer may hold
              halo rows around them (img_row0 / img_rows in `pair` place the buffer in the image)
    metrics_mask  False: run the fused stats unmasked even if a plane exists (the reference's
              all-False-mask fallback, run_codec.py:264)
    plane     a validity plane already computed for this buffer (skips dm_validity)
    """
    L = lib()
    dev = pair.ref.device
    st = _stream_ptr()
    P = out if out is not None else Partials.allocate(pair.bands, want.hist_bins, dev, pair.np_dtype)
    cp = pair.c_pair()
    cap_g, cap_z = want.err8_caps
    want_planes = want.errmax or cap_g is not None or cap_z is not None
    # the one-pass BIP kernel applies (see below); with a nodata value or a caller mask and EnMAP's 180 bands its
    # validity-folding build reads the pair ONCE where dm_validity + dm_fused_bip read it twice
    bip_one_pass = (want.stats and (want.moments or pair.bands == 180) and not want.hist_bins and not want.generic_stats
                    and not want.sid and (want_planes or want.sam or pair.bands == 180) and pair.layout == "bip"
                    and want.fused)
    scan = (plane is None and needs_plane(pair, valid) and metrics_mask and bip_one_pass and want.fused_scan
            and pair.bands == 180 and pair.npix >= 64 and pair.np_dtype in ("uint16", "int16")
            and (valid is None or valid.data_ptr() % 16 == 0))
    if plane is None and needs_plane(pair, valid):
        plane = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
        if not scan:
            check(L.dm_validity(C.byref(cp), _ptr(valid), _ptr(plane), _ptr(P.counts), st))
        P.planes["valid"] = plane
    use = plane if (plane is not None and metrics_mask) else None
    lut_g = lut_z = None
    pl_e = pl_g = pl_z = None               # the planes THIS call writes (P may carry planes of an earlier call)
    ws = workspace(dev) if (want.sam or want.sid) else None
    if want_planes or want.sam or want.sid:
        if want.errmax:
            pl_e = P.planes["errmax"] = torch.empty(pair.npix, dtype=torch.int16, device=dev)
        if cap_g is not None:
            lut_g = _lut_on_device(cap_g, dev)
            pl_g = P.planes["err8_g"] = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
        if cap_z is not None:
            lut_z = _lut_on_device(cap_z, dev)
            pl_z = P.planes["err8_z"] = torch.empty(pair.npix, dtype=torch.uint8, device=dev)
    spectral_args = (_ptr(pl_e),
                     _ptr(lut_g), 0 if lut_g is None else lut_g.numel() - 1, _ptr(pl_g), _ptr(P.hist8_g),
                     _ptr(lut_z), 0 if lut_z is None else lut_z.numel() - 1, _ptr(pl_z), _ptr(P.hist8_z))
    done_stats = done_spectral = False
    if scan:
        rc = L.dm_fused_bip_scan(C.byref(cp), _ptr(valid), _ptr(plane), _ptr(P.counts), _ptr(P.sums), _ptr(P.imax),
                                 *spectral_args, 1 if want.sam else 0, _ptr(P.spec), _ptr(ws), st)
        if rc == _lib.DM_OK:
            done_stats = done_spectral = True
            P.used_mask = True
        elif rc == _lib.DM_EUNSUPPORTED:
            check(L.dm_validity(C.byref(cp), _ptr(valid), _ptr(plane), _ptr(P.counts), st))
        else:
            check(rc)
    # one-pass BIP kernel: stats + error planes + SAM from a single read (the plane selects METRICS
    # pixels for the stats and QUICKLOOK / SPECTRAL pixels for the rest, so it needs use == plane)
    # (for EnMAP's 180 bands it is also the fastest stats-only kernel: its pixel warps then idle)
    if bip_one_pass and use is plane and not done_stats:
        rc = L.dm_fused_bip(C.byref(cp), _ptr(plane), _ptr(P.sums), _ptr(P.imax), *spectral_args,
                            1 if want.sam else 0, _ptr(P.spec), _ptr(ws), st)
        if rc == _lib.DM_OK:
            done_stats = done_spectral = True
            P.used_mask = use is not None
        elif rc != _lib.DM_EUNSUPPORTED:
            check(rc)
    # one-pass BSQ kernel for few-band cubes (Case A): stats + error planes from a single read
    if (want.stats and not done_stats and want.moments and not want.hist_bins and not want.generic_stats
            and not want.sam and not want.sid and want_planes and pair.layout == "bsq" and pair.bands <= 4
            and use is plane and want.fused):
        rc = L.dm_fused_bsq(C.byref(cp), _ptr(plane), _ptr(P.sums), _ptr(P.imax), *spectral_args, st)
        if rc == _lib.DM_OK:
            done_stats = done_spectral = True
            P.used_mask = use is not None
        elif rc != _lib.DM_EUNSUPPORTED:
            check(rc)
    if want.stats and not done_stats:
        flags = (0 if want.moments else _lib.DM_STATS_NO_MOMENTS) | (_lib.DM_STATS_GENERIC if want.generic_stats else 0)
        check(L.dm_fused_stats(C.byref(cp), _ptr(use), DM_VALID_METRICS, want.hist_bins, flags,
                               _ptr(P.sums), _ptr(P.imax), _ptr(P.hist) if want.hist_bins else None, st))
        P.used_mask = use is not None
    if (want_planes or want.sam or want.sid) and not done_spectral:
        check(L.dm_spectral(C.byref(cp), _ptr(plane), *spectral_args,
                            1 if want.sam else 0, 1 if want.sid else 0, _ptr(P.spec), _ptr(ws), st))
    if want.lmse or want.ssim_gauss:
        r0, r1 = rows if rows is not None else (0, pair.rows)
        ws = ws if ws is not None else workspace(dev)
        bsq = None
        if want.lmse:
            # the kernels add their block partials in a fixed order themselves (the blocks that finish last) and
            # accumulate into P.lmse: no reduction kernels behind them.  BIP cubes of 16-bit samples go straight in
            # (a thread owns two bands of the spectrum); the rest of the layouts are transposed first
            buf = _scratch(dev, "sobel", pair.bands * L.dm_sobel_nblocks())
            rc = L.dm_sobel_lmse(C.byref(cp), r0, r1, pair.img_row0, pair.img_rows, _ptr(buf), _ptr(P.lmse), _ptr(ws), st) \
                if pair.layout == "bip" else _lib.DM_EUNSUPPORTED
            if rc == _lib.DM_EUNSUPPORTED:
                bsq = pair.as_bsq()
                cb = bsq.c_pair()
                rc = L.dm_sobel_lmse(C.byref(cb), r0, r1, pair.img_row0, pair.img_rows, _ptr(buf), _ptr(P.lmse), _ptr(ws), st)
            check(rc)
        if want.ssim_gauss and bsq is None:
            bsq = pair.as_bsq()
            cb = bsq.c_pair()
        if want.ssim_gauss:
            if data_range is None:
                raise ValueError("ssim_gauss needs data_range (the peak L of the SSIM constants)")
            buf = _scratch(dev, "ssim", pair.bands * L.dm_ssim_nblocks() * 2)
            check(L.dm_ssim_gauss(C.byref(cb), float(data_range), r0, r1, pair.img_row0, pair.img_rows, _ptr(buf),
                                  _ptr(P.ssimw_sum), _ptr(P.ssimw_cnt), _ptr(ws), st))
    return P


def evaluate_host_pairs(pairs, want: Want, layout: str = "bip", np_dtype: Optional[str] = None, group=None,
                        share_ref: bool = False):
    """Pipelined evaluation of a sweep of HOST pairs (the decoded cubes of a rate sweep): a generator that takes
    (ref, tst) pinned host tensors and yields one HostPartials per pair, in order.

    A pair's life is upload (PCIe, ~13.6 ms for a Case-B pair), kernels (0.14 ms), exchange with the other
    ranks if any, read-back of the 31.5 KB partial vector and the host-side finish.  Done one pair at a time,
    everything after the upload leaves the link idle; here pair i+1 is uploaded on a second stream while pair
    i computes and pair i-1 is finished on the host, so the link -- the only real bound of the end-to-end
    path -- never waits.  Results are identical to evaluate() + to_host() pair by pair.

    share_ref=True: a rate sweep compares MANY decoded cubes with ONE original (run_codec.py:472-475 loops rates and
    reps over the same src_path).  When consecutive items carry the same pinned original (same storage), it is
    uploaded once and stays resident; only the decoded cube crosses the link, which halves the bytes per pair."""
    from collections import deque
    dev = require_cuda()
    comp = torch.cuda.current_stream(dev)
    up = torch.cuda.Stream(dev)
    it = iter(pairs)

    # three device buffer pairs owned by the sweep and reused in turn (pair k lives in slot k % 3): no allocator
    # in the loop -- a cudaMalloc for a 377 MB cube would synchronise the device in the middle of the pipeline
    slots = [None, None, None]              # [dr, dt, event "kernels that read this slot are done"]
    counter = [0]
    shared = {"key": None, "dev": None}     # share_ref: the resident original

    def start_upload(item):
        ref, tst = item
        name = np_dtype or _np_name(ref)
        dtype_code(name)
        hr = ref.view(torch.int16) if ref.dtype == torch.uint16 else ref
        ht = tst.view(torch.int16) if tst.dtype == torch.uint16 else tst
        k = counter[0] % 3
        counter[0] += 1
        slot = slots[k]
        if slot is None or slot[1].shape != ht.shape or slot[1].dtype != ht.dtype:
            slot = slots[k] = [None if share_ref else torch.empty(hr.shape, dtype=hr.dtype, device=dev),
                               torch.empty(ht.shape, dtype=ht.dtype, device=dev), None]
            up.wait_stream(comp)            # fresh memory may have had users on the compute stream
        new_ref = None
        if share_ref:
            key = (hr.data_ptr(), tuple(hr.shape), hr.dtype)
            if shared["key"] != key:
                # a new original gets its own buffer, allocated on the compute stream like the slots (the previous
                # one may still be read by queued kernels; the DevicePairs that reference it keep it alive)
                shared["dev"], shared["key"] = torch.empty(hr.shape, dtype=hr.dtype, device=dev), key
                up.wait_stream(comp)
                new_ref = shared["dev"]
        dref = shared["dev"] if share_ref else slot[0]
        with torch.cuda.stream(up):
            if slot[2] is not None:
                up.wait_event(slot[2])      # the pair that lived here three uploads ago has been evaluated
            if new_ref is not None:
                new_ref.copy_(hr, non_blocking=True)
            elif not share_ref:
                slot[0].copy_(hr, non_blocking=True)
            slot[1].copy_(ht, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(up)
        shape = tuple(hr.shape)
        B, H, W = (shape[0], shape[1], shape[2]) if layout == "bsq" else (shape[2], shape[0], shape[1])
        return DevicePair(dref, slot[1], name, layout, B, H, W), ev, slot

    def finish_one(entry):
        P, host, ev = entry
        ev.synchronize()
        flat = host.numpy()
        ni, nm = P.isum.numel(), P.imax.numel()
        return HostPartials(P.bands, P.hist_bins, flat[:ni].copy(), flat[ni:ni + nm].copy(),
                            flat[ni + nm:].view(np.float64).copy(), P.np_dtype, P.used_mask)

    inflight = deque()
    ring, slot = [None, None, None], 0      # pinned read-back buffers (at most two results are in flight)
    try:
        nxt = start_upload(next(it))
    except StopIteration:
        return
    while nxt is not None:
        pair, ev, dslot = nxt
        try:
            nxt = start_upload(next(it))            # the next pair's copies queue up behind this pair's
        except StopIteration:
            nxt = None
        comp.wait_event(ev)
        P = evaluate(pair, want)
        dslot[2] = torch.cuda.Event()
        dslot[2].record(comp)                       # the slot may be overwritten once these kernels are done
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            P.allreduce_(group)
        host = ring[slot % len(ring)]
        if host is None or host.numel() != P.flat.numel():
            host = ring[slot % len(ring)] = torch.empty(P.flat.numel(), dtype=torch.int64).pin_memory()
        slot += 1
        host.copy_(P.flat, non_blocking=True)
        done = torch.cuda.Event()
        done.record(comp)
        inflight.append((P, host, done))
        if len(inflight) > 1:                        # finish pair i-1 while pair i is on the GPU
            yield finish_one(inflight.popleft())
    while inflight:
        yield finish_one(inflight.popleft())
