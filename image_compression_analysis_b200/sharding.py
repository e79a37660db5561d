"""Row-strip sharding of one image across the GPUs of a box, and the combine step.

Every metric of the path is a sum / max over independent pixels, so the only
exchange is an allreduce of the flat partial vectors (`Partials.allreduce_`): int64 SUM, int64 MAX
and float64 SUM, a few KB, latency bound on NVLink.  Stencil metrics (Sobel LMSE, Gaussian SSIM)
need halo rows: 1 and 5 rows on every interior strip edge, replicated when the strips are cut.

The functions here are pure host logic (testable on CPU with gloo) plus thin calls into engine.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

HALO_SOBEL = 1
HALO_SSIM = 5


@dataclass(frozen=True)
class Strip:
    """Rows [row0,row1) of the image are COUNTED by this rank; [buf0,buf1) are resident (with halo)."""
    rank: int
    row0: int
    row1: int
    buf0: int
    buf1: int

    @property
    def rows(self) -> int:
        return self.row1 - self.row0

    @property
    def count_range(self) -> Tuple[int, int]:
        """Counted rows in buffer coordinates."""
        return self.row0 - self.buf0, self.row1 - self.buf0


def strips(img_rows: int, world: int, halo: int = 0, align: int = 1) -> List[Strip]:
    """Cut img_rows into `world` contiguous strips, sizes differing by at most `align` rows.

    align > 1 keeps every strip start a multiple of `align` rows (row pitches that are not a
    multiple of 16 bytes then still give 16-byte aligned strips for the vector kernels)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    units = (img_rows + align - 1) // align
    out = []
    for r in range(world):
        u0, u1 = units * r // world, units * (r + 1) // world
        row0, row1 = min(u0 * align, img_rows), min(u1 * align, img_rows)
        buf0 = max(0, row0 - halo) if row1 > row0 else row0
        buf1 = min(img_rows, row1 + halo) if row1 > row0 else row1
        out.append(Strip(r, row0, row1, buf0, buf1))
    return out


def cut_bsq(cube: np.ndarray, s: Strip) -> np.ndarray:
    """(B,H,W) -> the strip's resident rows, contiguous."""
    return np.ascontiguousarray(cube[:, s.buf0:s.buf1, :])


def cut_bip(cube: np.ndarray, s: Strip) -> np.ndarray:
    """(H,W,B) -> the strip's resident rows (already contiguous)."""
    return cube[s.buf0:s.buf1]


def evaluate_strip(ref, tst, s: Strip, img_rows: int, layout: str, want, valid=None, *, ref_nodata=None,
                   tst_nodata=None, data_range: Optional[float] = None, group=None, reduce: bool = True):
    """Evaluate this rank's strip and (optionally) allreduce the partials.

    ref/tst: the strip's RESIDENT rows (with halo) as host arrays; `valid`: the caller's mask for
    the same rows or None.  Point-wise kernels see only the counted rows; stencil kernels see the
    whole buffer and count [row0,row1)."""
    import torch
    from .engine import DevicePair, Partials, Want, evaluate, to_device
    full = DevicePair.from_arrays(ref, tst, layout, ref_nodata, tst_nodata)
    full.img_row0, full.img_rows = s.buf0, img_rows
    c0, c1 = s.count_range
    W, B = full.width, full.bands
    # the counted rows as a view of the same device buffers
    if layout == "bsq":
        core = DevicePair(full.ref.view(-1)[c0 * W:], full.tst.view(-1)[c0 * W:], full.np_dtype, "bsq", B, c1 - c0, W,
                          full.ref_nodata, full.tst_nodata, s.row0, img_rows, band_stride=full.rows * W)
    else:
        core = DevicePair(full.ref.view(-1)[c0 * W * B:], full.tst.view(-1)[c0 * W * B:], full.np_dtype, "bip", B,
                          c1 - c0, W, full.ref_nodata, full.tst_nodata, s.row0, img_rows)
    vdev = None
    if valid is not None:
        vdev = to_device(np.ascontiguousarray(np.asarray(valid)[c0:c1].astype(bool)).reshape(-1))
    point = Want(stats=want.stats, moments=want.moments, hist_bins=want.hist_bins, errmax=want.errmax,
                 err8_caps=want.err8_caps, sam=want.sam, sid=want.sid, generic_stats=want.generic_stats)
    P = Partials.allocate(B, want.hist_bins, full.ref.device, full.np_dtype)
    if c1 > c0:
        evaluate(core, point, vdev, out=P)
        if want.lmse or want.ssim_gauss:
            evaluate(full, Want(stats=False, lmse=want.lmse, ssim_gauss=want.ssim_gauss), out=P, rows=(c0, c1),
                     data_range=data_range)
    # run_codec.py:264: `use_mask = np.any(vm)` -- a mask with NO valid pixel anywhere in the image means "use every
    # pixel".  That is a GLOBAL condition: a strip without valid pixels says nothing, so the valid counts of all
    # strips are added first (one extra 8-byte all-reduce, only when a mask exists at all; every rank takes part,
    # also one whose strip is empty) and the statistics are redone unmasked on every rank when the total is zero.
    masked = want.stats and (valid is not None or full.ref_nodata is not None or full.tst_nodata is not None)
    if reduce and masked:
        import torch.distributed as dist
        tot = P.counts[0:1].clone()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=group)
        if int(tot.item()) == 0:
            redo_stats_unmasked(P, core if c1 > c0 else None, point, vdev)
    if reduce:
        P.allreduce_(group)
    return P


def redo_stats_unmasked(P, core, want, vdev=None) -> None:
    """The reference's all-False-mask fallback (run_codec.py:264, 271-274) for one strip: drop the (empty) masked
    statistics and evaluate them again over every pixel.  Callers that combine strips themselves (reduce=False) call
    this on every strip once the SUM of the strips' P.counts[0] turns out to be zero."""
    from .engine import Want, evaluate
    P.sums.zero_()
    P.imax.zero_()
    if P.hist_bins:
        P.hist.zero_()
    if core is not None:
        w2 = Want(stats=True, moments=want.moments, hist_bins=want.hist_bins, generic_stats=want.generic_stats)
        evaluate(core, w2, vdev, out=P, metrics_mask=False, plane=P.planes.get("valid"))
    P.used_mask = False


class PipelinedCombiner:
    """Overlap the combine of pair i with the kernels of pair i+1 (a sweep of independent pairs).

    The allreduce of a pair's partials is a few KB and latency bound (~tens of microseconds), the
    same order as a 64th of a kernel; issued on the compute stream it would serialise with the next
    pair's kernels.  Here it runs on a side stream behind an event, so the compute stream only waits
    when it is about to REUSE a partials buffer whose combine has not finished."""

    def __init__(self, group=None):
        import torch
        self.group = group
        self.stream = torch.cuda.Stream()
        self._done = {}

    def before_reuse(self, P) -> None:
        """Call before zeroing / accumulating into P again on the compute stream."""
        import torch
        ev = self._done.pop(id(P), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def combine(self, P):
        """Allreduce P after everything queued so far on the current (compute) stream."""
        import torch
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            P.allreduce_(self.group)
            done = torch.cuda.Event()
            done.record()
        self._done[id(P)] = done
        return P

    def wait_all(self) -> None:
        import torch
        torch.cuda.current_stream().wait_stream(self.stream)
        self._done.clear()


class RunCombiner:
    """Combine the pairs of a sweep in batches: the partial vectors of the sweep live in one contiguous
    run (`Partials.allocate_run`), and every `batch` finished pairs are exchanged with ONE all-gather
    + dm_combine_partials on a side stream, behind an event, while the next pairs' kernels run.  The
    exchange is a few KB per pair and latency bound, so batching divides its cost (and the host's
    launch work) by `batch` without delaying anything a sweep needs before its end."""

    def __init__(self, run, bands: int, hist_bins: int = 0, batch: int = 8, group=None):
        import torch
        self.run, self.bands, self.hist_bins, self.batch, self.group = run, bands, hist_bins, max(1, batch), group
        self.stream = torch.cuda.Stream()
        self._next = 0          # first record not yet combined

    def done(self, i: int) -> None:
        """Record i (and everything before it) has been queued on the current stream."""
        if i + 1 - self._next >= self.batch:
            self._flush(i + 1)

    def _flush(self, upto: int) -> None:
        import torch
        from .engine import Partials
        if upto <= self._next:
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            Partials.allreduce_run_(self.run, self._next, upto, self.bands, self.hist_bins, self.group)
        self._next = upto

    def finish(self, upto: int) -> None:
        """Combine what is left up to record `upto` and make the current stream wait for it."""
        import torch
        self._flush(upto)
        torch.cuda.current_stream().wait_stream(self.stream)


class P2PRunCombiner:
    """RunCombiner without a collective library: the partial vectors of a sweep are PUSHED over NVLink into the
    exchange buffer every peer keeps (libdm_b200's dm_p2p_*: cudaMalloc + CUDA IPC, one small kernel with a CTA
    per destination, system-scope release/acquire flags), and each rank reduces its local copy in rank order.
    Same interface and same results as RunCombiner (bit-identical: the reduction order is the same).

    Set-up needs `torch.distributed` only to hand the 64-byte IPC handles around.  Anything that fails here
    raises, and the caller falls back to RunCombiner (NCCL)."""

    MAX_WORLD = 16          # kMaxWorld of csrc/p2p.cu

    def __init__(self, run, bands: int, hist_bins: int = 0, batch: int = 8, group=None, timeout_s: float = 20.0):
        import ctypes as C
        import torch
        import torch.distributed as dist
        from ._lib import check, lib
        from .engine import Partials
        self.run, self.bands, self.hist_bins, self.batch, self.group = run, bands, hist_bins, max(1, batch), group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > self.MAX_WORLD:     # before the collective handshake: every rank sees the same world size
            raise RuntimeError(f"P2P exchange supports up to {self.MAX_WORLD} ranks (dm_p2p_push), not {self.world}")
        self.capacity, self.words = int(run.shape[0]), int(run.shape[1])
        self.sizes = Partials.sizes(bands, hist_bins)
        assert sum(self.sizes) == self.words
        self.timeout_s = timeout_s
        self._next = 0
        self._L = lib()
        first_error = None
        data_bytes = self.world * self.capacity * self.words * 8
        self._flags_off = (data_bytes + 255) // 256 * 256
        total = self._flags_off + 256
        # Set-up is collective and must FAIL collectively: a rank that cannot allocate or map still takes part
        # in the two object gathers, and everybody raises together (the caller then falls back to NCCL).
        base, handle = C.c_void_p(), (C.c_ubyte * 64)()
        mine = None
        try:
            check(self._L.dm_p2p_alloc(total, C.byref(base), handle))
            mine = bytes(handle)
        except Exception as e:          # noqa: BLE001
            first_error = e
        self._base = base.value
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        self._peers, opened = [], True
        if all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    if r == self.rank:
                        self._peers.append(self._base)
                        continue
                    p = C.c_void_p()
                    check(self._L.dm_p2p_open((C.c_ubyte * 64).from_buffer_copy(h), C.byref(p)))
                    self._peers.append(p.value)
            except Exception as e:      # noqa: BLE001
                first_error, opened = e, False
        else:
            opened = False
        oks = [None] * self.world
        dist.all_gather_object(oks, opened, group=group)
        if not all(oks):
            for r, p in enumerate(self._peers):
                if r != self.rank:
                    self._L.dm_p2p_close(p)
            if self._base:
                self._L.dm_p2p_free(self._base)
            self._peers = []
            raise RuntimeError(f"P2P exchange set-up failed on rank(s) {[r for r, o in enumerate(oks) if not o]}"
                               + (f": {first_error}" if first_error is not None else ""))
        self.stream = torch.cuda.Stream()
        self._status = torch.zeros(1, dtype=torch.int32, device=run.device)
        dist.barrier(group=group)           # every buffer exists and is mapped before anybody pushes

    def done(self, i: int) -> None:
        if i + 1 - self._next >= self.batch:
            self._flush(i + 1)

    def _flush(self, upto: int) -> None:
        import ctypes as C
        import torch
        from ._lib import check
        if upto < self._next:
            raise RuntimeError(f"P2PRunCombiner: record {upto} is behind the {self._next} already exchanged -- "
                               "call reset() (collective) before a second sweep over the same run")
        if upto == self._next:
            return
        i0, n = self._next, upto - self._next
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            st = C.c_void_p(self.stream.cuda_stream)
            slot = (self.rank * self.capacity + i0) * self.words * 8
            dst = (C.c_void_p * self.world)(*[p + slot for p in self._peers])
            flg = (C.c_void_p * self.world)(*[p + self._flags_off + 8 * self.rank for p in self._peers])
            src = self.run.data_ptr() + i0 * self.words * 8
            check(self._L.dm_p2p_push(C.c_void_p(src), n * self.words, dst, flg, self.world, upto, st))
            ni, nm, nf = self.sizes
            check(self._L.dm_p2p_combine(C.c_void_p(self._base), C.c_void_p(self._base + self._flags_off), self.world, upto,
                                         self.capacity, i0, n, ni, nm, nf, C.c_void_p(src), C.c_void_p(self._status.data_ptr()),
                                         float(self.timeout_s), st))
        self._next = upto

    def finish(self, upto: int, check: bool = False) -> None:
        """Exchange what is left up to record `upto`; the current stream waits for it.  check=True also
        synchronises and raises if a combine timed out, so that a sweep cannot consume an uncombined run."""
        import torch
        self._flush(upto)
        torch.cuda.current_stream().wait_stream(self.stream)
        if check:
            torch.cuda.current_stream().synchronize()
            self.check_status()

    def check_status(self) -> None:
        """Raises if a combine gave up waiting for a peer (call after a synchronisation point)."""
        if int(self._status.item()) != 0:
            raise RuntimeError("P2P exchange: a peer's partial vectors did not arrive within the time-out")

    def reset(self) -> None:
        """Make the object reusable for another sweep over the same run: the arrival flags hold the number of
        records delivered so far and only ever grow, so they are zeroed -- collectively, between two barriers, so
        that no peer is still pushing the old sweep or already pushing the new one."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        self.check_status()
        dist.barrier(group=self.group)
        flags = (self._base + self._flags_off)
        import ctypes as C
        rc = self._L.dm_p2p_zero(C.c_void_p(flags), 256, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        from ._lib import check as _check
        _check(rc)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self._next = 0

    def close(self) -> None:
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        dist.barrier(group=self.group)      # nobody still reads or writes a buffer that is about to go away
        for r, p in enumerate(self._peers):
            if r != self.rank:
                self._L.dm_p2p_close(p)
        self._L.dm_p2p_free(self._base)
        self._peers = []
