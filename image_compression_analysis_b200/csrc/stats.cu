// Fused single-pass integer reduction: per-band {N, Sx, Sy, Sxx, Syy, Sxy, S|d|, SSE}, maxima and
// |d| histograms for an original/decoded cube pair.
//
// Replaces the arithmetic of the band loop of compute_metrics
// (/root/reference/tools/run_codec.py:268-285) and of effective_data_range (:86-117).
// Every quantity is an exact integer, so the host can finish PSNR / SSIM / MAE bit-for-bit.
//
// The kernels are HBM-bound integer work with a budget of roughly ten issue slots per sample pair
// at 6.5 TB/s, so the arithmetic is SIMD-in-word on PACKED pairs of 16-bit samples:
//   |x-y|            VIMNMX.U16x2 max, min, one 32-bit subtract (no borrow: max >= min per half)
//   sums             IDP.2A  (dp2a: two 16-bit x 8-bit products + 32-bit accumulate per issue);
//                    x*y = x*lo8(y) + 256*x*hi8(y), the byte split done by one PRMT per word
//   maxima           VIMNMX(3).U16x2 on packed running maxima
// 32-bit partials are bounded (<= 128 dp2a per quadratic partial) and spilled to 64-bit totals.
// int16 cubes are mapped to offset binary (x ^ 0x8000) so that the same unsigned arithmetic
// applies; the exact signed sums are restored from the unsigned ones when a partial is flushed.
//
// Kernels:
//   stats_bsq_packed  (B,H,W): a word = two neighbouring pixels of one band; 16-byte streaming
//                     loads, 8 in flight per thread, per-band warp-shuffle flush + 64-bit REDs;
//   stats_bip_packed  (H,W,B): a thread owns the same 2 or 4 bands for the whole kernel and pairs
//                     two PIXELS per word with PRMT, so the packed arithmetic above still applies;
//   stats_generic     any dtype / layout / alignment, one sample per step: the cross-check and
//                     the path for uint8 BIP and odd band counts.  Correct, not fast.

#include <type_traits>

#include "dm_common.cuh"

namespace dm {

namespace {

struct StatsArgs {
  const void* ref;
  const void* tst;
  const uint8_t* plane;   // may be null
  int plane_bit;
  int plane_shift;        // left shift that moves plane_bit to bit 7 of its byte
  int64_t bands, npix, band_stride;  // npix = rows*width
  int hist_bins;          // 0 or power of two
  int64_t* sums;
  int64_t* maxs;
  int64_t* hist;
  const dm_batch_item_t* items;   // batch launch: blockIdx.y selects {ref, tst, sums, maxs}; else null
};

// ------------------------------------------------------------------------------------------------
// packed SIMD-in-word primitives
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t vminu2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// prmt.b32 in its generic form: selector nibble bit 3 replicates the sign of the selected byte
// (__byte_perm only documents the low three bits of every nibble)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
__device__ __forceinline__ int hmax2(uint32_t packed) {   // horizontal max of the two u16 halves
  return max((int)(packed & 0xffffu), (int)(packed >> 16));
}
__device__ __forceinline__ int hmin2(uint32_t packed) {
  return min((int)(packed & 0xffffu), (int)(packed >> 16));
}
__device__ __forceinline__ int hmax2s(uint32_t packed) {  // horizontal max of two s16 halves
  return max((int)(short)(packed & 0xffffu), (int)(short)(packed >> 16));
}

// Per-band packed accumulators of one thread.
template <bool MOMENTS>
struct BandAcc {
  uint32_t sabs, sx, sy;                      // <= 2*65535 per dp2a
  uint32_t xxl, xxh, yyl, yyh, xyl, xyh;      // <= 2*65535*255 per dp2a -> spill every <= 128
  uint32_t maxd;                              // packed u16x2 running max |d|
  unsigned long long t_abs, t_x, t_y, t_xx, t_yy, t_xy;   // t_xx doubles as SSE when !MOMENTS

  __device__ __forceinline__ void reset() {
    sabs = sx = sy = 0; xxl = xxh = yyl = yyh = xyl = xyh = 0; maxd = 0;
    t_abs = t_x = t_y = t_xx = t_yy = t_xy = 0;
  }
  __device__ __forceinline__ void spill() {
    t_abs += sabs; sabs = 0;
    t_xx += (unsigned long long)xxl + ((unsigned long long)xxh << 8); xxl = xxh = 0;
    if (MOMENTS) {
      t_x += sx; t_y += sy; sx = sy = 0;
      t_yy += (unsigned long long)yyl + ((unsigned long long)yyh << 8); yyl = yyh = 0;
      t_xy += (unsigned long long)xyl + ((unsigned long long)xyh << 8); xyl = xyh = 0;
    }
  }
};

// Cube-wide packed accumulators of one thread.
struct CubeAcc {
  uint32_t maxsel;   // U8/U16: packed u16x2 max(x,y) over selected; I16: packed s16x2 max |v| (wrapping abs)
  uint32_t umax;     // packed u16x2 max of the reference over ALL samples (offset binary for I16)
  uint32_t umin;     // I16 only: packed u16x2 min (offset binary)
  uint32_t orbits;   // OR of all raw reference words
  __device__ __forceinline__ void reset() { maxsel = 0; umax = 0; umin = 0xffffffffu; orbits = 0; }
};

// One packed word: xr/yr = two raw 16-bit samples of the SAME band (two pixels).
//   m     halfword select mask (0xffff per selected half), only read when MASK
//   PAIR  false: both halves hold the same sample and it must count once (scalar head/tail)
template <int DT, bool MASK, bool MOMENTS, bool PAIR>
__device__ __forceinline__ uint32_t word_op(BandAcc<MOMENTS>& a, CubeAcc& c, uint32_t xr, uint32_t yr,
                                            uint32_t m) {
  c.orbits |= xr;
  uint32_t x = xr, y = yr;
  if (DT == DM_I16) { x ^= 0x80008000u; y ^= 0x80008000u; }
  c.umax = vmaxu2(c.umax, x);
  if (DT == DM_I16) c.umin = vminu2(c.umin, x);
  if (MASK) { x &= m; y &= m; }
  const uint32_t mx = vmaxu2(x, y);
  const uint32_t mn = vminu2(x, y);
  const uint32_t d = mx - mn;                 // packed |x-y|, each half 0..65535
  if (DT == DM_I16) {
    const uint32_t xs = MASK ? (xr & m) : xr, ys = MASK ? (yr & m) : yr;
    c.maxsel = __vimax3_s16x2(c.maxsel, __vabs2(xs), __vabs2(ys));   // np.abs(-32768) stays negative
  } else {
    c.maxsel = vmaxu2(c.maxsel, mx);
  }
  a.maxd = vmaxu2(a.maxd, d);
  const uint32_t ones = PAIR ? 0x0101u : 0x0001u;
  a.sabs = dp2a_lo(d, ones, a.sabs);
  if (MOMENTS) {
    uint32_t px = __byte_perm(x, 0, 0x3120);  // (x0.lo8, x1.lo8, x0.hi8, x1.hi8)
    uint32_t py = __byte_perm(y, 0, 0x3120);
    if (!PAIR) { px &= 0x00ff00ffu; py &= 0x00ff00ffu; }
    a.sx = dp2a_lo(x, ones, a.sx);
    a.sy = dp2a_lo(y, ones, a.sy);
    a.xxl = dp2a_lo(x, px, a.xxl);
    a.xxh = dp2a_hi(x, px, a.xxh);
    a.yyl = dp2a_lo(y, py, a.yyl);
    a.yyh = dp2a_hi(y, py, a.yyh);
    a.xyl = dp2a_lo(x, py, a.xyl);
    a.xyh = dp2a_hi(x, py, a.xyh);
  } else {
    uint32_t pd = __byte_perm(d, 0, 0x3120);
    if (!PAIR) pd &= 0x00ff00ffu;
    a.xxl = dp2a_lo(d, pd, a.xxl);
    a.xxh = dp2a_hi(d, pd, a.xxh);
  }
  return d;
}

// Exact signed sums from offset-binary ones (x' = x + 32768) over n samples.
__device__ __forceinline__ void unbias_i16(long long n, long long& sx, long long& sy, long long& sxx,
                                           long long& syy, long long& sxy) {
  const long long c = 32768, c2 = 32768ll * 32768ll;
  const long long xx = sxx - 2 * c * sx + c2 * n;
  const long long yy = syy - 2 * c * sy + c2 * n;
  const long long xy = sxy - c * (sx + sy) + c2 * n;
  sx -= c * n; sy -= c * n;
  sxx = xx; syy = yy; sxy = xy;
}

// Store one band's combined partial (already reduced) with 64-bit REDs.
template <int DT, bool MOMENTS>
__device__ __forceinline__ void red_band(const StatsArgs& g, int band, long long n, long long s_abs,
                                         long long sx, long long sy, long long sxx, long long syy,
                                         long long sxy, int maxd) {
  int64_t* S = g.sums + (int64_t)band * DM_NSTAT;
  if (n) atomic_add_i64(S + DM_S_N, n);
  if (s_abs) atomic_add_i64(S + DM_S_ABS, s_abs);
  if (MOMENTS) {
    if (DT == DM_I16) unbias_i16(n, sx, sy, sxx, syy, sxy);
    if (sx) atomic_add_i64(S + DM_S_X, sx);
    if (sy) atomic_add_i64(S + DM_S_Y, sy);
    if (sxx) atomic_add_i64(S + DM_S_XX, sxx);
    if (syy) atomic_add_i64(S + DM_S_YY, syy);
    if (sxy) atomic_add_i64(S + DM_S_XY, sxy);
    const long long sse = sxx + syy - 2 * sxy;      // == sum (x-y)^2 exactly
    if (sse) atomic_add_i64(S + DM_S_SSE, sse);
  } else {
    if (sxx) atomic_add_i64(S + DM_S_SSE, sxx);     // t_xx holds sum d^2 in this variant
  }
  if (maxd) atomic_max_i64(g.maxs + (int64_t)band * DM_NSTAT + DM_M_MAXERR, maxd);
}

template <int DT>
__device__ __forceinline__ void red_cube(const StatsArgs& g, int maxsel, int umax, int umin, unsigned orbits,
                                         bool any) {
  if (!any) return;
  int64_t* M = g.maxs;                               // cube-wide values are reported on band 0
  if (DT == DM_I16) {
    const int hi = umax - 32768, lo = umin - 32768;
    if (hi > 0) atomic_max_i64(M + DM_M_UMAX, hi);
    if (lo < 0) atomic_max_i64(M + DM_M_UNEGMIN, -lo);
  } else {
    if (umax > 0) atomic_max_i64(M + DM_M_UMAX, umax);
  }
  if (maxsel > 0) atomic_max_i64(M + DM_M_ABSXY, maxsel);
  if (orbits & 0xFu) atomic_max_i64(M + DM_M_LOW4, 1);
  if (orbits & 0x3u) atomic_max_i64(M + DM_M_LOW2, 1);
}

// ------------------------------------------------------------------------------------------------
// BSQ packed kernel
// ------------------------------------------------------------------------------------------------

constexpr int kThreadsBsq = 256;
constexpr int kUnrollBsq = 4;           // 16-byte vectors per cube in flight per thread

// 8 consecutive samples -> 4 packed words
template <int DT>
__device__ __forceinline__ void load_vec8(const void* base, int64_t elem, uint32_t (&w)[4]) {
  if (DT == DM_U8) {
    const uint2 v = ldg_stream8(static_cast<const uint8_t*>(base) + elem);
    w[0] = __byte_perm(v.x, 0, 0x4140); w[1] = __byte_perm(v.x, 0, 0x4342);
    w[2] = __byte_perm(v.y, 0, 0x4140); w[3] = __byte_perm(v.y, 0, 0x4342);
  } else {
    const uint4 v = ldg_stream16(static_cast<const uint16_t*>(base) + elem);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
  }
}

// 8 mask bytes (pixels elem..elem+7) shifted so that the selecting bit is bit 7 of every byte
__device__ __forceinline__ uint2 load_mask8(const uint8_t* p, int shift) {
  uint2 mv;
  if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) {
    mv = ldg_stream8(p);
  } else {
    uint32_t lo = 0, hi = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lo |= (uint32_t)p[j] << (8 * j);
      hi |= (uint32_t)p[4 + j] << (8 * j);
    }
    mv = make_uint2(lo, hi);
  }
  mv.x <<= shift; mv.y <<= shift;
  return mv;
}

template <int DT>
__device__ __forceinline__ uint32_t scalar_word(const void* base, int64_t i) {
  const uint32_t v = DT == DM_U8 ? (uint32_t)static_cast<const uint8_t*>(base)[i]
                                 : (uint32_t)static_cast<const uint16_t*>(base)[i];
  return v | (v << 16);
}

template <bool HIST_ON>
__device__ __forceinline__ void hist_word(unsigned* hist_sh, int K, int lane, uint32_t d, uint32_t m,
                                          bool masked, bool pair) {
  if (!HIST_ON) return;
  const uint32_t dk = vminu2(d, (uint32_t)(K - 1) * 0x10001u);
  if (!masked || (m & 0xffffu)) atomicAdd(&hist_sh[(dk & 0xffffu) * 32 + lane], 1u);
  if (pair && (!masked || (m >> 16))) atomicAdd(&hist_sh[(dk >> 16) * 32 + lane], 1u);
}

template <int DT, bool MASK, bool MOMENTS, bool HIST>
__global__ void __launch_bounds__(kThreadsBsq)
stats_bsq_packed(StatsArgs g, int64_t chunk_vecs, int64_t nchunk) {
  extern __shared__ unsigned hist_sh[];     // HIST: K*32 lane-private counters of the current band
  if (g.items) {                             // batch launch (dm_fused_stats_batch): this block row's pair
    const dm_batch_item_t it = g.items[blockIdx.y];
    g.ref = it.ref; g.tst = it.tst; g.sums = it.sums; g.maxs = it.maxs;
  }
  const int tid = threadIdx.x, lane = tid & 31;
  const int K = g.hist_bins;
  constexpr int EB = DT == DM_U8 ? 1 : 2;   // bytes per sample
  constexpr int VB = 8 * EB;                // bytes per 8-sample vector
  if (HIST) {
    for (int i = tid; i < K * 32; i += kThreadsBsq) hist_sh[i] = 0;
    __syncthreads();
  }
  const int64_t units = g.bands * nchunk;
  const int64_t u_begin = units * blockIdx.x / gridDim.x;
  const int64_t u_end = units * (blockIdx.x + 1) / gridDim.x;

  BandAcc<MOMENTS> a;
  CubeAcc c;
  a.reset();
  c.reset();
  int cur_band = -1;
  long long n = 0;        // selected samples of cur_band seen by this thread
  bool any = false;

  auto flush = [&]() {
    a.spill();
    long long v_n = warp_sum_ll(n), v_abs = warp_sum_ll((long long)a.t_abs);
    long long v_x = 0, v_y = 0, v_yy = 0, v_xy = 0;
    long long v_xx = warp_sum_ll((long long)a.t_xx);
    if (MOMENTS) {
      v_x = warp_sum_ll((long long)a.t_x); v_y = warp_sum_ll((long long)a.t_y);
      v_yy = warp_sum_ll((long long)a.t_yy); v_xy = warp_sum_ll((long long)a.t_xy);
    }
    const int v_maxd = (int)warp_max_ll(hmax2(a.maxd));
    if (lane == 0) red_band<DT, MOMENTS>(g, cur_band, v_n, v_abs, v_x, v_y, v_xx, v_yy, v_xy, v_maxd);
    a.reset();
    n = 0;
    if (HIST) {
      __syncthreads();
      int64_t* out = g.hist + (int64_t)cur_band * K;
      for (int k = tid; k < K; k += kThreadsBsq) {
        unsigned tot = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
          const int jj = (j + k) & 31;       // rotate so that the threads of a warp hit 32 banks
          tot += hist_sh[k * 32 + jj];
          hist_sh[k * 32 + jj] = 0;
        }
        if (tot) atomic_add_i64(out + k, (long long)tot);
      }
      __syncthreads();
    }
  };

  for (int64_t u = u_begin; u < u_end; ++u) {
    const int band = (int)(u / nchunk);
    const int64_t chunk = u - (int64_t)band * nchunk;
    if (band != cur_band) {
      if (cur_band >= 0) flush();
      cur_band = band;
    }
    const char* rb = static_cast<const char*>(g.ref) + (int64_t)band * g.band_stride * EB;
    const char* tb = static_cast<const char*>(g.tst) + (int64_t)band * g.band_stride * EB;
    // leading samples up to the first VB-aligned address (ref and tst are congruent mod VB)
    const int64_t head_al = (int64_t)(((VB - (int)(reinterpret_cast<uintptr_t>(rb) & (VB - 1))) & (VB - 1)) / EB);
    const int64_t head = head_al < g.npix ? head_al : g.npix;
    const int64_t nvec = (g.npix - head) / 8;
    const int64_t v0 = chunk * chunk_vecs;
    const int64_t v1 = min(v0 + chunk_vecs, nvec);
    const void* rv = rb + head * EB;
    const void* tv = tb + head * EB;
    const uint8_t* pv = MASK ? g.plane + head : nullptr;

    for (int64_t vb = v0 + tid; vb < v1; vb += (int64_t)kThreadsBsq * kUnrollBsq) {
      uint32_t xw[kUnrollBsq][4], yw[kUnrollBsq][4];
      uint2 mv[kUnrollBsq];
#pragma unroll
      for (int r = 0; r < kUnrollBsq; ++r) {
        const int64_t v = vb + (int64_t)r * kThreadsBsq;
        if (v < v1) {
          load_vec8<DT>(rv, v * 8, xw[r]);
          load_vec8<DT>(tv, v * 8, yw[r]);
          if (MASK) mv[r] = load_mask8(pv + v * 8, g.plane_shift);
        }
      }
#pragma unroll
      for (int r = 0; r < kUnrollBsq; ++r) {
        const int64_t v = vb + (int64_t)r * kThreadsBsq;
        if (v < v1) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            uint32_t m = 0;
            if (MASK) {
              const uint32_t mw = w < 2 ? mv[r].x : mv[r].y;
              m = prmt(mw, 0u, (w & 1) ? 0xbbaau : 0x9988u);   // sign-replicate two mask bytes
            }
            const uint32_t d = word_op<DT, MASK, MOMENTS, true>(a, c, xw[r][w], yw[r][w], m);
            hist_word<HIST>(hist_sh, K, lane, d, m, MASK, true);
          }
          if (MASK) n += __popc((mv[r].x & 0x80808080u)) + __popc((mv[r].y & 0x80808080u));
          else n += 8;
        }
      }
    }
    // scalar head and tail of the band, once per band (chunk 0)
    if (chunk == 0) {
      const int64_t ntail = g.npix - head - nvec * 8;
      const int64_t i = tid < head ? (int64_t)tid : (tid - head < ntail ? head + nvec * 8 + (tid - head) : -1);
      if (i >= 0 && tid < head + ntail) {
        const uint32_t xr = scalar_word<DT>(rb, i), yr = scalar_word<DT>(tb, i);
        uint32_t m = 0xffffffffu;
        if (MASK) m = (g.plane[i] & g.plane_bit) ? 0xffffffffu : 0u;
        const uint32_t d = word_op<DT, MASK, MOMENTS, false>(a, c, xr, yr, m);
        hist_word<HIST>(hist_sh, K, lane, d, m, MASK, false);
        n += (m != 0) ? 1 : 0;
      }
    }
    a.spill();          // chunk_vecs <= 128 words per accumulator between spills (host guarantees)
    any = true;
  }
  if (cur_band >= 0) flush();
  // cube-wide maxima, once per warp
  {
    int v_sel = DT == DM_I16 ? hmax2s(c.maxsel) : hmax2(c.maxsel);
    int v_umax = hmax2(c.umax), v_umin = hmin2(c.umin);
    unsigned v_or = (c.orbits | (c.orbits >> 16)) & 0xffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      v_sel = max(v_sel, __shfl_xor_sync(0xffffffffu, v_sel, o));
      v_umax = max(v_umax, __shfl_xor_sync(0xffffffffu, v_umax, o));
      v_umin = min(v_umin, __shfl_xor_sync(0xffffffffu, v_umin, o));
      v_or |= __shfl_xor_sync(0xffffffffu, v_or, o);
    }
    if (lane == 0) red_cube<DT>(g, v_sel, v_umax, v_umin, v_or, any);
  }
}

// ------------------------------------------------------------------------------------------------
// BIP packed kernel: thread <-> C consecutive 32-bit columns (2C bands) for the whole kernel;
// every step it loads the same columns of two pixels and pairs them band-wise with PRMT.
// ------------------------------------------------------------------------------------------------

struct BipGeom {
  int tpp;          // threads per pixel = (bands/2)/C
  int ppb;          // pixel rows per block step (block = tpp*ppb threads, a step covers 2*ppb pixels)
  int words;        // 32-bit words per pixel = bands/2
  int64_t ngroups;  // full steps of 2*ppb pixels
};

template <int C>
__device__ __forceinline__ void load_cols(const uint32_t* p, uint32_t (&w)[C]) {
  if (C == 2) { const uint2 v = ldg_stream8(p); w[0] = v.x; w[1] = v.y; }
  else { w[0] = ldg_stream4(p); }
}

template <int DT, bool MASK, bool MOMENTS, bool HIST, int C>
__global__ void __launch_bounds__(384)
stats_bip_packed(StatsArgs g, BipGeom geo) {
  constexpr int NB = 2 * C;      // bands per thread
  constexpr int U = 4;           // pixel-pair steps in flight
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int B = (int)g.bands, K = g.hist_bins;
  // [8 ints cube-wide | region shared by the histogram (main loop) and the combine partials (end)]
  int* sh_cube = reinterpret_cast<int*>(smem_raw);                                 // [8] (+ pad to 64 B)
  unsigned long long* sh_sums = reinterpret_cast<unsigned long long*>(smem_raw + 64);   // [8][NB][ppb*tpp] partials
  unsigned* sh_hist = reinterpret_cast<unsigned*>(smem_raw + 64);                  // HIST: [K][B]
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid < 8) sh_cube[tid] = tid == 2 ? 0x7fffffff : (tid == 0 ? (int)0x80000000 : 0);
  if (HIST) for (int i = tid; i < K * B; i += nt) sh_hist[i] = 0;
  __syncthreads();

  const int col = (tid % geo.tpp) * C;     // first 32-bit column of this thread
  const int prow = tid / geo.tpp;
  const uint32_t* ref = static_cast<const uint32_t*>(g.ref);
  const uint32_t* tst = static_cast<const uint32_t*>(g.tst);

  BandAcc<MOMENTS> a[NB];
  CubeAcc c;
#pragma unroll
  for (int j = 0; j < NB; ++j) a[j].reset();
  c.reset();
  long long n = 0;
  int since_spill = 0;
  bool any = false;

  // one step: pixels pa (low halves) and pb (high halves); pb < 0 => pa alone, counted once
  auto step = [&](const uint32_t (&xa)[C], const uint32_t (&xb)[C], const uint32_t (&ya)[C],
                  const uint32_t (&yb)[C], uint32_t m, auto pair_tag) {
    constexpr bool PAIR = decltype(pair_tag)::value;
#pragma unroll
    for (int k = 0; k < C; ++k) {
      const uint32_t x0 = __byte_perm(xa[k], xb[k], 0x5410), x1 = __byte_perm(xa[k], xb[k], 0x7632);
      const uint32_t y0 = __byte_perm(ya[k], yb[k], 0x5410), y1 = __byte_perm(ya[k], yb[k], 0x7632);
      const uint32_t d0 = word_op<DT, MASK, MOMENTS, PAIR>(a[2 * k], c, x0, y0, m);
      const uint32_t d1 = word_op<DT, MASK, MOMENTS, PAIR>(a[2 * k + 1], c, x1, y1, m);
      if (HIST) {
        const uint32_t kk = (uint32_t)(K - 1) * 0x10001u;
        const uint32_t e0 = vminu2(d0, kk), e1 = vminu2(d1, kk);
        unsigned* h0 = sh_hist + 2 * (col + k);
        if (!MASK || (m & 0xffffu)) { atomicAdd(h0 + (e0 & 0xffffu) * B, 1u); atomicAdd(h0 + 1 + (e1 & 0xffffu) * B, 1u); }
        if (PAIR && (!MASK || (m >> 16))) { atomicAdd(h0 + (e0 >> 16) * B, 1u); atomicAdd(h0 + 1 + (e1 >> 16) * B, 1u); }
      }
    }
  };
  using TrueT = std::true_type;
  using FalseT = std::false_type;

  if (prow < geo.ppb) {
    // running pointers: pixel A of this thread's first step; pixel B sits offb words further on, the
    // next step of this block gstep words further on (64-bit adds instead of index arithmetic)
    const int64_t offb = (int64_t)geo.ppb * geo.words;
    const int64_t gstep = (int64_t)gridDim.x * 2 * offb;
    const int64_t first = ((int64_t)blockIdx.x * 2 * geo.ppb + prow) * geo.words + col;
    const uint32_t* pr = ref + first;
    const uint32_t* pt = tst + first;
    const uint8_t* pm = MASK ? g.plane + (int64_t)blockIdx.x * 2 * geo.ppb + prow : nullptr;
    const int64_t mstep = (int64_t)gridDim.x * 2 * geo.ppb;
    int64_t left = geo.ngroups > blockIdx.x ? (geo.ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // steps of this block
    any = left > 0;
    auto run = [&](auto ucount) {
      constexpr int UU = decltype(ucount)::value;
      uint32_t xa[UU][C], xb[UU][C], ya[UU][C], yb[UU][C];
      uint32_t m[UU];
#pragma unroll
      for (int r = 0; r < UU; ++r) {
        load_cols<C>(pr + r * gstep, xa[r]);
        load_cols<C>(pr + r * gstep + offb, xb[r]);
        load_cols<C>(pt + r * gstep, ya[r]);
        load_cols<C>(pt + r * gstep + offb, yb[r]);
        if (MASK) {
          const uint32_t sa = (pm[r * mstep] & g.plane_bit) ? 0xffffu : 0u;
          const uint32_t sb = (pm[r * mstep + geo.ppb] & g.plane_bit) ? 0xffff0000u : 0u;
          m[r] = sa | sb;
        }
      }
#pragma unroll
      for (int r = 0; r < UU; ++r) {
        step(xa[r], xb[r], ya[r], yb[r], MASK ? m[r] : 0xffffffffu, TrueT());
        if (MASK) n += ((m[r] & 1u) ? 1 : 0) + ((m[r] >> 31) ? 1 : 0);
        else n += 2;
      }
      pr += UU * gstep; pt += UU * gstep;
      if (MASK) pm += UU * mstep;
      left -= UU;
      since_spill += UU;
      if (since_spill >= 128 - U) {
        since_spill = 0;
#pragma unroll
        for (int j = 0; j < NB; ++j) a[j].spill();
      }
    };
    while (left >= U) run(std::integral_constant<int, U>());
    while (left > 0) run(std::integral_constant<int, 1>());
    // leftover pixels (< 2*ppb), one at a time, by block 0
    if (blockIdx.x == 0) {
      const int64_t p0 = geo.ngroups * 2 * geo.ppb;
      for (int64_t p = p0 + prow; p < g.npix; p += geo.ppb) {
        uint32_t xa[C], ya[C];
        load_cols<C>(ref + p * geo.words + col, xa);
        load_cols<C>(tst + p * geo.words + col, ya);
        uint32_t m = 0xffffffffu;
        if (MASK) m = (g.plane[p] & g.plane_bit) ? 0xffffffffu : 0u;
        step(xa, xa, ya, ya, m, FalseT());
        n += m ? 1 : 0;
        any = true;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) a[j].spill();

  if (HIST) {       // flush the histogram first: its memory is reused by the combine below
    __syncthreads();
    for (int i = tid; i < K * B; i += nt) {
      const int k = i / B, b = i - k * B;
      const unsigned cnt = sh_hist[i];
      if (cnt) atomic_add_i64(g.hist + (int64_t)b * K + k, (long long)cnt);
    }
    __syncthreads();
  }
  // combine (once per kernel), without atomics: every thread stores its per-band values at
  // [quantity][j][pixel row][column group] (consecutive lanes -> consecutive 8-byte words), then one
  // thread per band adds the pixel rows
  {
    unsigned long long* sh_part = sh_sums;             // [8][NB][ppb*tpp] uint64 (sized by the host)
    const int lanes_used = geo.ppb * geo.tpp;
    const size_t q = (size_t)NB * lanes_used;
    if (prow < geo.ppb) {
      const int me = prow * geo.tpp + col / C;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        unsigned long long* d = sh_part + (size_t)j * lanes_used + me;
        d[0 * q] = (unsigned long long)n;
        d[1 * q] = a[j].t_x; d[2 * q] = a[j].t_y; d[3 * q] = a[j].t_xx; d[4 * q] = a[j].t_yy; d[5 * q] = a[j].t_xy;
        d[6 * q] = a[j].t_abs; d[7 * q] = (unsigned long long)hmax2(a[j].maxd);
      }
      if (any) {
        atomicMax(sh_cube + 0, DT == DM_I16 ? hmax2s(c.maxsel) : hmax2(c.maxsel));
        atomicMax(sh_cube + 1, hmax2(c.umax));
        atomicMin(sh_cube + 2, hmin2(c.umin));
        atomicOr(reinterpret_cast<unsigned*>(sh_cube + 3), (c.orbits | (c.orbits >> 16)) & 0xffffu);
        sh_cube[4] = 1;
      }
    }
    __syncthreads();
    for (int b = tid; b < B; b += nt) {
      const int cg = b / NB, j = b - cg * NB;
      unsigned long long v[7] = {0, 0, 0, 0, 0, 0, 0};
      int md = 0;
      for (int r = 0; r < geo.ppb; ++r) {
        const unsigned long long* d = sh_part + (size_t)j * lanes_used + r * geo.tpp + cg;
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] += d[k * q];
        md = max(md, (int)d[7 * q]);
      }
      red_band<DT, MOMENTS>(g, b, (long long)v[0], (long long)v[6], (long long)v[1], (long long)v[2], (long long)v[3],
                            (long long)v[4], (long long)v[5], md);
    }
  }
  if (tid == 0) red_cube<DT>(g, sh_cube[0], sh_cube[1], sh_cube[2], (unsigned)sh_cube[3], sh_cube[4] != 0);
}

// ------------------------------------------------------------------------------------------------
// generic scalar kernel: one block per (band, chunk); any dtype, layout, stride and alignment
// ------------------------------------------------------------------------------------------------

template <int DT, typename T>
__global__ void __launch_bounds__(256)
stats_generic(StatsArgs g, int64_t elem_stride, int64_t band_step, int64_t chunk, int64_t nchunk,
              int moments) {
  const int band = (int)(blockIdx.x / nchunk);
  const int64_t cidx = blockIdx.x - (int64_t)band * nchunk;
  const T* rb = static_cast<const T*>(g.ref) + (int64_t)band * band_step;
  const T* tb = static_cast<const T*>(g.tst) + (int64_t)band * band_step;
  const int64_t i0 = cidx * chunk, i1 = min(i0 + chunk, g.npix);
  const bool masked = g.plane != nullptr;
  const int K = g.hist_bins;
  long long n = 0, sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0, sabs = 0, sse = 0;
  int maxd = 0, maxsel = 0, umax = 0, unegmin = 0;
  unsigned orbits = 0;
  for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    const int x = (int)rb[i * elem_stride];
    const int y = (int)tb[i * elem_stride];
    umax = max(umax, x);
    unegmin = max(unegmin, -x);
    orbits |= (unsigned)x;
    if (masked && !(g.plane[i] & g.plane_bit)) continue;
    const int d = abs(x - y);
    n += 1; sx += x; sy += y;
    sxx += (long long)x * x; syy += (long long)y * y; sxy += (long long)x * y;
    sabs += d; sse += (long long)d * d;
    maxd = max(maxd, d);
    // np.abs on int16 wraps -32768 to itself, which never wins the max (run_codec.py:285)
    maxsel = max(maxsel, max(x == -32768 ? 0 : abs(x), y == -32768 ? 0 : abs(y)));
    if (K) atomic_add_i64(g.hist + (int64_t)band * K + min(d, K - 1), 1);
  }
  n = warp_sum_ll(n); sabs = warp_sum_ll(sabs); sse = warp_sum_ll(sse);
  sx = warp_sum_ll(sx); sy = warp_sum_ll(sy); sxx = warp_sum_ll(sxx); syy = warp_sum_ll(syy); sxy = warp_sum_ll(sxy);
  maxd = (int)warp_max_ll(maxd); maxsel = (int)warp_max_ll(maxsel);
  umax = (int)warp_max_ll(umax); unegmin = (int)warp_max_ll(unegmin);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) orbits |= __shfl_xor_sync(0xffffffffu, orbits, o);
  if ((threadIdx.x & 31) == 0) {
    int64_t* S = g.sums + (int64_t)band * DM_NSTAT;
    int64_t* M = g.maxs + (int64_t)band * DM_NSTAT;
    if (n) atomic_add_i64(S + DM_S_N, n);
    if (sabs) atomic_add_i64(S + DM_S_ABS, sabs);
    if (sse) atomic_add_i64(S + DM_S_SSE, sse);
    if (moments) {
      if (sx) atomic_add_i64(S + DM_S_X, sx);
      if (sy) atomic_add_i64(S + DM_S_Y, sy);
      if (sxx) atomic_add_i64(S + DM_S_XX, sxx);
      if (syy) atomic_add_i64(S + DM_S_YY, syy);
      if (sxy) atomic_add_i64(S + DM_S_XY, sxy);
    }
    if (maxd) atomic_max_i64(M + DM_M_MAXERR, maxd);
    if (maxsel) atomic_max_i64(M + DM_M_ABSXY, maxsel);
    if (umax) atomic_max_i64(M + DM_M_UMAX, umax);
    if (unegmin) atomic_max_i64(M + DM_M_UNEGMIN, unegmin);
    if (orbits & 0xFu) atomic_max_i64(M + DM_M_LOW4, 1);
    if (orbits & 0x3u) atomic_max_i64(M + DM_M_LOW2, 1);
  }
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------

template <typename KernelT>
int blocks_per_sm(KernelT kernel, int threads, size_t smem) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem) != cudaSuccess) nb = 1;
  return nb < 1 ? 1 : nb;
}

template <int DT, bool MASK, bool MOMENTS, bool HIST>
int run_bsq(const StatsArgs& g, cudaStream_t s, int n_items = 1) {
  auto kernel = stats_bsq_packed<DT, MASK, MOMENTS, HIST>;
  const size_t smem = HIST ? (size_t)g.hist_bins * 32 * sizeof(unsigned) : 0;
  if (smem > 48 * 1024)
    DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const int64_t max_blocks = (int64_t)sms * blocks_per_sm(kernel, kThreadsBsq, smem);
  const int64_t nvec = g.npix / 8;
  // a chunk is 1..8 trips of the unrolled inner loop (<= 128 packed words per accumulator between
  // spills); small cubes get small chunks so that every SM has work
  const int64_t trip = (int64_t)kThreadsBsq * kUnrollBsq;
  int64_t trips = (g.bands * nvec * n_items + max_blocks * trip - 1) / (max_blocks * trip);
  trips = trips < 1 ? 1 : (trips > 8 ? 8 : trips);
  const int64_t chunk_vecs = trips * trip;
  int64_t nchunk = (nvec + chunk_vecs - 1) / chunk_vecs;
  if (nchunk < 1) nchunk = 1;
  const int64_t units = g.bands * nchunk;
  int64_t per_item = max_blocks / n_items;              // a batch shares the resident blocks between its pairs
  if (per_item < 1) per_item = 1;
  const int64_t grid = units < per_item ? units : per_item;
  kernel<<<dim3((unsigned)grid, (unsigned)n_items), kThreadsBsq, smem, s>>>(g, chunk_vecs, nchunk);
  DM_LAUNCH_CHECK("stats_bsq_packed");
  return DM_OK;
}

template <int DT, bool MASK, bool MOMENTS, bool HIST, int C>
int run_bip(const StatsArgs& g, cudaStream_t s) {
  const int B = (int)g.bands;
  BipGeom geo;
  geo.words = B / 2;
  geo.tpp = geo.words / C;
  int ppb = 384 / geo.tpp;
  if (ppb < 1) ppb = 1;
  geo.ppb = ppb;
  const int threads = geo.tpp * ppb;
  geo.ngroups = g.npix / (2 * ppb);
  auto kernel = stats_bip_packed<DT, MASK, MOMENTS, HIST, C>;
  size_t smem = (size_t)8 * (2 * C) * threads * 8;
  if (HIST && (size_t)g.hist_bins * B * sizeof(unsigned) > smem) smem = (size_t)g.hist_bins * B * sizeof(unsigned);
  smem += 64;
  if (smem > 48 * 1024)
    DM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  int64_t grid = (int64_t)sms * blocks_per_sm(kernel, threads, smem);
  if (grid > geo.ngroups) grid = geo.ngroups;
  if (grid < 1) grid = 1;
  kernel<<<(unsigned)grid, threads, smem, s>>>(g, geo);
  DM_LAUNCH_CHECK("stats_bip_packed");
  return DM_OK;
}

template <int DT, bool MASK, bool MOMENTS, bool HIST>
int dispatch_layout(const StatsArgs& g, int bip_cols, cudaStream_t s) {
  if constexpr (DT != DM_U8) {
    if (bip_cols == 2) return run_bip<DT, MASK, MOMENTS, HIST, 2>(g, s);
    if (bip_cols == 1) return run_bip<DT, MASK, MOMENTS, HIST, 1>(g, s);
  }
  return run_bsq<DT, MASK, MOMENTS, HIST>(g, s);
}

template <int DT, bool MASK>
int dispatch_variant(const StatsArgs& g, int bip_cols, bool moments, cudaStream_t s) {
  const bool hist = g.hist_bins > 0;
  if (moments) return hist ? dispatch_layout<DT, MASK, true, true>(g, bip_cols, s)
                           : dispatch_layout<DT, MASK, true, false>(g, bip_cols, s);
  return hist ? dispatch_layout<DT, MASK, false, true>(g, bip_cols, s)
              : dispatch_layout<DT, MASK, false, false>(g, bip_cols, s);
}

template <int DT>
int dispatch_mask(const StatsArgs& g, int bip_cols, bool moments, cudaStream_t s) {
  return g.plane ? dispatch_variant<DT, true>(g, bip_cols, moments, s)
                 : dispatch_variant<DT, false>(g, bip_cols, moments, s);
}

template <int DT, typename T>
int run_generic(const StatsArgs& g, bool bip, bool moments, cudaStream_t s) {
  const int64_t chunk = 1 << 15;
  int64_t nchunk = (g.npix + chunk - 1) / chunk;
  if (nchunk < 1) nchunk = 1;
  const int64_t grid = g.bands * nchunk;
  if (grid > 0x7fffffffll) return fail(DM_EUNSUPPORTED, "dm_fused_stats: cube too large for the generic path");
  const int64_t elem_stride = bip ? g.bands : 1;
  const int64_t band_step = bip ? 1 : g.band_stride;
  stats_generic<DT, T><<<(unsigned)grid, 256, 0, s>>>(g, elem_stride, band_step, chunk, nchunk, moments ? 1 : 0);
  DM_LAUNCH_CHECK("stats_generic");
  return DM_OK;
}

}  // namespace

int launch_fused_stats(const dm_pair_t& p, const uint8_t* plane, int plane_bit, int hist_bins,
                       uint32_t flags, int64_t* sums, int64_t* maxs, int64_t* hist, cudaStream_t s) {
  if (!p.ref || !p.tst || !sums || !maxs) return fail(DM_EARG, "dm_fused_stats: null pointer");
  if (p.bands <= 0 || p.rows < 0 || p.width < 0) return fail(DM_EARG, "dm_fused_stats: bad geometry");
  if (hist_bins < 0 || hist_bins > 1024 || (hist_bins & (hist_bins - 1)))
    return fail(DM_EARG, "dm_fused_stats: hist_bins must be 0 or a power of two <= 1024");
  if (hist_bins && !hist) return fail(DM_EARG, "dm_fused_stats: hist is null");
  if (plane && !(plane_bit > 0 && plane_bit < 256 && (plane_bit & (plane_bit - 1)) == 0))
    return fail(DM_EARG, "dm_fused_stats: plane_bit must be a single bit 1..128");
  if (p.dtype != DM_U8 && p.dtype != DM_U16 && p.dtype != DM_I16) return fail(DM_EARG, "dm_fused_stats: bad dtype");
  const bool bip = p.layout == DM_BIP;
  if (!bip && p.layout != DM_BSQ) return fail(DM_EARG, "dm_fused_stats: bad layout");
  StatsArgs g;
  g.ref = p.ref; g.tst = p.tst; g.plane = plane; g.plane_bit = plane_bit; g.plane_shift = 0;
  if (plane) { int b = plane_bit; while (b < 128) { b <<= 1; ++g.plane_shift; } }
  g.bands = p.bands; g.npix = p.rows * p.width;
  g.band_stride = bip ? 1 : p.band_stride;
  g.hist_bins = hist_bins; g.sums = sums; g.maxs = maxs; g.hist = hist; g.items = nullptr;
  if (g.npix == 0) return DM_OK;
  if (!bip && p.band_stride < g.npix) return fail(DM_EARG, "dm_fused_stats: band_stride < rows*width");
  const bool moments = !(flags & DM_STATS_NO_MOMENTS);
  const uintptr_t ra = reinterpret_cast<uintptr_t>(p.ref), ta = reinterpret_cast<uintptr_t>(p.tst);
  const int eb = elem_bytes(p.dtype);

  bool packed = !(flags & DM_STATS_GENERIC);
  int bip_cols = 0;
  if (packed && !bip) {
    // ref and tst must be congruent modulo the vector size so that one head length serves both
    const uintptr_t vb = 8 * eb;
    packed = (ra % eb == 0) && (ta % eb == 0) && ((ra % vb) == (ta % vb)) &&
             (!hist_bins || (size_t)hist_bins * 32 * 4 <= 200 * 1024);
  } else if (packed) {
    const int B = (int)p.bands;
    packed = p.dtype != DM_U8 && (B % 2 == 0) && (ra % 4 == 0) && (ta % 4 == 0) && B <= 2048;
    if (packed) {
      bip_cols = (B % 4 == 0 && ra % 8 == 0 && ta % 8 == 0) ? 2 : 1;
      if ((B / 2) / bip_cols > 384) packed = false;
      if ((size_t)hist_bins * B * 4 > 200 * 1024) packed = false;
    }
  }
  if (!packed) {
    switch (p.dtype) {
      case DM_U8: return run_generic<DM_U8, uint8_t>(g, bip, moments, s);
      case DM_U16: return run_generic<DM_U16, uint16_t>(g, bip, moments, s);
      default: return run_generic<DM_I16, int16_t>(g, bip, moments, s);
    }
  }
  switch (p.dtype) {
    case DM_U8: return dispatch_mask<DM_U8>(g, 0, moments, s);
    case DM_U16: return dispatch_mask<DM_U16>(g, bip_cols, moments, s);
    default: return dispatch_mask<DM_I16>(g, bip_cols, moments, s);
  }
}

int launch_fused_stats_batch(const dm_pair_t& p, const dm_batch_item_t* items_dev, int n_items, uint32_t flags,
                             cudaStream_t s) {
  if (!items_dev || n_items < 0) return fail(DM_EARG, "dm_fused_stats_batch: null items / negative count");
  if (n_items > 65535) return fail(DM_EARG, "dm_fused_stats_batch: at most 65535 pairs per launch");
  if (p.layout != DM_BSQ) return fail(DM_EUNSUPPORTED, "dm_fused_stats_batch: DM_BSQ cubes only");
  if (p.dtype != DM_U8 && p.dtype != DM_U16 && p.dtype != DM_I16) return fail(DM_EARG, "dm_fused_stats_batch: bad dtype");
  if (p.bands <= 0 || p.rows < 0 || p.width < 0) return fail(DM_EARG, "dm_fused_stats_batch: bad geometry");
  if (p.band_stride < p.rows * p.width) return fail(DM_EARG, "dm_fused_stats_batch: band_stride < rows*width");
  if ((p.band_stride * elem_bytes(p.dtype)) % 16) return fail(DM_EUNSUPPORTED, "dm_fused_stats_batch: bands must start on 16-byte boundaries");
  if (flags & DM_STATS_GENERIC) return fail(DM_EUNSUPPORTED, "dm_fused_stats_batch: packed kernel only");
  StatsArgs g;
  g.ref = nullptr; g.tst = nullptr; g.plane = nullptr; g.plane_bit = 0; g.plane_shift = 0;
  g.bands = p.bands; g.npix = p.rows * p.width; g.band_stride = p.band_stride;
  g.hist_bins = 0; g.sums = nullptr; g.maxs = nullptr; g.hist = nullptr; g.items = items_dev;
  if (g.npix == 0 || n_items == 0) return DM_OK;
  const bool moments = !(flags & DM_STATS_NO_MOMENTS);
#define DM_BATCH(DT) (moments ? run_bsq<DT, false, true, false>(g, s, n_items) : run_bsq<DT, false, false, false>(g, s, n_items))
  switch (p.dtype) {
    case DM_U8: return DM_BATCH(DM_U8);
    case DM_U16: return DM_BATCH(DM_U16);
    default: return DM_BATCH(DM_I16);
  }
#undef DM_BATCH
}

}  // namespace dm
