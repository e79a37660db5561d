// One-pass BSQ kernel for few-band cubes (Sentinel-2 Case A: 4 bands): per-band integer statistics AND
// the per-pixel error quicklook from a single read of both cubes.
//
//   per band   {N, Sx, Sy, Sxx, Syy, Sxy, S|d|, SSE, max|d|}, data-range scan   run_codec.py:268-285, 86-117
//   per pixel  max_b |d| -> ERR8 planes + their 256-bin histograms             quicklooks.py:123-150, 175-184
//
// In BSQ a 16-byte vector is eight neighbouring pixels of ONE band, so a thread that loads the same
// vector of every band (2 x NB streaming 128-bit loads in flight) holds whole spectra of eight
// pixels: the packed |x-y| words feed the per-band dp2a accumulators (same arithmetic as stats.cu)
// and their element-wise maximum over the bands is the error map.  The ERR8 scaling is the host's
// LUT (bit-exact by construction); both planes share ONE histogram of min(e, 255) (lane-private
// shared-memory counters), which the block maps through the two LUTs at the end.
//
// 32-bit partials hold 128 dp2a steps; they spill into per-thread 64-bit slots in shared memory
// (no contention), reduced once per block.  HBM bound: 4 B per sample pair + 1 B per pixel and plane.

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBands = 4;
constexpr int kNQ = 6;                    // abs, x, y, xx, yy, xy

__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) { uint32_t r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t vminu2(uint32_t a, uint32_t b) { uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel)); return r; }
__device__ __forceinline__ int hmax2(uint32_t p) { return max((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmin2(uint32_t p) { return min((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmax2s(uint32_t p) { return max((int)(short)(p & 0xffffu), (int)(short)(p >> 16)); }

struct Args {
  const void* ref;
  const void* tst;
  const uint8_t* plane;       // may be null
  int64_t npix, band_stride;
  int64_t* sums;
  int64_t* maxs;
  uint16_t* errmax;
  const uint8_t* lut_g; int cap_g; uint8_t* err8_g; int64_t* hist8_g;
  const uint8_t* lut_z; int cap_z; uint8_t* err8_z; int64_t* hist8_z;
};

struct Acc {
  uint32_t sabs, sx, sy, xxl, xxh, yyl, yyh, xyl, xyh, maxd;
  __device__ __forceinline__ void zero_sums() { sabs = sx = sy = xxl = xxh = yyl = yyh = xyl = xyh = 0; }
};
struct Cube {
  uint32_t maxsel, umax, umin, orbits;
};

// one packed word of one band (two pixels, or one pixel twice when !PAIR); returns the UNMASKED |x-y|
template <int DT, bool MASK, bool PAIR>
__device__ __forceinline__ uint32_t word(Acc& a, Cube& c, uint32_t xr, uint32_t yr, uint32_t m) {
  c.orbits |= xr;                                   // data-range scan: raw reference words, unmasked
  uint32_t x = xr, y = yr;
  if (DT == DM_I16) { x ^= 0x80008000u; y ^= 0x80008000u; }
  c.umax = vmaxu2(c.umax, x);
  if (DT == DM_I16) c.umin = vminu2(c.umin, x);
  uint32_t draw = 0;
  if (MASK) {
    draw = vmaxu2(x, y) - vminu2(x, y);
    x &= m; y &= m;
  }
  const uint32_t mx = vmaxu2(x, y), mn = vminu2(x, y), d = mx - mn;
  if (!MASK) draw = d;
  if (DT == DM_I16) {
    const uint32_t xs = MASK ? (xr & m) : xr, ys = MASK ? (yr & m) : yr;
    c.maxsel = __vimax3_s16x2(c.maxsel, __vabs2(xs), __vabs2(ys));   // np.abs(-32768) stays negative
  } else {
    c.maxsel = vmaxu2(c.maxsel, mx);
  }
  a.maxd = vmaxu2(a.maxd, d);
  const uint32_t ones = PAIR ? 0x0101u : 0x0001u;
  uint32_t px = __byte_perm(x, 0, 0x3120), py = __byte_perm(y, 0, 0x3120);
  if (!PAIR) { px &= 0x00ff00ffu; py &= 0x00ff00ffu; }
  a.sabs = dp2a_lo(d, ones, a.sabs);
  a.sx = dp2a_lo(x, ones, a.sx);
  a.sy = dp2a_lo(y, ones, a.sy);
  a.xxl = dp2a_lo(x, px, a.xxl); a.xxh = dp2a_hi(x, px, a.xxh);
  a.yyl = dp2a_lo(y, py, a.yyl); a.yyh = dp2a_hi(y, py, a.yyh);
  a.xyl = dp2a_lo(x, py, a.xyl); a.xyh = dp2a_hi(x, py, a.xyh);
  return draw;
}

// halfword select masks of a packed word from two plane bytes (bit `bit` of each)
__device__ __forceinline__ uint32_t mask_word(uint32_t b0, uint32_t b1, uint32_t bit) {
  return ((b0 & bit) ? 0xffffu : 0u) | ((b1 & bit) ? 0xffff0000u : 0u);
}

template <int DT, bool MASK, bool ERR, int NB>
__global__ void __launch_bounds__(kThreads, 2)
fused_bsq_kernel(Args g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [NB*kNQ][kThreads] 64-bit per-thread totals | [256][32] lane-private e-histogram | two LUTs
  unsigned long long* tot = reinterpret_cast<unsigned long long*>(smem_raw);
  unsigned* hist = reinterpret_cast<unsigned*>(tot + NB * kNQ * kThreads);
  uint8_t* lut_g = reinterpret_cast<uint8_t*>(hist + (ERR ? 256 * 32 : 0));
  uint8_t* lut_z = lut_g + 256;
  __shared__ long long red_n[kThreads / 32];
  __shared__ int red_m[NB + 4][kThreads / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NB * kNQ * kThreads; i += kThreads) tot[i] = 0ull;
  if (ERR) {
    for (int i = tid; i < 256 * 32; i += kThreads) hist[i] = 0u;
    // LUT index min(e, cap) with cap <= 255: replicate the last entry up to 255
    lut_g[tid] = g.err8_g ? g.lut_g[min(tid, g.cap_g)] : (uint8_t)0;
    lut_z[tid] = g.err8_z ? g.lut_z[min(tid, g.cap_z)] : (uint8_t)0;
  }
  __syncthreads();

  Acc a[NB];
  Cube c;
#pragma unroll
  for (int b = 0; b < NB; ++b) { a[b].zero_sums(); a[b].maxd = 0; }
  c.maxsel = 0; c.umax = 0; c.umin = 0xffffffffu; c.orbits = 0;
  long long n = 0;
  int since_spill = 0;
  bool any = false;

  auto spill = [&]() {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      unsigned long long* t = tot + (size_t)(b * kNQ) * kThreads + tid;
      t[0 * kThreads] += a[b].sabs;
      t[1 * kThreads] += a[b].sx;
      t[2 * kThreads] += a[b].sy;
      t[3 * kThreads] += (unsigned long long)a[b].xxl + ((unsigned long long)a[b].xxh << 8);
      t[4 * kThreads] += (unsigned long long)a[b].yyl + ((unsigned long long)a[b].yyh << 8);
      t[5 * kThreads] += (unsigned long long)a[b].xyl + ((unsigned long long)a[b].xyh << 8);
      a[b].zero_sums();
    }
    since_spill = 0;
  };
  // ERR8 bytes + histogram for the two pixels of one packed error word; returns (g byte0, g byte1, z byte0, z byte1)
  auto err_pair = [&](uint32_t e2, uint32_t& outg, uint32_t& outz, int shift) {
    const uint32_t k2 = vminu2(e2, 0x00ff00ffu);            // min(e, 255) per pixel
    const uint32_t k0 = k2 & 0xffffu, k1 = k2 >> 16;
    outg |= ((uint32_t)lut_g[k0] << shift) | ((uint32_t)lut_g[k1] << (shift + 8));
    outz |= ((uint32_t)lut_z[k0] << shift) | ((uint32_t)lut_z[k1] << (shift + 8));
    atomicAdd(&hist[k0 * 32 + lane], 1u);
    atomicAdd(&hist[k1 * 32 + lane], 1u);
  };

  const int64_t nvec = g.npix / 8;
  const char* rbase = static_cast<const char*>(g.ref);
  const char* tbase = static_cast<const char*>(g.tst);
  const int64_t bstride = g.band_stride * 2;                // bytes
  for (int64_t v = (int64_t)blockIdx.x * kThreads + tid; v < nvec; v += (int64_t)gridDim.x * kThreads) {
    uint4 xv[NB], yv[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      xv[b] = ldg_stream16(rbase + b * bstride + v * 16);
      yv[b] = ldg_stream16(tbase + b * bstride + v * 16);
    }
    uint32_t mm[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, mq[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
    if (MASK) {
      const uint2 pv = ldg_stream8(g.plane + v * 8);
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t src = w < 2 ? pv.x : pv.y;
        const uint32_t b0 = (src >> (16 * (w & 1))) & 0xffu, b1 = (src >> (16 * (w & 1) + 8)) & 0xffu;
        mm[w] = mask_word(b0, b1, DM_VALID_METRICS);
        mq[w] = mask_word(b0, b1, DM_VALID_QUICKLOOK);
        n += (mm[w] & 1u) + (mm[w] >> 31);
      }
    } else {
      n += 8;
    }
    if (since_spill + 4 > 128) spill();
    since_spill += 4;
    any = true;
    uint32_t e[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint32_t xw[4] = {xv[b].x, xv[b].y, xv[b].z, xv[b].w}, yw[4] = {yv[b].x, yv[b].y, yv[b].z, yv[b].w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t d = word<DT, MASK, true>(a[b], c, xw[w], yw[w], mm[w]);
        if (ERR) e[w] = vmaxu2(e[w], d);
      }
    }
    if (ERR) {
      if (MASK) {
#pragma unroll
        for (int w = 0; w < 4; ++w) e[w] &= mq[w];                        // quicklooks.py:134
      }
      if (g.errmax) *reinterpret_cast<uint4*>(g.errmax + v * 8) = make_uint4(e[0], e[1], e[2], e[3]);
      uint32_t g0 = 0, g1 = 0, z0 = 0, z1 = 0;
      err_pair(e[0], g0, z0, 0); err_pair(e[1], g0, z0, 16);
      err_pair(e[2], g1, z1, 0); err_pair(e[3], g1, z1, 16);
      if (g.err8_g) *reinterpret_cast<uint2*>(g.err8_g + v * 8) = make_uint2(g0, g1);
      if (g.err8_z) *reinterpret_cast<uint2*>(g.err8_z + v * 8) = make_uint2(z0, z1);
    }
  }
  // the npix % 8 trailing pixels: one thread each, in block 0
  const int64_t tail0 = nvec * 8;
  if (blockIdx.x == 0 && tail0 + tid < g.npix) {
    const int64_t p = tail0 + tid;
    uint32_t m = 0xffffffffu, q = 0xffffffffu;
    if (MASK) {
      const uint32_t pb = g.plane[p];
      m = (pb & DM_VALID_METRICS) ? 0xffffffffu : 0u;
      q = (pb & DM_VALID_QUICKLOOK) ? 0xffffffffu : 0u;
    }
    n += m ? 1 : 0;
    spill();
    since_spill = 1;
    any = true;
    uint32_t e = 0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint32_t xs = static_cast<const uint16_t*>(g.ref)[b * g.band_stride + p];
      const uint32_t ys = static_cast<const uint16_t*>(g.tst)[b * g.band_stride + p];
      const uint32_t d = word<DT, MASK, false>(a[b], c, xs | (xs << 16), ys | (ys << 16), m);
      if (ERR) e = max(e, d & 0xffffu);
    }
    if (ERR) {
      e &= q & 0xffffu;
      if (g.errmax) g.errmax[p] = (uint16_t)e;
      const uint32_t k = min(e, 255u);
      if (g.err8_g) g.err8_g[p] = lut_g[k];
      if (g.err8_z) g.err8_z[p] = lut_z[k];
      atomicAdd(&hist[k * 32 + lane], 1u);
    }
  }
  spill();
  __syncthreads();

  // ---- block reduction: 64-bit totals per (band, quantity), counts, maxima
  n = warp_sum_ll(n);
  if (lane == 0) red_n[warp] = n;
  {
    int v[NB + 4];
#pragma unroll
    for (int b = 0; b < NB; ++b) v[b] = hmax2(a[b].maxd);
    v[NB + 0] = DT == DM_I16 ? hmax2s(c.maxsel) : hmax2(c.maxsel);
    v[NB + 1] = hmax2(c.umax);
    v[NB + 2] = -hmin2(c.umin);                               // max of the negated minimum
    v[NB + 3] = (int)((c.orbits | (c.orbits >> 16)) & 0xffffu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < NB + 3; ++k) v[k] = max(v[k], __shfl_xor_sync(0xffffffffu, v[k], o));
      v[NB + 3] |= __shfl_xor_sync(0xffffffffu, v[NB + 3], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NB + 4; ++k) red_m[k][warp] = v[k];
    }
  }
  const bool block_any = __syncthreads_or(any ? 1 : 0) != 0;
  if (warp < NB && block_any) {
    // warp b reduces band b: lanes stride over the per-thread slots of each quantity
    const int b = warp;
    long long q[kNQ];
#pragma unroll
    for (int k = 0; k < kNQ; ++k) {
      const unsigned long long* t = tot + (size_t)(b * kNQ + k) * kThreads;
      unsigned long long s = 0;
      for (int i = lane; i < kThreads; i += 32) s += t[i];
      q[k] = warp_sum_ll((long long)s);
    }
    if (lane == 0) {
      long long nn = 0;
      int md = 0;
      for (int w = 0; w < kThreads / 32; ++w) { nn += red_n[w]; md = max(md, red_m[b][w]); }
      long long sab = q[0], sx = q[1], sy = q[2], sxx = q[3], syy = q[4], sxy = q[5];
      if (DT == DM_I16) {
        const long long cc = 32768, c2 = 32768ll * 32768ll;
        const long long xx = sxx - 2 * cc * sx + c2 * nn, yy = syy - 2 * cc * sy + c2 * nn;
        const long long xy = sxy - cc * (sx + sy) + c2 * nn;
        sx -= cc * nn; sy -= cc * nn; sxx = xx; syy = yy; sxy = xy;
      }
      int64_t* O = g.sums + (int64_t)b * DM_NSTAT;
      if (nn) atomic_add_i64(O + DM_S_N, nn);
      if (sab) atomic_add_i64(O + DM_S_ABS, sab);
      if (sx) atomic_add_i64(O + DM_S_X, sx);
      if (sy) atomic_add_i64(O + DM_S_Y, sy);
      if (sxx) atomic_add_i64(O + DM_S_XX, sxx);
      if (syy) atomic_add_i64(O + DM_S_YY, syy);
      if (sxy) atomic_add_i64(O + DM_S_XY, sxy);
      const long long sse = sxx + syy - 2 * sxy;
      if (sse) atomic_add_i64(O + DM_S_SSE, sse);
      if (md) atomic_max_i64(g.maxs + (int64_t)b * DM_NSTAT + DM_M_MAXERR, md);
    }
  }
  if (tid == kThreads - 1 && block_any) {
    int msel = (int)0x80000000, umax = 0, nmin = (int)0x80000000, orb = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      msel = max(msel, red_m[NB + 0][w]); umax = max(umax, red_m[NB + 1][w]);
      nmin = max(nmin, red_m[NB + 2][w]); orb |= red_m[NB + 3][w];
    }
    int64_t* M = g.maxs;                                       // cube-wide values are reported on band 0
    if (DT == DM_I16) {
      const int hi = umax - 32768, lo = -nmin - 32768;
      if (hi > 0) atomic_max_i64(M + DM_M_UMAX, hi);
      if (lo < 0) atomic_max_i64(M + DM_M_UNEGMIN, -lo);
    } else if (umax > 0) {
      atomic_max_i64(M + DM_M_UMAX, umax);
    }
    if (msel > 0) atomic_max_i64(M + DM_M_ABSXY, msel);
    if (orb & 0xF) atomic_max_i64(M + DM_M_LOW4, 1);
    if (orb & 0x3) atomic_max_i64(M + DM_M_LOW2, 1);
  }
  if (ERR) {
    // histogram of min(e,255) -> the two planes' byte histograms through their LUTs
    unsigned cnt = 0;
#pragma unroll 8
    for (int j = 0; j < 32; ++j) cnt += hist[tid * 32 + ((j + tid) & 31)];
    if (cnt) {
      if (g.hist8_g) atomic_add_i64(g.hist8_g + lut_g[tid], cnt);
      if (g.hist8_z) atomic_add_i64(g.hist8_z + lut_z[tid], cnt);
    }
  }
}

template <int DT, bool MASK, bool ERR, int NB>
int run(const Args& g, cudaStream_t s) {
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const size_t smem = (size_t)NB * kNQ * kThreads * 8 + (ERR ? 256 * 32 * 4 + 512 : 0);
  auto k = fused_bsq_kernel<DT, MASK, ERR, NB>;
  DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nvec = g.npix / 8;
  int64_t grid = (nvec + kThreads - 1) / kThreads;
  if (grid > 2 * (int64_t)sms) grid = 2 * (int64_t)sms;
  if (grid < 1) grid = 1;
  k<<<(unsigned)grid, kThreads, smem, s>>>(g);
  DM_LAUNCH_CHECK("fused_bsq");
  return DM_OK;
}

template <int DT, bool MASK, bool ERR>
int run_nb(const Args& g, int nb, cudaStream_t s) {
  switch (nb) {
    case 1: return run<DT, MASK, ERR, 1>(g, s);
    case 2: return run<DT, MASK, ERR, 2>(g, s);
    case 3: return run<DT, MASK, ERR, 3>(g, s);
    case 4: return run<DT, MASK, ERR, 4>(g, s);
  }
  return fail(DM_EUNSUPPORTED, "dm_fused_bsq: 1..4 bands");
}

}  // namespace

int launch_fused_bsq(const dm_pair_t& p, const uint8_t* plane, int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                     const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z,
                     int cap_z, uint8_t* err8_z, int64_t* hist8_z, cudaStream_t s) {
  if (!p.ref || !p.tst || !sums || !maxs) return fail(DM_EARG, "dm_fused_bsq: null pointer");
  if (p.layout != DM_BSQ) return fail(DM_EUNSUPPORTED, "dm_fused_bsq: BSQ cubes only");
  if (p.dtype != DM_U16 && p.dtype != DM_I16) return fail(DM_EUNSUPPORTED, "dm_fused_bsq: 16-bit samples only");
  if (p.bands < 1 || p.bands > kMaxBands) return fail(DM_EUNSUPPORTED, "dm_fused_bsq: 1..4 bands (spectra are held in registers)");
  if (((reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) & 15) || (p.bands > 1 && (p.band_stride & 7)))
    return fail(DM_EUNSUPPORTED, "dm_fused_bsq: bands must start on 16-byte boundaries");
  if (err8_g && (!lut_g || cap_g < 0)) return fail(DM_EARG, "dm_fused_bsq: bad global LUT");
  if (err8_z && (!lut_z || cap_z < 0)) return fail(DM_EARG, "dm_fused_bsq: bad zoom LUT");
  if ((err8_g && cap_g > 255) || (err8_z && cap_z > 255))
    return fail(DM_EUNSUPPORTED, "dm_fused_bsq: caps above 255 take the separate passes");
  const int64_t npix = p.rows * p.width;
  if (npix <= 0) return DM_OK;
  if (plane && (reinterpret_cast<uintptr_t>(plane) & 7)) return fail(DM_EUNSUPPORTED, "dm_fused_bsq: plane must be 8-byte aligned");
  if (errmax_out && (reinterpret_cast<uintptr_t>(errmax_out) & 15)) return fail(DM_EUNSUPPORTED, "dm_fused_bsq: errmax must be 16-byte aligned");
  if ((err8_g && (reinterpret_cast<uintptr_t>(err8_g) & 7)) || (err8_z && (reinterpret_cast<uintptr_t>(err8_z) & 7)))
    return fail(DM_EUNSUPPORTED, "dm_fused_bsq: ERR8 planes must be 8-byte aligned");
  Args g;
  g.ref = p.ref; g.tst = p.tst; g.plane = plane; g.npix = npix; g.band_stride = p.band_stride;
  g.sums = sums; g.maxs = maxs; g.errmax = errmax_out;
  g.lut_g = lut_g; g.cap_g = cap_g; g.err8_g = err8_g; g.hist8_g = err8_g ? hist8_g : nullptr;
  g.lut_z = lut_z; g.cap_z = cap_z; g.err8_z = err8_z; g.hist8_z = err8_z ? hist8_z : nullptr;
  const bool err = errmax_out || err8_g || err8_z;
  const int nb = (int)p.bands;
  if (p.dtype == DM_U16) {
    if (plane) return err ? run_nb<DM_U16, true, true>(g, nb, s) : run_nb<DM_U16, true, false>(g, nb, s);
    return err ? run_nb<DM_U16, false, true>(g, nb, s) : run_nb<DM_U16, false, false>(g, nb, s);
  }
  if (plane) return err ? run_nb<DM_I16, true, true>(g, nb, s) : run_nb<DM_I16, true, false>(g, nb, s);
  return err ? run_nb<DM_I16, false, true>(g, nb, s) : run_nb<DM_I16, false, false>(g, nb, s);
}

}  // namespace dm
