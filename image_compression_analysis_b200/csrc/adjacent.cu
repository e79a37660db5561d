// Rows either side of the distortion path: the RGB quicklook, the baseline builders'
// requantisation and scene error maps, and the codec wrappers' reversible band differencing.  All of it
// is streaming integer work bounded by HBM; nothing here is staged through the host.
//
//   band_hist16        /root/reference/tools/quicklooks.py:51-72    exact 65536-bin value histograms (percentiles)
//   lut_bands          /root/reference/tools/quicklooks.py:81-89    stretch8 through host-built tables
//   requantize         /root/reference/tools/make_baseline_B.py:281-316 (k-LSB truncation, nodata kept)
//                      /root/reference/tools/make_baseline_A.py:166-167 (round to a multiple of 16)
//   scene_error        /root/reference/tools/make_baseline_B.py:324-419 (modes mean / rms / count3 / max / p95)
//   diff1              /root/reference/tools/codecs/ccsds121/ccsds121_wrap.py:66-85 (mod 2^16)
//                      /root/reference/tools/codecs/jpegls/jpegls_wrap.py:92-120   (mod 2^N, int16 saturating)

#include "dm_common.cuh"

namespace dm {

namespace {

// unsigned bin of a sample: int16 in offset binary so that bins are in value order
template <int DT> __device__ __forceinline__ unsigned bin_of(unsigned raw) {
  return DT == DM_I16 ? ((raw ^ 0x8000u) & 0xffffu) : raw;
}

// ------------------------------------------------------------------------------------------------
// Exact value histogram of selected bands.  One block keeps ALL 65536 bins in shared memory as packed
// 16-bit counters (128 KB) and flushes them to the int64 global histogram before any counter can
// overflow (a flush period covers at most 65535 pixels), so the pass reads every selected sample once
// and global atomics only see the distinct values of a period.
constexpr int kHistThreads = 1024;
constexpr int kHistPeriod = 61440;          // pixels per flush period (< 65536, multiple of kHistThreads)

struct HistArgs {
  const void* data;
  const uint8_t* plane;
  int plane_bit;
  int64_t npix, sb, sp;
  int sel[4];
  int64_t* hist;           // nsel x 65536
};

template <typename T, int DT>
__global__ void __launch_bounds__(kHistThreads, 1)
band_hist16_kernel(HistArgs g) {
  extern __shared__ unsigned packed[];            // 32768 words: bin v -> half (v & 1) of word v >> 1
  const T* src = static_cast<const T*>(g.data) + (int64_t)g.sel[blockIdx.y] * g.sb;
  int64_t* out = g.hist + (int64_t)blockIdx.y * 65536;
  for (int i = threadIdx.x; i < 32768; i += kHistThreads) packed[i] = 0u;
  __syncthreads();
  const int64_t nper = (g.npix + kHistPeriod - 1) / kHistPeriod;
  for (int64_t per = blockIdx.x; per < nper; per += gridDim.x) {
    const int64_t p0 = per * kHistPeriod;
    const int64_t p1 = p0 + kHistPeriod < g.npix ? p0 + kHistPeriod : g.npix;
    for (int64_t p = p0 + threadIdx.x; p < p1; p += kHistThreads) {
      if (g.plane && !(g.plane[p] & g.plane_bit)) continue;
      const unsigned v = bin_of<DT>((unsigned)(uint16_t)src[p * g.sp]);
      atomicAdd(&packed[v >> 1], 1u << (16 * (v & 1u)));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 32768; i += kHistThreads) {
      const unsigned w = packed[i];
      if (w) {
        if (w & 0xffffu) atomic_add_i64(out + 2 * i, (long long)(w & 0xffffu));
        if (w >> 16) atomic_add_i64(out + 2 * i + 1, (long long)(w >> 16));
        packed[i] = 0u;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// out[c][p] = lut[c][bin(sample(sel[c], p))]: the reference's float32 stretch is elementwise on an integer
// sample, so the host tabulates it with the reference's own expression and the planes are bit-exact.
struct LutArgs {
  const void* data;
  int64_t npix, sb, sp;
  int sel[4];
  int nsel;
  const uint8_t* luts;     // nsel x 65536
  uint8_t* out;            // nsel x npix
};

// The 64 KB table of the block's channel sits in shared memory (a gather there costs bank conflicts only;
// through L1 every divergent lane is a tag lookup); contiguous 16-bit bands are read 8 pixels per thread.
template <typename T, int DT>
__global__ void __launch_bounds__(512)
lut_bands_kernel(LutArgs g) {
  extern __shared__ __align__(16) uint8_t slut[];               // 65536 bytes
  const int c = blockIdx.y;
  const T* src = static_cast<const T*>(g.data) + (int64_t)g.sel[c] * g.sb;
  uint8_t* out = g.out + (int64_t)c * g.npix;
  {
    const uint4* gl = reinterpret_cast<const uint4*>(g.luts + (int64_t)c * 65536);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) reinterpret_cast<uint4*>(slut)[i] = __ldg(gl + i);
  }
  __syncthreads();
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (sizeof(T) == 2 && g.sp == 1 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) {
    const int64_t nv = g.npix >> 3;
    for (int64_t i = tid; i < nv; i += nth) {
      const uint4 v = ldg_stream16(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[2] = {0u, 0u};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const unsigned bin = bin_of<DT>((w[j >> 1] >> (16 * (j & 1))) & 0xffffu);
        o[j >> 2] |= (uint32_t)slut[bin] << (8 * (j & 3));
      }
      __stcs(reinterpret_cast<uint2*>(out) + i, make_uint2(o[0], o[1]));
    }
    done = nv << 3;
  }
  for (int64_t p = done + tid; p < g.npix; p += nth) out[p] = slut[bin_of<DT>((unsigned)(uint16_t)src[p * g.sp])];
}

// ------------------------------------------------------------------------------------------------
// Requantisation of 16-bit samples, two per 32-bit lane.
//   mode 0: ((u >> k) << k) on the uint16 view; samples equal to nodata keep their value
//   mode 1: ((u + 2^(k-1)) >> k) << k in uint16 arithmetic (the sum wraps, as numpy's does)
__device__ __forceinline__ unsigned requant2(unsigned w, int mode, unsigned keep, unsigned half2, unsigned nd2, bool has_nd) {
  unsigned r;
  if (mode == 0) {
    r = w & keep;
    if (has_nd) {
      const unsigned eq = __vcmpeq2(w, nd2);      // 0xffff per equal half
      r = (r & ~eq) | (w & eq);
    }
  } else {
    r = __vadd2(w, half2) & keep;
  }
  return r;
}

__global__ void __launch_bounds__(256)
requantize_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, int64_t n, int mode, int k,
                  int has_nd, unsigned nd) {
  const unsigned keep1 = (0xffffu >> k) << k, keep = keep1 | (keep1 << 16);
  const unsigned half1 = k > 0 ? (1u << (k - 1)) : 0u, half2 = half1 | (half1 << 16);
  const unsigned nd2 = (nd & 0xffffu) | (nd << 16);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  int64_t done = 0;
  if (aligned) {
    const int64_t nv = n >> 3;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int64_t i = tid; i < nv; i += nth) {
      uint4 v = ldg_stream16(s4 + i);
      v.x = requant2(v.x, mode, keep, half2, nd2, has_nd);
      v.y = requant2(v.y, mode, keep, half2, nd2, has_nd);
      v.z = requant2(v.z, mode, keep, half2, nd2, has_nd);
      v.w = requant2(v.w, mode, keep, half2, nd2, has_nd);
      __stcs(d4 + i, v);
    }
    done = nv << 3;
  }
  for (int64_t i = done + tid; i < n; i += nth) {
    const unsigned w = src[i];
    dst[i] = (uint16_t)requant2(w, mode, keep, half2, nd2, has_nd);
  }
}

// ------------------------------------------------------------------------------------------------
// Scene error maps: per pixel, walk the bands IN ORDER (the reference accumulates float32 planes band by
// band, make_baseline_B.py:345-361, so the rounding sequence of the rms mode is part of the result).
enum { EM_MEAN = 0, EM_RMS = 1, EM_COUNT3 = 2, EM_MAX = 3, EM_P95 = 4 };

template <int NW>                     // NW 64-bit words of packed 16-bit counters: 4 * NW histogram bins (p95)
struct PixelAcc {
  float acc;                          // mean: sum d   rms: sum d*d   (float32, rounded per band)
  unsigned cnt, mx;
  unsigned long long h[NW];
  __device__ __forceinline__ void init() {
    acc = 0.f; cnt = 0; mx = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) h[w] = 0ull;
  }
  template <int MODE> __device__ __forceinline__ void add(int d, int kmax) {
    if (MODE == EM_MEAN) acc = __double2float_rn((double)acc + (double)d);
    if (MODE == EM_RMS) acc = __double2float_rn((double)acc + (double)(int)((unsigned)d * (unsigned)d));   // int32 product wraps
    if (MODE == EM_COUNT3) cnt += d == kmax;
    if (MODE == EM_MAX) mx = max(mx, (unsigned)d & 0xffffu);
    if (MODE == EM_P95) {
      const int k = min(d, kmax);
      const unsigned long long inc = 1ull << (16 * (k & 3));
      if (NW == 1) h[0] += inc;
      else {
#pragma unroll
        for (int w = 0; w < NW; ++w) h[w] += (k >> 2) == w ? inc : 0ull;
      }
    }
  }
  template <int MODE> __device__ __forceinline__ float finish(int bands, int kmax, unsigned thr) const {
    if (MODE == EM_MEAN) return __fdiv_rn(acc, (float)bands);
    if (MODE == EM_RMS) return __fsqrt_rn(__fdiv_rn(acc, (float)bands));
    if (MODE == EM_COUNT3) return (float)(cnt & 0xffffu);
    if (MODE == EM_MAX) return (float)mx;
    // p95 (make_baseline_B.py:362-368): first bin k with cdf[k] >= thr, assigned only while the output
    // is still 0 -- so a hit at k = 0 leaves it open for k = 1 (kept as the reference has it)
    unsigned cdf = 0;
    float out = 0.f;
#pragma unroll
    for (int k = 0; k < 4 * NW; ++k) {
      if (k <= kmax) {
        cdf += (unsigned)(h[k >> 2] >> (16 * (k & 3))) & 0xffffu;
        if (cdf >= thr && out == 0.f) out = (float)k;
      }
    }
    return out;
  }
};

struct SceneArgs {
  const void* ref;
  const void* tst;
  const uint8_t* valid;      // (rows*width) nonzero = valid, or NULL
  int64_t bands, npix, sb;
  int kmax;
  unsigned thr;
  float* out;
  unsigned* out_max;         // bit pattern of the largest (non-negative) output value, atomicMax
};

__device__ __forceinline__ void block_max_to(unsigned* dst, float v) {
  // outputs are >= 0 (or NaN in the rms mode when the reference's int32 square wraps), so the bit
  // patterns order like the values
  unsigned b = __float_as_uint(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((threadIdx.x & 31) == 0 && b) atomicMax(dst, b);
}

// BSQ: thread per pixel (VEC2: per pixel PAIR, one 32-bit load per band and cube), every band coalesced
// across the warp; eight bands of both cubes in flight
template <typename T, int MODE, int NW, bool VEC2>
__global__ void __launch_bounds__(256)
scene_error_bsq(SceneArgs g) {
  const T* ref = static_cast<const T*>(g.ref);
  const T* tst = static_cast<const T*>(g.tst);
  const int B = (int)g.bands;
  float vmax = 0.f;
  if (VEC2) {
    const int64_t npair = g.npix >> 1;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < npair; q += (int64_t)gridDim.x * blockDim.x) {
      const int64_t p = 2 * q;
      const bool ok0 = !g.valid || g.valid[p], ok1 = !g.valid || g.valid[p + 1];
      PixelAcc<NW> a0, a1;
      a0.init(); a1.init();
      int b = 0;
      for (; b + 8 <= B; b += 8) {
        uint32_t x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          x[j] = ldg_stream4(ref + (b + j) * g.sb + p);
          y[j] = ldg_stream4(tst + (b + j) * g.sb + p);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int d0 = abs(sample16<sizeof(T) == 2 && T(-1) < T(0) ? DM_I16 : DM_U16>(x[j], 0) -
                             sample16<sizeof(T) == 2 && T(-1) < T(0) ? DM_I16 : DM_U16>(y[j], 0));
          const int d1 = abs(sample16<sizeof(T) == 2 && T(-1) < T(0) ? DM_I16 : DM_U16>(x[j], 1) -
                             sample16<sizeof(T) == 2 && T(-1) < T(0) ? DM_I16 : DM_U16>(y[j], 1));
          a0.template add<MODE>(ok0 ? d0 : 0, g.kmax);
          a1.template add<MODE>(ok1 ? d1 : 0, g.kmax);
        }
      }
      for (; b < B; ++b) {
        const int x0 = (int)__ldg(ref + b * g.sb + p), y0 = (int)__ldg(tst + b * g.sb + p);
        const int x1 = (int)__ldg(ref + b * g.sb + p + 1), y1 = (int)__ldg(tst + b * g.sb + p + 1);
        a0.template add<MODE>(ok0 ? abs(x0 - y0) : 0, g.kmax);
        a1.template add<MODE>(ok1 ? abs(x1 - y1) : 0, g.kmax);
      }
      const float v0 = a0.template finish<MODE>(B, g.kmax, g.thr), v1 = a1.template finish<MODE>(B, g.kmax, g.thr);
      *reinterpret_cast<float2*>(g.out + p) = make_float2(v0, v1);
      vmax = __uint_as_float(max(__float_as_uint(vmax), max(__float_as_uint(v0), __float_as_uint(v1))));
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < g.npix; p += (int64_t)gridDim.x * blockDim.x) {
      const bool ok = !g.valid || g.valid[p];
      PixelAcc<NW> a;
      a.init();
      int b = 0;
      for (; b + 8 <= B; b += 8) {
        int x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { x[j] = (int)__ldg(ref + (b + j) * g.sb + p); y[j] = (int)__ldg(tst + (b + j) * g.sb + p); }
#pragma unroll
        for (int j = 0; j < 8; ++j) a.template add<MODE>(ok ? abs(x[j] - y[j]) : 0, g.kmax);
      }
      for (; b < B; ++b) {
        const int x = (int)__ldg(ref + b * g.sb + p), y = (int)__ldg(tst + b * g.sb + p);
        a.template add<MODE>(ok ? abs(x - y) : 0, g.kmax);
      }
      const float v = a.template finish<MODE>(B, g.kmax, g.thr);
      g.out[p] = v;
      vmax = __uint_as_float(max(__float_as_uint(vmax), __float_as_uint(v)));
    }
  }
  block_max_to(g.out_max, vmax);
}

// BIP: a tile of kTile pixels (both cubes) is copied to shared memory with coalesced 16-byte (or 2-byte)
// loads, then thread t walks the spectrum of pixel t
constexpr int kSceneTile = 64;

template <typename T, int MODE, int NW>
__global__ void __launch_bounds__(kSceneTile)
scene_error_bip(SceneArgs g) {
  extern __shared__ __align__(16) unsigned char tile[];
  const int B = (int)g.bands;
  const int64_t tile_elems = (int64_t)kSceneTile * B;
  T* sa = reinterpret_cast<T*>(tile);
  T* sr = sa + tile_elems;
  const T* ref = static_cast<const T*>(g.ref);
  const T* tst = static_cast<const T*>(g.tst);
  const int64_t ntiles = (g.npix + kSceneTile - 1) / kSceneTile;
  const bool vec = ((reinterpret_cast<uintptr_t>(ref) | reinterpret_cast<uintptr_t>(tst)) & 15) == 0 &&
                   (tile_elems * sizeof(T)) % 16 == 0;
  float vmax = 0.f;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t p0 = t * kSceneTile;
    const int np = (int)((g.npix - p0) < kSceneTile ? (g.npix - p0) : kSceneTile);
    const int64_t e0 = p0 * B, ne = (int64_t)np * B;
    __syncthreads();
    if (vec && np == kSceneTile) {
      const int nv = (int)(ne * sizeof(T) / 16);
      const uint4* ga = reinterpret_cast<const uint4*>(ref + e0);
      const uint4* gr = reinterpret_cast<const uint4*>(tst + e0);
      for (int i = threadIdx.x; i < nv; i += kSceneTile) {
        reinterpret_cast<uint4*>(sa)[i] = ldg_stream16(ga + i);
        reinterpret_cast<uint4*>(sr)[i] = ldg_stream16(gr + i);
      }
    } else {
      for (int64_t i = threadIdx.x; i < ne; i += kSceneTile) { sa[i] = ref[e0 + i]; sr[i] = tst[e0 + i]; }
    }
    __syncthreads();
    if ((int)threadIdx.x < np) {
      const int64_t p = p0 + threadIdx.x;
      const bool ok = !g.valid || g.valid[p];
      PixelAcc<NW> a;
      a.init();
      const T* xa = sa + (int64_t)threadIdx.x * B;
      const T* xr = sr + (int64_t)threadIdx.x * B;
      int b = 0;
      if (sizeof(T) == 2 && (B & 1) == 0) {                   // two bands per shared-memory load, still in band order
        constexpr int SDT = T(-1) < T(0) ? DM_I16 : DM_U16;
        const uint32_t* wa = reinterpret_cast<const uint32_t*>(xa);
        const uint32_t* wr = reinterpret_cast<const uint32_t*>(xr);
        for (; b < B; b += 2) {
          const uint32_t x = wa[b >> 1], y = wr[b >> 1];
          a.template add<MODE>(ok ? abs(sample16<SDT>(x, 0) - sample16<SDT>(y, 0)) : 0, g.kmax);
          a.template add<MODE>(ok ? abs(sample16<SDT>(x, 1) - sample16<SDT>(y, 1)) : 0, g.kmax);
        }
      }
      for (; b < B; ++b) a.template add<MODE>(ok ? abs((int)xa[b] - (int)xr[b]) : 0, g.kmax);
      const float v = a.template finish<MODE>(B, g.kmax, g.thr);
      g.out[p] = v;
      vmax = __uint_as_float(max(__float_as_uint(vmax), __float_as_uint(v)));
    }
  }
  block_max_to(g.out_max, vmax);
}

// (np.clip(v, 0, emax) * (255.0/emax) + 0.5).astype(np.uint8): a float32 chain (make_baseline_B.py:417)
__global__ void __launch_bounds__(256)
scale_plane_u8_kernel(const float* __restrict__ plane, int64_t n, float emax, float scale, uint8_t* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float c = fminf(fmaxf(plane[i], 0.f), emax);
    out[i] = (uint8_t)(int)__fadd_rn(__fmul_rn(c, scale), 0.5f);
  }
}

// ------------------------------------------------------------------------------------------------
// Band differencing along the band axis of a BSQ cube.  Thread <-> one 16-byte vector of pixels (or one
// pixel in the unaligned fallback); the previous band stays in registers, so every sample is read once
// and written once.
//   arith 0: modulo 2^N   (uint16 / int16 view / uint8)      forward R[b] = X[b] - X[b-1], inverse running sum
//   arith 1: int16 saturating (jpegls_wrap.py:100-102, 114-116): R = clip(X[b] - X[b-1]), X[b] = clip(R[b] + X[b-1])
template <int EB, int ARITH> __device__ __forceinline__ unsigned vsub(unsigned a, unsigned b) {
  if (ARITH == 1) return __vsubss2(a, b);
  return EB == 2 ? __vsub2(a, b) : __vsub4(a, b);
}
template <int EB, int ARITH> __device__ __forceinline__ unsigned vadd(unsigned a, unsigned b) {
  if (ARITH == 1) return __vaddss2(a, b);
  return EB == 2 ? __vadd2(a, b) : __vadd4(a, b);
}

template <int EB, int ARITH, bool INVERSE>
__global__ void __launch_bounds__(256)
diff1_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t bands, int64_t nvec, int64_t sbv) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    uint4 prev = ldg_stream16(src + i);
    __stcs(dst + i, prev);
    int64_t b = 1;
    for (; b + 4 <= bands; b += 4) {
      uint4 c[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) c[j] = ldg_stream16(src + (b + j) * sbv + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 o;
        if (INVERSE) {
          o.x = vadd<EB, ARITH>(c[j].x, prev.x); o.y = vadd<EB, ARITH>(c[j].y, prev.y);
          o.z = vadd<EB, ARITH>(c[j].z, prev.z); o.w = vadd<EB, ARITH>(c[j].w, prev.w);
          prev = o;
        } else {
          o.x = vsub<EB, ARITH>(c[j].x, prev.x); o.y = vsub<EB, ARITH>(c[j].y, prev.y);
          o.z = vsub<EB, ARITH>(c[j].z, prev.z); o.w = vsub<EB, ARITH>(c[j].w, prev.w);
          prev = c[j];
        }
        __stcs(dst + (b + j) * sbv + i, o);
      }
    }
    for (; b < bands; ++b) {
      const uint4 c = ldg_stream16(src + b * sbv + i);
      uint4 o;
      if (INVERSE) {
        o.x = vadd<EB, ARITH>(c.x, prev.x); o.y = vadd<EB, ARITH>(c.y, prev.y);
        o.z = vadd<EB, ARITH>(c.z, prev.z); o.w = vadd<EB, ARITH>(c.w, prev.w);
        prev = o;
      } else {
        o.x = vsub<EB, ARITH>(c.x, prev.x); o.y = vsub<EB, ARITH>(c.y, prev.y);
        o.z = vsub<EB, ARITH>(c.z, prev.z); o.w = vsub<EB, ARITH>(c.w, prev.w);
        prev = c;
      }
      __stcs(dst + b * sbv + i, o);
    }
  }
}

template <typename T, int ARITH, bool INVERSE>
__global__ void __launch_bounds__(256)
diff1_scalar_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t bands, int64_t npix, int64_t sb) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    int prev = (int)src[p];
    dst[p] = (T)prev;
    for (int64_t b = 1; b < bands; ++b) {
      const int c = (int)src[b * sb + p];
      int o = INVERSE ? c + prev : c - prev;
      if (ARITH == 1) o = max(-32768, min(32767, o));
      dst[b * sb + p] = (T)o;                       // the cast wraps modulo 2^N
      prev = INVERSE ? (int)(T)o : c;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Interleave conversions.  BSQ (B,H,W) <-> BIL (H,B,W) moves whole rows; everything that involves BIP is
// a batched 2-D transpose [batch][R][C] -> [batch][C][R] through a padded shared-memory tile.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(const T* __restrict__ src, T* __restrict__ dst, int64_t R, int64_t Cn, int64_t tiles_fast, int r_fast,
                 int64_t batch0) {
  __shared__ T tile[64][65];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;      // 64 x 4
  const int64_t batch = batch0 + blockIdx.y;
  const T* s = src + batch * R * Cn;
  T* d = dst + batch * R * Cn;
  // consecutive blocks walk the SHORT axis (the bands): the partial lines a tile leaves on the interleaved side
  // are completed by its neighbours while they are still in L2
  const int64_t t_slow = (int64_t)blockIdx.x / tiles_fast, t_fast = (int64_t)blockIdx.x % tiles_fast;
  const int64_t r0 = (r_fast ? t_fast : t_slow) * 64, c0 = (r_fast ? t_slow : t_fast) * 64;
#pragma unroll 4
  for (int j = 0; j < 64; j += 4) {
    const int64_t r = r0 + ty + j, c = c0 + tx;
    if (r < R && c < Cn) tile[ty + j][tx] = s[r * Cn + c];
  }
  __syncthreads();
#pragma unroll 4
  for (int j = 0; j < 64; j += 4) {
    const int64_t c = c0 + ty + j, r = r0 + tx;
    if (r < R && c < Cn) d[c * R + r] = tile[tx][ty + j];
  }
}

// 16-bit elements, R and Cn even, 4-byte aligned cubes: global accesses are 32-bit words (two elements), so a
// warp moves 128 contiguous bytes per instruction on both sides; the tile is kept as 16-bit elements with an
// odd WORD pitch (66 elements = 33 words): row writes are conflict free, the column-pair reads two-way
__global__ void __launch_bounds__(256)
transpose16_pair_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t R, int64_t Cn, int64_t tiles_fast,
                        int r_fast, int64_t batch0) {
  __shared__ uint16_t tile[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 words x 8 rows
  const int64_t batch = batch0 + blockIdx.y;
  const uint32_t* s = src + batch * (R * Cn / 2);
  uint32_t* d = dst + batch * (R * Cn / 2);
  // consecutive blocks walk the SHORT axis (the bands): the partial lines a tile leaves on the interleaved side
  // are completed by its neighbours while they are still in L2
  const int64_t t_slow = (int64_t)blockIdx.x / tiles_fast, t_fast = (int64_t)blockIdx.x % tiles_fast;
  const int64_t r0 = (r_fast ? t_fast : t_slow) * 64, c0 = (r_fast ? t_slow : t_fast) * 64;
#pragma unroll
  for (int j = 0; j < 64; j += 8) {
    const int64_t r = r0 + ty + j, c = c0 + 2 * tx;
    if (r < R && c < Cn) *reinterpret_cast<uint32_t*>(&tile[ty + j][2 * tx]) = __ldg(s + (r * Cn + c) / 2);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 64; j += 8) {
    const int64_t c = c0 + ty + j, r = r0 + 2 * tx;
    if (r < R && c < Cn) d[(c * R + r) / 2] = (uint32_t)tile[2 * tx][ty + j] | ((uint32_t)tile[2 * tx + 1][ty + j] << 16);
  }
}

// dst row (i1, i0) <- src row (i0, i1), rows of `len` bytes: BSQ <-> BIL
__global__ void __launch_bounds__(256)
swap_rows_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, int64_t n0, int64_t n1, int64_t len) {
  const int64_t nrows = n0 * n1;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)len) & 15) == 0;
  for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
    const int64_t i0 = row / n1, i1 = row % n1;
    const unsigned char* s = src + row * len;
    unsigned char* d = dst + (i1 * n0 + i0) * len;
    if (vec) {
      for (int64_t i = threadIdx.x; i < (len >> 4); i += blockDim.x)
        __stcs(reinterpret_cast<uint4*>(d) + i, ldg_stream16(reinterpret_cast<const uint4*>(s) + i));
    } else {
      for (int64_t i = threadIdx.x; i < len; i += blockDim.x) d[i] = s[i];
    }
  }
}

int grid_for(int64_t work_items, int threads, int per_sm) {
  const int sms = sm_count();
  int64_t blocks = (work_items + threads - 1) / threads;
  const int64_t cap = (int64_t)(sms > 0 ? sms : 148) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int check_cube(const dm_cube_t& c, const char* who) {
  if (!c.data) return fail(DM_EARG, "%s: null cube", who);
  if (c.dtype != DM_U8 && c.dtype != DM_U16 && c.dtype != DM_I16) return fail(DM_EARG, "%s: bad dtype", who);
  if (c.layout != DM_BSQ && c.layout != DM_BIP) return fail(DM_EARG, "%s: layout must be DM_BSQ or DM_BIP", who);
  if (c.bands <= 0 || c.rows < 0 || c.width < 0) return fail(DM_EARG, "%s: bad geometry", who);
  if (c.layout == DM_BSQ && c.band_stride < c.rows * c.width) return fail(DM_EARG, "%s: band_stride < rows*width", who);
  return DM_OK;
}

}  // namespace

int launch_band_hist(const dm_cube_t& c, const int32_t* sel, int nsel, const uint8_t* plane, int plane_bit,
                     int64_t* hist, cudaStream_t s) {
  if (int rc = check_cube(c, "dm_band_hist")) return rc;
  if (!sel || !hist || nsel < 1 || nsel > 4) return fail(DM_EARG, "dm_band_hist: 1..4 selected bands and a histogram");
  HistArgs g{};
  g.data = c.data; g.plane = plane; g.plane_bit = plane_bit; g.npix = c.rows * c.width; g.hist = hist;
  g.sb = c.layout == DM_BSQ ? c.band_stride : 1;
  g.sp = c.layout == DM_BSQ ? 1 : c.bands;
  for (int i = 0; i < nsel; ++i) {
    if (sel[i] < 0 || sel[i] >= c.bands) return fail(DM_EARG, "dm_band_hist: band index %d out of range", (int)sel[i]);
    g.sel[i] = sel[i];
  }
  if (g.npix == 0) return DM_OK;
  const int64_t nper = (g.npix + kHistPeriod - 1) / kHistPeriod;
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const dim3 grid((unsigned)(nper < sms ? nper : sms), (unsigned)nsel);
  const size_t smem = 32768 * sizeof(unsigned);
#define DM_HIST(T, DT)                                                                                         \
  do {                                                                                                         \
    DM_CUDA(cudaFuncSetAttribute(band_hist16_kernel<T, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    band_hist16_kernel<T, DT><<<grid, kHistThreads, smem, s>>>(g);                                             \
  } while (0)
  if (c.dtype == DM_U8) DM_HIST(uint8_t, DM_U8);
  else if (c.dtype == DM_U16) DM_HIST(uint16_t, DM_U16);
  else DM_HIST(uint16_t, DM_I16);
#undef DM_HIST
  DM_LAUNCH_CHECK("band_hist16");
  return DM_OK;
}

int launch_lut_bands(const dm_cube_t& c, const int32_t* sel, int nsel, const uint8_t* luts, uint8_t* out, cudaStream_t s) {
  if (int rc = check_cube(c, "dm_lut_bands_u8")) return rc;
  if (!sel || !luts || !out || nsel < 1 || nsel > 4) return fail(DM_EARG, "dm_lut_bands_u8: 1..4 selected bands, tables and an output");
  LutArgs g{};
  g.data = c.data; g.npix = c.rows * c.width; g.nsel = nsel; g.luts = luts; g.out = out;
  g.sb = c.layout == DM_BSQ ? c.band_stride : 1;
  g.sp = c.layout == DM_BSQ ? 1 : c.bands;
  for (int i = 0; i < nsel; ++i) {
    if (sel[i] < 0 || sel[i] >= c.bands) return fail(DM_EARG, "dm_lut_bands_u8: band index %d out of range", (int)sel[i]);
    g.sel[i] = sel[i];
  }
  if (g.npix == 0) return DM_OK;
  const dim3 grid((unsigned)grid_for((g.npix + 7) / 8, 512, 3), (unsigned)nsel);      // three 64 KB tables per SM
  const size_t smem = 65536;
#define DM_LUT(T, DT)                                                                                          \
  do {                                                                                                         \
    DM_CUDA(cudaFuncSetAttribute(lut_bands_kernel<T, DT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    lut_bands_kernel<T, DT><<<grid, 512, smem, s>>>(g);                                                        \
  } while (0)
  if (c.dtype == DM_U8) DM_LUT(uint8_t, DM_U8);
  else if (c.dtype == DM_U16) DM_LUT(uint16_t, DM_U16);
  else DM_LUT(uint16_t, DM_I16);
#undef DM_LUT
  DM_LAUNCH_CHECK("lut_bands");
  return DM_OK;
}

int launch_requantize(const void* src, void* dst, int dtype, int64_t n, int mode, int k, int has_nodata, int nodata,
                      cudaStream_t s) {
  if (!src || !dst) return fail(DM_EARG, "dm_requantize: null pointer");
  if (dtype != DM_U16 && dtype != DM_I16) return fail(DM_EUNSUPPORTED, "dm_requantize: 16-bit samples only");
  if (n < 0 || k < 0 || k > 15 || (mode != 0 && mode != 1)) return fail(DM_EARG, "dm_requantize: bad arguments");
  if (n == 0) return DM_OK;
  requantize_kernel<<<grid_for((n + 7) / 8, 256, 8), 256, 0, s>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst),
                                                                   n, mode, k, has_nodata && mode == 0, (unsigned)nodata & 0xffffu);
  DM_LAUNCH_CHECK("requantize");
  return DM_OK;
}

template <typename T>
static int scene_dispatch(const dm_pair_t& p, SceneArgs& g, int mode, cudaStream_t s) {
  const bool bsq = p.layout == DM_BSQ;
  const size_t smem = 2 * (size_t)kSceneTile * (size_t)p.bands * sizeof(T);
  if (!bsq && smem > 200 * 1024) return fail(DM_EUNSUPPORTED, "dm_scene_error: too many bands for the BIP tile");
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  // two pixels per thread through 32-bit loads when the geometry allows it (16-bit samples, even strides)
  const bool vec2 = bsq && sizeof(T) == 2 && g.npix % 2 == 0 && g.sb % 2 == 0 &&
                    ((reinterpret_cast<uintptr_t>(g.ref) | reinterpret_cast<uintptr_t>(g.tst)) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(g.out) & 7) == 0;
  const bool small_hist = g.kmax <= 3;          // p95 with k_bits <= 2 (the reference's k): one counter word
#define DM_SCENE_K(MODE, NW)                                                                                    \
  do {                                                                                                          \
    if (bsq) {                                                                                                  \
      if (vec2) scene_error_bsq<T, MODE, NW, sizeof(T) == 2><<<grid_for(g.npix / 2, 256, 8), 256, 0, s>>>(g);   \
      else scene_error_bsq<T, MODE, NW, false><<<grid_for(g.npix, 256, 8), 256, 0, s>>>(g);                     \
    } else {                                                                                                    \
      DM_CUDA(cudaFuncSetAttribute(scene_error_bip<T, MODE, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      const int64_t fit = (int64_t)(200 * 1024) / (int64_t)smem, ntiles = (g.npix + kSceneTile - 1) / kSceneTile;      \
      const int64_t cap = (int64_t)sms * (fit < 1 ? 1 : (fit > 16 ? 16 : fit));                                         \
      scene_error_bip<T, MODE, NW><<<(unsigned)(ntiles < cap ? ntiles : cap), kSceneTile, smem, s>>>(g);                \
    }                                                                                                           \
  } while (0)
#define DM_SCENE(MODE)                                                                                          \
  do {                                                                                                          \
    if (MODE == EM_P95 && !small_hist) DM_SCENE_K(MODE, 4);                                                     \
    else DM_SCENE_K(MODE, 1);                                                                                   \
  } while (0)
  switch (mode) {
    case EM_MEAN: DM_SCENE(EM_MEAN); break;
    case EM_RMS: DM_SCENE(EM_RMS); break;
    case EM_COUNT3: DM_SCENE(EM_COUNT3); break;
    case EM_MAX: DM_SCENE(EM_MAX); break;
    case EM_P95: DM_SCENE(EM_P95); break;
    default: return fail(DM_EARG, "dm_scene_error: bad mode");
  }
#undef DM_SCENE
#undef DM_SCENE_K
  DM_LAUNCH_CHECK("scene_error");
  return DM_OK;
}

int launch_scene_error(const dm_pair_t& p, const uint8_t* valid, int mode, int k_bits, uint32_t p95_thr, float* out_plane,
                       uint32_t* out_max_bits, cudaStream_t s) {
  if (!p.ref || !p.tst || !out_plane || !out_max_bits) return fail(DM_EARG, "dm_scene_error: null pointer");
  if (p.layout != DM_BSQ && p.layout != DM_BIP) return fail(DM_EARG, "dm_scene_error: bad layout");
  if (p.bands <= 0 || p.bands > 65535 || p.rows < 0 || p.width < 0) return fail(DM_EARG, "dm_scene_error: bad geometry");
  if (k_bits < 0 || k_bits > 16) return fail(DM_EARG, "dm_scene_error: k_bits out of range");
  if (mode == EM_P95 && k_bits > 4) return fail(DM_EUNSUPPORTED, "dm_scene_error: p95 keeps at most 16 bins per pixel (k_bits <= 4)");
  SceneArgs g{};
  g.ref = p.ref; g.tst = p.tst; g.valid = valid; g.bands = p.bands; g.npix = p.rows * p.width;
  g.sb = p.layout == DM_BSQ ? p.band_stride : 1;
  g.kmax = (1 << k_bits) - 1; g.thr = p95_thr; g.out = out_plane; g.out_max = out_max_bits;
  if (g.npix == 0) return DM_OK;
  switch (p.dtype) {
    case DM_U8: return scene_dispatch<uint8_t>(p, g, mode, s);
    case DM_U16: return scene_dispatch<uint16_t>(p, g, mode, s);
    case DM_I16: return scene_dispatch<int16_t>(p, g, mode, s);
  }
  return fail(DM_EARG, "dm_scene_error: bad dtype");
}

int launch_scale_plane_u8(const float* plane, int64_t n, float emax, float scale, uint8_t* out, cudaStream_t s) {
  if (!plane || !out) return fail(DM_EARG, "dm_scale_plane_u8: null pointer");
  if (n <= 0) return DM_OK;
  scale_plane_u8_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(plane, n, emax, scale, out);
  DM_LAUNCH_CHECK("scale_plane_u8");
  return DM_OK;
}

int launch_diff1(const void* src, void* dst, int dtype, int arith, int inverse, int64_t bands, int64_t npix,
                 int64_t band_stride, cudaStream_t s) {
  if (!src || !dst) return fail(DM_EARG, "dm_diff1: null pointer");
  if (bands <= 0 || npix < 0 || band_stride < npix) return fail(DM_EARG, "dm_diff1: bad geometry");
  if (arith != 0 && arith != 1) return fail(DM_EARG, "dm_diff1: arith must be 0 (modulo) or 1 (saturating)");
  if (arith == 1 && dtype != DM_I16) return fail(DM_EARG, "dm_diff1: saturating arithmetic is the int16 variant");
  if (npix == 0) return DM_OK;
  const int eb = elem_bytes(dtype);
  const int per16 = 16 / eb;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 &&
                   band_stride % per16 == 0 && npix % per16 == 0;
#define DM_DIFF_VEC(EB, AR, INV)                                                                                  \
  diff1_vec_kernel<EB, AR, INV><<<grid_for(npix / per16, 256, 8), 256, 0, s>>>(                                   \
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), bands, npix / per16, band_stride / per16)
#define DM_DIFF_SC(T, AR, INV)                                                                                    \
  diff1_scalar_kernel<T, AR, INV><<<grid_for(npix, 256, 8), 256, 0, s>>>(static_cast<const T*>(src), static_cast<T*>(dst), \
                                                                        bands, npix, band_stride)
  if (vec) {
    if (eb == 1) { if (inverse) DM_DIFF_VEC(1, 0, true); else DM_DIFF_VEC(1, 0, false); }
    else if (arith == 0) { if (inverse) DM_DIFF_VEC(2, 0, true); else DM_DIFF_VEC(2, 0, false); }
    else { if (inverse) DM_DIFF_VEC(2, 1, true); else DM_DIFF_VEC(2, 1, false); }
  } else {
    if (eb == 1) { if (inverse) DM_DIFF_SC(uint8_t, 0, true); else DM_DIFF_SC(uint8_t, 0, false); }
    else if (arith == 0) { if (inverse) DM_DIFF_SC(uint16_t, 0, true); else DM_DIFF_SC(uint16_t, 0, false); }
    else { if (inverse) DM_DIFF_SC(int16_t, 1, true); else DM_DIFF_SC(int16_t, 1, false); }
  }
#undef DM_DIFF_VEC
#undef DM_DIFF_SC
  DM_LAUNCH_CHECK("diff1");
  return DM_OK;
}

int launch_interleave(const void* src, void* dst, int eb, int from, int to, int64_t bands, int64_t rows, int64_t width,
                      cudaStream_t s) {
  if (!src || !dst) return fail(DM_EARG, "dm_interleave: null pointer");
  if (eb != 1 && eb != 2) return fail(DM_EARG, "dm_interleave: elem_bytes must be 1 or 2");
  if (from < 0 || from > 2 || to < 0 || to > 2) return fail(DM_EARG, "dm_interleave: layouts are DM_BSQ / DM_BIP / DM_BIL");
  if (bands <= 0 || rows < 0 || width < 0) return fail(DM_EARG, "dm_interleave: bad geometry");
  const int64_t n = bands * rows * width;
  if (n == 0) return DM_OK;
  if (from == to) {
    DM_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * eb, cudaMemcpyDeviceToDevice, s));
    return DM_OK;
  }
  if ((from == DM_BSQ && to == DM_BIL) || (from == DM_BIL && to == DM_BSQ)) {
    const int64_t n0 = from == DM_BSQ ? bands : rows, n1 = from == DM_BSQ ? rows : bands;
    swap_rows_kernel<<<grid_for(n0 * n1 * 256, 256, 16), 256, 0, s>>>(static_cast<const unsigned char*>(src),
                                                                     static_cast<unsigned char*>(dst), n0, n1, width * eb);
    DM_LAUNCH_CHECK("swap_rows");
    return DM_OK;
  }
  // transposes: [batch][R][C] -> [batch][C][R]
  int64_t batch, R, Cn;
  if (from == DM_BSQ) { batch = 1; R = bands; Cn = rows * width; }            // -> BIP
  else if (from == DM_BIL) { batch = rows; R = bands; Cn = width; }           // -> BIP
  else if (to == DM_BSQ) { batch = 1; R = rows * width; Cn = bands; }         // BIP ->
  else { batch = rows; R = width; Cn = bands; }                               // BIP -> BIL
  const int64_t gx = (Cn + 63) / 64, gy = (R + 63) / 64;
  if (gx * gy > 0x7fffffffll) return fail(DM_EUNSUPPORTED, "dm_interleave: cube too large for one launch");
  const int r_fast = R < Cn ? 1 : 0;
  const int64_t tiles_fast = r_fast ? gy : gx;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const dim3 grid((unsigned)(gx * gy), (unsigned)((batch - b0) < 65535 ? (batch - b0) : 65535));
    if (eb == 1)
      transpose_kernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), R, Cn, tiles_fast, r_fast, b0);
    else if (R % 2 == 0 && Cn % 2 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0)
      transpose16_pair_kernel<<<grid, 256, 0, s>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), R, Cn, tiles_fast, r_fast, b0);
    else
      transpose_kernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<const uint16_t*>(src), static_cast<uint16_t*>(dst), R, Cn, tiles_fast, r_fast, b0);
    DM_LAUNCH_CHECK("transpose");
  }
  return DM_OK;
}

}  // namespace dm
