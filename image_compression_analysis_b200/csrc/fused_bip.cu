// One-pass BIP kernel: per-band integer statistics AND per-pixel spectral metrics from a single
// read of both cubes (4 bytes per sample pair, the algorithmic minimum of SURVEY.md 8d).
//
//   per band   {N, Sx, Sy, Sxx, Syy, Sxy, S|d|, SSE, max|d|}, data-range scan   run_codec.py:268-285, 86-117
//   per pixel  max_b |d| -> ERR8 planes (quicklooks.py:123-150), SAM (run_codec.py:328-332)
//
// The two reductions run along different axes of the same (pixels x bands) tile, so the tile is
// staged ONCE in shared memory and consumed by two specialised warp groups at the same time:
//
//   warp 0            producer: one elected lane streams tiles of P pixels (both cubes) into a
//                     4-stage shared-memory ring with cp.async.bulk (TMA, 1-D) + mbarrier tx counts,
//                     so up to three tiles (138 KB) are in flight per SM while one is consumed;
//   warps 1..6        "band" group: thread <-> 4 fixed bands (one 8-byte column pair), walks the
//                     tile's pixels two at a time and pairs them band-wise with PRMT so that the
//                     packed dp2a arithmetic of stats.cu applies; accumulators stay in registers
//                     for the whole kernel;
//   warps 7..10       "pixel" group: thread <-> pixel, walks the spectrum in natural (band-pair)
//                     words: packed |d| max, dp2a dot / |a|^2 / |r|^2 as lo/hi 32-bit partials that
//                     cannot overflow for B <= 256 bands, then float64 sqrt/div/acos per pixel.  A tile
//                     has 64 pixels, so the two halves of the group take alternate tiles.
//
// A stage is released (empty mbarrier) when every consumer warp that reads it has arrived.  The kernel is
// persistent: one CTA per SM, tiles strided over CTAs.  Shared-memory reads are conflict free for
// EnMAP's 180 bands (pixel pitch 45 x 8 B, odd).  See DESIGN.md for the instruction budget.

#include <cstdlib>
#include <type_traits>

#include "dm_common.cuh"

namespace dm {

namespace {

// Band-group threads own C 32-bit columns (2C bands).  C = 1 needs half the accumulator registers, so
// twelve band warps fit: with four pixel warps every scheduler (warp id mod 4) then hosts 3 band + 1
// pixel warp.  With C = 2 (six band warps) two of the four schedulers carry 2 band + 1 pixel warp and
// the others 1 + 1, and the tile barrier makes everyone wait for the loaded ones: measured 215 us vs
// the balanced layout.
constexpr int kPixelWarps = 4;
constexpr int kPixelThreads = kPixelWarps * 32;
template <int C> struct Shape {
  static constexpr int kBandWarps = C == 2 ? 6 : 12;
  static constexpr int kBandThreads = kBandWarps * 32;
  static constexpr int kThreads = 32 + kBandThreads + kPixelThreads;     // 352 / 544
};
constexpr int kMaxBandThreads = 12 * 32;
constexpr int kStages = 4;
constexpr int kTilePixels = 64;                                  // pixels per tile (one half of the pixel group)
constexpr int kStageBytesMax = 46080;                            // 2 cubes x 64 px x 180 bands x 2 B
constexpr int kMaxSpecBlocks = 1184;                             // == dm_spectral_nblocks()

struct FusedArgs {
  const void* ref;
  const void* tst;
  const uint8_t* plane;       // may be null
  int64_t npix;
  int bands;
  int P;                      // pixels per tile (even)
  int64_t ntiles;             // FULL tiles (TMA); the leftover pixels form one generic tile
  int tail_pixels;
  int64_t* sums;
  int64_t* maxs;
  uint16_t* errmax;
  const uint8_t* lut_g; int cap_g; uint8_t* err8_g; int64_t* hist8_g;
  const uint8_t* lut_z; int cap_z; uint8_t* err8_z; int64_t* hist8_z;
  int want_sam;
  int debug;                  // experiments only: 1 = producer skips the copies (compute-only timing)
  double* spec_out;           // [3 * kMaxSpecBlocks]
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) { uint32_t r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t vminu2(uint32_t a, uint32_t b) { uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ int hmax2(uint32_t p) { return max((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmin2(uint32_t p) { return min((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmax2s(uint32_t p) { return max((int)(short)(p & 0xffffu), (int)(short)(p >> 16)); }

// per-band packed accumulators (same scheme as stats.cu)
struct BandAcc {
  uint32_t sabs, sx, sy, xxl, xxh, yyl, yyh, xyl, xyh, maxd;
  unsigned long long t_abs, t_x, t_y, t_xx, t_yy, t_xy;
  __device__ __forceinline__ void reset() {
    sabs = sx = sy = xxl = xxh = yyl = yyh = xyl = xyh = maxd = 0;
    t_abs = t_x = t_y = t_xx = t_yy = t_xy = 0;
  }
  __device__ __forceinline__ void spill() {
    t_abs += sabs; t_x += sx; t_y += sy; sabs = sx = sy = 0;
    t_xx += (unsigned long long)xxl + ((unsigned long long)xxh << 8); xxl = xxh = 0;
    t_yy += (unsigned long long)yyl + ((unsigned long long)yyh << 8); yyl = yyh = 0;
    t_xy += (unsigned long long)xyl + ((unsigned long long)xyh << 8); xyl = xyh = 0;
  }
};

// one packed word of one band (two pixels); x,y already in the unsigned domain and masked.
// TRACK: also fold max(x,y) into maxsel_u (the masked / int16 variants; the plain uint16 variant
// takes the maxima of the natural words instead, two words per VIMNMX3)
template <bool PAIR, bool TRACK>
__device__ __forceinline__ void band_word(BandAcc& a, uint32_t x, uint32_t y, uint32_t& maxsel_u) {
  const uint32_t mx = vmaxu2(x, y), mn = vminu2(x, y), d = mx - mn;
  if (TRACK) maxsel_u = vmaxu2(maxsel_u, mx);
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
}

__device__ __forceinline__ void hist_add(unsigned* h, unsigned bin) {
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, bin);
  if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[bin], (unsigned)__popc(peers));
}

// arccos for the SAM mean.  Spectral angles of decoded imagery are small, so c sits next to 1 where
// 1-c is exact (Sterbenz) and acos(c) = 2*asin(sqrt((1-c)/2)); the odd series below is accurate to
// < 1e-16 relative for angles up to 0.12 rad.  Anything else takes libdevice's acos.
__device__ __forceinline__ double acos_sam(double c) {
  const double e = 1.0 - c;
  if (e >= 0.0 && e < 0.0072) {
    const double s2 = 0.5 * e, s = __dsqrt_rn(s2);
    double p = 945.0 / 42240.0;
    p = fma(p, s2, 105.0 / 3456.0);
    p = fma(p, s2, 15.0 / 336.0);
    p = fma(p, s2, 3.0 / 40.0);
    p = fma(p, s2, 1.0 / 6.0);
    p = fma(p, s2, 1.0);
    return 2.0 * s * p;
  }
  return acos(c);
}

template <int DT, bool MASK, bool ERR, int C>
__global__ void __launch_bounds__(Shape<C>::kThreads, 1)
fused_bip_kernel(FusedArgs g) {
  constexpr int kBandWarps = Shape<C>::kBandWarps, kBandThreads = Shape<C>::kBandThreads;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  __shared__ unsigned h8g[256], h8z[256];
  __shared__ double red[3][kPixelWarps];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int B = g.bands, W = B >> 1;                   // W: 32-bit words per pixel
  const int P = g.P;
  const uint32_t cube_bytes = (uint32_t)P * (uint32_t)B * 2u;     // one cube's share of a stage
  const uint32_t stage_bytes = 2u * cube_bytes;
  constexpr uint32_t OFS = DT == DM_I16 ? 0x80008000u : 0u;
  constexpr bool TRACK = MASK || DT == DM_I16;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kBandWarps + kPixelWarps / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 256) { h8g[tid] = 0; h8z[tid] = 0; }
  __syncthreads();

  // tiles of this CTA: full tiles t = blockIdx.x + k*gridDim.x, then (one CTA) the tail tile
  const int64_t total_tiles = g.ntiles + (g.tail_pixels ? 1 : 0);
  const char* ref8 = static_cast<const char*>(g.ref);
  const char* tst8 = static_cast<const char*>(g.tst);

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    int it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      mbar_wait(&empty_bar[s], ph ^ 1u);             // first pass over the ring: passes immediately
      unsigned char* dst = smem + (size_t)s * stage_bytes;
      const int64_t off = t * (int64_t)cube_bytes;
      if (g.debug == 1) {
        if (lane == 0) mbar_arrive(&full_bar[s]);
      } else if (t < g.ntiles) {
        if (lane == 0) {
          mbar_expect_tx(&full_bar[s], stage_bytes);
          bulk_g2s(dst, ref8 + off, cube_bytes, &full_bar[s]);
          bulk_g2s(dst + cube_bytes, tst8 + off, cube_bytes, &full_bar[s]);
        }
      } else {
        // tail tile (< P pixels, any size): generic 4-byte copies by the whole warp
        const int nwords = g.tail_pixels * W;
        const uint32_t* rs = reinterpret_cast<const uint32_t*>(ref8 + off);
        const uint32_t* ts = reinterpret_cast<const uint32_t*>(tst8 + off);
        uint32_t* d0 = reinterpret_cast<uint32_t*>(dst);
        uint32_t* d1 = reinterpret_cast<uint32_t*>(dst + cube_bytes);
        for (int i = lane; i < nwords; i += 32) { d0[i] = ldg_stream4(rs + i); d1[i] = ldg_stream4(ts + i); }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp <= kBandWarps) {
    // ------------------------------------------------------------------ band group
    constexpr int NB = 2 * C;                        // bands per thread
    const int ts = tid - 32;                         // 0..kBandThreads-1
    const int tpp = W / C;                           // threads per pixel (C 32-bit columns each)
    const int nslots = kBandThreads / tpp;           // pixel pairs processed side by side
    const int slot = ts / tpp;
    const bool active = slot < nslots;
    const int col = (ts - slot * tpp) * C;           // first 32-bit column
    BandAcc a[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) a[j].reset();
    uint32_t maxsel_u = 0, maxsel_s = 0, umax = 0, umin = 0xffffffffu, orbits = 0, ymax = 0;
    long long n = 0;
    int since_spill = 0;
    bool any = false;

    auto pair_step = [&](const uint32_t (&xaw)[C], const uint32_t (&xbw)[C], const uint32_t (&yaw)[C],
                         const uint32_t (&ybw)[C], uint32_t m, auto tag) {
      constexpr bool PAIR = decltype(tag)::value;
#pragma unroll
      for (int k = 0; k < C; ++k) {
        // data-range scan on the raw reference words (unmasked)
        orbits |= xaw[k] | xbw[k];
        const uint32_t ua = xaw[k] ^ OFS, ub = xbw[k] ^ OFS, va = yaw[k] ^ OFS, vb = ybw[k] ^ OFS;
        umax = __vimax3_u16x2(umax, ua, ub);
        if (DT == DM_I16) umin = __vimin3_u16x2(umin, ua, ub);
        if (!TRACK) ymax = __vimax3_u16x2(ymax, va, vb);
        uint32_t x0 = __byte_perm(ua, ub, 0x5410), x1 = __byte_perm(ua, ub, 0x7632);
        uint32_t y0 = __byte_perm(va, vb, 0x5410), y1 = __byte_perm(va, vb, 0x7632);
        if (MASK) { x0 &= m; x1 &= m; y0 &= m; y1 &= m; }
        if (DT == DM_I16) {
          // np.abs semantics on the signed samples (wrapping abs of -32768 never wins)
          const uint32_t mm = MASK ? m : 0xffffffffu;
          const uint32_t sx0 = (x0 ^ OFS) & mm, sx1 = (x1 ^ OFS) & mm, sy0 = (y0 ^ OFS) & mm, sy1 = (y1 ^ OFS) & mm;
          maxsel_s = __vimax3_s16x2(maxsel_s, __vabs2(sx0), __vabs2(sy0));
          maxsel_s = __vimax3_s16x2(maxsel_s, __vabs2(sx1), __vabs2(sy1));
        }
        band_word<PAIR, TRACK>(a[2 * k], x0, y0, maxsel_u);
        band_word<PAIR, TRACK>(a[2 * k + 1], x1, y1, maxsel_u);
      }
    };
    auto lds_cols = [&](const unsigned char* p, uint32_t (&w)[C]) {
      if (C == 2) { const uint2 v = *reinterpret_cast<const uint2*>(p); w[0] = v.x; w[C - 1] = v.y; }
      else { w[0] = *reinterpret_cast<const uint32_t*>(p); }
    };

    // shared-memory walk of this thread: pair q = slot, slot+nslots, ... ; pixel 2q sits at byte
    // 2q*W*4 of the cube's share, its partner W*4 bytes further, the next pair pstep further
    const uint32_t w4 = (uint32_t)W * 4u;
    const uint32_t pstep = 2u * (uint32_t)nslots * w4;
    const uint32_t first = (2u * (uint32_t)slot * (uint32_t)W + (uint32_t)col) * 4u;
    const int full_steps = active ? ((P >> 1) - slot + nslots - 1) / nslots : 0;   // steps of a full tile

    int it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int cnt = t < g.ntiles ? P : g.tail_pixels;
      mbar_wait(&full_bar[s], ph);
      if (active) {
        const unsigned char* xs = smem + (size_t)s * stage_bytes + first;
        const uint8_t* pl = MASK ? g.plane + t * (int64_t)P : nullptr;
        any = true;
        const int steps = cnt == P ? full_steps : ((cnt >> 1) - slot + nslots - 1) / nslots;
        if (since_spill + steps > 127) {
          since_spill = 0;
#pragma unroll
          for (int j = 0; j < NB; ++j) a[j].spill();
        }
        since_spill += steps > 0 ? steps : 0;
        if (!MASK) n += steps > 0 ? 2 * steps : 0;
#pragma unroll 4
        for (int i = 0; i < steps; ++i) {
          const unsigned char* px = xs + (uint32_t)i * pstep;
          uint32_t xa[C], xb[C], ya[C], yb[C];
          lds_cols(px, xa); lds_cols(px + w4, xb); lds_cols(px + cube_bytes, ya); lds_cols(px + cube_bytes + w4, yb);
          uint32_t m = 0xffffffffu;
          if (MASK) {
            const int pa = 2 * (slot + i * nslots);
            m = ((pl[pa] & DM_VALID_METRICS) ? 0xffffu : 0u) | ((pl[pa + 1] & DM_VALID_METRICS) ? 0xffff0000u : 0u);
            n += (m & 1u) + (m >> 31);
          }
          pair_step(xa, xb, ya, yb, m, std::true_type());
        }
        if ((cnt & 1) && slot == 0) {                // odd leftover pixel of the tail tile
          const int pa = cnt - 1;
          const unsigned char* px = smem + (size_t)s * stage_bytes + ((size_t)pa * W + col) * 4;
          uint32_t xa[C], ya[C];
          lds_cols(px, xa); lds_cols(px + cube_bytes, ya);
          uint32_t m = 0xffffffffu;
          if (MASK) m = (pl[pa] & DM_VALID_METRICS) ? 0xffffffffu : 0u;
          n += m ? 1 : 0;
#pragma unroll
          for (int j = 0; j < NB; ++j) a[j].spill();
          since_spill = 1;
          pair_step(xa, xa, ya, ya, m, std::false_type());
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) a[j].spill();
    if (!TRACK) maxsel_u = vmaxu2(umax, ymax);       // unmasked uint16: max over everything read
    // ---- combine the band group through shared memory (stage memory is free once every tile is done).
    // No atomics: every thread stores its 8 per-band values at [quantity][j][slot][column pair]
    // (consecutive lanes -> consecutive 8-byte words), then one thread per band adds the slots.
    asm volatile("bar.sync 1, %0;" ::"r"(kBandThreads + kPixelThreads));       // consumers only
    unsigned long long* sh_part = reinterpret_cast<unsigned long long*>(smem);      // [8][NB][nslots*tpp]
    const int lanes_used = nslots * tpp;
    __shared__ int sh_cube[8];
    if (ts < 8) sh_cube[ts] = ts == 2 ? 0x7fffffff : (ts == 0 ? (int)0x80000000 : 0);
    if (active) {
      const int me = slot * tpp + col / C;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        unsigned long long* d = sh_part + (size_t)j * lanes_used + me;
        const size_t q = (size_t)NB * lanes_used;
        d[0 * q] = (unsigned long long)n;
        d[1 * q] = a[j].t_x; d[2 * q] = a[j].t_y; d[3 * q] = a[j].t_xx; d[4 * q] = a[j].t_yy; d[5 * q] = a[j].t_xy;
        d[6 * q] = a[j].t_abs; d[7 * q] = (unsigned long long)hmax2(a[j].maxd);
      }
    }
    asm volatile("bar.sync 2, %0;" ::"r"(kBandThreads));
    if (active && any) {
      atomicMax(sh_cube + 0, DT == DM_I16 ? hmax2s(maxsel_s) : hmax2(maxsel_u));
      atomicMax(sh_cube + 1, hmax2(umax));
      atomicMin(sh_cube + 2, hmin2(umin));
      atomicOr(reinterpret_cast<unsigned*>(sh_cube + 3), (orbits | (orbits >> 16)) & 0xffffu);
      sh_cube[4] = 1;
    }
    for (int b = ts; b < B; b += kBandThreads) {
      const int cp = b / NB, j = b - cp * NB;
      const size_t q = (size_t)NB * lanes_used;
      unsigned long long v[7] = {0, 0, 0, 0, 0, 0, 0};
      int md = 0;
      for (int sl = 0; sl < nslots; ++sl) {
        const unsigned long long* d = sh_part + (size_t)j * lanes_used + sl * tpp + cp;
#pragma unroll
        for (int k = 0; k < 7; ++k) v[k] += d[k * q];
        md = max(md, (int)d[7 * q]);
      }
      long long nn = (long long)v[0], sx = (long long)v[1], sy = (long long)v[2];
      long long sxx = (long long)v[3], syy = (long long)v[4], sxy = (long long)v[5];
      if (DT == DM_I16) {
        const long long c = 32768, c2 = 32768ll * 32768ll;
        const long long xx = sxx - 2 * c * sx + c2 * nn, yy = syy - 2 * c * sy + c2 * nn;
        const long long xy = sxy - c * (sx + sy) + c2 * nn;
        sx -= c * nn; sy -= c * nn; sxx = xx; syy = yy; sxy = xy;
      }
      int64_t* O = g.sums + (int64_t)b * DM_NSTAT;
      if (nn) atomic_add_i64(O + DM_S_N, nn);
      if (v[6]) atomic_add_i64(O + DM_S_ABS, (long long)v[6]);
      if (sx) atomic_add_i64(O + DM_S_X, sx);
      if (sy) atomic_add_i64(O + DM_S_Y, sy);
      if (sxx) atomic_add_i64(O + DM_S_XX, sxx);
      if (syy) atomic_add_i64(O + DM_S_YY, syy);
      if (sxy) atomic_add_i64(O + DM_S_XY, sxy);
      const long long sse = sxx + syy - 2 * sxy;
      if (sse) atomic_add_i64(O + DM_S_SSE, sse);
      if (md) atomic_max_i64(g.maxs + (int64_t)b * DM_NSTAT + DM_M_MAXERR, md);
    }
    asm volatile("bar.sync 2, %0;" ::"r"(kBandThreads));
    if (ts == 0 && sh_cube[4]) {
      int64_t* M = g.maxs;
      if (DT == DM_I16) {
        const int hi = sh_cube[1] - 32768, lo = sh_cube[2] - 32768;
        if (hi > 0) atomic_max_i64(M + DM_M_UMAX, hi);
        if (lo < 0) atomic_max_i64(M + DM_M_UNEGMIN, -lo);
      } else if (sh_cube[1] > 0) {
        atomic_max_i64(M + DM_M_UMAX, sh_cube[1]);
      }
      if (sh_cube[0] > 0) atomic_max_i64(M + DM_M_ABSXY, sh_cube[0]);
      if (sh_cube[3] & 0xF) atomic_max_i64(M + DM_M_LOW4, 1);
      if (sh_cube[3] & 0x3) atomic_max_i64(M + DM_M_LOW2, 1);
    }
  } else {
    // ------------------------------------------------------------------ pixel group
    // thread <-> pixel; the group is two halves of kTilePixels threads and half h takes the tiles
    // with (it & 1) == h, so every lane is busy in the float64 finish of its own pixel.
    // (Measured alternatives, same data: two lanes per pixel with an early stage release 243 us,
    // the same with the finish deferred by one tile 227 us, this mapping 206 us.)
    const int tg = tid - 32 - kBandThreads;          // 0..127
    const int half = tg / kTilePixels;
    const int tp = tg - half * kTilePixels;          // pixel of the tile
    double s_acos = 0.0, s_n = 0.0;
    int it = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
      if ((it & 1) != half) continue;
      const int s = it % kStages;
      const uint32_t ph = (uint32_t)((it / kStages) & 1);
      const int cnt = t < g.ntiles ? P : g.tail_pixels;
      mbar_wait(&full_bar[s], ph);
      uint32_t emax = 0;
      uint32_t xxl = 0, xxh = 0, yyl = 0, yyh = 0, xyl = 0, xyh = 0, sx = 0, sy = 0;
      if (tp < cnt) {
        const unsigned char* xs = smem + (size_t)s * stage_bytes + (size_t)tp * W * 4;
        const unsigned char* ys = xs + cube_bytes;
#pragma unroll 5
        for (int j = 0; j < (W >> 1); ++j) {
          const uint2 xv = *reinterpret_cast<const uint2*>(xs + 8 * j);
          const uint2 yv = *reinterpret_cast<const uint2*>(ys + 8 * j);
          const uint32_t xw[2] = {xv.x ^ OFS, xv.y ^ OFS}, yw[2] = {yv.x ^ OFS, yv.y ^ OFS};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t x = xw[k], y = yw[k];
            if (ERR) emax = vmaxu2(emax, vmaxu2(x, y) - vminu2(x, y));
            const uint32_t px = __byte_perm(x, 0, 0x3120), py = __byte_perm(y, 0, 0x3120);
            xxl = dp2a_lo(x, px, xxl); xxh = dp2a_hi(x, px, xxh);
            yyl = dp2a_lo(y, py, yyl); yyh = dp2a_hi(y, py, yyh);
            xyl = dp2a_lo(x, py, xyl); xyh = dp2a_hi(x, py, xyh);
            if (DT == DM_I16) { sx = dp2a_lo(x, 0x0101u, sx); sy = dp2a_lo(y, 0x0101u, sy); }
          }
        }
      }
      // the stage is no longer needed: release it before the per-pixel float64 work
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
      if (tp < cnt) {
        const int64_t p = t * (int64_t)P + tp;
        const uint8_t v = MASK ? g.plane[p] : (uint8_t)0xff;
        if (ERR) {
          int e = (v & DM_VALID_QUICKLOOK) ? hmax2(emax) : 0;           // quicklooks.py:134
          if (g.errmax) g.errmax[p] = (uint16_t)e;
          if (g.err8_g) {
            const uint8_t e8 = __ldg(g.lut_g + min(e, g.cap_g));
            g.err8_g[p] = e8;
            if (g.hist8_g) hist_add(h8g, e8);
          }
          if (g.err8_z) {
            const uint8_t e8 = __ldg(g.lut_z + min(e, g.cap_z));
            g.err8_z[p] = e8;
            if (g.hist8_z) hist_add(h8z, e8);
          }
        }
        if (g.want_sam && (v & DM_VALID_SPECTRAL)) {
          // lo + 256*hi: both halves < 2^32 and the sum < 2^53, so float64 holds it exactly
          double na2 = fma((double)xxh, 256.0, (double)xxl);
          double nr2 = fma((double)yyh, 256.0, (double)yyl);
          double dot = fma((double)xyh, 256.0, (double)xyl);
          if (DT == DM_I16) {
            const double c = 32768.0, c2B = 32768.0 * 32768.0 * (double)B, fx = (double)sx, fy = (double)sy;
            dot = dot - c * (fx + fy) + c2B;             // all terms exact integers below 2^53
            na2 = na2 - 2.0 * c * fx + c2B;
            nr2 = nr2 - 2.0 * c * fy + c2B;
          }
          const double na = __dadd_rn(__dsqrt_rn(na2), 1e-12);
          const double nr = __dadd_rn(__dsqrt_rn(nr2), 1e-12);
          double c = __ddiv_rn(dot, __dmul_rn(na, nr));
          c = fmin(1.0, fmax(-1.0, c));
          s_acos += acos_sam(c);
          s_n += 1.0;
        }
      }
    }
    asm volatile("bar.sync 1, %0;" ::"r"(kBandThreads + kPixelThreads));       // pairs with the band group
    // deterministic reduction of the float partials: warp shuffle tree, then warps in order
    s_acos = warp_sum_f64(s_acos); s_n = warp_sum_f64(s_n);
    const int pw = warp - 1 - kBandWarps;
    if (lane == 0) { red[0][pw] = s_acos; red[2][pw] = s_n; }
    asm volatile("bar.sync 3, %0;" ::"r"(kPixelThreads));
    if (tg == 0 && g.spec_out) {
      double t0 = 0, t2 = 0;
      for (int w = 0; w < kPixelWarps; ++w) { t0 += red[0][w]; t2 += red[2][w]; }
      g.spec_out[3 * blockIdx.x + 0] = t0; g.spec_out[3 * blockIdx.x + 1] = 0.0; g.spec_out[3 * blockIdx.x + 2] = t2;
    }
    if (blockIdx.x == 0 && g.spec_out)       // unused slots of the fixed-size partial array
      for (int i = 3 * gridDim.x + tg; i < 3 * kMaxSpecBlocks; i += kPixelThreads) g.spec_out[i] = 0.0;
    for (int i = tg; i < 256; i += kPixelThreads) {
      if (g.hist8_g && h8g[i]) atomic_add_i64(g.hist8_g + i, h8g[i]);
      if (g.hist8_z && h8z[i]) atomic_add_i64(g.hist8_z + i, h8z[i]);
    }
  }
}

}  // namespace

int launch_fused_bip(const dm_pair_t& p, const uint8_t* plane, int64_t* sums, int64_t* maxs,
                     uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                     const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                     double* spectral_out, cudaStream_t s) {
  if (!p.ref || !p.tst || !sums || !maxs) return fail(DM_EARG, "dm_fused_bip: null pointer");
  if (p.layout != DM_BIP) return fail(DM_EUNSUPPORTED, "dm_fused_bip: BIP cubes only");
  if (p.dtype != DM_U16 && p.dtype != DM_I16) return fail(DM_EUNSUPPORTED, "dm_fused_bip: 16-bit samples only");
  const int64_t B = p.bands;
  // dp2a lo/hi partials of one pixel stay below 2^32 up to 256 bands; 4 <= B, B % 4 == 0 for the
  // 8-byte column pairs; at most 192 band-group threads per pixel slot
  if (B < 4 || B % 4 || B > 256) return fail(DM_EUNSUPPORTED, "dm_fused_bip: bands must be a multiple of 4 in 4..256");
  if ((reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) & 15)
    return fail(DM_EUNSUPPORTED, "dm_fused_bip: cubes must be 16-byte aligned");
  if (err8_g && (!lut_g || cap_g < 0 || cap_g > 65535)) return fail(DM_EARG, "dm_fused_bip: bad global LUT");
  if (err8_z && (!lut_z || cap_z < 0 || cap_z > 65535)) return fail(DM_EARG, "dm_fused_bip: bad zoom LUT");
  if (want_sam && !spectral_out) return fail(DM_EARG, "dm_fused_bip: spectral_out is null");
  FusedArgs g;
  g.ref = p.ref; g.tst = p.tst; g.plane = plane; g.npix = p.rows * p.width; g.bands = (int)B;
  int P = kTilePixels;                                 // 64 pixels per tile when they fit a stage
  while ((int64_t)P * B * 4 > kStageBytesMax) P >>= 1;
  g.P = P;                                             // P*B*2 is a multiple of 16 (P even, B % 4 == 0)
  g.ntiles = g.npix / P;
  g.tail_pixels = (int)(g.npix - g.ntiles * P);
  g.sums = sums; g.maxs = maxs; g.errmax = errmax_out;
  g.lut_g = lut_g; g.cap_g = cap_g; g.err8_g = err8_g; g.hist8_g = err8_g ? hist8_g : nullptr;
  g.lut_z = lut_z; g.cap_z = cap_z; g.err8_z = err8_z; g.hist8_z = err8_z ? hist8_z : nullptr;
  g.want_sam = want_sam; g.spec_out = spectral_out;
  { const char* e = getenv("DM_FUSED_DEBUG"); g.debug = e ? atoi(e) : 0; }
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const int64_t total = g.ntiles + (g.tail_pixels ? 1 : 0);
  int64_t grid = sms < kMaxSpecBlocks ? sms : kMaxSpecBlocks;
  if (grid > total) grid = total;
  if (grid < 1) grid = 1;
  size_t smem = (size_t)kStages * 2 * P * B * 2;
  const size_t need_combine = (size_t)8 * 2 * kMaxBandThreads * 8;   // [8][NB][<= band threads] uint64 partials
  if (smem < need_combine) smem = need_combine;
  // C = 1 (twelve band warps, balanced schedulers) when one pixel needs at most 96 band threads per
  // slot, i.e. up to 192 bands; C = 2 above that
  const int cols = (B / 2 <= 96) ? 1 : 2;
#define DM_FUSED(DT, MASK, ERR)                                                                       \
  do {                                                                                                \
    if (cols == 1) {                                                                                  \
      auto k = fused_bip_kernel<DT, MASK, ERR, 1>;                                                    \
      DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
      k<<<(unsigned)grid, Shape<1>::kThreads, smem, s>>>(g);                                          \
    } else {                                                                                          \
      auto k = fused_bip_kernel<DT, MASK, ERR, 2>;                                                    \
      DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
      k<<<(unsigned)grid, Shape<2>::kThreads, smem, s>>>(g);                                          \
    }                                                                                                 \
  } while (0)
  const bool err = errmax_out || err8_g || err8_z;
  if (p.dtype == DM_U16) {
    if (plane) { if (err) DM_FUSED(DM_U16, true, true); else DM_FUSED(DM_U16, true, false); }
    else { if (err) DM_FUSED(DM_U16, false, true); else DM_FUSED(DM_U16, false, false); }
  } else {
    if (plane) { if (err) DM_FUSED(DM_I16, true, true); else DM_FUSED(DM_I16, true, false); }
    else { if (err) DM_FUSED(DM_I16, false, true); else DM_FUSED(DM_I16, false, false); }
  }
#undef DM_FUSED
  DM_LAUNCH_CHECK("fused_bip");
  return DM_OK;
}

}  // namespace dm
