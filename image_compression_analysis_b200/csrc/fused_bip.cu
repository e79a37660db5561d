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
//
// That is the run-time-geometry kernel (fused_bip_kernel, any band count = 0 mod 4 up to 256).  EnMAP's 180 bands take
// fused_ct_kernel further down: compile-time geometry, no producer warp, ldmatrix-paired band warps, two lanes per
// pixel in the pixel warps, launches chained by programmatic dependent launch -- and, in its SCAN build
// (dm_fused_bip_scan), the validity rule of run_codec.py:249-263 evaluated by the pixel warps inside the same pass.

#include <cstdlib>
#include <type_traits>

#include "dm_common.cuh"

#ifndef DM_POLL_NS
#define DM_POLL_NS 64
#endif
// Experiment switches (tools/probe_fused.py: DM_FUSED_DEBUG / DM_POLL_NS / DM_NO_PDL environment variables, A/B
// branches inside the kernels) exist only in a -DDM_DEBUG_HOOKS build (DM_DEBUG_HOOKS=1 python csrc/build.py).
// The shipping library reads no environment variable and the branches fold away at compile time.
#ifdef DM_DEBUG_HOOKS
#define DM_DBG(g) ((g).debug)
#define DM_POLL(g) ((g).poll_ns)
#else
#define DM_DBG(g) 0
#define DM_POLL(g) kPollNs
#endif

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
  int debug;                  // experiments only (DM_FUSED_DEBUG): 1 = producer skips the copies,
                              // 2 = band group skips its arithmetic, 4 = pixel group skips its arithmetic
  uint32_t zero;              // 0 (an operand the compiler cannot fold, see band_word)
  uint32_t poll_ns;           // barrier polling interval (DM_POLL_NS, default kPollNs)
  double* spec_acc;           // {sum arccos, -, n}, accumulated (ordered_block_sum3)
  void* ws;
  // in-kernel validity scan (fused_ct_kernel<..., SCAN = true>, dm_fused_bip_scan): the rule of dm_validity
  const uint8_t* valid_in;    // caller mask, may be null (16-byte aligned: its tile bytes ride along by bulk copy)
  uint8_t* plane_out;         // may be null: the validity plane as a by-product
  int64_t* counts;            // 3 x int64, accumulated (may be null)
  int ref_has, ref_nd, tst_has, tst_nd;
  int scan;                   // host side only: launch the SCAN build
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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, do not spin
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1]) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) { uint32_t r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t vminu2(uint32_t a, uint32_t b) { uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ int hmax2(uint32_t p) { return max((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmin2(uint32_t p) { return min((int)(p & 0xffffu), (int)(p >> 16)); }
__device__ __forceinline__ int hmax2s(uint32_t p) { return max((int)(short)(p & 0xffffu), (int)(short)(p >> 16)); }

// per-band packed accumulators (same scheme as stats.cu)
struct BandAccS {
  uint32_t sabs, sx, sy, xxl, xxh, yyl, yyh, xyl, xyh, maxd;
  uint32_t zero;      // always 0, but loaded from the kernel arguments (see band_word)
};
struct BandAcc : BandAccS {
  unsigned long long t_abs, t_x, t_y, t_xx, t_yy, t_xy;
  __device__ __forceinline__ void reset() {
    sabs = sx = sy = xxl = xxh = yyl = yyh = xyl = xyh = maxd = 0; zero = 0;
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
__device__ __forceinline__ void band_word(BandAccS& a, uint32_t x, uint32_t y, uint32_t& maxsel_u) {
  // three-input add with an operand the compiler cannot fold: IADD3 on the ALU pipe instead of an
  // IMAD.IADD on the FMA pipe, which the dp2a stream already saturates
  const uint32_t mx = vmaxu2(x, y), mn = vminu2(x, y), d = mx - mn + a.zero;
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
      if (DM_DBG(g) == 1) {
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
    if (tg < 32 && g.spec_acc) {
      double t0 = 0, t2 = 0;
      for (int w = 0; w < kPixelWarps; ++w) { t0 += red[0][w]; t2 += red[2][w]; }
      ordered_block_sum3(t0, 0.0, t2, g.ws, g.spec_acc);
    }
    for (int i = tg; i < 256; i += kPixelThreads) {
      if (g.hist8_g && h8g[i]) atomic_add_i64(g.hist8_g + i, h8g[i]);
      if (g.hist8_z && h8z[i]) atomic_add_i64(g.hist8_z + i, h8z[i]);
    }
  }
}

// =================================================================================================
// Compile-time geometry variant (EnMAP: 180 bands).  Same band / pixel roles as the kernel above, but
// there is no producer warp: the consumer warp that finishes a tile LAST (shared-memory arrival
// counter) re-arms the stage's full barrier and issues the bulk copies of the tile four ahead.  Every
// shared-memory offset is an immediate, a tile is always full (the launcher hands a partial last tile
// to the generic kernel), band threads own FOUR bands (one LDS.64 per pixel and cube) and keep only
// the 32-bit partials in registers: the rare spill (every 128 dp2a steps) goes to 64-bit
// shared-memory totals, which also makes the end-of-kernel combine a plain per-band read.
constexpr int kPixelWarpsCT = 8;
constexpr int kPixelThreadsCT = kPixelWarpsCT * 32;
template <int BANDS, int MPW_ = 4> struct Geo {
  static constexpr int MPW = MPW_;                      // ldmatrix matrices (16-byte chunks) per band warp: 4 or 2
  static constexpr int W = BANDS / 2;                   // 32-bit words per pixel
  static constexpr int PIXB = BANDS * 2;                // bytes per pixel
  static constexpr int P = kTilePixels;                 // 64
  static constexpr int CUBE = P * PIXB;                 // one cube's share of a stage
  static constexpr int STAGE = 2 * CUBE;
  static constexpr int PITCH = STAGE + 128;             // a stage in the ring: both cubes' tiles + the tile's mask bytes
  static constexpr int CHUNKS = BANDS / 4;              // 16-byte chunks per pixel pair
  static constexpr int BAND_WARPS = (CHUNKS + MPW - 1) / MPW, BAND_THREADS = BAND_WARPS * 32;   // 12 / 23 for 180 bands
  static constexpr int ROWBLOCKS = P / 16;              // ldmatrix row blocks (8 pixel pairs) per tile
  static constexpr int THREADS = BAND_THREADS + kPixelThreadsCT;   // 640: 20 warps x 96 registers
  static constexpr int CONSUMERS = BAND_WARPS + kPixelWarpsCT / 2;    // warps that read one tile
  static constexpr int NQ = 6;                          // abs, x, y, xx, yy, xy
  static constexpr size_t TOT_BYTES = (size_t)(NQ + 1) * BANDS * 8 + (size_t)BANDS * 4;   // totals, N, max|d|
  static constexpr size_t SMEM = (size_t)kStages * PITCH + TOT_BYTES;
  static_assert(BANDS % 4 == 0 && BANDS <= 256, "dp2a lo/hi pixel partials need B <= 256");
  static_assert(SMEM <= 227 * 1024, "ring does not fit");
  static_assert((PIXB / 8) % 2 == 1, "conflict-free LDS.64 / ldmatrix walks need an odd pixel pitch in 8-byte units");
  static_assert(MPW == 2 || MPW == 4, "ldmatrix x2 / x4");
  static_assert(THREADS <= 1024, "too many band warps for one CTA");
};

// signed dp2a for int16 samples: a = two s16, selector bytes unsigned (low bytes) / signed (high bytes)
__device__ __forceinline__ uint32_t dp2a_lo_su(uint32_t a, uint32_t b, uint32_t c) { int r; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"((int)c)); return (uint32_t)r; }
__device__ __forceinline__ uint32_t dp2a_hi_ss(uint32_t a, uint32_t b, uint32_t c) { int r; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"((int)c)); return (uint32_t)r; }
__device__ __forceinline__ uint32_t vadd2_wrap(uint32_t a, uint32_t b) { return __vadd2(a, b); }   // per-half add, VIADD.16x2

// barrier helpers on raw shared-memory addresses (computed once per thread, not per tile)
constexpr unsigned kPollNs = DM_POLL_NS;
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity, uint32_t poll_ns) {
  // One try_wait (the hardware suspends the warp for a while when the phase is not complete), then
  // poll at a coarse interval: the bulk copy signals the barrier once per 2 KB granule, and a waiter
  // that wakes on every one of those spends issue slots that the working warps need.
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "WAIT_%=:\n"
      "nanosleep.u32 %2;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity), "r"(poll_ns) : "memory");
}
// arrive and tell whether this was the LAST pending arrival of the phase (exactly one arriver sees it)
__device__ __forceinline__ bool mbar_arrive_is_last_a(uint32_t bar) {
  uint64_t st;
  uint32_t pending;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(bar) : "memory");
  asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(pending) : "l"(st));
  return pending == 1u;
}

// SCAN (implies MASK): the validity plane is not an input but computed in the kernel.  The pixel group that
// serves a tile first sweeps it for the nodata rules of dm_validity (OR / MIN of the spectrum XORed with the
// packed nodata values: ~5 instructions per 8 bytes, a few hundred nanoseconds per tile), writes the tile's 64
// validity bytes where the bulk-copied plane of the MASK variant would sit and arrives on the stage's `mask`
// barrier; the band warps wait for THAT barrier instead of the full barrier, so they trail the data by the
// scan only, not by the pixel group's whole SAM sweep, and the 4-stage ring keeps its slack.  One read of
// the pair instead of two (validity pre-pass + masked kernel): run_codec.py:249-263 folded into :268-285.
template <class T> __device__ __forceinline__ constexpr int stage_of(T) { return T::value; }     // std::integral_constant
__device__ __forceinline__ constexpr int stage_of(int s) { return s; }

template <int BANDS, int DT, bool MASK, bool ERR, int MPW, bool SCAN = false>
__global__ void __launch_bounds__(Geo<BANDS, MPW>::THREADS, 1)
fused_ct_kernel(FusedArgs g) {
  using G = Geo<BANDS, MPW>;
  static_assert(!SCAN || MASK, "the in-kernel scan feeds the masked arithmetic");
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], mask_bar[kStages];
  __shared__ unsigned h8g[256], h8z[256];
  __shared__ double red[3][kPixelWarpsCT];
  __shared__ int sh_cube[8];
  unsigned long long* tot = reinterpret_cast<unsigned long long*>(smem + (size_t)kStages * G::PITCH);   // [NQ][BANDS]
  unsigned long long* tot_n = tot + G::NQ * BANDS;                                                      // [BANDS]
  unsigned* tot_maxd = reinterpret_cast<unsigned*>(tot_n + BANDS);                                      // [BANDS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t OFS = DT == DM_I16 ? 0x80008000u : 0u;
  constexpr bool TRACK = MASK || DT == DM_I16;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], G::CONSUMERS);
      mbar_init(&mask_bar[s], kPixelWarpsCT / 2);       // SCAN: the four pixel warps that scan a tile
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 256) { h8g[tid] = 0; h8z[tid] = 0; }
  if (tid < 8) sh_cube[tid] = tid == 2 ? 0x7fffffff : (tid == 0 ? (int)0x80000000 : 0);
  for (int i = tid; i < (G::NQ + 1) * BANDS; i += G::THREADS) tot[i] = 0ull;
  for (int i = tid; i < BANDS; i += G::THREADS) tot_maxd[i] = 0u;
  __syncthreads();

  // tiles of this CTA: local index it = 0 .. my_tiles-1 is global tile blockIdx.x + it*gridDim.x and
  // lives in stage it % kStages (full tiles only; the launcher hands a partial tile to the generic kernel)
  const int my_tiles = (int64_t)blockIdx.x < g.ntiles ? (int)((g.ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  // no per-pixel output requested (stats only): the pixel warps just hand their stages back
  const int dbg = DM_DBG(g) | ((ERR || g.want_sam) ? 0 : 4);
  const uint32_t poll_ns = DM_POLL(g);
  const uint32_t ring = smem_u32(smem);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), mask0 = smem_u32(&mask_bar[0]);
  const bool scan_vin = SCAN && g.valid_in != nullptr;
  // global writes inside the tile loop (error planes, the validity plane): no early start of the next launch
  const bool loop_writes = ERR || (SCAN && g.plane_out != nullptr);

  auto issue_tile = [&](int it) {
    if (it >= my_tiles) return;
    const int s = it & (kStages - 1);
    uint64_t* fb = &full_bar[s];
    if (dbg & 1) { mbar_arrive(fb); return; }
    const int64_t off = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * (int64_t)G::CUBE;
    unsigned char* dst = smem + (size_t)s * G::PITCH;
    mbar_expect_tx(fb, G::STAGE + (SCAN ? (scan_vin ? G::P : 0) : (MASK ? G::P : 0)));
    bulk_g2s(dst, static_cast<const char*>(g.ref) + off, G::CUBE, fb);
    bulk_g2s(dst + G::CUBE, static_cast<const char*>(g.tst) + off, G::CUBE, fb);
    // the tile's validity bytes ride along, so that no consumer touches global memory in its loop
    // (SCAN: the caller's mask bytes instead, behind the 64 bytes the pixel group will write)
    if (SCAN) { if (scan_vin) bulk_g2s(dst + G::STAGE + G::P, g.valid_in + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * G::P, G::P, fb); }
    else if (MASK) bulk_g2s(dst + G::STAGE, g.plane + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * G::P, G::P, fb);
  };
  // Called by lane 0 of a consumer warp when the warp has finished reading tile `it`.  There is no
  // producer warp: every consumer warp arrives on the stage's empty barrier, and the one whose arrival
  // completes the phase (the last of the CONSUMERS warps) refills the stage with tile it + kStages.
  // The (immediately successful) wait on the completed phase gives that thread acquire ordering on
  // the other warps' shared-memory reads before the asynchronous proxy overwrites the stage.
  auto release_tile = [&](int it) {
    const uint32_t eb = empty0 + 8u * (uint32_t)(it & (kStages - 1));
    if (mbar_arrive_is_last_a(eb)) {
      mbar_wait_a(eb, (uint32_t)((it / kStages) & 1), poll_ns);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_tile(it + kStages);
    }
  };
  if (tid == 0) {
    for (int it = 0; it < kStages; ++it) issue_tile(it);
  }
  // Programmatic dependent launch: a following launch of this kernel (run_ct sets the attribute while the
  // caller has launch chaining on) may start its CTAs as ours exit, instead of after the whole grid -- the
  // tail of this launch (CTAs with one tile fewer, the last block's ordered reduction) then overlaps the
  // next pair's first tiles.  The next launch only READS the cubes before its own griddepcontrol.wait
  // (below, ahead of every global write), so nothing it does early can race with what is still running
  // here.  Variants that write per-pixel planes do not trigger early: two launches may be given the same
  // planes.
  if (!loop_writes) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp < G::BAND_WARPS) {
    // ------------------------------------------------------------------ band group
    // A pixel PAIR (even pixel, odd pixel) is 2*PIXB bytes = CHUNKS 16-byte chunks, and 16-bit column c
    // of chunk k is always the same (pixel parity, band): sample column sc = 8k + c, parity sc / BANDS,
    // band sc % BANDS.  ldmatrix.trans over 8 pixel-pair rows x one chunk hands thread (c = lane/4,
    // r = lane%4) the word [row 2r+1 | row 2r] of column c -- two pixels of the SAME band, which is the
    // operand pairing dp2a needs, with no PRMT regrouping.  Warp w owns chunks 4w .. 4w+3 (x4: four
    // chunks per instruction), i.e. four fixed sample columns per thread, accumulated in registers.
    // Row addresses are 2*PIXB apart = 45 chunks, odd, so the eight rows of a matrix hit eight
    // different 16-byte bank groups (conflict free).
    const int ts = tid;
    const int c = lane >> 2, r = lane & 3;
    const int nmat = (G::CHUNKS - MPW * warp) < MPW ? (G::CHUNKS - MPW * warp) : MPW;     // >= 1 for every band warp
    const int sc0 = 8 * MPW * warp + c;                // sample column of matrix j: sc0 + 8j
    auto par_of = [&](int j) { return sc0 + 8 * j >= BANDS; };
    auto band_of = [&](int j) { const int sc = sc0 + 8 * j; return sc >= BANDS ? sc - BANDS : sc; };
    const int mlane = (lane >> 3) & (MPW - 1);
    const int mchunk = (MPW * warp + mlane) < G::CHUNKS ? (MPW * warp + mlane) : (G::CHUNKS - 1);
    const uint32_t ld_off = ring + (uint32_t)(lane & 7) * (2u * G::PIXB) + (uint32_t)mchunk * 16u;
    BandAccS a[MPW];
#pragma unroll
    for (int j = 0; j < MPW; ++j) { a[j].sabs = a[j].sx = a[j].sy = a[j].xxl = a[j].xxh = a[j].yyl = a[j].yyh = a[j].xyl = a[j].xyh = a[j].maxd = 0; a[j].zero = g.zero; }
    uint32_t maxsel_u = 0, minsel_m1 = 0xffffffffu, umax = 0, umin = 0xffffffffu, orbits = 0, ymax = 0;
    uint32_t n0 = 0, n1 = 0;                           // MASK: selected even / odd pixels of this thread's rows

    auto spill = [&]() {
#pragma unroll
      for (int j = 0; j < MPW; ++j) {
        if (j < nmat) {
          const int b = band_of(j);
          atomicAdd(tot + 0 * BANDS + b, (unsigned long long)a[j].sabs);
          atomicAdd(tot + 1 * BANDS + b, (unsigned long long)a[j].sx);
          atomicAdd(tot + 2 * BANDS + b, (unsigned long long)a[j].sy);
          atomicAdd(tot + 3 * BANDS + b, (unsigned long long)a[j].xxl + ((unsigned long long)a[j].xxh << 8));
          atomicAdd(tot + 4 * BANDS + b, (unsigned long long)a[j].yyl + ((unsigned long long)a[j].yyh << 8));
          atomicAdd(tot + 5 * BANDS + b, (unsigned long long)a[j].xyl + ((unsigned long long)a[j].xyh << 8));
        }
        a[j].sabs = a[j].sx = a[j].sy = a[j].xxl = a[j].xxh = a[j].yyl = a[j].yyh = a[j].xyl = a[j].xyh = 0;
      }
    };

    // NM = matrices (chunks) this warp really owns: 4, or fewer in the last band warp; a compile-time
    // count keeps the row block free of branches so that the four columns interleave.  Tiles run in
    // epochs of 128 / ROWBLOCKS: a 32-bit partial takes one dp2a per row block and holds 128 of them
    // (2 * 65535 * 255 each), so the spill to the 64-bit shared totals sits outside the tile loop.
    auto run = [&](auto nm_tag) {
      constexpr int NM = decltype(nm_tag)::value;
      constexpr int EPOCH = 128 / G::ROWBLOCKS;
      static_assert(EPOCH % kStages == 0, "an epoch starts in stage 0");
      // one tile in stage S
      auto tile = [&](auto s_tag, int it, uint32_t par) {
        // S: std::integral_constant (unrolled trips: offsets and barrier addresses fold into immediates) or a plain int
        const int S = stage_of(s_tag);
        // SCAN: the mask barrier completes after the pixel warps have seen the full barrier complete and released
        // their writes, so it orders the tile's data as well (one wait per tile instead of two: -3 us)
        mbar_wait_a((SCAN ? mask0 : full0) + 8u * (uint32_t)S, par, poll_ns);
        if (!(dbg & 2)) {
          const uint32_t xs = ld_off + (uint32_t)S * G::PITCH;
          const unsigned char* pl = smem + (size_t)S * G::PITCH + G::STAGE;      // MASK: the tile's validity bytes
#pragma unroll
          for (int rb = 0; rb < G::ROWBLOCKS; ++rb) {
            uint32_t xr[MPW], yr[MPW];
            if constexpr (MPW == 4) {
              ldsm_x4_trans(xr, xs + rb * (16 * G::PIXB));
              ldsm_x4_trans(yr, xs + rb * (16 * G::PIXB) + G::CUBE);
            } else {
              ldsm_x2_trans(xr, xs + rb * (16 * G::PIXB));
              ldsm_x2_trans(yr, xs + rb * (16 * G::PIXB) + G::CUBE);
            }
            uint32_t m0 = 0xffffffffu, m1 = 0xffffffffu;
            if (MASK) {
              // pixels 16rb + 4r .. +3 of the tile: even/odd pixel of row 2r, even/odd pixel of row 2r+1
              const uint32_t vb = *reinterpret_cast<const uint32_t*>(pl + 16 * rb + 4 * r);
              m0 = ((vb & DM_VALID_METRICS) ? 0xffffu : 0u) | ((vb & (DM_VALID_METRICS << 16)) ? 0xffff0000u : 0u);
              m1 = ((vb & (DM_VALID_METRICS << 8)) ? 0xffffu : 0u) | ((vb & (DM_VALID_METRICS << 24)) ? 0xffff0000u : 0u);
              n0 += (m0 & 1u) + (m0 >> 31);
              n1 += (m1 & 1u) + (m1 >> 31);
            }
            // data-range scan on the raw reference words (unmasked); a short last warp re-reads its
            // last chunk in the unused matrices, which changes neither an OR nor a maximum
#pragma unroll
            for (int j = 0; j < MPW; j += 2) orbits |= xr[j] | xr[j + 1];
#pragma unroll
            for (int j = 0; j < MPW; ++j) { xr[j] ^= OFS; yr[j] ^= OFS; }
#pragma unroll
            for (int j = 0; j < MPW; j += 2) {
              umax = __vimax3_u16x2(umax, xr[j], xr[j + 1]);
              if (DT == DM_I16) umin = __vimin3_u16x2(umin, xr[j], xr[j + 1]);
              if (!TRACK) ymax = __vimax3_u16x2(ymax, yr[j], yr[j + 1]);
            }
#pragma unroll
            for (int j = 0; j < NM; ++j) {
              uint32_t x = xr[j], y = yr[j];
              if (MASK) { const uint32_t m = par_of(j) ? m1 : m0; x &= m; y &= m; }
              band_word<true, TRACK>(a[j], x, y, maxsel_u);
              if (DT == DM_I16) {
                // max |v| with np.abs semantics (abs(-32768) wraps and never wins) from offset-binary
                // extremes: maxsel_u tracks max u (in band_word); here min over u > 0 as min of u-1
                // with per-half wrap (u = 0, i.e. -32768 or a masked sample, becomes 0xffff and drops out)
                minsel_m1 = vminu2(minsel_m1, vadd2_wrap(vminu2(x, y), 0xffffffffu));
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
          // release_tile(it) with the stage known: arrive on the stage's empty barrier; the arrival that completes the
          // phase refills the stage with tile it + kStages
          const uint32_t eb = empty0 + 8u * (uint32_t)S;
          if (mbar_arrive_is_last_a(eb)) {
            mbar_wait_a(eb, par, poll_ns);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_tile(it + kStages);
          }
        }
      };
      for (int it0 = 0; it0 < my_tiles; it0 += EPOCH) {
        const int it1 = it0 + EPOCH < my_tiles ? it0 + EPOCH : my_tiles;
        int it = it0;
        [[maybe_unused]] uint32_t par = (uint32_t)(it0 / kStages) & 1u;
        // The unrolled form pays where the tile body is small (the plain 23-band-warp build: 132.9 -> 131.4 us per
        // step); with the masked / four-matrix bodies four copies of it overflow the instruction cache (the int16 scan
        // build went from 186 to 249 us), so those keep the rolled loop.
        if constexpr (MPW == 2 && !MASK) {
          for (; it + kStages <= it1; it += kStages, par ^= 1u) {
            tile(std::integral_constant<int, 0>(), it, par);
            tile(std::integral_constant<int, 1>(), it + 1, par);
            tile(std::integral_constant<int, 2>(), it + 2, par);
            tile(std::integral_constant<int, 3>(), it + 3, par);
          }
        }
        for (; it < it1; ++it) tile(it & (kStages - 1), it, (uint32_t)(it / kStages) & 1u);
        spill();
      }
    };
    // (only two counts occur: MPW, and what is left for the last band warp)
    constexpr int LAST_NM = G::CHUNKS - MPW * (G::BAND_WARPS - 1);
    if (nmat == MPW) run(std::integral_constant<int, MPW>());
    else run(std::integral_constant<int, LAST_NM>());

    // ---- flush: per-band maxima and counts -> shared, then one thread per band -> global
    // (the preceding launch must have finished before anything global is written: no-op unless this
    // launch was started early through programmatic dependent launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (my_tiles > 0 && !(dbg & 2)) {
#pragma unroll
      for (int j = 0; j < MPW; ++j) {
        if (j < nmat) {
          atomicMax(tot_maxd + band_of(j), (unsigned)hmax2(a[j].maxd));
          if (MASK) atomicAdd(tot_n + band_of(j), (unsigned long long)(par_of(j) ? n1 : n0));
        }
      }
      if (!TRACK) maxsel_u = vmaxu2(umax, ymax);       // unmasked uint16: max over everything read
      if (DT == DM_I16) {
        const int hi = hmax2(maxsel_u) - 32768, lo = 32768 - (hmin2(minsel_m1) + 1);   // lo = -32768 when no u > 0
        atomicMax(sh_cube + 0, max(hi, lo));
      } else {
        atomicMax(sh_cube + 0, hmax2(maxsel_u));
      }
      atomicMax(sh_cube + 1, hmax2(umax));
      atomicMin(sh_cube + 2, hmin2(umin));
      atomicOr(reinterpret_cast<unsigned*>(sh_cube + 3), (orbits | (orbits >> 16)) & 0xffffu);
      sh_cube[4] = 1;
    }
    asm volatile("bar.sync 2, %0;" ::"r"(G::BAND_THREADS));
    const long long n_cta = (dbg & 2) ? 0 : (long long)my_tiles * G::P;      // pixels this CTA has read
    for (int b = ts; b < BANDS; b += G::BAND_THREADS) {
      long long nn = MASK ? (long long)tot_n[b] : n_cta;
      long long sab = (long long)tot[0 * BANDS + b], sx = (long long)tot[1 * BANDS + b], sy = (long long)tot[2 * BANDS + b];
      long long sxx = (long long)tot[3 * BANDS + b], syy = (long long)tot[4 * BANDS + b], sxy = (long long)tot[5 * BANDS + b];
      const int md = (int)tot_maxd[b];
      if (DT == DM_I16) {
        const long long c = 32768, c2 = 32768ll * 32768ll;
        const long long xx = sxx - 2 * c * sx + c2 * nn, yy = syy - 2 * c * sy + c2 * nn;
        const long long xy = sxy - c * (sx + sy) + c2 * nn;
        sx -= c * nn; sy -= c * nn; sxx = xx; syy = yy; sxy = xy;
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
    // Two 4-warp groups take alternate tiles; inside a group TWO lanes share a pixel (lane l and
    // l+16 walk the two halves of its spectrum), which halves the time a stage is held.  The halves
    // are combined with one shuffle per partial; the float64 finish is batched over two visits
    // (lanes 0-15 keep the pixels of the even visit, lanes 16-31 those of the odd one) so that all
    // 32 lanes of the warp work in it.
    if (loop_writes) asm volatile("griddepcontrol.wait;" ::: "memory");   // planes are written in the tile loop
    const int tg = tid - G::BAND_THREADS;             // 0..255
    const int grp = tg >> 7;                          // tile parity this group serves
    const int wq = (tg >> 5) & 3;                     // warp of the group: pixels 16wq .. 16wq+15
    const int hl = lane >> 4;                         // half of the spectrum this lane walks
    const int tp = 16 * wq + (lane & 15);
    constexpr int U = G::W / 2;                       // 8-byte units per pixel
    constexpr int H0 = (U + 1) / 2, H1 = U - H0;      // units of half 0 / half 1
    const unsigned char* lane_base = smem + (size_t)tp * G::PIXB + (hl ? H0 * 8 : 0);
    double s_acos = 0.0, s_n = 0.0;
    uint32_t k_xxl = 0, k_xxh = 0, k_yyl = 0, k_yyh = 0, k_xyl = 0, k_xyh = 0, k_e = 0;
    int k_it = 0;
    uint32_t k_v = 0xff;
    const uint32_t ndr = ((uint32_t)g.ref_nd & 0xffffu) * 0x10001u, ndt = ((uint32_t)g.tst_nd & 0xffffu) * 0x10001u;
    int c0 = 0, c1 = 0, c2 = 0;                       // SCAN: pixels with each validity bit (lanes 0-15)

    auto finish = [&]() {
      const int64_t p = ((int64_t)blockIdx.x + (int64_t)k_it * gridDim.x) * G::P + tp;
      const uint32_t v = k_v;
      if (ERR) {
        int e = (v & DM_VALID_QUICKLOOK) ? hmax2(k_e) : 0;             // quicklooks.py:134
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
        double na2, nr2, dot;
        if (DT == DM_I16) {                            // signed 32-bit partials
          na2 = fma((double)(int)k_xxh, 256.0, (double)(int)k_xxl);
          nr2 = fma((double)(int)k_yyh, 256.0, (double)(int)k_yyl);
          dot = fma((double)(int)k_xyh, 256.0, (double)(int)k_xyl);
        } else {
          na2 = fma((double)k_xxh, 256.0, (double)k_xxl);
          nr2 = fma((double)k_yyh, 256.0, (double)k_yyl);
          dot = fma((double)k_xyh, 256.0, (double)k_xyl);
        }
        const double na = __dadd_rn(__dsqrt_rn(na2), 1e-12);
        const double nr = __dadd_rn(__dsqrt_rn(nr2), 1e-12);
        double c = __ddiv_rn(dot, __dmul_rn(na, nr));
        c = fmin(1.0, fmax(-1.0, c));
        s_acos += acos_sam(c);
        s_n += 1.0;
      }
    };

    // SCAN: validity of tile `it` (the rule of validity_ct_kernel below).  Waits for the tile, publishes its 64
    // validity bytes in the stage, arrives on the stage's mask barrier and returns this lane's pixel's byte.
    // Lean first sweep: only the per-halfword MINIMUM of the spectrum XORed with nodata (zero <=> some band equals
    // nodata <=> the pixel leaves compute_metrics); when nodata is the type's lowest value in both files (EnMAP's
    // int16 -32768, Sentinel-2's 0) the minimum of the raw samples does it without the XOR.  The OR over the
    // spectrum (dataset_mask(): some band differs from nodata) can only be zero where that minimum is zero, so a
    // second sweep computes it just for warps that met a nodata sample (warp-uniform branch).
    const bool nd_lowest = SCAN && (!g.ref_has || (g.ref_nd & 0xffff) == (DT == DM_I16 ? 0x8000 : 0))
                                && (!g.tst_has || (g.tst_nd & 0xffff) == (DT == DM_I16 ? 0x8000 : 0));
    auto scan_tile = [&](int it) -> uint32_t {
      const int s = it & (kStages - 1);
      mbar_wait_a(full0 + 8u * (uint32_t)s, (uint32_t)((it / kStages) & 1), poll_ns);
      const unsigned char* xs = lane_base + (size_t)s * G::PITCH;
      uint32_t r_min, t_min;
      if (nd_lowest) {
        auto lo = [&](const unsigned char* base) -> uint32_t {
          uint32_t m = DT == DM_I16 ? 0x7fff7fffu : 0xffffffffu;
          auto unit = [&](int j) {
            const uint2 w = *reinterpret_cast<const uint2*>(base + 8 * j);
            m = DT == DM_I16 ? __vimin3_s16x2(m, w.x, w.y) : __vimin3_u16x2(m, w.x, w.y);
          };
#pragma unroll 11
          for (int j = 0; j < H1; ++j) unit(j);
          if (H0 > H1 && hl == 0) unit(H1);
          return m ^ OFS;                              // offset binary: zero <=> the lowest value was met
        };
        r_min = g.ref_has ? lo(xs) : 0xffffffffu;
        t_min = g.tst_has ? lo(xs + G::CUBE) : 0xffffffffu;
      } else {
        auto lo = [&](const unsigned char* base, uint32_t nd2) -> uint32_t {
          uint32_t m = 0xffffffffu;
          auto unit = [&](int j) {
            const uint2 w = *reinterpret_cast<const uint2*>(base + 8 * j);
            m = __vimin3_u16x2(m, w.x ^ nd2, w.y ^ nd2);
          };
#pragma unroll 11
          for (int j = 0; j < H1; ++j) unit(j);
          if (H0 > H1 && hl == 0) unit(H1);
          return m;
        };
        r_min = g.ref_has ? lo(xs, ndr) : 0xffffffffu;
        t_min = g.tst_has ? lo(xs + G::CUBE, ndt) : 0xffffffffu;
      }
      r_min = vminu2(r_min, __shfl_xor_sync(0xffffffffu, r_min, 16));
      t_min = vminu2(t_min, __shfl_xor_sync(0xffffffffu, t_min, 16));
      const bool rl = hmin2(r_min) != 0, tl = hmin2(t_min) != 0;   // no band equals nodata
      bool ds = true, band1 = true;
      if (__any_sync(0xffffffffu, !(rl && tl))) {
        // "some band differs from nodata" for the cube(s) in which this warp met a nodata sample: OR of the
        // XORed words, or -- nodata being the lowest value -- the maximum of the raw samples (one instruction
        // per 8 bytes instead of two)
        auto any = [&](const unsigned char* base, uint32_t nd2) -> bool {
          uint32_t o = nd_lowest ? (DT == DM_I16 ? 0x80008000u : 0u) : 0u;
          auto unit = [&](int j) {
            const uint2 w = *reinterpret_cast<const uint2*>(base + 8 * j);
            if (nd_lowest) o = DT == DM_I16 ? __vimax3_s16x2(o, w.x, w.y) : __vimax3_u16x2(o, w.x, w.y);
            else o |= (w.x ^ nd2) | (w.y ^ nd2);
          };
#pragma unroll 11
          for (int j = 0; j < H1; ++j) unit(j);
          if (H0 > H1 && hl == 0) unit(H1);
          if (nd_lowest) o ^= OFS;                     // nonzero <=> some sample above the lowest value
          o |= __shfl_xor_sync(0xffffffffu, o, 16);
          return o != 0;
        };
        bool r_any = true, t_any = true, r_b1 = true, t_b1 = true;
        if (__any_sync(0xffffffffu, !rl)) { r_any = any(xs, ndr); r_b1 = ((*reinterpret_cast<const uint32_t*>(xs) ^ ndr) & 0xffffu) != 0; }
        if (__any_sync(0xffffffffu, !tl)) { t_any = any(xs + G::CUBE, ndt); t_b1 = ((*reinterpret_cast<const uint32_t*>(xs + G::CUBE) ^ ndt) & 0xffffu) != 0; }
        // band 1 lives in half 0: lanes 16-31 take their partner's verdict
        const unsigned b1 = __ballot_sync(0xffffffffu, r_b1 && t_b1);
        band1 = (b1 >> (lane & 15)) & 1u;
        ds = r_any && t_any;
      }
      unsigned char* mk = smem + (size_t)s * G::PITCH + G::STAGE;
      const bool vin = scan_vin ? mk[G::P + tp] != 0 : true;
      uint32_t v = 0;
      if (ds && rl && tl && vin) v |= DM_VALID_METRICS;
      if (ds && band1) v |= DM_VALID_QUICKLOOK;
      if (scan_vin ? vin : ds) v |= DM_VALID_SPECTRAL;
      if (hl == 0) {
        mk[tp] = (unsigned char)v;
        if (g.plane_out) g.plane_out[((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * G::P + tp] = (uint8_t)v;
        c0 += (v & DM_VALID_METRICS) ? 1 : 0;
        c1 += (v & DM_VALID_QUICKLOOK) ? 1 : 0;
        c2 += (v & DM_VALID_SPECTRAL) ? 1 : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mask_bar[s]);
      return v;
    };

    // SCAN: a group scans its own tile when it gets to it.  Scanning the group's NEXT tile ahead of time (before the
    // sweep, half way through it, or between sweep and float64 finish, blocking or only when the tile has landed)
    // was measured slower every time (216-244 us against 207 us per EnMAP cube): it delays the group's own sweep,
    // and the stage that sweep holds is what the ring is short of.
    int visit = 0;
    for (int it = grp; it < my_tiles; it += 2, ++visit) {
      const int s = it & (kStages - 1);
      if (!SCAN) mbar_wait_a(full0 + 8u * (uint32_t)s, (uint32_t)((it / kStages) & 1), poll_ns);   // SCAN: scan_tile(it) waited
      uint32_t emax = 0, vcur = 0xffu;
      if (SCAN) vcur = scan_tile(it);
      uint32_t xxl = 0, xxh = 0, yyl = 0, yyh = 0, xyl = 0, xyh = 0;
      if (!(dbg & 4)) {
        const unsigned char* xs = lane_base + (size_t)s * G::PITCH;
        if (MASK && !SCAN) vcur = smem[(size_t)s * G::PITCH + G::STAGE + tp];
        auto unit = [&](int j) {
          const uint2 xv = *reinterpret_cast<const uint2*>(xs + 8 * j);
          const uint2 yv = *reinterpret_cast<const uint2*>(xs + G::CUBE + 8 * j);
          const uint32_t xw[2] = {xv.x, xv.y}, yw[2] = {yv.x, yv.y};
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint32_t x = xw[k], y = yw[k];
            if (ERR) { const uint32_t ux = x ^ OFS, uy = y ^ OFS; emax = vmaxu2(emax, vmaxu2(ux, uy) - vminu2(ux, uy)); }
            const uint32_t px = __byte_perm(x, 0, 0x3120), py = __byte_perm(y, 0, 0x3120);
            if (DT == DM_I16) {
              // signed samples directly: s16 x (unsigned low byte) and s16 x (signed high byte); a pixel's
              // partials stay below 2^31 up to 256 bands (2 * 32768 * 255 per dp2a)
              xxl = dp2a_lo_su(x, px, xxl); xxh = dp2a_hi_ss(x, px, xxh);
              yyl = dp2a_lo_su(y, py, yyl); yyh = dp2a_hi_ss(y, py, yyh);
              xyl = dp2a_lo_su(x, py, xyl); xyh = dp2a_hi_ss(x, py, xyh);
            } else {
              xxl = dp2a_lo(x, px, xxl); xxh = dp2a_hi(x, px, xxh);
              yyl = dp2a_lo(y, py, yyl); yyh = dp2a_hi(y, py, yyh);
              xyl = dp2a_lo(x, py, xyl); xyh = dp2a_hi(x, py, xyh);
            }
          }
        };
#pragma unroll 11
        for (int j = 0; j < H1; ++j) unit(j);
        if (H0 > H1 && hl == 0) unit(H1);              // the odd unit belongs to half 0
      }
      // the stage is no longer needed: release it before the shuffles and the float64 work
      __syncwarp();
      if (lane == 0) release_tile(it);
      if (!(dbg & 4)) {
        xxl += __shfl_xor_sync(0xffffffffu, xxl, 16); xxh += __shfl_xor_sync(0xffffffffu, xxh, 16);
        yyl += __shfl_xor_sync(0xffffffffu, yyl, 16); yyh += __shfl_xor_sync(0xffffffffu, yyh, 16);
        xyl += __shfl_xor_sync(0xffffffffu, xyl, 16); xyh += __shfl_xor_sync(0xffffffffu, xyh, 16);
        if (ERR) emax = vmaxu2(emax, __shfl_xor_sync(0xffffffffu, emax, 16));
        if (hl == (visit & 1)) {
          k_xxl = xxl; k_xxh = xxh; k_yyl = yyl; k_yyh = yyh; k_xyl = xyl; k_xyh = xyh; k_e = emax;
          k_it = it; k_v = vcur;
        }
        if (visit & 1) finish();
      }
    }
    if ((visit & 1) && hl == 0 && !(dbg & 4)) finish();      // pixels of an unpaired last visit
    __syncwarp();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // see the band group's flush
    if (SCAN && g.counts) {
      const long long t0 = warp_sum_ll(c0), t1 = warp_sum_ll(c1), t2 = warp_sum_ll(c2);
      if (lane == 0) {
        if (t0) atomic_add_i64(g.counts + 0, t0);
        if (t1) atomic_add_i64(g.counts + 1, t1);
        if (t2) atomic_add_i64(g.counts + 2, t2);
      }
    }
    s_acos = warp_sum_f64(s_acos); s_n = warp_sum_f64(s_n);
    const int pw = warp - G::BAND_WARPS;
    if (lane == 0) { red[0][pw] = s_acos; red[2][pw] = s_n; }
    asm volatile("bar.sync 3, %0;" ::"r"(kPixelThreadsCT));
    if (tg < 32 && g.spec_acc) {
      double t0 = 0, t2 = 0;
      for (int w = 0; w < kPixelWarpsCT; ++w) { t0 += red[0][w]; t2 += red[2][w]; }
      ordered_block_sum3(t0, 0.0, t2, g.ws, g.spec_acc);
    }
    for (int i = tg; i < 256; i += kPixelThreadsCT) {
      if (g.hist8_g && h8g[i]) atomic_add_i64(g.hist8_g + i, h8g[i]);
      if (g.hist8_z && h8z[i]) atomic_add_i64(g.hist8_z + i, h8z[i]);
    }
  }
}

// -------------------------------------------------------------------------------------------------
// Validity plane of a 180-band BIP pair at HBM speed (the pre-pass of the nodata / caller-mask path,
// run_codec.py:249-263, quicklooks.py:35-45, run_codec.py:314-319).  Same TMA ring and tile walk as
// the pixel group above: two lanes per pixel OR / MIN the words of their half of the spectrum XORed
// with the packed nodata value -- OR != 0 <=> some band differs from nodata (dataset_mask), MIN != 0
// per half-word <=> no band equals it (the all-bands test).  Only cubes that have a nodata value are
// read at all.
struct ValArgs {
  const void* ref;
  const void* tst;
  const uint8_t* valid_in;    // may be null
  uint8_t* plane;
  int64_t* counts;            // may be null
  int64_t ntiles;
  int ref_has, ref_nd, tst_has, tst_nd;
  uint32_t poll_ns;
};

template <int BANDS>
__global__ void __launch_bounds__(kPixelThreadsCT, 1)
validity_ct_kernel(ValArgs g) {
  using G = Geo<BANDS, 4>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kPixelWarpsCT / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int my_tiles = (int64_t)blockIdx.x < g.ntiles ? (int)((g.ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
  const uint32_t poll_ns = DM_POLL(g);
  const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
  const uint32_t tx_bytes = (g.ref_has ? G::CUBE : 0) + (g.tst_has ? G::CUBE : 0);

  auto issue_tile = [&](int it) {
    if (it >= my_tiles) return;
    const int s = it & (kStages - 1);
    uint64_t* fb = &full_bar[s];
    const int64_t off = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * (int64_t)G::CUBE;
    unsigned char* dst = smem + (size_t)s * G::STAGE;
    mbar_expect_tx(fb, tx_bytes);
    if (g.ref_has) bulk_g2s(dst, static_cast<const char*>(g.ref) + off, G::CUBE, fb);
    if (g.tst_has) bulk_g2s(dst + G::CUBE, static_cast<const char*>(g.tst) + off, G::CUBE, fb);
  };
  auto release_tile = [&](int it) {
    const uint32_t eb = empty0 + 8u * (uint32_t)(it & (kStages - 1));
    if (mbar_arrive_is_last_a(eb)) {
      mbar_wait_a(eb, (uint32_t)((it / kStages) & 1), poll_ns);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_tile(it + kStages);
    }
  };
  if (tid == 0) {
    for (int it = 0; it < kStages; ++it) issue_tile(it);
  }
  const int grp = tid >> 7, wq = (tid >> 5) & 3, hl = lane >> 4;
  const int tp = 16 * wq + (lane & 15);
  constexpr int U = G::W / 2, H0 = (U + 1) / 2, H1 = U - H0;
  const unsigned char* lane_base = smem + (size_t)tp * G::PIXB + (hl ? H0 * 8 : 0);
  const uint32_t ndr = ((uint32_t)g.ref_nd & 0xffffu) * 0x10001u, ndt = ((uint32_t)g.tst_nd & 0xffffu) * 0x10001u;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int it = grp; it < my_tiles; it += 2) {
    const int s = it & (kStages - 1);
    mbar_wait_a(full0 + 8u * (uint32_t)s, (uint32_t)((it / kStages) & 1), poll_ns);
    const unsigned char* xs = lane_base + (size_t)s * G::STAGE;
    uint32_t r_or = 0, r_min = 0xffffffffu, t_or = 0, t_min = 0xffffffffu, r_b1 = 1, t_b1 = 1;
    auto scan = [&](const unsigned char* base, uint32_t nd2, uint32_t& o, uint32_t& m) {
      auto unit = [&](int j) {
        const uint2 v = *reinterpret_cast<const uint2*>(base + 8 * j);
        const uint32_t a = v.x ^ nd2, b = v.y ^ nd2;
        o |= a | b;
        m = __vimin3_u16x2(m, a, b);
      };
#pragma unroll 11
      for (int j = 0; j < H1; ++j) unit(j);
      if (H0 > H1 && hl == 0) unit(H1);
    };
    if (g.ref_has) { scan(xs, ndr, r_or, r_min); r_b1 = ((*reinterpret_cast<const uint32_t*>(xs) ^ ndr) & 0xffffu) != 0; }
    if (g.tst_has) { scan(xs + G::CUBE, ndt, t_or, t_min); t_b1 = ((*reinterpret_cast<const uint32_t*>(xs + G::CUBE) ^ ndt) & 0xffffu) != 0; }
    __syncwarp();
    if (lane == 0) release_tile(it);
    r_or |= __shfl_xor_sync(0xffffffffu, r_or, 16); t_or |= __shfl_xor_sync(0xffffffffu, t_or, 16);
    r_min = vminu2(r_min, __shfl_xor_sync(0xffffffffu, r_min, 16));
    t_min = vminu2(t_min, __shfl_xor_sync(0xffffffffu, t_min, 16));
    if (hl == 0) {                                     // lanes 0-15 hold band 1 and write the pixel
      const int64_t p = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * G::P + tp;
      const bool ra = g.ref_has ? r_or != 0 : true, ta = g.tst_has ? t_or != 0 : true;
      const bool rl = g.ref_has ? hmin2(r_min) != 0 : true, tl = g.tst_has ? hmin2(t_min) != 0 : true;
      const bool vin = g.valid_in ? g.valid_in[p] != 0 : true;
      const bool ds = ra && ta;
      uint8_t v = 0;
      if (ds && rl && tl && vin) v |= DM_VALID_METRICS;
      if (ds && r_b1 && t_b1) v |= DM_VALID_QUICKLOOK;
      if (g.valid_in ? vin : ds) v |= DM_VALID_SPECTRAL;
      g.plane[p] = v;
      c0 += (v & DM_VALID_METRICS) ? 1 : 0;
      c1 += (v & DM_VALID_QUICKLOOK) ? 1 : 0;
      c2 += (v & DM_VALID_SPECTRAL) ? 1 : 0;
    }
  }
  const long long s0 = warp_sum_ll(c0), s1 = warp_sum_ll(c1), s2 = warp_sum_ll(c2);
  if (lane == 0 && g.counts) {
    if (s0) atomic_add_i64(g.counts + 0, s0);
    if (s1) atomic_add_i64(g.counts + 1, s1);
    if (s2) atomic_add_i64(g.counts + 2, s2);
  }
}

}  // namespace

namespace {

// generic kernel over g.npix pixels (full tiles by TMA, one partial tile by plain loads)
int run_generic(FusedArgs g, int dtype, cudaStream_t s) {
  const int64_t B = g.bands;
  int P = kTilePixels;                                 // 64 pixels per tile when they fit a stage
  while ((int64_t)P * B * 4 > kStageBytesMax) P >>= 1;
  g.P = P;                                             // P*B*2 is a multiple of 16 (P even, B % 4 == 0)
  g.ntiles = g.npix / P;
  g.tail_pixels = (int)(g.npix - g.ntiles * P);
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const int64_t total = g.ntiles + (g.tail_pixels ? 1 : 0);
  int64_t grid = sms < kMaxPartialBlocks ? sms : kMaxPartialBlocks;
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
  const bool err = g.errmax || g.err8_g || g.err8_z;
  if (dtype == DM_U16) {
    if (g.plane) { if (err) DM_FUSED(DM_U16, true, true); else DM_FUSED(DM_U16, true, false); }
    else { if (err) DM_FUSED(DM_U16, false, true); else DM_FUSED(DM_U16, false, false); }
  } else {
    if (g.plane) { if (err) DM_FUSED(DM_I16, true, true); else DM_FUSED(DM_I16, true, false); }
    else { if (err) DM_FUSED(DM_I16, false, true); else DM_FUSED(DM_I16, false, false); }
  }
#undef DM_FUSED
  DM_LAUNCH_CHECK("fused_bip");
  return DM_OK;
}

// compile-time geometry kernel over g.ntiles FULL tiles of 64 pixels
template <int BANDS, int MPW>
int run_ct(FusedArgs g, int dtype, cudaStream_t s) {
  using G = Geo<BANDS, MPW>;
  g.P = G::P; g.tail_pixels = 0;
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  int64_t grid = sms < kMaxPartialBlocks ? sms : kMaxPartialBlocks;
  if (grid > g.ntiles) grid = g.ntiles;
  // programmatic dependent launch only while the caller has switched launch chaining on (dm_launch_chaining):
  // a chained launch reads its inputs before the preceding kernel's writes are guaranteed to be flushed
#ifdef DM_DEBUG_HOOKS
  static const bool pdl_allowed = []() { const char* e = getenv("DM_NO_PDL"); return !(e && atoi(e)); }();
#else
  constexpr bool pdl_allowed = true;
#endif
  const bool pdl = pdl_allowed && launch_chaining();
#define DM_FUSED_CT(DT, MASK, ERR, SCAN)                                                              \
  do {                                                                                                \
    auto k = fused_ct_kernel<BANDS, DT, MASK, ERR, MPW, SCAN>;                                        \
    DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM));      \
    cudaLaunchConfig_t cfg = {};                                                                      \
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(G::THREADS);                              \
    cfg.dynamicSmemBytes = G::SMEM; cfg.stream = s;                                                   \
    cudaLaunchAttribute at[1];                                                                        \
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                    \
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;                                   \
    cfg.attrs = at; cfg.numAttrs = 1;                                                                 \
    DM_CUDA(cudaLaunchKernelEx(&cfg, k, g));                                                          \
  } while (0)
  const bool err = g.errmax || g.err8_g || g.err8_z;
  if (g.scan) {
    if (dtype == DM_U16) { if (err) DM_FUSED_CT(DM_U16, true, true, true); else DM_FUSED_CT(DM_U16, true, false, true); }
    else { if (err) DM_FUSED_CT(DM_I16, true, true, true); else DM_FUSED_CT(DM_I16, true, false, true); }
  } else if (dtype == DM_U16) {
    if (g.plane) { if (err) DM_FUSED_CT(DM_U16, true, true, false); else DM_FUSED_CT(DM_U16, true, false, false); }
    else { if (err) DM_FUSED_CT(DM_U16, false, true, false); else DM_FUSED_CT(DM_U16, false, false, false); }
  } else {
    if (g.plane) { if (err) DM_FUSED_CT(DM_I16, true, true, false); else DM_FUSED_CT(DM_I16, true, false, false); }
    else { if (err) DM_FUSED_CT(DM_I16, false, true, false); else DM_FUSED_CT(DM_I16, false, false, false); }
  }
#undef DM_FUSED_CT
  DM_LAUNCH_CHECK("fused_bip_ct");
  return DM_OK;
}

}  // namespace

namespace {

// Common body of dm_fused_bip (plane = input, may be null) and dm_fused_bip_scan (scan != null: the validity
// plane is computed in the kernel from the pair's nodata values and the caller's mask).
struct ScanSpec { const uint8_t* valid_in; uint8_t* plane_out; int64_t* counts; };

int fused_bip_impl(const dm_pair_t& p, const uint8_t* plane, const ScanSpec* scan, int64_t* sums, int64_t* maxs,
                   uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                   const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                   double* spectral_acc, void* workspace, cudaStream_t s) {
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
  if (want_sam && (!spectral_acc || !workspace)) return fail(DM_EARG, "dm_fused_bip: spectral_acc / workspace is null");
  FusedArgs g;
  g.ref = p.ref; g.tst = p.tst; g.plane = plane; g.npix = p.rows * p.width; g.bands = (int)B;
  g.P = 0; g.ntiles = 0; g.tail_pixels = 0; g.zero = 0;
  g.sums = sums; g.maxs = maxs; g.errmax = errmax_out;
  g.lut_g = lut_g; g.cap_g = cap_g; g.err8_g = err8_g; g.hist8_g = err8_g ? hist8_g : nullptr;
  g.lut_z = lut_z; g.cap_z = cap_z; g.err8_z = err8_z; g.hist8_z = err8_z ? hist8_z : nullptr;
  g.want_sam = want_sam; g.spec_acc = want_sam ? spectral_acc : nullptr; g.ws = workspace;
  g.debug = 0; g.poll_ns = kPollNs;
  g.valid_in = nullptr; g.plane_out = nullptr; g.counts = nullptr; g.scan = 0;
  g.ref_has = p.ref_has_nodata; g.ref_nd = p.ref_nodata; g.tst_has = p.tst_has_nodata; g.tst_nd = p.tst_nodata;
#ifdef DM_DEBUG_HOOKS
  { const char* e = getenv("DM_FUSED_DEBUG"); g.debug = e ? atoi(e) : 0; }
  { const char* e = getenv("DM_POLL_NS"); g.poll_ns = e ? (uint32_t)atoi(e) : kPollNs; }
#endif
  if (g.npix <= 0) return DM_OK;
  const int variant = fused_bip_variant();             // dm_fused_bip_variant(): 0 auto | 12 | 23 | 1 = run-time-geometry kernel
  const bool force_generic = variant == 1 || (g.debug & 8) != 0;
  if (scan) {
    // the in-kernel scan exists in the specialised 180-band kernel only; a partial last tile takes the two-pass
    // route (dm_validity's generic kernel on those < 64 pixels, then the generic one-pass kernel with that plane)
    if (B != 180 || g.npix < kTilePixels || force_generic) return fail(DM_EUNSUPPORTED, "dm_fused_bip_scan: 180-band cubes of at least 64 pixels");
    if (reinterpret_cast<uintptr_t>(scan->valid_in) & 15) return fail(DM_EUNSUPPORTED, "dm_fused_bip_scan: valid_in must be 16-byte aligned");
    if (!scan->plane_out && g.npix % kTilePixels) return fail(DM_EARG, "dm_fused_bip_scan: a partial last tile needs plane_out");
    g.scan = 1; g.valid_in = scan->valid_in; g.plane_out = scan->plane_out; g.counts = scan->counts;
  }
  // (the specialised kernel bulk-copies the tile's 64 mask bytes: the plane must be 16-byte aligned)
  if (B == 180 && g.npix >= kTilePixels && !force_generic && !(reinterpret_cast<uintptr_t>(plane) & 15)) {
    // full 64-pixel tiles through the specialised kernel, the partial last tile through the generic one
    g.ntiles = g.npix / kTilePixels;
    const int64_t done = g.ntiles * kTilePixels;
    // Two builds of the kernel: 12 band warps (ldmatrix.x4, 96 registers) or 23 (ldmatrix.x2, 64 registers).
    // Measured in a sweep (bench.py): plain stats + SAM 135.5 us with 23 band warps against 137.2 us; with
    // error planes or a validity plane the 12-warp build wins (163 / 176 us against 165 / 190 us).
    // dm_fused_bip_variant(12 | 23) pins the choice (A/B runs, and the tests cover both builds).
    bool narrow = (!plane && !scan && !errmax_out && !err8_g && !err8_z) != ((g.debug & 16) != 0);
    if (variant == 12) narrow = false; else if (variant == 23) narrow = true;
    int rc = narrow ? run_ct<180, 2>(g, p.dtype, s) : run_ct<180, 4>(g, p.dtype, s);
    if (rc != DM_OK || done == g.npix) return rc;
    g.ref = static_cast<const char*>(g.ref) + done * B * 2;
    g.tst = static_cast<const char*>(g.tst) + done * B * 2;
    if (g.plane) g.plane += done;
    if (g.errmax) g.errmax += done;
    if (g.err8_g) g.err8_g += done;
    if (g.err8_z) g.err8_z += done;
    g.npix -= done;
    if (scan) {
      dm_pair_t q = p;
      q.ref = g.ref; q.tst = g.tst; q.rows = 1; q.width = g.npix;
      rc = launch_validity(q, scan->valid_in ? scan->valid_in + done : nullptr, scan->plane_out + done, scan->counts, s);
      if (rc != DM_OK) return rc;
      g.plane = scan->plane_out + done;
      g.scan = 0; g.valid_in = nullptr; g.plane_out = nullptr; g.counts = nullptr;
    }
  }
  return run_generic(g, p.dtype, s);
}

}  // namespace

int launch_fused_bip(const dm_pair_t& p, const uint8_t* plane, int64_t* sums, int64_t* maxs,
                     uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                     const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                     double* spectral_acc, void* workspace, cudaStream_t s) {
  return fused_bip_impl(p, plane, nullptr, sums, maxs, errmax_out, lut_g, cap_g, err8_g, hist8_g,
                        lut_z, cap_z, err8_z, hist8_z, want_sam, spectral_acc, workspace, s);
}

int launch_fused_bip_scan(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts,
                          int64_t* sums, int64_t* maxs,
                          uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                          const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                          double* spectral_acc, void* workspace, cudaStream_t s) {
  const ScanSpec sc{valid_in, plane_out, counts};
  return fused_bip_impl(p, nullptr, &sc, sums, maxs, errmax_out, lut_g, cap_g, err8_g, hist8_g,
                        lut_z, cap_z, err8_z, hist8_z, want_sam, spectral_acc, workspace, s);
}

// full 64-pixel tiles of a 180-band, 16-bit BIP pair through validity_ct_kernel; returns the number of
// pixels covered (0 when the geometry does not qualify: the caller then runs the generic kernel on everything)
int64_t launch_validity_ct(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts,
                           cudaStream_t s, int* status) {
  *status = DM_OK;
  const int64_t npix = p.rows * p.width;
  if (p.layout != DM_BIP || p.bands != 180 || (p.dtype != DM_U16 && p.dtype != DM_I16) || npix < kTilePixels) return 0;
  if ((reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) & 15) return 0;
  if (!p.ref_has_nodata && !p.tst_has_nodata) return 0;          // nothing to read: the generic kernel is a plain fill
  using G = Geo<180, 4>;
  ValArgs g;
  g.ref = p.ref; g.tst = p.tst; g.valid_in = valid_in; g.plane = plane_out; g.counts = counts;
  g.ntiles = npix / kTilePixels;
  g.ref_has = p.ref_has_nodata; g.ref_nd = p.ref_nodata; g.tst_has = p.tst_has_nodata; g.tst_nd = p.tst_nodata;
  g.poll_ns = kPollNs;
#ifdef DM_DEBUG_HOOKS
  { const char* e = getenv("DM_POLL_NS"); g.poll_ns = e ? (uint32_t)atoi(e) : kPollNs; }
#endif
  const int sms = sm_count();
  if (sms < 0) { *status = DM_ECUDA; return 0; }
  int64_t grid = sms;
  if (grid > g.ntiles) grid = g.ntiles;
  const size_t smem = (size_t)kStages * G::STAGE;
  auto k = validity_ct_kernel<180>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { *status = cuda_fail(e, "cudaFuncSetAttribute(validity_ct)"); return 0; }
  k<<<(unsigned)grid, kPixelThreadsCT, smem, s>>>(g);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) { *status = cuda_fail(e, "validity_ct"); return 0; }
  return g.ntiles * kTilePixels;
}

}  // namespace dm
