// Per-pixel spectral pass: one sweep over the spectral axis of every pixel.
//
//   error quicklook   /root/reference/tools/quicklooks.py:123-150  max_b |A-B| -> 8-bit via LUT
//   SAM               /root/reference/tools/run_codec.py:328-332
//   SID               /root/reference/tools/run_codec.py:334-339
//
// SAM's dot / |a|^2 / |r|^2 are exact int64 sums (<= 180*65535^2 < 2^53), so cos(angle) is
// reproduced bit-for-bit with correctly rounded sqrt / mul / div; acos and log are libdevice
// (<= 2 ulp).  SID follows the reference's float64 expressions (see sid_term).  Floating-point
// partial sums are combined in a fixed order by the last block of the launch (ordered_block_sum3)
// and accumulated into the caller's {sum arccos, sum sid, n}.
//
// The ERR8 scaling is a float32 chain in the reference (clip((e-0)/(cap+1e-9),0,1)*255 -> uint8);
// the host tabulates it with the reference's own expression for e = 0..cap and the kernel only
// indexes lut[min(e,cap)], which makes the planes bit-exact by construction.

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kSpecBlocks = kMaxPartialBlocks;   // fixed grid: the summation order must not depend on the device
constexpr int kSpecThreads = 256;

struct SpecArgs {
  const void* ref;
  const void* tst;
  const uint8_t* plane;
  int64_t bands, npix;
  int64_t sb, sp;             // element strides: band, pixel
  uint16_t* errmax;
  const uint8_t* lut_g; int cap_g; uint8_t* err8_g; int64_t* hist8_g;
  const uint8_t* lut_z; int cap_z; uint8_t* err8_z; int64_t* hist8_z;
  int want_sam, want_sid;
  double* acc;                // {sum arccos, sum sid, n}, accumulated
  void* ws;
};

template <typename T>
__device__ __forceinline__ int ld(const T* p, int64_t i) { return (int)__ldg(p + i); }

// warp-aggregated shared-memory histogram update: error maps are mostly a handful of values, so a
// plain atomicAdd would serialise 32 lanes on one address
__device__ __forceinline__ void hist_add(unsigned* h, unsigned bin) {
  const unsigned act = __activemask();
  const unsigned peers = __match_any_sync(act, bin);
  if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[bin], (unsigned)__popc(peers));
}

// One SID term  ap*ln(a/r) + rp*ln(r/a)  with a = ap+1e-15, r = rp+1e-15  (run_codec.py:338-339), written
// as (ap - rp) * ln(a/r): identical in exact arithmetic, and within the rounding noise of the
// reference's own two logarithms in float64.  Decoded spectra are close to the original, so the ratio
// sits next to 1 where ln(a/r) = 2 atanh(z), z = (a-r)/(a+r): six odd terms are exact to < 1e-17
// relative for |z| < 0.05 (one division instead of two divisions and two libdevice logs).
__device__ __forceinline__ double sid_term(double ap, double rp) {
  const double a = ap + 1e-15, r = rp + 1e-15;
  const double z = (a - r) / (a + r);
  double L;
  if (fabs(z) < 0.05) {
    const double z2 = z * z;
    double p = 1.0 / 11.0;
    p = fma(p, z2, 1.0 / 9.0);
    p = fma(p, z2, 1.0 / 7.0);
    p = fma(p, z2, 1.0 / 5.0);
    p = fma(p, z2, 1.0 / 3.0);
    p = fma(p, z2, 1.0);
    L = 2.0 * z * p;
  } else {
    L = log(a / r);
  }
  return (ap - rp) * L;
}

// atanh series coefficients in constant memory: FP64 instructions take them as constant-bank operands
// (as literals every use costs two uniform moves)
__constant__ double kAtanhC[8] = {1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0, 1.0, 2e-15, 1e-15};

// fast path of one SID term; *slow is set when |d/s| >= 0.05 and the caller must take log()
__device__ __forceinline__ double sid_term_fast(double ap, double rp, bool* slow) {
  const double d = ap - rp, sden = (ap + rp) + kAtanhC[6];
  float qf;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(qf) : "f"((float)sden));      // ~2^-23 relative (sden >= 2e-15: no denormals)
  double q = (double)qf;
  q = fma(fma(-sden, q, kAtanhC[5]), q, q);                  // Newton: 2^-46
  q = fma(fma(-sden, q, kAtanhC[5]), q, q);                  // 2^-92 -> rounding only
  const double z = d * q;
  *slow = fabs(z) >= 0.05;
  const double z2 = z * z;
  double pz = kAtanhC[0];
  pz = fma(pz, z2, kAtanhC[1]);
  pz = fma(pz, z2, kAtanhC[2]);
  pz = fma(pz, z2, kAtanhC[3]);
  pz = fma(pz, z2, kAtanhC[4]);
  pz = fma(pz, z2, kAtanhC[5]);
  return (d + d) * (z * pz);
}

// one SID term: the fast form, log() only where the series does not apply
__device__ __forceinline__ double sid_term_auto(double ap, double rp) {
  bool slow;
  const double t = sid_term_fast(ap, rp, &slow);
  return slow ? sid_term(ap, rp) : t;
}

// thread per pixel; bands at stride sb (BSQ: coalesced across the warp for every band)
template <typename T>
__global__ void __launch_bounds__(kSpecThreads)
spectral_pixel(SpecArgs g) {
  __shared__ unsigned hg[256], hz[256];
  __shared__ double red[3][kSpecThreads / 32];
  const T* ref = static_cast<const T*>(g.ref);
  const T* tst = static_cast<const T*>(g.tst);
  const int tid = threadIdx.x;
  hg[tid] = 0; hz[tid] = 0;
  __syncthreads();
  const int B = (int)g.bands;
  double s_acos = 0.0, s_sid = 0.0, s_n = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * kSpecThreads + tid; p < g.npix; p += (int64_t)gridDim.x * kSpecThreads) {
    const uint8_t v = g.plane ? g.plane[p] : (uint8_t)0xff;
    long long dot = 0, na2 = 0, nr2 = 0, sa = 0, sr = 0;
    int emax = 0, amin = 0x7fffffff, rmin = 0x7fffffff;
    const int64_t base = p * g.sp;
    for (int b = 0; b < B; ++b) {
      const int a = ld(ref, base + b * g.sb), r = ld(tst, base + b * g.sb);
      emax = max(emax, abs(a - r));
      dot += (long long)a * r; na2 += (long long)a * a; nr2 += (long long)r * r;
      sa += a; sr += r; amin = min(amin, a); rmin = min(rmin, r);
    }
    if (!(v & DM_VALID_QUICKLOOK)) emax = 0;                  // quicklooks.py:134
    if (g.errmax) g.errmax[p] = (uint16_t)emax;
    if (g.err8_g) {
      const uint8_t e8 = __ldg(g.lut_g + min(emax, g.cap_g));
      g.err8_g[p] = e8;
      if (g.hist8_g) hist_add(hg, e8);
    }
    if (g.err8_z) {
      const uint8_t e8 = __ldg(g.lut_z + min(emax, g.cap_z));
      g.err8_z[p] = e8;
      if (g.hist8_z) hist_add(hz, e8);
    }
    if ((g.want_sam || g.want_sid) && (v & DM_VALID_SPECTRAL)) {
      s_n += 1.0;
      if (g.want_sam) {
        const double na = __dadd_rn(__dsqrt_rn((double)na2), 1e-12);
        const double nr = __dadd_rn(__dsqrt_rn((double)nr2), 1e-12);
        double c = __ddiv_rn((double)dot, __dmul_rn(na, nr));
        c = fmin(1.0, fmax(-1.0, c));
        s_acos += acos(c);
      }
      if (g.want_sid) {
        // Ap = (a - amin + 1e-12) / sum_b(a - amin + 1e-12); the integer part of the sum is exact
        const double SA = (double)(sa - (long long)B * amin) + (double)B * 1e-12;
        const double SR = (double)(sr - (long long)B * rmin) + (double)B * 1e-12;
        const double iSA = 1.0 / SA, iSR = 1.0 / SR;
        double t = 0.0;
        for (int b = 0; b < B; ++b) {
          const int a = ld(ref, base + b * g.sb), r = ld(tst, base + b * g.sb);
          t += sid_term_auto(((double)(a - amin) + 1e-12) * iSA, ((double)(r - rmin) + 1e-12) * iSR);
        }
        s_sid += t;
      }
    }
  }
  // deterministic block reduction of the float partials
  const int lane = tid & 31, warp = tid >> 5;
  s_acos = warp_sum_f64(s_acos); s_sid = warp_sum_f64(s_sid); s_n = warp_sum_f64(s_n);
  if (lane == 0) { red[0][warp] = s_acos; red[1][warp] = s_sid; red[2][warp] = s_n; }
  __syncthreads();
  if (tid < 32 && g.acc) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < kSpecThreads / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    ordered_block_sum3(t0, t1, t2, g.ws, g.acc);
  }
  if (g.hist8_g && hg[tid]) atomic_add_i64(g.hist8_g + tid, hg[tid]);
  if (g.hist8_z && hz[tid]) atomic_add_i64(g.hist8_z + tid, hz[tid]);
}

// BIP: one warp per pixel; lanes stride over the contiguous spectrum (coalesced), shuffles combine.
template <typename T>
__global__ void __launch_bounds__(kSpecThreads)
spectral_warp_bip(SpecArgs g) {
  __shared__ unsigned hg[256], hz[256];
  __shared__ double red[3][kSpecThreads / 32];
  const T* ref = static_cast<const T*>(g.ref);
  const T* tst = static_cast<const T*>(g.tst);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  hg[tid] = 0; hz[tid] = 0;
  __syncthreads();
  const int B = (int)g.bands;
  double s_acos = 0.0, s_sid = 0.0, s_n = 0.0;     // meaningful in lane 0
  const int64_t wstride = (int64_t)gridDim.x * (kSpecThreads / 32);
  for (int64_t p = (int64_t)blockIdx.x * (kSpecThreads / 32) + warp; p < g.npix; p += wstride) {
    const uint8_t v = g.plane ? g.plane[p] : (uint8_t)0xff;
    long long dot = 0, na2 = 0, nr2 = 0, sa = 0, sr = 0;
    int emax = 0, amin = 0x7fffffff, rmin = 0x7fffffff;
    const int64_t base = p * (int64_t)B;
    for (int b = lane; b < B; b += 32) {
      const int a = ld(ref, base + b), r = ld(tst, base + b);
      emax = max(emax, abs(a - r));
      dot += (long long)a * r; na2 += (long long)a * a; nr2 += (long long)r * r;
      sa += a; sr += r; amin = min(amin, a); rmin = min(rmin, r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      emax = max(emax, __shfl_xor_sync(0xffffffffu, emax, o));
      amin = min(amin, __shfl_xor_sync(0xffffffffu, amin, o));
      rmin = min(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
    }
    const bool spec = (g.want_sam || g.want_sid) && (v & DM_VALID_SPECTRAL);
    if (spec) {
      dot = warp_sum_ll(dot); na2 = warp_sum_ll(na2); nr2 = warp_sum_ll(nr2);
      sa = warp_sum_ll(sa); sr = warp_sum_ll(sr);
    }
    if (!(v & DM_VALID_QUICKLOOK)) emax = 0;
    if (lane == 0) {
      if (g.errmax) g.errmax[p] = (uint16_t)emax;
      if (g.err8_g) {
        const uint8_t e8 = __ldg(g.lut_g + min(emax, g.cap_g));
        g.err8_g[p] = e8;
        if (g.hist8_g) hist_add(hg, e8);
      }
      if (g.err8_z) {
        const uint8_t e8 = __ldg(g.lut_z + min(emax, g.cap_z));
        g.err8_z[p] = e8;
        if (g.hist8_z) hist_add(hz, e8);
      }
    }
    if (spec) {
      if (lane == 0) {
        s_n += 1.0;
        if (g.want_sam) {
          const double na = __dadd_rn(__dsqrt_rn((double)na2), 1e-12);
          const double nr = __dadd_rn(__dsqrt_rn((double)nr2), 1e-12);
          double c = __ddiv_rn((double)dot, __dmul_rn(na, nr));
          c = fmin(1.0, fmax(-1.0, c));
          s_acos += acos(c);
        }
      }
      if (g.want_sid) {
        const double SA = (double)(sa - (long long)B * amin) + (double)B * 1e-12;
        const double SR = (double)(sr - (long long)B * rmin) + (double)B * 1e-12;
        const double iSA = 1.0 / SA, iSR = 1.0 / SR;
        double t = 0.0;
        for (int b = lane; b < B; b += 32) {
          const int a = ld(ref, base + b), r = ld(tst, base + b);
          t += sid_term_auto(((double)(a - amin) + 1e-12) * iSA, ((double)(r - rmin) + 1e-12) * iSR);
        }
        t = warp_sum_f64(t);
        if (lane == 0) s_sid += t;
      }
    }
  }
  if (lane == 0) { red[0][warp] = s_acos; red[1][warp] = s_sid; red[2][warp] = s_n; }
  __syncthreads();
  if (tid < 32 && g.acc) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < kSpecThreads / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    ordered_block_sum3(t0, t1, t2, g.ws, g.acc);
  }
  if (g.hist8_g && hg[tid]) atomic_add_i64(g.hist8_g + tid, hg[tid]);
  if (g.hist8_z && hz[tid]) atomic_add_i64(g.hist8_z + tid, hz[tid]);
}


// BIP, 16-bit samples, SAM / SID only (no error planes): a GROUP of LPP lanes per pixel (32 / LPP pixels per warp)
// with the spectrum held in REGISTERS as 32-bit words (two bands per lane and load), so the second sweep SID needs
// (after the minimum and the sum are known) re-reads nothing.  With one warp per pixel the per-PIXEL work -- four
// reductions, the set-up of the SID constants, the float64 finish of SAM in one lane -- was most of the kernel for
// EnMAP's 180 bands (5.6 samples per lane); with 8 lanes per pixel the same instructions serve four pixels.
//
// SID from ONE exact numerator per sample.  With a' = a - amin, r' = r - rmin (integers >= 0), eps = 1e-12,
// SA = sum a' + B eps, SR = sum r' + B eps (the integer parts exact), the reference's (run_codec.py:334-339)
//   Ap - Rp = n / (SA SR),   n = (a' + eps) SR - (r' + eps) SA
//   (Ap + 1e-15) / (Rp + 1e-15) = (1 + z) / (1 - z),   z = n / D,   D = (a'+eps) SR + (r'+eps) SA + 2e-15 SA SR
// so the term (Ap - Rp) ln(..) = n * 2 atanh(z) / (SA SR): two fused multiply-adds give n and D straight from the
// integer samples (a' SR and r' SA are exact products < 2^53 up to the eps parts, so the cancellation in n costs
// nothing -- this is where the float64 quotients Ap, Rp of the reference lose their digits, not here), 1/D is a
// MUFU.RCP64H seed + one Newton step (2^-40), atanh is six odd terms (|z| < 0.12: next term < 7e-13).  Every term is
// >= 0, so a relative error of 1e-12 per term is 1e-12 on the sum.  14 FP64 operations and no conversion-unit instruction
// per sample; 1 / (SA SR) is applied once per pixel.  |z| >= 0.12 (the bands where a spectrum has its minimum, dark
// bands with a large relative error) takes the reference's own expression with log(): those samples -- typically
// one or two per pixel, scattered over lanes and register slots -- are COMPACTED through a group-private list and
// evaluated side by side, one per lane, behind a warp-uniform guard.
template <int LPP>
__device__ __forceinline__ int group_add(int v) {
  if (LPP == 32) return __reduce_add_sync(0xffffffffu, v);
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPP>
__device__ __forceinline__ int group_min(int v) {
  if (LPP == 32) return __reduce_min_sync(0xffffffffu, v);
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
template <int LPP>
__device__ __forceinline__ long long group_add_ll(long long v) {
#pragma unroll
  for (int o = LPP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// exact int -> float64 for 0 <= v < 2^31.  r02 (late): the conversion instruction again.  The 2^52 mantissa splice
// (__hiloint2double(0x43300000, v) - 2^52) saves an FP64-pipe slot per conversion but costs two register moves to
// pair the sample with the constant high word (IMAD.MOV was 10 % of the executed instructions, profiles/
// r02i_ncu_sid_16_lanes.txt), and the kernel is bound by issue slots (66 %), not by the FP64 pipe (44 %): I2F.F64 is
// one issue slot.  Measured on the Case-B cube: SID 565 -> 494 us with 8 lanes per pixel (whose 12 words per lane no
// longer spill), 518 -> 507 us with 16; results bit-identical.  -DDM_SPLICE_I2D brings the splice back.
__device__ __forceinline__ double splice_u32(int v) {
#ifdef DM_SPLICE_I2D
  return __hiloint2double(0x43300000, v) - 4503599627370496.0;
#else
  return (double)v;
#endif
}

__device__ __forceinline__ uint32_t dp2a_lo_uu(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t dp2a_hi_uu(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t dp2a_lo_su(uint32_t a, uint32_t b, uint32_t c) { int r; asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"((int)c)); return (uint32_t)r; }
__device__ __forceinline__ uint32_t dp2a_hi_ss(uint32_t a, uint32_t b, uint32_t c) { int r; asm("dp2a.hi.s32.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"((int)c)); return (uint32_t)r; }
__device__ __forceinline__ uint32_t vmin_u16x2(uint32_t a, uint32_t b) { uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

template <int DT, int NWL, int LPP>
__global__ void __launch_bounds__(kSpecThreads, 2)
spectral_warp_bip16(SpecArgs g) {
  constexpr int PPW = 32 / LPP;                                      // pixels per warp
  constexpr uint32_t OFS = DT == DM_I16 ? 0x80008000u : 0u;          // int16 -> offset binary, both halves of a word
  constexpr int WARPS = kSpecThreads / 32;
  __shared__ double red[3][WARPS];
  __shared__ uint32_t slow_list[WARPS][PPW][2 * NWL * LPP];         // SID: a' | r' << 16 (both < 2^16) of the samples that need log()
  const uint32_t* ref = static_cast<const uint32_t*>(g.ref);
  const uint32_t* tst = static_cast<const uint32_t*>(g.tst);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / LPP, sl = lane % LPP;                       // pixel of the warp, lane of the pixel's group
  const int B = (int)g.bands, WPX = B >> 1;
  double s_acos = 0.0, s_n = 0.0;                  // meaningful in the first lane of every group
  double s_sid = 0.0;                              // per LANE: SID is a plain sum over pixels and bands, so the
                                                   // lanes' shares meet once, at the end of the kernel
  const int64_t wstride = (int64_t)gridDim.x * WARPS * PPW;
  uint32_t x[NWL], y[NWL], nx[NWL], ny[NWL];
  auto fetch = [&](int64_t p, uint32_t (&a)[NWL], uint32_t (&b)[NWL]) {
    const int64_t base = p * (int64_t)WPX;
#pragma unroll
    for (int j = 0; j < NWL; ++j) {
      const int idx = sl + LPP * j;
      if (idx < WPX && p < g.npix) { a[j] = __ldg(ref + base + idx); b[j] = __ldg(tst + base + idx); }
      else { a[j] = 0; b[j] = 0; }
    }
  };
  int64_t p0 = ((int64_t)blockIdx.x * WARPS + warp) * PPW;          // first pixel of this warp's current visit
  fetch(p0 + sub, nx, ny);
  for (; p0 < g.npix; p0 += wstride) {
    const int64_t p = p0 + sub;
#pragma unroll
    for (int j = 0; j < NWL; ++j) { x[j] = nx[j]; y[j] = ny[j]; }
    fetch(p + wstride, nx, ny);                                      // the next pixel's spectrum, in flight during this one
    // a group whose pixel is past the end or masked out runs along on zeros and contributes nothing
    const bool ok = p < g.npix && ((g.plane ? g.plane[p] : (uint8_t)0xff) & DM_VALID_SPECTRAL);
    // First sweep, SIMD-in-word on the packed pairs of samples (int16 in offset binary, u = x ^ 0x8000: order and
    // differences are those of the signed values): sums by dp2a against 0x0101, minima by VIMNMX.U16x2; SAM's three
    // products as dp2a lo/hi byte partials (x * lo8(y) and x * hi8(y), PRMT splits the bytes) -- 2 + 4 instructions
    // per sample where scalar code spent ~8 and three quarter-rate 64-bit multiply-adds.
    uint32_t su_a = 0, su_r = 0, mn_a = 0xffffffffu, mn_r = 0xffffffffu;
    uint32_t xxl = 0, xxh = 0, yyl = 0, yyh = 0, xyl = 0, xyh = 0;
#pragma unroll
    for (int j = 0; j < NWL; ++j) {
      if (sl + LPP * j < WPX) {
        const uint32_t ux = x[j] ^ OFS, uy = y[j] ^ OFS;
        su_a = dp2a_lo_uu(ux, 0x0101u, su_a); su_r = dp2a_lo_uu(uy, 0x0101u, su_r);
        mn_a = vmin_u16x2(mn_a, ux); mn_r = vmin_u16x2(mn_r, uy);
        if (g.want_sam) {
          const uint32_t px = __byte_perm(x[j], 0, 0x3120), py = __byte_perm(y[j], 0, 0x3120);
          if (DT == DM_I16) {      // signed samples: s16 x (unsigned low byte) and s16 x (signed high byte)
            xxl = dp2a_lo_su(x[j], px, xxl); xxh = dp2a_hi_ss(x[j], px, xxh);
            yyl = dp2a_lo_su(y[j], py, yyl); yyh = dp2a_hi_ss(y[j], py, yyh);
            xyl = dp2a_lo_su(x[j], py, xyl); xyh = dp2a_hi_ss(x[j], py, xyh);
          } else {
            xxl = dp2a_lo_uu(x[j], px, xxl); xxh = dp2a_hi_uu(x[j], px, xxh);
            yyl = dp2a_lo_uu(y[j], py, yyl); yyh = dp2a_hi_uu(y[j], py, yyh);
            xyl = dp2a_lo_uu(x[j], py, xyl); xyh = dp2a_hi_uu(x[j], py, xyh);
          }
        }
      }
    }
    const int sa = group_add<LPP>((int)su_a), sr = group_add<LPP>((int)su_r);                 // sums of u (< 2^31)
    const int amin = group_min<LPP>((int)min(mn_a & 0xffffu, mn_a >> 16));
    const int rmin = group_min<LPP>((int)min(mn_r & 0xffffu, mn_r >> 16));
    if (sl == 0 && ok) s_n += 1.0;
    if (g.want_sam) {
      // lo + 256 * hi: each part fits 32 bits up to 256 bands x 2 lanes' shares, the sum is an exact float64
      const int t_xxl = group_add<LPP>((int)xxl), t_xxh = group_add<LPP>((int)xxh), t_yyl = group_add<LPP>((int)yyl);
      const int t_yyh = group_add<LPP>((int)yyh), t_xyl = group_add<LPP>((int)xyl), t_xyh = group_add<LPP>((int)xyh);
      if (sl == 0 && ok) {
        double na2, nr2, dot;
        if (DT == DM_I16) {
          na2 = fma((double)t_xxh, 256.0, (double)t_xxl); nr2 = fma((double)t_yyh, 256.0, (double)t_yyl);
          dot = fma((double)t_xyh, 256.0, (double)t_xyl);
        } else {
          na2 = fma((double)(uint32_t)t_xxh, 256.0, (double)(uint32_t)t_xxl); nr2 = fma((double)(uint32_t)t_yyh, 256.0, (double)(uint32_t)t_yyl);
          dot = fma((double)(uint32_t)t_xyh, 256.0, (double)(uint32_t)t_xyl);
        }
        const double na = __dadd_rn(__dsqrt_rn(na2), 1e-12);
        const double nr = __dadd_rn(__dsqrt_rn(nr2), 1e-12);
        double c = __ddiv_rn(dot, __dmul_rn(na, nr));
        c = fmin(1.0, fmax(-1.0, c));
        s_acos += acos(c);
      }
    }
    if (g.want_sid) {
      // sum a' <= 512 * 65535 < 2^31: the integer parts go through the mantissa splice like the samples
      const double SA = splice_u32(sa - B * amin) + (double)B * 1e-12, SR = splice_u32(sr - B * rmin) + (double)B * 1e-12;
      const double cE = -1e-12 * (SR - SA);                            // P2 = r' SA + cE
      const double cD = 2.0 * SR * fma(1e-15, SA, 1e-12);              // D = a' SR + P2 + cD
      const uint32_t amin2 = (uint32_t)amin * 0x10001u, rmin2 = (uint32_t)rmin * 0x10001u;
      double t = 0.0;
      unsigned slow_mask = 0;                                          // bit 2j+h: this lane's sample needs log()
#pragma unroll
      for (int j = 0; j < NWL; ++j) {
        const bool have = ok && sl + LPP * j < WPX;
        // a' and r' of both samples of the word with one subtraction each (no borrow: every half >= its minimum)
        const uint32_t dx = (x[j] ^ OFS) - amin2, dy = (y[j] ^ OFS) - rmin2;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // int -> float64 by splicing the non-negative integer into the mantissa of 2^52 (one exact subtraction;
          // the conversion instruction costs two FP64-pipe slots, tools/ubench_fp64.cu)
          const double ad = splice_u32((int)(h ? dx >> 16 : dx & 0xffffu)), rd = splice_u32((int)(h ? dy >> 16 : dy & 0xffffu));
          const double p2 = fma(rd, SA, cE);
          const double n = fma(ad, SR, -p2);
          const double D = fma(ad, SR, p2) + cD;
          double q;
          asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(D));
          q = fma(fma(-D, q, kAtanhC[5]), q, q);                       // seed 2^-20 (measured) -> 2^-40: plenty for a 1e-6 gate
          const double z = n * q, z2 = z * z, z4 = z2 * z2;
          const bool slow = __double2hiint(z2) >= 0x3f8d7dbf;         // z^2 >= 0.0144 (high word of 0.0144), i.e. |z| >= 0.12
          // six odd terms of atanh(z) / z in z^2, Estrin form (three short chains instead of five dependent steps);
          // the next term is < 0.12^12 / 13 = 7e-13
          const double e0 = fma(z2, kAtanhC[4], kAtanhC[5]);           // 1 + z2/3
          const double e1 = fma(z2, kAtanhC[2], kAtanhC[3]);           // 1/5 + z2/7
          const double e2 = fma(z2, kAtanhC[0], kAtanhC[1]);           // 1/9 + z2/11
          const double pz = fma(z4, fma(z4, e2, e1), e0);
          if (have && !slow) t = fma(n, z * pz, t);
          slow_mask |= (have && slow) ? (1u << (2 * j + h)) : 0u;
        }
      }
      // 2 / (SA SR): seed + two Newton steps (SA SR >= (B eps)^2 ~ 3e-20 and <= 2^62: no exponent trouble)
      {
        const double ss = SA * SR;
        double q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(ss));
        q = fma(fma(-ss, q, 1.0), q, q);
        q = fma(fma(-ss, q, 1.0), q, q);
        t *= 2.0 * q;
      }
      if (__any_sync(0xffffffffu, slow_mask != 0)) {
        uint32_t* list = slow_list[warp][sub];
        const unsigned gmask = LPP == 32 ? 0xffffffffu : (((1u << LPP) - 1u) << (sub * LPP));
        const unsigned below = gmask & ((1u << lane) - 1u);          // lanes of this group in front of this one
        int base = 0;
#pragma unroll
        for (int j = 0; j < NWL; ++j) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const bool mine = (slow_mask >> (2 * j + h)) & 1u;
            const unsigned b = __ballot_sync(0xffffffffu, mine);
            if (mine) {
              const uint32_t dx = (x[j] ^ OFS) - amin2, dy = (y[j] ^ OFS) - rmin2;
              list[base + __popc(b & below)] = h ? (dx >> 16) | (dy & 0xffff0000u) : (dx & 0xffffu) | (dy << 16);
            }
            base += __popc(b & gmask);
          }
        }
        __syncwarp();
        if (base > 0) {
          const double iSA = 1.0 / SA, iSR = 1.0 / SR;
          const double eA = 1e-12 * iSA, eR = 1e-12 * iSR;
          for (int i = sl; i < base; i += LPP) {
            const uint32_t v = list[i];
            const double ap = fma((double)(v & 0xffffu), iSA, eA), rp = fma((double)(v >> 16), iSR, eR);   // run_codec.py:334-337
            t += (ap - rp) * log((ap + 1e-15) / (rp + 1e-15));                                   // :338-339
          }
        }
        __syncwarp();
      }
      s_sid += t;
    }
  }
  s_sid = warp_sum_f64(s_sid);
  s_acos = warp_sum_f64(s_acos); s_n = warp_sum_f64(s_n);           // (zero outside the groups' first lanes)
  if (lane == 0) { red[0][warp] = s_acos; red[1][warp] = s_sid; red[2][warp] = s_n; }
  __syncthreads();
  if (tid < 32 && g.acc) {
    double t0 = 0, t1 = 0, t2 = 0;
    for (int w = 0; w < WARPS; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
    ordered_block_sum3(t0, t1, t2, g.ws, g.acc);
  }
}

template <typename T>
int run_spectral(const SpecArgs& g, bool bip, cudaStream_t s) {
  if (bip && g.bands >= 16)
    spectral_warp_bip<T><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);
  else
    spectral_pixel<T><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);
  DM_LAUNCH_CHECK("spectral");
  return DM_OK;
}

}  // namespace

int launch_spectral(const dm_pair_t& p, const uint8_t* plane, uint16_t* errmax_out, const uint8_t* lut_g,
                    int cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z, int cap_z,
                    uint8_t* err8_z, int64_t* hist8_z, int want_sam, int want_sid, double* spectral_acc,
                    void* workspace, cudaStream_t s) {
  if (!p.ref || !p.tst) return fail(DM_EARG, "dm_spectral: null pointer");
  if (p.bands <= 0 || p.rows < 0 || p.width < 0) return fail(DM_EARG, "dm_spectral: bad geometry");
  if (p.layout != DM_BSQ && p.layout != DM_BIP) return fail(DM_EARG, "dm_spectral: bad layout");
  if (err8_g && (!lut_g || cap_g < 0 || cap_g > 65535)) return fail(DM_EARG, "dm_spectral: bad global LUT");
  if (err8_z && (!lut_z || cap_z < 0 || cap_z > 65535)) return fail(DM_EARG, "dm_spectral: bad zoom LUT");
  if ((want_sam || want_sid) && (!spectral_acc || !workspace))
    return fail(DM_EARG, "dm_spectral: spectral_acc / workspace is null");
  SpecArgs g;
  g.ref = p.ref; g.tst = p.tst; g.plane = plane; g.bands = p.bands; g.npix = p.rows * p.width;
  const bool bip = p.layout == DM_BIP;
  g.sb = bip ? 1 : p.band_stride; g.sp = bip ? p.bands : 1;
  g.errmax = errmax_out;
  g.lut_g = lut_g; g.cap_g = cap_g; g.err8_g = err8_g; g.hist8_g = err8_g ? hist8_g : nullptr;
  g.lut_z = lut_z; g.cap_z = cap_z; g.err8_z = err8_z; g.hist8_z = err8_z ? hist8_z : nullptr;
  g.want_sam = want_sam; g.want_sid = want_sid;
  g.acc = (want_sam || want_sid) ? spectral_acc : nullptr; g.ws = workspace;
  // SAM / SID only on a 16-bit BIP cube: the register-resident warp kernel
  if (bip && !errmax_out && !err8_g && !err8_z && (want_sam || want_sid) && p.dtype != DM_U8 && p.bands % 2 == 0 &&
      p.bands <= 256 /* 32-bit dp2a partials of a pixel */ && ((reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) & 3) == 0) {
    // lanes per pixel: 8 while a pixel's words fit 12 per lane (<= 192 bands: EnMAP's 180 -> four pixels per warp),
    // then 16, then the whole warp
    const int wpx = (int)(p.bands / 2);
    int lpp = spectral_lanes_per_pixel();                                // dm_spectral_lanes_per_pixel(): 0 = auto
    // measured on the Case-B cube: SID 494 / 507 / 506 us and SAM + SID 581 / 653 / 769 us with 8 / 16 / 32 lanes per pixel
    // (r02e, with the mantissa splice and 12 spilling words per lane at 8 lanes: 627 / 510 / 509 and 760 / 674 / 832)
    if (lpp == 0) lpp = wpx <= 8 * 12 ? 8 : (wpx <= 16 * 8 ? 16 : 32);
    if ((lpp == 8 && wpx > 8 * 12) || (lpp == 16 && wpx > 16 * 8)) lpp = 32;
#define DM_SPEC16(DT)                                                                                   \
    do {                                                                                                \
      if (lpp == 8 && wpx <= 8 * 6) spectral_warp_bip16<DT, 6, 8><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);          \
      else if (lpp == 8) spectral_warp_bip16<DT, 12, 8><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);        \
      else if (lpp == 16 && wpx <= 16 * 6) spectral_warp_bip16<DT, 6, 16><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);  \
      else if (lpp == 16) spectral_warp_bip16<DT, 8, 16><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);       \
      else if (wpx <= 32 * 3) spectral_warp_bip16<DT, 3, 32><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);   \
      else spectral_warp_bip16<DT, 8, 32><<<kSpecBlocks, kSpecThreads, 0, s>>>(g);                      \
    } while (0)
    if (p.dtype == DM_I16) DM_SPEC16(DM_I16); else DM_SPEC16(DM_U16);
#undef DM_SPEC16
    DM_LAUNCH_CHECK("spectral_bip16");
    return DM_OK;
  }
  switch (p.dtype) {
    case DM_U8: return run_spectral<uint8_t>(g, bip, s);
    case DM_U16: return run_spectral<uint16_t>(g, bip, s);
    case DM_I16: return run_spectral<int16_t>(g, bip, s);
  }
  return fail(DM_EARG, "dm_spectral: bad dtype");
}

}  // namespace dm
