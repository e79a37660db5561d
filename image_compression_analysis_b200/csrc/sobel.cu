// Sobel LMSE: for every band, sum over pixels of (|grad ref| - |grad tst|)^2.
//
// Replaces sobel_mag + mse in the LMSE loop of compute_sam_sid_lmse_caseB
// (/root/reference/tools/run_codec.py:123-137, 341-346).  gx, gy are exact integers
// (|g| <= 4*65535), gx^2+gy^2 < 2^38 is exact in float64; a term is evaluated with one square root and no
// cancellation (lmse_term: ~1e-12 relative, all terms >= 0; the gate is 1e-6 on the band's sum); the order of the
// final float64 sum differs (block-ordered partials, added in a fixed order by the blocks that finish last and
// accumulated into the caller's per-band sums: no follow-up reduction kernel).
//
// BSQ: shared-memory tiled, a block stages a (32+2) x (32+2) tile of both cubes once and every sample is
// read from HBM ~1.13 times.  BIP (16-bit samples, even band count): no transposition -- a thread owns one
// 32-bit word of the spectrum (two bands), so a warp's loads are the contiguous spectrum of one pixel, and
// marches along x over two image rows with the 3x3 window kept in registers as separable column terms
// (vertical [1,2,1] sums and top-bottom differences of the last three columns).

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kSobBlocks = 296;   // blocks per band (fixed: the summation order must not depend on the device)
constexpr int kSobGroup = 8;      // BIP kernel: blocks per group of the two-level final sum
constexpr int kSobGroups = (kSobBlocks + kSobGroup - 1) / kSobGroup;
static_assert(kSobGroups <= kMaxGroups, "Workspace::group_counter too small");
constexpr int TW = 32, TH = 32;

// One LMSE term (|grad a| - |grad r|)^2 from the four integer gradients (|g| <= 4*65535), with ONE square root
// and no cancellation:   with ga = gxa^2 + gya^2, gr = gxr^2 + gyr^2 (exact float64 integers < 2^38)
//     (sqrt ga - sqrt gr)^2 = (ga - gr)^2 / (sqrt ga + sqrt gr)^2 = (ga - gr)^2 / (ga + gr + 2 sqrt(ga gr)).
// ga - gr is exact, the denominator is a sum of non-negative terms, so the quotient is good to the accuracy of
// sqrt(ga gr) and of the reciprocal -- MUFU.RSQ64H / MUFU.RCP64H seeds (2^-20, tools/ubench_fp64.cu) + one Newton step
// each: ~1e-12 relative, and every term is >= 0, so that is also the relative error of the band's sum (gate 1e-6).
// The reference's two correctly rounded roots (run_codec.py:137, 344) cost twice the FP64-pipe slots, and the
// kernel is issue bound: 20 FP64 operations + 2 MUFU per sample pair here against 40 + conversions before.  The
// integers enter float64 through the 2^52 mantissa splice (one exact subtraction; an I2F.F64 costs two pipe slots).
// 1e-30 is added to both sums of squares: it vanishes next to any non-zero sum and keeps 0/0 out of flat areas
// (ga = gr = 0 gives exactly 0).
__device__ __forceinline__ double splice_s32(int g) {          // exact for |g| < 2^20
  return __hiloint2double(0x43300000, g + (1 << 20)) - (4503599627370496.0 + 1048576.0);
}
__device__ __forceinline__ double lmse_term(int gxa, int gya, int gxr, int gyr) {
  const double fxa = splice_s32(gxa), fya = splice_s32(gya), fxr = splice_s32(gxr), fyr = splice_s32(gyr);
  const double ga = fma(fxa, fxa, fma(fya, fya, 1e-30));
  const double gr = fma(fxr, fxr, fma(fyr, fyr, 1e-30));
  const double dg = ga - gr, p = ga * gr;
  double y, q;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  double sq = p * y;                                            // ~ sqrt(ga gr)
  sq = fma(fma(-sq, sq, p), 0.5 * y, sq);                       // one Newton step
  const double den = fma(2.0, sq, ga + gr);
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(den));
  q = fma(fma(-den, q, 1.0), q, q);
  return (dg * dg) * q;
}

template <typename T>
__global__ void __launch_bounds__(256)
sobel_lmse_kernel(const T* __restrict__ ref, const T* __restrict__ tst, int64_t band_stride, int64_t width,
                  int64_t row_begin, int64_t row_end, int64_t img_row0, int64_t img_rows, double* scratch,
                  double* lmse_acc, void* workspace) {
  __shared__ int sa[TH + 2][TW + 2 + 1];
  __shared__ int sr[TH + 2][TW + 2 + 1];
  __shared__ double red[32];
  const int band = blockIdx.y;
  const T* A = ref + (int64_t)band * band_stride;
  const T* R = tst + (int64_t)band * band_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t nrows = row_end - row_begin;
  const int64_t tiles_x = (width + TW - 1) / TW, tiles_y = (nrows + TH - 1) / TH;
  const int64_t ntiles = tiles_x * tiles_y;
  double acc = 0.0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = row_begin + (t / tiles_x) * TH, c0 = (t % tiles_x) * TW;
    __syncthreads();
    // stage: warp ty takes rows ty, ty+8, ...; lane = column (lanes 0/1 also take columns 32/33).
    // np.pad(mode="edge"): clamp to the IMAGE, then map back to the buffer row
    {
      int64_t ca = c0 + tx - 1, cb = c0 + tx + 32 - 1;
      ca = ca < 0 ? 0 : (ca >= width ? width - 1 : ca);
      cb = cb < 0 ? 0 : (cb >= width ? width - 1 : cb);
      for (int lr = ty; lr < TH + 2; lr += 8) {
        int64_t ir = img_row0 + r0 + lr - 1;
        ir = ir < 0 ? 0 : (ir >= img_rows ? img_rows - 1 : ir);
        const T* ar = A + (ir - img_row0) * width;
        const T* rr = R + (ir - img_row0) * width;
        sa[lr][tx] = (int)ar[ca];
        sr[lr][tx] = (int)rr[ca];
        if (tx < 2) { sa[lr][tx + 32] = (int)ar[cb]; sr[lr][tx + 32] = (int)rr[cb]; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < TH; j += 8) {
      const int lr = ty + j;
      if (r0 + lr < row_end && c0 + tx < width) {
        const int (*s)[TW + 3] = sa;
        int gx[2], gy[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int p00 = s[lr][tx], p01 = s[lr][tx + 1], p02 = s[lr][tx + 2];
          const int p10 = s[lr + 1][tx], p12 = s[lr + 1][tx + 2];
          const int p20 = s[lr + 2][tx], p21 = s[lr + 2][tx + 1], p22 = s[lr + 2][tx + 2];
          gx[q] = (p00 - p02) + 2 * (p10 - p12) + (p20 - p22);             // |g| <= 4*65535: int32
          gy[q] = (p00 - p20) + 2 * (p01 - p21) + (p02 - p22);
          s = sr;
        }
        acc += lmse_term(gx[0], gy[0], gx[1], gy[1]);
      }
    }
  }
  acc = warp_sum_f64(acc);
  if (tx == 0) red[ty] = acc;
  __syncthreads();
  double t[1] = {0.0};
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) t[0] += red[w];
  }
  __syncthreads();
  double* const accs[1] = {lmse_acc};
  ordered_band_sum<1>(t, scratch, static_cast<Workspace*>(workspace)->band_counter, accs, red);
}


// ---- BIP ------------------------------------------------------------------------------------------
constexpr int kBipThreads = 192;      // groups of WPpad threads (WPpad = words per pixel rounded up to a warp)
constexpr int kBipCols = 30;          // columns per work item (a multiple of the 3-column trip); an item is two image rows x kBipCols columns

template <int DT>
struct ColTerms {                     // separable Sobel terms of one image column, two output rows, two bands
  int v[2][2], d[2][2];               // [output row][band]:  v = p(y-1) + 2 p(y) + p(y+1),  d = p(y-1) - p(y+1)
  __device__ __forceinline__ void set(const uint32_t (&w)[4]) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p0 = sample16<DT>(w[0], h), p1 = sample16<DT>(w[1], h), p2 = sample16<DT>(w[2], h), p3 = sample16<DT>(w[3], h);
      v[0][h] = p0 + 2 * p1 + p2; d[0][h] = p0 - p2;
      v[1][h] = p1 + 2 * p2 + p3; d[1][h] = p1 - p3;
    }
  }
};


template <int DT>
__global__ void __launch_bounds__(kBipThreads, 2)
sobel_lmse_bip_kernel(const uint32_t* __restrict__ ref, const uint32_t* __restrict__ tst, int wp, int wpad, int64_t width,
                      int64_t row_begin, int64_t row_end, int64_t img_row0, int64_t img_rows, double* scratch,
                      double* lmse_acc, void* workspace) {
  __shared__ double red[kBipThreads][2];
  __shared__ unsigned s_flag;
  const int groups = kBipThreads / wpad;
  const int grp = threadIdx.x / wpad, w = threadIdx.x - grp * wpad;
  const bool active = grp < groups && w < wp;
  const int64_t nrows = row_end - row_begin;
  const int64_t items_x = (width + kBipCols - 1) / kBipCols, items_y = (nrows + 1) / 2;
  const int64_t nitems = items_x * items_y;
  double acc0 = 0.0, acc1 = 0.0;
  if (active) {
    for (int64_t it = (int64_t)blockIdx.x * groups + grp; it < nitems; it += (int64_t)gridDim.x * groups) {
      const int64_t y0 = row_begin + (it / items_x) * 2;
      const int c0 = (int)((it % items_x) * kBipCols);
      const int c1 = c0 + kBipCols < (int)width ? c0 + kBipCols : (int)width;
      const bool two = y0 + 1 < row_end;
      // the four window rows, clamped to the IMAGE (np.pad mode="edge"), as 32-bit word offsets from the item's first
      // row: every load is then one IMAD.WIDE.U32 (base + 4 * offset) instead of five instructions of 64-bit address
      // arithmetic (r02 ncu: 40 of 185 instructions per column step were addresses)
      int64_t rows[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        int64_t ir = img_row0 + y0 + k - 1;
        ir = ir < 0 ? 0 : (ir >= img_rows ? img_rows - 1 : ir);
        rows[k] = ir - img_row0;
      }
      const int64_t base = rows[0] * width * wp;                  // rows[] is non-decreasing
      const uint32_t* refb = ref + base;
      const uint32_t* tstb = tst + base;
      uint32_t ro[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) ro[k] = (uint32_t)((rows[k] - rows[0]) * width * wp) + (uint32_t)w;
      // raw words of one column (four window rows, both cubes); loads are issued one column AHEAD of their
      // use so that their latency hides behind the previous column's square roots
      auto fetch = [&](int c, uint32_t (&wa)[4], uint32_t (&wr)[4]) {
        c = c < 0 ? 0 : (c >= (int)width ? (int)width - 1 : c);
        const uint32_t cw = (uint32_t)c * (uint32_t)wp;
#pragma unroll
        for (int k = 0; k < 4; ++k) { wa[k] = __ldg(refb + (ro[k] + cw)); wr[k] = __ldg(tstb + (ro[k] + cw)); }
      };
      auto emit = [&](const ColTerms<DT>& aL, const ColTerms<DT>& aC, const ColTerms<DT>& aR,
                      const ColTerms<DT>& rL, const ColTerms<DT>& rC, const ColTerms<DT>& rR, bool live) {
#pragma unroll
        for (int o = 0; o < 2; ++o) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const double term = lmse_term(aL.v[o][h] - aR.v[o][h], aL.d[o][h] + 2 * aC.d[o][h] + aR.d[o][h],
                                          rL.v[o][h] - rR.v[o][h], rL.d[o][h] + 2 * rC.d[o][h] + rR.d[o][h]);
            const double t = (live && (o == 0 || two)) ? term : 0.0;          // column past the item / row past the strip
            if (h == 0) acc0 += t; else acc1 += t;
          }
        }
      };
      ColTerms<DT> a0, a1, a2, r0, r1, r2;
      // Raw words of three columns in flight, one buffer per step of the three-column trip: column c lives in buffer
      // (c - c0) mod 3, is consumed as the window's right edge one step after ... three steps after its loads were issued
      // and refilled on the spot with the column three ahead -- the buffers rotate by NAME like the window (no register
      // copies: 8 moves per step in the one-buffer form), and a load has three steps (~550 instructions) to land (the
      // one-column lead left 0.94 stall cycles per issued instruction on the scoreboard, profiles/r02f_ncu_lmse.txt).
      uint32_t wa0[4], wr0[4], wa1[4], wr1[4], wa2[4], wr2[4];
      fetch(c0 - 1, wa0, wr0); a0.set(wa0); r0.set(wr0);
      fetch(c0, wa0, wr0); a1.set(wa0); r1.set(wr0);
      fetch(c0 + 1, wa1, wr1);                                 // right edges of output columns c0, c0 + 1, c0 + 2
      fetch(c0 + 2, wa2, wr2);
      fetch(c0 + 3, wa0, wr0);
#define DM_SOBEL_STEP(L_, C_, R_, x_, B_)                                               \
      do {                                                                              \
        a##R_.set(wa##B_); r##R_.set(wr##B_);                                           \
        fetch((x_) + 4, wa##B_, wr##B_);                                                \
        emit(a##L_, a##C_, a##R_, r##L_, r##C_, r##R_, (x_) < c1);                      \
      } while (0)
      // the last trip may run one or two columns past the item: their (clamped) loads are harmless and
      // their terms are dropped -- cheaper than a second copy of the step code for the tail
      for (int x = c0; x < c1; x += 3) {
        DM_SOBEL_STEP(0, 1, 2, x, 1);
        DM_SOBEL_STEP(1, 2, 0, x + 1, 2);
        DM_SOBEL_STEP(2, 0, 1, x + 2, 0);
      }
#undef DM_SOBEL_STEP
    }
  }
  red[threadIdx.x][0] = acc0; red[threadIdx.x][1] = acc1;
  __syncthreads();
  // Two-level ordered final sum.  Every block holds a partial for EVERY band, so a single "last block" would have
  // to add gridDim.x x bands values by itself; instead the block that finishes last in each group of kSobGroup
  // blocks adds the group's partials (level 1, in block order) and the group that finishes last adds the group
  // sums (level 2, in group order) into the caller's per-band accumulators.  Thread w owns bands 2w, 2w+1.
  Workspace* ws = static_cast<Workspace*>(workspace);
  const int nb = 2 * wp, grp_id = blockIdx.x / kSobGroup;
  const int ngroups = (gridDim.x + kSobGroup - 1) / kSobGroup;
  const int gsize = min(kSobGroup, (int)gridDim.x - grp_id * kSobGroup);
  double* l1 = scratch;                                    // [gridDim.x][bands]
  double* l2 = scratch + (size_t)gridDim.x * nb;           // [ngroups][bands]
  if (threadIdx.x < wp) {
    double t0 = 0.0, t1 = 0.0;
    for (int gi = 0; gi < groups; ++gi) { t0 += red[gi * wpad + threadIdx.x][0]; t1 += red[gi * wpad + threadIdx.x][1]; }
    __stcg(reinterpret_cast<double2*>(l1 + (size_t)blockIdx.x * nb) + threadIdx.x, make_double2(t0, t1));
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_flag = atomicAdd(&ws->group_counter[grp_id], 1u) == (unsigned)gsize - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (threadIdx.x < wp) {
    double u0 = 0.0, u1 = 0.0;
    for (int k = 0; k < gsize; ++k) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(l1 + (size_t)(grp_id * kSobGroup + k) * nb) + threadIdx.x);
      u0 += v.x; u1 += v.y;
    }
    __stcg(reinterpret_cast<double2*>(l2 + (size_t)grp_id * nb) + threadIdx.x, make_double2(u0, u1));
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ws->group_counter[grp_id] = 0;
    s_flag = atomicAdd(&ws->counter[1], 1u) == (unsigned)ngroups - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (threadIdx.x < wp) {
    double u0 = 0.0, u1 = 0.0;
    for (int k = 0; k < ngroups; ++k) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(l2 + (size_t)k * nb) + threadIdx.x);
      u0 += v.x; u1 += v.y;
    }
    lmse_acc[2 * threadIdx.x] += u0;
    lmse_acc[2 * threadIdx.x + 1] += u1;
  }
  if (threadIdx.x == 0) ws->counter[1] = 0;
}

// ---- magnitude map (the reference's sobel_mag as a function of its own, run_codec.py:123-137) -----------
// thread per pixel, neighbours through L1 / L2 (an (H,W) plane is read ~once from HBM), edge replication by
// index clamping, exact integer gradients, correctly rounded square root: bit-identical to the reference.
template <typename T>
__global__ void __launch_bounds__(256)
sobel_mag_kernel(const T* __restrict__ img, int64_t rows, int64_t width, double* __restrict__ out) {
  const int64_t n = rows * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t y = i / width, x = i - y * width;
    const int64_t y0 = y > 0 ? y - 1 : 0, y2 = y + 1 < rows ? y + 1 : rows - 1;
    const int64_t x0 = x > 0 ? x - 1 : 0, x2 = x + 1 < width ? x + 1 : width - 1;
    const T* r0 = img + y0 * width;
    const T* r1 = img + y * width;
    const T* r2 = img + y2 * width;
    const int p00 = (int)__ldg(r0 + x0), p01 = (int)__ldg(r0 + x), p02 = (int)__ldg(r0 + x2);
    const int p10 = (int)__ldg(r1 + x0), p12 = (int)__ldg(r1 + x2);
    const int p20 = (int)__ldg(r2 + x0), p21 = (int)__ldg(r2 + x), p22 = (int)__ldg(r2 + x2);
    const int gx = (p00 - p02) + 2 * (p10 - p12) + (p20 - p22);
    const int gy = (p00 - p20) + 2 * (p01 - p21) + (p02 - p22);
    const double fx = (double)gx, fy = (double)gy;
    out[i] = __dsqrt_rn(fma(fx, fx, fy * fy));
  }
}

}  // namespace

int launch_sobel_mag(const void* img, int dtype, int64_t rows, int64_t width, double* out, cudaStream_t s) {
  if (!img || !out) return fail(DM_EARG, "dm_sobel_mag: null pointer");
  if (rows < 0 || width < 0) return fail(DM_EARG, "dm_sobel_mag: bad geometry");
  if (rows * width == 0) return DM_OK;
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  int64_t grid = (rows * width + 255) / 256;
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  switch (dtype) {
    case DM_U8: sobel_mag_kernel<uint8_t><<<(unsigned)grid, 256, 0, s>>>(static_cast<const uint8_t*>(img), rows, width, out); break;
    case DM_U16: sobel_mag_kernel<uint16_t><<<(unsigned)grid, 256, 0, s>>>(static_cast<const uint16_t*>(img), rows, width, out); break;
    case DM_I16: sobel_mag_kernel<int16_t><<<(unsigned)grid, 256, 0, s>>>(static_cast<const int16_t*>(img), rows, width, out); break;
    default: return fail(DM_EARG, "dm_sobel_mag: bad dtype");
  }
  DM_LAUNCH_CHECK("sobel_mag");
  return DM_OK;
}

int sobel_nblocks() { return kSobBlocks + kSobGroups; }   // scratch slots per band: block partials + group sums

int launch_sobel(const dm_pair_t& p, int64_t row_begin, int64_t row_end, int64_t img_row0, int64_t img_rows,
                 double* scratch, double* lmse_acc, void* workspace, cudaStream_t s) {
  if (!p.ref || !p.tst || !scratch || !lmse_acc || !workspace) return fail(DM_EARG, "dm_sobel_lmse: null pointer");
  if ((reinterpret_cast<uintptr_t>(scratch) & 15) != 0) return fail(DM_EARG, "dm_sobel_lmse: scratch must be 16-byte aligned");
  if (p.layout != DM_BSQ && p.layout != DM_BIP) return fail(DM_EARG, "dm_sobel_lmse: bad layout");
  // (four window rows of a work item are addressed by 32-bit word offsets: 4 * width * bands / 2 must stay below 2^32)
  if (p.layout == DM_BIP && (p.dtype == DM_U8 || p.bands % 2 != 0 || p.bands > 2 * kBipThreads ||
                             p.width * p.bands * 2 >= (int64_t)1 << 32 ||
                             (reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) % 4 != 0))
    return fail(DM_EUNSUPPORTED, "dm_sobel_lmse: BIP needs 16-bit samples, an even band count and 4-byte aligned cubes "
                                 "(otherwise transpose with dm_bip_to_bsq)");
  if (p.bands <= 0 || p.bands > kMaxCounterBands || p.width <= 0) return fail(DM_EARG, "dm_sobel_lmse: bad geometry (1..2048 bands)");
  if (row_begin < 0 || row_end > p.rows || row_begin > row_end || img_row0 < 0 || img_row0 + p.rows > img_rows)
    return fail(DM_EARG, "dm_sobel_lmse: bad row range");
  // halo rows must be present unless the strip touches the image border
  if ((row_begin == 0 && img_row0 > 0) || (row_end == p.rows && img_row0 + p.rows < img_rows))
    return fail(DM_EARG, "dm_sobel_lmse: strip lacks its halo row");
  if (p.layout == DM_BIP) {
    const int wp = (int)(p.bands / 2), wpad = (wp + 31) / 32 * 32;
    if (p.dtype == DM_I16)
      sobel_lmse_bip_kernel<DM_I16><<<kSobBlocks, kBipThreads, 0, s>>>(static_cast<const uint32_t*>(p.ref), static_cast<const uint32_t*>(p.tst),
                                                                      wp, wpad, p.width, row_begin, row_end, img_row0, img_rows, scratch, lmse_acc, workspace);
    else
      sobel_lmse_bip_kernel<DM_U16><<<kSobBlocks, kBipThreads, 0, s>>>(static_cast<const uint32_t*>(p.ref), static_cast<const uint32_t*>(p.tst),
                                                                      wp, wpad, p.width, row_begin, row_end, img_row0, img_rows, scratch, lmse_acc, workspace);
    DM_LAUNCH_CHECK("sobel_lmse_bip");
    return DM_OK;
  }
  const dim3 grid(kSobBlocks, (unsigned)p.bands);
#define DM_SOBEL(T)                                                                                          \
  sobel_lmse_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(p.ref), static_cast<const T*>(p.tst),      \
                                            p.band_stride, p.width, row_begin, row_end, img_row0, img_rows, scratch, \
                                            lmse_acc, workspace)
  switch (p.dtype) {
    case DM_U8: DM_SOBEL(uint8_t); break;
    case DM_U16: DM_SOBEL(uint16_t); break;
    case DM_I16: DM_SOBEL(int16_t); break;
    default: return fail(DM_EARG, "dm_sobel_lmse: bad dtype");
  }
#undef DM_SOBEL
  DM_LAUNCH_CHECK("sobel_lmse");
  return DM_OK;
}

}  // namespace dm
