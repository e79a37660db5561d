// Gaussian-window SSIM per band (addition x1 of SURVEY.md section 8a; the reference has no windowed
// SSIM -- parity is pinned to the scipy restatement in oracle/distortion_oracle.py).
//
// skimage.metrics.structural_similarity(gaussian_weights=True, sigma=1.5,
// use_sample_covariance=False, data_range=L) semantics: 11-tap separable Gaussian (radius 5) of
// x, y, x^2, y^2, xy in float64, S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),
// mean over the image cropped by 5 px.  Because of the crop no counted window touches the image
// border, so no boundary rule is needed.  The kernel filters x^2 + y^2 as ONE plane (only vx + vy
// is used), i.e. four planes instead of five.
//
// Shared-memory tiled, separable, FP64-FMA bound (4 planes x 11 taps x 2 passes = 88 DFMA per pixel, so
// <= ~0.55 TB/s of pair bytes at the measured 62.7 DFMA lanes/clk/SM); the tiling keeps everything else
// off the FP64 pipe's back:
//   stage   (54+10) x (32+10) samples of both cubes as int32 in shared memory
//   pass H  a thread makes FOUR neighbouring outputs of one row: 14 inputs are converted to float64 and
//           squared/multiplied ONCE (not once per tap), 176 DFMA, results to four float64 planes in
//           shared memory; 64 rows x 8 groups = 512 items = exactly two per thread
//   pass V  a thread makes SEVEN vertically neighbouring outputs of one column, plane by plane: 17
//           shared-memory loads feed 77 DFMA (2.4 loads per output and plane instead of 11), then the
//           SSIM formula; block-ordered partial sums
// Column index is the fastest thread index in both passes, so shared-memory accesses are conflict free.
// (Tried: 38-row tiles with 192 threads and 71 KB, three CTAs per SM -- 6.87 ms per scene against 6.20 ms: the
// extra halo rows of the shorter tile cost more than the third CTA's overlap buys.)

#include <cmath>

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kSsimBlocks = 296;
constexpr int SW = 32, SH = 54, RAD = 5;
static_assert(SH + 2 * RAD == 64, "pass H decodes its item index with shifts");
constexpr int IW = SW + 2 * RAD, IH = SH + 2 * RAD;     // 42 x 64 staged samples
constexpr int VSEG = 7;                                 // outputs per thread in pass V (8 segments >= 54 rows)
constexpr int HP = SW + 1;                              // row pitch of the float64 planes (odd: see pass H)
constexpr int PH = IH + 2;                              // plane rows incl. two never-written rows that only
                                                        // discarded outputs of the last segment read
constexpr int kSsimSmem = 2 * IH * (IW + 1) * 4 + 4 * PH * HP * 8;

struct Taps { double w[2 * RAD + 1]; };

template <typename T>
__global__ void __launch_bounds__(256, 2)
ssim_gauss_kernel(const T* __restrict__ ref, const T* __restrict__ tst, int64_t band_stride, int64_t width,
                  int64_t r_lo, int64_t r_hi, int64_t buf_rows, Taps taps, double c1, double c2, double* scratch,
                  double* sum_acc, double* cnt_acc, void* workspace) {
  extern __shared__ __align__(16) unsigned char ssim_smem[];
  int (*xs)[IW + 1] = reinterpret_cast<int (*)[IW + 1]>(ssim_smem);
  int (*ys)[IW + 1] = reinterpret_cast<int (*)[IW + 1]>(ssim_smem + IH * (IW + 1) * 4);
  double (*hp)[PH][HP] = reinterpret_cast<double (*)[PH][HP]>(ssim_smem + 2 * IH * (IW + 1) * 4);   // [4][PH][HP]
  __shared__ double red[2][32];
  const int band = blockIdx.y;
  const T* A = ref + (int64_t)band * band_stride;
  const T* R = tst + (int64_t)band * band_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c_lo = RAD, c_hi = width - RAD;        // counted columns [c_lo, c_hi)
  const int64_t ncols = c_hi - c_lo, nrows = r_hi - r_lo;
  double acc = 0.0, cnt = 0.0;
  // the window is symmetric: six distinct taps, kept in registers (indices are compile-time after unrolling)
  double ws[RAD + 1];
#pragma unroll
  for (int k = 0; k <= RAD; ++k) ws[k] = taps.w[k];
#define DM_TAP(k) ws[(k) <= RAD ? (k) : 2 * RAD - (k)]
  if (ncols > 0 && nrows > 0) {
    const int64_t tiles_x = (ncols + SW - 1) / SW, tiles_y = (nrows + SH - 1) / SH;
    const int64_t ntiles = tiles_x * tiles_y;
    // Software pipeline over the tiles of this block: the (64 x 42) samples of both cubes of the NEXT tile are
    // fetched into registers (one ref/tst pair per register, eleven per thread) right after the current
    // tile has been staged, so that their DRAM latency passes during the two filter passes instead of in
    // front of them.  Element e = threadIdx.x + 256 k of the tile is row e / 42, column e % 42; clamped
    // indices are only reached by outputs that are discarded.
    constexpr int NPRE = (IH * IW + 255) / 256;          // 11
    uint32_t pre[NPRE];
    auto fetch_tile = [&](int64_t t) {
      const int64_t r0 = r_lo + (t / tiles_x) * SH, c0 = c_lo + (t % tiles_x) * SW;
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int e = threadIdx.x + 256 * k;
        const int lr = e / IW, lc = e - lr * IW;
        int64_t r = r0 + lr - RAD, c = c0 + lc - RAD;
        r = r < 0 ? 0 : (r >= buf_rows ? buf_rows - 1 : r);
        c = c < 0 ? 0 : (c >= width ? width - 1 : c);
        const uint32_t a = e < IH * IW ? (uint32_t)(uint16_t)A[r * width + c] : 0u;
        const uint32_t b = e < IH * IW ? (uint32_t)(uint16_t)R[r * width + c] : 0u;
        pre[k] = a | (b << 16);
      }
    };
    if ((int64_t)blockIdx.x < ntiles) fetch_tile(blockIdx.x);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int64_t r0 = r_lo + (t / tiles_x) * SH, c0 = c_lo + (t % tiles_x) * SW;
      const int rows_here = (int)(r_hi - r0 < SH ? r_hi - r0 : SH), cols_here = (int)(c_hi - c0 < SW ? c_hi - c0 : SW);
      __syncthreads();
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int e = threadIdx.x + 256 * k;
        if (e < IH * IW) {
          const int lr = e / IW, lc = e - lr * IW;
          // (T)(...) restores the sample's sign for int16 cubes
          xs[lr][lc] = (int)(T)(pre[k] & 0xffffu);
          ys[lr][lc] = (int)(T)(pre[k] >> 16);
        }
      }
      __syncthreads();
      if (t + gridDim.x < ntiles) fetch_tile(t + gridDim.x);
      // ---- pass H: item = (group of 4 output columns, row lr); 512 items, two per thread.  The ROW is
      // the fastest thread index: the lanes of a warp read the same columns of 32 rows (43 words apart)
      // and write the same columns of 32 plane rows (33 doubles apart) -- both strides odd, conflict free.
#pragma unroll 1
      for (int item = threadIdx.x; item < IH * (SW / 4); item += 256) {
        const int lr = item & (IH - 1), g4 = (item >> 6) * 4;
        double vx[14], vy[14], vq[14], vp[14];
#pragma unroll
        for (int k = 0; k < 14; ++k) {
          const double x = (double)xs[lr][g4 + k], y = (double)ys[lr][g4 + k];
          vx[k] = x; vy[k] = y; vq[k] = fma(x, x, y * y); vp[k] = x * y;
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          double hx = 0.0, hy = 0.0, hq = 0.0, hxy = 0.0;
#pragma unroll
          for (int k = 0; k <= 2 * RAD; ++k) {
            const double w = DM_TAP(k);
            hx = fma(w, vx[o + k], hx); hy = fma(w, vy[o + k], hy);
            hq = fma(w, vq[o + k], hq); hxy = fma(w, vp[o + k], hxy);
          }
          hp[0][lr][g4 + o] = hx; hp[1][lr][g4 + o] = hy; hp[2][lr][g4 + o] = hq; hp[3][lr][g4 + o] = hxy;
        }
      }
      __syncthreads();
      // ---- pass V: thread = (column tx, rows 7*ty .. 7*ty+6), plane by plane
      {
        const int l0 = ty * VSEG;
        double u[4][VSEG];
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
          double v[VSEG + 2 * RAD];
#pragma unroll
          for (int k = 0; k < VSEG + 2 * RAD; ++k) v[k] = hp[pl][l0 + k][tx];
#pragma unroll
          for (int o = 0; o < VSEG; ++o) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k <= 2 * RAD; ++k) s = fma(DM_TAP(k), v[o + k], s);
            u[pl][o] = s;
          }
        }
#pragma unroll
        for (int o = 0; o < VSEG; ++o) {
          const int lr = l0 + o;
          if (lr < rows_here && tx < cols_here) {
            const double ux = u[0][o], uy = u[1][o], uq = u[2][o], uxy = u[3][o];
            const double mm = ux * ux + uy * uy;
            const double vsum = uq - mm, vxy = uxy - ux * uy;
            const double num = (2.0 * ux * uy + c1) * (2.0 * vxy + c2);
            const double den = (mm + c1) * (vsum + c2);
            acc += num / den;
            cnt += 1.0;
          }
        }
      }
    }
  }
  acc = warp_sum_f64(acc); cnt = warp_sum_f64(cnt);
  if (tx == 0) { red[0][ty] = acc; red[1][ty] = cnt; }
  __syncthreads();
  double t[2] = {0.0, 0.0};
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) { t[0] += red[0][w]; t[1] += red[1][w]; }
  }
  __syncthreads();
  // the band's last block adds the band's block partials in a fixed order and accumulates {sum of S, count}
  double* const accs[2] = {sum_acc, cnt_acc};
  ordered_band_sum<2>(t, scratch, static_cast<Workspace*>(workspace)->band_counter, accs, &red[0][0]);
}

#undef DM_TAP

}  // namespace

int ssim_nblocks() { return kSsimBlocks; }

int launch_ssim_gauss(const dm_pair_t& p, double L, int64_t row_begin, int64_t row_end, int64_t img_row0,
                      int64_t img_rows, double* scratch, double* sum_acc, double* cnt_acc, void* workspace,
                      cudaStream_t s) {
  if (!p.ref || !p.tst || !scratch || !sum_acc || !cnt_acc || !workspace) return fail(DM_EARG, "dm_ssim_gauss: null pointer");
  if (p.layout != DM_BSQ) return fail(DM_EUNSUPPORTED, "dm_ssim_gauss: BSQ only (transpose with dm_bip_to_bsq)");
  if (p.bands <= 0 || p.bands > kMaxCounterBands || p.width <= 0) return fail(DM_EARG, "dm_ssim_gauss: bad geometry (1..2048 bands)");
  if (row_begin < 0 || row_end > p.rows || row_begin > row_end || img_row0 < 0 || img_row0 + p.rows > img_rows)
    return fail(DM_EARG, "dm_ssim_gauss: bad row range");
  // counted buffer rows: inside [row_begin,row_end) and inside the 5-px crop of the image
  int64_t r_lo = row_begin, r_hi = row_end;
  if (img_row0 + r_lo < RAD) r_lo = RAD - img_row0;
  if (img_row0 + r_hi > img_rows - RAD) r_hi = img_rows - RAD - img_row0;
  if (r_hi > r_lo && (r_lo - RAD < 0 || r_hi + RAD > p.rows))
    return fail(DM_EARG, "dm_ssim_gauss: strip lacks its 5 halo rows");
  Taps taps;
  double sum = 0.0;
  for (int k = -RAD; k <= RAD; ++k) { taps.w[k + RAD] = std::exp(-0.5 / (1.5 * 1.5) * (double)(k * k)); sum += taps.w[k + RAD]; }
  for (int k = 0; k <= 2 * RAD; ++k) taps.w[k] /= sum;
  const double c1 = (0.01 * L) * (0.01 * L), c2 = (0.03 * L) * (0.03 * L);
  const dim3 grid(kSsimBlocks, (unsigned)p.bands);
#define DM_SSIM(T)                                                                                           \
  do {                                                                                                       \
    DM_CUDA(cudaFuncSetAttribute(ssim_gauss_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSsimSmem)); \
    ssim_gauss_kernel<T><<<grid, 256, kSsimSmem, s>>>(static_cast<const T*>(p.ref), static_cast<const T*>(p.tst), \
                                                      p.band_stride, p.width, r_lo, r_hi, p.rows, taps, c1, c2, scratch, \
                                                      sum_acc, cnt_acc, workspace);                           \
  } while (0)
  switch (p.dtype) {
    case DM_U8: DM_SSIM(uint8_t); break;
    case DM_U16: DM_SSIM(uint16_t); break;
    case DM_I16: DM_SSIM(int16_t); break;
    default: return fail(DM_EARG, "dm_ssim_gauss: bad dtype");
  }
#undef DM_SSIM
  DM_LAUNCH_CHECK("ssim_gauss");
  return DM_OK;
}

}  // namespace dm
