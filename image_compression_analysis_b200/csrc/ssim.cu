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
// Shared-memory tiled, separable: stage a (16+10) x (32+10) tile of both cubes, horizontal pass into
// four float64 planes, vertical pass + SSIM formula + block-ordered partial sums.  FP64-FMA bound
// (~130 DFMA per pixel), not HBM bound; see DESIGN.md.

#include <cmath>

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kSsimBlocks = 296;
constexpr int SW = 32, SH = 16, RAD = 5;

struct Taps { double w[2 * RAD + 1]; };

template <typename T>
__global__ void __launch_bounds__(256)
ssim_gauss_kernel(const T* __restrict__ ref, const T* __restrict__ tst, int64_t band_stride, int64_t width,
                  int64_t r_lo, int64_t r_hi, int64_t buf_rows, Taps taps, double c1, double c2, double* out) {
  __shared__ int xs[SH + 2 * RAD][SW + 2 * RAD + 1];
  __shared__ int ys[SH + 2 * RAD][SW + 2 * RAD + 1];
  __shared__ double hp[4][SH + 2 * RAD][SW];
  __shared__ double red[2][8];
  const int band = blockIdx.y;
  const T* A = ref + (int64_t)band * band_stride;
  const T* R = tst + (int64_t)band * band_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c_lo = RAD, c_hi = width - RAD;        // counted columns [c_lo, c_hi)
  const int64_t ncols = c_hi - c_lo, nrows = r_hi - r_lo;
  double acc = 0.0, cnt = 0.0;
  if (ncols > 0 && nrows > 0) {
    const int64_t tiles_x = (ncols + SW - 1) / SW, tiles_y = (nrows + SH - 1) / SH;
    const int64_t ntiles = tiles_x * tiles_y;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int64_t r0 = r_lo + (t / tiles_x) * SH, c0 = c_lo + (t % tiles_x) * SW;
      __syncthreads();
      for (int i = threadIdx.x; i < (SH + 2 * RAD) * (SW + 2 * RAD); i += 256) {
        const int lr = i / (SW + 2 * RAD), lc = i - lr * (SW + 2 * RAD);
        int64_t r = r0 + lr - RAD, c = c0 + lc - RAD;
        r = r < 0 ? 0 : (r >= buf_rows ? buf_rows - 1 : r);     // only reached by discarded outputs
        c = c < 0 ? 0 : (c >= width ? width - 1 : c);
        xs[lr][lc] = (int)A[r * width + c];
        ys[lr][lc] = (int)R[r * width + c];
      }
      __syncthreads();
      // horizontal pass: (SH+10) rows x SW columns
      for (int i = threadIdx.x; i < (SH + 2 * RAD) * SW; i += 256) {
        const int lr = i / SW, lc = i - lr * SW;
        double hx = 0.0, hy = 0.0, hq = 0.0, hxy = 0.0;
#pragma unroll
        for (int k = 0; k <= 2 * RAD; ++k) {
          const double x = (double)xs[lr][lc + k], y = (double)ys[lr][lc + k], w = taps.w[k];
          hx = fma(w, x, hx); hy = fma(w, y, hy);
          hq = fma(w, fma(x, x, y * y), hq); hxy = fma(w, x * y, hxy);
        }
        hp[0][lr][lc] = hx; hp[1][lr][lc] = hy; hp[2][lr][lc] = hq; hp[3][lr][lc] = hxy;
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < SH; j += 8) {
        const int lr = ty + j;
        if (r0 + lr < r_hi && c0 + tx < c_hi) {
          double ux = 0.0, uy = 0.0, uq = 0.0, uxy = 0.0;
#pragma unroll
          for (int k = 0; k <= 2 * RAD; ++k) {
            const double w = taps.w[k];
            ux = fma(w, hp[0][lr + k][tx], ux); uy = fma(w, hp[1][lr + k][tx], uy);
            uq = fma(w, hp[2][lr + k][tx], uq); uxy = fma(w, hp[3][lr + k][tx], uxy);
          }
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
  acc = warp_sum_f64(acc); cnt = warp_sum_f64(cnt);
  if (tx == 0) { red[0][ty] = acc; red[1][ty] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    out[((int64_t)band * kSsimBlocks + blockIdx.x) * 2 + 0] = t0;
    out[((int64_t)band * kSsimBlocks + blockIdx.x) * 2 + 1] = t1;
  }
}

}  // namespace

int ssim_nblocks() { return kSsimBlocks; }

int launch_ssim_gauss(const dm_pair_t& p, double L, int64_t row_begin, int64_t row_end, int64_t img_row0,
                      int64_t img_rows, double* out, cudaStream_t s) {
  if (!p.ref || !p.tst || !out) return fail(DM_EARG, "dm_ssim_gauss: null pointer");
  if (p.layout != DM_BSQ) return fail(DM_EUNSUPPORTED, "dm_ssim_gauss: BSQ only (transpose with dm_bip_to_bsq)");
  if (p.bands <= 0 || p.bands > 65535 || p.width <= 0) return fail(DM_EARG, "dm_ssim_gauss: bad geometry");
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
  ssim_gauss_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(p.ref), static_cast<const T*>(p.tst),      \
                                            p.band_stride, p.width, r_lo, r_hi, p.rows, taps, c1, c2, out)
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
