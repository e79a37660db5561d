// Sobel LMSE: for every band, sum over pixels of (|grad ref| - |grad tst|)^2.
//
// Replaces sobel_mag + mse in the LMSE loop of compute_sam_sid_lmse_caseB
// (/root/reference/tools/run_codec.py:123-137, 341-346).  gx, gy are exact integers
// (|g| <= 4*65535), gx^2+gy^2 < 2^38 is exact in float64 and sqrt is correctly rounded, so every
// per-pixel term equals the reference's bit for bit; only the order of the final float64 sum
// differs (block-ordered partials, reduced on the host).
//
// BSQ only (the host transposes BIP cubes first).  Shared-memory tiled: a block stages a
// (32+2) x (32+2) tile of both cubes once and every sample is read from HBM ~1.13 times.

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kSobBlocks = 296;   // partial slots per band (fixed: layout must not depend on the device)
constexpr int TW = 32, TH = 32;

template <typename T>
__global__ void __launch_bounds__(256)
sobel_lmse_kernel(const T* __restrict__ ref, const T* __restrict__ tst, int64_t band_stride, int64_t width,
                  int64_t row_begin, int64_t row_end, int64_t img_row0, int64_t img_rows, double* out) {
  __shared__ int sa[TH + 2][TW + 2 + 1];
  __shared__ int sr[TH + 2][TW + 2 + 1];
  __shared__ double red[8];
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
        double mag[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int p00 = s[lr][tx], p01 = s[lr][tx + 1], p02 = s[lr][tx + 2];
          const int p10 = s[lr + 1][tx], p12 = s[lr + 1][tx + 2];
          const int p20 = s[lr + 2][tx], p21 = s[lr + 2][tx + 1], p22 = s[lr + 2][tx + 2];
          const int gx = (p00 - p02) + 2 * (p10 - p12) + (p20 - p22);      // |g| <= 4*65535: int32
          const int gy = (p00 - p20) + 2 * (p01 - p21) + (p02 - p22);
          const double fx = (double)gx, fy = (double)gy;                   // exact; fx^2 + fy^2 < 2^38 exact
          mag[q] = __dsqrt_rn(fma(fx, fx, fy * fy));
          s = sr;
        }
        const double e = __dsub_rn(mag[0], mag[1]);
        acc += __dmul_rn(e, e);
      }
    }
  }
  acc = warp_sum_f64(acc);
  if (tx == 0) red[ty] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[(int64_t)band * kSobBlocks + blockIdx.x] = t;
  }
}

}  // namespace

int sobel_nblocks() { return kSobBlocks; }

int launch_sobel(const dm_pair_t& p, int64_t row_begin, int64_t row_end, int64_t img_row0, int64_t img_rows,
                 double* out, cudaStream_t s) {
  if (!p.ref || !p.tst || !out) return fail(DM_EARG, "dm_sobel_lmse: null pointer");
  if (p.layout != DM_BSQ) return fail(DM_EUNSUPPORTED, "dm_sobel_lmse: BSQ only (transpose with dm_bip_to_bsq)");
  if (p.bands <= 0 || p.bands > 65535 || p.width <= 0) return fail(DM_EARG, "dm_sobel_lmse: bad geometry");
  if (row_begin < 0 || row_end > p.rows || row_begin > row_end || img_row0 < 0 || img_row0 + p.rows > img_rows)
    return fail(DM_EARG, "dm_sobel_lmse: bad row range");
  // halo rows must be present unless the strip touches the image border
  if ((row_begin == 0 && img_row0 > 0) || (row_end == p.rows && img_row0 + p.rows < img_rows))
    return fail(DM_EARG, "dm_sobel_lmse: strip lacks its halo row");
  const dim3 grid(kSobBlocks, (unsigned)p.bands);
#define DM_SOBEL(T)                                                                                          \
  sobel_lmse_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(p.ref), static_cast<const T*>(p.tst),      \
                                            p.band_stride, p.width, row_begin, row_end, img_row0, img_rows, out)
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
