// Layout helper dm_bip_to_bsq ((rows, width, bands) -> (bands, rows, width); the transpose itself lives in
// adjacent.cu) and the multi-GPU combine of partial vectors.

#include "dm_common.cuh"

namespace dm {

namespace {

// world gathered runs of `records` partial vectors -> one run (see dm_combine_partials in dm_b200.h)
__global__ void __launch_bounds__(256)
combine_partials_kernel(const long long* __restrict__ gathered, int world, int64_t records, int64_t n_sum, int64_t n_max,
                        int64_t n_f64, long long* __restrict__ out) {
  const int64_t len = n_sum + n_max + n_f64, total = records * len;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t off = i % len;
    if (off < n_sum) {
      long long v = 0;
      for (int r = 0; r < world; ++r) v += gathered[(int64_t)r * total + i];
      out[i] = v;
    } else if (off < n_sum + n_max) {
      long long v = gathered[i];
      for (int r = 1; r < world; ++r) v = max(v, gathered[(int64_t)r * total + i]);
      out[i] = v;
    } else {
      double v = 0.0;
      for (int r = 0; r < world; ++r) v += __longlong_as_double(gathered[(int64_t)r * total + i]);   // rank order
      out[i] = __double_as_longlong(v);
    }
  }
}

}  // namespace

int launch_combine_partials(const void* gathered, int world, int64_t records, int64_t n_sum, int64_t n_max, int64_t n_f64,
                            void* out, cudaStream_t s) {
  if (!gathered || !out) return fail(DM_EARG, "dm_combine_partials: null pointer");
  if (world < 1 || records < 0 || n_sum < 0 || n_max < 0 || n_f64 < 0) return fail(DM_EARG, "dm_combine_partials: bad sizes");
  const int64_t total = records * (n_sum + n_max + n_f64);
  if (total == 0) return DM_OK;
  int64_t grid = (total + 255) / 256;
  if (grid > 1024) grid = 1024;
  combine_partials_kernel<<<(unsigned)grid, 256, 0, s>>>(static_cast<const long long*>(gathered), world, records, n_sum,
                                                        n_max, n_f64, static_cast<long long*>(out));
  DM_LAUNCH_CHECK("combine_partials");
  return DM_OK;
}

int launch_bip_to_bsq(const void* src, void* dst, int elem_bytes, int64_t bands, int64_t rows, int64_t width,
                      cudaStream_t s) {
  if (!src || !dst) return fail(DM_EARG, "dm_bip_to_bsq: null pointer");
  if (elem_bytes != 1 && elem_bytes != 2) return fail(DM_EARG, "dm_bip_to_bsq: elem_bytes must be 1 or 2");
  const int64_t npix = rows * width;
  if (npix <= 0 || bands <= 0) return DM_OK;
  // the batched 64x64 transpose of dm_interleave (32-bit accesses, tiles ordered along the band axis)
  return launch_interleave(src, dst, elem_bytes, DM_BIP, DM_BSQ, bands, rows, width, s);
}

}  // namespace dm
