// Shared host/device helpers for libdm_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "dm_b200.h"

namespace dm {

// ---- host side --------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);           // records dm_last_error(), returns code
int cuda_fail(cudaError_t e, const char* what);     // DM_ECUDA with the CUDA error text
int sm_count();                                     // SMs of the current device (cached), <0 on error
void count_launch();                                // one more kernel launched (dm_launch_count)
bool launch_chaining();                             // dm_launch_chaining() state of this thread
int fused_bip_variant();                            // dm_fused_bip_variant() state of this thread
int ssim_variant();                                 // dm_ssim_variant() state of this thread
int spectral_lanes_per_pixel();                     // dm_spectral_lanes_per_pixel() state of this thread

#define DM_CUDA(expr)                                              \
  do {                                                             \
    cudaError_t dm_e_ = (expr);                                    \
    if (dm_e_ != cudaSuccess) return ::dm::cuda_fail(dm_e_, #expr); \
  } while (0)

#define DM_LAUNCH_CHECK(name)                                       \
  do {                                                              \
    ::dm::count_launch();                                           \
    cudaError_t dm_e_ = cudaGetLastError();                         \
    if (dm_e_ != cudaSuccess) return ::dm::cuda_fail(dm_e_, name);  \
  } while (0)

static inline int elem_bytes(int dtype) { return dtype == DM_U8 ? 1 : 2; }

// Caller-provided device scratch (dm_workspace_bytes(), zeroed once by the caller): per-block float64
// partials and the arrival counter of the in-kernel ordered final reduction.  The last block of a
// launch resets the counter, so consecutive launches on one stream can share a workspace.
constexpr int kMaxPartialBlocks = 1184;
constexpr int kMaxCounterBands = 2048;     // per-band arrival counters of the (blocks, bands)-grid stencil kernels
constexpr int kMaxGroups = 64;             // block groups of the two-level reduction (BIP Sobel kernel)
struct Workspace {
  unsigned counter[16];                    // [0] spectral / one-pass kernels, [1] BIP Sobel kernel (final level)
  double part[3 * kMaxPartialBlocks];
  unsigned band_counter[kMaxCounterBands];
  unsigned group_counter[kMaxGroups];
};

// launchers implemented in the kernel translation units
int launch_fused_stats(const dm_pair_t& p, const uint8_t* plane, int plane_bit, int hist_bins,
                       uint32_t flags, int64_t* sums, int64_t* maxs, int64_t* hist, cudaStream_t s);
int launch_fused_stats_batch(const dm_pair_t& p, const dm_batch_item_t* items_dev, int n_items, uint32_t flags,
                             cudaStream_t s);
int launch_validity(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out,
                    int64_t* counts, cudaStream_t s);
int launch_spectral(const dm_pair_t& p, const uint8_t* plane, uint16_t* errmax_out,
                    const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                    const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z,
                    int want_sam, int want_sid, double* spectral_acc, void* workspace, cudaStream_t s);
int launch_fused_bip(const dm_pair_t& p, const uint8_t* plane, int64_t* sums, int64_t* maxs,
                     uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                     const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                     double* spectral_acc, void* workspace, cudaStream_t s);
int launch_fused_bip_scan(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts,
                          int64_t* sums, int64_t* maxs,
                          uint16_t* errmax_out, const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g,
                          const uint8_t* lut_z, int cap_z, uint8_t* err8_z, int64_t* hist8_z, int want_sam,
                          double* spectral_acc, void* workspace, cudaStream_t s);
int64_t launch_validity_ct(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts,
                           cudaStream_t s, int* status);
int launch_fused_bsq(const dm_pair_t& p, const uint8_t* plane, int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                     const uint8_t* lut_g, int cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z,
                     int cap_z, uint8_t* err8_z, int64_t* hist8_z, cudaStream_t s);
int launch_sobel_mag(const void* img, int dtype, int64_t rows, int64_t width, double* out, cudaStream_t s);
int sobel_nblocks();
int ssim_nblocks();
int launch_sobel(const dm_pair_t& p, int64_t row_begin, int64_t row_end, int64_t img_row0,
                 int64_t img_rows, double* scratch, double* lmse_acc, void* workspace, cudaStream_t s);
int launch_ssim_gauss(const dm_pair_t& p, double L, int64_t row_begin, int64_t row_end,
                      int64_t img_row0, int64_t img_rows, double* scratch, double* sum_acc, double* cnt_acc,
                      void* workspace, cudaStream_t s);
int launch_combine_partials(const void* gathered, int world, int64_t records, int64_t n_sum, int64_t n_max,
                            int64_t n_f64, void* out, cudaStream_t s);
int launch_bip_to_bsq(const void* src, void* dst, int elem_bytes, int64_t bands, int64_t rows,
                      int64_t width, cudaStream_t s);

int launch_band_hist(const dm_cube_t& c, const int32_t* sel, int nsel, const uint8_t* plane, int plane_bit,
                     int64_t* hist, cudaStream_t s);
int launch_lut_bands(const dm_cube_t& c, const int32_t* sel, int nsel, const uint8_t* luts, uint8_t* out, cudaStream_t s);
int launch_requantize(const void* src, void* dst, int dtype, int64_t n, int mode, int k, int has_nodata, int nodata,
                      cudaStream_t s);
int launch_scene_error(const dm_pair_t& p, const uint8_t* valid, int mode, int k_bits, uint32_t p95_thr, float* out_plane,
                       uint32_t* out_max_bits, cudaStream_t s);
int launch_scale_plane_u8(const float* plane, int64_t n, float emax, float scale, uint8_t* out, cudaStream_t s);
int launch_diff1(const void* src, void* dst, int dtype, int arith, int inverse, int64_t bands, int64_t npix,
                 int64_t band_stride, cudaStream_t s);
int launch_interleave(const void* src, void* dst, int eb, int from, int to, int64_t bands, int64_t rows, int64_t width,
                      cudaStream_t s);

int p2p_alloc(int64_t bytes, void** ptr, void* handle64);
int p2p_open(const void* handle64, void** ptr);
int p2p_close(void* ptr);
int p2p_free(void* ptr);
int p2p_zero(void* ptr, int64_t bytes, cudaStream_t s);
int launch_p2p_push(const void* src, int64_t total_words, void* const* peer_dst, void* const* peer_flag, int world,
                    uint64_t flag_value, cudaStream_t s);
int launch_p2p_combine(const void* gathered, const void* flags, int world, uint64_t need, int64_t capacity, int64_t rec0,
                       int64_t nrec, int64_t n_sum, int64_t n_max, int64_t n_f64, void* out, uint32_t* status,
                       double timeout_s, cudaStream_t s);

// ---- device side ------------------------------------------------------------------------------
#ifdef __CUDACC__

// streaming 128/64/32-bit loads: read-only path, do not allocate in L1 (data is touched once)
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// sample decode: element j (0/1) of a 32-bit word holding two 16-bit samples
template <int DT>
__device__ __forceinline__ int sample16(uint32_t w, int j) {
  if (DT == DM_I16) return j ? ((int)w >> 16) : ((int)(w << 16) >> 16);
  return j ? (int)(w >> 16) : (int)(w & 0xffffu);
}

template <int DT, typename T>
__device__ __forceinline__ int sample_at(const T* p, int64_t i) {
  return (int)p[i];
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ long long warp_max_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reductions through a small shared scratch (>= 32 x 8 bytes).  All threads call.
// Result valid in thread 0.  blockDim.x may be any size <= 1024 (partial last warp allowed
// only if every thread of the block participates, which __shfl with full mask requires:
// callers keep blockDim.x a multiple of 32).
__device__ __forceinline__ long long block_sum_ll(long long v, long long* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_ll(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? scratch[lane] : 0;
    v = warp_sum_ll(v);
  }
  return v;
}
__device__ __forceinline__ long long block_max_ll(long long v, long long* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max_ll(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? scratch[lane] : (long long)0x8000000000000000ll;
    v = warp_max_ll(v);
  }
  return v;
}
__device__ __forceinline__ double block_sum_f64(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum_f64(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < nw ? scratch[lane] : 0.0;
    v = warp_sum_f64(v);
  }
  return v;
}

// Ordered final reduction of per-block float64 partials {t0,t1,t2} (valid in lane 0).  Called by ONE
// full warp of every block at the end of the kernel.  Each block stores its partials; the block that
// arrives last adds them up in a fixed order (lane l takes blocks l, l+32, ...; then the shuffle tree)
// and ACCUMULATES the result into acc[0..2], so the sum is bit-identical from run to run for a given
// grid size, and a tail launch on the same stream composes with the main one.
__device__ __forceinline__ void ordered_block_sum3(double t0, double t1, double t2, void* workspace, double* acc) {
  Workspace* ws = static_cast<Workspace*>(workspace);
  const int lane = threadIdx.x & 31;
  unsigned last = 0;
  if (lane == 0) {
    double* d = ws->part + 3 * (size_t)blockIdx.x;
    __stcg(d + 0, t0); __stcg(d + 1, t1); __stcg(d + 2, t2);
    __threadfence();
    last = atomicAdd(&ws->counter[0], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (unsigned i = lane; i < gridDim.x; i += 32) {
    const double* d = ws->part + 3 * (size_t)i;
    a0 += __ldcg(d + 0); a1 += __ldcg(d + 1); a2 += __ldcg(d + 2);
  }
  a0 = warp_sum_f64(a0); a1 = warp_sum_f64(a1); a2 = warp_sum_f64(a2);
  if (lane == 0) {
    acc[0] += a0; acc[1] += a1; acc[2] += a2;
    ws->counter[0] = 0;
  }
}

// Ordered final reduction for kernels whose grid is (blocks, bands) and whose blocks each hold NV float64
// partials of ONE band (valid in thread 0).  Block (x, band) stores them at scratch[(band * gridDim.x + x) * NV + v];
// the block of a band that arrives last adds that band's gridDim.x partials in a fixed order (thread i takes
// entries i, i + blockDim.x, ...; shuffle tree; warps in order) and ACCUMULATES into acc[v][band].  The result does
// not depend on which block is last, every band is reduced by a different block (the tail is one short pass over
// gridDim.x values, not over the whole grid), and no follow-up reduction kernel is needed.  All threads call;
// `red` is shared scratch of >= NV * 32 doubles.  counter: Workspace::band_counter (zero on entry, reset here).
template <int NV>
__device__ __forceinline__ void ordered_band_sum(const double (&t)[NV], double* scratch, unsigned* counter,
                                                 double* const (&acc)[NV], double* red) {
  __shared__ unsigned s_last;
  const int band = blockIdx.y;
  double* mine = scratch + ((size_t)band * gridDim.x + blockIdx.x) * NV;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) __stcg(mine + v, t[v]);
    __threadfence();
    s_last = atomicAdd(&counter[band], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double* base = scratch + (size_t)band * gridDim.x * NV;
  double a[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) a[v] = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
#pragma unroll
    for (int v = 0; v < NV; ++v) a[v] += __ldcg(base + (size_t)i * NV + v);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    a[v] = warp_sum_f64(a[v]);
    if (lane == 0) red[v * 32 + warp] = a[v];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double tot = 0.0;
      for (int w = 0; w < nw; ++w) tot += red[v * 32 + w];
      acc[v][band] += tot;
    }
    counter[band] = 0;
  }
}

__device__ __forceinline__ void atomic_add_i64(int64_t* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}
__device__ __forceinline__ void atomic_max_i64(int64_t* p, long long v) {
  atomicMax(reinterpret_cast<long long*>(p), v);
}

#endif  // __CUDACC__

}  // namespace dm
