// libdm_b200.so: C-ABI entry points (include/dm_b200.h), error reporting and device queries.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "dm_common.cuh"

namespace dm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local bool g_chaining = false;
bool launch_chaining() { return g_chaining; }
void set_launch_chaining(bool on) { g_chaining = on; }
static thread_local int g_bip_variant = 0;
int fused_bip_variant() { return g_bip_variant; }
static thread_local int g_ssim_variant = 0;
int ssim_variant() { return g_ssim_variant; }
static thread_local int g_spec_lpp = 0;
int spectral_lanes_per_pixel() { return g_spec_lpp; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(DM_ECUDA, "CUDA error in %s: %s", what, cudaGetErrorString(e));
}

void set_launch_chaining(bool on);

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -cuda_fail(e, "cudaGetDevice");
  if (dev != cached_dev) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return -cuda_fail(e, "cudaDeviceGetAttribute");
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace dm

using namespace dm;

extern "C" {

int dm_abi_version(void) { return DM_ABI_VERSION; }
void dm_launch_chaining(int32_t on) { dm::set_launch_chaining(on != 0); }
int dm_fused_bip_variant(int32_t v) {
  if (v != 0 && v != 1 && v != 12 && v != 23) return fail(DM_EARG, "dm_fused_bip_variant: 0 (auto), 1 (run-time geometry), 12 or 23 band warps");
  dm::g_bip_variant = v;
  return DM_OK;
}
const char* dm_last_error(void) { return g_err; }
int dm_device_sm_count(void) { return sm_count(); }
int64_t dm_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

int dm_validity(const dm_pair_t* p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts_out,
                void* stream) {
  if (!p) return fail(DM_EARG, "dm_validity: null pair");
  return launch_validity(*p, valid_in, plane_out, counts_out, static_cast<cudaStream_t>(stream));
}

int dm_fused_stats(const dm_pair_t* p, const uint8_t* plane, int32_t plane_bit, int32_t hist_bins,
                   uint32_t flags, int64_t* sums, int64_t* maxs, int64_t* hist, void* stream) {
  if (!p) return fail(DM_EARG, "dm_fused_stats: null pair");
  return launch_fused_stats(*p, plane, plane_bit, hist_bins, flags, sums, maxs, hist,
                            static_cast<cudaStream_t>(stream));
}

int dm_fused_stats_batch(const dm_pair_t* geom, const dm_batch_item_t* items_dev, int32_t n_items, uint32_t flags,
                         void* stream) {
  if (!geom) return fail(DM_EARG, "dm_fused_stats_batch: null geometry");
  return launch_fused_stats_batch(*geom, items_dev, n_items, flags, static_cast<cudaStream_t>(stream));
}

int64_t dm_workspace_bytes(void) { return (int64_t)sizeof(Workspace); }

int dm_spectral(const dm_pair_t* p, const uint8_t* plane, uint16_t* errmax_out, const uint8_t* lut_g,
                int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z, int32_t cap_z,
                uint8_t* err8_z, int64_t* hist8_z, int32_t want_sam, int32_t want_sid,
                double* spectral_acc, void* workspace, void* stream) {
  if (!p) return fail(DM_EARG, "dm_spectral: null pair");
  return launch_spectral(*p, plane, errmax_out, lut_g, cap_g, err8_g, hist8_g, lut_z, cap_z, err8_z,
                         hist8_z, want_sam, want_sid, spectral_acc, workspace, static_cast<cudaStream_t>(stream));
}

int dm_fused_bip(const dm_pair_t* p, const uint8_t* plane, int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                 const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z,
                 int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z, int32_t want_sam, double* spectral_acc,
                 void* workspace, void* stream) {
  if (!p) return fail(DM_EARG, "dm_fused_bip: null pair");
  return launch_fused_bip(*p, plane, sums, maxs, errmax_out, lut_g, cap_g, err8_g, hist8_g, lut_z, cap_z, err8_z,
                          hist8_z, want_sam, spectral_acc, workspace, static_cast<cudaStream_t>(stream));
}

int dm_fused_bip_scan(const dm_pair_t* p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts_out,
                      int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                      const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z,
                      int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z, int32_t want_sam, double* spectral_acc,
                      void* workspace, void* stream) {
  if (!p) return fail(DM_EARG, "dm_fused_bip_scan: null pair");
  return launch_fused_bip_scan(*p, valid_in, plane_out, counts_out, sums, maxs, errmax_out, lut_g, cap_g, err8_g, hist8_g,
                               lut_z, cap_z, err8_z, hist8_z, want_sam, spectral_acc, workspace,
                               static_cast<cudaStream_t>(stream));
}

int dm_fused_bsq(const dm_pair_t* p, const uint8_t* plane, int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                 const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g, const uint8_t* lut_z,
                 int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z, void* stream) {
  if (!p) return fail(DM_EARG, "dm_fused_bsq: null pair");
  return launch_fused_bsq(*p, plane, sums, maxs, errmax_out, lut_g, cap_g, err8_g, hist8_g, lut_z, cap_z, err8_z,
                          hist8_z, static_cast<cudaStream_t>(stream));
}

int dm_spectral_lanes_per_pixel(int32_t v) {
  if (v != 0 && v != 8 && v != 16 && v != 32) return fail(DM_EARG, "dm_spectral_lanes_per_pixel: 0 (auto), 8, 16 or 32");
  dm::g_spec_lpp = v;
  return DM_OK;
}

int dm_ssim_variant(int32_t v) {
  if (v < 0 || v > 3) return fail(DM_EARG, "dm_ssim_variant: 0 (auto), 1 (integer streaming kernel), 2 (tiled kernel), 3 (ring kernel)");
  dm::g_ssim_variant = v;
  return DM_OK;
}

int dm_sobel_mag(const void* img, int32_t dtype, int64_t rows, int64_t width, double* out, void* stream) {
  return launch_sobel_mag(img, dtype, rows, width, out, static_cast<cudaStream_t>(stream));
}

int dm_sobel_nblocks(void) { return sobel_nblocks(); }

int dm_sobel_lmse(const dm_pair_t* p, int64_t row_begin, int64_t row_end, int64_t img_row0,
                  int64_t img_rows, double* scratch, double* lmse_acc, void* workspace, void* stream) {
  if (!p) return fail(DM_EARG, "dm_sobel_lmse: null pair");
  return launch_sobel(*p, row_begin, row_end, img_row0, img_rows, scratch, lmse_acc, workspace,
                      static_cast<cudaStream_t>(stream));
}

int dm_ssim_nblocks(void) { return ssim_nblocks(); }

int dm_ssim_gauss(const dm_pair_t* p, double data_range, int64_t row_begin, int64_t row_end,
                  int64_t img_row0, int64_t img_rows, double* scratch, double* sum_acc, double* cnt_acc,
                  void* workspace, void* stream) {
  if (!p) return fail(DM_EARG, "dm_ssim_gauss: null pair");
  return launch_ssim_gauss(*p, data_range, row_begin, row_end, img_row0, img_rows, scratch, sum_acc, cnt_acc,
                           workspace, static_cast<cudaStream_t>(stream));
}

int dm_combine_partials(const void* gathered, int32_t world, int64_t records, int64_t n_sum, int64_t n_max, int64_t n_f64,
                        void* out, void* stream) {
  return launch_combine_partials(gathered, world, records, n_sum, n_max, n_f64, out, static_cast<cudaStream_t>(stream));
}

int dm_bip_to_bsq(const void* src, void* dst, int32_t elem_bytes, int64_t bands, int64_t rows,
                  int64_t width, void* stream) {
  return launch_bip_to_bsq(src, dst, elem_bytes, bands, rows, width, static_cast<cudaStream_t>(stream));
}

int dm_band_hist(const dm_cube_t* c, const int32_t* sel_bands, int32_t nsel, const uint8_t* plane, int32_t plane_bit,
                 int64_t* hist, void* stream) {
  if (!c) return fail(DM_EARG, "dm_band_hist: null cube");
  return launch_band_hist(*c, sel_bands, nsel, plane, plane_bit, hist, static_cast<cudaStream_t>(stream));
}

int dm_lut_bands_u8(const dm_cube_t* c, const int32_t* sel_bands, int32_t nsel, const uint8_t* luts, uint8_t* out,
                    void* stream) {
  if (!c) return fail(DM_EARG, "dm_lut_bands_u8: null cube");
  return launch_lut_bands(*c, sel_bands, nsel, luts, out, static_cast<cudaStream_t>(stream));
}

int dm_requantize(const void* src, void* dst, int32_t dtype, int64_t n, int32_t mode, int32_t k, int32_t has_nodata,
                  int32_t nodata, void* stream) {
  return launch_requantize(src, dst, dtype, n, mode, k, has_nodata, nodata, static_cast<cudaStream_t>(stream));
}

int dm_scene_error(const dm_pair_t* p, const uint8_t* valid, int32_t mode, int32_t k_bits, uint32_t p95_thr,
                   float* out_plane, uint32_t* out_max_bits, void* stream) {
  if (!p) return fail(DM_EARG, "dm_scene_error: null pair");
  return launch_scene_error(*p, valid, mode, k_bits, p95_thr, out_plane, out_max_bits, static_cast<cudaStream_t>(stream));
}

int dm_scale_plane_u8(const float* plane, int64_t n, float emax, float scale, uint8_t* out, void* stream) {
  return launch_scale_plane_u8(plane, n, emax, scale, out, static_cast<cudaStream_t>(stream));
}

int dm_diff1(const void* src, void* dst, int32_t dtype, int32_t arith, int32_t inverse, int64_t bands, int64_t npix,
             int64_t band_stride, void* stream) {
  return launch_diff1(src, dst, dtype, arith, inverse, bands, npix, band_stride, static_cast<cudaStream_t>(stream));
}

int dm_interleave(const void* src, void* dst, int32_t elem_bytes, int32_t from_layout, int32_t to_layout, int64_t bands,
                  int64_t rows, int64_t width, void* stream) {
  return launch_interleave(src, dst, elem_bytes, from_layout, to_layout, bands, rows, width, static_cast<cudaStream_t>(stream));
}

int dm_p2p_alloc(int64_t bytes, void** ptr, void* handle64) { return p2p_alloc(bytes, ptr, handle64); }
int dm_p2p_open(const void* handle64, void** ptr) { return p2p_open(handle64, ptr); }
int dm_p2p_close(void* ptr) { return p2p_close(ptr); }
int dm_p2p_free(void* ptr) { return p2p_free(ptr); }
int dm_p2p_zero(void* ptr, int64_t bytes, void* stream) { return p2p_zero(ptr, bytes, static_cast<cudaStream_t>(stream)); }

int dm_p2p_push(const void* src, int64_t total_words, void* const* peer_dst, void* const* peer_flag, int32_t world,
                uint64_t flag_value, void* stream) {
  return launch_p2p_push(src, total_words, peer_dst, peer_flag, world, flag_value, static_cast<cudaStream_t>(stream));
}

int dm_p2p_combine(const void* gathered, const void* flags, int32_t world, uint64_t need, int64_t capacity, int64_t rec0,
                   int64_t nrec, int64_t n_sum, int64_t n_max, int64_t n_f64, void* out, uint32_t* status, double timeout_s,
                   void* stream) {
  return launch_p2p_combine(gathered, flags, world, need, capacity, rec0, nrec, n_sum, n_max, n_f64, out, status, timeout_s,
                            static_cast<cudaStream_t>(stream));
}

}  // extern "C"
