// Per-pixel validity plane for the three masks of the distortion path.
//
//   DM_VALID_METRICS    /root/reference/tools/run_codec.py:249-263   (compute_metrics)
//   DM_VALID_QUICKLOOK  /root/reference/tools/quicklooks.py:35-45    (write_error_max8)
//   DM_VALID_SPECTRAL   /root/reference/tools/run_codec.py:314-319   (compute_sam_sid_lmse_caseB)
//
// rasterio's dataset_mask() (third party) is: valid where ANY band differs from nodata, all valid
// when the dataset has no nodata.  The nodata tests need every band of both cubes before any band
// can be reduced, hence this pre-pass; it is only launched when a nodata value or a caller mask
// exists (otherwise the metric kernels run with plane == NULL).

#include "dm_common.cuh"

namespace dm {

namespace {

template <typename T>
__device__ __forceinline__ void scan_pixel_bsq(const T* cube, int64_t pix, int64_t bands, int64_t stride,
                                               int nodata, bool& any_ne, bool& all_ne, bool& b1_ne) {
  any_ne = false; all_ne = true; b1_ne = true;
  for (int64_t b = 0; b < bands; ++b) {
    const bool ne = (int)cube[b * stride + pix] != nodata;
    any_ne |= ne; all_ne &= ne;
    if (b == 0) b1_ne = ne;
  }
}

// BSQ: one thread per pixel (coalesced along the row for every band)
template <typename T>
__global__ void __launch_bounds__(256)
validity_bsq(const T* ref, const T* tst, int64_t bands, int64_t npix, int64_t stride, int ref_has, int ref_nd,
             int tst_has, int tst_nd, const uint8_t* valid_in, uint8_t* plane, int64_t* counts) {
  int c0 = 0, c1 = 0, c2 = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (int64_t)gridDim.x * blockDim.x) {
    bool ra = true, rl = true, r1 = true, ta = true, tl = true, t1 = true;
    if (ref_has) scan_pixel_bsq(ref, p, bands, stride, ref_nd, ra, rl, r1);
    if (tst_has) scan_pixel_bsq(tst, p, bands, stride, tst_nd, ta, tl, t1);
    const bool vin = valid_in ? valid_in[p] != 0 : true;
    const bool ds = ra && ta;
    uint8_t v = 0;
    if (ds && rl && tl && vin) v |= DM_VALID_METRICS;
    if (ds && r1 && t1) v |= DM_VALID_QUICKLOOK;
    if (valid_in ? vin : ds) v |= DM_VALID_SPECTRAL;
    plane[p] = v;
    c0 += (v & DM_VALID_METRICS) ? 1 : 0;
    c1 += (v & DM_VALID_QUICKLOOK) ? 1 : 0;
    c2 += (v & DM_VALID_SPECTRAL) ? 1 : 0;
  }
  long long s0 = warp_sum_ll(c0), s1 = warp_sum_ll(c1), s2 = warp_sum_ll(c2);
  if ((threadIdx.x & 31) == 0 && counts) {
    if (s0) atomic_add_i64(counts + 0, s0);
    if (s1) atomic_add_i64(counts + 1, s1);
    if (s2) atomic_add_i64(counts + 2, s2);
  }
}

// BIP: one warp per pixel, lanes stride over the (contiguous) spectrum
template <typename T>
__global__ void __launch_bounds__(256)
validity_bip(const T* ref, const T* tst, int64_t bands, int64_t npix, int ref_has, int ref_nd, int tst_has,
             int tst_nd, const uint8_t* valid_in, uint8_t* plane, int64_t* counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int c0 = 0, c1 = 0, c2 = 0;
  for (int64_t p = warp; p < npix; p += nwarps) {
    bool ra = false, rl = true, ta = false, tl = true, r1 = true, t1 = true;
    for (int64_t b = lane; b < bands; b += 32) {
      if (ref_has) { const bool ne = (int)ref[p * bands + b] != ref_nd; ra |= ne; rl &= ne; if (b == 0) r1 = ne; }
      if (tst_has) { const bool ne = (int)tst[p * bands + b] != tst_nd; ta |= ne; tl &= ne; if (b == 0) t1 = ne; }
    }
    const unsigned full = 0xffffffffu;
    ra = ref_has ? __any_sync(full, ra) : true;
    ta = tst_has ? __any_sync(full, ta) : true;
    rl = __all_sync(full, rl); tl = __all_sync(full, tl);
    r1 = __all_sync(full, r1); t1 = __all_sync(full, t1);
    if (lane == 0) {
      const bool vin = valid_in ? valid_in[p] != 0 : true;
      const bool ds = ra && ta;
      uint8_t v = 0;
      if (ds && rl && tl && vin) v |= DM_VALID_METRICS;
      if (ds && r1 && t1) v |= DM_VALID_QUICKLOOK;
      if (valid_in ? vin : ds) v |= DM_VALID_SPECTRAL;
      plane[p] = v;
      c0 += (v & DM_VALID_METRICS) ? 1 : 0;
      c1 += (v & DM_VALID_QUICKLOOK) ? 1 : 0;
      c2 += (v & DM_VALID_SPECTRAL) ? 1 : 0;
    }
  }
  if (lane == 0 && counts) {
    if (c0) atomic_add_i64(counts + 0, c0);
    if (c1) atomic_add_i64(counts + 1, c1);
    if (c2) atomic_add_i64(counts + 2, c2);
  }
}

template <typename T>
int run_validity(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane, int64_t* counts, cudaStream_t s) {
  const int64_t npix = p.rows * p.width;
  const int sms = sm_count();
  if (sms < 0) return DM_ECUDA;
  const T* ref = static_cast<const T*>(p.ref);
  const T* tst = static_cast<const T*>(p.tst);
  if (p.layout == DM_BSQ) {
    int64_t grid = (npix + 255) / 256;
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    validity_bsq<T><<<(unsigned)grid, 256, 0, s>>>(ref, tst, p.bands, npix, p.band_stride, p.ref_has_nodata,
                                                    p.ref_nodata, p.tst_has_nodata, p.tst_nodata, valid_in, plane, counts);
  } else {
    int64_t grid = (npix + 7) / 8;
    if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
    validity_bip<T><<<(unsigned)grid, 256, 0, s>>>(ref, tst, p.bands, npix, p.ref_has_nodata, p.ref_nodata,
                                                    p.tst_has_nodata, p.tst_nodata, valid_in, plane, counts);
  }
  DM_LAUNCH_CHECK("validity");
  return DM_OK;
}

}  // namespace

int launch_validity(const dm_pair_t& p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts,
                    cudaStream_t s) {
  if (!p.ref || !p.tst || !plane_out) return fail(DM_EARG, "dm_validity: null pointer");
  if (p.bands <= 0 || p.rows < 0 || p.width < 0) return fail(DM_EARG, "dm_validity: bad geometry");
  if (p.layout != DM_BSQ && p.layout != DM_BIP) return fail(DM_EARG, "dm_validity: bad layout");
  if (p.rows * p.width == 0) return DM_OK;
  // EnMAP BIP cubes: full 64-pixel tiles through the TMA-staged kernel, the rest through the generic one
  dm_pair_t q = p;
  {
    int st = DM_OK;
    const int64_t done = launch_validity_ct(p, valid_in, plane_out, counts, s, &st);
    if (st != DM_OK) return st;
    if (done == p.rows * p.width) return DM_OK;
    if (done > 0) {
      const int64_t eb = 2;
      q.ref = static_cast<const char*>(p.ref) + done * p.bands * eb;
      q.tst = static_cast<const char*>(p.tst) + done * p.bands * eb;
      q.rows = 1; q.width = p.rows * p.width - done;
      if (valid_in) valid_in += done;
      plane_out += done;
    }
  }
  switch (q.dtype) {
    case DM_U8: return run_validity<uint8_t>(q, valid_in, plane_out, counts, s);
    case DM_U16: return run_validity<uint16_t>(q, valid_in, plane_out, counts, s);
    case DM_I16: return run_validity<int16_t>(q, valid_in, plane_out, counts, s);
  }
  return fail(DM_EARG, "dm_validity: bad dtype");
}

}  // namespace dm
