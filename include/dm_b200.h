/*
 * dm_b200.h -- C ABI of libdm_b200.so: B200 (sm_100a) reconstruction-distortion kernels.
 *
 * This is the drop-in boundary for the one hot path of Angela0110/Image-compression-analysis:
 * the metric evaluation tools/run_codec.py runs after every codec decode.  The reference is
 * pure Python/numpy and has no FFI of its own; each entry point below names the reference
 * function (file:line under /root/reference) whose ARITHMETIC it replaces.  The reference-side
 * binding is ctypes; the host mirror of the reference's Python signatures
 * lives in image_compression_analysis_b200/metrics.py and quicklooks.py.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer owned by the caller unless the name says `host`;
 *    the library allocates nothing persistent; `stream` is a cudaStream_t passed as void*
 *    (0 = legacy default stream);
 *  - every call is asynchronous on `stream`; integer outputs ACCUMULATE into caller-zeroed
 *    buffers (sums add, maxima max), so row strips / band groups / GPUs compose exactly;
 *  - return value: 0 = ok, otherwise a DM_E* code; dm_last_error() gives the text (thread-local);
 *  - there is no CPU path: without a CUDA device every compute entry point fails with DM_ECUDA.
 *
 * Cube geometry (dm_pair_t): `layout` DM_BSQ = (bands, rows, width) with `band_stride` elements
 * between bands (rows contiguous), DM_BIP = (rows, width, bands) contiguous.  `rows` is the number
 * of image rows present in the buffer; stencil kernels additionally take the image row index of
 * buffer row 0 and the full image height so that row strips with halos shard across GPUs.
 */
#ifndef DM_B200_H
#define DM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_ABI_VERSION 9

/* status codes */
enum { DM_OK = 0, DM_EARG = 1, DM_ECUDA = 2, DM_EUNSUPPORTED = 3 };

/* sample types (the reference path sees uint8 / uint16 / int16: run_codec.py:89-113) */
enum { DM_U8 = 0, DM_U16 = 1, DM_I16 = 2 };

/* in-memory interleave (DM_BIL, (rows, bands, width), only appears in dm_interleave) */
enum { DM_BSQ = 0, DM_BIP = 1, DM_BIL = 2 };

/* bits of the per-pixel validity plane written by dm_validity() */
enum {
  DM_VALID_METRICS = 1,   /* run_codec.py:249-263  compute_metrics mask            */
  DM_VALID_QUICKLOOK = 2, /* quicklooks.py:35-45,128-130  write_error_max8 mask   */
  DM_VALID_SPECTRAL = 4   /* run_codec.py:314-319  compute_sam_sid_lmse_caseB mask */
};

/* per-band integer partials written by dm_fused_stats(): two int64 arrays of DM_NSTAT per band */
enum { DM_NSTAT = 8 };
/* sums[band*DM_NSTAT + k], combined across shards with SUM */
enum { DM_S_N = 0, DM_S_X = 1, DM_S_Y = 2, DM_S_XX = 3, DM_S_YY = 4, DM_S_XY = 5, DM_S_ABS = 6, DM_S_SSE = 7 };
/* maxs[band*DM_NSTAT + k], combined across shards with MAX.  MAXERR is per band; the others are
 * cube-wide quantities that the kernels may report on any band (the host takes the max over bands). */
enum {
  DM_M_MAXERR = 0,  /* max |x-y| over selected pixels of this band          (run_codec.py:275-276) */
  DM_M_ABSXY = 1,   /* max(np.abs(x), np.abs(y)) over selected pixels; int16 -32768 wraps (:285)   */
  DM_M_UMAX = 2,    /* max(0, max x) over ALL pixels of the reference        (run_codec.py:97,108)  */
  DM_M_UNEGMIN = 3, /* max(0, -min x) over ALL pixels of the reference       (run_codec.py:107)     */
  DM_M_LOW4 = 4,    /* 1 if any reference sample has (x & 0xF) != 0          (run_codec.py:98)      */
  DM_M_LOW2 = 5,    /* 1 if any reference sample has (x & 0x3) != 0          (run_codec.py:109)     */
  DM_M_SPARE0 = 6,
  DM_M_SPARE1 = 7
};

/* dm_fused_stats flags */
enum {
  DM_STATS_NO_MOMENTS = 1, /* PSNR-only variant: S_X,S_Y,S_XX,S_YY,S_XY untouched, S_SSE direct  */
  DM_STATS_GENERIC = 2     /* force the scalar any-layout kernel (cross-check / A-B timing)     */
};

typedef struct dm_pair {
  const void* ref;      /* original cube (device)                    */
  const void* tst;      /* decoded cube (device), same geometry      */
  int32_t dtype;        /* DM_U8 / DM_U16 / DM_I16                   */
  int32_t layout;       /* DM_BSQ / DM_BIP                           */
  int64_t bands;
  int64_t rows;         /* rows present in the buffer                */
  int64_t width;
  int64_t band_stride;  /* DM_BSQ: elements between bands (>= rows*width); ignored for DM_BIP */
  int32_t ref_has_nodata, ref_nodata;  /* integer nodata of the original (rasterio ds.nodata) */
  int32_t tst_has_nodata, tst_nodata;
} dm_pair_t;

/* library / device ------------------------------------------------------------------------ */
int dm_abi_version(void);
const char* dm_last_error(void);
/* number of SMs, or a negative DM_E* code when no CUDA device is usable */
int dm_device_sm_count(void);
/* kernels this library has launched in this process so far (bench.py reports the difference over
 * its timed region as "gpu_launches") */
int64_t dm_launch_count(void);
/* size in bytes of the device scratch ("workspace") that dm_spectral / dm_fused_bip take: per-block
 * float64 partials plus the arrival counter of their in-kernel ordered final reduction.  The caller
 * allocates it, zeroes it ONCE, and may reuse it for any number of launches on one stream (the last
 * block of every launch resets the counter); launches on different streams need different workspaces. */
int64_t dm_workspace_bytes(void);

/* masks -------------------------------------------------------------------------------------
 * Per-pixel validity plane (uint8, rows*width), bit set = pixel selected:
 *   DM_VALID_METRICS   = valid_in AND every band != nodata in BOTH cubes   (run_codec.py:249-263)
 *   DM_VALID_QUICKLOOK = dataset masks AND band 1 != nodata in both cubes  (quicklooks.py:35-45)
 *   DM_VALID_SPECTRAL  = valid_in if given, else the two dataset masks     (run_codec.py:314-319)
 * where a dataset mask is rasterio's dataset_mask(): any band != nodata (all valid without nodata).
 * valid_in (uint8, nonzero = valid) may be NULL.  counts_out (3 x int64, accumulated) receives
 * the number of pixels with each bit set. */
int dm_validity(const dm_pair_t* p, const uint8_t* valid_in, uint8_t* plane_out,
                int64_t* counts_out, void* stream);

/* fused single-pass integer reduction ---------------------------------------------------------
 * Replaces the arithmetic of compute_metrics' band loop (run_codec.py:268-285) and of
 * effective_data_range (run_codec.py:86-117): reads both cubes once and ACCUMULATES, per band,
 * sums[DM_NSTAT] (add) and maxs[DM_NSTAT] (max) as laid out above, plus, when hist_bins > 0,
 * hist[band*hist_bins + min(|x-y|, hist_bins-1)] (add).  `plane` (NULL = all pixels) selects
 * pixels whose byte has the single bit `plane_bit` set; DM_M_UMAX.. are always over all pixels.
 * hist_bins must be 0 or a power of two <= 1024. */
int dm_fused_stats(const dm_pair_t* p, const uint8_t* plane, int32_t plane_bit, int32_t hist_bins,
                   uint32_t flags, int64_t* sums, int64_t* maxs, int64_t* hist, void* stream);

/* the same reduction for a BATCH of pairs of one geometry in ONE launch (the tiles of a manifest, the decoded
 * tiles of a rate sweep: run_codec.py:448 -> 472 -> 475 calls compute_metrics once per tile, rate and rep).  A
 * Case-A tile pair is 16.8 MB -- 2.6 us at HBM speed, less than a kernel launch -- so a sweep over tiles is launch
 * bound pair by pair; here grid.y walks the pairs.  geom gives dtype / DM_BSQ / bands / rows / width / band_stride
 * (its ref / tst are ignored); items_dev is a DEVICE array of n_items {ref, tst, sums, maxs}; every item's sums /
 * maxs are accumulated exactly as by dm_fused_stats(plane = NULL, hist_bins = 0).  All cubes must be 16-byte
 * aligned (the caller's contract: the array lives on the device); DM_BSQ only. */
typedef struct dm_batch_item {
  const void* ref;
  const void* tst;
  int64_t* sums;   /* DM_NSTAT x bands, accumulated */
  int64_t* maxs;
} dm_batch_item_t;
int dm_fused_stats_batch(const dm_pair_t* geom, const dm_batch_item_t* items_dev, int32_t n_items, uint32_t flags,
                         void* stream);

/* per-pixel spectral pass --------------------------------------------------------------------
 * One pass over the spectral axis of every pixel.  Replaces the arithmetic of
 *   write_error_max8   quicklooks.py:123-150  (max_b |A-B|, invalid -> 0, float32 scaling -> uint8)
 *   SAM / SID          run_codec.py:328-339
 * Outputs (each may be NULL to skip):
 *   errmax_out  uint16 plane of max_b|x-y| (0 where the QUICKLOOK bit is clear)
 *   err8_g/err8_z  uint8 planes lut_g[min(e,cap_g)] / lut_z[min(e,cap_z)]; the LUTs (cap+1 bytes,
 *               device) are built by the host with the reference's own float32 expression
 *   hist8_g/hist8_z  256-bin int64 histograms of the uint8 planes (accumulated; give the
 *               STATISTICS_MEAN / STATISTICS_STDDEV tags of quicklooks.py:175-184 exactly)
 *   spectral_acc  double[3] {sum arccos, sum sid, n} over pixels with the SPECTRAL bit (all pixels
 *               when plane == NULL), ACCUMULATED: per-block partials are added up in a fixed order
 *               by the last block of the launch (deterministic), through `workspace`
 *               (dm_workspace_bytes(), see above)
 * want_sid = 0 skips SID (its partial is 0). */
int dm_spectral(const dm_pair_t* p, const uint8_t* plane,
                uint16_t* errmax_out,
                const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g,
                const uint8_t* lut_z, int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z,
                int32_t want_sam, int32_t want_sid, double* spectral_acc, void* workspace, void* stream);

/* lanes that share one pixel in the register-resident SAM / SID kernel of dm_spectral (16-bit BIP cubes, no planes):
 * 0 = by band count (8 lanes up to 192 bands -- four pixels per warp for EnMAP's 180 --, 16 up to 256, else 32),
 * 8 / 16 / 32 = pin (a choice the band count does not fit falls back to 32).  Thread-local; for A/B measurements
 * and the parity tests, which cover every grouping. */
int dm_spectral_lanes_per_pixel(int32_t lanes);

/* one-pass BIP kernel: dm_fused_stats (moments, no histogram) + dm_spectral (error planes, SAM) from a
 * SINGLE read of both cubes -- the tile is staged once in shared memory by TMA bulk copies and
 * consumed by a per-band and a per-pixel warp group.  Same outputs and conventions as the two
 * entry points above (spectral_acc: double[3] accumulated, SID slot untouched).  Supports DM_BIP,
 * 16-bit samples, bands a multiple of 4 in 4..256, 16-byte aligned cubes; anything else returns
 * DM_EUNSUPPORTED and the caller uses the two separate passes.  180-band cubes (EnMAP) take a kernel
 * specialised at compile time for that pixel pitch; a partial last tile goes through the generic one. */
/* Launch chaining for sweeps (thread-local switch, default OFF).  While it is on, launches of the 180-band
 * one-pass kernel carry the programmatic-stream-serialization attribute: the next launch's CTAs start as the
 * previous launch's CTAs exit, so that tail and ramp-up overlap (0.79 -> 0.85 of the HBM peak in a sweep).  A
 * chained launch READS its cubes (and validity plane) before the preceding kernel in the stream is guaranteed to
 * have flushed its writes -- every global WRITE of the launch waits (griddepcontrol.wait) -- so switch it on only
 * when the inputs of a launch are not produced by the kernel issued immediately before it on the same stream
 * (a sweep over cubes that are already resident: engine.PreparedFused). */
void dm_launch_chaining(int32_t on);
/* Which build of the one-pass BIP kernel dm_fused_bip launches for 180-band cubes (thread-local, default 0):
 * 0 = the measured choice (23 band warps for plain statistics + SAM, 12 with planes or a validity plane),
 * 12 / 23 = pin that build, 1 = the run-time-geometry kernel that serves every other band count.  For A/B
 * measurements and for the parity tests, which cover every build. */
int dm_fused_bip_variant(int32_t variant);
int dm_fused_bip(const dm_pair_t* p, const uint8_t* plane, int64_t* sums, int64_t* maxs,
                 uint16_t* errmax_out,
                 const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g,
                 const uint8_t* lut_z, int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z,
                 int32_t want_sam, double* spectral_acc, void* workspace, void* stream);

/* The same one-pass kernel with the VALIDITY RULE FOLDED IN (ABI 9): for pairs whose files carry a nodata value
 * (the real EnMAP products: int16, nodata -32768 in both files) and / or a caller mask, dm_validity + dm_fused_bip
 * read the pair twice; here the pixel warps that serve a tile first sweep it for the three rules of dm_validity
 * (run_codec.py:249-263, quicklooks.py:35-45, run_codec.py:314-319), hand the tile's validity bytes to the band
 * warps through shared memory and only then start their own spectral sweep -- one read of the pair.
 * valid_in (uint8, nonzero = valid, 16-byte aligned) may be NULL; plane_out (may be NULL when rows*width is a
 * multiple of 64) receives the plane exactly as dm_validity writes it, for the kernels that follow (dm_spectral's
 * SID, dm_sobel_lmse take it as `plane`); counts_out (3 x int64, accumulated, may be NULL) the pixels per bit.
 * The reference's all-False-mask rule (run_codec.py:264: no valid pixel -> every pixel counts) is the caller's:
 * counts_out[0] == 0 after the launch means "zero the statistics and rerun dm_fused_bip with plane = NULL".
 * 180-band DM_BIP cubes of at least 64 pixels only; anything else returns DM_EUNSUPPORTED and the caller runs
 * dm_validity + dm_fused_bip.  Everything else as dm_fused_bip. */
int dm_fused_bip_scan(const dm_pair_t* p, const uint8_t* valid_in, uint8_t* plane_out, int64_t* counts_out,
                      int64_t* sums, int64_t* maxs, uint16_t* errmax_out,
                      const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g,
                      const uint8_t* lut_z, int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z,
                      int32_t want_sam, double* spectral_acc, void* workspace, void* stream);

/* one-pass BSQ kernel for few-band cubes (Sentinel-2 Case A): dm_fused_stats (moments, no histogram)
 * + the error planes of dm_spectral from a SINGLE read -- a thread loads the same 8-pixel vector of
 * every band, so it holds whole spectra.  Same outputs and conventions as those two entry points.
 * Supports DM_BSQ, 16-bit samples, 1..4 bands, bands starting on 16-byte boundaries, caps <= 255;
 * anything else returns DM_EUNSUPPORTED and the caller uses the two separate passes. */
int dm_fused_bsq(const dm_pair_t* p, const uint8_t* plane, int64_t* sums, int64_t* maxs,
                 uint16_t* errmax_out,
                 const uint8_t* lut_g, int32_t cap_g, uint8_t* err8_g, int64_t* hist8_g,
                 const uint8_t* lut_z, int32_t cap_z, uint8_t* err8_z, int64_t* hist8_z, void* stream);

/* Sobel LMSE -----------------------------------------------------------------------------------
 * Replaces sobel_mag + mse in the LMSE loop (run_codec.py:123-137, 341-346): for every band,
 * sum over pixels of (|grad ref| - |grad tst|)^2 with the 3x3 Sobel pair and edge replication.
 * Counts buffer rows [row_begin,row_end); buffer row 0 is image row img_row0 of img_rows, and the
 * buffer must hold one halo row on each side that is not an image border.
 * lmse_acc: double[bands], ACCUMULATED (caller-zeroed): the per-band sums.  The blocks' float64 partials are
 * added in a fixed order by the blocks that finish last (inside the kernel), so the result is reproducible
 * from run to run and no follow-up reduction is needed.
 * scratch: double[bands * dm_sobel_nblocks()], 16-byte aligned, contents irrelevant; workspace: dm_workspace_bytes()
 * bytes, zeroed once by the caller (the kernels leave it zeroed), one per stream.  1..2048 bands. */
/* sobel_mag as a function of its own (run_codec.py:123-137): float64 magnitude map of one (rows, width) plane,
 * 3x3 Sobel pair with edge replication; bit-identical to the reference for 8/16-bit integer samples. */
int dm_sobel_mag(const void* img, int32_t dtype, int64_t rows, int64_t width, double* out, void* stream);
int dm_sobel_nblocks(void);
int dm_sobel_lmse(const dm_pair_t* p, int64_t row_begin, int64_t row_end, int64_t img_row0,
                  int64_t img_rows, double* scratch, double* lmse_acc, void* workspace, void* stream);

/* Gaussian-window SSIM (addition) --------------------------------------------
 * 11-tap separable Gaussian (sigma 1.5, truncate 3.5), float64, skimage semantics
 * (use_sample_covariance=False), mean over the image cropped by 5 px.  DM_BSQ only.
 * sum_acc / cnt_acc: double[bands] each, ACCUMULATED (caller-zeroed): sum of S and number of pixels over the
 * pixels of buffer rows [row_begin,row_end) that lie inside the crop; the band's block partials are added in
 * a fixed order by the band's last block (inside the kernel).
 * scratch: double[bands * 2 * dm_ssim_nblocks()], contents irrelevant; workspace as for dm_sobel_lmse.
 * The buffer must hold 5 halo rows on each side that is not an image border.  1..2048 bands. */
int dm_ssim_nblocks(void);
/* which Gaussian-SSIM kernel dm_ssim_gauss launches (thread-local):
 *   0 / 2  the shared-memory tiled all-FP64 kernel (any geometry; the measured choice: 5.71 ms per 10980^2 x 4 scene)
 *   3  the row-streaming "ring" kernel where it applies (16-bit cubes of even width; all float64, 128-column strips
 *      streamed 11 rows at a time, cp.async staging, vertical pass as a register scatter: 5.86 ms), else tiled
 *   1  the warp-streaming kernel with an exact integer horizontal pass (8.98 ms: IMAD.WIDE runs at a third of the
 *      DFMA rate on B200)
 * Three independent implementations of one definition; the parity tests check each against both oracles. */
int dm_ssim_variant(int32_t variant);
int dm_ssim_gauss(const dm_pair_t* p, double data_range, int64_t row_begin, int64_t row_end,
                  int64_t img_row0, int64_t img_rows, double* scratch, double* sum_acc, double* cnt_acc,
                  void* workspace, void* stream);

/* multi-GPU combine ----------------------------------------------------------------------------
 * After ONE all-gather of every rank's run of `records` flat partial vectors, each
 * [n_sum int64 | n_max int64 | n_f64 double] (the layout of engine.Partials; a run is the partials of
 * `records` consecutive pairs of a sweep), this reduces the `world` gathered runs into `out` (one run,
 * same layout): int64 sums added, int64 maxima maxed, float64 sums added in RANK ORDER, so the result
 * is bit-identical on every rank and from run to run.
 * gathered: world x records x (n_sum+n_max+n_f64) 8-byte words. */
int dm_combine_partials(const void* gathered, int32_t world, int64_t records, int64_t n_sum, int64_t n_max,
                        int64_t n_f64, void* out, void* stream);

/* layout helper: (rows,width,bands) -> (bands,rows,width), same dtype (1 or 2 bytes/sample) */
int dm_bip_to_bsq(const void* src, void* dst, int32_t elem_bytes, int64_t bands, int64_t rows,
                  int64_t width, void* stream);


/* ===== rows either side of the path ========================================= */

/* one cube (same geometry fields as dm_pair_t) */
typedef struct dm_cube {
  const void* data;     /* device */
  int32_t dtype;        /* DM_U8 / DM_U16 / DM_I16 */
  int32_t layout;       /* DM_BSQ / DM_BIP */
  int64_t bands, rows, width;
  int64_t band_stride;  /* DM_BSQ: elements between bands; ignored for DM_BIP */
} dm_cube_t;

/* RGB quicklook, part 1 -- replaces the sort behind np.percentile in stretch_params_from_baseline
 * (quicklooks.py:51-72): exact value histograms of up to 4 selected bands (0-based indices in the HOST
 * array sel_bands), hist[i*65536 + bin] ACCUMULATED, bin = the sample (int16: sample + 32768, so bins
 * are in value order).  `plane`/`plane_bit` as in dm_fused_stats (NULL = all pixels).  The host turns the
 * histogram into numpy's linearly interpolated percentiles (finish.percentiles_from_hist). */
int dm_band_hist(const dm_cube_t* c, const int32_t* sel_bands, int32_t nsel, const uint8_t* plane,
                 int32_t plane_bit, int64_t* hist, void* stream);

/* RGB quicklook, part 2 -- replaces stretch8 in write_rgb_8bit (quicklooks.py:81-89):
 * out[i*rows*width + p] = luts[i*65536 + bin(sample(sel_bands[i], p))].  The float32 stretch is an
 * elementwise function of an integer sample, so the host tabulates it with the reference's own
 * expression (finish.stretch8_lut) and the planes are bit-exact by construction. */
int dm_lut_bands_u8(const dm_cube_t* c, const int32_t* sel_bands, int32_t nsel, const uint8_t* luts,
                    uint8_t* out, void* stream);

/* Baseline builders' requantisation of n 16-bit samples (src may equal dst):
 *   mode DM_REQ_TRUNC  ((u >> k) << k) on the uint16 view, samples equal to `nodata` untouched
 *                      (trunc_uint16 / write_truncated_copy, make_baseline_B.py:281-316; "14-in-16" is k = 2)
 *   mode DM_REQ_ROUND  ((u + 2^(k-1)) >> k) << k in uint16 arithmetic (to_12in16, make_baseline_A.py:166-167, k = 4) */
enum { DM_REQ_TRUNC = 0, DM_REQ_ROUND = 1 };
int dm_requantize(const void* src, void* dst, int32_t dtype, int64_t n, int32_t mode, int32_t k,
                  int32_t has_nodata, int32_t nodata, void* stream);

/* Scene error maps of make_scene_error_map (make_baseline_B.py:324-419): per pixel over the bands IN ORDER,
 * d = |ref - cmp| (0 where `valid`, uint8 rows*width nonzero = valid, is clear):
 *   DM_EM_MEAN   float32 sum of d / bands            DM_EM_RMS   sqrt(float32 sum of int32(d*d) / bands)
 *   DM_EM_COUNT3 #{d == 2^k_bits - 1}                DM_EM_MAX   max d
 *   DM_EM_P95    the reference's per-pixel histogram percentile (bins 0..2^k_bits-1, k_bits <= 4,
 *                p95_thr = uint32(bands * 0.95) computed by the host)
 * out_plane: float32 rows*width (written); out_max_bits: bit pattern of the largest output value
 * (uint32, atomicMax into a caller-zeroed word; outputs are non-negative).
 * dm_scale_plane_u8 is the final (clip(v, 0, emax) * scale + 0.5) -> uint8 of make_baseline_B.py:417 with
 * scale = float32(255.0 / emax) from the host. */
enum { DM_EM_MEAN = 0, DM_EM_RMS = 1, DM_EM_COUNT3 = 2, DM_EM_MAX = 3, DM_EM_P95 = 4 };
int dm_scene_error(const dm_pair_t* p, const uint8_t* valid, int32_t mode, int32_t k_bits, uint32_t p95_thr,
                   float* out_plane, uint32_t* out_max_bits, void* stream);
int dm_scale_plane_u8(const float* plane, int64_t n, float emax, float scale, uint8_t* out, void* stream);

/* Reversible spectral differencing of the CCSDS-121 / JPEG-LS wrappers along the band axis of a BSQ cube
 * (bands x npix, `band_stride` elements between bands; src may equal dst):
 *   arith DM_DIFF_MODULO    R[b] = X[b] - X[b-1] mod 2^N, inverse = running sum mod 2^N
 *                           (_diff1_bsq_signed/_unsigned, _int1_bsq_*: ccsds121_wrap.py:66-85;
 *                            _diff1_forward/_inverse uint16, uint8: jpegls_wrap.py:96-99, 110-113, 103-105, 117-119)
 *   arith DM_DIFF_SATURATE  int16 only: R = clip(X[b] - X[b-1]), X[b] = clip(R[b] + X[b-1]) (jpegls_wrap.py:100-102, 114-116)
 * inverse = 0 forward, 1 inverse. */
enum { DM_DIFF_MODULO = 0, DM_DIFF_SATURATE = 1 };
int dm_diff1(const void* src, void* dst, int32_t dtype, int32_t arith, int32_t inverse, int64_t bands,
             int64_t npix, int64_t band_stride, void* stream);

/* Raw interleave conversion of the wrappers (_write_raw_interleaved / _read_raw_interleaved,
 * ccsds121_wrap.py:44-64, ccsds123_wrap.py:43-63): any of DM_BSQ / DM_BIL / DM_BIP to any other,
 * contiguous cubes of 1- or 2-byte samples. */
int dm_interleave(const void* src, void* dst, int32_t elem_bytes, int32_t from_layout, int32_t to_layout,
                  int64_t bands, int64_t rows, int64_t width, void* stream);

/* ===== NVLink peer-memory exchange of partial vectors (alternative to the NCCL all-gather) =====
 * Every rank owns one exchange buffer  [world][capacity] partial vectors | world uint64 arrival flags,
 * allocated by dm_p2p_alloc (cudaMalloc, zeroed) and exported as a 64-byte CUDA IPC handle that the host
 * layer hands to the other ranks (any transport; torch.distributed.all_gather_object here); dm_p2p_open maps a
 * peer's buffer.  dm_p2p_push copies `total_words` 8-byte words from src into peer_dst[r] for every r (one
 * CTA per destination, plain stores over NVLink; r = own rank is a local copy) and then stores flag_value to
 * peer_flag[r] with system-scope release.  dm_p2p_combine waits (system-scope acquire, at most timeout_s,
 * then *status = 1 and nothing is written) until all `world` LOCAL flags are >= need and reduces records
 * [rec0, rec0+nrec) of the world local copies into out exactly as dm_combine_partials does.
 * peer_dst / peer_flag are HOST arrays of `world` device pointers. */
int dm_p2p_alloc(int64_t bytes, void** ptr, void* handle64);
int dm_p2p_open(const void* handle64, void** ptr);
int dm_p2p_close(void* ptr);
int dm_p2p_free(void* ptr);
/* zero `bytes` of an exchange buffer on `stream` (the arrival flags, between two sweeps; the caller brackets it
 * with barriers so that no peer is pushing) */
int dm_p2p_zero(void* ptr, int64_t bytes, void* stream);
int dm_p2p_push(const void* src, int64_t total_words, void* const* peer_dst, void* const* peer_flag, int32_t world,
                uint64_t flag_value, void* stream);
int dm_p2p_combine(const void* gathered, const void* flags, int32_t world, uint64_t need, int64_t capacity,
                   int64_t rec0, int64_t nrec, int64_t n_sum, int64_t n_max, int64_t n_f64, void* out,
                   uint32_t* status, double timeout_s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DM_B200_H */
