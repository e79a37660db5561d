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
#include <type_traits>

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
constexpr int XP = IW + 2;                              // row pitch of the staged samples: 44 ints = 176 bytes, so that pass H reads its
                                                        // 14 inputs as three LDS.128 + one LDS.64 (rows 176 B apart: conflict free)
constexpr int kSsimSmem = 2 * IH * XP * 4 + 4 * PH * HP * 8;

struct Taps { double w[2 * RAD + 1]; };

template <typename T>
__global__ void __launch_bounds__(256, 2)
ssim_gauss_kernel(const T* __restrict__ ref, const T* __restrict__ tst, int64_t band_stride, int64_t width,
                  int64_t r_lo, int64_t r_hi, int64_t buf_rows, Taps taps, double c1, double c2, double* scratch,
                  double* sum_acc, double* cnt_acc, void* workspace) {
  extern __shared__ __align__(16) unsigned char ssim_smem[];
  int (*xs)[XP] = reinterpret_cast<int (*)[XP]>(ssim_smem);
  int (*ys)[XP] = reinterpret_cast<int (*)[XP]>(ssim_smem + IH * XP * 4);
  double (*hp)[PH][HP] = reinterpret_cast<double (*)[PH][HP]>(ssim_smem + 2 * IH * XP * 4);   // [4][PH][HP]
  __shared__ double red[2][32];
  const int band = blockIdx.y;
  const T* A = ref + (int64_t)band * band_stride;
  const T* R = tst + (int64_t)band * band_stride;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  constexpr uint32_t OFS = sizeof(T) == 2 && T(-1) < T(0) ? 0x8000u : 0u;      // int16 -> offset binary
  constexpr double kSplice = 4503599627370496.0 + (OFS ? 32768.0 : 0.0);       // 2^52 (+ the int16 offset)
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
    // the two samples of an element stay in SEPARATE registers until they are staged: packing them right behind the
    // loads (one register per element) made the pack instruction wait for the loads -- the "prefetch" was a
    // synchronous load in front of pass H (ncu r02l: 20 % of the stall samples on those four IMADs)
    uint32_t pre_a[NPRE], pre_b[NPRE];
    // element k of this thread sits at (lr, lc) of every tile: its offset from the tile's first staged sample is fixed.
    // Tiles whose staged window lies inside the buffer (all but the last tile row / column) need no clamping, and
    // the clamps with their 64-bit compares were ~35 instructions per output pixel of a kernel that waits for
    // its non-FP64 phases (r0 >= 5 and c0 >= 5 always: only the upper edges can stick out).
    uint32_t off[NPRE];
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int e = threadIdx.x + 256 * k;
      const int lr = e / IW, lc = e - lr * IW;
      off[k] = (uint32_t)lr * (uint32_t)width + (uint32_t)lc;
    }
    const bool fits32 = width < (int64_t)(1 << 24);
    auto fetch_tile = [&](int64_t t) {
      const int64_t r0 = r_lo + (t / tiles_x) * SH, c0 = c_lo + (t % tiles_x) * SW;
      if (fits32 && r0 - RAD + IH <= buf_rows && c0 - RAD + IW <= width) {
        const T* a0 = A + (r0 - RAD) * width + (c0 - RAD);
        const T* b0 = R + (r0 - RAD) * width + (c0 - RAD);
#pragma unroll
        for (int k = 0; k < NPRE; ++k) {
          const bool have = k < NPRE - 1 || threadIdx.x + 256 * k < IH * IW;
          pre_a[k] = have ? (uint32_t)(uint16_t)a0[off[k]] : 0u;
          pre_b[k] = have ? (uint32_t)(uint16_t)b0[off[k]] : 0u;
        }
        return;
      }
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        const int e = threadIdx.x + 256 * k;
        const int lr = e / IW, lc = e - lr * IW;
        int64_t r = r0 + lr - RAD, c = c0 + lc - RAD;
        r = r < 0 ? 0 : (r >= buf_rows ? buf_rows - 1 : r);
        c = c < 0 ? 0 : (c >= width ? width - 1 : c);
        pre_a[k] = e < IH * IW ? (uint32_t)(uint16_t)A[r * width + c] : 0u;
        pre_b[k] = e < IH * IW ? (uint32_t)(uint16_t)R[r * width + c] : 0u;
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
          // samples go to shared memory in offset binary (int16: x ^ 0x8000), i.e. as non-negative integers
          // below 2^16, which pass H turns into float64 with the 2^52 splice (see there)
          xs[lr][lc] = (int)(pre_a[k] ^ OFS);
          ys[lr][lc] = (int)(pre_b[k] ^ OFS);
        }
      }
      __syncthreads();
      if (t + gridDim.x < ntiles) fetch_tile(t + gridDim.x);
      // ---- pass H: item = (group of 4 output columns, row lr); 512 items, two per thread.  The ROW is
      // the fastest thread index: the lanes of a warp read the same columns of 32 rows (44 words apart, 16-byte vectors)
      // and write the same columns of 32 plane rows (33 doubles apart) -- both strides odd, conflict free.
#pragma unroll 1
      for (int item = threadIdx.x; item < IH * (SW / 4); item += 256) {
        const int lr = item & (IH - 1), g4 = (item >> 6) * 4;
        double vx[14], vy[14], vq[14], vp[14];
        int xi[16], yi[16];              // the item's 14 inputs (+2) of both cubes: three LDS.128 + one LDS.64 each
        {
          const int4* qx = reinterpret_cast<const int4*>(&xs[lr][g4]);
          const int4* qy = reinterpret_cast<const int4*>(&ys[lr][g4]);
          const int4 a = qx[0], b = qx[1], c = qx[2], d = qy[0], e = qy[1], f = qy[2];
          const int2 gx = *reinterpret_cast<const int2*>(&xs[lr][g4 + 12]), gy = *reinterpret_cast<const int2*>(&ys[lr][g4 + 12]);
          xi[0] = a.x; xi[1] = a.y; xi[2] = a.z; xi[3] = a.w; xi[4] = b.x; xi[5] = b.y; xi[6] = b.z; xi[7] = b.w;
          xi[8] = c.x; xi[9] = c.y; xi[10] = c.z; xi[11] = c.w; xi[12] = gx.x; xi[13] = gx.y;
          yi[0] = d.x; yi[1] = d.y; yi[2] = d.z; yi[3] = d.w; yi[4] = e.x; yi[5] = e.y; yi[6] = e.z; yi[7] = e.w;
          yi[8] = f.x; yi[9] = f.y; yi[10] = f.z; yi[11] = f.w; yi[12] = gy.x; yi[13] = gy.y;
        }
#pragma unroll
        for (int k = 0; k < 14; ++k) {
          // int -> float64 without the conversion unit (I2F.F64 costs two FP64-pipe slots here, measured): the
          // sample spliced into the mantissa of 2^52 IS 2^52 + u; one exact subtraction gives u (or u - 32768)
          const double x = __hiloint2double(0x43300000, xi[k]) - kSplice;
          const double y = __hiloint2double(0x43300000, yi[k]) - kSplice;
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
            // quotient: MUFU.RCP64H seed (~2^-20) + two Newton steps (full precision), no slow-path branch;
            // den >= c1 * c2 > 0 and far from the exponent limits
            double q;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(den));
            q = fma(fma(-den, q, 1.0), q, q);
            q = fma(fma(-den, q, 1.0), q, q);
            acc = fma(num, q, acc);
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

// ---------------------------------------------------------------------------------------------------------------
// Streaming kernel (dm_ssim_variant(1); NOT the default -- measured slower, kept as an independent second
// implementation that the parity tests check against the oracle, and as the record of the experiment).
// Idea: no block barrier, no halo recomputation, 64 FP64 operations per pixel instead of ~130:
//
//   * a WARP owns a strip of 32 output columns and streams down the rows of a segment; lane = column
//   * horizontal pass in EXACT INTEGER arithmetic on the other pipe: the window is the Gaussian quantised to
//     31-bit fixed point (W_k, sum exactly 2^31 -- a normalised window that differs from the ideal taps by < 2^-32,
//     far inside the 1e-6 gate; what matters for the variance terms is that all planes see the SAME window and
//     that nothing is rounded before the products are summed).  Per input sample the lanes store three 32-bit
//     words in a warp-private row buffer -- p = x | y << 16, x*y, (x-y)^2 -- and every output is 11 taps x
//     4 IMAD.WIDE.U32 (sums of W x, W p, W xy, W d^2 < 2^63; W y = (W p - W x) >> 16).  int16 samples are taken
//     in offset binary (x ^ 0x8000): variances and covariances do not see the shift, the means get it back.
//   * one I2F.F64.U64 per plane, then the vertical pass as a SCATTER into 11 pending outputs per lane held in
//     registers (44 accumulators; the row loop is unrolled 11 x so that every index is static): each horizontal
//     result is used 11 times, nothing is computed twice, no shared-memory traffic in FP64
//   * SSIM from {E x, E y, E xy, E d^2}: sigma_x^2 + sigma_y^2 = var(d) + 2 cov, so neither x^2 nor y^2 is
//     filtered; the quotient is a MUFU.RCP64H seed + two Newton steps (no library slow-path branch)
//   * the next rows' samples are fetched two steps ahead into registers
// Measured (r02, tools/probe_ssim.py, tools/ubench_fp64.cu): 10980^2 x 4 scene 8.98 ms against 6.19 ms for the tiled
// kernel; results agree to 2e-12.  Why: IMAD.WIDE issues at ~23 lane-ops/clk/SM here (ptxas also splits the 64-bit
// accumulate into IMAD.WIDE + IADD3 + IADD3.X: 252 warp instructions per 32 pixels) while DFMA runs at 61.5 -- on
// this chip the FP64 pipe IS the fast wide multiplier, and an exact-integer horizontal pass costs three times what
// the rounded one does.  What carries over to the tiled kernel: the var(d) + 2 cov form and the branch-free quotient.
struct StreamArgs {
  const void* ref;
  const void* tst;
  int64_t band_stride, width, buf_rows;
  int64_t r_lo, r_hi;                 // counted buffer rows
  int seg_rows, strips_x;
  int64_t nseg;
  uint32_t w[2 * RAD + 1];            // fixed-point taps, sum 2^31
  double wv[2 * RAD + 1];             // w[k] * 2^-62: vertical weight including the horizontal scale
  double c1, c2;
  double* scratch; double* sum_acc; double* cnt_acc; void* workspace;
};

constexpr int kStreamWarps = 4;
constexpr int kStreamBuf = 48;        // words per array of a row buffer (42 used)

__device__ __forceinline__ unsigned long long madw(uint32_t a, uint32_t b, unsigned long long c) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kStreamWarps * 32, 3)
ssim_stream_kernel(const StreamArgs g) {
  constexpr int NT = 2 * RAD + 1;
  constexpr uint32_t OFS = sizeof(T) == 2 && T(-1) < T(0) ? 0x8000u : 0u;      // int16 -> offset binary
  __shared__ uint32_t rowbuf[kStreamWarps][2][3][kStreamBuf];
  __shared__ double red[2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int band = blockIdx.y;
  const T* A = static_cast<const T*>(g.ref) + (int64_t)band * g.band_stride;
  const T* R = static_cast<const T*>(g.tst) + (int64_t)band * g.band_stride;
  const int64_t c_lo = RAD, c_hi = g.width - RAD;
  double acc_s = 0.0;
  unsigned acc_n = 0;
  double va[NT][4];                    // pending vertical sums: slot a started at the step with phase a
  const int64_t ntasks = g.nseg * g.strips_x;
  for (int64_t task = (int64_t)blockIdx.x * kStreamWarps + warp; task < ntasks; task += (int64_t)gridDim.x * kStreamWarps) {
    const int64_t seg = task / g.strips_x;
    const int strip = (int)(task - seg * g.strips_x);
    const int64_t r0 = g.r_lo + seg * g.seg_rows;
    const int64_t r1 = r0 + g.seg_rows < g.r_hi ? r0 + g.seg_rows : g.r_hi;
    const int64_t c0 = c_lo + (int64_t)strip * 32;
    const bool col_ok = c0 + lane < c_hi;
    const int n_in = (int)(r1 - r0) + 2 * RAD;           // input rows r0-5 .. r1+4
    // this lane's input columns: c0-5+lane, and (lanes 0..9) c0+27+lane; clamped columns only feed dropped outputs
    int64_t ca = c0 - RAD + lane, cb = c0 + 32 - RAD + lane;
    ca = ca >= g.width ? g.width - 1 : ca;
    cb = cb >= g.width ? g.width - 1 : cb;
    const bool second = lane < 2 * RAD;
    const T* pa = A + (r0 - RAD) * g.width;
    const T* pr = R + (r0 - RAD) * g.width;
    uint32_t f0[4], f1[4];               // fetched samples of the next two rows {x(ca), y(ca), x(cb), y(cb)}
    auto fetch = [&](int s, uint32_t (&f)[4]) {
      if (s < n_in) {
        const T* qa = pa + (int64_t)s * g.width;
        const T* qr = pr + (int64_t)s * g.width;
        f[0] = (uint32_t)(uint16_t)__ldg(qa + ca); f[1] = (uint32_t)(uint16_t)__ldg(qr + ca);
        if (second) { f[2] = (uint32_t)(uint16_t)__ldg(qa + cb); f[3] = (uint32_t)(uint16_t)__ldg(qr + cb); }
      }
    };
    fetch(0, f0);
    fetch(1, f1);
    // one row: stage the fetched samples, horizontal pass (integer), vertical scatter with compile-time phase PH
    auto step = [&](auto ph_tag, int s, uint32_t (&fc)[4]) {
      constexpr int PH = decltype(ph_tag)::value;
      uint32_t (*buf)[kStreamBuf] = rowbuf[warp][s & 1];
      {
        const uint32_t x = fc[0] ^ OFS, y = fc[1] ^ OFS;
        const uint32_t d = x > y ? x - y : y - x;
        buf[0][lane] = x | (y << 16); buf[1][lane] = x * y; buf[2][lane] = d * d;
        if (second) {
          const uint32_t x2 = fc[2] ^ OFS, y2 = fc[3] ^ OFS;
          const uint32_t d2 = x2 > y2 ? x2 - y2 : y2 - x2;
          buf[0][lane + 32] = x2 | (y2 << 16); buf[1][lane + 32] = x2 * y2; buf[2][lane + 32] = d2 * d2;
        }
      }
      fetch(s + 2, fc);                  // this buffer is free again: the row after next goes into it
      __syncwarp();
      unsigned long long hx = 0, hp = 0, hxy = 0, hd = 0;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const uint32_t p = buf[0][lane + k], q = buf[1][lane + k], e = buf[2][lane + k];
        const uint32_t wk = g.w[k];
        hx = madw(p & 0xffffu, wk, hx); hp = madw(p, wk, hp); hxy = madw(q, wk, hxy); hd = madw(e, wk, hd);
      }
      const unsigned long long hy = (hp - hx) >> 16;
      const double h[4] = {__ull2double_rn(hx), __ull2double_rn(hy), __ull2double_rn(hxy), __ull2double_rn(hd)};
#pragma unroll
      for (int a = 0; a < NT; ++a) {
        constexpr int dummy = 0; (void)dummy;
        const int t = (PH - a + NT) % NT;                 // tap this slot receives now (static after unrolling)
        const double wt = g.wv[t];
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) va[a][pl] = t == 0 ? wt * h[pl] : fma(wt, h[pl], va[a][pl]);
      }
      if (s >= 2 * RAD && col_ok) {                       // the slot started 10 steps ago is complete
        constexpr int a = (PH + 1) % NT;
        const double ux1 = va[a][0], uy1 = va[a][1], uxy = va[a][2], ud2 = va[a][3];
        const double ux = OFS ? ux1 - 32768.0 : ux1, uy = OFS ? uy1 - 32768.0 : uy1;
        const double vxy = fma(-ux1, uy1, uxy);           // covariance (shift invariant)
        const double dm = ux1 - uy1;
        const double vsum = fma(2.0, vxy, fma(-dm, dm, ud2));     // var x + var y = var(x - y) + 2 cov
        const double num = fma(2.0, ux * uy, g.c1) * fma(2.0, vxy, g.c2);
        const double den = fma(ux, ux, fma(uy, uy, g.c1)) * (vsum + g.c2);
        double q;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(den));
        q = fma(fma(-den, q, 1.0), q, q);
        q = fma(fma(-den, q, 1.0), q, q);
        acc_s = fma(num, q, acc_s);
        acc_n += 1u;
      }
    };
#define DM_STEP(PH_, F_)                                                                  \
    if (s0 + PH_ < n_in) step(std::integral_constant<int, PH_>(), s0 + PH_, F_)
    // 11 steps per trip (static phases); two fetch buffers alternate, 11 is odd, so a second trip swaps them
    for (int s0 = 0; s0 < n_in; s0 += 2 * NT) {
      DM_STEP(0, f0); DM_STEP(1, f1); DM_STEP(2, f0); DM_STEP(3, f1); DM_STEP(4, f0); DM_STEP(5, f1);
      DM_STEP(6, f0); DM_STEP(7, f1); DM_STEP(8, f0); DM_STEP(9, f1); DM_STEP(10, f0);
      s0 += NT;
      DM_STEP(0, f1); DM_STEP(1, f0); DM_STEP(2, f1); DM_STEP(3, f0); DM_STEP(4, f1); DM_STEP(5, f0);
      DM_STEP(6, f1); DM_STEP(7, f0); DM_STEP(8, f1); DM_STEP(9, f0); DM_STEP(10, f1);
      s0 -= NT;
    }
#undef DM_STEP
    __syncwarp();                        // the next task's first row reuses rowbuf[warp][0]
  }
  acc_s = warp_sum_f64(acc_s);
  double cnt = warp_sum_f64((double)acc_n);
  if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = cnt; }
  __syncthreads();
  double t[2] = {0.0, 0.0};
  if (threadIdx.x == 0) {
    for (int w = 0; w < kStreamWarps; ++w) { t[0] += red[0][w]; t[1] += red[1][w]; }
  }
  __syncthreads();
  double* const accs[2] = {g.sum_acc, g.cnt_acc};
  ordered_band_sum<2>(t, g.scratch, static_cast<Workspace*>(g.workspace)->band_counter, accs, &red[0][0]);
}

// ---------------------------------------------------------------------------------------------------------------
// Ring kernel (dm_ssim_variant(3); 16-bit cubes with an even width).  What the captures of the tiled kernel say
// (profiles/r02l_ncu_ssim_tiled.txt): 282 instructions per band pixel of which 136 on the FP64 pipe, pipe 58 % busy,
// issue slots 60 % -- neither saturated; the time goes to three block barriers per tile at two CTAs per SM, to
// dependency waits, and to work done twice (18 % halo rows in the horizontal pass, conversions and products per
// 4-output group).  Same arithmetic (all float64, 11-tap separable), different schedule:
//
//   * a 128-thread block owns a strip of 128 output columns and STREAMS down a row segment in steps of 11 input
//     rows: no row is filtered twice (a segment pays 10 warm-up rows once), three independent blocks per SM
//   * stage: cp.async copies of 32-bit words (two pixels) one step ahead, straight into a double-buffered 12.7 KB
//     block of shared memory (no registers held across the passes)
//   * pass H: one item = (row, 4 neighbouring columns): seven words of each cube by three LDS.64 + one LDS.32, float64
//     through the 2^52 mantissa splice (int16 in offset binary), products once per input and item, 176 DFMA,
//     results to a 45 KB plane block
//     in a lane-major order (column 4 g + o lives at o * 32 + g) so that writes and reads are conflict free
//   * pass V: thread = column, the 11 rows of the step in order, each a SCATTER into 11 pending outputs held in
//     registers (44 accumulators; with 11 rows per step every phase is static): each horizontal result is read once
//     (4 LDS.64) and used 11 times -- against 9.7 shared loads per output and plane, and 4 % recomputation, before
//   * SSIM from {E x, E y, E(x^2+y^2), E xy}; the quotient as MUFU.RCP64H + two Newton steps
// 157 instructions per band pixel (122 FP64), two barriers per 11 rows.  Measured (r02q): 5.87 ms per 10980^2 x 4 scene
// against 6.16 ms -- 5 %, not the 35 % the instruction count promised: ncu (profiles/r02n_ncu_ssim_ring_first.txt) shows
// the FP64 pipe 49 % busy at three warps per scheduler against 58 % at four in the tiled kernel; a pure DFMA stream needs
// neither (tools/ubench_dfma_occ.cu: 12 warps per SM with 16 chains reach 95 %), so what is left are the fixed-latency
// waits between the shared-memory loads and the DFMA blocks that consume them, uneven rows per warp (11 rows on 4 warps)
// and the barriers.  With the tiled kernel's prefetch repaired (it packed the loaded samples right behind the loads,
// i.e. waited for them: 20 % of its stall samples) the tiled kernel is the faster one again: 5.71 ms.
constexpr int kRingThreads = 128;
constexpr int RS = 128;                 // output columns per strip
constexpr int RR = 2 * RAD + 1;         // input rows per step
constexpr int RWORDS = (RS + 2 * RAD) / 2;                 // 69 32-bit words (two pixels) per row and cube
constexpr int RPITCH = 72;              // words per staged row and cube (69 used; 288 bytes: rows stay 16-byte aligned)
constexpr int RPRE = (RR * RWORDS + kRingThreads - 1) / kRingThreads;   // 6 words per thread, cube and step
constexpr int kRingSmem = 2 * 2 * RR * RPITCH * 4 + RR * 4 * RS * 8;

struct RingArgs {
  const void* ref;
  const void* tst;
  int64_t band_stride, width, buf_rows;
  int64_t r_lo, r_hi;                 // counted buffer rows
  int seg_rows, strips_x;
  int64_t nseg;
  double w[2 * RAD + 1];
  double c1, c2;
  double* scratch; double* sum_acc; double* cnt_acc; void* workspace;
};

template <bool SIGNED>
__global__ void __launch_bounds__(kRingThreads, 3)
ssim_ring_kernel(const RingArgs g) {
  extern __shared__ __align__(16) unsigned char ring_smem[];
  uint32_t (*raw)[2][RR][RPITCH] = reinterpret_cast<uint32_t (*)[2][RR][RPITCH]>(ring_smem);              // [2 buffers][x | y]
  double (*hres)[4][RS] = reinterpret_cast<double (*)[4][RS]>(ring_smem + 2 * 2 * RR * RPITCH * 4);         // [RR]
  __shared__ double red[2][32];
  constexpr uint32_t OFS2 = SIGNED ? 0x80008000u : 0u;
  constexpr double kSplice = 4503599627370496.0;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int band = blockIdx.y;
  const uint32_t* A = reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(g.ref) + (int64_t)band * g.band_stride);
  const uint32_t* R = reinterpret_cast<const uint32_t*>(static_cast<const uint16_t*>(g.tst) + (int64_t)band * g.band_stride);
  const int64_t c_lo = RAD, c_hi = g.width - RAD;
  const int64_t wrow = g.width >> 1;                       // words per image row
  double acc_s = 0.0;
  unsigned acc_n = 0;
  double va[RR][4];                    // pending vertical sums: slot a started at the row with phase a
  const uint32_t raw_base = (uint32_t)__cvta_generic_to_shared(ring_smem);
  const int64_t ntasks = g.nseg * g.strips_x;
  for (int64_t task = blockIdx.x; task < ntasks; task += gridDim.x) {
    const int64_t seg = task / g.strips_x;
    const int strip = (int)(task - seg * g.strips_x);
    const int64_t r0 = g.r_lo + seg * g.seg_rows;
    const int64_t r1 = r0 + g.seg_rows < g.r_hi ? r0 + g.seg_rows : g.r_hi;
    const int64_t c0 = c_lo + (int64_t)strip * RS;
    const int64_t col = c0 + 4 * lane + warp;              // pass V: the column whose plane values sit at position t
    const bool col_ok = col < c_hi;
    const int n_in = (int)(r1 - r0) + 2 * RAD;             // input rows r0-5 .. r1+4
    const int nsteps = (n_in + RR - 1) / RR;
    const int64_t w0 = (c0 - RAD) >> 1;                    // first word of the strip in a row (c0 - 5 is even)
    const int64_t wmax = wrow - 1;
    // stage the words of step `st` (rows st*11 .. +10) straight into shared memory with cp.async: no registers are
    // held across the filter passes (a register prefetch spilled, and a spilled load is a synchronous load).
    // Rows / words past the data are clamped: they only feed outputs that are never counted.
    auto stage = [&](int st) {
      if (st < nsteps) {
#pragma unroll
        for (int k = 0; k < RPRE; ++k) {
          const int e = t + kRingThreads * k;
          if (e < RR * RWORDS) {
            const int rr = e / RWORDS, wc = e - rr * RWORDS;
            int64_t ri = r0 - RAD + (int64_t)st * RR + rr;
            ri = ri >= g.buf_rows ? g.buf_rows - 1 : ri;
            int64_t wi = w0 + wc;
            wi = wi > wmax ? wmax : wi;
            const uint32_t dst = raw_base + (uint32_t)((((st & 1) * 2 + 0) * RR + rr) * RPITCH + wc) * 4u;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(A + ri * wrow + wi) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + (uint32_t)(RR * RPITCH * 4)), "l"(R + ri * wrow + wi) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    __syncthreads();                   // the previous task's last pass H has read its rows
    stage(0);
    for (int st = 0; st < nsteps; ++st) {
      stage(st + 1);                   // buffer (st+1)&1 was last read by pass H of step st-1, before that step's barrier
      asm volatile("cp.async.wait_group 1;" ::: "memory");      // this thread's copies of step st have landed
      __syncthreads();                 // ... everybody's have; and pass V of the previous step has read the plane block
      // ---- pass H: items (row, group of 4 columns); row = warp + 4 k, group = lane
      {
        const uint32_t (*bx)[RPITCH] = raw[st & 1][0];
        const uint32_t (*by)[RPITCH] = raw[st & 1][1];
#pragma unroll 1
        for (int row = warp; row < RR; row += kRingThreads / 32) {
          // 14 pixels from pixel 4 * lane: seven words of each cube from word 2 * lane (8-byte aligned)
          uint32_t xw[7], yw[7];
          {
            const uint2* qx = reinterpret_cast<const uint2*>(&bx[row][2 * lane]);
            const uint2* qy = reinterpret_cast<const uint2*>(&by[row][2 * lane]);
            const uint2 a = qx[0], b = qx[1], c = qx[2];
            const uint2 d = qy[0], e = qy[1], f = qy[2];
            xw[0] = a.x; xw[1] = a.y; xw[2] = b.x; xw[3] = b.y; xw[4] = c.x; xw[5] = c.y; xw[6] = bx[row][2 * lane + 6];
            yw[0] = d.x; yw[1] = d.y; yw[2] = e.x; yw[3] = e.y; yw[4] = f.x; yw[5] = f.y; yw[6] = by[row][2 * lane + 6];
          }
          double x[14], y[14];
#pragma unroll
          for (int k = 0; k < 14; ++k) {
            const uint32_t wx = xw[k >> 1] ^ OFS2, wy = yw[k >> 1] ^ OFS2;
            x[k] = __hiloint2double(0x43300000, (int)((k & 1) ? wx >> 16 : wx & 0xffffu)) - kSplice;
            y[k] = __hiloint2double(0x43300000, (int)((k & 1) ? wy >> 16 : wy & 0xffffu)) - kSplice;
          }
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) {
            double a4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 14; ++k) {
              const double v = pl == 0 ? x[k] : pl == 1 ? y[k] : pl == 2 ? fma(x[k], x[k], y[k] * y[k]) : x[k] * y[k];
#pragma unroll
              for (int o = 0; o < 4; ++o)
                if (k - o >= 0 && k - o <= 2 * RAD) a4[o] = fma(g.w[k - o], v, a4[o]);
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) hres[row][pl][o * 32 + lane] = a4[o];
          }
        }
      }
      __syncthreads();
      // ---- pass V: thread = column; the rows of the step in order, static phases
#pragma unroll
      for (int rr = 0; rr < RR; ++rr) {
        const int sidx = st * RR + rr;
        if (sidx < n_in) {
          double h[4];
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) h[pl] = hres[rr][pl][t];
#pragma unroll
          for (int a = 0; a < RR; ++a) {
            const int tap = (rr - a + RR) % RR;
            const double wt = g.w[tap];
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) va[a][pl] = tap == 0 ? wt * h[pl] : fma(wt, h[pl], va[a][pl]);
          }
          if (sidx >= 2 * RAD && col_ok) {
            constexpr int dummy = 0; (void)dummy;
            const int a = (rr + 1) % RR;
            const double ux1 = va[a][0], uy1 = va[a][1], uq = va[a][2], up = va[a][3];
            const double ux = SIGNED ? ux1 - 32768.0 : ux1, uy = SIGNED ? uy1 - 32768.0 : uy1;
            const double vsum = uq - fma(ux1, ux1, uy1 * uy1);          // var x + var y (shift invariant)
            const double vxy = fma(-ux1, uy1, up);                     // covariance (shift invariant)
            const double num = fma(2.0, ux * uy, g.c1) * fma(2.0, vxy, g.c2);
            const double den = fma(ux, ux, fma(uy, uy, g.c1)) * (vsum + g.c2);
            double q;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(den));
            q = fma(fma(-den, q, 1.0), q, q);
            q = fma(fma(-den, q, 1.0), q, q);
            acc_s = fma(num, q, acc_s);
            acc_n += 1u;
          }
        }
      }
    }
  }
  acc_s = warp_sum_f64(acc_s);
  double cnt = warp_sum_f64((double)acc_n);
  __syncthreads();
  if (lane == 0) { red[0][warp] = acc_s; red[1][warp] = cnt; }
  __syncthreads();
  double tt[2] = {0.0, 0.0};
  if (t == 0) {
    for (int w = 0; w < kRingThreads / 32; ++w) { tt[0] += red[0][w]; tt[1] += red[1][w]; }
  }
  __syncthreads();
  double* const accs[2] = {g.sum_acc, g.cnt_acc};
  ordered_band_sum<2>(tt, g.scratch, static_cast<Workspace*>(g.workspace)->band_counter, accs, &red[0][0]);
}

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
  // the window: exp(-k^2 / (2 sigma^2)) / sum, sigma 1.5, radius 5 (scipy.ndimage.gaussian_filter, truncate 3.5)
  Taps taps;
  double sum = 0.0;
  for (int k = -RAD; k <= RAD; ++k) { taps.w[k + RAD] = std::exp(-0.5 / (1.5 * 1.5) * (double)(k * k)); sum += taps.w[k + RAD]; }
  for (int k = 0; k <= 2 * RAD; ++k) taps.w[k] /= sum;
  const double c1 = (0.01 * L) * (0.01 * L), c2 = (0.03 * L) * (0.03 * L);
  const int variant = ssim_variant();
  const bool ring_ok = p.dtype != DM_U8 && (p.width & 1) == 0 && (p.band_stride & 1) == 0 && p.width >= 2 * RAD + 2 &&
                       ((reinterpret_cast<uintptr_t>(p.ref) | reinterpret_cast<uintptr_t>(p.tst)) & 3) == 0;
  // blocks per band of the ring kernel: a fixed function of the band count (the summation order must not depend on
  // the device); 444 = 3 resident blocks on each of 148 SMs
  int nbx = (int)(444 / p.bands);
  nbx = nbx < 2 ? 2 : (nbx > kSsimBlocks ? kSsimBlocks : nbx);
  // variant 0 (auto) is the tiled kernel: with its prefetch fixed (r02r) it does the 10980^2 x 4 scene in 5.71 ms against
  // 5.86 ms for the ring kernel, and a 1024^2 x 4 tile in 70 us against 93 us (the ring kernel pays 10 warm-up rows per
  // segment, which small images cannot amortise).  The ring kernel runs on request (dm_ssim_variant(3)).
  if (variant == 3 && ring_ok) {
    RingArgs g;
    g.ref = p.ref; g.tst = p.tst; g.band_stride = p.band_stride; g.width = p.width; g.buf_rows = p.rows;
    g.r_lo = r_lo; g.r_hi = r_hi > r_lo ? r_hi : r_lo;
    for (int k = 0; k <= 2 * RAD; ++k) g.w[k] = taps.w[k];
    g.c1 = c1; g.c2 = c2;
    g.scratch = scratch; g.sum_acc = sum_acc; g.cnt_acc = cnt_acc; g.workspace = workspace;
    const int64_t ncols = p.width - 2 * RAD, nrows = g.r_hi - g.r_lo;
    g.strips_x = ncols > 0 ? (int)((ncols + RS - 1) / RS) : 0;
    // segments of 11 m - 10 rows (11 m input rows: whole steps); long ones pay their 10 warm-up rows once, short
    // ones give every block of a small image something to do
    int m = 24;
    while (m > 3 && (int64_t)g.strips_x * ((nrows + (RR * m - 2 * RAD) - 1) / (RR * m - 2 * RAD)) < 3 * (int64_t)nbx) m -= 3;
    g.seg_rows = RR * m - 2 * RAD;
    g.nseg = nrows > 0 ? (nrows + g.seg_rows - 1) / g.seg_rows : 0;
    const dim3 rgrid((unsigned)nbx, (unsigned)p.bands);
    if (p.dtype == DM_I16) {
      DM_CUDA(cudaFuncSetAttribute(ssim_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));
      ssim_ring_kernel<true><<<rgrid, kRingThreads, kRingSmem, s>>>(g);
    } else {
      DM_CUDA(cudaFuncSetAttribute(ssim_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingSmem));
      ssim_ring_kernel<false><<<rgrid, kRingThreads, kRingSmem, s>>>(g);
    }
    DM_LAUNCH_CHECK("ssim_ring");
    return DM_OK;
  }
  if (variant == 1) {
    StreamArgs g;
    g.ref = p.ref; g.tst = p.tst; g.band_stride = p.band_stride; g.width = p.width; g.buf_rows = p.rows;
    g.r_lo = r_lo; g.r_hi = r_hi > r_lo ? r_hi : r_lo;
    // 31-bit fixed-point taps, symmetric, summing to exactly 2^31 (the centre tap absorbs the rounding)
    int64_t tot = 0;
    for (int k = 0; k <= 2 * RAD; ++k) { g.w[k] = (uint32_t)std::llround(std::ldexp(taps.w[k], 31)); tot += g.w[k]; }
    g.w[RAD] = (uint32_t)((int64_t)g.w[RAD] + ((int64_t)1 << 31) - tot);
    for (int k = 0; k <= 2 * RAD; ++k) g.wv[k] = std::ldexp((double)g.w[k], -62);
    g.c1 = c1; g.c2 = c2;
    g.scratch = scratch; g.sum_acc = sum_acc; g.cnt_acc = cnt_acc; g.workspace = workspace;
    const int64_t ncols = p.width - 2 * RAD, nrows = g.r_hi - g.r_lo;
    g.strips_x = ncols > 0 ? (int)((ncols + 31) / 32) : 0;
    // blocks per band: a fixed function of the band count (the summation order must not depend on the device);
    // 444 = 3 resident blocks on each of 148 SMs
    int nbx = (int)(444 / p.bands);
    nbx = nbx < 2 ? 2 : (nbx > kSsimBlocks ? kSsimBlocks : nbx);
    // row segments: long enough that the 10 warm-up rows of a segment are small change, short enough that every
    // warp gets a few tasks
    int seg = 256;
    const int64_t warps = (int64_t)nbx * kStreamWarps;
    while (seg > 32 && (int64_t)g.strips_x * ((nrows + seg - 1) / seg) < 3 * warps) seg >>= 1;
    g.seg_rows = seg;
    g.nseg = nrows > 0 ? (nrows + seg - 1) / seg : 0;
    const dim3 sgrid((unsigned)nbx, (unsigned)p.bands);
    switch (p.dtype) {
      case DM_U8: ssim_stream_kernel<uint8_t><<<sgrid, kStreamWarps * 32, 0, s>>>(g); break;
      case DM_U16: ssim_stream_kernel<uint16_t><<<sgrid, kStreamWarps * 32, 0, s>>>(g); break;
      case DM_I16: ssim_stream_kernel<int16_t><<<sgrid, kStreamWarps * 32, 0, s>>>(g); break;
      default: return fail(DM_EARG, "dm_ssim_gauss: bad dtype");
    }
    DM_LAUNCH_CHECK("ssim_stream");
    return DM_OK;
  }
  // 296 blocks per band (four "waves" for a 4-band image): the hardware hands out blocks as slots free up, so many short
  // blocks balance better than one resident wave of long ones (74 per band, tried r02t: scene 6.01 ms against 5.71 ms)
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
