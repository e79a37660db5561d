// FP64-side instruction throughput on B200 (sm_100a): what the Gaussian-SSIM / SID / Sobel kernels can afford.
// Measurement tool only (not part of libdm_b200.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
// Reports lane-ops per clock and SM for: DFMA, IMAD.WIDE (64-bit accumulate), int->double conversions (s32 / u32 /
// s64), the reciprocal / rsqrt seeds (MUFU.RCP64H / RSQ64H), LDS.32, and MIXES of DFMA with the others issued from
// the same warp -- a mix that runs as fast as its slower half means the two halves use different pipes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 2048;
constexpr int NACC = 8;

enum Op { DFMA = 0, IMADW, I2D_S32, I2D_U32, I2D_S64, RCP64H, RSQ64H, LDS32, DFMA_IMADW, DFMA_I2D_S64, DFMA_LDS32, DFMA_I2D_S32,
          DFMA_IMADW_LDS, NOPS };
const char* kNames[NOPS] = {"DFMA", "IMAD.WIDE", "I2F.F64.S32", "I2F.F64.U32", "I2F.F64.S64", "MUFU.RCP64H", "MUFU.RSQ64H", "LDS.32",
                            "DFMA + IMAD.WIDE (1:1)", "DFMA + I2F.F64.S64 (8:1)", "DFMA + LDS.32 (1:1)", "DFMA + I2F.F64.S32 (4:1)",
                            "DFMA + IMAD.WIDE + LDS.32 (1:1:1)"};
// operations counted per inner trip (per accumulator)
const double kOpsPerTrip[NOPS] = {1, 1, 1, 1, 1, 1, 1, 1, 2, 1.125, 2, 1.25, 3};

template <int OP>
__global__ void __launch_bounds__(256) k_ops(uint32_t* out, uint32_t seed) {
  __shared__ uint32_t sh[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i * seed;
  __syncthreads();
  double d[NACC];
  uint64_t w[NACC];
  int32_t s[NACC];
  long long q[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { d[i] = 1.0 + (seed + threadIdx.x + i) * 1e-9; w[i] = seed + i + threadIdx.x * 977u; s[i] = seed * 3 + threadIdx.x + i; q[i] = ((long long)s[i] << 20) + i; }
  uint32_t b = seed * 3 + threadIdx.x, c = (seed ^ 0x01010101u) + threadIdx.x;
  double dsum = 0.0;
  const uint32_t* lp = sh + (threadIdx.x & 31);
  for (int it = 0; it < ITERS; ++it) {
    b = b * 5u + 0x9E37u;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      constexpr bool dfma = OP == DFMA || OP == DFMA_IMADW || OP == DFMA_I2D_S64 || OP == DFMA_LDS32 || OP == DFMA_I2D_S32 || OP == DFMA_IMADW_LDS;
      if (dfma) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(0.5));
      if (OP == IMADW || OP == DFMA_IMADW || OP == DFMA_IMADW_LDS) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((int)(b + i)), "r"(c));
      if (OP == I2D_S32 || (OP == DFMA_I2D_S32 && (i & 3) == 0)) { double t; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t) : "r"(s[i])); dsum += t; s[i] += 3; }
      if (OP == I2D_U32) { double t; asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(t) : "r"((uint32_t)s[i])); dsum += t; s[i] += 3; }
      if (OP == I2D_S64 || (OP == DFMA_I2D_S64 && i == 0)) { double t; asm volatile("cvt.rn.f64.s64 %0, %1;" : "=d"(t) : "l"(q[i])); d[i] = t; q[i] ^= (long long)it; }
      if (OP == RCP64H) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[i]));
      if (OP == RSQ64H) asm volatile("rsqrt.approx.ftz.f64 %0, %0;" : "+d"(d[i]));
      if (OP == LDS32 || OP == DFMA_LDS32 || OP == DFMA_IMADW_LDS) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(lp + ((i * 32 + it) & 511)))); b ^= v; }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += (uint32_t)d[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)q[i];
  r += (uint32_t)dsum + b;
  if (r == 0x12345678u) out[0] = r;
}

template <int OP>
double run_op(int blocks_per_sm, int sms, uint32_t* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_ops<OP><<<sms * blocks_per_sm, 256>>>(out, 1);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_ops<OP><<<sms * blocks_per_sm, 256>>>(out, 2);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return (double)sms * blocks_per_sm * 256 * ITERS * NACC * kOpsPerTrip[OP] / (ms * 1e-3);
}

template <int OP>
void report(int sms, double khz, uint32_t* out) {
  double best = 0;
  for (int bps : {2, 4, 8}) { const double r = run_op<OP>(bps, sms, out); if (r > best) best = r; }
  printf("%-40s %12.1f Gops/s(lane) %8.2f lane-ops/clk/SM\n", kNames[OP], best / 1e9, best / (sms * khz * 1e3));
}

// accuracy of the reciprocal / rsqrt seeds (what one or two Newton steps start from)
__global__ void k_seed_err(double* out) {
  double worst_rcp = 0.0, worst_rsq = 0.0;
  unsigned long long st = 0x9E3779B97F4A7C15ull * (threadIdx.x + 1 + blockIdx.x * blockDim.x);
  for (int i = 0; i < 4096; ++i) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    const double m = 1.0 + (double)(st >> 11) * (1.0 / 9007199254740992.0);       // [1, 2)
    const double x = ldexp(m, (int)((st >> 3) % 200) - 100);
    double r, q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
    worst_rcp = fmax(worst_rcp, fabs(r * x - 1.0));
    worst_rsq = fmax(worst_rsq, fabs(q * q * x - 1.0) * 0.5);
  }
  for (int o = 16; o > 0; o >>= 1) {
    worst_rcp = fmax(worst_rcp, __shfl_xor_sync(0xffffffffu, worst_rcp, o));
    worst_rsq = fmax(worst_rsq, __shfl_xor_sync(0xffffffffu, worst_rsq, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMax((unsigned long long*)out, __double_as_longlong(worst_rcp)); atomicMax((unsigned long long*)out + 1, __double_as_longlong(worst_rsq)); }
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  printf("device %s, %d SMs, clock %d kHz (rates below are per NOMINAL clock)\n", p.name, p.multiProcessorCount, khz);
  uint32_t* out; CK(cudaMalloc(&out, 64));
  const int sms = p.multiProcessorCount;
  report<DFMA>(sms, khz, out); report<IMADW>(sms, khz, out); report<I2D_S32>(sms, khz, out); report<I2D_U32>(sms, khz, out);
  report<I2D_S64>(sms, khz, out); report<RCP64H>(sms, khz, out); report<RSQ64H>(sms, khz, out); report<LDS32>(sms, khz, out);
  report<DFMA_IMADW>(sms, khz, out); report<DFMA_I2D_S64>(sms, khz, out); report<DFMA_LDS32>(sms, khz, out);
  report<DFMA_I2D_S32>(sms, khz, out); report<DFMA_IMADW_LDS>(sms, khz, out);
  double* err; CK(cudaMalloc(&err, 16)); CK(cudaMemset(err, 0, 16));
  k_seed_err<<<64, 256>>>(err);
  double h[2]; CK(cudaMemcpy(h, err, 16, cudaMemcpyDeviceToHost));
  printf("max relative error of the seeds over 6.7e7 random inputs: rcp.approx.ftz.f64 %.3e (2^%.1f)   rsqrt.approx.ftz.f64 %.3e (2^%.1f)\n",
         h[0], log2(h[0]), h[1], log2(h[1]));
  return 0;
}
