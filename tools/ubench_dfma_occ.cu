// DFMA throughput against occupancy and instruction-level parallelism on B200 (sm_100a): how many warps per SM
// (and independent chains per thread) the FP64 pipe needs before it runs at its 64 lanes per clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_dfma_occ tools/ubench_dfma_occ.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 4096;
template <int NACC, bool REGOPS>
__global__ void __launch_bounds__(128) k(double* out, double a, double b) {
  extern __shared__ double pad[];
  double d[NACC], m = a + threadIdx.x * 1e-12, c = b + threadIdx.x * 1e-13;
#pragma unroll
  for (int i = 0; i < NACC; ++i) d[i] = 1.0 + (threadIdx.x + i) * 1e-9;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (REGOPS) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(m), "d"(c));          // three register operands
      else asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(0.5));       // two immediates / constants
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += d[i];
  if (r == 12345.678) out[0] = r + pad[0];
}
template <int NACC, bool REGOPS>
double run(int warps_per_sm, int sms, double* out) {
  // one 128-thread block = 4 warps; occupancy is pinned by the number of blocks launched (one wave)
  const int blocks = sms * warps_per_sm / 4;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<NACC, REGOPS><<<blocks, 128>>>(out, 1.0000001, 0.5);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k<NACC, REGOPS><<<blocks, 128>>>(out, 1.0000001, 0.5);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return (double)blocks * 128 * ITERS * NACC / (ms * 1e-3);
}
template <int NACC, bool REGOPS>
void row(int sms, double khz, double* out) {
  printf("%-10s chains/thread %2d :", REGOPS ? "reg,reg,reg" : "reg,imm,imm", NACC);
  for (int w : {4, 8, 12, 16, 24, 32, 48, 64}) printf(" %5.1f", run<NACC, REGOPS>(w, sms, out) / (sms * khz * 1e3));
  printf("\n");
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  double* out; CK(cudaMalloc(&out, 64));
  const int sms = p.multiProcessorCount;
  printf("%s: DFMA lane-ops per NOMINAL clock (%d kHz) and SM; columns = resident warps per SM: 4 8 12 16 24 32 48 64\n", p.name, khz);
  row<1, false>(sms, khz, out); row<2, false>(sms, khz, out); row<4, false>(sms, khz, out); row<8, false>(sms, khz, out); row<16, false>(sms, khz, out);
  row<1, true>(sms, khz, out); row<4, true>(sms, khz, out); row<8, true>(sms, khz, out); row<16, true>(sms, khz, out); row<44, true>(sms, khz, out);
  return 0;
}
