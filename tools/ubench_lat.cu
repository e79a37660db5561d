// Dependent-issue latency / single-warp throughput of the packed-arithmetic instructions on B200.
// Measurement tool only.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_lat tools/ubench_lat.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1;} } while (0)

constexpr int ITERS = 2048;

// NCH independent dp2a chains, one warp per SMSP (blockDim = 32*WPS*4): cycles per IDP
template <int NCH, int OP>
__global__ void k_chain(uint32_t* out, long long* cyc, uint32_t seed) {
  uint32_t a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = seed + threadIdx.x + i;
  uint32_t b = seed * 3 + threadIdx.x, c = seed ^ 0x01010101u;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (OP == 0) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == 1) asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a[i]) : "r"(b));
      if (OP == 2) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == 3) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) r += a[i];
  if (r == 0x12345678u) out[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// the pixel-group inner loop: per word 2 PRMT + 6 dp2a into NSETS x 6 accumulators, words from shared memory
template <int NSETS>
__global__ void k_pixel(uint32_t* out, long long* cyc, uint32_t seed) {
  __shared__ uint2 sm[2][64 * 45 + 16];
  for (int i = threadIdx.x; i < 64 * 45; i += blockDim.x) { sm[0][i] = make_uint2(seed + i, seed * 7 + i); sm[1][i] = make_uint2(seed + 3 * i, seed * 5 + i); }
  __syncthreads();
  uint32_t acc[NSETS][6];
#pragma unroll
  for (int s = 0; s < NSETS; ++s)
#pragma unroll
    for (int q = 0; q < 6; ++q) acc[s][q] = 0;
  const int tp = threadIdx.x & 63;
  long long t0 = clock64();
  for (int rep = 0; rep < 64; ++rep) {
    const uint2* xs = &sm[0][tp * 45];
    const uint2* ys = &sm[1][tp * 45];
#pragma unroll 9
    for (int j = 0; j < 45; ++j) {
      const uint2 xv = xs[j], yv = ys[j];
      const uint32_t xw[2] = {xv.x, xv.y}, yw[2] = {yv.x, yv.y};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const uint32_t x = xw[k], y = yw[k];
        uint32_t px, py;
        asm("prmt.b32 %0, %1, 0, 0x3120;" : "=r"(px) : "r"(x));
        asm("prmt.b32 %0, %1, 0, 0x3120;" : "=r"(py) : "r"(y));
        uint32_t* a = acc[(NSETS == 1) ? 0 : (NSETS == 2 ? k : (2 * (j & 1) + k) % NSETS)];
        asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[0]) : "r"(x), "r"(px));
        asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a[1]) : "r"(x), "r"(px));
        asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[2]) : "r"(y), "r"(py));
        asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a[3]) : "r"(y), "r"(py));
        asm("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[4]) : "r"(x), "r"(py));
        asm("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a[5]) : "r"(x), "r"(py));
      }
    }
  }
  long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int s = 0; s < NSETS; ++s)
#pragma unroll
    for (int q = 0; q < 6; ++q) r += acc[s][q];
  if (r == 0x12345678u) out[0] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NCH, int OP> int run_chain(const char* name, int threads, uint32_t* out, long long* cyc) {
  k_chain<NCH, OP><<<148, threads>>>(out, cyc, 1);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  printf("%-10s chains=%2d warps/SMSP=%d : %6.2f cycles per instruction per warp, %6.2f cycles per instr per SMSP\n", name, NCH, threads / 128,
         (double)h / (ITERS * NCH), (double)h / (ITERS * NCH) / (threads / 128));
  return 0;
}
template <int NSETS> int run_pixel(int threads, uint32_t* out, long long* cyc) {
  k_pixel<NSETS><<<148, threads>>>(out, cyc, 1);
  CK(cudaDeviceSynchronize());
  long long h; CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  const double words = 64.0 * 90;
  printf("pixel loop  sets=%d warps/SMSP=%d : %7.2f cycles per word per warp (9 instr: 6 IDP, 2 PRMT, 1/2+1/2 LDS.64)\n", NSETS, threads / 128, (double)h / words);
  return 0;
}

int main() {
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64));
  run_chain<1, 0>("IDP.2A", 128, out, cyc); run_chain<2, 0>("IDP.2A", 128, out, cyc); run_chain<3, 0>("IDP.2A", 128, out, cyc);
  run_chain<4, 0>("IDP.2A", 128, out, cyc); run_chain<6, 0>("IDP.2A", 128, out, cyc); run_chain<8, 0>("IDP.2A", 128, out, cyc);
  run_chain<12, 0>("IDP.2A", 128, out, cyc); run_chain<6, 0>("IDP.2A", 256, out, cyc); run_chain<6, 0>("IDP.2A", 512, out, cyc);
  run_chain<1, 1>("PRMT", 128, out, cyc); run_chain<4, 1>("PRMT", 128, out, cyc); run_chain<8, 1>("PRMT", 128, out, cyc);
  run_chain<1, 2>("VIMNMX", 128, out, cyc); run_chain<8, 2>("VIMNMX", 128, out, cyc);
  run_chain<1, 3>("IMAD", 128, out, cyc); run_chain<8, 3>("IMAD", 128, out, cyc);
  run_chain<1, 4>("IADD", 128, out, cyc); run_chain<8, 4>("IADD", 128, out, cyc);
  run_pixel<1>(128, out, cyc); run_pixel<2>(128, out, cyc); run_pixel<4>(128, out, cyc);
  run_pixel<1>(256, out, cyc); run_pixel<2>(256, out, cyc);
  run_pixel<1>(512, out, cyc);
  return 0;
}
