// Instruction-throughput and read-stream microbenchmarks for B200 (sm_100a).
// Measurement tool only (not part of libdm_b200.so): tells which integer SIMD-in-word
// instructions the fused-stats kernel can afford per sample pair at HBM speed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NACC = 8;

enum Op { IADD = 0, LOP, IMAD, IMADWIDE, IDP2A, IDP4A, VMNMX2, VMNMX3, PRMT, SHF, POPC, ATOMS, DFMA, FFMA, MIX, NOPS };
const char* kNames[NOPS] = {"IADD3", "LOP3", "IMAD", "IMAD.WIDE(64acc)", "IDP.2A", "IDP.4A", "VIMNMX.U16x2", "VIMNMX3.U16x2",
                            "PRMT", "SHF", "POPC", "ATOMS(lane-private)", "DFMA", "FFMA", "MIX(2 VIMNMX+IADD+9 IDP+2 PRMT)"};

template <int OP>
__global__ void __launch_bounds__(1024) k_ops(uint32_t* out, uint32_t seed) {
  __shared__ uint32_t sh[256 * 8];
  uint32_t a[NACC];
  uint64_t w[NACC];
  double d[NACC];
  float f[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { a[i] = seed + threadIdx.x * 7 + i; w[i] = a[i]; d[i] = a[i]; f[i] = a[i]; }
  for (int i = threadIdx.x; i < 256 * 8; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  uint32_t b = seed * 3 + threadIdx.x, c = seed ^ 0x01010101u;
  uint32_t e0 = seed, e1 = seed;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == IMADWIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b), "r"(c));
      if (OP == IDP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == IDP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b), "r"(c));
      if (OP == VMNMX2) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
      if (OP == VMNMX3) a[i] = __vimax3_u16x2(a[i], b, c);
      if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x3120;" : "+r"(a[i]) : "r"(b));
      if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(a[i]) : "r"(b));
      if (OP == POPC) asm volatile("{ .reg .u32 t; popc.b32 t, %0; add.u32 %0, t, %1; }" : "+r"(a[i]) : "r"(b));
      if (OP == ATOMS) atomicAdd(&sh[((a[i] + it) & 7) * 256 + (threadIdx.x & 255)], 1u);
      if (OP == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(0.5));
      if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(1.0001f), "f"(0.5f));
    }
    if (OP == MIX) {
      // the per-word instruction mix of the packed fused-stats kernel, two words per trip
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        uint32_t x = b + it + i, y = c + it * 3 + i;
        uint32_t mx, mn;
        asm volatile("max.u16x2 %0, %1, %2;" : "=r"(mx) : "r"(x), "r"(y));
        asm volatile("min.u16x2 %0, %1, %2;" : "=r"(mn) : "r"(x), "r"(y));
        uint32_t dd = mx - mn;
        uint32_t px, py;
        asm volatile("prmt.b32 %0, %1, %1, 0x3120;" : "=r"(px) : "r"(x));
        asm volatile("prmt.b32 %0, %1, %1, 0x3120;" : "=r"(py) : "r"(y));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[0]) : "r"(dd), "r"(0x0101u));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[1]) : "r"(x), "r"(0x0101u));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[2]) : "r"(y), "r"(0x0101u));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[3]) : "r"(x), "r"(px));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a[4]) : "r"(x), "r"(px));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[5]) : "r"(y), "r"(py));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(a[6]) : "r"(y), "r"(py));
        asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(a[7]) : "r"(x), "r"(py));
        asm volatile("dp2a.hi.u32.u32 %0, %1, %2, %0;" : "+r"(e0) : "r"(x), "r"(py));
        e1 = __vimax3_u16x2(e1, dd, mx);
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += a[i] + (uint32_t)w[i] + (uint32_t)d[i] + (uint32_t)f[i];
  r += sh[threadIdx.x & 255] + e0 + e1;
  if (r == 0x12345678u) out[0] = r;
}

// MIX at a given residency (warps per SM): how much thread-level parallelism the packed mix needs
double run_mix(int blocks_per_sm, int threads, int sms, uint32_t* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_ops<MIX><<<sms * blocks_per_sm, threads>>>(out, 1);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_ops<MIX><<<sms * blocks_per_sm, threads>>>(out, 2);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return (double)sms * blocks_per_sm * threads * ITERS * 2.0 / (ms * 1e-3);
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
  const double per_iter = OP == MIX ? 2.0 : (double)NACC;   // MIX: words per trip
  return (double)sms * blocks_per_sm * 256 * ITERS * per_iter / (ms * 1e-3);
}

// ---- read-stream kernels: two arrays, 16-byte streaming loads, xor-reduce --------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) k_read2(const uint4* __restrict__ a, const uint4* __restrict__ b, int64_t n, uint32_t* out) {
  uint32_t acc = 0;
  const int64_t stride = (int64_t)gridDim.x * 256 * UNROLL;
  for (int64_t i = (int64_t)blockIdx.x * 256 * UNROLL + threadIdx.x; i < n; i += stride) {
    uint4 x[UNROLL], y[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t j = i + (int64_t)u * 256;
      if (j < n) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[u].x), "=r"(x[u].y), "=r"(x[u].z), "=r"(x[u].w) : "l"(a + j));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(y[u].x), "=r"(y[u].y), "=r"(y[u].z), "=r"(y[u].w) : "l"(b + j));
      } else { x[u] = make_uint4(0, 0, 0, 0); y[u] = x[u]; }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc ^= x[u].x ^ x[u].y ^ x[u].z ^ x[u].w ^ y[u].x ^ y[u].y ^ y[u].z ^ y[u].w;
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int UNROLL>
void run_read(const uint4* a, const uint4* b, int64_t n, int grid, uint32_t* out) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int w = 0; w < 2; ++w) k_read2<UNROLL><<<grid, 256>>>(a, b, n, out);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_read2<UNROLL><<<grid, 256>>>(a, b, n, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("read2  unroll=%d grid=%5d : %8.1f GB/s (best of 5, %.3f ms)\n", UNROLL, grid, 2.0 * n * 16 / (best * 1e-3) / 1e9, best);
}

// ---- TMA bulk-copy streaming: persistent CTA per SM, one producer lane, S-stage ring, consumers only
// wait/arrive.  Shows how many bytes must be in flight per SM for cp.async.bulk to reach HBM speed.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k_tma_stream(const char* a, const char* b, int64_t ntiles, int tile_bytes, int stages,
                                                   uint32_t* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[16], empty_bar[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty_bar[s])), "r"(3));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t acc = 0;
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % stages;
    const uint32_t ph = (uint32_t)((it / stages) & 1);
    unsigned char* dst = smem + (size_t)s * 2 * tile_bytes;
    if (warp == 0) {
      if (lane == 0) {
        uint32_t done = 0;
        while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&empty_bar[s])), "r"(ph ^ 1u) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(2 * tile_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(a + t * (int64_t)tile_bytes), "r"(tile_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst + tile_bytes)), "l"(b + t * (int64_t)tile_bytes), "r"(tile_bytes), "r"(smem_u32(&full_bar[s])) : "memory");
      }
    } else {
      uint32_t done = 0;
      while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&full_bar[s])), "r"(ph) : "memory");
      acc ^= reinterpret_cast<const uint32_t*>(dst)[tid];
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

void run_tma(const char* a, const char* b, int64_t bytes, int tile_bytes, int stages, int sms, uint32_t* out) {
  const int64_t ntiles = bytes / tile_bytes;
  const size_t smem = (size_t)stages * 2 * tile_bytes;
  CK(cudaFuncSetAttribute(k_tma_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k_tma_stream<<<sms, 128, smem>>>(a, b, ntiles, tile_bytes, stages, out);
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; ++r) {
    CK(cudaEventRecord(e0));
    k_tma_stream<<<sms, 128, smem>>>(a, b, ntiles, tile_bytes, stages, out);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  printf("tma_stream tile=2x%6d B stages=%2d (%6.1f KB ring/SM): %8.1f GB/s\n", tile_bytes, stages, smem / 1024.0,
         2.0 * ntiles * tile_bytes / (best * 1e-3) / 1e9);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  const int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  uint32_t* out; CK(cudaMalloc(&out, 64));
  printf("%-36s %14s %14s\n", "op", "Gops/s(lane)", "lane-ops/clk/SM@1.965GHz");
#define RUN(OP) { double r = run_op<OP>(8, sms, out); printf("%-36s %14.1f %14.2f\n", kNames[OP], r / 1e9, r / sms / 1.965e9); }
  RUN(IADD) RUN(LOP) RUN(IMAD) RUN(IMADWIDE) RUN(IDP2A) RUN(IDP4A) RUN(VMNMX2) RUN(VMNMX3) RUN(PRMT) RUN(SHF) RUN(POPC) RUN(ATOMS) RUN(DFMA) RUN(FFMA) RUN(MIX)
  for (int threads : {128, 256, 352, 512, 608, 704, 1024}) {
    double r = run_mix(1, threads, sms, out);
    printf("MIX 1 block x %4d threads/SM (%2d warps): %8.2f words/clk/SM@1.965GHz -> %6.2f TB/s of pair bytes\n", threads,
           threads / 32, r / sms / 1.965e9, r * 8 / 1e12);
  }
  const int64_t bytes = 1ll << 30;   // per array
  uint4 *a, *b;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  const int64_t n = bytes / 16;
  for (int tb : {5760, 11520, 23040, 46080})
    for (int st : {2, 3, 4, 6, 8, 16})
      if ((size_t)st * 2 * tb <= 200 * 1024) run_tma((const char*)a, (const char*)b, bytes, tb, st, sms, out);
  for (int mult : {2, 4, 8, 16, 32}) {
    run_read<1>(a, b, n, sms * mult, out);
    run_read<2>(a, b, n, sms * mult, out);
    run_read<4>(a, b, n, sms * mult, out);
  }
  return 0;
}
