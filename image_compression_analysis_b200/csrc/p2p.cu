// NVLink peer-memory exchange of partial vectors (an alternative to the NCCL all-gather of the multi-GPU
// combine).  Every rank owns an exchange buffer
//     [world][capacity] partial vectors of `words` 8-byte words  |  world x uint64 arrival flags
// allocated with cudaMalloc and exported through CUDA IPC; after opening its peers' buffers a rank PUSHES its
// finished partial vectors straight into the slot every peer keeps for it (plain stores over NVLink from one
// small kernel, one CTA per destination), publishes the number of records it has delivered so far in the
// peer's flag word (system-scope release), and the combine kernel of each rank waits on its LOCAL flags
// (system-scope acquire) before it reduces its local copy in rank order.  No collective library call, no
// rendezvous on the host, nothing on the wire but the partial vectors themselves.

#include "dm_common.cuh"

namespace dm {

namespace {

constexpr int kMaxWorld = 16;

struct PushArgs {
  const unsigned long long* src;          // nrec x words, contiguous
  unsigned long long* dst[kMaxWorld];     // where peer r keeps this rank's records [rec0, rec0 + nrec)
  unsigned long long* flag[kMaxWorld];    // this rank's flag word in peer r's buffer
  long long total_words;
  unsigned long long flag_value;
  int world;
};

__global__ void __launch_bounds__(512)
p2p_push_kernel(PushArgs a) {
  const int r = blockIdx.x;               // destination
  unsigned long long* dst = a.dst[r];
  // 16-byte vectors when both sides allow it (they do: slots are multiples of 8 words)
  if (((reinterpret_cast<uintptr_t>(a.src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && (a.total_words & 1) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(a.src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (long long i = threadIdx.x; i < a.total_words / 2; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (long long i = threadIdx.x; i < a.total_words; i += blockDim.x) dst[i] = a.src[i];
  }
  __threadfence_system();                 // every thread's stores are visible system-wide ...
  __syncthreads();
  if (threadIdx.x == 0) {                 // ... before the flag that announces them
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[r]), "l"(a.flag_value) : "memory");
  }
}

// wait until every rank has delivered `need` records into this rank's buffer, then reduce records
// [rec0, rec0 + nrec) of the `world` local copies into out (layout and arithmetic of combine_partials_kernel)
__global__ void __launch_bounds__(256)
p2p_combine_kernel(const long long* __restrict__ gathered, const unsigned long long* flags, int world, unsigned long long need,
                   long long capacity, long long rec0, long long nrec, long long n_sum, long long n_max, long long n_f64,
                   long long* __restrict__ out, unsigned* status, unsigned long long timeout_ns) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    int good = 1;
    for (int r = 0; r < world && good; ++r) {
      for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + r) : "memory");
        if (v >= need) break;
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) { good = 0; break; }
      }
    }
    if (!good) atomicExch(status, 1u);
    ok = good;
  }
  __syncthreads();
  if (!ok) return;
  const long long len = n_sum + n_max + n_f64, total = nrec * len;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long rec = rec0 + i / len, off = i % len;
    const long long* g0 = gathered + rec * len + off;               // rank 0's copy; rank r is capacity*len further
    const long long rs = capacity * len;
    if (off < n_sum) {
      long long v = 0;
      for (int r = 0; r < world; ++r) v += __ldcg(g0 + r * rs);
      out[i] = v;
    } else if (off < n_sum + n_max) {
      long long v = __ldcg(g0);
      for (int r = 1; r < world; ++r) v = max(v, __ldcg(g0 + r * rs));
      out[i] = v;
    } else {
      double v = 0.0;
      for (int r = 0; r < world; ++r) v += __longlong_as_double(__ldcg(g0 + r * rs));     // rank order
      out[i] = __double_as_longlong(v);
    }
  }
}

}  // namespace

int p2p_alloc(int64_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes <= 0) return fail(DM_EARG, "dm_p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  DM_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), p);
  if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "dm_p2p_alloc"); }
  *ptr = p;
  return DM_OK;
}

int p2p_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return fail(DM_EARG, "dm_p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  DM_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DM_OK;
}

int p2p_close(void* ptr) { if (ptr) DM_CUDA(cudaIpcCloseMemHandle(ptr)); return DM_OK; }
int p2p_free(void* ptr) { if (ptr) DM_CUDA(cudaFree(ptr)); return DM_OK; }
int p2p_zero(void* ptr, int64_t bytes, cudaStream_t s) {
  if (!ptr || bytes < 0) return fail(DM_EARG, "dm_p2p_zero: bad arguments");
  DM_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, s));
  return DM_OK;
}

int launch_p2p_push(const void* src, int64_t total_words, void* const* peer_dst, void* const* peer_flag, int world,
                    uint64_t flag_value, cudaStream_t s) {
  if (!src || !peer_dst || !peer_flag || world < 1 || world > kMaxWorld || total_words < 0)
    return fail(DM_EARG, "dm_p2p_push: bad arguments (world <= %d)", kMaxWorld);
  PushArgs a{};
  a.src = static_cast<const unsigned long long*>(src);
  a.total_words = total_words; a.flag_value = flag_value; a.world = world;
  for (int r = 0; r < world; ++r) {
    if (!peer_dst[r] || !peer_flag[r]) return fail(DM_EARG, "dm_p2p_push: null peer pointer");
    a.dst[r] = static_cast<unsigned long long*>(peer_dst[r]);
    a.flag[r] = static_cast<unsigned long long*>(peer_flag[r]);
  }
  p2p_push_kernel<<<world, 512, 0, s>>>(a);
  DM_LAUNCH_CHECK("p2p_push");
  return DM_OK;
}

int launch_p2p_combine(const void* gathered, const void* flags, int world, uint64_t need, int64_t capacity, int64_t rec0,
                       int64_t nrec, int64_t n_sum, int64_t n_max, int64_t n_f64, void* out, uint32_t* status,
                       double timeout_s, cudaStream_t s) {
  if (!gathered || !flags || !out || !status) return fail(DM_EARG, "dm_p2p_combine: null pointer");
  if (world < 1 || world > kMaxWorld || nrec < 0 || rec0 < 0 || rec0 + nrec > capacity)
    return fail(DM_EARG, "dm_p2p_combine: bad sizes");
  const int64_t total = nrec * (n_sum + n_max + n_f64);
  if (total == 0) return DM_OK;
  int64_t grid = (total + 255) / 256;
  if (grid > 64) grid = 64;               // the exchange is a few hundred KB: keep the spinning blocks few
  p2p_combine_kernel<<<(unsigned)grid, 256, 0, s>>>(static_cast<const long long*>(gathered), static_cast<const unsigned long long*>(flags),
                                                   world, need, capacity, rec0, nrec, n_sum, n_max, n_f64,
                                                   static_cast<long long*>(out), status,
                                                   (unsigned long long)(timeout_s * 1e9));
  DM_LAUNCH_CHECK("p2p_combine");
  return DM_OK;
}

}  // namespace dm
