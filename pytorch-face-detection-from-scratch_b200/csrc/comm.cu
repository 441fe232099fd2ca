// Gradient all-reduce of the data-parallel train step over NVLink peer memory (one process per GPU).
//
// The reference trains on one GPU (train_model.py:47-53, Trainer(gpus=1)); its loss is a SUM over the batch
// (models/ModelMeta.py:173-176,215), so the data-parallel gradient is the plain sum of the shard gradients.
// The unit of exchange is the engine's flat fp32 gradient buffer (3 MB): far below the size where a ring pays,
// so this is a latency problem.  ONE kernel, no host involvement, capturable in the step's CUDA graph:
//
//   phase 1  scatter-push : rank r stores slice p of its local gradient into peer p's window, slot r   (NVLink writes)
//   barrier  (per-CTA flags in peer memory, st.release.sys / ld.acquire.sys)
//   phase 2  reduce + broadcast-push : rank r sums the `world` slots of its slice in rank order (every rank gets
//            bit-identical results: each element is summed exactly once, by its owner) and stores the sum into the
//            result window of every peer
//   barrier
//   phase 3  copy the result window back into the local gradient buffer (local HBM traffic only)
//
// CTA b of every rank touches the same element subset in every phase, so the two barriers are between CTA b of
// all ranks only -- no grid-wide synchronisation and no co-residency requirement.  No barrier is needed at entry:
// a peer's scatter window is free once the previous call's second barrier passed, and its result window is only
// written after the next call's first barrier, by which time that peer finished its phase 3.
#include <cstdlib>

#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kCommBlocks = 64;
constexpr int kCommThreads = 512;
// window layout (bytes): [flags: 2 phases x kCommBlocks x kMaxPeers u32 | epoch counters: kCommBlocks u32 | pad to 8 KB]
//                        [scatter slots: world x chunk floats][result: world x chunk floats]
constexpr size_t kFlagBytes = 8192;
constexpr int kErrWord = 2 * kCommBlocks * kMaxPeers + kCommBlocks;      // u32 index: 0 = ok, 1 + peer = timed out on peer
constexpr long long kSpinLimitClk = 40000000000LL;                       // ~20 s at 2 GHz: start-up skew between ranks is fine, a dead peer is not

struct Peers {
  unsigned char* win[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// L1-bypassing load of data that a peer GPU wrote into this GPU's memory
__device__ __forceinline__ float4 ld_peer_written(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void cta_barrier_across_ranks(const Peers& pp, int phase, int rank, int world, unsigned epoch,
                                                         unsigned* s_bad) {
  // The CTA's peer stores happen-before the bar.sync; the flag store below is a system-scope RELEASE by a thread that
  // passed that barrier, and release is cumulative -- the same publish pattern as NCCL's (barrier, then ONE thread
  // fences and posts), without a system fence in every thread.
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < world) {
    const int peer = threadIdx.x;
    // no separate __threadfence_system(): st.release.sys below IS the (cumulative) system-scope release of everything
    // that happened before the bar.sync; a second fence only added another NVLink round trip per barrier
    const size_t slot = (static_cast<size_t>(phase) * kCommBlocks + blockIdx.x) * kMaxPeers;
    st_release_sys(reinterpret_cast<unsigned*>(pp.win[peer]) + slot + rank, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(pp.win[rank]) + slot + peer;
    const long long t_start = clock64();
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
      if (clock64() - t_start > kSpinLimitClk) {        // a peer never arrived (died / not launched): do not hang the GPU
        reinterpret_cast<unsigned*>(pp.win[rank])[kErrWord] = 1u + peer;
        *s_bad = 1u + peer;          // this call's sums are invalid: phase 3 poisons the gradient instead of copying them
        break;
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCommThreads)
allreduce_sum_kernel(Peers pp, float* __restrict__ data, long n4, long chunk4, int rank, int world) {
  pdl_trigger();
  pdl_wait();                        // the local gradient is complete
  __shared__ unsigned s_epoch, s_bad;
  unsigned char* my = pp.win[rank];
  if (threadIdx.x == 0) {
    unsigned* ctr = reinterpret_cast<unsigned*>(my) + 2 * kCommBlocks * kMaxPeers + blockIdx.x;
    s_epoch = *ctr + 1;
    *ctr = s_epoch;
    // sticky: once an exchange of this window failed, the ranks' epochs may be out of step -- every later call is
    // poisoned too until the host notices (fd_comm_status) and rebuilds the windows
    s_bad = *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(my) + kErrWord);
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const long t0 = static_cast<long>(blockIdx.x) * kCommThreads + threadIdx.x;
  const long stride = static_cast<long>(gridDim.x) * kCommThreads;
  float4* data4 = reinterpret_cast<float4*>(data);

  // phase 1: scatter-push (destinations staggered so that the ranks do not all hit the same peer at once)
  for (int k = 0; k < world; ++k) {
    const int p = (rank + k) % world;
    const long lo = p * chunk4, len = min(chunk4, n4 - lo);
    float4* dst = reinterpret_cast<float4*>(pp.win[p] + kFlagBytes) + rank * chunk4;
    for (long j = t0; j < len; j += stride) dst[j] = data4[lo + j];
  }
  cta_barrier_across_ranks(pp, 0, rank, world, epoch, &s_bad);

  // phase 2: reduce my slice in rank order, push the sum into every peer's result window
  {
    const long lo = rank * chunk4, len = min(chunk4, n4 - lo);
    const float4* slots = reinterpret_cast<const float4*>(my + kFlagBytes);
    for (long j = t0; j < len; j += stride) {
      float4 acc = ld_peer_written(slots + j);
      for (int q = 1; q < world; ++q) {
        const float4 v = ld_peer_written(slots + q * chunk4 + j);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      for (int k = 0; k < world; ++k) {
        const int p = (rank + k) % world;
        reinterpret_cast<float4*>(pp.win[p] + kFlagBytes)[world * chunk4 + lo + j] = acc;
      }
    }
  }
  cta_barrier_across_ranks(pp, 1, rank, world, epoch, &s_bad);

  // phase 3: result window -> local gradient buffer.  After a barrier time-out the sums are garbage: the gradient is
  // POISONED with NaN instead, so that the optimizer step / the loss surface the failure in the same step (a silent
  // update with un-reduced gradients would let the replicas drift apart); the host reads the cause via fd_comm_status.
  const float4* res = reinterpret_cast<const float4*>(my + kFlagBytes) + world * chunk4;
  const bool bad = s_bad != 0u;
  const float qnan = __int_as_float(0x7fc00000);
  for (int q = 0; q < world; ++q) {
    const long lo = q * chunk4, len = min(chunk4, n4 - lo);
    for (long j = t0; j < len; j += stride)
      data4[lo + j] = bad ? make_float4(qnan, qnan, qnan, qnan) : ld_peer_written(res + lo + j);
  }
}

// ONE-SHOT variant for small buffers (the "late" region of the split exchange: ~0.7 MB): every rank pushes its WHOLE
// buffer into slot `rank` of every peer, ONE cross-rank barrier, then every rank sums the `world` contributions locally in
// rank order (bit-identical on all ranks).  One barrier instead of two and no second push / copy-back; (world - 1) x the
// bytes on the wire, which is why it is only used below kOneShotMaxN.  The slots are double buffered by call parity: a
// fast peer's pushes of call k+1 land in the other buffer while this rank may still be summing call k, and buffer k & 1 is
// only reused by call k+2, whose pushes follow barrier k+1 -- which this rank reaches after it finished call k.
constexpr long kOneShotMaxN = 256 * 1024;      // floats (1 MB)
__global__ void __launch_bounds__(kCommThreads)
allreduce_oneshot_kernel(Peers pp, float* __restrict__ data, long n4, int rank, int world) {
  pdl_trigger();
  pdl_wait();                        // the local gradient is complete
  __shared__ unsigned s_epoch, s_bad;
  unsigned char* my = pp.win[rank];
  if (threadIdx.x == 0) {
    unsigned* ctr = reinterpret_cast<unsigned*>(my) + 2 * kCommBlocks * kMaxPeers + blockIdx.x;
    s_epoch = *ctr + 1;
    *ctr = s_epoch;
    s_bad = *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(my) + kErrWord);
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const long t0 = static_cast<long>(blockIdx.x) * kCommThreads + threadIdx.x;
  const long stride = static_cast<long>(gridDim.x) * kCommThreads;
  float4* data4 = reinterpret_cast<float4*>(data);
  const long buf = static_cast<long>(epoch & 1u) * world * n4;          // float4 offset of this call's slot set
  for (int k = 1; k < world; ++k) {
    const int p = (rank + k) % world;
    float4* dst = reinterpret_cast<float4*>(pp.win[p] + kFlagBytes) + buf + rank * n4;
    for (long j = t0; j < n4; j += stride) dst[j] = data4[j];
  }
  cta_barrier_across_ranks(pp, 0, rank, world, epoch, &s_bad);
  const float4* slots = reinterpret_cast<const float4*>(my + kFlagBytes) + buf;
  const bool bad = s_bad != 0u;
  const float qnan = __int_as_float(0x7fc00000);
  for (long j = t0; j < n4; j += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {
      const float4 v = q == rank ? data4[j] : ld_peer_written(slots + q * n4 + j);
      if (q == 0) acc = v;
      else { acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    data4[j] = bad ? make_float4(qnan, qnan, qnan, qnan) : acc;
  }
}

inline long chunk4_of(long n, int world) { return ((n / 4) + world - 1) / world; }
inline bool use_oneshot(long n) {
  static const bool off = getenv("FD_COMM_TWO_SHOT") != nullptr;
  return !off && n <= kOneShotMaxN;
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" long fd_comm_window_bytes(long n, int world) {
  if (n <= 0 || n % 4 != 0 || world < 1 || world > kMaxPeers) return -1;
  if (use_oneshot(n)) return static_cast<long>(kFlagBytes) + 2L * world * n * 4;      // two slot sets of world x n floats
  return static_cast<long>(kFlagBytes) + 2L * world * chunk4_of(n, world) * 16;
}

extern "C" int fd_comm_alloc(long bytes, void** ptr) {
  if (!ptr || bytes <= 0) return FD_EINVAL;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));      // a plain cudaMalloc allocation: IPC-exportable
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { (void)cudaGetLastError(); cudaFree(p); return static_cast<int>(e); }
  *ptr = p;
  return FD_OK;
}

extern "C" int fd_comm_free(void* ptr) {
  if (!ptr) return FD_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) (void)cudaGetLastError();
  return e == cudaSuccess ? FD_OK : static_cast<int>(e);
}

extern "C" int fd_comm_export(void* ptr, unsigned char* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  if (!ptr || !handle64) return FD_EINVAL;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  memcpy(handle64, &h, 64);
  return FD_OK;
}

extern "C" int fd_comm_import(const unsigned char* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) return FD_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  *peer_ptr = p;
  return FD_OK;
}

extern "C" int fd_comm_release(void* peer_ptr) {
  if (!peer_ptr) return FD_OK;
  cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
  if (e != cudaSuccess) (void)cudaGetLastError();
  return e == cudaSuccess ? FD_OK : static_cast<int>(e);
}

extern "C" int fd_comm_error_offset(void) { return kErrWord * 4; }

// Status word of the LOCAL window: 0 = every barrier of every call completed; 1 + r = a wait on rank r ran into the
// 20 s spin limit (the peer died or never launched its kernel) and the sums of that call are garbage.  Synchronises the
// device (host-side health check, not on the data path).
extern "C" int fd_comm_status(void* window, int* status) {
  if (!window || !status) return FD_EINVAL;
  unsigned v = 0;
  cudaError_t e = cudaMemcpy(&v, static_cast<unsigned char*>(window) + kErrWord * 4, 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  *status = static_cast<int>(v);
  return FD_OK;
}

extern "C" int fd_allreduce_sum_f32(void* const* windows, int rank, int world, float* data, long n, void* stream) {
  return fd_allreduce_sum_f32_blocks(windows, rank, world, data, n, kCommBlocks, stream);
}

// Same exchange on `blocks` (1..64) thread blocks.  An exchange that runs BESIDE compute kernels (the early half of
// parallel.SplitAllReduce) uses few blocks: a block that waits for a late peer occupies its SM, and the persistent
// convolution kernels need a whole SM per CTA.  Every rank must pass the same `blocks` for a given window.
extern "C" int fd_allreduce_sum_f32_blocks(void* const* windows, int rank, int world, float* data, long n, int blocks,
                                           void* stream) {
  if (!windows || !data || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n <= 0) return FD_EINVAL;
  if (blocks < 1 || blocks > kCommBlocks) return FD_EINVAL;
  if (n % 4 != 0) return FD_EUNSUPPORTED;
  if (world == 1) return FD_OK;
  Peers pp = {};
  for (int i = 0; i < world; ++i) {
    if (!windows[i]) return FD_EINVAL;
    pp.win[i] = static_cast<unsigned char*>(windows[i]);
  }
  if (use_oneshot(n))
    launch_k(allreduce_oneshot_kernel, dim3(blocks), dim3(kCommThreads), 0, static_cast<cudaStream_t>(stream), pp, data, n / 4,
             rank, world);
  else
    launch_k(allreduce_sum_kernel, dim3(blocks), dim3(kCommThreads), 0, static_cast<cudaStream_t>(stream), pp, data,
             n / 4, chunk4_of(n, world), rank, world);
  count_launch();
  return launch_status();
}
