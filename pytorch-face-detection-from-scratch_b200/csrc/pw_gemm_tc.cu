// Pointwise (1x1) convolution of the MobilenetV3 backbone (reference models/MobilenetV3Backbone.py:33-39: timm
// tf_mobilenetv3_small_100; conv_pw / conv_pwl / ConvBnAct of the archive's efficientnet_blocks.py) as a tcgen05 GEMM
//
//     out[m, n] = act( sum_k x[m, k] * w[n, k] + bias[n] ) (+ residual[m, n])        m = pixel (B*H*W), k = Cin, n = Cout
//
// with BatchNorm (eval) folded into w / bias by the caller.  Channel counts are the network's own (16 ... 576, all
// multiples of 8), NOT padded in HBM: activations are NHWC bf16 rows of C*2 bytes, so the layer moves exactly the
// algorithmic bytes.  These layers are HBM-bound (AI 8..90 FLOP/B against a ridge of ~220): the tensor core is used
// because CUDA-core FMA throughput (72 TFLOP/s) would make them compute-bound, and because TMA + one MMA-issuing
// thread leaves every other warp free for the epilogue.
//
// One CTA per SM, persistent over (128-row M tile, N tile) work items:
//   warp 0   TMA producer: per 64-channel K slab one box {64 ch, 128 rows} of x (channels beyond Cin are zero-filled
//            by TMA) and one box {64, N16} of the packed weights into a ring of stages
//   warp 1   MMA issuer: per slab ceil(valid K / 16) tcgen05.mma (M=128, N=N16, K=16, both operands K-major SW128);
//            accumulators double-buffered in TMEM so the epilogue of tile i overlaps the loads + MMAs of tile i+1
//   warps 2-9  epilogue: tcgen05.ld -> + bias -> ReLU / Hardswish -> (+ residual) -> bf16 -> swizzled staging tile ->
//            TMA tensor store (clipped at Cout and at the last row)
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kPwThreads = 10 * 32;
constexpr int kPwEpiThreads = 8 * 32;
constexpr uint32_t kPwCtl = 4096;          // mbarriers + TMEM slot (first 1 KB), bias (up to 768 fp32)
constexpr int kPwMaxStages = 8;

struct PwParams {
  long M;
  int K, N, N16, nsplit, kslabs, last_ksteps, stages, nchunks, act, out_bufs;
  long num_tiles;
  uint32_t wslab_bytes, stage_bytes, tmem_cols, idesc;
  const float* bias;              // [nsplit * N16] fp32 (zero padded)
  const __nv_bfloat16* residual;  // [M, N] or null
};

__device__ __forceinline__ uint32_t pw_swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kPwEpiThreads) : "memory"); }

__device__ __forceinline__ float pw_act(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);                                       // ReLU
  if (act == 2) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);    // Hardswish: x * relu6(x + 3) / 6
  return v;
}

__global__ void __launch_bounds__(kPwThreads, 1)
pw_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
               const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ PwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full = bars;                          // [stages]
  uint64_t* empty = bars + kPwMaxStages;          // [stages]
  uint64_t* acc_full = bars + 2 * kPwMaxStages;   // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_bias = reinterpret_cast<float*>(smem + 1024);
  uint8_t* ring = smem + kPwCtl;
  uint8_t* staging = ring + static_cast<size_t>(p.stages) * p.stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < p.nsplit * p.N16; i += kPwThreads) s_bias[i] = __ldg(p.bias + i);
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx = 16384u + static_cast<uint32_t>(p.N16) * 128u;
      for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const long mt = tile / p.nsplit;
        const int nt = static_cast<int>(tile - mt * p.nsplit);
        const int m0 = static_cast<int>(mt * 128);
        for (int slab = 0; slab < p.kslabs; ++slab) {
          mbar_wait(empty + stage, phase ^ 1u);
          uint8_t* st = ring + static_cast<size_t>(stage) * p.stage_bytes;
          mbar_expect_tx(full + stage, tx);
          tma_load_2d(st, &tm_a, full + stage, slab * 64, m0);
          tma_load_2d(st + 16384, &tm_w, full + stage, 0, (nt * p.kslabs + slab) * p.N16);
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one_sync()) {
      uint32_t stage = 0, phase = 0;
      long it = 0;
      for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = static_cast<uint32_t>(it & 1), aphase = static_cast<uint32_t>((it >> 1) & 1);
        mbar_wait(acc_empty + buf, aphase ^ 1u);          // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * static_cast<uint32_t>(p.N16);
        for (int slab = 0; slab < p.kslabs; ++slab) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          uint8_t* st = ring + static_cast<size_t>(stage) * p.stage_bytes;
          const uint32_t a_lo = sdesc_lo(smem_u32(st), 16), b_lo = sdesc_lo(smem_u32(st + 16384), 16);
          const int ksteps = slab == p.kslabs - 1 ? p.last_ksteps : 4;
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d_tmem, sdesc_sw128(a_lo + 2 * k), sdesc_sw128(b_lo + 2 * k), p.idesc,
                      (slab > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty + stage);                     // stage reusable once these MMAs have read it
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
        umma_commit(acc_full + buf);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;                    // TMEM lane quadrant (hardware rule: warp % 4)
    const int half = (warp - 2) >> 2;          // the two warps of a quadrant take alternating 16-column chunks
    const bool leader = threadIdx.x == 64;
    const int nch16 = p.N16 >> 4;
    long it = 0;
    for (long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = static_cast<uint32_t>(it & 1), aphase = static_cast<uint32_t>((it >> 1) & 1);
      const long mt = tile / p.nsplit;
      const int nt = static_cast<int>(tile - mt * p.nsplit);
      const long m0 = mt * 128;
      const int n0 = nt * p.N16;
      // staging buffer (it % out_bufs) was last read by the TMA stores of tile it - out_bufs
      if (leader) {
        if (p.out_bufs == 2) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
      }
      epi_bar_sync();
      uint8_t* stg = staging + static_cast<size_t>(it % p.out_bufs) * p.nchunks * 16384;
      mbar_wait(acc_full + buf, aphase);
      tc_fence_after();
      const int row = q * 32 + lane;
      const long grow = m0 + row;
      for (int c16 = half; c16 < nch16; c16 += 2) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * static_cast<uint32_t>(p.N16) +
                               static_cast<uint32_t>(c16 * 16), acc);
        tmem_ld_wait();
        float v[16];
        const float* b = s_bias + n0 + c16 * 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = pw_act(__uint_as_float(acc[j]) + b[j], p.act);
        if (p.residual != nullptr && grow < p.M) {
          const int col = n0 + c16 * 16;
          const __nv_bfloat16* rp = p.residual + grow * p.N + col;
          if (col < p.N) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp));
            v[0] += bf16lo(u.x); v[1] += bf16hi(u.x); v[2] += bf16lo(u.y); v[3] += bf16hi(u.y);
            v[4] += bf16lo(u.z); v[5] += bf16hi(u.z); v[6] += bf16lo(u.w); v[7] += bf16hi(u.w);
          }
          if (col + 8 < p.N) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(rp + 8));
            v[8] += bf16lo(u.x); v[9] += bf16hi(u.x); v[10] += bf16lo(u.y); v[11] += bf16hi(u.y);
            v[12] += bf16lo(u.z); v[13] += bf16hi(u.z); v[14] += bf16lo(u.w); v[15] += bf16hi(u.w);
          }
        }
        uint4 u0, u1;
        u0.x = pack_bf16x2(v[0], v[1]);   u0.y = pack_bf16x2(v[2], v[3]);
        u0.z = pack_bf16x2(v[4], v[5]);   u0.w = pack_bf16x2(v[6], v[7]);
        u1.x = pack_bf16x2(v[8], v[9]);   u1.y = pack_bf16x2(v[10], v[11]);
        u1.z = pack_bf16x2(v[12], v[13]); u1.w = pack_bf16x2(v[14], v[15]);
        uint8_t* chunk = stg + static_cast<size_t>(c16 >> 2) * 16384;
        const uint32_t off = static_cast<uint32_t>(row) * 128u + static_cast<uint32_t>(c16 & 3) * 32u;
        *reinterpret_cast<uint4*>(chunk + pw_swz(off)) = u0;
        *reinterpret_cast<uint4*>(chunk + pw_swz(off + 16u)) = u1;
      }
      tc_fence_before();
      fence_proxy_async();
      epi_bar_sync();
      if (leader) {
        mbar_arrive(acc_empty + buf);                   // every epilogue thread's TMEM reads are complete
        for (int c = 0; c < p.nchunks; ++c) {
          if (n0 + c * 64 < p.N) tma_store_2d(&tm_out, stg + static_cast<size_t>(c) * 16384, n0 + c * 64, static_cast<int>(m0));
        }
        tma_store_commit();
      }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, p.tmem_cols);
}

// w fp32 [N][K] (nn.Conv2d 1x1 layout) * scale[n] (folded BatchNorm, nullable)  ->  bf16 [nsplit][kslabs][N16][64],
// zero padded: slab s of N tile t is one K-major 128B-swizzle-ready tile (the TMA load applies the swizzle).
__global__ void pw_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale, int N, int K, int N16,
                               int nsplit, int kslabs, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long total = static_cast<long>(nsplit) * kslabs * N16 * 64;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i & 63);
    long r = i >> 6;
    const int row = static_cast<int>(r % N16);
    r /= N16;
    const int slab = static_cast<int>(r % kslabs);
    const int t = static_cast<int>(r / kslabs);
    const int n = t * N16 + row, k = slab * 64 + c;
    float v = 0.f;
    if (n < N && k < K) v = w[static_cast<long>(n) * K + k] * (scale ? scale[n] : 1.f);
    out[i] = __float2bfloat16_rn(v);
  }
}

inline void pw_split(int N, int* N16, int* nsplit) {
  if (N <= 256) { *N16 = (N + 15) / 16 * 16; *nsplit = 1; return; }
  // several N tiles: the tile width must be a multiple of 64 so that the 64-column TMA stores of one tile never reach
  // into the columns of the next
  const int t = (N % 192 == 0) ? 192 : 128;
  *N16 = t;
  *nsplit = (N + t - 1) / t;
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" long fd_pw_packed_elems(int N, int K) {
  if (N <= 0 || K <= 0 || N % 8 || K % 8) return -1;
  int N16, nsplit;
  pw_split(N, &N16, &nsplit);
  return static_cast<long>(nsplit) * ((K + 63) / 64) * N16 * 64;
}
extern "C" int fd_pw_padded_n(int N) {
  if (N <= 0 || N % 8) return -1;
  int N16, nsplit;
  pw_split(N, &N16, &nsplit);
  return N16 * nsplit;
}

extern "C" int fd_pw_pack(const float* w, const float* scale, int N, int K, fd_bf16* out, void* stream) {
  if (!w || !out || N <= 0 || K <= 0) return FD_EINVAL;
  if (N % 8 || K % 8 || N > 768) return FD_EUNSUPPORTED;
  int N16, nsplit;
  pw_split(N, &N16, &nsplit);
  launch_k(pw_pack_kernel, dim3(64), dim3(256), 0, static_cast<cudaStream_t>(stream), w, scale, N, K, N16, nsplit,
           (K + 63) / 64, reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_pw_conv(const fd_bf16* x, const fd_bf16* w_packed, const float* bias_padded, long M, int K, int N,
                          int act, const fd_bf16* residual, fd_bf16* out, void* stream) {
  if (!x || !w_packed || !bias_padded || !out || M <= 0 || K <= 0 || N <= 0) return FD_EINVAL;
  if (K % 8 || N % 8 || N > 768 || act < 0 || act > 2 || M > 0x7fffff00L) return FD_EUNSUPPORTED;
  PwParams p;
  pw_split(N, &p.N16, &p.nsplit);
  p.M = M; p.K = K; p.N = N; p.act = act;
  p.kslabs = (K + 63) / 64;
  p.last_ksteps = ((K - (p.kslabs - 1) * 64) + 15) / 16;
  p.nchunks = (p.N16 + 63) / 64;
  p.wslab_bytes = static_cast<uint32_t>((p.N16 * 128 + 1023) / 1024 * 1024);
  p.stage_bytes = 16384u + p.wslab_bytes;
  const size_t cap = 225 * 1024;
  // two output staging tiles (the epilogue of tile i does not wait for the TMA stores of tile i-1) when at least three
  // ring stages still fit beside them
  p.out_bufs = (kPwCtl + 2 * static_cast<size_t>(p.nchunks) * 16384 + 1024 + 3 * static_cast<size_t>(p.stage_bytes) <= cap) ? 2 : 1;
  const size_t fixed = kPwCtl + static_cast<size_t>(p.out_bufs) * p.nchunks * 16384 + 1024;
  int stages = static_cast<int>((cap - fixed) / p.stage_bytes);
  if (stages > kPwMaxStages) stages = kPwMaxStages;
  if (stages < 2) return FD_EUNSUPPORTED;
  p.stages = stages;
  const long mtiles = (M + 127) / 128;
  p.num_tiles = mtiles * p.nsplit;
  const int cols = 2 * p.N16;
  p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  p.idesc = make_idesc_bf16(128, p.N16, 0, 0);
  p.bias = bias_padded;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  const size_t smem = fixed + static_cast<size_t>(stages) * p.stage_bytes;

  CUtensorMap tm_a, tm_w, tm_out;
  int rc = make_tmap_2d_bf16(&tm_a, x, static_cast<int>(M), K, 128, 64);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_w, w_packed, p.nsplit * p.kslabs * p.N16, 64, p.N16, 64);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_out, out, static_cast<int>(M), N, 128, 64);
  if (rc != FD_OK) return rc;
  cudaError_t e = set_max_dyn_smem(pw_gemm_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int nsm = sm_count();
  const int grid = p.num_tiles < nsm ? static_cast<int>(p.num_tiles) : nsm;
  e = launch_k(pw_gemm_kernel, dim3(grid), dim3(kPwThreads), smem, static_cast<cudaStream_t>(stream), tm_a, tm_w, tm_out, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}
