// Stem of the "standard" Resnet backbone (models/Resnet.py:64-70: Conv2d(3 -> 64, 3x3, stride 2, pad 1)), forward and
// weight gradient on the tensor cores.  The generic CUDA-core kernels in layers.cu took 451 us + 1974 us of the
// 3.4 ms Resnet train step (batch 16): 71 % of the step for 5 % of its FLOPs.
//
// K = Cin*3*3 = 27 is tiny, so the convolution is a GEMM over an im2col "patch tile" that is BUILT IN SHARED MEMORY
// (never materialised in HBM): one output pixel = one 128-byte row = 64 bf16 "channels", channel k < 27 = input
// (c, ky, kx), k = 27 = the constant 1, k > 27 = 0 -- exactly the pixel-row layout (128B swizzle) of every other
// tensor-core kernel in this library, so the operand descriptors are the ones of conv3x3_tc.cu / wgrad3x3_tc.cu:
//
//   forward : y[px, co]  = patch[px, k] * Wt[co, k]          K-major A (patch) and B (weights), M=128, N=64, 2 K-steps
//             Wt[co][27] = bias[co]: the ones column adds the bias inside the GEMM
//   wgrad   : D[co, k]   = sum_px g[px, co] * patch[px, k]   MN-major A (g tile, TMA) and B (patch), M=64, N=64, K=16 px
//             D[co][27] = sum_px g[px, co] = dbias: again the ones column
//
// One task = one output row of one image (Wo <= 256 pixels = two 128-row blocks).  The 9 input rows (3 channels x 3
// kernel rows) are staged as bf16 with coalesced float4 / uchar4 loads (uint8: /255 fused, PoolResnet.py:95), the
// patch rows are gathered from them with conflict-free 2-byte shared loads.  Phases of a task are serial inside a
// CTA; two CTAs per SM overlap each other.  The weight gradient accumulates in TMEM over all tasks of a CTA and is
// drained once (64 x 28 atomics per CTA).
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kBuildWarps = 8;
constexpr int kThreads = (kBuildWarps + 1) * 32;   // warps 0..7: staging / patch build / epilogue; warp 8: TMA + MMA issue
constexpr int kRowElems = 512;                     // staged bf16 row: element j holds input column j - 2
constexpr int kKTaps = 27;                         // Cin * 3 * 3
constexpr uint32_t kTile = 128 * 128;              // one 128-pixel operand tile

struct S2Params {
  int B, Hin, Win, Ho, Wo, ntask;
  const void* x;
  const float* w;       // [64][3][3][3] fp32
  const float* bias;    // [64]
  __nv_bfloat16* y;     // [B,Ho,Wo,64]
  float* dw;            // [64][27] accumulated
  float* dbias;         // [64] accumulated (nullable)
};

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// x / 255.0f, correctly rounded (see stem_tc.cu: bit-identical to the IEEE division for all 256 inputs)
__device__ __forceinline__ float div255(unsigned char b) {
  const float x = static_cast<float>(b), r = 1.0f / 255.0f;
  const float q = x * r;
  return fmaf(fmaf(-q, 255.0f, x), r, q);
}

// 9 input rows (c, ky) of output row oy -> sRows[r][kRowElems] bf16, element j = input column j - 2 (zeros outside).
// All global loads of a thread are issued before the first conversion (5 independent 16-byte loads in flight per
// thread; a rolled load -> convert -> store loop exposed one HBM round trip per iteration: 4000 clk per task).
constexpr int kStageIters = (9 * (kRowElems / 4) + kBuildWarps * 32 - 1) / (kBuildWarps * 32);      // 5 for Win <= 512
template <typename TIn>
__device__ __forceinline__ void stage_rows(const S2Params& p, int n, int oy, uint8_t* sRows, int tid) {
  const int quads = p.Win >> 2;
  const int total = 9 * quads;
  uint4 raw[kStageIters];
  int dst[kStageIters];
#pragma unroll
  for (int t = 0; t < kStageIters; ++t) {
    const int idx = tid + t * kBuildWarps * 32;
    dst[t] = -1;
    raw[t] = make_uint4(0, 0, 0, 0);
    if (idx < total) {
      const int r = idx / quads, qd = idx - r * quads;
      const int c = r / 3, ky = r - c * 3;
      const int iy = 2 * oy + ky - 1;
      dst[t] = r * (kRowElems * 2) + (4 * qd + 2) * 2;
      if (iy >= 0 && iy < p.Hin) {
        const size_t off = ((static_cast<size_t>(n) * 3 + c) * p.Hin + iy) * p.Win + 4 * qd;
        if constexpr (sizeof(TIn) == 4) {
          raw[t] = __ldg(reinterpret_cast<const uint4*>(static_cast<const float*>(p.x) + off));
        } else {
          raw[t].x = __ldg(reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(p.x) + off));
          raw[t].y = 1u;            // marks "loaded" for the uint8 path (a zero word is a valid pixel quad)
        }
      } else {
        dst[t] = -2 - dst[t];       // out-of-image row: store zeros
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kStageIters; ++t) {
    if (dst[t] == -1) continue;
    uint32_t a0 = 0, a1 = 0;
    int d = dst[t];
    if (d >= 0) {
      float v0, v1, v2, v3;
      if constexpr (sizeof(TIn) == 4) {
        v0 = __uint_as_float(raw[t].x); v1 = __uint_as_float(raw[t].y);
        v2 = __uint_as_float(raw[t].z); v3 = __uint_as_float(raw[t].w);
      } else {
        const uint32_t u = raw[t].x;
        v0 = div255(u & 0xFF); v1 = div255((u >> 8) & 0xFF); v2 = div255((u >> 16) & 0xFF); v3 = div255(u >> 24);
      }
      a0 = pack_bf16x2(v0, v1);
      a1 = pack_bf16x2(v2, v3);
    } else {
      d = -2 - d;
    }
    uint32_t* q = reinterpret_cast<uint32_t*>(sRows + d);
    q[0] = a0;
    q[1] = a1;
  }
}

// Patch rows of both 128-pixel blocks of the task.  Thread = (pixel m of the block, chunk pair JP): 16 of the 32 live
// k columns -> two 16-byte chunks.  Pixels >= Wo get all-zero rows (incl. the ones column).
template <int JP>
__device__ __forceinline__ void build_patch(const S2Params& p, const uint8_t* sRows, uint8_t* sA, int m) {
#pragma unroll
  for (int blk = 0; blk < 2; ++blk) {
    const int px = blk * 128 + m;
    const bool valid = px < p.Wo;
    const uint16_t* base = reinterpret_cast<const uint16_t*>(sRows) + 2 * px + 1;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      uint32_t u[4] = {0, 0, 0, 0};
      if (valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = (2 * JP + jj) * 8 + i;
          uint32_t v = 0;
          if (k < kKTaps) {
            const int c = k / 9, ky = (k % 9) / 3, kx = k % 3;
            v = base[(c * 3 + ky) * kRowElems + kx];
          } else if (k == kKTaps) {
            v = 0x3F80u;            // bf16 1.0: bias (forward) / dbias (weight gradient) ride on this column
          }
          u[i >> 1] |= v << ((i & 1) * 16);
        }
      }
      *reinterpret_cast<uint4*>(sA + blk * kTile + swz(static_cast<uint32_t>(m) * 128u + (2 * JP + jj) * 16u)) =
          make_uint4(u[0], u[1], u[2], u[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward
// smem: [W 8 KB][A0 16 KB][A1 16 KB][rows 9 KB][barriers]
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 2)
stem_s2_fwd_kernel(const __grid_constant__ CUtensorMap tm_y, const S2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = smem + 8192;
  uint8_t* sRows = sA + 2 * kTile;
  uint64_t* acc_full = reinterpret_cast<uint64_t*>(sRows + 9 * kRowElems * 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kBuildWarps) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_y);
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // zero A tiles (k >= 32 stays zero forever) and the staged rows (left pad / tail stay zero forever)
  for (uint32_t i = threadIdx.x * 16u; i < 2 * kTile + 9 * kRowElems * 2; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  pdl_trigger();
  pdl_wait();
  // weights: Wt[co][k] bf16, K-major rows of 128 B, 128B swizzle; k = 27 carries the bias
  for (int idx = threadIdx.x; idx < kC * 8; idx += kThreads) {
    const int co = idx >> 3, j = idx & 7;
    uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = j * 8 + i;
      float v = 0.f;
      if (k < kKTaps) v = __ldg(p.w + co * kKTaps + k);
      else if (k == kKTaps) v = __ldg(p.bias + co);
      u[i >> 1] |= static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v))) << ((i & 1) * 16);
    }
    *reinterpret_cast<uint4*>(sW + swz(static_cast<uint32_t>(co) * 128u + j * 16u)) = make_uint4(u[0], u[1], u[2], u[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);

  int it = 0;
  for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
    const int n = task / p.Ho, oy = task - n * p.Ho;
    if (warp < kBuildWarps) {
      stage_rows<TIn>(p, n, oy, sRows, threadIdx.x);
    } else {
      if (elect_one_sync()) tma_store_wait_read<0>();   // the previous task's output stores have drained the A tiles
      __syncwarp();
    }
    __syncthreads();
    if (warp < kBuildWarps) {
      if (warp < 4) build_patch<0>(p, sRows, sA, threadIdx.x & 127);
      else build_patch<1>(p, sRows, sA, threadIdx.x & 127);
      fence_proxy_async();          // patch rows (generic proxy) -> visible to the tensor core
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kBuildWarps) {
      if (elect_one_sync()) {
        const uint32_t b_lo = sdesc_lo(smem_u32(sW), 16);
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint32_t a_lo = sdesc_lo(smem_u32(sA + blk * kTile), 16);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)         // k < 32: two K steps; columns 32..63 of the tiles are zero
            umma_bf16(tmem_base + blk * kC, sdesc_sw128(a_lo + 2 * ks), sdesc_sw128(b_lo + 2 * ks), idesc, ks);
        }
        umma_commit(acc_full);
        mbar_wait(acc_full, it & 1);             // one thread polls; the others sleep in the CTA barrier
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // epilogue: the patch tiles are dead (MMAs complete) -> they become the dense, swizzled staging tiles of the output
    // rows, which leave through TMA stores (a row-per-thread global store touches 32 lines per instruction)
    if (warp < kBuildWarps) {
      const int q = warp & 3, h = warp >> 2;     // TMEM lane quadrant (= warp % 4), 32-channel half
#pragma unroll
      for (int blk = 0; blk < 2; ++blk) {
        const uint32_t row = static_cast<uint32_t>(q * 32 + lane) * 128u;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + blk * kC + h * 32, acc);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(acc[8 * i + 0]), __uint_as_float(acc[8 * i + 1]));
          u.y = pack_bf16x2(__uint_as_float(acc[8 * i + 2]), __uint_as_float(acc[8 * i + 3]));
          u.z = pack_bf16x2(__uint_as_float(acc[8 * i + 4]), __uint_as_float(acc[8 * i + 5]));
          u.w = pack_bf16x2(__uint_as_float(acc[8 * i + 6]), __uint_as_float(acc[8 * i + 7]));
          *reinterpret_cast<uint4*>(sA + blk * kTile + swz(row + (h * 4 + i) * 16u)) = u;
        }
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();               // accumulators and staged rows are free; the A tiles hold the output rows
    tc_fence_after();
    if (warp == kBuildWarps) {
      if (elect_one_sync()) {
        tma_store_4d(&tm_y, sA, 0, 0, oy, n);                        // pixels 0..127 of the row
        if (p.Wo > 128) tma_store_4d(&tm_y, sA + kTile, 0, 128, oy, n);   // pixels >= Wo: clipped
        tma_store_commit();
      }
      __syncwarp();
    }
  }
  if (warp == kBuildWarps) {
    if (elect_one_sync()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBuildWarps) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// smem: [A0 16 KB][A1 16 KB][G0 16 KB][G1 16 KB][rows 9 KB][barriers]
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 2)
stem_s2_wgrad_kernel(const __grid_constant__ CUtensorMap tm_g, const S2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sG = smem + 2 * kTile;
  uint8_t* sRows = sG + 2 * kTile;
  uint64_t* g_full = reinterpret_cast<uint64_t*>(sRows + 9 * kRowElems * 2);
  uint64_t* mma_done = g_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kBuildWarps) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_g);
      mbar_init(g_full, 1);
      mbar_init(mma_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  for (uint32_t i = threadIdx.x * 16u; i < 2 * kTile; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  for (uint32_t i = threadIdx.x * 16u; i < 9 * kRowElems * 2; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sRows + i) = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  constexpr uint32_t idesc = make_idesc_bf16(64, kC, 1, 1);      // both operands MN-major (pixel rows)

  int it = 0;
  for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
    const int n = task / p.Ho, oy = task - n * p.Ho;
    if (warp == kBuildWarps) {
      if (elect_one_sync()) {
        // g rows of this output row: 256 consecutive pixel rows of the flattened [B*Ho*Wo, 64] tensor (the rows beyond
        // Wo belong to the next output row / are zero-filled past the end; their patch rows are zero)
        const int row0 = task * p.Wo;
        mbar_expect_tx(g_full, 2 * kTile);
        tma_load_2d(sG, &tm_g, g_full, 0, row0);
        tma_load_2d(sG + kTile, &tm_g, g_full, 0, row0 + 128);
      }
      __syncwarp();
    } else {
      stage_rows<TIn>(p, n, oy, sRows, threadIdx.x);
    }
    __syncthreads();
    if (warp < kBuildWarps) {
      if (warp < 4) build_patch<0>(p, sRows, sA, threadIdx.x & 127);
      else build_patch<1>(p, sRows, sA, threadIdx.x & 127);
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kBuildWarps) {
      if (elect_one_sync()) {
        mbar_wait(g_full, it & 1);
        tc_fence_after();
        const uint32_t nblk = p.Wo > 128 ? 2u : 1u;
        for (uint32_t blk = 0; blk < nblk; ++blk) {
          const uint32_t g_lo = sdesc_lo(smem_u32(sG + blk * kTile), 1024);
          const uint32_t a_lo = sdesc_lo(smem_u32(sA + blk * kTile), 1024);
#pragma unroll
          for (uint32_t ks = 0; ks < 8; ++ks)   // 16 pixel rows = 2048 B per K step
            umma_bf16(tmem_base, sdesc_sw128(g_lo + ks * 128), sdesc_sw128(a_lo + ks * 128), idesc,
                      (static_cast<uint32_t>(it) | blk | ks) != 0 ? 1u : 0u);
        }
        umma_commit(mma_done);
        mbar_wait(mma_done, it & 1);             // operands free again; one thread polls
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  // drain: M = 64 accumulator, row co lives in TMEM lane (co % 16) + 32 * (co / 16); columns = k
  if (warp < 4 && it > 0) {
    uint32_t acc[32];
    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16), acc);
    tmem_ld_wait();
    if (lane < 16) {
      const int co = warp * 16 + lane;
#pragma unroll
      for (int k = 0; k < kKTaps; ++k) atomicAdd(p.dw + co * kKTaps + k, __uint_as_float(acc[k]));
      if (p.dbias) atomicAdd(p.dbias + co, __uint_as_float(acc[kKTaps]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kBuildWarps) tmem_dealloc(tmem_base, 64);
}

inline bool s2_shape_ok(int Cin, int Hin, int Win, int C, int K, int stride, int pad) {
  return Cin == 3 && K == 3 && stride == 2 && pad == 1 && C == kC && Win % 4 == 0 && Win >= 4 &&
         (Win + 2 * pad - K) / stride + 1 <= 256 && Win + 2 <= kRowElems - 2 && Hin >= 1;
}

inline S2Params s2_params(const void* x, int B, int Hin, int Win) {
  S2Params p{};
  p.B = B; p.Hin = Hin; p.Win = Win;
  p.Ho = (Hin + 2 - 3) / 2 + 1;
  p.Wo = (Win + 2 - 3) / 2 + 1;
  p.ntask = B * p.Ho;
  p.x = x;
  return p;
}

}  // namespace

// Called from layers.cu's fd_stem_fwd / fd_stem_wgrad after the stride-8 stem declined; FD_EUNSUPPORTED = not the
// 3x3 / stride-2 / pad-1 / 3 -> 64 stem.
int stem_s2_fwd_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win, int C,
                   int K, int stride, int pad, fd_bf16* y, cudaStream_t st) {
  if (!s2_shape_ok(Cin, Hin, Win, C, K, stride, pad)) return FD_EUNSUPPORTED;
  S2Params p = s2_params(x, B, Hin, Win);
  p.w = w; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  CUtensorMap tm_y;
  {
    const int rc = make_tmap_nhwc_bf16(&tm_y, y, B, p.Ho, p.Wo, kC, 128, 1);     // box = 128 pixels of one output row
    if (rc != FD_OK) return rc;
  }
  const size_t smem = 8192 + 2 * kTile + 9 * kRowElems * 2 + 64 + 1024;
  const int grid = min(p.ntask, 2 * sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_s2_fwd_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_s2_fwd_kernel<uint8_t>, dim3(grid), dim3(kThreads), smem, st, tm_y, p);
  } else {
    e = set_max_dyn_smem(stem_s2_fwd_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_s2_fwd_kernel<float>, dim3(grid), dim3(kThreads), smem, st, tm_y, p);
  }
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

int stem_s2_wgrad_tc(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                     int stride, int pad, float* dw, float* dbias, cudaStream_t st) {
  if (!s2_shape_ok(Cin, Hin, Win, C, K, stride, pad)) return FD_EUNSUPPORTED;
  S2Params p = s2_params(x, B, Hin, Win);
  p.dw = dw; p.dbias = dbias;
  CUtensorMap tm_g;
  const int rc = make_tmap_2d_bf16(&tm_g, g, B * p.Ho * p.Wo, kC, 128, kC);
  if (rc != FD_OK) return rc;
  const size_t smem = 4 * kTile + 9 * kRowElems * 2 + 64 + 1024;
  const int grid = min(p.ntask, 2 * sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_s2_wgrad_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_s2_wgrad_kernel<uint8_t>, dim3(grid), dim3(kThreads), smem, st, tm_g, p);
  } else {
    e = set_max_dyn_smem(stem_s2_wgrad_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_s2_wgrad_kernel<float>, dim3(grid), dim3(kThreads), smem, st, tm_g, p);
  }
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

}  // namespace fd
