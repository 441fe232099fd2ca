// Stem of the MobilenetV3 backbone on the tensor cores: Conv2dSame 3x3 stride 2, 3 -> 16, + folded BatchNorm bias +
// Hardswish (feature_extractor.0-2 of the official archive; models/MobilenetV3Backbone.py:33-39).
//
// K = 27 is tiny, so the convolution becomes a GEMM over an im2col PATCH TILE built in shared memory (the construction
// of stem_s2_tc.cu, for N = 16): one output pixel = one 128-byte row, column k < 27 = input (c, ky, kx), k = 27 = the
// constant 1 (weight row k = 27 holds the bias, so the bias add happens inside the GEMM), k = 28..31 zero.
// M = 128 pixels (8 x 16 tile), N = 16, K = 32 = two tcgen05.mma.  The CUDA-core version spent ~1080 instructions per
// pixel (432 FMAs + per-tap bounds / address arithmetic) and ran at 0.19 of HBM; here a pixel costs ~160.
//
// The fp32 (or uint8) NCHW image patch of a tile (3 planes x 17 rows x 33 columns) is TMA-loaded with zero fill outside
// the image (= the TF "SAME" padding).  Two groups of 4 warps alternate tiles: while one group waits for its MMA and
// runs the epilogue (tcgen05.ld -> Hardswish -> bf16 -> 32-byte coalesced stores), the other builds its patch tile.
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kTH = 8, kTW = 16;                 // output tile (rows x columns) = 128 GEMM rows
constexpr int kPR = 2 * kTH + 1;                 // 17 patch rows
constexpr int kThreadsS = 9 * 32;                // 8 worker warps (2 groups) + 1 control warp

struct StemTcParams {
  int B, H, W, Ho, Wo, pad_t, pad_l, tiles_x, tiles_y;
  long num_tiles;
  int pc;                      // patch row pitch in ELEMENTS (fp32: 36, uint8: 48)
  uint32_t patch_bytes;        // 3 planes x 17 rows x pitch
  uint32_t patch_stride;       // patch_bytes rounded up to 128
  const float* w;              // [16][3][3][3] fp32, BatchNorm folded
  const float* bias;           // [16]
  __nv_bfloat16* out;          // [B,Ho,Wo,16]
};

__device__ __forceinline__ void tma_load_3d_s(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ float hswish(float v) { return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f); }

template <bool kU8>
__global__ void __launch_bounds__(kThreadsS, 2)
mbv3_stem_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ StemTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;                                  // 16 rows x 128 B (K-major, 128B swizzle)
  uint8_t* sA = smem + 2048;                           // [group][128 rows x 128 B]; 1024-aligned
  uint8_t* sP = sA + 2 * 16384;                        // [group][stage] patches
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 4 * p.patch_stride);
  uint64_t* patch_full = bars;                         // [2 groups][2 stages]
  uint64_t* a_ready = bars + 4;                        // [2]
  uint64_t* acc_full = bars + 6;                       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_x);
    for (int i = 0; i < 4; ++i) mbar_init(patch_full + i, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(a_ready + g, 128);
      mbar_init(acc_full + g, 1);
    }
    fence_barrier_init();
  }
  if (warp == 8) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  // weight tile: element (n, k) at n*128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2; k = c*9 + ky*3 + kx, k = 27: bias
  for (int i = threadIdx.x; i < 16 * 64; i += kThreadsS) {
    const int n = i >> 6, k = i & 63;
    float v = 0.f;
    if (k < 27) v = __ldg(p.w + n * 27 + k);
    else if (k == 27) v = __ldg(p.bias + n);
    *reinterpret_cast<__nv_bfloat16*>(sW + n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) * 2))) = __float2bfloat16_rn(v);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  pdl_trigger();
  pdl_wait();

  auto tile_of = [&](int g, long k) { return static_cast<long>(blockIdx.x) + (2 * k + g) * static_cast<long>(gridDim.x); };
  auto coords = [&](long tile, int& n, int& oy0, int& ox0) {
    n = static_cast<int>(tile / tiles_per_img);
    const int rem = static_cast<int>(tile - static_cast<long>(n) * tiles_per_img);
    const int ty = rem / p.tiles_x;
    oy0 = ty * kTH;
    ox0 = (rem - ty * p.tiles_x) * kTW;
  };

  if (warp == 8) {
    // ------------------------------------------------------------------ control: TMA patch loads + MMA issue
    if (elect_one_sync()) {
      auto load_patch = [&](int g, long k) {
        const long tile = tile_of(g, k);
        if (tile >= p.num_tiles) return;
        int n, oy0, ox0;
        coords(tile, n, oy0, ox0);
        const int s = static_cast<int>(k & 1);
        uint64_t* bar = patch_full + g * 2 + s;
        uint8_t* dst = sP + (g * 2 + s) * p.patch_stride;
        mbar_expect_tx(bar, p.patch_bytes);
        tma_load_3d_s(dst, &tm_x, bar, ox0 * 2 - p.pad_l, oy0 * 2 - p.pad_t, n * 3);   // one box = the 3 planes
      };
      for (int g = 0; g < 2; ++g) {
        load_patch(g, 0);
        load_patch(g, 1);
      }
      constexpr uint32_t idesc = make_idesc_bf16(128, 16, 0, 0);
      const uint32_t b_lo = sdesc_lo(smem_u32(sW), 16);
      // this CTA's tiles: blockIdx.x + j * gridDim.x, j = 0 .. J-1; tile j belongs to group j & 1 (its (j >> 1)-th tile)
      const long J = (p.num_tiles - static_cast<long>(blockIdx.x) + gridDim.x - 1) / gridDim.x;
      for (long j = 0; j < J; ++j) {
        const int g = static_cast<int>(j & 1);
        const long k = j >> 1;
        mbar_wait(a_ready + g, static_cast<uint32_t>(k & 1));
        tc_fence_after();
        const uint32_t a_lo = sdesc_lo(smem_u32(sA + g * 16384), 16);
        umma_bf16(tmem_base + static_cast<uint32_t>(g * 16), sdesc_sw128(a_lo), sdesc_sw128(b_lo), idesc, 0u);
        umma_bf16(tmem_base + static_cast<uint32_t>(g * 16), sdesc_sw128(a_lo + 2), sdesc_sw128(b_lo + 2), idesc, 1u);
        umma_commit(acc_full + g);
        load_patch(g, k + 2);              // stage k & 1 has been consumed by the builders (they arrived on a_ready)
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ workers: build the patch tile, then the epilogue
    const int g = warp >> 2;                       // group
    const int m = (warp & 3) * 32 + lane;          // GEMM row = TMEM lane (quadrant = warp % 4)
    const int ty = m >> 4, tx = m & 15;
    uint8_t* arow = sA + g * 16384 + m * 128;
    const uint32_t sw = static_cast<uint32_t>(m & 7);
    for (long k = 0;; ++k) {
      const long tile = tile_of(g, k);
      if (tile >= p.num_tiles) break;
      int n, oy0, ox0;
      coords(tile, n, oy0, ox0);
      const int s = static_cast<int>(k & 1);
      mbar_wait(patch_full + g * 2 + s, static_cast<uint32_t>((k >> 1) & 1));
      const uint8_t* patch = sP + (g * 2 + s) * p.patch_stride;
      float v[32];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int idx = (c * kPR + 2 * ty + ky) * p.pc + 2 * tx + kx;
            if (kU8) v[c * 9 + ky * 3 + kx] = __fdiv_rn(static_cast<float>(patch[idx]), 255.f);    // x / 255.0 (:52)
            else v[c * 9 + ky * 3 + kx] = reinterpret_cast<const float*>(patch)[idx];
          }
        }
      }
      v[27] = 1.f;
      v[28] = v[29] = v[30] = v[31] = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint4 u;
        u.x = pack_bf16x2(v[c4 * 8 + 0], v[c4 * 8 + 1]);
        u.y = pack_bf16x2(v[c4 * 8 + 2], v[c4 * 8 + 3]);
        u.z = pack_bf16x2(v[c4 * 8 + 4], v[c4 * 8 + 5]);
        u.w = pack_bf16x2(v[c4 * 8 + 6], v[c4 * 8 + 7]);
        *reinterpret_cast<uint4*>(arow + ((static_cast<uint32_t>(c4) ^ sw) << 4)) = u;
      }
      fence_proxy_async();                 // patch-tile writes (generic proxy) -> visible to the tensor core
      mbar_arrive(a_ready + g);
      mbar_wait(acc_full + g, static_cast<uint32_t>(k & 1));
      tc_fence_after();
      uint32_t acc[16];
      tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + static_cast<uint32_t>(g * 16), acc);
      tmem_ld_wait();
      tc_fence_before();                   // the next MMA of this group is issued after this thread's next arrive
      const int oy = oy0 + ty, ox = ox0 + tx;
      if (oy < p.Ho && ox < p.Wo) {
        uint32_t u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          u[j] = pack_bf16x2(hswish(__uint_as_float(acc[2 * j])), hswish(__uint_as_float(acc[2 * j + 1])));
        uint4* dst = reinterpret_cast<uint4*>(p.out + ((static_cast<long>(n) * p.Ho + oy) * p.Wo + ox) * 16);
        dst[0] = make_uint4(u[0], u[1], u[2], u[3]);
        dst[1] = make_uint4(u[4], u[5], u[6], u[7]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 32);
}

}  // namespace

int mbv3_stem_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int H, int W, int pad_t, int pad_l,
                 int Ho, int Wo, fd_bf16* out, cudaStream_t st) {
  StemTcParams p;
  p.B = B; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.pad_t = pad_t; p.pad_l = pad_l;
  p.tiles_x = (Wo + kTW - 1) / kTW;
  p.tiles_y = (Ho + kTH - 1) / kTH;
  p.num_tiles = static_cast<long>(B) * p.tiles_x * p.tiles_y;
  p.pc = x_is_u8 ? 48 : 36;                                   // >= 2*16 + 1 columns, row bytes a multiple of 16
  p.patch_bytes = static_cast<uint32_t>(3 * kPR * p.pc * (x_is_u8 ? 1 : 4));
  p.patch_stride = (p.patch_bytes + 127u) & ~127u;
  p.w = w; p.bias = bias;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  if ((x_is_u8 ? W : W * 4) % 16 != 0) return FD_EUNSUPPORTED;     // TMA: global row pitch must be a multiple of 16 bytes
  CUtensorMap tm_x;
  int rc = make_tmap_3d(&tm_x, x, x_is_u8 ? 1 : 4, x_is_u8, W, H, B * 3, p.pc, kPR, 3);
  if (rc != FD_OK) return rc;
  const size_t smem = 2048 + 2 * 16384 + 4 * static_cast<size_t>(p.patch_stride) + 256 + 1024;
  auto kern = x_is_u8 ? mbv3_stem_tc_kernel<true> : mbv3_stem_tc_kernel<false>;
  cudaError_t e = set_max_dyn_smem(kern, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const long cap = 2L * sm_count();
  const int grid = static_cast<int>(p.num_tiles < cap ? p.num_tiles : cap);
  e = launch_k(kern, dim3(grid), dim3(kThreadsS), smem, st, tm_x, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}

}  // namespace fd
