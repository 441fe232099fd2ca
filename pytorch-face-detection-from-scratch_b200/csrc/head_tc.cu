// Backward of the detection head (models/PoolResnet.py:83-89,100-102: Dropout2d -> KxK conv 64 -> 5 -> sigmoid) on the
// tensor cores.  The CUDA-core kernel in layers.cu needed 42 us for 0.3 GFLOP (8 % of the PoolResnet train step).
//
// With dz = dy * y * (1 - y) (5 x Ho x Wo values per image), both gradients are GEMMs over ONE im2col "patch tile" that
// is built in shared memory per image -- row = input pixel px, column k = (o, tap), o-major so that four consecutive
// columns are four consecutive taps of one dw[o][c][:] row (16-byte vector reductions in the drain):
//
//     P[px][o*K*K + tap] = dz[o][iy + pad - ky][ix + pad - kx]      (0 outside the output map / beyond the image)
//
//   dx[px][c]       = sum_k P[px][k] * Wt[c][k]          K-major A (P) x K-major B (weights):  M=128, N=64
//   dw[c][(tap,o)]  = sum_px x[px][c] * P[px][k]         MN-major A (x tile, TMA) x MN-major B (P):  M=64, N=64 per atom
//
// (the weight gradient re-indexed by INPUT pixel: dw[o][c][tap] = sum_q dz[o][q] x[q + tap][c] = sum_px x[px][c] dz[o][px - tap]).
// P rows are the usual 128-byte pixel rows (64 bf16, 128B swizzle), K = 5*K*K <= 192 = up to three 64-column atoms, so
// every operand descriptor is one already used by conv3x3_tc.cu / wgrad3x3_tc.cu.  dz and the weights enter in bf16
// (fp32 accumulation): relative error ~2e-3, the precision every other gradient of the backbone has.
// One CTA per image (persistent over the batch); Dropout2d multipliers are applied in the epilogues.
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kWorkWarps = 8;
constexpr int kThreads = (kWorkWarps + 1) * 32;    // warp 8: TMA + MMA issue
constexpr uint32_t kTile = 128 * 128;              // one 128-row operand tile

struct HeadParams {
  int B, H, W, Ho, Wo, npx, npo, natoms, nblk;
  const float* cs;          // [B,64] Dropout2d multipliers of the head input (nullable)
  const float* w;           // [5][64][K][K]
  const float* y;           // [B,5,Ho,Wo] sigmoid output
  const float* dy;
  __nv_bfloat16* dx;        // nullable
  __nv_bfloat16* dx2;       // nullable: dx * cs2 * (mask bit ? 1 : slope)
  const uint32_t* mask_bits;
  const float* cs2;
  float slope;
  float* dw;                // [5][64][K][K] accumulated
  float* dbias;             // [5] accumulated
};

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

// smem: [P: natoms x 2 x 16 KB][X: 2 x 16 KB][Wt: natoms x 8 KB][dz: 5*npo bf16][barriers]
template <int K, int PAD>
__global__ void __launch_bounds__(kThreads, 1)
head_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const HeadParams p) {
  constexpr int KK = K * K, NK = KK * 5;
  constexpr int NATOMS = (NK + 63) / 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sP = smem;                                   // [atom][blk][128 rows][128 B]
  uint8_t* sX = sP + NATOMS * 2 * kTile;
  uint8_t* sWt = sX + 2 * kTile;                        // [atom][64 rows c][128 B]
  uint16_t* sDz = reinterpret_cast<uint16_t*>(sWt + NATOMS * 8192);
  uint64_t* x_full = reinterpret_cast<uint64_t*>(sWt + NATOMS * 8192 + ((5 * p.npo * 2 + 15) & ~15));
  uint64_t* mma_done = x_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kWorkWarps) {
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      mbar_init(x_full, 1);
      mbar_init(mma_done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // one-time clear of the patch tiles: the columns >= 5*K*K of the last atom are never written
  for (uint32_t i = threadIdx.x * 16u; i < NATOMS * 2 * kTile; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sP + i) = make_uint4(0, 0, 0, 0);
  for (uint32_t i = threadIdx.x * 16u; i < NATOMS * 8192; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sWt + i) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  pdl_trigger();
  pdl_wait();
  // weights Wt[c][k] = w[o][c][tap], k = o*K*K + tap, bf16, K-major rows of 128 B per atom, 128B swizzle.  Read in w's own
  // memory order (coalesced), scattered into shared memory with 2-byte stores; k >= 5*K*K stays zero from the clear.
  for (int i = threadIdx.x; i < 5 * kC * KK; i += kThreads) {
    const int oc = i / KK, tap = i - oc * KK;
    const int o = oc / kC, c = oc - o * kC;
    const int k = o * KK + tap;
    *reinterpret_cast<uint16_t*>(sWt + (k >> 6) * 8192 + swz(static_cast<uint32_t>(c) * 128u + (k & 63) * 2u)) =
        __bfloat16_as_ushort(__float2bfloat16_rn(__ldg(p.w + i)));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_dx = tmem_base, tm_dw = tmem_base + 2 * kC;      // [2 blocks x 64] | [NATOMS x 64]

  int it = 0;
  for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
    // ---------------------------------------------------------------- phase 0: x tile (TMA), dz, dbias
    if (warp == kWorkWarps) {
      if (elect_one_sync()) {
        mbar_expect_tx(x_full, 2 * kTile);
        tma_load_2d(sX, &tm_x, x_full, 0, n * p.npx);
        tma_load_2d(sX + kTile, &tm_x, x_full, 0, n * p.npx + 128);
      }
      __syncwarp();
    } else {
      const float* yn = p.y + static_cast<size_t>(n) * 5 * p.npo;
      const float* dyn = p.dy + static_cast<size_t>(n) * 5 * p.npo;
      for (int i = threadIdx.x; i < 5 * p.npo; i += kWorkWarps * 32) {
        const float yv = yn[i];
        sDz[i] = __bfloat16_as_ushort(__float2bfloat16_rn(dyn[i] * yv * (1.f - yv)));
      }
      if (warp < 5) {            // dbias[o] += sum dz[o] (fp32 dz, not the rounded copy)
        float t = 0.f;
        for (int i = lane; i < p.npo; i += 32) {
          const float yv = yn[warp * p.npo + i];
          t += dyn[warp * p.npo + i] * yv * (1.f - yv);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) atomicAdd(p.dbias + warp, t);
      }
    }
    __syncthreads();
    // ---------------------------------------------------------------- phase 1: patch tile, one pixel row per thread
    if (warp < kWorkWarps) {
      const int px = threadIdx.x;
      const bool live = px < p.npx;
      const int iy = px / p.W, ix = px - iy * p.W;
      int rowoff[K];
      bool colok[K];
      int colx[K];
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const int oy = iy + PAD - t, ox = ix + PAD - t;
        rowoff[t] = (live && oy >= 0 && oy < p.Ho) ? oy * p.Wo : -1;
        colok[t] = ox >= 0 && ox < p.Wo;
        colx[t] = ox;
      }
      uint8_t* prow = sP + (px >> 7) * kTile;
      const uint32_t r128 = static_cast<uint32_t>(px & 127) * 128u;
#pragma unroll
      for (int a = 0; a < NATOMS; ++a) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t u[4] = {0, 0, 0, 0};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = a * 64 + j * 8 + i;
            if (k < NK) {
              const int o = k / KK, tap = k - o * KK, ky = tap / K, kx = tap - ky * K;
              uint32_t v = 0;
              if (rowoff[ky] >= 0 && colok[kx]) v = sDz[o * p.npo + rowoff[ky] + colx[kx]];
              u[i >> 1] |= v << ((i & 1) * 16);
            }
          }
          if (a * 64 + j * 8 < NK)       // chunks entirely beyond NK stay zero from the one-time clear
            *reinterpret_cast<uint4*>(prow + a * 2 * kTile + swz(r128 + j * 16u)) = make_uint4(u[0], u[1], u[2], u[3]);
        }
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---------------------------------------------------------------- phase 2: MMAs
    if (warp == kWorkWarps) {
      if (elect_one_sync()) {
        mbar_wait(x_full, it & 1);
        tc_fence_after();
        constexpr uint32_t idesc_dx = make_idesc_bf16(128, kC, 0, 0);
        constexpr uint32_t idesc_dw = make_idesc_bf16(64, kC, 1, 1);
        for (int blk = 0; blk < p.nblk; ++blk) {
#pragma unroll
          for (int a = 0; a < NATOMS; ++a) {
            const uint32_t a_lo = sdesc_lo(smem_u32(sP + (a * 2 + blk) * kTile), 16);
            const uint32_t b_lo = sdesc_lo(smem_u32(sWt + a * 8192), 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tm_dx + blk * kC, sdesc_sw128(a_lo + 2 * ks), sdesc_sw128(b_lo + 2 * ks), idesc_dx,
                        (a | ks) != 0 ? 1u : 0u);
          }
        }
#pragma unroll
        for (int a = 0; a < NATOMS; ++a) {
          for (int blk = 0; blk < p.nblk; ++blk) {
            const uint32_t x_lo = sdesc_lo(smem_u32(sX + blk * kTile), 1024);
            const uint32_t p_lo = sdesc_lo(smem_u32(sP + (a * 2 + blk) * kTile), 1024);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)      // 16 pixel rows = 2048 B per K step
              umma_bf16(tm_dw + a * kC, sdesc_sw128(x_lo + ks * 128), sdesc_sw128(p_lo + ks * 128), idesc_dw,
                        (blk | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(mma_done);
        mbar_wait(mma_done, it & 1);            // one thread polls; the workers sleep in the CTA barrier
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---------------------------------------------------------------- phase 3: epilogues
    if (warp < kWorkWarps) {
      const int q = warp & 3, h = warp >> 2;     // TMEM lane quadrant, 32-column half
      const float* csn = p.cs ? p.cs + n * kC : nullptr;
      const float* cs2n = p.cs2 ? p.cs2 + n * kC : nullptr;
      for (int blk = 0; blk < p.nblk; ++blk) {
        const int px = blk * 128 + q * 32 + lane;
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tm_dx + (static_cast<uint32_t>(q * 32) << 16) + blk * kC + h * 32, acc);
        tmem_ld_wait();
        if (px < p.npx) {
          const size_t gi = (static_cast<size_t>(n) * p.npx + px) * kC + h * 32;
          const uint32_t mk = p.dx2 ? __ldg(p.mask_bits + (gi >> 5)) : 0u;
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(acc[e]) * (csn ? __ldg(csn + h * 32 + e) : 1.f);
          if (p.dx) {
            uint4* d = reinterpret_cast<uint4*>(p.dx + gi);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              d[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
          if (p.dx2) {
#pragma unroll
            for (int e = 0; e < 32; ++e)
              v[e] = v[e] * (((mk >> e) & 1u) ? 1.f : p.slope) * (cs2n ? __ldg(cs2n + h * 32 + e) : 1.f);
            uint4* d = reinterpret_cast<uint4*>(p.dx2 + gi);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              d[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                                pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
          }
        }
        __syncwarp();
      }
      // weight gradient: M = 64 accumulator, row c lives in TMEM lane (c % 16) + 32 * (c / 16); columns = k of the atom
#pragma unroll
      for (int a = 0; a < NATOMS; ++a) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tm_dw + (static_cast<uint32_t>(q * 32) << 16) + a * kC + h * 32, acc);
        tmem_ld_wait();
        if (lane < 16) {
          const int c = q * 16 + lane;
          const float s = csn ? __ldg(csn + c) : 1.f;
#pragma unroll
          for (int e = 0; e < 32; e += 4) {        // 4 consecutive k = 4 consecutive taps of dw[o][c][:] (KK % 4 == 0)
            const int k = a * 64 + h * 32 + e;
            if (k < NK) {
              const int o = k / KK, tap = k - o * KK;
              float* dst = p.dw + (static_cast<size_t>(o) * kC + c) * KK + tap;
              if constexpr (KK % 4 == 0) {
                atomicAdd(reinterpret_cast<float4*>(dst),
                          make_float4(__uint_as_float(acc[e]) * s, __uint_as_float(acc[e + 1]) * s,
                                      __uint_as_float(acc[e + 2]) * s, __uint_as_float(acc[e + 3]) * s));
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int ki = k + i;
                  if (ki < NK) {
                    const int oi = ki / KK, ti = ki - oi * KK;
                    atomicAdd(p.dw + (static_cast<size_t>(oi) * kC + c) * KK + ti, __uint_as_float(acc[e + i]) * s);
                  }
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (warp == kWorkWarps) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Returns FD_EUNSUPPORTED for shapes this kernel is not instantiated for (the caller falls back to the CUDA-core kernel).
int head_bwd_tc(const fd_bf16* x, const float* chan_scale, const float* w, const float* y, const float* dy, int B, int H,
                int W, int C, int K, int pad, fd_bf16* dx, const uint32_t* mask_bits, const float* chan_scale2, float slope,
                fd_bf16* dx2, float* dw, float* dbias, cudaStream_t st) {
  if (C != kC || H * W > 256 || !((K == 6 && pad == 0) || (K == 3 && pad == 1))) return FD_EUNSUPPORTED;
  const int Ho = H + 2 * pad - K + 1, Wo = W + 2 * pad - K + 1;
  if (Ho <= 0 || Wo <= 0) return FD_EUNSUPPORTED;
  HeadParams p{};
  p.B = B; p.H = H; p.W = W; p.Ho = Ho; p.Wo = Wo; p.npx = H * W; p.npo = Ho * Wo;
  p.natoms = (K * K * 5 + 63) / 64;
  p.nblk = p.npx > 128 ? 2 : 1;
  p.cs = chan_scale; p.w = w; p.y = y; p.dy = dy;
  p.dx = reinterpret_cast<__nv_bfloat16*>(dx); p.dx2 = reinterpret_cast<__nv_bfloat16*>(dx2);
  p.mask_bits = mask_bits; p.cs2 = chan_scale2; p.slope = slope; p.dw = dw; p.dbias = dbias;
  CUtensorMap tm_x;
  const int rc = make_tmap_2d_bf16(&tm_x, x, B * p.npx, kC, 128, kC);
  if (rc != FD_OK) return rc;
  const size_t smem = static_cast<size_t>(p.natoms) * 2 * kTile + 2 * kTile + static_cast<size_t>(p.natoms) * 8192 +
                      ((5 * p.npo * 2 + 15) & ~15) + 64 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = B < sm_count() ? B : sm_count();
  cudaError_t e;
  if (K == 6) {
    e = set_max_dyn_smem(head_bwd_tc_kernel<6, 0>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(head_bwd_tc_kernel<6, 0>, dim3(grid), dim3(kThreads), smem, st, tm_x, p);
  } else {
    e = set_max_dyn_smem(head_bwd_tc_kernel<3, 1>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(head_bwd_tc_kernel<3, 1>, dim3(grid), dim3(kThreads), smem, st, tm_x, p);
  }
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

}  // namespace fd
