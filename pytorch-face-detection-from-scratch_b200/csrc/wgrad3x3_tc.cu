// Weight gradient of the 3x3 / stride 1 / pad 1, 64 -> 64 convolution on tcgen05.
//
// Replaces the weight-gradient half of autograd's conv2d backward for
// models/PoolResnet.py:35,37 of the reference.
//
//   dW[t][ci][co] = sum_pixels  xpad[p + off_t][ci] * g[p][co],   off_t = ky*Wp + kx
//
// is a GEMM whose reduction (K) dimension is the pixel index.  Both operands live in shared memory
// exactly as TMA delivers NHWC tiles -- one pixel per 128-byte row, channels contiguous -- which is
// the "MN-major" operand form of tcgen05.mma, so no transposes are needed:
//   A (M side) = the halo input tile, start address shifted by off_t rows; two taps are stacked
//                into one M=128 instruction by using the shift difference as the distance between the
//                two 64-channel atoms (descriptor LBO),
//   B (N side) = the gradient tile (64 output channels).
// The halo column of g (x == W) is zero-filled by TMA and the K padding rows are kept zero, so
// junk rows contribute nothing.  5 stacked-tap MMAs per 16 pixels; the fp32 accumulators
// (5 x 64 TMEM columns) live in TMEM across ALL tiles of the persistent CTA and are reduced
// into global memory once per CTA with vector fp32 reductions.  The bias gradient (column sums
// of g) is accumulated by the otherwise idle epilogue warps from the same smem tiles.
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;

struct WgradParams {
  int B, H, W, R, TW, Wp, tiles_w, tiles_per_img, num_tiles, ksteps;
  uint32_t x_bytes, g_bytes;          // bytes delivered by the two TMA boxes
  uint32_t x_buf_bytes, g_buf_bytes;  // reserved per stage (multiples of 1024)
  float* dw;                          // [9][ci][co] fp32, accumulated
  float* dbias;                       // [co] fp32, accumulated (nullable)
  int flags, dbg;
  // multi-problem launch: x / g hold `nprob` stacked [B,H,W,C] tensors; problem q accumulates into
  // dw + q * dw_stride and dbias + q * dbias_stride.  A CTA never straddles two problems.
  int nprob, tiles_per_prob, ctas_per_prob, per;
  long dw_stride, dbias_stride;
  uint32_t bar_off;                   // barriers live behind max(stages, drain staging tiles)
  int dw_row0, dw_rows_per_prob;      // rows of the [.,64] fp32 view of dw: first row, rows between problems
};

// Optional timestamps of CTA 0 (FD_WGRAD_TIMING=1).
__device__ unsigned long long g_wgrad_dbg[16];
#define FD_WTS(slot) do { if (p.dbg && blockIdx.x == 0) g_wgrad_dbg[slot] = clock64(); } while (0)

// One gradient map per column strip: its W extent ends at the strip, so the two junk columns of the
// K index (x >= TW) are zero-filled by TMA and contribute nothing.
constexpr int kMaxStrips = 8;
struct GMaps { CUtensorMap m[kMaxStrips]; };

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ GMaps gmaps,
                   const __grid_constant__ CUtensorMap tm_dw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const uint32_t stage_bytes = p.x_buf_bytes + p.g_buf_bytes;
  uint8_t* sStage = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* full = bars + 0;      // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sBias = reinterpret_cast<float*>(bars + 6);  // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) FD_WTS(0);

  // K-padding rows of g and the over-read tail of x are never written by TMA and must stay zero.
  for (uint32_t i = threadIdx.x * 16u; i < 2 * stage_bytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sStage + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&gmaps.m[0]);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1 + 4);  // MMA commit + the four column-sum warps
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) FD_WTS(1);

  // contiguous chunk of tiles (of ONE problem) for this CTA
  const int prob = blockIdx.x / p.ctas_per_prob;
  const int chunk = blockIdx.x - prob * p.ctas_per_prob;
  const int tile_begin = prob * p.tiles_per_prob + min(p.tiles_per_prob, chunk * p.per);
  const int tile_end = prob * p.tiles_per_prob + min(p.tiles_per_prob, (chunk + 1) * p.per);
  float* const dbias_out = p.dbias ? p.dbias + prob * p.dbias_stride : nullptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / p.tiles_per_img;
        const int rem = tile - n * p.tiles_per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * p.R, w0 = tw * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, p.x_bytes + p.g_bytes);
        tma_load_4d(sStage + s * stage_bytes, &tm_x, full + s, 0, w0 - 1, h0 - 1, n);
        tma_load_4d(sStage + s * stage_bytes + p.x_buf_bytes, &gmaps.m[tw], full + s, 0, 0, h0, n);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kC, 1, 1);  // both operands MN-major
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t x_addr = smem_u32(sStage + s * stage_bytes);
        const uint32_t g_lo = sdesc_lo(x_addr + p.x_buf_bytes, 1024);
        // tap pairs (0,1) (2,3) (4,5) (6,7) (7,8): the last pair recomputes tap 7 (dropped in the
        // epilogue) so that every instruction is a regular two-atom M=128 MMA.  Pair j reads the
        // halo tile at row offset off0 with the second atom (off1 - off0) rows further (LBO).
        uint32_t a_lo[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int t0 = (j < 4) ? 2 * j : 7, t1 = t0 + 1;
          const int off0 = (t0 / 3) * p.Wp + (t0 % 3);
          const int off1 = (t1 / 3) * p.Wp + (t1 % 3);
          a_lo[j] = sdesc_lo(x_addr + static_cast<uint32_t>(off0 * 128), static_cast<uint32_t>((off1 - off0) * 128));
        }
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t bd = sdesc_sw128(g_lo + ks * 128);      // 16 pixel rows = 2048 B per K step
#pragma unroll
          for (int j = 0; j < 5; ++j)
            umma_bf16(tmem_base + j * kC, sdesc_sw128(a_lo[j] + ks * 128), bd, idesc, (it | ks) != 0 ? 1u : 0u);
        }
        umma_commit(empty + s);
        if (tile + 1 == tile_end) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps: (1) bias gradient from the smem g tiles while the MMAs run
    const int et = threadIdx.x - 64;      // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum = 0.f;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      if (p.dbias) {
        const uint8_t* g = sStage + s * stage_bytes + p.x_buf_bytes;
        const int rows = p.R * p.Wp;
        for (int r = rpar; r < rows; r += 2) {
          const uint32_t off = r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
          bsum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(g + off));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    // (2) drain the accumulators
    if (tile_begin < tile_end) {
      const int q = warp & 3;
      if (et == 0) FD_WTS(2);
      mbar_wait(acc_full, 0);
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");      // all column sums done before the staging tiles overwrite the stages
      if (et == 0) FD_WTS(3);
      // TMEM -> registers -> fp32 staging tiles in shared memory (the stage buffers are free now) -> TMA tensor
      // REDUCE-stores (cp.reduce.async.bulk.tensor .add: the fp32 adds happen in L2, 16 KB per instruction)
      // instead of 41k per-thread red.global.add per CTA.  Staging tile = [128 rows][32 fp32] with the 128B
      // swizzle of the tensor map, so the row-per-thread writes are bank-conflict free.
      const int row = q * 32 + lane;           // = tsel * 64 + ci: row of the [2 taps][64 ci] block
      const uint32_t sw = static_cast<uint32_t>(row) & 7u;
#pragma unroll 1
      for (int j = 0; j < 5; ++j) {
        // pairs 0..3 = taps (2j, 2j+1); pair 4 = taps (7, 8) where tap 7 was already produced by pair 3: add zeros
        const bool dup = (j == 4) && (row < 64);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(j * kC + half * 32),
                             acc);
          tmem_ld_wait();
          uint8_t* tile = sStage + static_cast<size_t>(j * 2 + half) * (128 * 128) + row * 128;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            uint4 u = make_uint4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            if (dup) u = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(tile + ((static_cast<uint32_t>(v) ^ sw) << 4)) = u;
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          const int tap0 = (j < 4) ? 2 * j : 7;
          const int grow = p.dw_row0 + prob * p.dw_rows_per_prob + tap0 * kC;
#pragma unroll
          for (int half = 0; half < 2; ++half)
            asm volatile(
                "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&tm_dw)),
                "r"(smem_u32(sStage + static_cast<size_t>(j * 2 + half) * (128 * 128))), "r"(half * 32), "r"(grow)
                : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (et == 0) FD_WTS(4);
      if (p.dbias) {
        sBias[et] = bsum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < 64) atomicAdd(dbias_out + et, sBias[et] + sBias[et + 64]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) FD_WTS(5);
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------- N = 192 formulation
// All nine taps from TWO instructions per 16 pixels.  Moving the column shift of a tap onto the GRADIENT operand,
//   dW[dy][dx][ci][co] = sum_q xpad[q + dy*Wp + dx][ci] g[q][co] = sum_q' xpad[q' + dy*Wp][ci] g[q' - dx][co],
// makes the M side the row shifts (atoms of 64 cin, Wp rows apart) and the N side the column shifts (atoms of 64 cout, ONE
// row apart): D[(dy, ci), (dx, co)] with M = 128 = 2 row shifts and N = 192 = 3 column shifts.  Instruction 1 = dy 0, 1;
// instruction 2 = dy 2 (its second atom is a fourth row shift: computed, never drained).  Per 16 pixels: 2 x 96 clk of tensor
// work for 9 taps (75 % useful) and 10 KB of operand reads per instruction (107 B/clk: below the shared-memory bound), where
// the tap-pair form issues five M=128, N=64 instructions of ~73 clk (6 KB each: operand bound, 10 tap slots for 9 taps).
// The gradient tile is loaded two pixels further right (TMA box starts at x = -2: zero filled), so g[q' - dx] is the tile read
// (2 - dx) rows from its start; the K range grows by two pixels.  Accumulators: 2 x 192 TMEM columns across all tiles of the CTA.
__global__ void __launch_bounds__(kThreads, 1)
wgrad3x3_n192_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ GMaps gmaps,
                     const __grid_constant__ CUtensorMap tm_dw, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const uint32_t stage_bytes = p.x_buf_bytes + p.g_buf_bytes;
  uint8_t* sStage = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* full = bars + 0;      // [2]
  uint64_t* empty = bars + 2;     // [2]
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sBias = reinterpret_cast<float*>(bars + 6);  // [128]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // K-padding rows of g and the over-read tail of x are never written by TMA and must stay zero.
  for (uint32_t i = threadIdx.x * 16u; i < 2 * stage_bytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sStage + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&gmaps.m[0]);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1 + 4);  // MMA commit + the four column-sum warps
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int prob = blockIdx.x / p.ctas_per_prob;
  const int chunk = blockIdx.x - prob * p.ctas_per_prob;
  const int tile_begin = prob * p.tiles_per_prob + min(p.tiles_per_prob, chunk * p.per);
  const int tile_end = prob * p.tiles_per_prob + min(p.tiles_per_prob, (chunk + 1) * p.per);
  float* const dbias_out = p.dbias ? p.dbias + prob * p.dbias_stride : nullptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / p.tiles_per_img;
        const int rem = tile - n * p.tiles_per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * p.R, w0 = tw * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, p.x_bytes + p.g_bytes);
        tma_load_4d(sStage + s * stage_bytes, &tm_x, full + s, 0, w0 - 1, h0 - 1, n);
        tma_load_4d(sStage + s * stage_bytes + p.x_buf_bytes, &gmaps.m[tw], full + s, 0, -2, h0, n);   // two zero pixels first
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 3 * kC, 1, 1);  // both operands MN-major
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t x_addr = smem_u32(sStage + s * stage_bytes);
        // N atom j = column shift dx = 2 - j: the tile read j rows from its start, atoms 128 B apart (LBO)
        const uint32_t g_lo = sdesc_lo(x_addr + p.x_buf_bytes, 128);
        // M atoms = row shifts dy, dy + 1: Wp rows apart
        const uint32_t a_lo0 = sdesc_lo(x_addr, static_cast<uint32_t>(p.Wp) * 128u);
        const uint32_t a_lo1 = sdesc_lo(x_addr + static_cast<uint32_t>(2 * p.Wp) * 128u, static_cast<uint32_t>(p.Wp) * 128u);
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t bd = sdesc_sw128(g_lo + ks * 128);      // 16 pixel rows = 2048 B per K step
          const uint32_t acc = (it | ks) != 0 ? 1u : 0u;
          umma_bf16(tmem_base, sdesc_sw128(a_lo0 + ks * 128), bd, idesc, acc);
          umma_bf16(tmem_base + 3 * kC, sdesc_sw128(a_lo1 + ks * 128), bd, idesc, acc);
        }
        umma_commit(empty + s);
        if (tile + 1 == tile_end) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // epilogue warps: (1) bias gradient from the smem g tiles while the MMAs run (zero rows add nothing)
    const int et = threadIdx.x - 64;      // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum = 0.f;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      if (p.dbias) {
        const uint8_t* g = sStage + s * stage_bytes + p.x_buf_bytes;
        const int rows = p.R * p.Wp;
        for (int r = rpar; r < rows; r += 2) {
          const uint32_t off = r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
          bsum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(g + off));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    // (2) drain: accumulator a (rows dy = 2a, 2a + 1), column block j (dx = 2 - j) -> tap (dy, dx); 18 staging tiles
    // [64 ci][32 fp32] (tap, column half) with the 128B swizzle of the tensor map, then TMA reduce-stores into [9][ci][co]
    if (tile_begin < tile_end) {
      const int q = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // the staging tiles below overwrite the stage buffers: every warp must have finished its column sums of the LAST tile
      // first (with the faster MMAs a warp can get here while a sibling still reads the gradient tile: NaN bias gradients)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int row = q * 32 + lane;           // = (dy & 1) * 64 + ci
      const int ci = row & 63;
      const uint32_t sw = static_cast<uint32_t>(ci) & 7u;
#pragma unroll 1
      for (int a = 0; a < 2; ++a) {
        const int dy = 2 * a + (row >> 6);
        if (dy < 3) {                          // warp-uniform: a warp is one lane quadrant = one row shift
#pragma unroll 1
          for (int j = 0; j < 3; ++j) {
            const int tap = dy * 3 + (2 - j);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t acc[32];
              tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                     static_cast<uint32_t>(a * 3 * kC + j * kC + half * 32),
                                 acc);
              tmem_ld_wait();
              uint8_t* tile = sStage + static_cast<size_t>(tap * 2 + half) * (64 * 128) + ci * 128;
#pragma unroll
              for (int v = 0; v < 8; ++v)
                *reinterpret_cast<uint4*>(tile + ((static_cast<uint32_t>(v) ^ sw) << 4)) =
                    make_uint4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
            }
          }
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        const int grow = p.dw_row0 + prob * p.dw_rows_per_prob;
        for (int t2 = 0; t2 < 18; ++t2)
          asm volatile(
              "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                  reinterpret_cast<uint64_t>(&tm_dw)),
              "r"(smem_u32(sStage + static_cast<size_t>(t2) * (64 * 128))), "r"((t2 & 1) * 32), "r"(grow + (t2 >> 1) * kC)
              : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      }
      if (p.dbias) {
        sBias[et] = bsum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < 64) atomicAdd(dbias_out + et, sBias[et] + sBias[et + 64]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace fd

extern "C" FD_API int fd_debug_wgrad_timing(unsigned long long* out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fd::g_wgrad_dbg, sizeof(unsigned long long) * n));
}

extern "C" int fd_conv3x3_wgrad_multi(const fd_bf16* x, const fd_bf16* g, int nprob, int B, int H, int W, int C,
                                      float* dw_packed, long dw_stride, float* dbias, long dbias_stride, int flags,
                                      void* stream) {
  using namespace fd;
  if (!x || !g || !dw_packed || nprob <= 0 || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != kC) return FD_EUNSUPPORTED;
  const int nsm = sm_count();
  // column strips of TW <= 62 pixels; K index p = y*Wp + x with Wp = TW + 2 (two junk columns per row)
  const int tiles_w = (W + 61) / 62;
  if (tiles_w > kMaxStrips) return FD_EUNSUPPORTED;
  const int TW = (W + tiles_w - 1) / tiles_w;
  const int Wp = TW + 2;
  const size_t smem_cap = 227 * 1024;
  // N = 192 formulation (two instructions per 16 pixels) unless FD_WGRAD_TAPPAIR=1 asks for the five tap-pair instructions
  static const bool n192 = getenv("FD_WGRAD_TAPPAIR") == nullptr;
  const int g_extra = n192 ? 2 : 0;            // the column-shifted reads of the gradient tile reach two rows further

  // Rows per tile: among the heights whose two stages of (x halo tile + g tile) fit in shared memory, pick the
  // one that minimises the work of the busiest CTA (tiles per CTA x max(MMA time, TMA fill time) per tile).
  // Measured: one M=128,N=64,K=16 MMA with MN-major operands takes ~73 clk; five of them per 16 pixels.
  int bestR = 0;
  double best = 1e30;
  for (int R = 1; R <= H && R + 2 <= 256; ++R) {
    const int ksteps = (R * Wp + 15) / 16;
    const size_t xb = (static_cast<size_t>(ksteps * 16 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024;
    const size_t gb = (static_cast<size_t>(ksteps * 16 + g_extra) * 128 + 1023) / 1024 * 1024;
    if (2 * (xb + gb) + 1024 + 1024 > smem_cap) break;
    const long tiles_per_prob = static_cast<long>(B) * ((H + R - 1) / R) * tiles_w;
    long cpp = nsm / nprob;
    if (cpp < 1) cpp = 1;
    if (cpp > tiles_per_prob) cpp = tiles_per_prob;
    const long per = (tiles_per_prob + cpp - 1) / cpp;
    const double mma = n192 ? ksteps * 2 * 110.0 : ksteps * 5 * 73.0;
    const double fill = static_cast<double>(2 * R + 2) * Wp * 128 / 48.0;
    const double cost = per * ((mma > fill ? mma : fill) + 1500.0);   // + per-tile pipeline bubble (measured)
    if (cost < best) { best = cost; bestR = R; }
  }
  if (bestR == 0) return FD_EUNSUPPORTED;

  WgradParams p;
  p.B = nprob * B; p.H = H; p.W = W; p.R = bestR; p.TW = TW; p.Wp = Wp; p.tiles_w = tiles_w;
  p.tiles_per_img = ((H + bestR - 1) / bestR) * tiles_w;
  p.num_tiles = p.B * p.tiles_per_img;
  p.ksteps = (bestR * Wp + 15) / 16;
  p.x_bytes = static_cast<uint32_t>((bestR + 2) * Wp * 128);
  p.g_bytes = static_cast<uint32_t>(bestR * Wp * 128);
  p.x_buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.ksteps * 16 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024);
  p.g_buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.ksteps * 16 + g_extra) * 128 + 1023) / 1024 * 1024);
  p.dw = dw_packed; p.dbias = dbias; p.flags = flags;
  { const char* d = getenv("FD_WGRAD_TIMING"); p.dbg = d ? atoi(d) : 0; }
  p.nprob = nprob;
  p.tiles_per_prob = B * p.tiles_per_img;
  p.dw_stride = dw_stride; p.dbias_stride = dbias_stride;
  // CTAs per problem: one wave over the SMs in total; few CTAs when there are few tiles so that the
  // per-CTA reduction into global memory (36864 fp32 adds) stays amortised (>= 4 tiles per CTA)
  int cpp = nsm / nprob;                       // one wave over all SMs; the drain is a handful of TMA reduces per CTA
  if (cpp < 1) cpp = 1;
  if (cpp > p.tiles_per_prob) cpp = p.tiles_per_prob;
  p.per = (p.tiles_per_prob + cpp - 1) / cpp;
  p.ctas_per_prob = (p.tiles_per_prob + p.per - 1) / p.per;

  CUtensorMap tm_x;
  GMaps gmaps;
  int rc = make_tmap_nhwc_bf16(&tm_x, x, p.B, H, W, C, Wp, bestR + 2);
  if (rc != FD_OK) return rc;
  for (int tw = 0; tw < kMaxStrips; ++tw) {
    const int w0 = (tw < tiles_w ? tw : 0) * TW;
    const int wext = (W - w0 < TW) ? W - w0 : TW;
    rc = make_tmap_nhwc_bf16_strided(&gmaps.m[tw], g + static_cast<size_t>(w0) * C, p.B, H, wext, W, C, Wp, bestR);
    if (rc != FD_OK) return rc;
  }

  if (nprob > 1 && dw_stride % kC != 0) return FD_EINVAL;
  CUtensorMap tm_dw;
  p.dw_row0 = 0;
  p.dw_rows_per_prob = static_cast<int>(dw_stride / kC);
  rc = make_tmap_2d_f32(&tm_dw, dw_packed, static_cast<long>(nprob - 1) * p.dw_rows_per_prob + 9 * kC, kC, n192 ? 64 : 128, 32);
  if (rc != FD_OK) return rc;
  const size_t stages = 2 * static_cast<size_t>(p.x_buf_bytes + p.g_buf_bytes);
  const size_t drain = n192 ? 18 * 64 * 128     // 9 taps x 2 column halves x [64 rows][128 B]
                            : 10 * 128 * 128;   // 5 tap pairs x 2 column halves x [128 rows][128 B]
  p.bar_off = static_cast<uint32_t>(stages > drain ? stages : drain);
  const size_t smem = p.bar_off + 1024 + 1024;
  if (smem > smem_cap) return FD_EUNSUPPORTED;
  auto kern = n192 ? wgrad3x3_n192_kernel : wgrad3x3_tc_kernel;
  cudaError_t e = set_max_dyn_smem(kern, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = nprob * p.ctas_per_prob;
  e = launch_k(kern, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), tm_x, gmaps, tm_dw, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}

extern "C" int fd_conv3x3_wgrad(const fd_bf16* x, const fd_bf16* g, int B, int H, int W, int C, float* dw_packed,
                                float* dbias, int flags, void* stream) {
  return fd_conv3x3_wgrad_multi(x, g, 1, B, H, W, C, dw_packed, 0, dbias, 0, flags, stream);
}
