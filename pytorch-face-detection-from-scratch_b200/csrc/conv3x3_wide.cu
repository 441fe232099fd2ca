// 3x3 / stride 1 / pad 1 convolution for the WIDE models (filters = 128: the width train_model.py:17 trains; the
// 128- and 256-channel blocks of models/SSD.py:164-189), 64*gin -> 128 output channels per launch, as an implicit GEMM
// on the sm_100a tensor cores with the two SMs of a TPC working as ONE (tcgen05.mma.cta_group::2).
//
// Replaces aten::conv2d (+ leaky_relu / dropout2d / residual add) at models/PoolResnet.py:35-40 and -- with
// dgrad-packed weights -- the input-gradient half of its backward, for channel counts above the 64-channel kernels
// (conv3x3_tc.cu), on the channel-PLANE layout of engine_planar.py (every tensor = planes of [B,H,W,64] bf16).
//
// Why a kernel of its own instead of four 64 -> 64 launches per layer:
//   * N = 128 halves the A-operand shared-memory traffic per FLOP.  A M=128,N=64,K=16 MMA reads 6 KB of operands for
//     32 clk of tensor work (128 B/clk of smem bandwidth => 48 clk: operand bound); here one CTA reads A (4 KB) and
//     HALF of B (2 KB) per 64 clk of tensor work: the pair's B operand (128 couts x 16 cin) is split between the two
//     CTAs' shared memories (each holds 64 couts), so the tensor pipe, not shared memory, is the bound.
//   * The partial sums over input planes accumulate in TMEM (fp32) instead of travelling through HBM as bf16
//     between chained launches, and the activation / mask / dropout / skip epilogue is fused (no fd_act_mask pass).
//   * 288 KB of weights per 128 couts do not fit in shared memory: they STREAM from L2 through a ring of
//     (input plane, tap) chunks -- 8 KB per CTA and chunk -- once per pair of 256-row tiles.
//
// Halo-tile formulation as in conv3x3_tc.cu: one TMA box per input plane lands the zero-padded (R+2) x Wp patch;
// GEMM row m = y*Wp + x; tap (ky,kx) reads the same tile through a descriptor shifted by (ky*Wp + kx)*128 bytes.
// A tile has up to two 128-row blocks; its accumulators are 2 x 128 TMEM columns, double buffered (512 columns).
//
// CTA pair protocol (rank 0 = leader issues every MMA for both CTAs):
//   in_full[b], w_full[s]  : leader's barriers; BOTH CTAs' TMA loads complete_tx on them (cta_group::2 loads)
//   in_empty[b], w_empty[s], acc_full[a] : per-CTA barriers, signalled by MULTICAST tcgen05.commit
//   acc_empty[a]           : leader's barrier, 2 arrivals (each CTA's epilogue; the peer arrives remotely)
// Maps with fewer tiles than SM pairs (15x15 at batch 64: 64 two-block tiles) run in SHARED-TILE mode: the pair works on
// ONE tile, CTA r owns 128-row block r -- its copy of the halo tile is loaded 128 rows (16 KB) lower in shared memory, so the
// pair's common A descriptor addresses block 0 in CTA 0 and block 1 in CTA 1 -- and twice as many SMs are busy.
// Warp roles (672 threads): 0 = input loads, 1 = MMA issuer + TMEM owner, 2..17 = epilogue, 18 = TMA stores,
// 19 = weight stream, 20 = residual loads (each stream of loads waits on its own barriers: none delays another).
// Weights are handed over one KERNEL ROW (3 taps, 24 KB per CTA) at a time: the MMA thread's queue is only ~2 MMAs deep,
// so every barrier test (~100-250 clk even when the data is long there) drains the tensor pipe; 6 tests per tile
// instead of 18 (measured: 4400 -> ~1500 clk of stall per 9200-clk tile).
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;                       // channels per plane
constexpr int kNOut = 128;                   // TMEM / constant-table stride of one block: the widest instantiation (kN = 128)
constexpr int kInBufs = 3;                   // ring of input-plane tiles
constexpr int kMaxWSlots = 8;                // ring of weight chunks (kernel rows): as many as fit (WideParams::wslots)
constexpr int kEpiWarpsW = 16;
constexpr int kEpiThreadsW = kEpiWarpsW * 32;
constexpr int kStoreWarpW = 2 + kEpiWarpsW;  // 18
constexpr int kWeightWarp = kStoreWarpW + 1; // 19
constexpr int kResWarp = kWeightWarp + 1;    // 20
constexpr int kThreadsW = (kResWarp + 1) * 32;      // 672
constexpr int kMaxGin = 4;

struct WideMaps {
  CUtensorMap in[kMaxGin];
  CUtensorMap res[2];
  CUtensorMap out[2];
  CUtensorMap w;
};

struct WideParams {
  int B, H, W, R, TW, Wp, nblk, tiles_w, tiles_h, num_tiles, gin;
  int tap_lo, tap_hi;           // taps issued: [0,9) for 3x3, [4,5) for the centre-tap (1x1) mode
  uint32_t in_bytes;            // bytes of one input-plane TMA box
  uint32_t in_buf_bytes;        // bytes reserved per input ring buffer
  uint32_t stg_bytes, stg_buf_bytes;   // one staging UNIT: a whole plane-tile, or (split) the rows of one 128-row block
  int wslots;                   // weight ring depth
  int resident;                 // every weight chunk of a tile has its own slot: loaded ONCE per CTA, no ring hand-shake
  int share;                    // SHARED-TILE mode (fewer tiles than SM pairs): the pair works on ONE two-block tile, CTA r
                                // owns block r; outputs / residual through plain vector loads / stores, no staging
  const __nv_bfloat16* res_ptr[2];
  __nv_bfloat16* out_ptr[2];
  __nv_bfloat16* out2_ptr[2];   // shared-tile mode only: BOTH outputs (out = the sum, out2 = its masked / scaled copy)
  int tpg, ngrp;                // taps per weight chunk (3: a kernel row; 1: centre-tap mode) and chunks per input plane
  int split, rpb, units;        // split: 128 % Wp == 0, a block = rpb whole tile rows = one staging unit; units per plane-tile
  uint32_t inv_wp;
  int flags, has_res, staged_out2, dbg;
  float slope;
  const float* bias;            // [128]
  const float* chan_scale[2];   // per output plane [B,64] or null
  const float* chan_scale2[2];
  const uint16_t* mask_in[2];   // per output plane, [pixel][4] 16-channel units
  uint16_t* mask_out[2];
};

// FD_WIDE_TIMING=1: cycles the MMA thread of CTA 0 spent waiting for [0] free accumulators, [1] input planes, [2] weight
// chunks, [3] total loop cycles, [4] tiles; epilogue thread 0: [5] waiting for acc_full, [6] waiting for staging, [7] total
__device__ unsigned long long g_wide_dbg[16];   // [8] kernel entry -> MMA loop start, [9] MMA loop end -> CTA exit, [10] entry -> after pdl_wait, [11] launches
#define FD_WTE(slot, expr) do { if (p.dbg && blockIdx.x == 0 && et == 0) { const long long t0_ = clock64(); expr; g_wide_dbg[slot] += clock64() - t0_; } else { expr; } } while (0)
#define FD_WT(slot, expr) do { if (p.dbg && blockIdx.x == 0) { const long long t0_ = clock64(); expr; g_wide_dbg[slot] += clock64() - t0_; } else { expr; } } while (0)

__device__ __forceinline__ void bar_sync_epi_w() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreadsW) : "memory"); }

// ---- cta_group-dependent primitives (kCg = 1: one CTA on its own; kCg = 2: the CTA pair) ----
template <int kCg>
__device__ __forceinline__ void tmem_alloc_g(uint32_t* smem_dst, uint32_t ncols) {
  if (kCg == 2)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
                 : "memory");
}
template <int kCg>
__device__ __forceinline__ void tmem_relinquish_g() {
  if (kCg == 2) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCg>
__device__ __forceinline__ void tmem_dealloc_g(uint32_t taddr, uint32_t ncols) {
  if (kCg == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
template <int kCg>
__device__ __forceinline__ void umma_g(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (kCg == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    umma_bf16(tmem_d, adesc, bdesc, idesc, accumulate);
}
// arrive on the barrier at this shared-memory offset in EVERY CTA of the group once the MMAs issued so far are done
template <int kCg>
__device__ __forceinline__ void umma_commit_g(uint64_t* bar) {
  if (kCg == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  else
    umma_commit(bar);
}
// the leader CTA's copy of a barrier (shared::cluster address of the same offset in CTA rank 0)
template <int kCg>
__device__ __forceinline__ uint32_t leader_bar(uint64_t* bar) {
  return kCg == 2 ? mapa_shared(smem_u32(bar), 0) : smem_u32(bar);
}
template <int kCg>
__device__ __forceinline__ void tma_load_2d_g(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1) {
  if (kCg == 2)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}
template <int kCg>
__device__ __forceinline__ void tma_load_4d_g(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1, int c2,
                                              int c3) {
  if (kCg == 2)
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// kN = output channels per launch: 128 (two planes) or 64 (one plane: the 64-channel layers on CTA pairs)
template <int kCg, int kN>
__global__ void __launch_bounds__(kThreadsW, 1)
conv3x3_wide_kernel(const __grid_constant__ WideMaps maps, const __grid_constant__ WideParams p) {
  extern __shared__ uint8_t smem_raw[];
  const long long t_entry = clock64();
  __shared__ long long s_mma_end, s_after_wait;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  // Layout: input ring | weight ring | staging 0 | staging 1 | constants | barriers.  Junk GEMM rows of the last block read a
  // few rows BEHIND their input buffer (into the next ring buffer / the weight ring): harmless, they are never stored.
  constexpr int kNP = kN / 64;                                              // output planes
  constexpr uint32_t kTapBytes = static_cast<uint32_t>(kCg == 2 ? kN / 2 : kN) * 128u;   // this CTA's part of one (plane, tap)
  const uint32_t kChunkBytes = kTapBytes * static_cast<uint32_t>(p.tpg);  // one ring slot = one kernel row of taps
  uint8_t* sIn = smem + (p.share ? 16384u : 0u);      // shared-tile mode: CTA 1 loads its tile 16 KB lower (guard space)
  uint8_t* sW = sIn + ((kInBufs * p.in_buf_bytes + 1023u) & ~1023u);
  uint8_t* sStg = sW + p.wslots * kChunkBytes;
  float* sConst = reinterpret_cast<float*>(sStg + 2 * p.stg_buf_bytes);     // bias[128] | chan_scale[128] | chan_scale2[128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sConst + 3 * kNOut);
  uint64_t* in_full = bars;                       // [3]  (leader's copy is the live one)
  uint64_t* in_empty = bars + 3;                  // [3]
  uint64_t* w_full = bars + 6;                    // [8]  (leader)
  uint64_t* w_empty = bars + 14;                  // [8]
  uint64_t* acc_full = bars + 22;                 // [2]
  uint64_t* acc_empty = bars + 24;                // [2]  (leader, 2 arrivals)
  uint64_t* res_full = bars + 26;                 // [2]
  uint64_t* stg_free = bars + 28;                 // [2]
  uint64_t* stg_ready = bars + 30;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kCg == 2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.gin; ++g) tma_prefetch_desc(&maps.in[g]);
    tma_prefetch_desc(&maps.w);
    tma_prefetch_desc(&maps.out[0]);
    for (int i = 0; i < kInBufs; ++i) {
      mbar_init(in_full + i, 1);
      mbar_init(in_empty + i, 1);
    }
    for (int i = 0; i < kMaxWSlots; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_empty + i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, kCg);
      mbar_init(res_full + s, 1);
      mbar_init(stg_free + s, 1);
      mbar_init(stg_ready + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_g<kCg>(tmem_slot, 512);
    tmem_relinquish_g<kCg>();
  }
  tc_fence_before();
  __syncthreads();
  if (kCg == 2) cluster_sync_all();           // the peer's barriers exist before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  // work items: a pair-tile j covers tiles kCg*j + rank; this group handles j = group, group + ngroups, ...
  const int ngroups = static_cast<int>(gridDim.x) / kCg;
  const int group = static_cast<int>(blockIdx.x) / kCg;
  const int njobs = p.share ? p.num_tiles : (p.num_tiles + kCg - 1) / kCg;
  pdl_trigger();
  pdl_wait();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) { g_wide_dbg[10] += clock64() - t_entry; g_wide_dbg[11] += 1; }
  if (p.dbg && threadIdx.x == 0) s_after_wait = clock64();

  auto tile_coords = [&](int tile, int& n, int& h0, int& w0) {
    n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    const int th = rem / p.tiles_w;
    h0 = th * p.R;
    w0 = (rem - th * p.tiles_w) * p.TW;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ input loads
    if (elect_one_sync()) {
      uint32_t b = 0, ph = 0;              // ring buffer and its phase
      for (int j = group; j < njobs; j += ngroups) {
        const int tile = p.share ? j : j * kCg + static_cast<int>(rank);   // >= num_tiles: dummy (all coordinates out of bounds -> zeros)
        int n, h0, w0;
        tile_coords(tile, n, h0, w0);
        for (int kh = 0; kh < p.gin; ++kh) {
          mbar_wait_sleep(in_empty + b, ph ^ 1u);
          if (leader) mbar_expect_tx(in_full + b, p.in_bytes * kCg);
          tma_load_4d_g<kCg>(sIn + b * p.in_buf_bytes - (p.share ? rank * 16384u : 0u), &maps.in[kh], leader_bar<kCg>(in_full + b),
                             0, w0 - 1, h0 - 1, n);
          if (++b == kInBufs) { b = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == kResWarp) {
    // ------------------------------------------------------------------ residual loads into the staging units
    if (p.has_res && !p.share && elect_one_sync()) {
      uint32_t q = 0;
      for (int j = group; j < njobs; j += ngroups) {
        const int tile = p.share ? j : j * kCg + static_cast<int>(rank);
        int n, h0, w0;
        tile_coords(tile, n, h0, w0);
        for (int g = 0; g < kNP; ++g) {
          for (int u = 0; u < p.units; ++u, ++q) {
            const uint32_t sb = q & 1u, ph = (q >> 1) & 1u;
            mbar_wait_sleep(stg_free + sb, ph ^ 1u);        // the store two units ago has drained this buffer
            mbar_expect_tx(res_full + sb, p.stg_bytes);
            tma_load_4d(sStg + sb * p.stg_buf_bytes, &maps.res[g], res_full + sb, 0, w0, h0 + u * p.rpb, n);
          }
        }
      }
    }
  } else if (warp == kWeightWarp) {
    // ------------------------------------------------------------------ weight stream: chunk (kh, tap) = [128 cout][64 cin]
    if (elect_one_sync()) {
      uint32_t s = 0, ph = 0;              // ring slot and its phase
      for (int j = group; j < njobs; j += ngroups) {
        if (p.resident && j != group) break;          // resident weights: the first tile's loads serve every tile
        for (int kh = 0; kh < p.gin; ++kh) {
          for (int r = 0; r < p.ngrp; ++r) {
            if (!p.resident) mbar_wait_sleep(w_empty + s, ph ^ 1u);
            if (leader) mbar_expect_tx(w_full + s, kChunkBytes * kCg);
            for (int i = 0; i < p.tpg; ++i)
              tma_load_2d_g<kCg>(sW + s * kChunkBytes + i * kTapBytes, &maps.w, leader_bar<kCg>(w_full + s), 0,
                                 (kh * 9 + p.tap_lo + r * p.tpg + i) * kN + static_cast<int>(rank) * (kN / 2));
            if (++s == static_cast<uint32_t>(p.wslots)) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && elect_one_sync()) {
      const long long t_start = clock64();
      constexpr uint32_t idesc = make_idesc_bf16(128 * kCg, kN, 0, 0);
      const uint32_t wp_units = static_cast<uint32_t>(p.Wp) * 8u;
      const int mma_blocks = p.share ? 1 : p.nblk;     // shared-tile mode: one instruction covers block 0 (CTA 0) and block 1 (CTA 1)
      uint32_t ib = 0, iph = 0;            // input ring buffer and its phase
      uint32_t s = 0, wph = 0;             // weight ring slot and its phase
      int it = 0;
      for (int j = group; j < njobs; j += ngroups, ++it) {
        const uint32_t a = static_cast<uint32_t>(it) & 1u, aph = (static_cast<uint32_t>(it) >> 1) & 1u;
        FD_WT(0, mbar_wait_cluster(acc_empty + a, aph ^ 1u));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * 256u;
        uint32_t accum = 0;
        for (int kh = 0; kh < p.gin; ++kh) {
          FD_WT(1, mbar_wait(in_full + ib, iph));
          tc_fence_after();
          // first tap: (0,0), or the centre tap (1,1) in 1x1 mode
          uint32_t a_lo = sdesc_lo(smem_u32(sIn + ib * p.in_buf_bytes), 16) + (p.tap_lo != 0 ? wp_units + 8u : 0u);
#pragma unroll 1
          for (int r = 0; r < p.ngrp; ++r) {
            if (!p.resident || it == 0) {
              FD_WT(2, mbar_wait(w_full + s, wph));
              tc_fence_after();
            }
            uint32_t b_lo = sdesc_lo(smem_u32(sW + s * kChunkBytes), 16);
#pragma unroll 1
            for (int i = 0; i < p.tpg; ++i) {
#pragma unroll 1
              for (int mb = 0; mb < mma_blocks; ++mb) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_g<kCg>(d_tmem + static_cast<uint32_t>(mb) * kNOut, sdesc_sw128(a_lo + static_cast<uint32_t>(mb) * 1024u + 2 * k),
                              sdesc_sw128(b_lo + 2 * k), idesc, (k != 0) ? 1u : accum);
              }
              accum = 1;                             // the tile's first tap overwrites the accumulators of EVERY block
              a_lo += 8u;                            // next column
              b_lo += kTapBytes >> 4;
            }
            if (!p.resident) umma_commit_g<kCg>(w_empty + s);     // weight slot free (in both CTAs) once these MMAs have read it
            if (++s == static_cast<uint32_t>(p.wslots)) { s = 0; wph ^= 1u; }
            a_lo += wp_units - 24u;                  // next kernel row
          }
          umma_commit_g<kCg>(in_empty + ib);         // input plane buffer free
          if (++ib == kInBufs) { ib = 0; iph ^= 1u; }
        }
        umma_commit_g<kCg>(acc_full + a);            // the tile's accumulators are final
      }
      if (p.dbg && blockIdx.x == 0) { g_wide_dbg[3] += clock64() - t_start; g_wide_dbg[4] += it; g_wide_dbg[8] += t_start - t_entry; s_mma_end = clock64(); g_wide_dbg[6] += s_mma_end - t_entry; }
    }
    __syncwarp();
  } else if (warp == kStoreWarpW) {
    // ------------------------------------------------------------------ TMA store issuer
    if (!p.share && elect_one_sync()) {
      int it = 0;
      for (int j = group; j < njobs; j += ngroups, ++it) {
        const int tile = p.share ? j : j * kCg + static_cast<int>(rank);
        int n, h0, w0;
        tile_coords(tile, n, h0, w0);
        uint32_t q = static_cast<uint32_t>(it) * static_cast<uint32_t>(kNP) * static_cast<uint32_t>(p.units);
        for (int g = 0; g < kNP; ++g) {
          for (int u = 0; u < p.units; ++u, ++q) {
            const uint32_t sb = q & 1u, ph = (q >> 1) & 1u;
            mbar_wait_sleep(stg_ready + sb, ph);
            if (tile < p.num_tiles && h0 + u * p.rpb < p.H) {
              tma_store_4d(&maps.out[g], sStg + sb * p.stg_buf_bytes, 0, w0, h0 + u * p.rpb, n);    // beyond the image: clipped
              tma_store_commit();
              tma_store_wait_read<0>();
            }
            mbar_arrive(stg_free + sb);
          }
        }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (16)
    // One thread = one GEMM row (pixel) x 16 channels of one output plane.  TMEM lane quadrant = warp % 4,
    // channel quarter = (warp - 2) / 4; the two output planes (TMEM column halves) are processed one after the other.
    const int q4 = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int c0 = cq * 16;
    const int et = threadIdx.x - 64;
    const uint64_t slope2 = pk2(p.slope, p.slope);
    const bool lrelu = (p.flags & FD_EPI_LRELU) != 0;
    const bool has_cs = p.chan_scale[0] != nullptr, has_cs2 = p.chan_scale2[0] != nullptr;
    if (et < kN) sConst[et] = p.bias ? __ldg(p.bias + et) : 0.f;
    const long long t_epi0 = clock64();
    int it = 0, last_n = -1;
    for (int j = group; j < njobs; j += ngroups, ++it) {
      const int tile = p.share ? j : j * kCg + static_cast<int>(rank);
      const bool live = tile < p.num_tiles;
      int n, h0, w0;
      tile_coords(tile, n, h0, w0);
      if (n != last_n) {          // per-image Dropout2d multipliers (uniform branch)
        bar_sync_epi_w();
        if (live && et >= kNOut && et < kNOut + kN) {
          const int ch = et - kNOut;
          sConst[et] = has_cs ? __ldg(p.chan_scale[ch >> 6] + n * kC + (ch & 63)) : 1.f;
        } else if (live && et >= 2 * kNOut && et < 2 * kNOut + kN) {
          const int ch = et - 2 * kNOut;
          sConst[et] = has_cs2 ? __ldg(p.chan_scale2[ch >> 6] + n * kC + (ch & 63)) : 1.f;
        }
        bar_sync_epi_w();
        last_n = n;
      }
      const uint32_t a = static_cast<uint32_t>(it) & 1u, aph = (static_cast<uint32_t>(it) >> 1) & 1u;
      if (p.share) {
        // shared-tile mode: this CTA owns GEMM rows [128 * rank, 128 * rank + 128) of the tile (TMEM block 0); residual and
        // output go through 16-byte vector loads / stores (a 15x15 map is 29 KB per plane: no staging, no TMA round trip)
        const int m = static_cast<int>(rank) * 128 + q4 * 32 + lane;
        const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
        const int x = m - y * p.Wp;
        const int oy = h0 + y, ox = w0 + x;
        const bool valid = live && (y < p.R) && (x < p.TW) && (oy < p.H) && (ox < p.W);
        const size_t pix = (static_cast<size_t>(n) * p.H + oy) * p.W + ox;
        // residual / mask operands of both planes are fetched BEFORE the accumulators are awaited: their global-memory
        // latency hides behind the tile's MMAs (with one tile per CTA nothing else would)
        uint4 rr[2][2];
        uint32_t mb2[2] = {0xffffu, 0xffffu};
#pragma unroll
        for (int g = 0; g < kNP; ++g) {
          rr[g][0] = rr[g][1] = make_uint4(0, 0, 0, 0);
          if (valid && p.has_res) {
            const uint4* rp = reinterpret_cast<const uint4*>(p.res_ptr[g] + pix * kC + c0);
            rr[g][0] = __ldg(rp);
            rr[g][1] = __ldg(rp + 1);
          }
          if (valid && p.mask_in[g]) mb2[g] = __ldg(p.mask_in[g] + pix * 4 + cq);
        }
        FD_WTE(5, mbar_wait_sleep(acc_full + a, aph, 1000));
        tc_fence_after();
        if (p.dbg && blockIdx.x == 0 && et == 0) g_wide_dbg[14] += clock64() - t_entry;     // entry -> accumulators seen by the epilogue
#pragma unroll
        for (int g = 0; g < kNP; ++g) {
          uint32_t acc[16];
          tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * 256u + static_cast<uint32_t>(g * kC + c0), acc);
          tmem_ld_wait();
          if (valid) {
            uint64_t v2[8];
            epi_bias_act16(acc, sConst + g * kC + c0, sConst + kNOut + g * kC + c0, lrelu, has_cs, slope2, v2);
            if (p.mask_out[g]) p.mask_out[g][pix * 4 + cq] = static_cast<uint16_t>(epi_sign_bits16(v2));
            if (p.has_res) epi_add_bf16x16(v2, rr[g][0], rr[g][1]);
            uint4 u0, u1;
            if (!p.staged_out2 || p.out2_ptr[g]) {
              epi_pack16(v2, u0, u1);
              uint4* op = reinterpret_cast<uint4*>(p.out_ptr[g] + pix * kC + c0);
              op[0] = u0;
              op[1] = u1;
            }
            if (p.staged_out2 || p.out2_ptr[g]) {
              uint64_t o2[8];
              epi_masked16(v2, mb2[g], p.slope, sConst + 2 * kNOut + g * kC + c0, has_cs2, o2);
              epi_pack16(o2, u0, u1);
              uint4* op = reinterpret_cast<uint4*>((p.out2_ptr[g] ? p.out2_ptr[g] : p.out_ptr[g]) + pix * kC + c0);
              op[0] = u0;
              op[1] = u1;
            }
          }
        }
        tc_fence_before();
        bar_sync_epi_w();
        if (p.dbg && blockIdx.x == 0 && et == 0) g_wide_dbg[15] += clock64() - t_entry;     // entry -> epilogue of the tile done
        if (et == 0) {             // every column of this tile's accumulators has been read by every warp
          if (kCg == 2 && !leader) mbar_arrive_remote(mapa_shared(smem_u32(acc_empty + a), 0));
          else mbar_arrive(acc_empty + a);
        }
        continue;
      }
      uint32_t qq = static_cast<uint32_t>(it) * static_cast<uint32_t>(kNP) * static_cast<uint32_t>(p.units);
      for (int g = 0; g < kNP; ++g) {
        const uint16_t* mask_in = p.mask_in[g];
        uint16_t* mask_out = p.mask_out[g];
        uint8_t* stg = nullptr;
        uint32_t sb = 0;
#pragma unroll 1
        for (int mb = 0; mb < p.nblk; ++mb) {
          if (p.split || mb == 0) {          // a staging unit begins: its residual has landed / the store two units ago has drained it
            sb = qq & 1u;
            const uint32_t sph = (qq >> 1) & 1u;
            stg = sStg + sb * p.stg_buf_bytes;
            if (p.has_res) FD_WTE(6, mbar_wait_sleep(res_full + sb, sph));
            else FD_WTE(6, mbar_wait_sleep(stg_free + sb, sph ^ 1u));
          }
          if (g == 0 && mb == 0) {
            FD_WTE(5, mbar_wait_sleep(acc_full + a, aph, 1000));
            tc_fence_after();
          }
          const int m = mb * 128 + q4 * 32 + lane;
          const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
          const int x = m - y * p.Wp;
          const int oy = h0 + y, ox = w0 + x;
          const bool valid = live && (y < p.R) && (x < p.TW) && (oy < p.H) && (ox < p.W);
          const size_t pix = (static_cast<size_t>(n) * p.H + oy) * p.W + ox;
          uint32_t mbits = 0xffffu;
          if (valid && mask_in) mbits = __ldg(mask_in + pix * 4 + cq);
          const uint32_t d = static_cast<uint32_t>((p.split ? y - mb * p.rpb : y) * p.TW + x);    // row of the dense staging unit
          uint8_t* row = stg + d * 128u;
          const uint32_t ch0 = ((static_cast<uint32_t>(cq) * 2u) ^ (d & 7u)) << 4;
          const uint32_t ch1 = ((static_cast<uint32_t>(cq) * 2u + 1u) ^ (d & 7u)) << 4;
          uint32_t acc[16];
          tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * 256u +
                                 static_cast<uint32_t>(mb * kNOut + g * kC + c0),
                             acc);
          tmem_ld_wait();
          if (valid) {
            uint64_t v2[8];
            epi_bias_act16(acc, sConst + g * kC + c0, sConst + kNOut + g * kC + c0, lrelu, has_cs, slope2, v2);
            if (mask_out) mask_out[pix * 4 + cq] = static_cast<uint16_t>(epi_sign_bits16(v2));
            if (p.has_res)
              epi_add_bf16x16(v2, *reinterpret_cast<const uint4*>(row + ch0), *reinterpret_cast<const uint4*>(row + ch1));
            uint4 u0, u1;
            if (!p.staged_out2) {
              epi_pack16(v2, u0, u1);
            } else {
              uint64_t o2[8];
              epi_masked16(v2, mbits, p.slope, sConst + 2 * kNOut + g * kC + c0, has_cs2, o2);
              epi_pack16(o2, u0, u1);
            }
            *reinterpret_cast<uint4*>(row + ch0) = u0;
            *reinterpret_cast<uint4*>(row + ch1) = u1;
          }
          if (p.split || mb == p.nblk - 1) {     // the unit is complete
            const bool tile_done = g == kNP - 1 && mb == p.nblk - 1;
            fence_proxy_async();       // staging writes (generic proxy) -> visible to the TMA store
            if (tile_done) tc_fence_before();
            bar_sync_epi_w();
            if (et == 0) {
              if (tile_done) {         // every column of this tile's accumulators has been read by every warp
                if (kCg == 2 && !leader) mbar_arrive_remote(mapa_shared(smem_u32(acc_empty + a), 0));
                else mbar_arrive(acc_empty + a);
              }
              mbar_arrive(stg_ready + sb);
            }
            ++qq;
          }
        }
      }
    }
    if (p.dbg && blockIdx.x == 0 && et == 0) g_wide_dbg[7] += clock64() - t_epi0;
  }

  tc_fence_before();
  __syncthreads();
  if (p.dbg && blockIdx.x == 0 && threadIdx.x == 0) g_wide_dbg[9] += clock64() - s_mma_end;
  if (p.dbg && threadIdx.x == 0) {       // slowest CTA of the launch: after pdl_wait -> exit, and entry -> exit
    atomicMax(&g_wide_dbg[12], static_cast<unsigned long long>(clock64() - s_after_wait));
    atomicMax(&g_wide_dbg[13], static_cast<unsigned long long>(clock64() - t_entry));
  }
  if (kCg == 2) cluster_sync_all();           // nobody exits (or frees TMEM) while the pair's MMAs / commits may still touch it
  if (warp == 1) tmem_dealloc_g<kCg>(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// A RUN of 3x3 convolutions in one launch (fd_conv3x3_wide_chain).  On maps small enough that one CTA pair holds a whole
// image (15x15 at filters = 128: 16 of the 20 convolutions of a PoolResnet pass, each only ~2.5 us of tensor work) a
// launch per layer spends more time leaving and entering the GPU than computing: exit + cluster tear-down, the
// grid-wide dependency, the next launch's set-up, and only then the first TMA round trip (measured 17.5 k clk per layer for
// 5.7 k clk of MMAs).  Here the pair keeps its image: after the epilogue of layer l has written its planes (plain stores,
// L2-resident) the two CTAs meet on a cluster-scope mbarrier and load them back as layer l + 1's halo tiles; barriers,
// TMEM and the weight ring stay alive, and the weight stream runs ahead through the ring while the epilogue works.
// No grid-wide synchronisation: images are independent.  Same MMAs and the same epilogue arithmetic as the one-layer kernel
// in shared-tile mode (bit-identical outputs, tests/test_gpu_wide.py).
// Inputs of every layer are slabs of two STACKED plane buffers [n_stack][B][H][W][64] (engine_planar.py keeps a run's
// activations / gradients that way for the multi-problem weight-gradient kernel); everything else is a per-layer pointer.
constexpr int kMaxChain = 16;
constexpr int kChainInBufs = 2;              // one layer's two input planes: the next layer's cannot start before they are free
struct ChainLayer {
  int in_n0;            // first image of this layer's input slab in the stacked planes (slab index * B)
  int w_row0;           // first row of its packed weights in the weight map
  int lrelu, has_res;
  const float* bias;    // [128] or null
  const __nv_bfloat16* res_ptr[2];
  __nv_bfloat16* out_ptr[2];        // nullable: conv (+ bias, activation, dropout, skip)
  __nv_bfloat16* out2_ptr[2];       // nullable: the masked / scaled copy (input-gradient chains)
  const float* chan_scale[2];
  const float* chan_scale2[2];
  const uint16_t* mask_in[2];
  uint16_t* mask_out[2];
};
struct ChainParams {
  int B, H, W, Wp, nlayers, wslots;
  uint32_t in_bytes, in_buf_bytes, inv_wp;
  float slope;
  ChainLayer layer[kMaxChain];
};
struct ChainMaps {
  CUtensorMap in[2];
  CUtensorMap w;
};

__global__ void __launch_bounds__(kThreadsW, 1)
conv3x3_wide_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int kCg = 2, kN = 128, kNP = 2;
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr uint32_t kTapBytes = (kN / 2) * 128u;          // this CTA's 64 couts of one (plane, tap)
  constexpr uint32_t kChunkBytes = kTapBytes * 3u;         // one ring slot = one kernel row of taps
  uint8_t* sIn = smem + 16384u;                            // CTA 1 loads its tile 16 KB lower (guard space)
  uint8_t* sW = sIn + ((kChainInBufs * p.in_buf_bytes + 1023u) & ~1023u);
  float* sConst = reinterpret_cast<float*>(sW + p.wslots * kChunkBytes);    // bias[128] | chan_scale[128] | chan_scale2[128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sConst + 3 * kNOut);
  uint64_t* in_full = bars;                       // [2]  (leader's copy is the live one)
  uint64_t* in_empty = bars + 2;                  // [2]
  uint64_t* w_full = bars + 4;                    // [8]  (leader)
  uint64_t* w_empty = bars + 12;                  // [8]
  uint64_t* acc_full = bars + 20;                 // [2]
  uint64_t* acc_empty = bars + 22;                // [2]  (leader, 2 arrivals)
  uint64_t* layer_done = bars + 24;               // [1]  both CTAs' epilogues have written (and fenced) a layer's outputs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n = static_cast<int>(blockIdx.x) / kCg;         // the pair's image

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.in[0]);
    tma_prefetch_desc(&maps.in[1]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < kChainInBufs; ++i) {
      mbar_init(in_full + i, 1);
      mbar_init(in_empty + i, 1);
    }
    for (int i = 0; i < kMaxWSlots; ++i) {
      mbar_init(w_full + i, 1);
      mbar_init(w_empty + i, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, kCg);
    }
    mbar_init(layer_done, kCg);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_g<kCg>(tmem_slot, 512);
    tmem_relinquish_g<kCg>();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer's barriers exist before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ input loads: layer l's planes, once layer l - 1 is out
    if (elect_one_sync()) {
      uint32_t b = 0, ph = 0;
      for (int l = 0; l < p.nlayers; ++l) {
        if (l > 0) mbar_wait_cluster(layer_done, static_cast<uint32_t>(l - 1) & 1u);
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait_sleep(in_empty + b, ph ^ 1u);
          if (leader) mbar_expect_tx(in_full + b, p.in_bytes * kCg);
          tma_load_4d_g<kCg>(sIn + b * p.in_buf_bytes - rank * 16384u, &maps.in[kh], leader_bar<kCg>(in_full + b), 0, -1, -1,
                             p.layer[l].in_n0 + n);
          if (++b == kChainInBufs) { b = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == kWeightWarp) {
    // ------------------------------------------------------------------ weight stream, running ahead through the ring
    if (elect_one_sync()) {
      uint32_t s = 0, ph = 0;
      for (int l = 0; l < p.nlayers; ++l) {
        const int row0 = p.layer[l].w_row0 + static_cast<int>(rank) * (kN / 2);
        for (int c = 0; c < 6; ++c) {           // chunk = (input plane c / 3, kernel row c % 3): taps 3c .. 3c + 2 of the [2][9] list
          mbar_wait_sleep(w_empty + s, ph ^ 1u);
          if (leader) mbar_expect_tx(w_full + s, kChunkBytes * kCg);
          for (int i = 0; i < 3; ++i)
            tma_load_2d_g<kCg>(sW + s * kChunkBytes + i * kTapBytes, &maps.w, leader_bar<kCg>(w_full + s), 0,
                               row0 + (c * 3 + i) * kN);
          if (++s == static_cast<uint32_t>(p.wslots)) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one thread)
    if (leader && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(128 * kCg, kN, 0, 0);
      const uint32_t wp_units = static_cast<uint32_t>(p.Wp) * 8u;
      uint32_t ib = 0, iph = 0, s = 0, wph = 0;
      for (int l = 0; l < p.nlayers; ++l) {
        const uint32_t a = static_cast<uint32_t>(l) & 1u, aph = (static_cast<uint32_t>(l) >> 1) & 1u;
        mbar_wait_cluster(acc_empty + a, aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * 256u;
        uint32_t accum = 0;
        for (int kh = 0; kh < 2; ++kh) {
          mbar_wait(in_full + ib, iph);
          tc_fence_after();
          uint32_t a_lo = sdesc_lo(smem_u32(sIn + ib * p.in_buf_bytes), 16);
#pragma unroll 1
          for (int r = 0; r < 3; ++r) {
            mbar_wait(w_full + s, wph);
            tc_fence_after();
            uint32_t b_lo = sdesc_lo(smem_u32(sW + s * kChunkBytes), 16);
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_g<kCg>(d_tmem, sdesc_sw128(a_lo + 2 * k), sdesc_sw128(b_lo + 2 * k), idesc, (k != 0) ? 1u : accum);
              accum = 1;
              a_lo += 8u;                            // next column
              b_lo += kTapBytes >> 4;
            }
            umma_commit_g<kCg>(w_empty + s);         // weight slot free (in both CTAs) once these MMAs have read it
            if (++s == static_cast<uint32_t>(p.wslots)) { s = 0; wph ^= 1u; }
            a_lo += wp_units - 24u;                  // next kernel row
          }
          umma_commit_g<kCg>(in_empty + ib);         // input plane buffer free
          if (++ib == kChainInBufs) { ib = 0; iph ^= 1u; }
        }
        umma_commit_g<kCg>(acc_full + a);            // the layer's accumulators are final
      }
    }
    __syncwarp();
  } else if (warp >= 2 && warp < 2 + kEpiWarpsW) {
    // ------------------------------------------------------------------ epilogue warps (16): as the one-layer kernel's
    // shared-tile branch -- this CTA owns GEMM rows [128 rank, 128 rank + 128), one thread = one pixel x 16 channels per plane
    const int q4 = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int c0 = cq * 16;
    const int et = threadIdx.x - 64;
    const uint64_t slope2 = pk2(p.slope, p.slope);
    const int m = static_cast<int>(rank) * 128 + q4 * 32 + lane;
    const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
    const int x = m - y * p.Wp;
    const bool valid = (y < p.H) && (x < p.W);
    const size_t pix = (static_cast<size_t>(n) * p.H + y) * p.W + x;
    for (int l = 0; l < p.nlayers; ++l) {
      const ChainLayer& L = p.layer[l];
      const bool has_cs = L.chan_scale[0] != nullptr, has_cs2 = L.chan_scale2[0] != nullptr;
      bar_sync_epi_w();               // nobody still reads the previous layer's constants
      if (et < kN) {
        sConst[et] = L.bias ? __ldg(L.bias + et) : 0.f;
      } else if (et < 2 * kN) {
        const int ch = et - kN;
        sConst[et] = has_cs ? __ldg(L.chan_scale[ch >> 6] + n * kC + (ch & 63)) : 1.f;
      } else if (et < 3 * kN) {
        const int ch = et - 2 * kN;
        sConst[et] = has_cs2 ? __ldg(L.chan_scale2[ch >> 6] + n * kC + (ch & 63)) : 1.f;
      }
      bar_sync_epi_w();
      // skip / mask operands before the accumulators are awaited.  The skip planes may have been written earlier in THIS
      // launch (by this very thread: same pixel, same channels): plain loads, not the read-only path.
      uint4 rr[2][2];
      uint32_t mb2[2] = {0xffffu, 0xffffu};
#pragma unroll
      for (int g = 0; g < kNP; ++g) {
        rr[g][0] = rr[g][1] = make_uint4(0, 0, 0, 0);
        if (valid && L.has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(L.res_ptr[g] + pix * kC + c0);
          rr[g][0] = rp[0];
          rr[g][1] = rp[1];
        }
        if (valid && L.mask_in[g]) mb2[g] = __ldg(L.mask_in[g] + pix * 4 + cq);
      }
      const uint32_t a = static_cast<uint32_t>(l) & 1u, aph = (static_cast<uint32_t>(l) >> 1) & 1u;
      mbar_wait_sleep(acc_full + a, aph, 1000);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < kNP; ++g) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + a * 256u + static_cast<uint32_t>(g * kC + c0), acc);
        tmem_ld_wait();
        if (valid) {
          uint64_t v2[8];
          epi_bias_act16(acc, sConst + g * kC + c0, sConst + kNOut + g * kC + c0, L.lrelu != 0, has_cs, slope2, v2);
          if (L.mask_out[g]) L.mask_out[g][pix * 4 + cq] = static_cast<uint16_t>(epi_sign_bits16(v2));
          if (L.has_res) epi_add_bf16x16(v2, rr[g][0], rr[g][1]);
          uint4 u0, u1;
          if (L.out_ptr[g]) {
            epi_pack16(v2, u0, u1);
            uint4* op = reinterpret_cast<uint4*>(L.out_ptr[g] + pix * kC + c0);
            op[0] = u0;
            op[1] = u1;
          }
          if (L.out2_ptr[g]) {
            uint64_t o2[8];
            epi_masked16(v2, mb2[g], p.slope, sConst + 2 * kNOut + g * kC + c0, has_cs2, o2);
            epi_pack16(o2, u0, u1);
            uint4* op = reinterpret_cast<uint4*>(L.out2_ptr[g] + pix * kC + c0);
            op[0] = u0;
            op[1] = u1;
          }
        }
      }
      // this thread's stores: ordered at GPU scope and against the async proxy (the next layer's TMA loads, either CTA's)
      __threadfence();
      fence_proxy_async_all();
      tc_fence_before();
      bar_sync_epi_w();
      if (et == 0) {
        if (!leader) mbar_arrive_remote(mapa_shared(smem_u32(acc_empty + a), 0));
        else mbar_arrive(acc_empty + a);
        if (l + 1 < p.nlayers) {
          mbar_arrive_remote(mapa_shared(smem_u32(layer_done), rank ^ 1u));
          mbar_arrive_remote(mapa_shared(smem_u32(layer_done), rank));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // nobody exits (or frees TMEM) while the pair's MMAs / commits / arrives may still touch it
  if (warp == 1) tmem_dealloc_g<kCg>(tmem_base, 512);
}

inline size_t wide_in_buf_bytes(int R, int Wp) { return static_cast<size_t>((R + 2) * Wp) * 128; }
inline bool wide_split(int R, int Wp) { return 128 % Wp == 0 && (R * Wp) % 128 == 0; }
inline int wide_unit_rows(int R, int Wp) { return wide_split(R, Wp) ? 128 / Wp : R; }
inline size_t wide_stg_buf_bytes(int R, int Wp, int TW) {
  return (static_cast<size_t>(wide_unit_rows(R, Wp)) * TW * 128 + 1023) / 1024 * 1024;
}
constexpr int kMinWSlots = 2;
// shared memory without the weight ring
inline size_t wide_smem_fixed(int R, int Wp, int TW) {
  const size_t in = (kInBufs * wide_in_buf_bytes(R, Wp) + 1023) / 1024 * 1024;
  return in + 2 * wide_stg_buf_bytes(R, Wp, TW) + 3 * kNOut * 4 + 512 + 1024;
}
inline size_t wide_chunk_bytes(int cg, int tpg, int nout = 128) { return static_cast<size_t>(cg == 2 ? nout / 2 : nout) * 128 * tpg; }
inline int wide_slots(int cg, int tpg, int R, int Wp, int TW, size_t cap, int nout = 128) {
  const size_t fixed = wide_smem_fixed(R, Wp, TW);
  if (fixed + kMinWSlots * wide_chunk_bytes(cg, tpg, nout) > cap) return 0;
  const size_t n = (cap - fixed) / wide_chunk_bytes(cg, tpg, nout);
  return n > kMaxWSlots ? kMaxWSlots : static_cast<int>(n);
}

template <int kCg, int kN>
int launch_wide(const WideMaps& maps, const WideParams& p, size_t smem, cudaStream_t st) {
  cudaError_t e = set_max_dyn_smem(conv3x3_wide_kernel<kCg, kN>, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int nsm = sm_count();
  const int njobs = p.share ? p.num_tiles : (p.num_tiles + kCg - 1) / kCg;
  const int max_groups = nsm / kCg;
  const int groups = njobs < max_groups ? njobs : max_groups;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * kCg);
  cfg.blockDim = dim3(kThreadsW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (kCg == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, conv3x3_wide_kernel<kCg, kN>, maps, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}

// [Cout, Cin, k, k] fp32 (torch; k = 3, or k = 1: only the centre tap is written) -> the two streamed bf16 packings.
// One thread per destination element (coalesced writes, gathered reads: the weights are L2 resident).
//   forward : wf[(co / 128)][ci / 64][tap][co % 128][ci % 64]
//   dgrad   : wd[(ci / 128)][co / 64][8 - tap][ci % 128][co % 64]        (flipped taps, roles of ci / co swapped)
__global__ void __launch_bounds__(256)
pack_conv_wide_kernel(const float* __restrict__ w, int n_layers, int Cout, int Cin, int ksize, __nv_bfloat16* __restrict__ wf,
                      __nv_bfloat16* __restrict__ wd) {
  pdl_trigger();
  pdl_wait();
  const int kk = ksize * ksize;
  const long src_layer = static_cast<long>(Cout) * Cin * kk;
  const long dst_layer = static_cast<long>(Cout) * Cin * 9;
  const long total = static_cast<long>(Cout) * Cin * kk * n_layers;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * 256L) {
    const long l = i / src_layer;
    const long r = i - l * src_layer;
    if (wf) {
      const int ci_lo = static_cast<int>(r % 64);
      const int co_lo = static_cast<int>((r / 64) % 128);
      const int tt = static_cast<int>((r / (64 * 128)) % kk);
      const int gi = static_cast<int>((r / (64L * 128 * kk)) % (Cin / 64));
      const int go = static_cast<int>(r / (64L * 128 * kk * (Cin / 64)));
      const int co = go * 128 + co_lo, ci = gi * 64 + ci_lo;
      const int t = ksize == 3 ? tt : 4;
      wf[l * dst_layer + (((static_cast<long>(go) * (Cin / 64) + gi) * 9 + t) * 128 + co_lo) * 64 + ci_lo] =
          __float2bfloat16(__ldg(w + l * src_layer + (static_cast<long>(co) * Cin + ci) * kk + tt));
    }
    if (wd) {
      const int co_lo = static_cast<int>(r % 64);
      const int ci_lo = static_cast<int>((r / 64) % 128);
      const int tt = static_cast<int>((r / (64 * 128)) % kk);
      const int go = static_cast<int>((r / (64L * 128 * kk)) % (Cout / 64));
      const int gi = static_cast<int>(r / (64L * 128 * kk * (Cout / 64)));
      const int co = go * 64 + co_lo, ci = gi * 128 + ci_lo;
      const int t = ksize == 3 ? tt : 4;               // destination tap; the source tap is the flipped one
      wd[l * dst_layer + (((static_cast<long>(gi) * (Cout / 64) + go) * 9 + t) * 128 + ci_lo) * 64 + co_lo] =
          __float2bfloat16(__ldg(w + l * src_layer + (static_cast<long>(co) * Cin + ci) * kk + (ksize == 3 ? 8 - tt : 0)));
    }
  }
}

}  // namespace
}  // namespace fd

extern "C" FD_API int fd_debug_wide_timing(unsigned long long* out, int reset) {
  unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  cudaError_t e = cudaMemcpyFromSymbol(out, fd::g_wide_dbg, sizeof(z));
  if (e == cudaSuccess && reset) e = cudaMemcpyToSymbol(fd::g_wide_dbg, z, sizeof(z));
  return static_cast<int>(e);
}

namespace fd {
namespace {
// Tiled 3x3 pack: one block = (layer, 32 output channels, 32 input channels).  The 32 x 288 source floats are read as
// contiguous runs, transposed through shared memory, and written as 64-byte runs of both packings (forward: 32 cins of one
// cout and tap; dgrad: 32 couts of one cin and tap).  The element-per-thread kernel gathers with a stride of 36 B (forward)
// / 4.6 KB (dgrad) and took 46 us for the 20 layers of PoolResnet(128).
constexpr int kPkT = 32;
__global__ void __launch_bounds__(256)
pack_conv3x3_wide_tiled_kernel(const float* __restrict__ w, int Cout, int Cin, __nv_bfloat16* __restrict__ wf,
                               __nv_bfloat16* __restrict__ wd) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm_pk[kPkT][kPkT * 9 + 1];
  const int tci = Cin / kPkT, tco = Cout / kPkT;
  const int l = blockIdx.x / (tci * tco), rem = blockIdx.x % (tci * tco);
  const int co0 = (rem / tci) * kPkT, ci0 = (rem % tci) * kPkT;
  const long per_layer = static_cast<long>(Cout) * Cin * 9;
  for (int i = threadIdx.x; i < kPkT * kPkT * 9; i += 256) {
    const int co = i / (kPkT * 9), r = i - co * (kPkT * 9);                  // r = ci_local * 9 + t: contiguous in the source
    sm_pk[co][r] = __ldg(w + l * per_layer + (static_cast<long>(co0 + co) * Cin + ci0) * 9 + r);
  }
  __syncthreads();
  if (wf) {       // wf[l][co / 128][ci / 64][t][co % 128][ci % 64]
    for (int i = threadIdx.x; i < kPkT * kPkT * 9; i += 256) {
      const int ci = i % kPkT, co = (i / kPkT) % kPkT, t = i / (kPkT * kPkT);
      const int cog = co0 + co, cig = ci0 + ci;
      wf[l * per_layer + (((static_cast<long>(cog / 128) * (Cin / 64) + cig / 64) * 9 + t) * 128 + cog % 128) * 64 + cig % 64] =
          __float2bfloat16(sm_pk[co][ci * 9 + t]);
    }
  }
  if (wd) {       // wd[l][ci / 128][co / 64][8 - t][ci % 128][co % 64]
    for (int i = threadIdx.x; i < kPkT * kPkT * 9; i += 256) {
      const int co = i % kPkT, ci = (i / kPkT) % kPkT, t = i / (kPkT * kPkT);
      const int cog = co0 + co, cig = ci0 + ci;
      wd[l * per_layer + (((static_cast<long>(cig / 128) * (Cout / 64) + cog / 64) * 9 + (8 - t)) * 128 + cig % 128) * 64 + cog % 64] =
          __float2bfloat16(sm_pk[co][ci * 9 + t]);
    }
  }
}
}  // namespace
}  // namespace fd

static int pack_conv_wide(const float* w, int n_layers, int Cout, int Cin, int ksize, fd_bf16* w_fwd, fd_bf16* w_dgrad,
                          void* stream) {
  using namespace fd;
  if (!w || n_layers <= 0 || (!w_fwd && !w_dgrad)) return FD_EINVAL;
  if (Cout <= 0 || Cin <= 0 || Cout % 64 != 0 || Cin % 64 != 0) return FD_EUNSUPPORTED;
  if ((w_fwd && Cout % 128 != 0) || (w_dgrad && Cin % 128 != 0)) return FD_EUNSUPPORTED;
  if (ksize == 3 && !getenv("FD_PACK_WIDE_GATHER")) {         // channel counts are multiples of 64: 32 x 32 tiles always fit
    launch_k(pack_conv3x3_wide_tiled_kernel, dim3(n_layers * (Cout / kPkT) * (Cin / kPkT)), dim3(256), 0,
             static_cast<cudaStream_t>(stream), w, Cout, Cin, reinterpret_cast<__nv_bfloat16*>(w_fwd),
             reinterpret_cast<__nv_bfloat16*>(w_dgrad));
    count_launch();
    return launch_status();
  }
  const long total = static_cast<long>(n_layers) * Cout * Cin * ksize * ksize;
  const long blocks = (total + 255) / 256;
  launch_k(pack_conv_wide_kernel, dim3(static_cast<unsigned>(blocks < 148 * 16 ? blocks : 148 * 16)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), w, n_layers, Cout, Cin, ksize, reinterpret_cast<__nv_bfloat16*>(w_fwd),
           reinterpret_cast<__nv_bfloat16*>(w_dgrad));
  count_launch();
  return launch_status();
}

extern "C" int fd_pack_conv3x3_wide(const float* w, int n_layers, int Cout, int Cin, fd_bf16* w_fwd, fd_bf16* w_dgrad,
                                    void* stream) {
  return pack_conv_wide(w, n_layers, Cout, Cin, 3, w_fwd, w_dgrad, stream);
}

extern "C" int fd_pack_conv1x1_wide(const float* w, int n_layers, int Cout, int Cin, fd_bf16* w_fwd, fd_bf16* w_dgrad,
                                    void* stream) {
  return pack_conv_wide(w, n_layers, Cout, Cin, 1, w_fwd, w_dgrad, stream);
}

namespace fd {
namespace {
constexpr size_t kWideSmemCap = 227 * 1024 - 64;     // dynamic shared memory: the kernel also has 16 bytes of static (timing) state
inline int wide_cta_group() {
  static const int cg_env = [] { const char* e = getenv("FD_WIDE_CTA_GROUP"); return e ? atoi(e) : 2; }();
  return cg_env == 1 ? 1 : 2;
}
inline bool wide_share_allowed(int cg) { return cg == 2 && !getenv("FD_WIDE_NO_SHARE"); }
// Tiling: TW <= 62 output columns, R rows with R * (TW + 2) <= 256 GEMM rows (two 128-row blocks = 2 x 128 TMEM columns,
// double buffered).  Cost model: tensor time of the padded blocks times the number of waves over the CTA groups; a map
// with fewer two-block tiles than SM pairs runs in shared-tile mode (one block per CTA).
inline void wide_pick_tiling(int cg, int nout, int tpg, int B, int H, int W, int* bestR, int* bestTW) {
  const int nsm = sm_count();
  const bool share_ok = wide_share_allowed(cg);
  double best = 1e30;
  *bestR = *bestTW = 0;
  const int min_tw_tiles = (W + 61) / 62;
  for (int tw_tiles = min_tw_tiles; tw_tiles <= min_tw_tiles + 2; ++tw_tiles) {
    const int TW = (W + tw_tiles - 1) / tw_tiles;
    if (TW > 62 || TW < 1) continue;
    const int Wp = TW + 2;
    for (int R = 1; R <= H && R + 2 <= 256; ++R) {
      const int nblk = (R * Wp + 127) / 128;
      if (nblk > 2) break;
      if (wide_slots(cg, tpg, R, Wp, TW, kWideSmemCap, nout) == 0) break;
      const long tiles = static_cast<long>(B) * ((H + R - 1) / R) * tw_tiles;
      const long waves = (tiles + nsm - 1) / nsm;
      double cost = waves * (2400.0 * nblk + 600.0);
      if (share_ok && nblk == 2 && tiles <= nsm / 2) cost = 2400.0 + 600.0;     // shared-tile mode: one block per CTA
      if (cost < best) { best = cost; *bestR = R; *bestTW = TW; }
    }
  }
}
inline bool wide_is_shared_tile(int cg, int nout, int tpg, int B, int H, int W) {
  int R = 0, TW = 0;
  wide_pick_tiling(cg, nout, tpg, B, H, W, &R, &TW);
  if (R == 0) return false;
  const int Wp = TW + 2, nblk = (R * Wp + 127) / 128;
  const long tiles = static_cast<long>(B) * ((H + R - 1) / R) * ((W + TW - 1) / TW);
  return wide_share_allowed(cg) && nblk == 2 && tiles <= sm_count() / 2;
}
}  // namespace
}  // namespace fd

// 1 when fd_conv3x3_wide will run this shape in SHARED-TILE mode (fewer two-block tiles than SM pairs: the pair works on one
// tile) -- the mode in which it can write BOTH outputs (out and out2) in one launch.
extern "C" int fd_conv3x3_wide_shared_tile(int B, int H, int W, int flags) {
  using namespace fd;
  if (B <= 0 || H <= 0 || W <= 0) return 0;
  return wide_is_shared_tile(wide_cta_group(), 128, (flags & FD_CONV_1X1) ? 1 : 3, B, H, W) ? 1 : 0;
}

// nout = 128: the wide API below; nout = 64: one output plane -- the 64-channel layers of conv3x3_tc.cu on CTA pairs
// (w_packed = the tap-major [9][64][64] packing of fd_pack_conv3x3, every per-plane array has ONE entry)
int fd::conv3x3_pairs(int nout, const fd_bf16* const* x, int gin, const fd_bf16* w_packed, int B, int H, int W,
                      const float* bias, float slope, const float* const* chan_scale, const fd_bf16* const* residual,
                      uint32_t* const* mask_out, fd_bf16* const* out, const uint32_t* const* mask_in,
                      const float* const* chan_scale2, fd_bf16* const* out2, int flags, void* stream) {
  using namespace fd;
  if (!x || !w_packed || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (gin < 1 || gin > kMaxGin || (nout != 64 && nout != 128)) return FD_EUNSUPPORTED;
  const int np = nout / 64;
  for (int g = 0; g < gin; ++g)
    if (!x[g]) return FD_EINVAL;
  auto all_set = [np](auto arr) {
    if (!arr) return false;
    for (int g = 0; g < np; ++g)
      if (!arr[g]) return false;
    return true;
  };
  bool has_out = all_set(out), has_out2 = all_set(out2);
  if (!has_out && !has_out2) return FD_EINVAL;
  const bool dual = has_out && has_out2;                     // both outputs: the shared-tile mode only (checked below)
  if (mask_in && !has_out2) return FD_EINVAL;
  if (has_out2 && !all_set(mask_in)) return FD_EINVAL;
  if (dual) has_out2 = false;                                // `out` is the primary (staged) output
  if (!(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;
  const int cg = wide_cta_group();
  const int nsm = sm_count();
  const size_t smem_cap = kWideSmemCap;
  const int tpg = (flags & FD_CONV_1X1) ? 1 : 3;
  const bool share_ok = wide_share_allowed(cg);
  int bestR = 0, bestTW = 0;
  wide_pick_tiling(cg, nout, tpg, B, H, W, &bestR, &bestTW);
  if (bestR == 0) return FD_EUNSUPPORTED;

  WideParams p;
  p.B = B; p.H = H; p.W = W; p.R = bestR; p.TW = bestTW; p.Wp = bestTW + 2;
  p.nblk = (bestR * p.Wp + 127) / 128;
  p.tiles_w = (W + bestTW - 1) / bestTW;
  p.tiles_h = (H + bestR - 1) / bestR;
  p.num_tiles = B * p.tiles_w * p.tiles_h;
  p.gin = gin;
  p.tap_lo = (flags & FD_CONV_1X1) ? 4 : 0;
  p.tap_hi = (flags & FD_CONV_1X1) ? 5 : 9;
  p.in_bytes = static_cast<uint32_t>((bestR + 2) * p.Wp * 128);
  p.in_buf_bytes = static_cast<uint32_t>(wide_in_buf_bytes(bestR, p.Wp));
  p.split = wide_split(bestR, p.Wp) ? 1 : 0;
  p.rpb = wide_unit_rows(bestR, p.Wp);
  p.units = p.split ? p.nblk : 1;
  p.stg_bytes = static_cast<uint32_t>(p.rpb * bestTW * 128);
  p.stg_buf_bytes = static_cast<uint32_t>(wide_stg_buf_bytes(bestR, p.Wp, bestTW));
  // shared-tile mode: two-block tiles, fewer of them than SM pairs, and the pair kernel
  p.share = (share_ok && p.nblk == 2 && p.num_tiles <= nsm / 2) ? 1 : 0;
  fd_bf16* const* staged_planes = has_out ? out : out2;
  const bool has_res = all_set(residual);
  if (dual && !p.share) return FD_EUNSUPPORTED;              // fd_conv3x3_wide_shared_tile() tells the caller beforehand
  for (int g = 0; g < 2; ++g) {
    const int gs = g < np ? g : 0;
    p.res_ptr[g] = has_res ? reinterpret_cast<const __nv_bfloat16*>(residual[gs]) : nullptr;
    p.out_ptr[g] = reinterpret_cast<__nv_bfloat16*>(staged_planes[gs]);
    p.out2_ptr[g] = dual ? reinterpret_cast<__nv_bfloat16*>(out2[gs]) : nullptr;
  }
  p.tpg = tpg;
  p.ngrp = (flags & FD_CONV_1X1) ? 1 : 3;
  p.wslots = wide_slots(cg, tpg, bestR, p.Wp, bestTW, smem_cap, nout);
  p.inv_wp = static_cast<uint32_t>((65536 + p.Wp - 1) / p.Wp);
  p.flags = flags;
  { static const int dbg = [] { const char* d = getenv("FD_WIDE_TIMING"); return d ? atoi(d) : 0; }(); p.dbg = dbg; }
  p.slope = slope;
  p.bias = bias;
  p.has_res = has_res ? 1 : 0;
  p.staged_out2 = has_out2 ? 1 : 0;
  for (int g = 0; g < 2; ++g) {
    const int gs = g < np ? g : 0;
    p.chan_scale[g] = chan_scale ? chan_scale[gs] : nullptr;
    p.chan_scale2[g] = chan_scale2 ? chan_scale2[gs] : nullptr;
    p.mask_in[g] = mask_in ? reinterpret_cast<const uint16_t*>(mask_in[gs]) : nullptr;
    p.mask_out[g] = mask_out ? reinterpret_cast<uint16_t*>(mask_out[gs]) : nullptr;
  }
  if ((p.chan_scale[0] == nullptr) != (p.chan_scale[1] == nullptr)) return FD_EINVAL;
  if ((p.chan_scale2[0] == nullptr) != (p.chan_scale2[1] == nullptr)) return FD_EINVAL;
  if ((p.mask_out[0] == nullptr) != (p.mask_out[1] == nullptr)) return FD_EINVAL;

  WideMaps maps;
  int rc;
  for (int g = 0; g < kMaxGin; ++g) {
    rc = make_tmap_nhwc_bf16(&maps.in[g], x[g < gin ? g : 0], B, H, W, kC, p.Wp, bestR + 2);
    if (rc != FD_OK) return rc;
  }
  rc = make_tmap_2d_bf16(&maps.w, w_packed, gin * 9 * nout, kC, cg == 2 ? nout / 2 : nout, kC);
  if (rc != FD_OK) return rc;
  fd_bf16* const* staged = has_out ? out : out2;
  for (int g = 0; g < 2; ++g) {
    const int gs = g < np ? g : 0;
    rc = make_tmap_nhwc_bf16(&maps.out[g], staged[gs], B, H, W, kC, bestTW, p.rpb);
    if (rc != FD_OK) return rc;
    rc = make_tmap_nhwc_bf16(&maps.res[g], p.has_res ? residual[gs] : x[0], B, H, W, kC, bestTW, p.rpb);
    if (rc != FD_OK) return rc;
  }
  p.resident = 0;
  {
    // all gin x ngrp chunks fit beside the buffers: keep them resident (the 64-output instantiation: 36 KB per CTA)
    const int all = gin * p.ngrp;
    if (!p.share && all <= kMaxWSlots && wide_smem_fixed(bestR, p.Wp, bestTW) + all * wide_chunk_bytes(cg, tpg, nout) <= smem_cap &&
        !getenv("FD_WIDE_NO_RESIDENT")) {
      p.resident = 1;
      p.wslots = all;
    }
  }
  size_t smem = wide_smem_fixed(bestR, p.Wp, bestTW) + p.wslots * wide_chunk_bytes(cg, tpg, nout);
  if (p.share) {
    // no staging buffers (outputs leave through plain stores); 16 KB of guard space below the input ring instead
    const size_t fixed = wide_smem_fixed(bestR, p.Wp, bestTW) - 2 * wide_stg_buf_bytes(bestR, p.Wp, bestTW) + 16384;
    size_t nslots = (smem_cap - fixed) / wide_chunk_bytes(cg, tpg, nout);
    if (nslots > kMaxWSlots) nslots = kMaxWSlots;
    p.wslots = static_cast<int>(nslots);
    p.stg_buf_bytes = 0;
    smem = fixed + nslots * wide_chunk_bytes(cg, tpg, nout);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nout == 64) return cg == 2 ? launch_wide<2, 64>(maps, p, smem, st) : launch_wide<1, 64>(maps, p, smem, st);
  return cg == 2 ? launch_wide<2, 128>(maps, p, smem, st) : launch_wide<1, 128>(maps, p, smem, st);
}

extern "C" int fd_conv3x3_wide(const fd_bf16* const* x, int gin, const fd_bf16* w_packed, int B, int H, int W,
                               const float* bias, float slope, const float* const* chan_scale,
                               const fd_bf16* const* residual, uint32_t* const* mask_out, fd_bf16* const* out,
                               const uint32_t* const* mask_in, const float* const* chan_scale2, fd_bf16* const* out2,
                               int flags, void* stream) {
  return fd::conv3x3_pairs(128, x, gin, w_packed, B, H, W, bias, slope, chan_scale, residual, mask_out, out, mask_in,
                           chan_scale2, out2, flags, stream);
}

// ---- fd_conv3x3_wide_chain: host side ----
namespace fd {
namespace {
// the chain kernel needs a whole image per CTA pair (one two-block tile) and a pair per image
inline bool wide_chain_ok(int B, int H, int W) {
  if (wide_cta_group() != 2 || !wide_share_allowed(2) || getenv("FD_WIDE_NO_CHAIN")) return false;
  const int rows = H * (W + 2);
  return B >= 1 && B <= sm_count() / 2 && rows > 128 && rows <= 256;
}
}  // namespace
}  // namespace fd

extern "C" int fd_conv3x3_wide_chain_ok(int B, int H, int W) { return fd::wide_chain_ok(B, H, W) ? 1 : 0; }

extern "C" int fd_conv3x3_wide_chain(const fd_bf16* const* x_stack, int n_stack, const fd_bf16* w_packed, int n_w_layers, int B,
                                     int H, int W, float slope, const fd_wide_chain_layer* layers, int n_layers, void* stream) {
  using namespace fd;
  if (!x_stack || !x_stack[0] || !x_stack[1] || !w_packed || !layers || B <= 0 || H <= 0 || W <= 0 || n_stack <= 0 ||
      n_w_layers <= 0 || n_layers <= 0)
    return FD_EINVAL;
  if (n_layers > kMaxChain || !wide_chain_ok(B, H, W) || !(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;
  ChainParams p;
  p.B = B; p.H = H; p.W = W; p.Wp = W + 2; p.nlayers = n_layers;
  p.in_bytes = static_cast<uint32_t>((H + 2) * p.Wp * 128);
  p.in_buf_bytes = static_cast<uint32_t>(wide_in_buf_bytes(H, p.Wp));
  p.inv_wp = static_cast<uint32_t>((65536 + p.Wp - 1) / p.Wp);
  p.slope = slope;
  for (int l = 0; l < n_layers; ++l) {
    const fd_wide_chain_layer& s = layers[l];
    ChainLayer& d = p.layer[l];
    if (s.in_index < 0 || s.in_index >= n_stack || s.w_index < 0 || s.w_index >= n_w_layers) return FD_EINVAL;
    const bool has_out = s.out[0] && s.out[1], has_out2 = s.out2[0] && s.out2[1], has_res = s.residual[0] && s.residual[1];
    if (!has_out && !has_out2) return FD_EINVAL;
    if ((s.out[0] == nullptr) != (s.out[1] == nullptr) || (s.out2[0] == nullptr) != (s.out2[1] == nullptr) ||
        (s.residual[0] == nullptr) != (s.residual[1] == nullptr) || (s.chan_scale[0] == nullptr) != (s.chan_scale[1] == nullptr) ||
        (s.chan_scale2[0] == nullptr) != (s.chan_scale2[1] == nullptr) || (s.mask_out[0] == nullptr) != (s.mask_out[1] == nullptr) ||
        (s.mask_in[0] == nullptr) != (s.mask_in[1] == nullptr))
      return FD_EINVAL;
    if (has_out2 && !s.mask_in[0]) return FD_EINVAL;          // the second output is the MASKED copy
    d.in_n0 = s.in_index * B;
    d.w_row0 = s.w_index * 2 * 9 * 128;
    d.lrelu = (s.flags & FD_EPI_LRELU) ? 1 : 0;
    d.has_res = has_res ? 1 : 0;
    d.bias = s.bias;
    for (int g = 0; g < 2; ++g) {
      d.res_ptr[g] = reinterpret_cast<const __nv_bfloat16*>(s.residual[g]);
      d.out_ptr[g] = reinterpret_cast<__nv_bfloat16*>(s.out[g]);
      d.out2_ptr[g] = reinterpret_cast<__nv_bfloat16*>(s.out2[g]);
      d.chan_scale[g] = s.chan_scale[g];
      d.chan_scale2[g] = s.chan_scale2[g];
      d.mask_in[g] = reinterpret_cast<const uint16_t*>(s.mask_in[g]);
      d.mask_out[g] = reinterpret_cast<uint16_t*>(s.mask_out[g]);
    }
  }
  const size_t fixed = 16384 + ((kChainInBufs * static_cast<size_t>(p.in_buf_bytes) + 1023) / 1024 * 1024) + 3 * kNOut * 4 + 512 + 1024;
  const size_t chunk = wide_chunk_bytes(2, 3, 128);
  if (fixed + kMinWSlots * chunk > kWideSmemCap) return FD_EUNSUPPORTED;
  size_t nslots = (kWideSmemCap - fixed) / chunk;
  if (nslots > kMaxWSlots) nslots = kMaxWSlots;
  p.wslots = static_cast<int>(nslots);
  const size_t smem = fixed + nslots * chunk;

  ChainMaps maps;
  int rc;
  for (int g = 0; g < 2; ++g) {
    rc = make_tmap_nhwc_bf16(&maps.in[g], x_stack[g], n_stack * B, H, W, kC, p.Wp, H + 2);
    if (rc != FD_OK) return rc;
  }
  rc = make_tmap_2d_bf16(&maps.w, w_packed, n_w_layers * 2 * 9 * 128, kC, 64, kC);
  if (rc != FD_OK) return rc;
  cudaError_t e = set_max_dyn_smem(conv3x3_wide_chain_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * B);
  cfg.blockDim = dim3(kThreadsW);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, conv3x3_wide_chain_kernel, maps, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}
