// Weight gradient of the 3x3 convolution for the WIDE layers (>= 128 channels on 64-channel planes: filters = 128 of
// train_model.py:17, the 128 / 256-channel blocks of models/SSD.py:164-189) on tcgen05.mma.cta_group::2.
//
//   dW[t][ci][co] = sum_pixels  xpad[p + off_t][ci] * g[p][co],   off_t = ky*Wp + kx        (wgrad3x3_tc.cu)
//
// One CTA PAIR works on two input planes and two gradient planes at once -- a 128 x 128 channel block of dW:
//   * CTA r loads the halo tile of x plane (2hh + r) and the tile of g plane (2gg + r) -- exactly the loads of the
//     64-channel kernel -- and ONE M=256, N=128, K=16 instruction multiplies them all: the M rows of CTA r are a stacked
//     tap pair of ITS x plane (two 64-channel MN-major atoms, LBO = the taps' row distance), the N columns 64c..64c+63
//     come from the g tile in CTA c's shared memory (the pair's B operand is split between the two CTAs).  Each CTA
//     reads 4 KB (A) + 2 KB (its half of B) per 64 clk of tensor work; the four 64 x 64 launches this replaces read
//     6 KB per 32 clk and load every tile twice.
//   * TMEM holds 128 lanes x 512 columns = four tap pairs x 128 couts: EIGHT taps.  A 128 x 64 x 9 fp32 block does not
//     fit (295 KB > 256 KB), so the ninth tap is a second pass of the same kernel with the single pair (7, 8), whose tap-7
//     half is dropped.  Pass 1: pairs (0,1) (2,3) (4,5) (6,7), 4 MMAs per 16 pixels; pass 2: 1 MMA per 16 pixels.
//   * accumulators live in TMEM across all tiles of the persistent pair and are reduced into the packed fp32 gradient
//     blocks [sub-block (g, h)][tap][ci][co] once per CTA with TMA reduce-stores.
//   * the bias gradient (column sums of g) is taken by the epilogue warps of the CTA that holds that g plane, from the
//     shared-memory tile, once the MMAs of the stage are done and before the stage is released.
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kThreadsGW = 192;
constexpr int kMaxStripsW = 8;

struct WgradWideParams {
  int H, W, R, TW, Wp, tiles_w, tiles_per_img, ksteps;
  uint32_t x_bytes, g_bytes, x_buf_bytes, g_buf_bytes;
  int nprob, tiles_per_prob, pairs_per_prob, per;
  int npairs;                 // tap pairs of this pass
  int pair_t0[4];             // first tap of every pair (the second is t0 + 1)
  int drop_first;             // bit j: the first tap of pair j is a duplicate (the pair (7, 8)): its rows are reduced as zeros
  int sub_row[2][2];          // [r][c]: row of sub-block (g = 2gg + c, h = 2hh + r) in the [., 64] fp32 view of dw, problem 0
  int dw_rows_per_prob;
  float* dbias[2];            // [r]: bias gradient of g plane 2gg + r, problem 0 (nullable)
  long dbias_stride;
  uint32_t bar_off;
};

struct WgradWideMaps {
  CUtensorMap x[2];
  CUtensorMap g[2][kMaxStripsW];
  CUtensorMap dw;
};

__device__ __forceinline__ void tma_load_4d_2cta(void* smem_dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}

__global__ void __launch_bounds__(kThreadsGW, 1)
wgrad3x3_wide_kernel(const __grid_constant__ WgradWideMaps maps, const __grid_constant__ WgradWideParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const uint32_t stage_bytes = p.x_buf_bytes + p.g_buf_bytes;
  uint8_t* sStage = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* full = bars + 0;        // [2]  leader's copy is live: both CTAs' loads
  uint64_t* mma_done = bars + 2;    // [2]  per CTA, multicast commit: the MMAs have read stage s
  uint64_t* empty = bars + 4;       // [2]  per CTA: the four epilogue warps are done with the g tile of stage s
  uint64_t* acc_full = bars + 6;    //      per CTA, multicast commit
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  float* sBias = reinterpret_cast<float*>(bars + 8);  // [128]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  // K-padding rows of g and the over-read tail of x are never written by TMA and must stay zero.
  for (uint32_t i = threadIdx.x * 16u; i < 2 * stage_bytes; i += kThreadsGW * 16u)
    *reinterpret_cast<uint4*>(sStage + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.x[rank]);
    tma_prefetch_desc(&maps.g[rank][0]);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, 1);
      mbar_init(mma_done + s, 1);
      mbar_init(empty + s, 4);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers and zeroed buffers exist before any remote completion / MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // contiguous chunk of tiles (of ONE problem) for this pair
  const int pair = static_cast<int>(blockIdx.x) >> 1;
  const int prob = pair / p.pairs_per_prob;
  const int chunk = pair - prob * p.pairs_per_prob;
  const int tile_begin = prob * p.tiles_per_prob + min(p.tiles_per_prob, chunk * p.per);
  const int tile_end = prob * p.tiles_per_prob + min(p.tiles_per_prob, (chunk + 1) * p.per);
  float* const dbias_out = p.dbias[rank] ? p.dbias[rank] + prob * p.dbias_stride : nullptr;

  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t full_leader0 = mapa_shared(smem_u32(full), 0);
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / p.tiles_per_img;
        const int rem = tile - n * p.tiles_per_img;
        const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
        const int h0 = th * p.R, w0 = tw * p.TW;
        mbar_wait(empty + s, ph ^ 1);
        if (leader) mbar_expect_tx(full + s, 2u * (p.x_bytes + p.g_bytes));
        tma_load_4d_2cta(sStage + s * stage_bytes, &maps.x[rank], full_leader0 + 8u * s, 0, w0 - 1, h0 - 1, n);
        tma_load_4d_2cta(sStage + s * stage_bytes + p.x_buf_bytes, &maps.g[rank][tw], full_leader0 + 8u * s, 0, 0, h0, n);
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(256, 128, 1, 1);  // both operands MN-major
      int it = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t x_addr = smem_u32(sStage + s * stage_bytes);
        const uint32_t g_lo = sdesc_lo(x_addr + p.x_buf_bytes, 1024);
        // pair j reads the halo tile at row offset off0 with the second 64-channel atom (off1 - off0) rows further (LBO)
        uint32_t a_lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int t0 = p.pair_t0[j < p.npairs ? j : 0], t1 = t0 + 1;
          const int off0 = (t0 / 3) * p.Wp + (t0 % 3);
          const int off1 = (t1 / 3) * p.Wp + (t1 % 3);
          a_lo[j] = sdesc_lo(x_addr + static_cast<uint32_t>(off0 * 128), static_cast<uint32_t>((off1 - off0) * 128));
        }
        for (int ks = 0; ks < p.ksteps; ++ks) {
          const uint64_t bd = sdesc_sw128(g_lo + ks * 128);      // 16 pixel rows = 2048 B per K step
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < p.npairs)
              umma_2cta(tmem_base + j * 128, sdesc_sw128(a_lo[j] + ks * 128), bd, idesc, (it | ks) != 0 ? 1u : 0u);
        }
        umma_commit_2cta(mma_done + s);
        if (tile + 1 == tile_end) umma_commit_2cta(acc_full);
      }
    }
    __syncwarp();
  } else {
    // epilogue warps: (1) once the MMAs have read a stage: bias gradient from its g tile, then release the stage
    const int et = threadIdx.x - 64;      // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum = 0.f;
    int it = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(mma_done + s, ph);
      if (dbias_out) {
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
    // (2) drain this CTA's accumulators: TMEM -> registers -> fp32 staging tiles [128 rows][32 fp32] (128B swizzle of the
    // tensor map) -> TMA reduce-stores (.add in L2) into the packed gradient blocks.  Two 64 KB halves of the (free) stage
    // buffers alternate between tap pairs.
    if (tile_begin < tile_end) {
      const int q = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      asm volatile("bar.sync 1, 128;" ::: "memory");      // every warp's column sums of the last tile are done: the staging
                                                           // tiles below overwrite the stage buffers
      const int row = q * 32 + lane;           // = tsel * 64 + ci
      const uint32_t sw = static_cast<uint32_t>(row) & 7u;
#pragma unroll 1
      for (int j = 0; j < p.npairs; ++j) {
        const bool dup = ((p.drop_first >> j) & 1) && row < 64;
        uint8_t* base = sStage + static_cast<size_t>(j & 1) * (4 * 128 * 128);
        if (j >= 2) {          // the reduce-stores of pair j - 2 must have read this half
          if (et == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {       // column quarter: g plane 2gg + qq / 2, couts (qq % 2) * 32 ...
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(j * 128 + qq * 32), acc);
          tmem_ld_wait();
          uint8_t* tile = base + static_cast<size_t>(qq) * (128 * 128) + row * 128;
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
          const int tap0 = p.pair_t0[j];
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const int grow = p.sub_row[rank][qq >> 1] + prob * p.dw_rows_per_prob + tap0 * kC;
            asm volatile(
                "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                    reinterpret_cast<uint64_t>(&maps.dw)),
                "r"(smem_u32(base + static_cast<size_t>(qq) * (128 * 128))), "r"((qq & 1) * 32), "r"(grow)
                : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      if (dbias_out) {
        sBias[et] = bsum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < 64) atomicAdd(dbias_out + et, sBias[et] + sBias[et + 64]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // nobody exits (or frees TMEM) while the pair's MMAs / commits may still touch its memory
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

}  // namespace
}  // namespace fd

// x0, x1: input planes 2hh, 2hh+1; g0, g1: gradient planes 2gg, 2gg+1; each holds `nprob` stacked [B,H,W,64] tensors.
// dw_packed: base of the packed fp32 gradient blocks; sub_off[r][c] = ELEMENT offset of sub-block (g = 2gg + c, h = 2hh + r)
// ([9][64][64] fp32 each) for problem 0; problem q adds q * dw_stride elements.  dbias0 / dbias1 (nullable): [64] bias
// gradients of the two g planes, problem q at + q * dbias_stride.
extern "C" int fd_conv3x3_wgrad_wide(const fd_bf16* x0, const fd_bf16* x1, const fd_bf16* g0, const fd_bf16* g1, int nprob, int B,
                                     int H, int W, float* dw_packed, const long* sub_off, long dw_stride, float* dbias0,
                                     float* dbias1, long dbias_stride, int flags, void* stream) {
  using namespace fd;
  (void)flags;
  if (!x0 || !x1 || !g0 || !g1 || !dw_packed || !sub_off || nprob <= 0 || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  for (int i = 0; i < 4; ++i)
    if (sub_off[i] < 0 || sub_off[i] % kC != 0) return FD_EINVAL;
  if (nprob > 1 && dw_stride % kC != 0) return FD_EINVAL;
  const int nsm = sm_count();
  const int tiles_w = (W + 61) / 62;
  if (tiles_w > kMaxStripsW) return FD_EUNSUPPORTED;
  const int TW = (W + tiles_w - 1) / tiles_w;
  const int Wp = TW + 2;
  const size_t smem_cap = 227 * 1024;
  const size_t drain = 2 * 4 * 128 * 128;       // two halves of four [128 rows][128 B] staging tiles

  // rows per tile: the tallest tile whose two stages fit minimises the halo re-reads; among those, the one that
  // balances the pairs best (tiles per pair x K steps per tile)
  int bestR = 0;
  double best = 1e30;
  for (int R = 1; R <= H && R + 2 <= 256; ++R) {
    const int ksteps = (R * Wp + 15) / 16;
    const size_t xb = (static_cast<size_t>(ksteps * 16 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024;
    const size_t gb = (static_cast<size_t>(ksteps * 16) * 128 + 1023) / 1024 * 1024;
    const size_t stages = 2 * (xb + gb);
    if ((stages > drain ? stages : drain) + 1024 + 1024 > smem_cap) break;
    const long tiles_per_prob = static_cast<long>(B) * ((H + R - 1) / R) * tiles_w;
    long ppp = (nsm / 2) / nprob;
    if (ppp < 1) ppp = 1;
    if (ppp > tiles_per_prob) ppp = tiles_per_prob;
    const long per = (tiles_per_prob + ppp - 1) / ppp;
    const double mma = ksteps * 4 * 70.0;
    const double fill = static_cast<double>(2 * R + 2) * Wp * 128 / 48.0;
    const double cost = per * ((mma > fill ? mma : fill) + 1500.0);
    if (cost < best) { best = cost; bestR = R; }
  }
  if (bestR == 0) return FD_EUNSUPPORTED;

  WgradWideParams p;
  p.H = H; p.W = W; p.R = bestR; p.TW = TW; p.Wp = Wp; p.tiles_w = tiles_w;
  p.tiles_per_img = ((H + bestR - 1) / bestR) * tiles_w;
  p.ksteps = (bestR * Wp + 15) / 16;
  p.x_bytes = static_cast<uint32_t>((bestR + 2) * Wp * 128);
  p.g_bytes = static_cast<uint32_t>(bestR * Wp * 128);
  p.x_buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.ksteps * 16 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024);
  p.g_buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.ksteps * 16) * 128 + 1023) / 1024 * 1024);
  p.nprob = nprob;
  p.tiles_per_prob = B * p.tiles_per_img;
  int ppp = (nsm / 2) / nprob;
  if (ppp < 1) ppp = 1;
  if (ppp > p.tiles_per_prob) ppp = p.tiles_per_prob;
  p.per = (p.tiles_per_prob + ppp - 1) / ppp;
  p.pairs_per_prob = (p.tiles_per_prob + p.per - 1) / p.per;
  for (int r = 0; r < 2; ++r)
    for (int c = 0; c < 2; ++c) p.sub_row[r][c] = static_cast<int>(sub_off[r * 2 + c] / kC);
  p.dw_rows_per_prob = static_cast<int>(dw_stride / kC);
  p.dbias[0] = dbias0; p.dbias[1] = dbias1;
  p.dbias_stride = dbias_stride;
  const size_t stages = 2 * static_cast<size_t>(p.x_buf_bytes + p.g_buf_bytes);
  p.bar_off = static_cast<uint32_t>(stages > drain ? stages : drain);
  const size_t smem = p.bar_off + 1024 + 1024;

  WgradWideMaps maps;
  const fd_bf16* xs[2] = {x0, x1};
  const fd_bf16* gs[2] = {g0, g1};
  int rc;
  for (int r = 0; r < 2; ++r) {
    rc = make_tmap_nhwc_bf16(&maps.x[r], xs[r], nprob * B, H, W, kC, Wp, bestR + 2);
    if (rc != FD_OK) return rc;
    for (int tw = 0; tw < kMaxStripsW; ++tw) {
      const int w0 = (tw < tiles_w ? tw : 0) * TW;
      const int wext = (W - w0 < TW) ? W - w0 : TW;
      rc = make_tmap_nhwc_bf16_strided(&maps.g[r][tw], gs[r] + static_cast<size_t>(w0) * kC, nprob * B, H, wext, W, kC, Wp, bestR);
      if (rc != FD_OK) return rc;
    }
  }
  long max_row = 0;
  for (int i = 0; i < 4; ++i)
    if (sub_off[i] / kC > max_row) max_row = sub_off[i] / kC;
  rc = make_tmap_2d_f32(&maps.dw, dw_packed, static_cast<long>(nprob - 1) * p.dw_rows_per_prob + max_row + 9 * kC, kC, 128, 32);
  if (rc != FD_OK) return rc;

  cudaError_t e = set_max_dyn_smem(wgrad3x3_wide_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  // Nine taps = five tap PAIRS (M = 256 = 2 taps x 128 cin), the last one (7, 8) with its tap-7 half dropped; TMEM holds four
  // pair accumulators, so two passes: 4 + 1 pairs.  The second pass re-reads x and g for a single MMA per 16 pixels (memory
  // bound); a balanced 3 + 2 split (FD_WGRAD_WIDE_SPLIT=32) was measured and is SLOWER: 140.1 vs 136.4 us at 60x60, 54.4 vs
  // 50.5 at 30x30, equal at 15x15 (batch 64) -- the first pass is not purely tensor-bound, so shortening it gains less than
  // the second MMA of pass 2 costs.
  static const bool split41 = [] { const char* e = getenv("FD_WGRAD_WIDE_SPLIT"); return !(e && atoi(e) == 32); }();
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 0) {
      p.npairs = split41 ? 4 : 3;
      for (int j = 0; j < 4; ++j) p.pair_t0[j] = 2 * j;
      p.drop_first = 0;
    } else {
      p.npairs = split41 ? 1 : 2;
      p.pair_t0[0] = split41 ? 7 : 6;
      p.pair_t0[1] = p.pair_t0[2] = p.pair_t0[3] = 7;
      p.drop_first = split41 ? 1 : 2;
      p.dbias[0] = p.dbias[1] = nullptr;        // counted in pass 1
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nprob * p.pairs_per_prob);
    cfg.blockDim = dim3(kThreadsGW);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = static_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, wgrad3x3_wide_kernel, maps, p);
    if (e != cudaSuccess) return static_cast<int>(e);
    count_launch();
  }
  return launch_status();
}
