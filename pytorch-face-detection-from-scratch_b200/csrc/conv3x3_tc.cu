// 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, as an implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM), fed and drained by TMA.
//
// Replaces aten::conv2d (+ leaky_relu / dropout2d / residual add) at models/PoolResnet.py:35-40 and
// models/Resnet.py:29-36 of the reference, and -- with dgrad-packed weights -- the input-gradient
// half of its backward.
//
// "Halo-tile" formulation (no im2col, no per-tap reloads):
//   * one CTA tile = R output rows x TW output columns of one image.  ONE TMA box
//     {64ch, Wp = TW+2, R+2, 1} starting at (w0-1, h0-1) lands the zero-padded input patch in shared
//     memory as (R+2)*Wp consecutive 128-byte rows (one pixel = 64 bf16 = one 128B-swizzle row; TMA
//     zero-fills out-of-bounds pixels = the convolution padding).
//   * GEMM row m = y*Wp + x.  For tap (ky,kx) the A operand of output row m is input row
//     m + ky*Wp + kx: the *same* shared-memory tile, read through a matrix descriptor whose start
//     address is shifted by (ky*Wp+kx)*128 bytes.  9 taps x 4 K-steps = 36 MMAs (M=128,N=64,K=16)
//     per 128-row block; weights (72 KB) stay in smem for the whole persistent CTA.
//   * rows with x >= TW or y >= R are junk and are never stored.
//   * the epilogue writes the output tile DENSE ([R][TW] pixels, 128B-swizzled) into a staging buffer
//     that one TMA tensor store moves to HBM; a residual operand is TMA-loaded into the same staging
//     buffer beforehand and updated in place.  No per-thread global loads/stores of activations: a
//     row-per-thread access pattern touches 32 different 128-byte lines per instruction and makes
//     the LSU, not the tensor pipe, the bottleneck.
//   * LeakyReLU' masks travel as sign bits (uint32[B,H,W,2]) instead of bf16 tensors.
//
// Measured on B200: a M=128,N=64,K=16 SS-mode MMA takes ~53 clk (6 KB of smem operands at 128 B/clk),
// not the 32 clk of the tensor-pipe floor, i.e. ~60 % of dense peak is the ceiling of this shape.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 =
// epilogue (TMEM lane quadrant = warp % 4, channel half = (warp - 2) / 4).  Input tiles and TMEM
// accumulators are double buffered: the epilogue of tile i overlaps the MMAs of tile i+1.
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kWBytes = 9 * kC * 128;  // 73728: [tap][cout][cin] bf16, K-major, 128B swizzle
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kStoreWarp = 2 + kEpiWarps;          // warp 0 = loads, 1 = MMA, 2..17 = epilogue, 18 = stores
constexpr int kThreads = (kStoreWarp + 1) * 32;    // 608
constexpr uint32_t kTmemCols = 512;

struct ConvParams {
  int B, H, W, R, TW, Wp, nblk, tiles_w, tiles_h, num_tiles;
  uint32_t in_bytes;      // bytes delivered by one input TMA box
  uint32_t in_buf_bytes;  // bytes reserved per input stage (multiple of 1024)
  uint32_t stg_bytes;     // bytes of one dense output / residual tile
  uint32_t stg_buf_bytes; // the same, rounded up to 1024
  uint32_t inv_wp;        // ceil(65536 / Wp): m / Wp == (m * inv_wp) >> 16 for m < 512
  int flags;
  int dbg;
  int has_res;            // a residual tile is TMA-loaded into the staging buffer
  int staged_out2;        // the staged (TMA-stored) value is the masked second output, not `out`
  int pool;               // fuse MaxPool2d(2): the staged tile is max-pooled in place, the POOLED tile is stored
  uint16_t* amax;         // pooled mode: 2-bit window positions of the maxima [B,H/2,W/2,8] uint16 (nullable)
  float slope;
  const float* bias;
  const float* chan_scale;
  const float* chan_scale2;
  const uint16_t* mask_in;      // sign-bit masks, addressed in 16-channel units: [pixel][4]
  uint16_t* mask_out;
  __nv_bfloat16* out2_direct;   // both outputs requested: the second one leaves through plain stores
};

// Optional per-tile timestamps of CTA 0 (FD_CONV_TIMING=1): [tile iteration][8] clock64 values.
__device__ unsigned long long g_conv_dbg[64 * 8];
#define FD_TS(slot) do { if (p.dbg && blockIdx.x == 0 && it < 64) g_conv_dbg[it * 8 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_res, const __grid_constant__ CUtensorMap tm_out,
                  const __grid_constant__ ConvParams p) {
  // tm_out: the tensor the staged tile is stored to -- in pooled mode the POOLED output [B,H/2,W/2,64], box {64,TW/2,R/2,1}
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  // Layout: weights | input stage 0 | input stage 1 | staging 0 | staging 1 | constants | barriers.
  // The junk rows of the last 128-row block read a few rows BEHIND their input stage (into the next
  // buffer): harmless, those GEMM rows are never stored, so the stages need no zero-filled tail.
  const uint32_t stage_bytes = p.in_buf_bytes;
  uint8_t* sW = smem;
  uint8_t* sIn = smem + kWBytes;
  uint8_t* sStg = sIn + ((2 * stage_bytes + 1023u) & ~1023u);     // staging rows are swizzled by (row & 7): 1024-aligned
  float* sConst = reinterpret_cast<float*>(sStg + 2 * p.stg_buf_bytes);   // [3][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sConst + 3 * kC);
  uint64_t* w_full = bars + 0;
  uint64_t* in_full = bars + 1;     // [2]
  uint64_t* in_empty = bars + 3;    // [2]
  uint64_t* acc_empty = bars + 5;   // [2]
  uint64_t* res_full = bars + 7;    // [2]
  uint64_t* stg_free = bars + 9;    // [2]
  uint64_t* stg_ready = bars + 11;  // [2]
  uint64_t* acc_full = bars + 13;   // [2][4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_out);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full + s, 1);
      mbar_init(in_empty + s, 1);
      mbar_init(acc_empty + s, 1);
      mbar_init(res_full + s, 1);
      mbar_init(stg_free + s, 1);
      mbar_init(stg_ready + s, 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(acc_full + i, 1);
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
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  pdl_trigger();
  pdl_wait();          // everything above overlapped the tail of the previous kernel

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_expect_tx(w_full, kWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * 8192, &tm_w, w_full, 0, t * kC);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int h0 = th * p.R, w0 = (rem - th * p.tiles_w) * p.TW;
        mbar_wait_sleep(in_empty + s, ph ^ 1);
        FD_TS(0);
        mbar_expect_tx(in_full + s, p.in_bytes);
        tma_load_4d(sIn + s * stage_bytes, &tm_in, in_full + s, 0, w0 - 1, h0 - 1, n);
        if (p.has_res) {
          mbar_wait_sleep(stg_free + s, ph ^ 1);        // the store of tile it-2 has drained this buffer
          mbar_expect_tx(res_full + s, p.stg_bytes);
          tma_load_4d(sStg + s * p.stg_buf_bytes, &tm_res, res_full + s, 0, w0, h0, n);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread runs the whole loop)
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
      const uint32_t w_lo = sdesc_lo(smem_u32(sW), 16);
      const uint32_t wp_units = static_cast<uint32_t>(p.Wp) * 8u;
      mbar_wait(w_full, 0);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(acc_empty + s, ph ^ 1);
        FD_TS(1);
        mbar_wait(in_full + s, ph);
        tc_fence_after();
        FD_TS(2);
        const uint32_t in_lo = sdesc_lo(smem_u32(sIn + s * stage_bytes), 16);
#pragma unroll 1
        for (int mb = 0; mb < p.nblk; ++mb) {
          if (p.flags & FD_CONV_1X1) {
            // 1x1 convolution (pointwise layers of the separable backbone on channel planes): only the centre tap --
            // A rows shifted by Wp + 1, B = tap 4 of the packed weights -- 4 MMAs instead of 36
            const uint32_t a_c = in_lo + static_cast<uint32_t>(mb * 1024) + wp_units + 8u, b_c = w_lo + 4u * 512u;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + static_cast<uint32_t>((s * 4 + mb) * kC), sdesc_sw128(a_c + 2 * k),
                        sdesc_sw128(b_c + 2 * k), idesc, k != 0 ? 1u : 0u);
          } else {
            issue_conv3x3_block(tmem_base + static_cast<uint32_t>((s * 4 + mb) * kC),
                                in_lo + static_cast<uint32_t>(mb * 1024), w_lo, wp_units, idesc, [](int) {});
          }
          umma_commit(acc_full + s * 4 + mb);   // this 128-row block is ready for the epilogue
        }
        umma_commit(in_empty + s);              // input tile free once these MMAs have read it
        FD_TS(3);
      }
    }
    __syncwarp();
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------------ TMA store issuer
    // A thread of its own, because draining a bulk store (wait_group.read) blocks the issuing thread.
    if (elect_one_sync()) {
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / tiles_per_img;
        const int rem = tile - n * tiles_per_img;
        const int th = rem / p.tiles_w;
        const int h0 = th * p.R, w0 = (rem - th * p.tiles_w) * p.TW;
        mbar_wait_sleep(stg_ready + s, ph);
        if (p.pool) tma_store_4d(&tm_out, sStg + s * p.stg_buf_bytes, 0, w0 >> 1, h0 >> 1, n);
        else tma_store_4d(&tm_out, sStg + s * p.stg_buf_bytes, 0, w0, h0, n);   // beyond the image: clipped
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(stg_free + s);
        FD_TS(7);
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (16)
    // One thread = one GEMM row (pixel) x 16 channels.  TMEM lane quadrant = warp % 4 (hardware rule),
    // channel quarter = (warp - 2) / 4.
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int c0 = cq * 16;
    const int et = threadIdx.x - 64;
    const uint64_t slope2 = pk2(p.slope, p.slope);
    const bool lrelu = (p.flags & FD_EPI_LRELU) != 0;
    const bool has_cs = p.chan_scale != nullptr, has_cs2 = p.chan_scale2 != nullptr;
    if (et < kC) sConst[et] = p.bias ? __ldg(p.bias + et) : 0.f;
    int it = 0, last_n = -1;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int n = tile / tiles_per_img;
      const int rem = tile - n * tiles_per_img;
      const int th = rem / p.tiles_w;
      const int h0 = th * p.R, w0 = (rem - th * p.tiles_w) * p.TW;
      if (n != last_n) {          // per-image Dropout2d multipliers (uniform branch: n depends on the tile only)
        bar_sync_epi();           // nobody still reads the previous image's values
        if (et >= kC && et < 2 * kC) sConst[et] = has_cs ? __ldg(p.chan_scale + n * kC + et - kC) : 1.f;
        else if (et >= 2 * kC && et < 3 * kC) sConst[et] = has_cs2 ? __ldg(p.chan_scale2 + n * kC + et - 2 * kC) : 1.f;
        bar_sync_epi();
        last_n = n;
      }
      uint8_t* stg = sStg + s * p.stg_buf_bytes;
      // staging buffer: either the residual tile has landed in it, or the store of tile it-2 has drained it
      if (p.has_res) mbar_wait_sleep(res_full + s, ph);
      else mbar_wait_sleep(stg_free + s, ph ^ 1);
      if (et == 0) FD_TS(4);
#pragma unroll 1
      for (int mb = 0; mb < p.nblk; ++mb) {
        const int m = mb * 128 + q * 32 + lane;
        const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
        const int x = m - y * p.Wp;
        const int oy = h0 + y, ox = w0 + x;
        const bool valid = (y < p.R) && (x < p.TW) && (oy < p.H) && (ox < p.W);
        const size_t pix = (static_cast<size_t>(n) * p.H + oy) * p.W + ox;
        uint32_t mbits = 0xffffu;
        if (valid && p.mask_in) mbits = __ldg(p.mask_in + pix * 4 + cq);     // latency hidden by the MMA wait
        const uint32_t d = static_cast<uint32_t>(y * p.TW + x);              // row of the dense staging tile
        uint8_t* row = stg + d * 128u;
        const uint32_t ch0 = ((static_cast<uint32_t>(cq) * 2u) ^ (d & 7u)) << 4;
        const uint32_t ch1 = ((static_cast<uint32_t>(cq) * 2u + 1u) ^ (d & 7u)) << 4;
        mbar_wait_sleep(acc_full + s * 4 + mb, ph, 1000);
        tc_fence_after();
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>((s * 4 + mb) * kC + c0),
                           acc);
        tmem_ld_wait();
        if (valid) {
          uint64_t v2[8];
          epi_bias_act16(acc, sConst + c0, sConst + kC + c0, lrelu, has_cs, slope2, v2);
          if (p.mask_out) p.mask_out[pix * 4 + cq] = static_cast<uint16_t>(epi_sign_bits16(v2));
          if (p.has_res)
            epi_add_bf16x16(v2, *reinterpret_cast<const uint4*>(row + ch0), *reinterpret_cast<const uint4*>(row + ch1));
          uint4 u0, u1;
          if (!p.staged_out2) {
            epi_pack16(v2, u0, u1);
            *reinterpret_cast<uint4*>(row + ch0) = u0;
            *reinterpret_cast<uint4*>(row + ch1) = u1;
          }
          if (p.staged_out2 || p.out2_direct) {
            uint64_t o2[8];
            epi_masked16(v2, mbits, p.slope, sConst + 2 * kC + c0, has_cs2, o2);
            epi_pack16(o2, u0, u1);
            if (p.staged_out2) {
              *reinterpret_cast<uint4*>(row + ch0) = u0;
              *reinterpret_cast<uint4*>(row + ch1) = u1;
            } else {
              uint4* gp = reinterpret_cast<uint4*>(p.out2_direct + pix * kC + c0);
              gp[0] = u0;
              gp[1] = u1;
            }
          }
        }
      }
      if (et == 0) FD_TS(5);
      if (p.pool) {
        // MaxPool2d(2) of the staged tile (models/PoolResnet.py:41-42), in place: every warp max-pools its share of the
        // (R/2) x (TW/2) pooled pixels into registers (lane = channel pair), and only after a barrier -- all reads done --
        // writes them over rows [0, npool) of the same buffer.  The un-pooled sum never reaches HBM; the positions of the
        // maxima (first maximum in (dy,dx) order, like ATen) go out as 2 bits per channel for the backward pass.
        tc_fence_before();
        bar_sync_epi();                       // the un-pooled tile is complete, every TMEM read of this tile is done
        if (et == 0) mbar_arrive(acc_empty + s);
        const int PW = p.TW >> 1, npool = (p.R >> 1) * PW;
        const int wk = warp - 2;
        const uint32_t chunk = static_cast<uint32_t>(lane) >> 2, inb = (static_cast<uint32_t>(lane) & 3u) * 4u;
        uint32_t pooled[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int i = wk + j * kEpiWarps;
          pooled[j] = 0;
          if (i < npool) {
            const int py = i / PW, px = i - py * PW;
            const uint32_t d00 = static_cast<uint32_t>((2 * py) * p.TW + 2 * px);
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t dr = d00 + static_cast<uint32_t>((k >> 1) * p.TW + (k & 1));
              v[k] = *reinterpret_cast<const uint32_t*>(stg + dr * 128u + (((chunk ^ (dr & 7u)) << 4) | inb));
            }
            float lo = bf16lo(v[0]), hi = bf16hi(v[0]);
            uint32_t alo = 0, ahi = 0;
#pragma unroll
            for (int k = 1; k < 4; ++k) {
              const float l2 = bf16lo(v[k]), h2 = bf16hi(v[k]);
              if (l2 > lo) { lo = l2; alo = k; }
              if (h2 > hi) { hi = h2; ahi = k; }
            }
            pooled[j] = pack_bf16x2(lo, hi);
            if (p.amax) {
              uint32_t a = (alo | (ahi << 2)) << (4u * (static_cast<uint32_t>(lane) & 3u));   // channels 2*(lane&3), +1 of the group
              a |= __shfl_xor_sync(0xffffffffu, a, 1);
              a |= __shfl_xor_sync(0xffffffffu, a, 2);
              const int oy = (h0 >> 1) + py, ox = (w0 >> 1) + px;
              if ((lane & 3) == 0 && oy < (p.H >> 1) && ox < (p.W >> 1))
                p.amax[((static_cast<size_t>(n) * (p.H >> 1) + oy) * (p.W >> 1) + ox) * 8 + chunk] = static_cast<uint16_t>(a);
            }
          }
        }
        bar_sync_epi();                       // every read of the un-pooled tile is done: overwrite it
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t i = static_cast<uint32_t>(wk + j * kEpiWarps);
          if (static_cast<int>(i) < npool)
            *reinterpret_cast<uint32_t*>(stg + i * 128u + (((chunk ^ (i & 7u)) << 4) | inb)) = pooled[j];
        }
      }
      fence_proxy_async();       // staging writes (generic proxy) -> visible to the TMA store
      tc_fence_before();
      bar_sync_epi();
      if (et == 0) {
        FD_TS(6);
        if (!p.pool) mbar_arrive(acc_empty + s);     // TMEM of this tile has been read by everyone
        mbar_arrive(stg_ready + s);     // staging tile complete: the store warp takes over
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// Exactly the TMA box.  The 128B swizzle of TMA and of the UMMA descriptors is a function of the absolute
// shared-memory address bits, so a stage only needs 128-byte (not 1024-byte) alignment.
inline size_t in_buf_bytes_for(int R, int Wp) { return static_cast<size_t>((R + 2) * Wp) * 128; }
inline size_t stg_buf_bytes_for(int R, int TW) { return (static_cast<size_t>(R) * TW * 128 + 1023) / 1024 * 1024; }
inline size_t smem_for(int nblk, int Wp, int R, int TW) {
  const size_t in_buf = in_buf_bytes_for(R, Wp);
  const size_t bufs = (2 * in_buf + 1023) / 1024 * 1024 + 2 * stg_buf_bytes_for(R, TW);
  // the last 128-row block of stage 1 reads up to row nblk*128 - 1 + 2*Wp + 2 of its stage: the allocation
  // must cover that (normally it ends inside the staging buffers)
  const size_t reach = in_buf + static_cast<size_t>(nblk * 128 + 2 * Wp + 2) * 128;
  const size_t body = bufs + 3 * kC * 4 + 256;
  return kWBytes + (body > reach ? body : reach) + 1024;
}

}  // namespace
}  // namespace fd

extern "C" FD_API int fd_debug_conv_timing(unsigned long long* out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fd::g_conv_dbg, sizeof(unsigned long long) * n));
}

static int conv3x3_impl(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                        float slope, const float* chan_scale, const fd_bf16* residual, uint32_t* mask_out,
                        fd_bf16* out, const uint32_t* mask_in, const float* chan_scale2, fd_bf16* out2,
                        int flags, fd_bf16* pooled, uint16_t* argmax, void* stream);

extern "C" int fd_conv3x3(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                          float slope, const float* chan_scale, const fd_bf16* residual, uint32_t* mask_out,
                          fd_bf16* out, const uint32_t* mask_in, const float* chan_scale2, fd_bf16* out2,
                          int flags, void* stream) {
  return conv3x3_impl(x, w_packed, B, H, W, C, bias, slope, chan_scale, residual, mask_out, out, mask_in, chan_scale2, out2,
                      flags, nullptr, nullptr, stream);
}

extern "C" int fd_conv3x3_pool(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                               float slope, const float* chan_scale, const fd_bf16* residual, uint32_t* mask_out,
                               fd_bf16* pooled, uint16_t* argmax, int flags, void* stream) {
  if (!pooled) return FD_EINVAL;
  if ((H | W) & 1) return FD_EUNSUPPORTED;       // floor-mode pooling of odd maps: fd_conv3x3 + fd_maxpool2x2_fwd
  return conv3x3_impl(x, w_packed, B, H, W, C, bias, slope, chan_scale, residual, mask_out, nullptr, nullptr, nullptr,
                      nullptr, flags, pooled, argmax, stream);
}

static int conv3x3_impl(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                        float slope, const float* chan_scale, const fd_bf16* residual, uint32_t* mask_out,
                        fd_bf16* out, const uint32_t* mask_in, const float* chan_scale2, fd_bf16* out2,
                        int flags, fd_bf16* pooled, uint16_t* argmax, void* stream) {
  using namespace fd;
  if (!x || !w_packed || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != kC) return FD_EUNSUPPORTED;
  if (!out && !out2 && !pooled) return FD_EINVAL;
  const bool pool = pooled != nullptr;
  if (mask_in && !out2) return FD_EINVAL;
  if (!(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;   // LeakyReLU is evaluated as max(v, slope*v)
  {
    // CTA-PAIR instantiation (conv3x3_wide.cu with 64 outputs: tcgen05.mma.cta_group::2, M = 256, the B operand split between
    // the two shared memories -- 5 KB instead of 6 KB of operand reads per MMA): one staged output, no fused pooling.
    // FD_CONV_PAIR=0 keeps this kernel (A/B runs); maps with fewer than two waves of tiles stay here as well.
    static const int pair_env = [] { const char* e = getenv("FD_CONV_PAIR"); return e ? atoi(e) : 0; }();
    if (pair_env && !pool && !(out && out2) && static_cast<long>(B) * H * W >= pair_env) {
      const fd_bf16* xs[1] = {x};
      const float* cs[1] = {chan_scale};
      const float* cs2[1] = {chan_scale2};
      const fd_bf16* rs[1] = {residual};
      uint32_t* mo[1] = {mask_out};
      const uint32_t* mi[1] = {mask_in};
      fd_bf16* o1[1] = {out};
      fd_bf16* o2[1] = {out2};
      const int rc = conv3x3_pairs(64, xs, 1, w_packed, B, H, W, bias, slope, chan_scale ? cs : nullptr, residual ? rs : nullptr,
                                   mask_out ? mo : nullptr, out ? o1 : nullptr, mask_in ? mi : nullptr,
                                   chan_scale2 ? cs2 : nullptr, out2 ? o2 : nullptr, flags & ~FD_CONV_ONE_TAP, stream);
      if (rc != FD_EUNSUPPORTED) return rc;
    }
  }
  const int nsm = sm_count();
  const size_t smem_cap = 227 * 1024;

  // Tiling: TW <= 62 output columns (Wp = TW + 2 <= 64 GEMM rows per image row), R rows with R*Wp <= 512.
  // Pick the (column tiles, R) pair that minimises a simple cycle model: MMA time of the padded 128-row
  // blocks (smem-operand bound, ~1900 clk each) times the number of waves over the SMs.
  int bestR = 0, bestTW = 0;
  double best = 1e30;
  const int min_tw_tiles = (W + 61) / 62;
  for (int tw_tiles = min_tw_tiles; tw_tiles <= min_tw_tiles + 1; ++tw_tiles) {
    int TW = (W + tw_tiles - 1) / tw_tiles;
    if (pool) TW = (TW + 1) & ~1;                  // pooled tiles start on even rows / columns
    if (TW > 62) continue;
    const int Wp = TW + 2;
    for (int R = pool ? 2 : 1; R <= (pool ? H + 1 : H) && R + 2 <= 256; R += pool ? 2 : 1) {
      const int nblk = (R * Wp + 127) / 128;
      if (nblk > 4) break;
      if (smem_for(nblk, Wp, R, TW) > smem_cap) break;
      const long tiles = static_cast<long>(B) * ((H + R - 1) / R) * tw_tiles;
      const long waves = (tiles + nsm - 1) / nsm;
      const double cost = waves * (1900.0 * nblk + 400.0);
      if (cost < best) { best = cost; bestR = R; bestTW = TW; }
    }
  }
  if (bestR == 0) return FD_EUNSUPPORTED;

  ConvParams p;
  p.B = B; p.H = H; p.W = W; p.R = bestR; p.TW = bestTW; p.Wp = bestTW + 2;
  p.nblk = (bestR * p.Wp + 127) / 128;
  p.tiles_w = (W + bestTW - 1) / bestTW;
  p.tiles_h = (H + bestR - 1) / bestR;
  p.num_tiles = B * p.tiles_w * p.tiles_h;
  p.in_bytes = static_cast<uint32_t>((bestR + 2) * p.Wp * 128);
  p.in_buf_bytes = static_cast<uint32_t>(in_buf_bytes_for(bestR, p.Wp));
  p.stg_bytes = static_cast<uint32_t>(bestR * bestTW * 128);
  p.stg_buf_bytes = static_cast<uint32_t>(stg_buf_bytes_for(bestR, bestTW));
  p.inv_wp = static_cast<uint32_t>((65536 + p.Wp - 1) / p.Wp);
  p.flags = flags;
  { const char* d = getenv("FD_CONV_TIMING"); p.dbg = d ? atoi(d) : 0; }
  p.slope = slope;
  p.bias = bias; p.chan_scale = chan_scale; p.chan_scale2 = chan_scale2;
  p.mask_in = reinterpret_cast<const uint16_t*>(mask_in); p.mask_out = reinterpret_cast<uint16_t*>(mask_out);
  p.has_res = residual != nullptr;
  p.staged_out2 = (out == nullptr) && !pool;
  p.out2_direct = (out && out2) ? reinterpret_cast<__nv_bfloat16*>(out2) : nullptr;
  p.pool = pool ? 1 : 0;
  p.amax = argmax;
  const size_t smem = smem_for(p.nblk, p.Wp, bestR, bestTW);

  CUtensorMap tm_in, tm_w, tm_res, tm_out;
  int rc = make_tmap_nhwc_bf16(&tm_in, x, B, H, W, C, p.Wp, bestR + 2);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_w, w_packed, 9 * kC, kC, kC, kC);
  if (rc != FD_OK) return rc;
  const fd_bf16* staged = out ? out : out2;
  rc = pool ? make_tmap_nhwc_bf16(&tm_out, pooled, B, H / 2, W / 2, C, bestTW / 2, bestR / 2)
            : make_tmap_nhwc_bf16(&tm_out, staged, B, H, W, C, bestTW, bestR);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_res, residual ? residual : x, B, H, W, C, bestTW, bestR);
  if (rc != FD_OK) return rc;

  cudaError_t e = set_max_dyn_smem(conv3x3_tc_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = p.num_tiles < nsm ? p.num_tiles : nsm;
  e = launch_k(conv3x3_tc_kernel, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), tm_in, tm_w, tm_res,
               tm_out, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}
