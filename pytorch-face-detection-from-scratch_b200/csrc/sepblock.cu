// Depthwise-separable residual block of the reference's SeparableCNN (models/SeparableCNN.py:10-51), forward,
// 64 channels, as ONE kernel per block:
//
//     t1 = lrelu(pw1(x))          1x1 conv, no bias   -> tcgen05 GEMM [pixels,64] x [64,64]
//     t2 = lrelu(dw3x3(t1))       depthwise, pad 1    -> CUDA cores, fp32 accumulate, from shared memory
//     y  = pw2(t2) + x            1x1 conv, no bias   -> tcgen05 GEMM, skip added in the epilogue
//
// (Dropout2d is the identity in eval mode; the MaxPool2d(2) that follows while H > 16 is fd_maxpool2x2_fwd.)
// The reference runs this as 3 cuDNN convolutions + 4 elementwise kernels with every intermediate going through
// HBM in fp32; the block is memory-bound (AI = 17.5 kFLOP / 256 B per pixel = 68 FLOP/B, ridge 219), so the
// intermediates t1 / t2 never leave shared memory here: HBM traffic = read x once (+ halo), write y once.
//
// Tile = R x TW output pixels of one image.  One TMA box {64ch, Wp = TW+2, R+2} lands the zero-padded halo patch of
// x as (R+2)*Wp rows of 128 B (pixel = 64 bf16 = one 128B-swizzle row = the K-major A operand of pw1).  pw1 is
// evaluated on ALL halo pixels (pw1 has no bias, so out-of-image pixels give lrelu(0) = 0 = the zero padding the
// depthwise conv expects).  The depthwise stage reads t1 with lane = channel pair (a warp reads one 128-B pixel row:
// conflict free) and a 3-column sliding window in registers, and writes t2 in the swizzled K-major layout that is
// the A operand of pw2 (GEMM row = y*Wp + x; rows with x >= TW are junk and never stored).  The output tile is
// staged densely in shared memory and leaves through one TMA tensor store.
//
// The phases of a tile are serial inside a CTA (pw1 -> epilogue -> depthwise -> pw2 -> epilogue); TWO CTAs per SM
// (<= 110 KB of shared memory, <= 256 TMEM columns each) overlap each other's phases and keep HBM busy.  The next
// tile's halo patch is prefetched into the second input stage while the current tile is processed.
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kWorkWarps = 16;
constexpr int kThreads = (1 + kWorkWarps) * 32;      // warp 0: TMA + MMA issue; warps 1..16: epilogues + depthwise
constexpr uint32_t kT1Pitch = 144;                   // bytes between t1 pixel rows (128 + 16: no swizzle needed)
constexpr uint32_t kConstBytes = 3072;               // depthwise weights [9][64] fp32 + mbarriers + TMEM slot

struct SepParams {
  int B, H, W, R, TW, Wp, nblk1, nblk3, tiles_w, tiles_h, num_tiles, nseg;
  uint32_t in_bytes;        // bytes of one halo TMA box
  uint32_t in_buf_bytes;    // per input stage (multiple of 1024)
  uint32_t t1_bytes;        // (R+2)*Wp rows, rounded to 1 KB
  uint32_t t2_bytes;        // R*Wp rows, rounded to 1 KB (also the dense output staging tile)
  uint32_t t1_rows;
  uint32_t inv_wp;          // ceil(65536 / Wp)
  uint32_t tmem_cols;
  int dbg;
  int pool;                 // fuse MaxPool2d(2): `out` is [B,H/2,W/2,64]
  float slope;
  const float* dw;          // [9][64] fp32, tap-major
};

// Optional per-tile timestamps of CTA 0 (FD_SEP_TIMING=1): [tile iteration][8] clock64 values.
__device__ unsigned long long g_sep_dbg[64 * 8];
#define SEP_TS(slot) do { if (p.dbg && blockIdx.x == 0 && it < 64) g_sep_dbg[it * 8 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__global__ void __launch_bounds__(kThreads, 2)
sepblock_fwd_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
                    const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out,
                    const __grid_constant__ SepParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  // layout (every region 1024-aligned): W1 8K | W2 8K | dw weights + barriers 3K | x stage 0 | x stage 1 | t2 / staging | t1
  // The last 128-row block of a GEMM reads A rows beyond its tile (into the region that follows): harmless, those
  // GEMM rows are junk and never stored; the host sizes the allocation so that the reads stay inside it.
  uint8_t* sW1 = smem;
  uint8_t* sW2 = smem + 8192;
  float* sDw = reinterpret_cast<float*>(smem + 16384);              // [9][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDw + 9 * kC);
  uint8_t* sX = smem + 16384 + kConstBytes;
  uint8_t* sT2 = sX + 2 * p.in_buf_bytes;
  uint8_t* sT1 = sT2 + p.t2_bytes;
  uint64_t* w_full = bars + 0;
  uint64_t* x_full = bars + 1;      // [2]
  uint64_t* acc1_full = bars + 3;
  uint64_t* acc2_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w1);
    tma_prefetch_desc(&tm_w2);
    tma_prefetch_desc(&tm_out);
    mbar_init(w_full, 1);
    mbar_init(x_full + 0, 1);
    mbar_init(x_full + 1, 1);
    mbar_init(acc1_full, 1);
    mbar_init(acc2_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  pdl_trigger();
  pdl_wait();

  for (int i = threadIdx.x; i < 9 * kC; i += kThreads) sDw[i] = __ldg(p.dw + i);

  auto tile_coords = [&](int tile, int& n, int& h0, int& w0) {
    n = tile / tiles_per_img;
    const int rem = tile - n * tiles_per_img;
    const int th = rem / p.tiles_w;
    h0 = th * p.R;
    w0 = (rem - th * p.tiles_w) * p.TW;
  };
  // Every single-thread region branches on elect.sync directly: ptxas then knows it is executed by one thread, keeps
  // descriptors in uniform registers and emits each tcgen05.mma / TMA as ONE instruction (a `lane == 0` or a saved
  // bool costs a divergence loop per instruction: measured 1200 clk for the 8 MMAs of phase 0).
  if (warp == 0 && elect_one_sync()) {
    mbar_expect_tx(w_full, 16384);
    tma_load_2d(sW1, &tm_w1, w_full, 0, 0);
    tma_load_2d(sW2, &tm_w2, w_full, 0, 0);
    if (static_cast<int>(blockIdx.x) < p.num_tiles) {
      int n, h0, w0;
      tile_coords(blockIdx.x, n, h0, w0);
      mbar_expect_tx(x_full + 0, p.in_bytes);
      tma_load_4d(sX, &tm_x, x_full + 0, 0, w0 - 1, h0 - 1, n);
    }
  }
  __syncthreads();          // depthwise weights visible

  constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
  const int wk = warp - 1;                 // 0..15 for the working warps
  const int q = warp & 3;                  // TMEM lane quadrant of this warp (hardware rule: warp % 4)
  const int cq = wk >> 2;                  // 16-channel quarter handled in the epilogues
  const float slope = p.slope;
  const uint64_t slope2 = pk2(slope, slope);

  int it = 0;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    int n, h0, w0;
    tile_coords(tile, n, h0, w0);
    uint8_t* xs = sX + s * p.in_buf_bytes;

    // ------------------------------------------------------------------ phase 0: pw1 on every halo pixel
    if (warp == 0) {
      if (elect_one_sync()) {
        // the previous tile's store has finished reading its source (the t2 / staging tile, or -- pooled -- stage s^1)
        SEP_TS(0);
        tma_store_wait_read<0>();
        const int next = tile + gridDim.x;
        if (next < p.num_tiles) {          // stage s^1 was last read by the skip-add of tile it-1 (behind a CTA barrier)
          int nn, nh0, nw0;
          tile_coords(next, nn, nh0, nw0);
          mbar_expect_tx(x_full + (s ^ 1), p.in_bytes);
          tma_load_4d(sX + (s ^ 1) * p.in_buf_bytes, &tm_x, x_full + (s ^ 1), 0, nw0 - 1, nh0 - 1, nn);
        }
        if (it == 0) mbar_wait(w_full, 0);
        mbar_wait(x_full + s, (it >> 1) & 1);
        SEP_TS(1);
        tc_fence_after();
        const uint32_t a_lo = sdesc_lo(smem_u32(xs), 16), b_lo = sdesc_lo(smem_u32(sW1), 16);
#pragma unroll 1
        for (int mb = 0; mb < p.nblk1; ++mb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + static_cast<uint32_t>(mb * kC), sdesc_sw128(a_lo + mb * 1024 + 2 * k),
                      sdesc_sw128(b_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(acc1_full);
        SEP_TS(2);
        // ONE thread polls the MMA-completion barrier; the 16 working warps sleep in the hardware CTA barrier below.
        // (512 threads polling try_wait starve this warp of issue slots -- measured 1000 clk to issue 8 MMAs -- and a
        // suspend-time hint costs microseconds per wait.)
        mbar_wait(acc1_full, it & 1);
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp != 0) {
      if (threadIdx.x == 32) SEP_TS(3);
#pragma unroll 1
      for (int mb = 0; mb < p.nblk1; ++mb) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mb * kC + cq * 16), acc);
        tmem_ld_wait();
        uint32_t u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float a, b, ta, tb;
          const uint64_t v = pk2u(acc[2 * i], acc[2 * i + 1]);
          upk2(v, a, b);
          upk2(mul2(v, slope2), ta, tb);
          u[i] = pack_bf16x2(fmaxf(a, ta), fmaxf(b, tb));
        }
        const uint32_t r = static_cast<uint32_t>(mb * 128 + q * 32 + lane);
        if (r < p.t1_rows) {             // t1 rows are kT1Pitch = 144 B apart: conflict-free here AND for the row-wide reads below
          uint8_t* row = sT1 + r * kT1Pitch + cq * 32u;
          *reinterpret_cast<uint4*>(row) = make_uint4(u[0], u[1], u[2], u[3]);
          *reinterpret_cast<uint4*>(row + 16) = make_uint4(u[4], u[5], u[6], u[7]);
        }
        __syncwarp();
      }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ------------------------------------------------------------------ phase 1: depthwise 3x3 + LeakyReLU, t1 -> t2
    if (threadIdx.x == 32) SEP_TS(4);
    if (warp != 0) {
      uint64_t w9[9];             // packed fp32x2 weights of this lane's channel pair: FFMA2 does both channels at once
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float2 w = *reinterpret_cast<const float2*>(sDw + t * kC + 2 * lane);
        w9[t] = pk2(w.x, w.y);
      }
      // One item = 8 consecutive output pixels of one row; lane = channel pair.  The 10 x 3 input columns are loaded
      // with immediate offsets (no address arithmetic, no swizzle: pitch-144 rows) and the 3x3 window rotates through
      // registers by full unrolling.  Columns beyond the tile are junk reads; their results are not stored.
      const int items = p.R * p.nseg;
      for (int item = wk; item < items; item += kWorkWarps) {
        const int y = item / p.nseg;
        const int x0 = (item - y * p.nseg) * 8;
        const uint8_t* src = sT1 + static_cast<uint32_t>(y * p.Wp + x0) * kT1Pitch + 4u * lane;
        const uint32_t wp_b = static_cast<uint32_t>(p.Wp) * kT1Pitch;
        const uint32_t row0 = static_cast<uint32_t>(y * p.Wp + x0);
        uint64_t c[3][3];
        auto load_col = [&](int j, int slot) {
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + ky * wp_b + j * kT1Pitch);
            c[ky][slot] = pk2u(v << 16, v & 0xFFFF0000u);
          }
        };
        load_col(0, 0);
        load_col(1, 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int s0 = j % 3, s1 = (j + 1) % 3, s2 = (j + 2) % 3;
          load_col(j + 2, s2);
          // three independent accumulation chains (one per kernel row), summed at the end
          uint64_t a0 = mul2(w9[0], c[0][s0]), a1 = mul2(w9[3], c[1][s0]), a2 = mul2(w9[6], c[2][s0]);
          a0 = fma2(w9[1], c[0][s1], a0); a1 = fma2(w9[4], c[1][s1], a1); a2 = fma2(w9[7], c[2][s1], a2);
          a0 = fma2(w9[2], c[0][s2], a0); a1 = fma2(w9[5], c[1][s2], a1); a2 = fma2(w9[8], c[2][s2], a2);
          const uint64_t v = add2(add2(a0, a1), a2);
          float ax, ay, tx, ty;
          upk2(v, ax, ay);
          upk2(mul2(v, slope2), tx, ty);
          if (x0 + j < p.TW) {
            const uint32_t o = (row0 + j) * 128u + 4u * lane;
            *reinterpret_cast<uint32_t*>(sT2 + swz(o)) = pack_bf16x2(fmaxf(ax, tx), fmaxf(ay, ty));
          }
        }
      }
      fence_proxy_async();        // t2 (generic-proxy writes) -> visible to the tensor core's operand reads
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ------------------------------------------------------------------ phase 2: pw2 + skip -> dense staging tile
    if (threadIdx.x == 32) SEP_TS(5);
    if (warp == 0) {
      if (elect_one_sync()) {
        const uint32_t a_lo = sdesc_lo(smem_u32(sT2), 16), b_lo = sdesc_lo(smem_u32(sW2), 16);
#pragma unroll 1
        for (int mb = 0; mb < p.nblk3; ++mb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + static_cast<uint32_t>((p.nblk1 + mb) * kC), sdesc_sw128(a_lo + mb * 1024 + 2 * k),
                      sdesc_sw128(b_lo + 2 * k), idesc, k > 0 ? 1u : 0u);
        }
        umma_commit(acc2_full);
        mbar_wait(acc2_full, it & 1);            // also: the MMAs have finished reading t2, it may become the staging tile
      }
      __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp != 0) {
      if (threadIdx.x == 32) SEP_TS(6);
      mbar_wait(x_full + s, (it >> 1) & 1);         // (completed long ago) acquire the TMA-written halo patch for the skip reads
      // all accumulator blocks into registers first: the staging tile aliases t2 rows of OTHER threads' blocks only
      // after every MMA has completed (acc2_full), so writes may start right away
#pragma unroll 1
      for (int mb = 0; mb < p.nblk3; ++mb) {
        const int m = mb * 128 + q * 32 + lane;
        const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
        const int x = m - y * p.Wp;
        const bool valid = (y < p.R) && (x < p.TW);
        uint32_t acc[16];
        __syncwarp();
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>((p.nblk1 + mb) * kC + cq * 16), acc);
        tmem_ld_wait();
        if (valid) {
          const uint32_t xrow = static_cast<uint32_t>(m + p.Wp + 1) * 128u;      // centre pixel of the halo patch
          const uint4 s0 = *reinterpret_cast<const uint4*>(xs + swz(xrow + cq * 32u));
          const uint4 s1 = *reinterpret_cast<const uint4*>(xs + swz(xrow + cq * 32u + 16u));
          uint64_t v2[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v2[i] = pk2u(acc[2 * i], acc[2 * i + 1]);
          epi_add_bf16x16(v2, s0, s1);
          uint4 u0, u1;
          epi_pack16(v2, u0, u1);
          const uint32_t drow = static_cast<uint32_t>(y * p.TW + x) * 128u;
          *reinterpret_cast<uint4*>(sT2 + swz(drow + cq * 32u)) = u0;
          *reinterpret_cast<uint4*>(sT2 + swz(drow + cq * 32u + 16u)) = u1;
        }
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 32) SEP_TS(7);
    if (!p.pool) {
      if (warp == 0) {
        if (elect_one_sync()) {
          tma_store_4d(&tm_out, sT2, 0, w0, h0, n);        // beyond the image: clipped
          tma_store_commit();
        }
        __syncwarp();
      }
    } else {
      // ---------------------------------------------------------------- phase 3: MaxPool2d(2) of the staged tile
      // (models/SeparableCNN.py:49-50).  Pooled pixels go, densely and swizzled, into input stage s (dead: its last
      // readers were the skip reads above); lane = channel pair, one pooled pixel per warp iteration.
      if (warp != 0) {
        const int PW = p.TW >> 1, npool = (p.R >> 1) * PW;
        for (int i = wk; i < npool; i += kWorkWarps) {
          const int py = i / PW, px = i - py * PW;
          const uint32_t r00 = static_cast<uint32_t>((2 * py) * p.TW + 2 * px) * 128u + 4u * lane;
          const uint32_t r10 = r00 + static_cast<uint32_t>(p.TW) * 128u;
          const uint32_t a = *reinterpret_cast<const uint32_t*>(sT2 + swz(r00));
          const uint32_t b = *reinterpret_cast<const uint32_t*>(sT2 + swz(r00 + 128u));
          const uint32_t c = *reinterpret_cast<const uint32_t*>(sT2 + swz(r10));
          const uint32_t d = *reinterpret_cast<const uint32_t*>(sT2 + swz(r10 + 128u));
          const float lo = fmaxf(fmaxf(bf16lo(a), bf16lo(b)), fmaxf(bf16lo(c), bf16lo(d)));
          const float hi = fmaxf(fmaxf(bf16hi(a), bf16hi(b)), fmaxf(bf16hi(c), bf16hi(d)));
          *reinterpret_cast<uint32_t*>(xs + swz(static_cast<uint32_t>(i) * 128u + 4u * lane)) = pack_bf16x2(lo, hi);
        }
        fence_proxy_async();
      }
      __syncthreads();
      if (warp == 0) {
        if (elect_one_sync()) {
          tma_store_4d(&tm_out, xs, 0, w0 >> 1, h0 >> 1, n);
          tma_store_commit();
        }
        __syncwarp();
      }
    }
  }
  if (warp == 0) {
    if (elect_one_sync()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// pointwise weights fp32 [L][64][64] -> bf16 (same [cout][cin] order = the K-major B operand); depthwise weights
// fp32 [L][64][3][3] -> [L][9][64] (tap-major, so that a warp's lanes read consecutive channels)
__global__ void sep_pack_kernel(const float* __restrict__ pw, long n_pw, __nv_bfloat16* __restrict__ pw_out,
                                const float* __restrict__ dw, int n_dw_layers, float* __restrict__ dw_out) {
  pdl_trigger();
  pdl_wait();
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_pw; i += stride)
    pw_out[i] = __float2bfloat16_rn(pw[i]);
  const long n_dw = static_cast<long>(n_dw_layers) * 9 * kC;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n_dw; i += stride) {
    const long l = i / (9 * kC);
    const int r = static_cast<int>(i - l * 9 * kC);
    const int t = r / kC, c = r - t * kC;
    dw_out[i] = dw[(l * kC + c) * 9 + t];
  }
}

inline uint32_t round1k(size_t v) { return static_cast<uint32_t>((v + 1023) / 1024 * 1024); }

// dynamic shared memory of a tile shape (incl. 1 KB alignment slack); covers the junk-row over-reads of both GEMMs
// t1: (R+2)*Wp rows of kT1Pitch bytes + slack for the junk-column reads of the last 8-pixel segment (up to 8 rows past)
inline uint32_t t1_bytes_for(int R, int Wp) { return round1k(static_cast<size_t>((R + 2) * Wp + 10) * kT1Pitch); }

inline size_t smem_for(int R, int Wp) {
  const int nblk1 = ((R + 2) * Wp + 127) / 128, nblk3 = (R * Wp + 127) / 128;
  const size_t in_buf = round1k(static_cast<size_t>(R + 2) * Wp * 128);
  const size_t t2 = round1k(static_cast<size_t>(R) * Wp * 128), t1 = t1_bytes_for(R, Wp);
  const size_t x0 = 16384 + kConstBytes;
  size_t end = x0 + 2 * in_buf + t2 + t1;
  const size_t reach1 = x0 + in_buf + static_cast<size_t>(nblk1) * 16384;      // pw1 over stage 1
  const size_t reach3 = x0 + 2 * in_buf + static_cast<size_t>(nblk3) * 16384;  // pw2 over t2
  if (reach1 > end) end = reach1;
  if (reach3 > end) end = reach3;
  return end + 1024;
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" FD_API int fd_debug_sep_timing(unsigned long long* out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fd::g_sep_dbg, sizeof(unsigned long long) * n));
}

extern "C" int fd_sep_pack(const float* pw, long n_pw, fd_bf16* pw_out, const float* dw, int n_dw_layers, float* dw_out,
                           void* stream) {
  if ((n_pw > 0 && (!pw || !pw_out)) || (n_dw_layers > 0 && (!dw || !dw_out)) || (n_pw <= 0 && n_dw_layers <= 0))
    return FD_EINVAL;
  launch_k(sep_pack_kernel, dim3(64), dim3(256), 0, static_cast<cudaStream_t>(stream), pw, n_pw,
           reinterpret_cast<__nv_bfloat16*>(pw_out), dw, n_dw_layers, dw_out);
  count_launch();
  return launch_status();
}

extern "C" int fd_sepblock_fwd(const fd_bf16* x, const fd_bf16* w_pw1, const float* w_dw, const fd_bf16* w_pw2, int B,
                               int H, int W, int C, float slope, int pool, fd_bf16* out, void* stream) {
  if (!x || !w_pw1 || !w_dw || !w_pw2 || !out || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != kC) return FD_EUNSUPPORTED;
  if (!(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;
  if (pool && (H < 2 || W < 2)) return FD_EINVAL;
  const int nsm = sm_count();
  const size_t smem_cap = 113 * 1024;              // two CTAs per SM
  int bestR = 0, bestTW = 0;
  double best = 1e30;
  const int min_tw_tiles = (W + 61) / 62;
  for (int tw_tiles = min_tw_tiles; tw_tiles <= min_tw_tiles + 2; ++tw_tiles) {
    int TW = (W + tw_tiles - 1) / tw_tiles;
    if (pool) TW = (TW + 1) & ~1;                  // pooled tiles start on even rows / columns
    if (TW > 62) continue;
    const int Wp = TW + 2;
    for (int R = pool ? 2 : 1; R <= (pool ? H + 1 : H); R += pool ? 2 : 1) {
      const int nblk1 = ((R + 2) * Wp + 127) / 128, nblk3 = (R * Wp + 127) / 128;
      if (nblk1 + nblk3 > 4) break;
      if (smem_for(R, Wp) > smem_cap) break;
      const long tiles = static_cast<long>(B) * ((H + R - 1) / R) * ((W + TW - 1) / TW);
      const long waves = (tiles + 2 * nsm - 1) / (2 * nsm);
      // per-tile cycle model: fixed + pw1 blocks (MMA + epilogue) + depthwise per pixel + pw2 blocks
      const double cost = waves * (600.0 + 450.0 * nblk1 + 6.0 * R * TW + 500.0 * nblk3);
      if (cost < best) { best = cost; bestR = R; bestTW = TW; }
    }
  }
  if (bestR == 0) return FD_EUNSUPPORTED;

  SepParams p;
  p.B = B; p.H = H; p.W = W; p.R = bestR; p.TW = bestTW; p.Wp = bestTW + 2;
  p.nblk1 = ((bestR + 2) * p.Wp + 127) / 128;
  p.nblk3 = (bestR * p.Wp + 127) / 128;
  p.tiles_w = (W + bestTW - 1) / bestTW;
  p.tiles_h = (H + bestR - 1) / bestR;
  p.num_tiles = B * p.tiles_w * p.tiles_h;
  p.nseg = (bestTW + 7) / 8;
  p.in_bytes = static_cast<uint32_t>((bestR + 2) * p.Wp * 128);
  p.in_buf_bytes = round1k(p.in_bytes);
  p.t1_rows = static_cast<uint32_t>((bestR + 2) * p.Wp);
  p.t1_bytes = t1_bytes_for(bestR, p.Wp);
  p.t2_bytes = round1k(static_cast<size_t>(bestR) * p.Wp * 128);
  p.inv_wp = static_cast<uint32_t>((65536 + p.Wp - 1) / p.Wp);
  const int cols = (p.nblk1 + p.nblk3) * kC;
  p.tmem_cols = cols <= 64 ? 64 : cols <= 128 ? 128 : 256;
  p.slope = slope;
  p.pool = pool ? 1 : 0;
  { const char* d = getenv("FD_SEP_TIMING"); p.dbg = d ? atoi(d) : 0; }
  p.dw = w_dw;
  const size_t smem = smem_for(bestR, p.Wp);

  CUtensorMap tm_x, tm_w1, tm_w2, tm_out;
  int rc = make_tmap_nhwc_bf16(&tm_x, x, B, H, W, C, p.Wp, bestR + 2);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_w1, w_pw1, kC, kC, kC, kC);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_w2, w_pw2, kC, kC, kC, kC);
  if (rc != FD_OK) return rc;
  rc = pool ? make_tmap_nhwc_bf16(&tm_out, out, B, H / 2, W / 2, C, bestTW / 2, bestR / 2)
            : make_tmap_nhwc_bf16(&tm_out, out, B, H, W, C, bestTW, bestR);
  if (rc != FD_OK) return rc;

  cudaError_t e = set_max_dyn_smem(sepblock_fwd_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = p.num_tiles < 2 * nsm ? p.num_tiles : 2 * nsm;
  e = launch_k(sepblock_fwd_kernel, dim3(grid), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), tm_x, tm_w1, tm_w2,
               tm_out, p);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return launch_status();
}
