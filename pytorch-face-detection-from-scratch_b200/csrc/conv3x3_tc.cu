// 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels, as an implicit GEMM on the sm_100a
// tensor cores (tcgen05.mma, fp32 accumulators in TMEM), fed by TMA.
//
// Replaces aten::conv2d (+ leaky_relu / dropout2d / residual add) at models/PoolResnet.py:35-40
// of the reference, and -- with dgrad-packed weights -- the input-gradient half of its backward.
//
// "Halo-tile" formulation (no im2col, no per-tap reloads):
//   * one CTA tile = R output rows x W columns of one image.  ONE TMA box {64ch, Wp=W+1, R+2, 1}
//     starting at (w=-1, h=h0-1) lands the zero-padded input patch in shared memory as
//     (R+2)*Wp consecutive 128-byte rows (one pixel = 64 bf16 = one 128B-swizzle row).  The
//     single halo column serves as right halo of row y and left halo of row y+1.
//   * GEMM row m = y*Wp + x.  For tap (ky,kx) the A operand of output row m is input row
//     m + ky*Wp + kx: the *same* shared-memory tile, read through a matrix descriptor whose start
//     address is shifted by (ky*Wp+kx)*128 bytes.  9 taps x 4 K-steps = 36 MMAs (M=128,N=64,K=16)
//     per 128-row block, all operands already resident; weights (72 KB) stay in smem for the
//     whole persistent CTA.
//   * rows with x == W or y >= R are junk and are simply not stored.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..9 = epilogue (TMEM -> registers -> fused bias/LeakyReLU/dropout/residual/mask -> global).
// Input tiles and TMEM accumulators are double buffered so the epilogue of tile i overlaps the
// MMAs of tile i+1 and the TMA of tile i+2.
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kWBytes = 9 * kC * 128;  // 73728: [tap][cout][cin] bf16, K-major, 128B swizzle
constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr uint32_t kTmemCols = 512;

struct ConvParams {
  int B, H, W, R, Wp, nblk, tiles_per_img, num_tiles;
  uint32_t in_bytes;      // bytes delivered by one input TMA box
  uint32_t in_buf_bytes;  // bytes reserved per input copy (multiple of 1024)
  int flags;
  float slope;
  const float* bias;
  const float* chan_scale;
  const __nv_bfloat16* residual;
  __nv_bfloat16* aux_out;
  __nv_bfloat16* out;
  const __nv_bfloat16* mask_src;
  const float* chan_scale2;
  __nv_bfloat16* out2;
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
    u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
    u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
    u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
    d[i] = u;
  }
}
__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float (&v)[32]) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u = __ldg(s + i);
    v[8 * i + 0] = bf16lo(u.x); v[8 * i + 1] = bf16hi(u.x);
    v[8 * i + 2] = bf16lo(u.y); v[8 * i + 3] = bf16hi(u.y);
    v[8 * i + 4] = bf16lo(u.z); v[8 * i + 5] = bf16hi(u.z);
    v[8 * i + 6] = bf16lo(u.w); v[8 * i + 7] = bf16hi(u.w);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                  const ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const uint32_t stage_bytes = p.in_buf_bytes;
  uint8_t* sW = smem;
  uint8_t* sIn = smem + kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sIn + 2 * stage_bytes);
  uint64_t* w_full = bars + 0;
  uint64_t* in_full = bars + 1;    // [2]
  uint64_t* in_empty = bars + 3;   // [2]
  uint64_t* acc_full = bars + 5;   // [2]
  uint64_t* acc_empty = bars + 7;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // Rows the TMA never writes (tail of each buffer) must read as zero: the tap (2,2) of the last
  // valid pixel of a tile wraps onto the first row behind the box.
  for (uint32_t i = threadIdx.x * 16u; i < 2 * stage_bytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sIn + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(in_full + s, 1);
      mbar_init(in_empty + s, 1);
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 8);
    }
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      mbar_expect_tx(w_full, kWBytes);
      for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * 8192, &tm_w, w_full, 0, t * kC);
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int n = tile / p.tiles_per_img;
        const int h0 = (tile - n * p.tiles_per_img) * p.R;
        mbar_wait(in_empty + s, ph ^ 1);
        mbar_expect_tx(in_full + s, p.in_bytes);
        tma_load_4d(sIn + s * stage_bytes, &tm_in, in_full + s, 0, -1, h0 - 1, n);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
    const uint32_t w_lo = sdesc_lo(smem_u32(sW), 16);
    mbar_wait(w_full, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc_empty + s, ph ^ 1);
      mbar_wait(in_full + s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t in_lo = sdesc_lo(smem_u32(sIn + s * stage_bytes), 16);
        for (int mb = 0; mb < p.nblk; ++mb) {
          const uint32_t d_tmem = tmem_base + static_cast<uint32_t>((s * p.nblk + mb) * kC);
          const uint32_t a_blk = in_lo + static_cast<uint32_t>(mb * 128 * 8);   // 128 rows x 128 B, in 16-B units
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int ky = t / 3, kx = t - 3 * ky;
            // tap (ky,kx): the same tile, start address shifted by (ky*Wp + kx) rows
            const uint32_t a_tap = a_blk + static_cast<uint32_t>((ky * p.Wp + kx) * 8);
            const uint32_t b_tap = w_lo + t * 512;
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(d_tmem, sdesc_sw128(a_tap + 2 * k), sdesc_sw128(b_tap + 2 * k), idesc, (t | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(in_empty + s);   // input tile free once these MMAs have read it
        umma_commit(acc_full + s);   // accumulators ready for the epilogue
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (8)
    // warp -> TMEM lane quadrant q (hardware rule: warp % 4) and channel half hf: each thread owns
    // 32 channels (64 B) of one output row.  Residual / mask rows of the next two blocks are
    // prefetched into registers so that their L2/HBM latency overlaps the MMAs and the math.
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int c0 = hf * 32;
    const bool need_res = p.residual != nullptr;
    const bool need_mask = p.out2 != nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int n = tile / p.tiles_per_img;
      const int h0 = (tile - n * p.tiles_per_img) * p.R;
      uint4 pre_r[2][4], pre_m[2][4];
      auto geometry = [&](int mb, bool& valid, size_t& pix) {
        const int m = mb * 128 + q * 32 + lane;
        const int y = m / p.Wp;
        const int x = m - y * p.Wp;
        const int oy = h0 + y;
        valid = (y < p.R) && (x < p.W) && (oy < p.H);
        pix = (static_cast<size_t>(n) * p.H + oy) * p.W + x;
      };
#define FD_PREFETCH(MB, SLOT)                                                                        \
  {                                                                                                  \
    bool v_;                                                                                         \
    size_t px_;                                                                                      \
    geometry(MB, v_, px_);                                                                           \
    if (v_) {                                                                                        \
      if (need_res) {                                                                                \
        const uint4* r_ = reinterpret_cast<const uint4*>(p.residual + px_ * kC + c0);                \
        _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) pre_r[SLOT][i_] = __ldg(r_ + i_);           \
      }                                                                                              \
      if (need_mask) {                                                                               \
        const uint4* m_ = reinterpret_cast<const uint4*>(p.mask_src + px_ * kC + c0);                \
        _Pragma("unroll") for (int i_ = 0; i_ < 4; ++i_) pre_m[SLOT][i_] = __ldg(m_ + i_);           \
      }                                                                                              \
    }                                                                                                \
  }
      FD_PREFETCH(0, 0)
      if (p.nblk > 1) FD_PREFETCH(1, 1)
      mbar_wait(acc_full + s, ph);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < 4; ++mb) {
        if (mb < p.nblk) {
          bool valid;
          size_t pix;
          geometry(mb, valid, pix);
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>((s * p.nblk + mb) * kC + c0),
                             acc);
          tmem_ld_wait();
          if (valid) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __ldg(p.bias + c0 + j);
            }
            if (p.flags & FD_EPI_LRELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
            }
            if (p.chan_scale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= __ldg(p.chan_scale + n * kC + c0 + j);
            }
            if (p.aux_out) store_bf16x32(p.aux_out + pix * kC + c0, v);
            if (need_res) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = pre_r[mb & 1][i];
                v[8 * i + 0] += bf16lo(u.x); v[8 * i + 1] += bf16hi(u.x);
                v[8 * i + 2] += bf16lo(u.y); v[8 * i + 3] += bf16hi(u.y);
                v[8 * i + 4] += bf16lo(u.z); v[8 * i + 5] += bf16hi(u.z);
                v[8 * i + 6] += bf16lo(u.w); v[8 * i + 7] += bf16hi(u.w);
              }
            }
            if (p.out) store_bf16x32(p.out + pix * kC + c0, v);
            if (need_mask) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 u = pre_m[mb & 1][i];
                const float mk[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y),
                                     bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  float g = v[8 * i + e] * (mk[e] > 0.f ? 1.f : p.slope);
                  if (p.chan_scale2) g *= __ldg(p.chan_scale2 + n * kC + c0 + 8 * i + e);
                  v[8 * i + e] = g;
                }
              }
              store_bf16x32(p.out2 + pix * kC + c0, v);
            }
          }
          if (mb + 2 < p.nblk) FD_PREFETCH(mb + 2, mb & 1)
        }
      }
#undef FD_PREFETCH
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + s);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace
}  // namespace fd

extern "C" int fd_conv3x3(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                          float slope, const float* chan_scale, const fd_bf16* residual, fd_bf16* aux_out,
                          fd_bf16* out, const fd_bf16* mask_src, const float* chan_scale2, fd_bf16* out2,
                          int flags, void* stream) {
  using namespace fd;
  if (!x || !w_packed || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != kC || W + 8 > 256) return FD_EUNSUPPORTED;
  if ((out2 != nullptr) != (mask_src != nullptr)) return FD_EINVAL;
  if (!out && !aux_out && !out2) return FD_EINVAL;
  const int nsm = sm_count();
  const int Wp = W + 1;
  const int ncopies = 1;
  const size_t smem_cap = 227 * 1024;

  // Pick the rows-per-tile R that minimises a simple cycle model: MMA time of the padded
  // row blocks plus the TMA fill, times the number of waves over the SMs.
  int bestR = 0;
  double best = 1e30;
  for (int R = 1; R <= H && R + 2 <= 256; ++R) {
    const int nblk = (R * Wp + 127) / 128;
    if (nblk > 4) break;
    const size_t in_buf = (static_cast<size_t>(nblk * 128 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024;
    const size_t need = kWBytes + 2 * ncopies * in_buf + 256 + 1024;
    if (need > smem_cap) break;
    const long tiles = static_cast<long>(B) * ((H + R - 1) / R);
    const long waves = (tiles + nsm - 1) / nsm;
    const double per_tile = 1152.0 * nblk + 0.35 * ncopies * (R + 2) * Wp * 128 / 32.0 + 700.0;
    const double cost = waves * per_tile;
    if (cost < best) { best = cost; bestR = R; }
  }
  if (bestR == 0) return FD_EUNSUPPORTED;

  ConvParams p;
  p.B = B; p.H = H; p.W = W; p.R = bestR; p.Wp = Wp;
  p.nblk = (bestR * Wp + 127) / 128;
  p.tiles_per_img = (H + bestR - 1) / bestR;
  p.num_tiles = B * p.tiles_per_img;
  p.in_bytes = static_cast<uint32_t>((bestR + 2) * Wp * 128);
  p.in_buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.nblk * 128 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024);
  p.flags = flags;
  p.slope = slope;
  p.bias = bias; p.chan_scale = chan_scale;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  p.aux_out = reinterpret_cast<__nv_bfloat16*>(aux_out);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask_src);
  p.chan_scale2 = chan_scale2;
  p.out2 = reinterpret_cast<__nv_bfloat16*>(out2);

  CUtensorMap tm_in, tm_w;
  int rc = make_tmap_nhwc_bf16(&tm_in, x, B, H, W, C, Wp, bestR + 2);
  if (rc != FD_OK) return rc;
  rc = make_tmap_2d_bf16(&tm_w, w_packed, 9 * kC, kC, kC, kC);
  if (rc != FD_OK) return rc;

  const size_t smem = kWBytes + 2 * static_cast<size_t>(ncopies) * p.in_buf_bytes + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int grid = p.num_tiles < nsm ? p.num_tiles : nsm;
  conv3x3_tc_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(tm_in, tm_w, p);
  count_launch();
  return launch_status();
}
