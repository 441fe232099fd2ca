// A run of residual blocks of one spatial shape as ONE persistent kernel (forward, or the input-
// gradient chain of the backward pass).
//
// Replaces the per-conv launches of models/PoolResnet.py:33-43 (and models/Resnet.py:27-40) for the
// blocks behind the last pooling stage (blocks 2..9 of PoolResnet: 16 convolutions at 15x15) where a
// whole image (15x15x64 bf16 = 28.8 KB) fits in shared memory:
//   * one CTA owns one image; its activations never leave shared memory between layers.  They live in
//     the zero-padded halo layout of conv3x3_tc.cu ((H+2) x (W+1) pixels, one pixel = one 128-byte
//     128B-swizzled row), so the output tile of layer l *is* the A operand of layer l+1;
//   * the weights (72 KB per layer) stream from L2 through a 9-slot tap ring: slot t is refilled with
//     tap t of the next layer as soon as the last MMA that reads it has been committed;
//   * everything that has to reach HBM for the backward pass is written by TMA tensor stores straight
//     from those shared-memory tiles (out-of-bounds halo elements are clipped by the TMA unit);
//   * LeakyReLU' of the backward pass comes from 1-bit sign masks written by the forward pass
//     (8 bytes per pixel instead of a second 128-byte bf16 tensor).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..9 =
// epilogue (TMEM lane quadrant = warp % 4, channel half = (warp - 2) / 4).
#include "fd_host.h"
#include "fd_ptx.cuh"
#include <cstdlib>

namespace fd {
namespace {

constexpr int kC = 64;
constexpr int kTapBytes = kC * 128;        // one tap: [cout][cin] bf16, K-major, 128B swizzle
constexpr int kWBytes = 9 * kTapBytes;     // 73728
constexpr int kThreads = 320;
constexpr int kEpiThreads = 256;
constexpr int kMaxLayers = 24;
constexpr int kMaxMaps = 40;
constexpr int kNumBufs = 3;
constexpr uint32_t kTmemCols = 256;

struct ChainLayer {
  const float* bias;          // [C] or null
  const float* chan_scale;    // [B,C] or null (Dropout2d multiplier, after LeakyReLU)
  const float* chan_scale2;   // [B,C] or null (multiplier of the masked second output)
  const uint32_t* mask_in;    // [B,H,W,2] sign bits selecting 1 / slope for the second output, or null
  uint32_t* mask_v;           // [B,H,W,2] sign bits of the value before the residual add, or null
  int w_row;                  // first row of this layer's weights in the weight tensor map
  int flags;                  // FD_EPI_LRELU
  int8_t in_buf;              // smem buffer holding the conv input
  int8_t res_buf;             // smem buffer holding the running residual (read + updated in place), -1 none
  int8_t v_buf;               // smem buffer receiving the pre-residual value (no residual: the output), -1 none
  int8_t out2_buf;            // smem buffer receiving the masked second output, -1 none
  int8_t map_v, map_res, map_out2;  // TMA store maps for those buffers (-1: not written to HBM)
  int8_t pad_;
};

struct ChainParams {
  int B, H, W, Wp, nblk, n_layers, n_init, dbg;
  uint32_t buf_bytes, box_bytes;
  float slope;
  ChainLayer L[kMaxLayers];
  CUtensorMap maps[kMaxMaps];
};

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

__device__ __forceinline__ uint32_t sign_bits(const float (&v)[32]) {
  // bit j = (bf16(v[j]) > 0): the stored activation, not the fp32 value, decides (same rule as the
  // bf16 mask_src path of conv3x3_tc.cu)
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j) m |= (__bfloat162float(__float2bfloat16_rn(v[j])) > 0.f ? 1u : 0u) << j;
  return m;
}

__global__ void __launch_bounds__(kThreads, 1)
resblock_chain_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in0,
                      const __grid_constant__ CUtensorMap tm_in1, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* sW = smem;
  uint8_t* sBuf = smem + kWBytes;                               // kNumBufs x buf_bytes
  float* sConst = reinterpret_cast<float*>(sBuf + kNumBufs * p.buf_bytes);   // [3][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sConst + 3 * kC);
  uint64_t* w_full = bars;            // [9]
  uint64_t* w_empty = bars + 9;       // [9]
  uint64_t* in_full = bars + 18;
  uint64_t* act_ready = bars + 19;
  uint64_t* acc_full = bars + 20;     // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // halo pixels and the rows behind the box are never written again: they must read as zero
  for (uint32_t i = threadIdx.x * 16u; i < kNumBufs * p.buf_bytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sBuf + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_w);
    tma_prefetch_desc(&tm_in0);
    for (int t = 0; t < 9; ++t) {
      mbar_init(w_full + t, 1);
      mbar_init(w_empty + t, 1);
    }
    mbar_init(in_full, 1);
    mbar_init(act_ready, 1);
    for (int m = 0; m < 4; ++m) mbar_init(acc_full + m, 1);
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
      int g = 0, it = 0;
      for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
        for (int l = 0; l < p.n_layers; ++l, ++g) {
          const int row0 = p.L[l].w_row;
          for (int t = 0; t < 9; ++t) {
            mbar_wait(w_empty + t, (g & 1) ^ 1);
            mbar_expect_tx(w_full + t, kTapBytes);
            tma_load_2d(sW + t * kTapBytes, &tm_w, w_full + t, 0, row0 + t * kC);
          }
          if (l == 0) {
            // the buffers are free once the last epilogue of the previous image (and its stores) are done
            if (g > 0) mbar_wait(act_ready, (g - 1) & 1);
            mbar_expect_tx(in_full, p.box_bytes * p.n_init);
            tma_load_4d(sBuf, &tm_in0, in_full, 0, -1, -1, n);
            if (p.n_init > 1) tma_load_4d(sBuf + p.buf_bytes, &tm_in1, in_full, 0, -1, -1, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
    const uint32_t w_lo = sdesc_lo(smem_u32(sW), 16);
    int g = 0, it = 0;
    for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        if (l == 0) mbar_wait(in_full, it & 1);
        if (g > 0) mbar_wait(act_ready, (g - 1) & 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t in_lo = sdesc_lo(smem_u32(sBuf + p.L[l].in_buf * p.buf_bytes), 16);
          for (int mb = 0; mb < p.nblk; ++mb) {
            const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(mb * kC);
            const uint32_t a_blk = in_lo + static_cast<uint32_t>(mb * 128 * 8);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
              const int ky = t / 3, kx = t - 3 * ky;
              if (mb == 0) mbar_wait(w_full + t, g & 1);
              const uint32_t a_tap = a_blk + static_cast<uint32_t>((ky * p.Wp + kx) * 8);
              const uint32_t b_tap = w_lo + t * (kTapBytes / 16);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(d_tmem, sdesc_sw128(a_tap + 2 * k), sdesc_sw128(b_tap + 2 * k), idesc, (t | k) != 0 ? 1u : 0u);
              if (mb == p.nblk - 1) umma_commit(w_empty + t);   // slot t may be refilled with the next layer's tap
            }
            umma_commit(acc_full + mb);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (8)
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int c0 = hf * 32;
    const int et = threadIdx.x - 64;
    int g = 0, it = 0;
    for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
      mbar_wait(in_full, it & 1);
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        const ChainLayer& L = p.L[l];
        // stores issued two layers ago have finished reading the buffers this layer overwrites
        if (et == 0) tma_store_wait_read<1>();
        if (et < kC) sConst[et] = L.bias ? __ldg(L.bias + et) : 0.f;
        else if (et < 2 * kC) sConst[et] = L.chan_scale ? __ldg(L.chan_scale + n * kC + et - kC) : 1.f;
        else if (et < 3 * kC) sConst[et] = L.chan_scale2 ? __ldg(L.chan_scale2 + n * kC + et - 2 * kC) : 1.f;
        uint32_t mk[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
        if (L.mask_in) {
#pragma unroll
          for (int mb = 0; mb < 4; ++mb) {
            const int m = mb * 128 + q * 32 + lane;
            const int y = m / p.Wp, x = m - y * p.Wp;
            if (mb < p.nblk && y < p.H && x < p.W)
              mk[mb] = __ldg(L.mask_in + ((static_cast<size_t>(n) * p.H + y) * p.W + x) * 2 + hf);
          }
        }
        bar_sync_epi();
        uint8_t* vbuf = L.v_buf >= 0 ? sBuf + L.v_buf * p.buf_bytes : nullptr;
        uint8_t* rbuf = L.res_buf >= 0 ? sBuf + L.res_buf * p.buf_bytes : nullptr;
        uint8_t* obuf = L.out2_buf >= 0 ? sBuf + L.out2_buf * p.buf_bytes : nullptr;
#pragma unroll
        for (int mb = 0; mb < 4; ++mb) {
          if (mb < p.nblk) {
            const int m = mb * 128 + q * 32 + lane;
            const int y = m / p.Wp, x = m - y * p.Wp;
            const bool valid = y < p.H && x < p.W;
            const uint32_t r = static_cast<uint32_t>(m + p.Wp + 1);      // smem row of pixel (y, x)
            const uint32_t roff = r * 128u;
            const uint32_t sw = r & 7u;
            mbar_wait(acc_full + mb, g & 1);
            tc_fence_after();
            uint32_t acc[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mb * kC + c0),
                               acc);
            tmem_ld_wait();
            if (valid) {
              float v[32];
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                const float4 b4 = *reinterpret_cast<const float4*>(sConst + c0 + 4 * j4);
                v[4 * j4 + 0] = __uint_as_float(acc[4 * j4 + 0]) + b4.x;
                v[4 * j4 + 1] = __uint_as_float(acc[4 * j4 + 1]) + b4.y;
                v[4 * j4 + 2] = __uint_as_float(acc[4 * j4 + 2]) + b4.z;
                v[4 * j4 + 3] = __uint_as_float(acc[4 * j4 + 3]) + b4.w;
              }
              if (L.flags & FD_EPI_LRELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
              }
              if (L.chan_scale) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 s4 = *reinterpret_cast<const float4*>(sConst + kC + c0 + 4 * j4);
                  v[4 * j4 + 0] *= s4.x; v[4 * j4 + 1] *= s4.y; v[4 * j4 + 2] *= s4.z; v[4 * j4 + 3] *= s4.w;
                }
              }
              if (L.mask_v) L.mask_v[((static_cast<size_t>(n) * p.H + y) * p.W + x) * 2 + hf] = sign_bits(v);
              if (vbuf) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4 u;
                  u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                  u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                  *reinterpret_cast<uint4*>(vbuf + roff + (((hf * 4 + i) ^ sw) << 4)) = u;
                }
              }
              if (rbuf) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4* rp = reinterpret_cast<uint4*>(rbuf + roff + (((hf * 4 + i) ^ sw) << 4));
                  uint4 u = *rp;
                  v[8 * i + 0] += bf16lo(u.x); v[8 * i + 1] += bf16hi(u.x);
                  v[8 * i + 2] += bf16lo(u.y); v[8 * i + 3] += bf16hi(u.y);
                  v[8 * i + 4] += bf16lo(u.z); v[8 * i + 5] += bf16hi(u.z);
                  v[8 * i + 6] += bf16lo(u.w); v[8 * i + 7] += bf16hi(u.w);
                  u.x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]); u.y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                  u.z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]); u.w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                  *rp = u;
                }
              }
              if (obuf) {
                const uint32_t mbits = mk[mb];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float4 s0 = *reinterpret_cast<const float4*>(sConst + 2 * kC + c0 + 8 * i);
                  const float4 s1 = *reinterpret_cast<const float4*>(sConst + 2 * kC + c0 + 8 * i + 4);
                  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                  float o[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    float t = v[8 * i + e] * (((mbits >> (8 * i + e)) & 1u) ? 1.f : p.slope);
                    if (L.chan_scale2) t *= sc[e];
                    o[e] = t;
                  }
                  uint4 u;
                  u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
                  u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
                  *reinterpret_cast<uint4*>(obuf + roff + (((hf * 4 + i) ^ sw) << 4)) = u;
                }
              }
            }
          }
        }
        fence_proxy_async();     // generic-proxy smem writes -> visible to tcgen05.mma and the TMA stores
        tc_fence_before();
        bar_sync_epi();
        if (et == 0) {
          if (!(p.dbg & 1)) {
          const uint32_t o0 = static_cast<uint32_t>(p.Wp + 1) * 128u;   // smem row of pixel (0,0)
          if (L.map_v >= 0) tma_store_4d(&p.maps[L.map_v], vbuf + o0, 0, 0, 0, n);
          if (L.map_res >= 0) tma_store_4d(&p.maps[L.map_res], rbuf + o0, 0, 0, 0, n);
          if (L.map_out2 >= 0) tma_store_4d(&p.maps[L.map_out2], obuf + o0, 0, 0, 0, n);
          }
          tma_store_commit();
          if (l == p.n_layers - 1) tma_store_wait_read<0>();   // the next image reloads the buffers
          mbar_arrive(act_ready);
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

struct ChainBuilder {
  ChainParams p;
  int n_maps = 0;
  int rc = FD_OK;
  int B, H, W;
  int add_map(const void* ptr) {
    if (!ptr) return -1;
    if (n_maps >= kMaxMaps) { rc = FD_EUNSUPPORTED; return -1; }
    // store box = rows 0..H-1 of the padded tile, starting at pixel (0,0): only upper-bound clipping (column W)
    int r = make_tmap_nhwc_bf16(&p.maps[n_maps], ptr, B, H, W, kC, W + 1, H);
    if (r != FD_OK) { rc = r; return -1; }
    return n_maps++;
  }
};

int launch_chain(ChainBuilder& cb, const fd_bf16* w, int w_layers, const fd_bf16* in0, const fd_bf16* in1, int B, int H,
                 int W, float slope, cudaStream_t st) {
  if (cb.rc != FD_OK) return cb.rc;
  ChainParams& p = cb.p;
  const int Wp = W + 1;
  p.B = B; p.H = H; p.W = W; p.Wp = Wp;
  p.nblk = (H * Wp + 127) / 128;
  if (p.nblk > 4) return FD_EUNSUPPORTED;
  p.n_init = in1 ? 2 : 1;
  p.box_bytes = static_cast<uint32_t>((H + 2) * Wp * 128);
  p.buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.nblk * 128 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024);
  p.slope = slope;
  { const char* d = getenv("FD_CHAIN_DBG"); p.dbg = d ? atoi(d) : 0; }
  const size_t smem = kWBytes + static_cast<size_t>(kNumBufs) * p.buf_bytes + 3 * kC * 4 + 256 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  CUtensorMap tm_w, tm_in0, tm_in1;
  int rc = make_tmap_2d_bf16(&tm_w, w, w_layers * 9 * kC, kC, kC, kC);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_in0, in0, B, H, W, kC, Wp, H + 2);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_in1, in1 ? in1 : in0, B, H, W, kC, Wp, H + 2);
  if (rc != FD_OK) return rc;
  cudaError_t e = cudaFuncSetAttribute(resblock_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int nsm = sm_count();
  const int grid = B < nsm ? B : nsm;
  resblock_chain_kernel<<<grid, kThreads, smem, st>>>(tm_w, tm_in0, tm_in1, p);
  count_launch();
  return launch_status();
}

}  // namespace
}  // namespace fd

extern "C" int fd_resblock_chain_shape_ok(int H, int W, int C) {
  using namespace fd;
  if (C != kC || H <= 0 || W <= 0 || W + 1 > 256 || H + 2 > 256) return 0;
  const int Wp = W + 1;
  const int nblk = (H * Wp + 127) / 128;
  if (nblk > 4) return 0;
  const size_t buf = (static_cast<size_t>(nblk * 128 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024;
  return kWBytes + kNumBufs * buf + 3 * kC * 4 + 256 + 1024 <= 227 * 1024 ? 1 : 0;
}

extern "C" int fd_resblock_chain_fwd(const fd_bf16* x, const fd_bf16* w_fwd, const fd_chain_fwd_block* blocks,
                                     int n_blocks, int B, int H, int W, int C, float slope, void* stream) {
  using namespace fd;
  if (!x || !w_fwd || !blocks || n_blocks <= 0 || B <= 0) return FD_EINVAL;
  if (!fd_resblock_chain_shape_ok(H, W, C) || 2 * n_blocks > kMaxLayers) return FD_EUNSUPPORTED;
  if (!blocks[n_blocks - 1].out) return FD_EINVAL;
  ChainBuilder cb;
  cb.B = B; cb.H = H; cb.W = W;
  ChainParams& p = cb.p;
  p.n_layers = 2 * n_blocks;
  // buffers: 0 = X (block input / running residual / block output), 1 = a, 2 = b (pre-residual, only
  // materialised when it has to be stored)
  for (int k = 0; k < n_blocks; ++k) {
    const fd_chain_fwd_block& bk = blocks[k];
    ChainLayer& c1 = p.L[2 * k];
    c1 = ChainLayer{};
    c1.bias = bk.bias1; c1.mask_v = bk.mask_a;
    c1.w_row = (2 * k) * 9 * kC; c1.flags = FD_EPI_LRELU;
    c1.in_buf = 0; c1.res_buf = -1; c1.v_buf = 1; c1.out2_buf = -1;
    c1.map_v = static_cast<int8_t>(cb.add_map(bk.a)); c1.map_res = -1; c1.map_out2 = -1;
    ChainLayer& c2 = p.L[2 * k + 1];
    c2 = ChainLayer{};
    c2.bias = bk.bias2; c2.chan_scale = bk.chan_scale; c2.mask_v = bk.mask_b;
    c2.w_row = (2 * k + 1) * 9 * kC; c2.flags = FD_EPI_LRELU;
    c2.in_buf = 1; c2.res_buf = 0; c2.v_buf = bk.b ? 2 : -1; c2.out2_buf = -1;
    c2.map_v = static_cast<int8_t>(cb.add_map(bk.b)); c2.map_res = static_cast<int8_t>(cb.add_map(bk.out));
    c2.map_out2 = -1;
  }
  return launch_chain(cb, w_fwd, 2 * n_blocks, x, nullptr, B, H, W, slope, static_cast<cudaStream_t>(stream));
}

extern "C" int fd_resblock_chain_bwd(const fd_bf16* g_out, const fd_bf16* gp2_last, const fd_bf16* w_dgrad,
                                     const fd_chain_bwd_block* blocks, int n_blocks, int B, int H, int W, int C,
                                     float slope, void* stream) {
  using namespace fd;
  if (!g_out || !gp2_last || !w_dgrad || !blocks || n_blocks <= 0 || B <= 0) return FD_EINVAL;
  if (!fd_resblock_chain_shape_ok(H, W, C) || 2 * n_blocks > kMaxLayers) return FD_EUNSUPPORTED;
  ChainBuilder cb;
  cb.B = B; cb.H = H; cb.W = W;
  ChainParams& p = cb.p;
  p.n_layers = 2 * n_blocks;
  // buffers: 0 = G (gradient w.r.t. the block output, updated in place), 1 = gp2, 2 = gp1.
  // blocks[] is in FORWARD order; execution runs from the last block to the first.
  for (int j = 0; j < n_blocks; ++j) {
    const int k = n_blocks - 1 - j;
    const fd_chain_bwd_block& bk = blocks[k];
    if (!bk.mask_a) return FD_EINVAL;
    const bool last = (k == 0);
    if (!last && (!bk.mask_b_prev || !bk.gp2_prev)) return FD_EINVAL;
    ChainLayer& l1 = p.L[2 * j];         // gp1 = dgrad_conv2(gp2) * lrelu'(a)
    l1 = ChainLayer{};
    l1.mask_in = bk.mask_a;
    l1.w_row = (2 * k + 1) * 9 * kC;
    l1.in_buf = 1; l1.res_buf = -1; l1.v_buf = -1; l1.out2_buf = 2;
    l1.map_v = -1; l1.map_res = -1; l1.map_out2 = static_cast<int8_t>(cb.add_map(bk.gp1));
    ChainLayer& l2 = p.L[2 * j + 1];     // G' = dgrad_conv1(gp1) + G ; gp2_prev = G' * drop * lrelu'(b_prev)
    l2 = ChainLayer{};
    l2.mask_in = bk.mask_b_prev; l2.chan_scale2 = bk.chan_scale_prev;
    l2.w_row = (2 * k) * 9 * kC;
    l2.in_buf = 2; l2.res_buf = 0; l2.v_buf = -1; l2.out2_buf = bk.gp2_prev ? 1 : -1;
    l2.map_v = -1; l2.map_res = static_cast<int8_t>(cb.add_map(bk.g_in));
    l2.map_out2 = static_cast<int8_t>(cb.add_map(bk.gp2_prev));
  }
  return launch_chain(cb, w_dgrad, 2 * n_blocks, g_out, gp2_last, B, H, W, slope, static_cast<cudaStream_t>(stream));
}
