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

// Optional per-layer timestamps of CTA 0 (FD_CHAIN_TIMING=1): [layer][16] clock64 values.
__device__ unsigned long long g_chain_dbg[kMaxLayers * 16];
#define FD_TS(slot) do { if (p.dbg && blockIdx.x == 0 && (lane == 0 || warp == 1)) g_chain_dbg[l * 16 + (slot)] = clock64(); } while (0)

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
    mbar_init(w_full, 1);
    for (int t = 0; t < 9; ++t) mbar_init(w_empty + t, 1);
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
            mbar_wait_sleep(w_empty + t, (g & 1) ^ 1);
            if (t == 0) mbar_expect_tx(w_full, kWBytes);      // one "weights of this layer landed" barrier
            tma_load_2d(sW + t * kTapBytes, &tm_w, w_full, 0, row0 + t * kC);
          }
          if (l == 0) {
            // the buffers are free once the last epilogue of the previous image (and its stores) are done
            if (g > 0) mbar_wait_sleep(act_ready, (g - 1) & 1);
            mbar_expect_tx(in_full, p.box_bytes * p.n_init);
            tma_load_4d(sBuf, &tm_in0, in_full, 0, -1, -1, n);
            if (p.n_init > 1) tma_load_4d(sBuf + p.buf_bytes, &tm_in1, in_full, 0, -1, -1, n);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread runs the whole loop nest)
    if (elect_one_sync()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, kC, 0, 0);
      const uint32_t w_lo = sdesc_lo(smem_u32(sW), 16);
      const uint32_t wp_units = static_cast<uint32_t>(p.Wp) * 8u;
      const int last_mb = p.nblk - 1;
      int g = 0, it = 0;
      for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
        for (int l = 0; l < p.n_layers; ++l, ++g) {
          if (l == 0) mbar_wait(in_full, it & 1);
          if (g > 0) mbar_wait(act_ready, (g - 1) & 1);
          mbar_wait(w_full, g & 1);
          tc_fence_after();
          FD_TS(0);
          const uint32_t in_lo = sdesc_lo(smem_u32(sBuf + p.L[l].in_buf * p.buf_bytes), 16);
#pragma unroll 1
          for (int mb = 0; mb < p.nblk; ++mb) {
            const bool release = (mb == last_mb);
            issue_conv3x3_block(tmem_base + static_cast<uint32_t>(mb * kC), in_lo + static_cast<uint32_t>(mb * 1024),
                                w_lo, wp_units, idesc, [&](int t) {
                                  if (release) umma_commit(w_empty + t);   // slot t may be refilled
                                });
            umma_commit(acc_full + mb);
          }
          FD_TS(1);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (8)
    // Issue-slot bound (4 SMSPs x 1 instr/clk), so the arithmetic runs on packed fp32x2 (FADD2/FMUL2),
    // LeakyReLU is max(v, slope*v) (0 <= slope <= 1) and the mb loop is rolled (I-cache).
    const int q = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int c0 = hf * 32;
    const int et = threadIdx.x - 64;
    const uint64_t slope2 = pk2(p.slope, p.slope);
    const float4* sBias4 = reinterpret_cast<const float4*>(sConst + c0);
    const float4* sCs4 = reinterpret_cast<const float4*>(sConst + kC + c0);
    int g = 0, it = 0;
    for (int n = blockIdx.x; n < p.B; n += gridDim.x, ++it) {
      mbar_wait_sleep(in_full, it & 1);
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        const ChainLayer& L = p.L[l];
        // stores issued two layers ago have finished reading the buffers this layer overwrites
        if (warp == 2) FD_TS(2);
        if (et == 0) tma_store_wait_read<1>();
        if (et < kC) sConst[et] = L.bias ? __ldg(L.bias + et) : 0.f;
        else if (et < 2 * kC) sConst[et] = L.chan_scale ? __ldg(L.chan_scale + n * kC + et - kC) : 1.f;
        else if (et < 3 * kC) sConst[et] = L.chan_scale2 ? __ldg(L.chan_scale2 + n * kC + et - 2 * kC) : 1.f;
        bar_sync_epi();
        uint8_t* vbuf = L.v_buf >= 0 ? sBuf + L.v_buf * p.buf_bytes : nullptr;
        uint8_t* rbuf = L.res_buf >= 0 ? sBuf + L.res_buf * p.buf_bytes : nullptr;
        uint8_t* obuf = L.out2_buf >= 0 ? sBuf + L.out2_buf * p.buf_bytes : nullptr;
        const bool lrelu = (L.flags & FD_EPI_LRELU) != 0;
        const bool has_cs = L.chan_scale != nullptr, has_cs2 = L.chan_scale2 != nullptr;
#pragma unroll 1
        for (int mb = 0; mb < p.nblk; ++mb) {
          const int m = mb * 128 + q * 32 + lane;
          const int y = m / p.Wp, x = m - y * p.Wp;
          const bool valid = y < p.H && x < p.W;
          const size_t mword = ((static_cast<size_t>(n) * p.H + y) * p.W + x) * 2 + hf;
          uint32_t mbits = 0xffffffffu;
          if (valid && L.mask_in) mbits = __ldg(L.mask_in + mword);      // latency hidden by the MMA wait
          const uint32_t r = static_cast<uint32_t>(m + p.Wp + 1);         // smem row of pixel (y, x)
          const uint32_t roff = r * 128u;
          const uint32_t sw = r & 7u;
          mbar_wait_sleep(acc_full + mb, g & 1, 1000);
          tc_fence_after();
          if (warp == 2) FD_TS(3 + 4 * mb);
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mb * kC + c0),
                             acc);
          tmem_ld_wait();
          if (warp == 2) FD_TS(4 + 4 * mb);
          if (valid) {
            uint64_t v2[16];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b4 = sBias4[i];
              v2[2 * i] = add2(pk2u(acc[4 * i], acc[4 * i + 1]), pk2(b4.x, b4.y));
              v2[2 * i + 1] = add2(pk2u(acc[4 * i + 2], acc[4 * i + 3]), pk2(b4.z, b4.w));
            }
            if (lrelu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a0, a1, t0, t1;
                upk2(v2[i], a0, a1);
                upk2(mul2(v2[i], slope2), t0, t1);
                v2[i] = pk2(fmaxf(a0, t0), fmaxf(a1, t1));
              }
            }
            if (has_cs) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 s4 = sCs4[i];
                v2[2 * i] = mul2(v2[2 * i], pk2(s4.x, s4.y));
                v2[2 * i + 1] = mul2(v2[2 * i + 1], pk2(s4.z, s4.w));
              }
            }
            if (L.mask_v) {
              uint32_t mv = 0;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a0, a1;
                upk2(v2[i], a0, a1);
                mv |= (__float_as_uint(a0) >> 31) << (2 * i);
                mv |= (__float_as_uint(a1) >> 31) << (2 * i + 1);
              }
              L.mask_v[mword] = ~mv;      // bit set iff the sign bit is clear (same rule as conv3x3_tc.cu)
            }
            if (vbuf) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4 u;
                float a0, a1;
                upk2(v2[4 * i + 0], a0, a1); u.x = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 1], a0, a1); u.y = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 2], a0, a1); u.z = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 3], a0, a1); u.w = pack_bf16x2(a0, a1);
                *reinterpret_cast<uint4*>(vbuf + roff + (((hf * 4 + i) ^ sw) << 4)) = u;
              }
            }
            if (rbuf) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint4* rp = reinterpret_cast<uint4*>(rbuf + roff + (((hf * 4 + i) ^ sw) << 4));
                uint4 u = *rp;
                v2[4 * i + 0] = add2(v2[4 * i + 0], pk2u(u.x << 16, u.x & 0xFFFF0000u));
                v2[4 * i + 1] = add2(v2[4 * i + 1], pk2u(u.y << 16, u.y & 0xFFFF0000u));
                v2[4 * i + 2] = add2(v2[4 * i + 2], pk2u(u.z << 16, u.z & 0xFFFF0000u));
                v2[4 * i + 3] = add2(v2[4 * i + 3], pk2u(u.w << 16, u.w & 0xFFFF0000u));
                float a0, a1;
                upk2(v2[4 * i + 0], a0, a1); u.x = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 1], a0, a1); u.y = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 2], a0, a1); u.z = pack_bf16x2(a0, a1);
                upk2(v2[4 * i + 3], a0, a1); u.w = pack_bf16x2(a0, a1);
                *rp = u;
              }
            }
            if (obuf) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float a0, a1;
                  upk2(v2[4 * i + e], a0, a1);
                  const int j = 8 * i + 2 * e;
                  a0 = ((mbits >> j) & 1u) ? a0 : a0 * p.slope;
                  a1 = ((mbits >> (j + 1)) & 1u) ? a1 : a1 * p.slope;
                  if (has_cs2) {
                    const float2 s2 = *reinterpret_cast<const float2*>(sConst + 2 * kC + c0 + j);
                    float b0, b1;
                    upk2(mul2(pk2(a0, a1), pk2(s2.x, s2.y)), b0, b1);
                    a0 = b0; a1 = b1;
                  }
                  w[e] = pack_bf16x2(a0, a1);
                }
                *reinterpret_cast<uint4*>(obuf + roff + (((hf * 4 + i) ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
              }
            }
          }
        }
        if (warp == 2) FD_TS(11);
        fence_proxy_async();     // generic-proxy smem writes -> visible to tcgen05.mma and the TMA stores
        tc_fence_before();
        if (warp == 2) FD_TS(12);
        bar_sync_epi();
        if (warp == 2) FD_TS(13);
        if (et == 0) {
          const bool last_layer = (l == p.n_layers - 1);
          if (!last_layer) mbar_arrive(act_ready);             // the MMAs only read the buffers, like the stores
          const uint32_t o0 = static_cast<uint32_t>(p.Wp + 1) * 128u;   // smem row of pixel (0,0)
          if (L.map_v >= 0) tma_store_4d(&p.maps[L.map_v], vbuf + o0, 0, 0, 0, n);
          if (L.map_res >= 0) tma_store_4d(&p.maps[L.map_res], rbuf + o0, 0, 0, 0, n);
          if (L.map_out2 >= 0) tma_store_4d(&p.maps[L.map_out2], obuf + o0, 0, 0, 0, n);
          tma_store_commit();
          if (last_layer) {
            tma_store_wait_read<0>();   // the next image reloads the buffers
            mbar_arrive(act_ready);
          }
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
  if (!(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;   // LeakyReLU is evaluated as max(v, slope*v)
  ChainParams& p = cb.p;
  const int Wp = W + 1;
  p.B = B; p.H = H; p.W = W; p.Wp = Wp;
  p.nblk = (H * Wp + 127) / 128;
  if (p.nblk > 4) return FD_EUNSUPPORTED;
  p.n_init = in1 ? 2 : 1;
  p.box_bytes = static_cast<uint32_t>((H + 2) * Wp * 128);
  p.buf_bytes = static_cast<uint32_t>((static_cast<size_t>(p.nblk * 128 + 2 * Wp + 2) * 128 + 1023) / 1024 * 1024);
  p.slope = slope;
  { const char* d = getenv("FD_CHAIN_TIMING"); p.dbg = d ? atoi(d) : 0; }
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

extern "C" FD_API int fd_debug_chain_timing(unsigned long long* out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fd::g_chain_dbg, sizeof(unsigned long long) * n));
}

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
  // buffers: 0 = X (block input / running residual / block output), 1 = a
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
    c2.in_buf = 1; c2.res_buf = 0; c2.v_buf = -1; c2.out2_buf = -1;
    c2.map_v = -1; c2.map_res = static_cast<int8_t>(cb.add_map(bk.out));
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
