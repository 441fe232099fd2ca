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
constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = (2 + kEpiWarps) * 32;   // 576: TMA producer, MMA issuer, 16 epilogue warps
constexpr int kMaxLayers = 24;
constexpr int kMaxMaps = 40;
constexpr int kNumBufs = 3;
constexpr uint32_t kTmemCols = 256;

struct ChainLayer {
  const float* bias;          // [C] or null
  const float* chan_scale;    // [B,C] or null (Dropout2d multiplier, after LeakyReLU)
  const float* chan_scale2;   // [B,C] or null (multiplier of the masked second output)
  const uint16_t* mask_in;    // sign bits selecting 1 / slope for the second output ([pixel][4] x 16 channels), or null
  uint16_t* mask_v;           // sign bits of the value before the residual add, or null
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
  int rows_per_blk;           // split mode: image rows per 128-row block (128 / Wp)
  uint32_t buf_bytes, box_bytes;
  uint32_t inv_wp;            // ceil(65536 / Wp)
  float slope;
  ChainLayer L[kMaxLayers];
  CUtensorMap maps[kMaxMaps];
};

// Optional per-layer timestamps of CTA 0 (FD_CHAIN_TIMING=1): [layer][16] clock64 values.
__device__ unsigned long long g_chain_dbg[2 * kMaxLayers * 16];
#define FD_TS(slot) do { if (p.dbg && blockIdx.x < 2) g_chain_dbg[(blockIdx.x * kMaxLayers + l) * 16 + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void bar_sync_epi() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// kSplit = false: one CTA per image, all 128-row blocks.
// kSplit = true : a 2-CTA cluster per image, CTA rank r owns block r (the image is exactly two blocks and a block
//                 is a whole number of image rows).  Both CTAs keep full-size buffers; an epilogue thread whose
//                 output row lies within Wp+1 rows of the block boundary also writes it into the peer's buffer
//                 through DSMEM (those rows are the halo of the peer's next convolution), and the per-layer
//                 "activations ready" barrier of each CTA collects one local and one remote arrival.
template <bool kSplit>
__global__ void __launch_bounds__(kThreads, 1)
resblock_chain_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in0,
                      const __grid_constant__ CUtensorMap tm_in1, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  uint8_t* sW = smem;
  uint8_t* sBuf = smem + kWBytes;                               // kNumBufs x buf_bytes
  float* sConst = reinterpret_cast<float*>(sBuf + kNumBufs * p.buf_bytes);   // [2 layers][3][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sConst + 2 * 3 * kC);
  uint64_t* w_full = bars;            // [1] (+8 unused)
  uint64_t* w_empty = bars + 9;       // [9]
  uint64_t* in_full = bars + 18;
  uint64_t* act_ready = bars + 19;
  uint64_t* acc_full = bars + 20;     // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = kSplit ? cluster_ctarank() : 0u;
  const uint32_t peer = rank ^ 1u;
  const int img0 = kSplit ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int img_step = kSplit ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int mb_lo = kSplit ? static_cast<int>(rank) : 0;
  const int mb_hi = kSplit ? static_cast<int>(rank) + 1 : p.nblk;

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
  if (kSplit) cluster_sync_all();     // the peer's barriers and zeroed buffers exist before anyone touches them
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one_sync()) {
      int g = 0, it = 0;
      for (int n = img0; n < p.B; n += img_step, ++it) {
        for (int l = 0; l < p.n_layers; ++l, ++g) {
          const int row0 = p.L[l].w_row;
          for (int t = 0; t < 9; ++t) {
            mbar_wait_sleep(w_empty + t, (g & 1) ^ 1);
            if (t == 0) mbar_expect_tx(w_full, kWBytes);      // one "weights of this layer landed" barrier
            tma_load_2d(sW + t * kTapBytes, &tm_w, w_full, 0, row0 + t * kC);
          }
          if (l == 0) {
            // the buffers are free once the last epilogues of the previous image (both CTAs) and its stores are done
            if (g > 0) {
              if (kSplit) mbar_wait_cluster(act_ready, (g - 1) & 1);
              else mbar_wait_sleep(act_ready, (g - 1) & 1);
            }
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
      int g = 0, it = 0;
      for (int n = img0; n < p.B; n += img_step, ++it) {
        for (int l = 0; l < p.n_layers; ++l, ++g) {
          if (l == 0) mbar_wait(in_full, it & 1);
          if (g > 0) {
            if (kSplit) {
              mbar_wait_cluster(act_ready, (g - 1) & 1);   // local epilogue arrived AND the peer's halo bytes landed
              FD_TS(4);
            } else {
              mbar_wait(act_ready, (g - 1) & 1);
            }
          }
          mbar_wait(w_full, g & 1);
          tc_fence_after();
          FD_TS(0);
          const uint32_t in_lo = sdesc_lo(smem_u32(sBuf + p.L[l].in_buf * p.buf_bytes), 16);
#pragma unroll 1
          for (int mb = mb_lo; mb < mb_hi; ++mb) {
            const bool release = (mb == mb_hi - 1);
            issue_conv3x3_block(tmem_base + static_cast<uint32_t>(mb * kC), in_lo + static_cast<uint32_t>(mb * 1024),
                                w_lo, wp_units, idesc, [&](int t) {
                                  if (release) umma_commit(w_empty + t);   // slot t may be refilled
                                });
            umma_commit(acc_full + mb);
          }
          FD_TS(1);
          if (p.dbg) { mbar_wait(acc_full + mb_hi - 1, g & 1); FD_TS(5); }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (16)
    // One thread = one GEMM row (pixel) x 16 channels; TMEM lane quadrant = warp % 4, channel quarter = (warp-2)/4.
    // The epilogue of the last block of a layer is on the critical path (the next layer needs every pixel), so the
    // arithmetic is packed fp32x2 and the work is spread over 16 warps for latency hiding.
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;
    const int c0 = cq * 16;
    const int et = threadIdx.x - 64;
    const uint64_t slope2 = pk2(p.slope, p.slope);
    const uint32_t act_ready_peer = kSplit ? mapa_shared(smem_u32(act_ready), peer) : 0u;
    // rows whose output is part of the peer's halo: the last Wp+1 rows of block 0 / the first Wp+1 rows of block 1.
    // They are mirrored into the peer's buffer with st.async, which reports its bytes to the PEER's act_ready
    // barrier; each CTA therefore expects 128 B per valid incoming halo row and written buffer, every layer.
    const int mir_lo = rank == 0 ? 128 - (p.Wp + 1) : 128;
    const int mir_hi = rank == 0 ? 128 : 128 + p.Wp + 1;
    uint32_t in_rows = 0;              // valid rows the peer mirrors into this CTA
    if (kSplit) {
      const int lo = rank == 0 ? 128 : 128 - (p.Wp + 1), hi = rank == 0 ? 128 + p.Wp + 1 : 128;
      for (int m = lo; m < hi; ++m) {
        const int y = m / p.Wp, x = m - y * p.Wp;
        in_rows += (y < p.H && x < p.W) ? 1u : 0u;
      }
    }
    // Per-layer constants (bias, Dropout2d multipliers) live in a double-buffered smem table: the values of layer
    // g+1 are fetched while layer g waits for its MMAs, so no global-load latency sits between two layers.
    auto load_const = [&](int layer, int img) -> float {
      const ChainLayer& Ln = p.L[layer];
      if (et < kC) return Ln.bias ? __ldg(Ln.bias + et) : 0.f;
      if (et < 2 * kC) return Ln.chan_scale ? __ldg(Ln.chan_scale + img * kC + et - kC) : 1.f;
      if (et < 3 * kC) return Ln.chan_scale2 ? __ldg(Ln.chan_scale2 + img * kC + et - 2 * kC) : 1.f;
      return 0.f;
    };
    if (img0 < p.B && et < 3 * kC) sConst[et] = load_const(0, img0);
    bar_sync_epi();
    int g = 0, it = 0;
    for (int n = img0; n < p.B; n += img_step, ++it) {
      mbar_wait_sleep(in_full, it & 1);
      for (int l = 0; l < p.n_layers; ++l, ++g) {
        const ChainLayer& L = p.L[l];
        if (et == 0) FD_TS(2);
        const float* sC = sConst + (g & 1) * 3 * kC;
        // constants of the next layer (or of layer 0 of this CTA's next image): loads in flight during the MMA wait
        const bool last_layer_of_img = (l == p.n_layers - 1);
        const int nl = last_layer_of_img ? 0 : l + 1;
        const int nn = last_layer_of_img ? n + img_step : n;
        float next_c = 0.f;
        if (et < 3 * kC && nn < p.B) next_c = load_const(nl, nn);
        uint8_t* vbuf = L.v_buf >= 0 ? sBuf + L.v_buf * p.buf_bytes : nullptr;
        uint8_t* rbuf = L.res_buf >= 0 ? sBuf + L.res_buf * p.buf_bytes : nullptr;
        uint8_t* obuf = L.out2_buf >= 0 ? sBuf + L.out2_buf * p.buf_bytes : nullptr;
        const bool lrelu = (L.flags & FD_EPI_LRELU) != 0;
        const bool has_cs = L.chan_scale != nullptr, has_cs2 = L.chan_scale2 != nullptr;
#pragma unroll 1
        for (int mb = mb_lo; mb < mb_hi; ++mb) {
          const int m = mb * 128 + q * 32 + lane;
          const int y = static_cast<int>((static_cast<uint32_t>(m) * p.inv_wp) >> 16);
          const int x = m - y * p.Wp;
          const bool valid = y < p.H && x < p.W;
          const size_t mword = ((static_cast<size_t>(n) * p.H + y) * p.W + x) * 4 + cq;
          uint32_t mbits = 0xffffu;
          if (valid && L.mask_in) mbits = __ldg(L.mask_in + mword);       // latency hidden by the MMA wait
          const uint32_t r = static_cast<uint32_t>(m + p.Wp + 1);         // smem row of pixel (y, x)
          const uint32_t o0 = r * 128u + (((static_cast<uint32_t>(cq) * 2u) ^ (r & 7u)) << 4);
          const uint32_t o1 = r * 128u + (((static_cast<uint32_t>(cq) * 2u + 1u) ^ (r & 7u)) << 4);
          const bool mirror = kSplit && m >= mir_lo && m < mir_hi;
          if (et == 0) FD_TS(8);
          if (lane == 0) mbar_wait(acc_full + mb, g & 1);     // one poller per warp
          __syncwarp();
          tc_fence_after();
          if (et == 0) FD_TS(3);
          uint32_t acc[16];
          tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(mb * kC + c0),
                             acc);
          tmem_ld_wait();
          if (valid) {
            uint64_t v2[8];
            epi_bias_act16(acc, sC + c0, sC + kC + c0, lrelu, has_cs, slope2, v2);
            if (L.mask_v) L.mask_v[mword] = static_cast<uint16_t>(epi_sign_bits16(v2));
            uint4 u0, u1;
            if (vbuf) {
              epi_pack16(v2, u0, u1);
              *reinterpret_cast<uint4*>(vbuf + o0) = u0;
              *reinterpret_cast<uint4*>(vbuf + o1) = u1;
              if (mirror) {
                st_async_cluster_v4(mapa_shared(smem_u32(vbuf + o0), peer), u0, act_ready_peer);
                st_async_cluster_v4(mapa_shared(smem_u32(vbuf + o1), peer), u1, act_ready_peer);
              }
            }
            if (rbuf) {
              epi_add_bf16x16(v2, *reinterpret_cast<const uint4*>(rbuf + o0), *reinterpret_cast<const uint4*>(rbuf + o1));
              epi_pack16(v2, u0, u1);
              *reinterpret_cast<uint4*>(rbuf + o0) = u0;
              *reinterpret_cast<uint4*>(rbuf + o1) = u1;
              if (mirror) {
                st_async_cluster_v4(mapa_shared(smem_u32(rbuf + o0), peer), u0, act_ready_peer);
                st_async_cluster_v4(mapa_shared(smem_u32(rbuf + o1), peer), u1, act_ready_peer);
              }
            }
            if (obuf) {
              uint64_t o2[8];
              epi_masked16(v2, mbits, p.slope, sC + 2 * kC + c0, has_cs2, o2);
              epi_pack16(o2, u0, u1);
              *reinterpret_cast<uint4*>(obuf + o0) = u0;
              *reinterpret_cast<uint4*>(obuf + o1) = u1;
              if (mirror) {
                st_async_cluster_v4(mapa_shared(smem_u32(obuf + o0), peer), u0, act_ready_peer);
                st_async_cluster_v4(mapa_shared(smem_u32(obuf + o1), peer), u1, act_ready_peer);
              }
            }
          }
        }
        if (et < 3 * kC) sConst[((g + 1) & 1) * 3 * kC + et] = next_c;
        if (et == 0) FD_TS(11);
        // local generic-proxy smem writes -> visible to tcgen05.mma / the TMA stores (the mirrored halo rows travel
        // as st.async and are covered by the transaction count of the peer's barrier)
        fence_proxy_async();
        tc_fence_before();
        bar_sync_epi();
        if (et == 0) {
          FD_TS(13);
          const bool last_layer = (l == p.n_layers - 1);
          // The stores issued one layer ago have (long) finished reading their buffers; confirming it BEFORE the
          // arrive orders the next layer's smem writes behind it (arrive -> MMAs -> acc_full -> epilogue threads).
          tma_store_wait_read<0>();
          const uint32_t in_bytes = in_rows * 128u * static_cast<uint32_t>((vbuf != nullptr) + (rbuf != nullptr) +
                                                                           (obuf != nullptr));
          if (!last_layer) {                                     // the MMAs only read the buffers, like the stores
            if (kSplit) mbar_expect_tx(act_ready, in_bytes);     // = arrive + expect the peer's halo bytes
            else mbar_arrive(act_ready);
            FD_TS(14);
          }
          // pixel (0,0) sits one padded row + one pixel into the tile; in split mode this CTA stores its own rows
          const uint32_t src0 = static_cast<uint32_t>(p.Wp + 1) * 128u + (kSplit ? rank * 128u * 128u : 0u);
          const int h0 = kSplit ? static_cast<int>(rank) * p.rows_per_blk : 0;
          if (L.map_v >= 0) tma_store_4d(&p.maps[L.map_v], vbuf + src0, 0, 0, h0, n);
          if (L.map_res >= 0) tma_store_4d(&p.maps[L.map_res], rbuf + src0, 0, 0, h0, n);
          if (L.map_out2 >= 0) tma_store_4d(&p.maps[L.map_out2], obuf + src0, 0, 0, h0, n);
          tma_store_commit();
          if (last_layer) {
            tma_store_wait_read<0>();   // the next image reloads the buffers
            if (kSplit) mbar_expect_tx(act_ready, in_bytes);
            else mbar_arrive(act_ready);
          }
        }
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (kSplit) cluster_sync_all();     // nobody exits while the peer may still write into its shared memory
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

inline bool chain_split_ok(int H, int W) {
  const int Wp = W + 1;
  return (128 % Wp == 0) && ((H * Wp + 127) / 128 == 2);     // exactly two blocks, each a whole number of image rows
}

struct ChainBuilder {
  ChainParams p;
  int n_maps = 0;
  int rc = FD_OK;
  int B, H, W;
  int add_map(const void* ptr) {
    if (!ptr) return -1;
    if (n_maps >= kMaxMaps) { rc = FD_EUNSUPPORTED; return -1; }
    // store box = the image rows of one CTA (all H rows, or 128/Wp rows in split mode), starting at column 0:
    // only upper-bound clipping (column W, rows >= H)
    const int box_h = chain_split_ok(H, W) ? 128 / (W + 1) : H;
    int r = make_tmap_nhwc_bf16(&p.maps[n_maps], ptr, B, H, W, kC, W + 1, box_h);
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
  const size_t smem = kWBytes + static_cast<size_t>(kNumBufs) * p.buf_bytes + 6 * kC * 4 + 256 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  CUtensorMap tm_w, tm_in0, tm_in1;
  int rc = make_tmap_2d_bf16(&tm_w, w, w_layers * 9 * kC, kC, kC, kC);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_in0, in0, B, H, W, kC, Wp, H + 2);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_in1, in1 ? in1 : in0, B, H, W, kC, Wp, H + 2);
  if (rc != FD_OK) return rc;
  p.inv_wp = static_cast<uint32_t>((65536 + Wp - 1) / Wp);
  p.rows_per_blk = 128 / Wp;
  const int nsm = sm_count();
  if (chain_split_ok(H, W)) {
    cudaError_t e = set_max_dyn_smem(resblock_chain_kernel<true>, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int nclusters = B < nsm / 2 ? B : nsm / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * nclusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    e = cudaLaunchKernelEx(&cfg, resblock_chain_kernel<true>, tm_w, tm_in0, tm_in1, p);
    if (e != cudaSuccess) return static_cast<int>(e);
  } else {
    cudaError_t e = set_max_dyn_smem(resblock_chain_kernel<false>, static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    const int grid = B < nsm ? B : nsm;
    e = launch_k(resblock_chain_kernel<false>, dim3(grid), dim3(kThreads), smem, st, tm_w, tm_in0, tm_in1, p);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
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
  return kWBytes + kNumBufs * buf + 6 * kC * 4 + 256 + 1024 <= 227 * 1024 ? 1 : 0;
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
    c1.bias = bk.bias1; c1.mask_v = reinterpret_cast<uint16_t*>(bk.mask_a);
    c1.w_row = (2 * k) * 9 * kC; c1.flags = FD_EPI_LRELU;
    c1.in_buf = 0; c1.res_buf = -1; c1.v_buf = 1; c1.out2_buf = -1;
    c1.map_v = static_cast<int8_t>(cb.add_map(bk.a)); c1.map_res = -1; c1.map_out2 = -1;
    ChainLayer& c2 = p.L[2 * k + 1];
    c2 = ChainLayer{};
    c2.bias = bk.bias2; c2.chan_scale = bk.chan_scale; c2.mask_v = reinterpret_cast<uint16_t*>(bk.mask_b);
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
    l1.mask_in = reinterpret_cast<const uint16_t*>(bk.mask_a);
    l1.w_row = (2 * k + 1) * 9 * kC;
    l1.in_buf = 1; l1.res_buf = -1; l1.v_buf = -1; l1.out2_buf = 2;
    l1.map_v = -1; l1.map_res = -1; l1.map_out2 = static_cast<int8_t>(cb.add_map(bk.gp1));
    ChainLayer& l2 = p.L[2 * j + 1];     // G' = dgrad_conv1(gp1) + G ; gp2_prev = G' * drop * lrelu'(b_prev)
    l2 = ChainLayer{};
    l2.mask_in = reinterpret_cast<const uint16_t*>(bk.mask_b_prev); l2.chan_scale2 = bk.chan_scale_prev;
    l2.w_row = (2 * k) * 9 * kC;
    l2.in_buf = 2; l2.res_buf = 0; l2.v_buf = -1; l2.out2_buf = bk.gp2_prev ? 1 : -1;
    l2.map_v = -1; l2.map_res = static_cast<int8_t>(cb.add_map(bk.g_in));
    l2.map_out2 = static_cast<int8_t>(cb.add_map(bk.gp2_prev));
  }
  return launch_chain(cb, w_dgrad, 2 * n_blocks, g_out, gp2_last, B, H, W, slope, static_cast<cudaStream_t>(stream));
}
