// Stem convolution of PoolResnet (models/PoolResnet.py:70-76,98: 10x10, stride 8, pad 2, 3 -> 64)
// forward and weight gradient on tcgen05, without ever materialising an im2col matrix.
//
// Trick: with stride 8 and bf16 data, the im2col row of output column x for one (channel, ky)
// input row is the 16-element window starting at element 8x of that input row -- and 8 bf16 are
// exactly 16 bytes, the row pitch of a tcgen05 *no-swizzle core matrix* (8 rows x 16 B).  So the raw
// bf16 input row, sitting contiguously in shared memory, already IS a valid K-major A operand
// (rows = output columns, K = kx padded 10 -> 16, SBO = 128 B, LBO = 16 B: overlapping windows),
// and, read MN-major (N = kx, K = output column), a valid B operand for the weight gradient.
//
//   one task      = one output row `oy` of an image PAIR (n0, n0+1)
//   smem A tile   = [c 3][ky 10][img 2] input rows, 1024 B each (480 px + left pad 2, bf16)
//   forward       : 30 MMAs (M=128 = 2 images x 64 columns, N=64, K=16), one per (c,ky);
//                   B = weights [c,ky][kx 16][co 64] bf16 resident in smem
//   weight grad   : A = g^T (TMA tile [img,x][co], MN-major, 128B swizzle), B = the same input
//                   rows (MN-major, no swizzle); 30 accumulators M=64(co) x N=16(kx) stay in TMEM
//                   (480 columns) across all tasks of the persistent CTA.
// Both kernels are bound by reading the fp32 images once from HBM (2.76 MB/image); the loader warps
// convert fp32 (or uint8 / 255) to bf16 on the way into shared memory.
#include <cstdlib>

#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kCo = 64;
constexpr int kRowBytes = 1024;            // one input row in smem: 512 bf16
constexpr int kLoaderWarps = 8;             // converter warps: fp32/u8 staging -> bf16 A tile
constexpr int kMmaWarp = kLoaderWarps;      // MMA issuer + TMEM owner
constexpr int kTmaWarp = kLoaderWarps + 1;  // TMA producer of the image chunks
constexpr int kEpiWarp0 = kLoaderWarps + 2; // 4 epilogue warps (10..13 -> TMEM quadrants 2,3,0,1)
constexpr int kThreads = (kLoaderWarps + 2 + 4) * 32;   // 448
// Converter warps of the forward.  Measured (FD_STEM_TIMING=1, tools/stem_debug.py) history of this kernel:
//  * the converters set the pace, not HBM: ~1400-2800 clk per chunk (TMA wait, dependent LDS -> F2FP -> STS/STG
//    chains, hand-over of the A rows), 3 chunks per warp with 8 warps = 7500 clk per task = 74 us;
//  * every converter warp now issues the TMA for ITS OWN slot right after it has pulled the slot's data into
//    registers (program order = the hand-shake: no "empty" barrier, no producer warp whose per-lane issues
//    serialise), division-free;
//  * 16 converter warps, one slot each (slot = chunk % 16 = warp; warps 0..7 drain chunks j and j + 16 of a task).
//    A slot must always be drained by the SAME warp: mbarrier parity waits only distinguish adjacent phases
//    (24 warps sharing 16 slots: launch failure).
constexpr int kFwdLoaders = 16;
constexpr int kFwdMmaWarp = kFwdLoaders;
constexpr int kFwdTmaWarp = kFwdLoaders + 1;
constexpr int kFwdEpiWarp0 = kFwdLoaders + 2;           // 18..21 -> TMEM quadrants 2,3,0,1
constexpr int kFwdThreads = (kFwdLoaders + 2 + 4) * 32; // 704
constexpr int kSlotBytes = 4864;            // one staging slot: K/2 rows x Win/2 pixels (<= 4800 B), 128B aligned
constexpr int kChunksPerTask = 24;          // 3 ch x 2 img x 2 row groups x 2 column halves
constexpr int kKH = 5;                      // rows per chunk  (K = 10)
constexpr int kQPR = 60;                    // 4-pixel quads per half row (Win = 480)
constexpr int kMaxCK = 30;                 // Cin * K rows per image per task

struct StemParams {
  int B, Cin, Hin, Win, K, stride, pad, Ho, Wo;
  int CK;           // Cin * K
  int npairs;       // ceil(B / 2)
  int ntask;        // npairs * Ho
  uint32_t a_bytes; // CK * 2 * 1024
  int elem_bytes;   // 4 (fp32 image) or 1 (uint8 image)
  int dbg;
  const void* x;
  const float* w;        // [64][Cin][K][K] fp32 (forward)
  const float* bias;     // [64]
  __nv_bfloat16* y;      // [B,Ho,Wo,64]
  float* dw;             // [64][Cin][K][K] fp32, accumulated (wgrad)
  float* dbias;          // [64] accumulated (wgrad)
  // bf16 copy of the images in the operand layout of the A tile: [pair][c][y][img][512] (element e = column e - pad,
  // pad columns zero).  Written by the forward converter warps, TMA-loaded by the weight-gradient kernel, so the
  // fp32 images are read from HBM once per step and converted once.
  __nv_bfloat16* xbf;
};

// no-swizzle descriptor (layout type 0)
__device__ __forceinline__ uint64_t make_sdesc_none(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <typename TIn>
struct Px4;
template <>
struct Px4<float> {
  static __device__ __forceinline__ void load(const uint8_t* slot, int quad, float (&v)[4]) {
    const float4 f = *reinterpret_cast<const float4*>(slot + quad * 16);
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
};
template <>
struct Px4<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* slot, int quad, float (&v)[4]) {
    const uchar4 u = *reinterpret_cast<const uchar4*>(slot + quad * 4);
    v[0] = div255(u.x); v[1] = div255(u.y); v[2] = div255(u.z); v[3] = div255(u.w);         // PoolResnet.py:95
  }
  // x / 255.0f, correctly rounded, without the ~12-instruction IEEE division: q = x * (1/255), one FMA residual and
  // one FMA correction.  Bit-identical to the division for all 256 inputs (tests/test_host_cpu.py checks the table).
  static __device__ __forceinline__ float div255(unsigned char b) {
    const float x = static_cast<float>(b), r = 1.0f / 255.0f;
    const float q = x * r;
    return fmaf(fmaf(-q, 255.0f, x), r, q);
  }
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// The image reaches shared memory through a ring of 8 TMA staging slots (one per converter warp):
// chunk j of a task = (channel c, image img, row group rg, column half) -> one box {Win/2, K/2, 1} of
// the [B*Cin, Hin, Win] input, zero filled above the image.  Converter warp (j % 8) turns it into
// bf16 rows of the A tile.  Element e of an A row holds input column e - pad.
// Optional timestamps of CTA 0 (FD_STEM_TIMING=1): [task iteration][16] clock64 values.
__device__ unsigned long long g_stem_dbg[32 * 16];
#define STEM_TS(slot) do { if (p.dbg && blockIdx.x == 0 && it < 32) g_stem_dbg[it * 16 + (slot)] = clock64(); } while (0)

struct ChunkCoord {
  int c, img, rg, half;
};
__device__ __forceinline__ ChunkCoord chunk_coord(int j) {
  ChunkCoord k;
  k.half = j & 1; k.rg = (j >> 1) & 1; k.img = (j >> 2) & 1; k.c = j >> 3;
  return k;
}

// TMA producer warp: lane j issues chunk j of every task (24 lanes busy), so the ~100-cycle mbarrier
// round trips of the slot hand-shake overlap across lanes instead of serialising in one thread.
template <int NS, typename Sched>
__device__ __forceinline__ void produce_chunks(const StemParams& p, const CUtensorMap* tm_x, uint8_t* sStage,
                                               uint64_t* stg_full, uint64_t* stg_empty, Sched sched, int lane) {
  const uint32_t bytes = static_cast<uint32_t>((p.K / 2) * (p.Win / 2) * p.elem_bytes);
  int it = 0;
  for (int task = sched.begin; task < sched.end; task += sched.step, ++it) {
    const int pair = task / p.Ho, oy = task % p.Ho;
    if (lane == 0) STEM_TS(10);
    // batches of NS chunks: inside a batch every lane owns a different slot, so no lane ever waits
    // for a slot that a sibling lane of the same batch still has to fill (that would deadlock the warp)
#pragma unroll 1
    for (int j0 = 0; j0 < kChunksPerTask; j0 += NS) {
      const int j = j0 + lane;
      if (lane < NS && j < kChunksPerTask) {
        const ChunkCoord k = chunk_coord(j);
        const int g = it * kChunksPerTask + j;                   // chunk sequence number of this CTA
        const int slot = g % NS;
        const int use = g / NS;                                  // how many times this slot was used before
        mbar_wait(stg_empty + slot, (use & 1) ^ 1);
        mbar_expect_tx(stg_full + slot, bytes);
        tma_load_3d(sStage + slot * kSlotBytes, tm_x, stg_full + slot, k.half * (p.Win / 2),
                    oy * p.stride - p.pad + k.rg * (p.K / 2), (pair * 2 + k.img) * p.Cin + k.c);
      }
      __syncwarp();
      if (lane == 0) STEM_TS(j0 == 0 ? 11 : 12);
    }
  }
}

// Self-service TMA of the forward: chunk j of the task (pair, oy) into staging slot `slot`.  Called by ONE elected
// lane of the converter warp that owns the slot, after it has emptied the slot.  No divisions on this path: it sits in
// the converters' critical loop (measured 550 clk per issue with the chunk number decoded by / and %).
__device__ __forceinline__ void issue_chunk(const StemParams& p, const CUtensorMap* tm_x, uint8_t* sStage,
                                            uint64_t* stg_full, int j, int pair, int oy, int slot) {
  const ChunkCoord k = chunk_coord(j);
  mbar_expect_tx(stg_full + slot, static_cast<uint32_t>((p.K / 2) * (p.Win / 2) * p.elem_bytes));
  tma_load_3d(sStage + slot * kSlotBytes, tm_x, stg_full + slot, k.half * (p.Win / 2),
              oy * p.stride - p.pad + k.rg * (p.K / 2), (pair * 2 + k.img) * p.Cin + k.c);
}

// Converter warp `warp`: its chunks of task iteration `it` -> A tile `abuf`.
template <typename TIn, int NS, int NW = kLoaderWarps>
__device__ __forceinline__ void convert_chunks(const StemParams& p, uint8_t* abuf, const uint8_t* sStage,
                                               uint64_t* stg_full, uint64_t* stg_empty, int it, int warp, int lane,
                                               int task = 0, uint64_t* grp_full = nullptr, uint64_t* grp_empty = nullptr,
                                               const CUtensorMap* self_tm = nullptr, int next_pair = -1, int next_oy = 0) {
  const int task_pair = task / p.Ho, task_oy = task - task_pair * p.Ho;
  // NS == NW == 16 (forward): slot = chunk % 16 = the owning warp, warps 0..7 drain two chunks per task (j, j + 16)
  constexpr bool kOwnSlot = (NS == 16 && NW == 16);
#pragma unroll 1
  for (int j = warp; j < kChunksPerTask; j += NW) {
    const int jj = j / NW;
    const int g = it * kChunksPerTask + j;
    const int sl = kOwnSlot ? (j & 15) : g % NS;
    const int use = kOwnSlot ? ((j & 15) < 8 ? 2 * it + (j >> 4) : it) : g / NS;
    const ChunkCoord k = chunk_coord(j);
    if (warp == 0 && lane == 0 && jj == 0) STEM_TS(0);
    if (warp == 0 && lane == 0 && jj == 1) STEM_TS(15);
    mbar_wait(stg_full + sl, use & 1);
    if (warp == 0 && lane == 0 && jj == 0) STEM_TS(1);
    const uint8_t* slot = sStage + sl * kSlotBytes;
    // kKH rows x kQPR quads, fully unrolled: all shared loads are issued before the first convert
    float v[kKH][2][4];
#pragma unroll
    for (int r = 0; r < kKH; ++r) {
      Px4<TIn>::load(slot, r * kQPR + lane, v[r][0]);
      if (lane + 32 < kQPR) Px4<TIn>::load(slot, r * kQPR + lane + 32, v[r][1]);
    }
    // The staging slot is free as soon as its data sits in registers: release it NOW so that the producer's next TMA
    // into this slot overlaps the conversion and the stores below (released at the end of the chunk, the HBM round
    // trip of the next chunk was exposed).  The MOVs make every lane wait for its last (in-order) shared load.
    {
      uint32_t sink;
      asm volatile("mov.b32 %0, %1;" : "=r"(sink) : "r"(__float_as_uint(v[kKH - 1][0][3])) : "memory");
      if (lane + 32 < kQPR) asm volatile("mov.b32 %0, %1;" : "=r"(sink) : "r"(__float_as_uint(v[kKH - 1][1][3])) : "memory");
      __syncwarp();
      if (warp == 0 && lane == 0 && jj == 1) STEM_TS(10);
      if (self_tm != nullptr) {
        // the next user of this slot is chunk j + NS of this task, or chunk j + NS - 24 of the CTA's next task
        if (elect_one_sync()) {
          if (j + NS < kChunksPerTask) {
            issue_chunk(p, self_tm, const_cast<uint8_t*>(sStage), stg_full, j + NS, task_pair, task_oy, sl);
          } else if (next_pair >= 0) {
            issue_chunk(p, self_tm, const_cast<uint8_t*>(sStage), stg_full, kOwnSlot ? (j & 15) : j + NS - kChunksPerTask,
                        next_pair, next_oy, sl);
          }
        }
        __syncwarp();
      } else if (lane == 0) {
        mbar_arrive(stg_empty + sl);
      }
      if (warp == 0 && lane == 0 && jj == 1) STEM_TS(11);
    }
    uint8_t* dst0 = abuf + ((k.c * p.K + k.rg * kKH) * 2 + k.img) * kRowBytes + (k.half * (p.Win / 2) + p.pad) * 2;
    // Row-group hand-off (forward): the 5 A rows of group (c, rg) may be overwritten as soon as the 5 MMAs of the
    // PREVIOUS task that read them have completed -- not only after all 30.
    if (grp_empty != nullptr) mbar_wait(grp_empty + k.c * 2 + k.rg, (it & 1) ^ 1);
    if (warp == 0 && lane == 0 && jj == 0) STEM_TS(2);
    if (warp == 0 && lane == 0 && jj == 1) STEM_TS(14);
    // global bf16 copy (forward only): row y of plane (pair, c), image slot img
    const int pair = task_pair, oy = task_oy;
    const int y0 = oy * p.stride - p.pad + k.rg * kKH;
    uint8_t* gdst0 = p.xbf == nullptr ? nullptr
                                      : reinterpret_cast<uint8_t*>(p.xbf) +
                                            ((static_cast<size_t>(pair) * p.Cin + k.c) * p.Hin * 2 + k.img) * kRowBytes +
                                            (k.half * (p.Win / 2) + p.pad) * 2 + lane * 8;
#pragma unroll
    for (int r = 0; r < kKH; ++r) {
      uint32_t* d = reinterpret_cast<uint32_t*>(dst0 + r * 2 * kRowBytes + lane * 8);
      const uint32_t a0 = pack_bf16x2(v[r][0][0], v[r][0][1]), a1 = pack_bf16x2(v[r][0][2], v[r][0][3]);
      d[0] = a0;
      d[1] = a1;
      uint32_t b0 = 0, b1 = 0;
      if (lane + 32 < kQPR) {
        b0 = pack_bf16x2(v[r][1][0], v[r][1][1]);
        b1 = pack_bf16x2(v[r][1][2], v[r][1][3]);
        d[64] = b0;
        d[65] = b1;
      }
      const int y = y0 + r;
      if (gdst0 != nullptr && y >= 0 && y < p.Hin) {
        uint8_t* grow = gdst0 + static_cast<size_t>(y) * 2 * kRowBytes;
        uint32_t* gw = reinterpret_cast<uint32_t*>(grow);      // rows start 2 pad elements in: 4-byte aligned only
        gw[0] = a0;
        gw[1] = a1;
        if (lane + 32 < kQPR) {
          gw[64] = b0;
          gw[65] = b1;
        }
      }
    }
    if (warp == 0 && lane == 0 && jj == 1) STEM_TS(12);
    if (grp_full != nullptr) fence_proxy_async();     // A rows (generic proxy) -> visible to the tensor core
    if (warp == 0 && lane == 0 && jj == 1) STEM_TS(13);
    __syncwarp();
    if (lane == 0 && grp_full != nullptr) mbar_arrive(grp_full + k.c * 2 + k.rg);
    if (warp == 0 && lane == 0) STEM_TS(jj == 0 ? 3 : 4);
  }
}

struct StrideSched { int begin, end, step; };

// bf16 B operand of the forward in shared memory, per (c,ky): [k1 2][co 64][k0 8] (K-major, no swizzle: LBO = 1024,
// SBO = 128; kx = 8 k1 + k0, zero for kx >= K) from w[co][c][ky][kx] fp32.  Reads are coalesced float4s in SOURCE order
// and scattered into the operand: the element-per-thread gather it replaces (stride K*CK floats between lanes) cost
// ~45 dependent L2 round trips per thread in the prologue of every stem launch.
__device__ __forceinline__ void build_stem_weights(uint8_t* sW, const StemParams& p, int tid, int nthreads) {
  __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(sW);
  const int KK = p.CK * p.K;
  if ((reinterpret_cast<uintptr_t>(p.w) & 15u) != 0 || ((kCo * KK) & 3) != 0) {       // unaligned view: plain gather
    for (int i = tid; i < p.CK * 1024; i += nthreads) {
      const int k0 = i & 7, co = (i >> 3) & 63, k1 = (i >> 9) & 1, cky = i >> 10;
      const int kx = k1 * 8 + k0;
      w16[i] = __float2bfloat16(kx < p.K ? p.w[static_cast<size_t>(co) * KK + cky * p.K + kx] : 0.f);
    }
    return;
  }
  for (int i = tid; i < p.CK * 512; i += nthreads)      // the padding kx >= K: k1 = 1, k0 >= K - 8 (disjoint from the scatter)
    if (8 + (i & 7) >= p.K) w16[(i >> 9) * 1024 + 512 + (i & 511)] = __float2bfloat16(0.f);
  const float4* w4 = reinterpret_cast<const float4*>(p.w);
  const int n4 = (kCo * KK) >> 2;
#pragma unroll 4
  for (int f = tid; f < n4; f += nthreads) {
    const float4 v = __ldg(w4 + f);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const int e = f * 4;
    int co = e / KK;
    const int rem = e - co * KK;
    int cky = rem / p.K, kx = rem - cky * p.K;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      w16[cky * 1024 + (kx >> 3) * 512 + co * 8 + (kx & 7)] = __float2bfloat16(vv[j]);
      if (++kx == p.K) {
        kx = 0;
        if (++cky == p.CK) { cky = 0; ++co; }
      }
    }
  }
}


// ------------------------------------------------------------------------------------- forward
// smem: [weights CK*2048][A tile (single stage)][slack 128][16 staging slots][barriers]
constexpr int kFwdSlots = 16;               // one per converter warp
constexpr int kWgSlots = 8;
template <typename TIn>
__global__ void __launch_bounds__(kFwdThreads, 1)
stem_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = sW + p.CK * 2048;
  uint8_t* sStage = sA + p.a_bytes + 128;              // staging slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + kFwdSlots * kSlotBytes);
  // bars + 0..3: unused (the A tile is handed over per row group, see grp_full / grp_empty)
  uint64_t* acc_full = bars + 4;   // [2]
  uint64_t* acc_empty = bars + 6;  // [2]
  uint64_t* stg_full = bars + 8;               // [kFwdSlots]
  uint64_t* stg_empty = bars + 8 + kFwdSlots;  // [kFwdSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kFwdSlots);
  // The A tile is handed over per ROW GROUP g = (c, rg) = 5 of the 30 (c,ky) rows x 2 images (4 chunks, 4 converter
  // warps): grp_full[g] (count 4) releases the 5 MMAs of the group, grp_empty[g] (their commit) lets the converters
  // refill those rows for the next task.  With one barrier for the whole tile the converters idled during the 30
  // MMAs, the staging ring filled up, and every task exposed one HBM round trip (11 k clk per task, 74 us).
  uint64_t* grp_full = bars + 8 + 2 * kFwdSlots + 2;     // [6]
  uint64_t* grp_empty = grp_full + 6;                    // [6]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one-time: zero the A stages (pad columns stay zero forever) and build the bf16 weight operand
  for (uint32_t i = threadIdx.x * 16u; i < p.a_bytes + 128; i += kFwdThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  build_stem_weights(sW, p, threadIdx.x, kFwdThreads);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    for (int g = 0; g < 6; ++g) {
      mbar_init(grp_full + g, 4);
      mbar_init(grp_empty + g, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 4);
    }
    for (int s = 0; s < kFwdSlots; ++s) {
      mbar_init(stg_full + s, 1);
      mbar_init(stg_empty + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == kFwdMmaWarp) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp < kFwdLoaders) {
    // prologue: this warp's first two chunks (slots warp and warp + 8); afterwards every chunk's successor in the same
    // slot is issued by convert_chunks itself
    if (lane == 0 && static_cast<int>(blockIdx.x) < p.ntask) {
      const int pair0 = blockIdx.x / p.Ho, oy0 = blockIdx.x - pair0 * p.Ho;
      issue_chunk(p, &tm_x, sStage, stg_full, warp, pair0, oy0, warp);        // one slot per warp: its first chunk
    }
    __syncwarp();
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      // single A stage, handed over per row group (see grp_full / grp_empty above)
      const int nt = task + gridDim.x;
      const int np = nt < p.ntask ? nt / p.Ho : -1;
      convert_chunks<TIn, kFwdSlots, kFwdLoaders>(p, sA, sStage, stg_full, stg_empty, it, warp, lane, task, grp_full,
                                                  grp_empty, &tm_x, np, nt - np * p.Ho);
    }
  } else if (warp == kFwdTmaWarp) {
    // (idle: the converter warps issue their own TMA loads)
  } else if (warp == kFwdMmaWarp) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kCo, 0, 0);
    const uint32_t w_addr = smem_u32(sW);
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc_empty + s, ph ^ 1);
      tc_fence_after();
      if (elect_one_sync()) {
        STEM_TS(5);
        const uint64_t ad0 = make_sdesc_none(smem_u32(sA), 16, 128);
        const uint64_t bd0 = make_sdesc_none(w_addr, 1024, 128);
#pragma unroll 1
        for (int g = 0; g < 6; ++g) {
          mbar_wait(grp_full + g, it & 1);
          if (g == 0) STEM_TS(6);
          tc_fence_after();
#pragma unroll
          for (int r = 0; r < kKH; ++r) {       // +2048 B per (c,ky) on both operands = +128 in the address field
            const int cky = g * kKH + r;
            umma_bf16(tmem_base + s * kCo, ad0 + static_cast<uint64_t>(cky * 128), bd0 + static_cast<uint64_t>(cky * 128),
                      idesc, cky != 0 ? 1u : 0u);
          }
          umma_commit(grp_empty + g);            // these 5 A rows may be refilled
        }
        umma_commit(acc_full + s);
        STEM_TS(7);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int pair = task / p.Ho, oy = task % p.Ho;
      mbar_wait(acc_full + s, ph);
      if (q == 0 && lane == 0) STEM_TS(8);
      tc_fence_after();
      const int r = q * 32 + lane;
      const int img = r >> 6, ox = r & 63;
      const int n = pair * 2 + img;
      const bool valid = ox < p.Wo && n < p.B;
      __nv_bfloat16* dst = p.y + ((static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox) * kCo;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * kCo + half * 32, acc);
        tmem_ld_wait();
        if (valid) {
          uint4* d = reinterpret_cast<uint4*>(dst + half * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = half * 32 + i * 8;
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(acc[8 * i + 0]) + __ldg(p.bias + c0 + 0),
                              __uint_as_float(acc[8 * i + 1]) + __ldg(p.bias + c0 + 1));
            u.y = pack_bf16x2(__uint_as_float(acc[8 * i + 2]) + __ldg(p.bias + c0 + 2),
                              __uint_as_float(acc[8 * i + 3]) + __ldg(p.bias + c0 + 3));
            u.z = pack_bf16x2(__uint_as_float(acc[8 * i + 4]) + __ldg(p.bias + c0 + 4),
                              __uint_as_float(acc[8 * i + 5]) + __ldg(p.bias + c0 + 5));
            u.w = pack_bf16x2(__uint_as_float(acc[8 * i + 6]) + __ldg(p.bias + c0 + 6),
                              __uint_as_float(acc[8 * i + 7]) + __ldg(p.bias + c0 + 7));
            d[i] = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + s);
      if (q == 0 && lane == 0) STEM_TS(9);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFwdMmaWarp) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------- weight gradient
// smem: [A stage 0][A stage 1][slack 1024][g stage 0 (16 KB)][g stage 1][barriers][bias scratch]
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1)
stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                     const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sG = sA + 2 * p.a_bytes + 1024;
  constexpr uint32_t kGBytes = 128 * 128;
  uint8_t* sStage = sG + 2 * kGBytes;                  // 8 staging slots
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 8 * kSlotBytes);
  uint64_t* full = bars + 0;     // [2] count = loader warps + 1 (TMA expect_tx)
  uint64_t* empty = bars + 2;    // [2] count = 1 (MMA commit) + 4 (bias warps)
  uint64_t* acc_full = bars + 4;
  uint64_t* stg_full = bars + 5;   // [8]
  uint64_t* stg_empty = bars + 13; // [8]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  float* sBias = reinterpret_cast<float*>(bars + 22);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x * 16u; i < 2 * p.a_bytes + 1024 + 2 * kGBytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    tma_prefetch_desc(&tm_x);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, kLoaderWarps + 1);
      mbar_init(empty + s, 1 + 4);
    }
    mbar_init(acc_full, 1);
    for (int s = 0; s < 8; ++s) {
      mbar_init(stg_full + s, 1);
      mbar_init(stg_empty + s, 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  // contiguous chunk of tasks per CTA
  const int per = (p.ntask + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(p.ntask, t_begin + per);

  if (warp < kLoaderWarps) {
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(empty + s, ph ^ 1);
      convert_chunks<TIn, kWgSlots>(p, sA + s * p.a_bytes, sStage, stg_full, stg_empty, it, warp, lane);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
    }
  } else if (warp == kTmaWarp) {
    produce_chunks<kWgSlots>(p, &tm_x, sStage, stg_full, stg_empty, StrideSched{t_begin, t_end, 1}, lane);
  } else if (warp == kMmaWarp) {
    constexpr uint32_t idesc = make_idesc_bf16(64, 16, 1, 1);
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      if (elect_one_sync()) {
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, 2u * p.Wo * 128u);
        // one box {64 ch, Wo, 1, 1} per image into 64-row slots; rows Wo..63 stay zero; n >= B is zero filled
        tma_load_4d(sG + s * kGBytes, &tm_g, full + s, 0, 0, task % p.Ho, (task / p.Ho) * 2);
        tma_load_4d(sG + s * kGBytes + 64 * 128, &tm_g, full + s, 0, 0, task % p.Ho, (task / p.Ho) * 2 + 1);
      }
      __syncwarp();
      mbar_wait(full + s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        // A = g^T: MN-major (M = co), 128B swizzle, 16 K-rows (positions) per MMA
        const uint64_t ad0 = make_sdesc_sw128(smem_u32(sG + s * kGBytes), 1024, 1024, 0);
        // B = input windows: MN-major (N = kx), no swizzle: kx chunks 16 B apart, 8-position groups 128 B apart
        const uint64_t bd0 = make_sdesc_none(smem_u32(sA + s * p.a_bytes), 128, 16);
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = ad0 + static_cast<uint64_t>(ks * 128);          // +2048 B
          const uint64_t bdk = bd0 + static_cast<uint64_t>(ks * 16);          // +256 B
#pragma unroll 6
          for (int cky = 0; cky < p.CK; ++cky)
            umma_bf16(tmem_base + cky * 16, ad, bdk + static_cast<uint64_t>(cky * 128), idesc, (it | ks) != 0 ? 1u : 0u);
        }
        umma_commit(empty + s);
        if (task + 1 == t_end) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int et = threadIdx.x - kEpiWarp0 * 32;   // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum = 0.f;
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      if (p.dbias && !(p.dbg & 2)) {
        const uint8_t* g = sG + s * kGBytes;
        for (int r = rpar; r < 128; r += 2) {
          const uint32_t off = r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
          bsum += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(g + off));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    if (t_begin < t_end) {
      const int q = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // M = 64 accumulator layout: row co lives in lane (co % 16) + 32 * (co / 16)
      const int co = q * 16 + lane;
      const int KK = p.CK * p.K;
      for (int cky = 0; cky < p.CK; ++cky) {
        uint32_t acc[16];
        tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cky * 16, acc);
        tmem_ld_wait();
        if (lane < 16) {
          float* dst = p.dw + static_cast<size_t>(co) * KK + cky * p.K;
#pragma unroll
          for (int kx = 0; kx < 16; ++kx)
            if (kx < p.K) atomicAdd(dst + kx, __uint_as_float(acc[kx]));
        }
      }
      if (p.dbias) {
        sBias[et] = bsum;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et < 64) atomicAdd(p.dbias + et, sBias[et] + sBias[et + 64]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------------------------- weight gradient from the bf16 copy
// Same MMAs as stem_wgrad_tc_kernel, but the A-tile rows arrive as three TMA boxes per task ({512 e, 2 img, K rows} of
// plane (pair, c)) straight from the bf16 image copy the forward wrote: no fp32 re-read, no converter warps, no
// staging ring.  Warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner, 2..5 = bias sums + drain.
constexpr int kWg2Threads = 6 * 32;
// kNP = 2: the weight gradients of TWO 64-channel output planes (the stem of a 128-filter model) from one pass over the
// image rows -- a second gradient tile per stage, 8 accumulators (all 512 TMEM columns); dw / dbias of plane 1 follow
// those of plane 0 (rows 64..127 of the [128][Cin][K][K] gradient).
template <int kNP>
__global__ void __launch_bounds__(kWg2Threads, 1)
stem_wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tm_xb, const __grid_constant__ CUtensorMap tm_g,
                       const __grid_constant__ CUtensorMap tm_g1, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  constexpr uint32_t kGBytes = 128 * 128;
  uint8_t* sA = smem;                                   // 2 stages x a_bytes
  uint8_t* sG = sA + 2 * p.a_bytes + 1024;              // 2 stages x kNP planes x 16 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * kNP * kGBytes);
  uint64_t* full = bars + 0;     // [2]
  uint64_t* empty = bars + 2;    // [2] count = 1 (MMA commit) + 4 (bias warps)
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sBias = reinterpret_cast<float*>(bars + 6);    // [kNP][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // g rows Wo..63 of each 64-row slot and the slack behind the A stages are never written by TMA: keep them zero
  for (uint32_t i = threadIdx.x * 16u; i < 2 * p.a_bytes + 1024 + 2 * kNP * kGBytes; i += kWg2Threads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    if (kNP > 1) tma_prefetch_desc(&tm_g1);
    tma_prefetch_desc(&tm_xb);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1 + 4);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int per = (p.ntask + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(p.ntask, t_begin + per);

  if (warp == 0) {
    if (elect_one_sync()) {
      int it = 0;
      for (int task = t_begin; task < t_end; ++task, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int pair = task / p.Ho, oy = task - pair * p.Ho;
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, ((p.dbg & 8) ? 0u : p.a_bytes) + kNP * 2u * p.Wo * 128u);
        for (int c = 0; c < ((p.dbg & 8) ? 0 : p.Cin); ++c)       // {256 e, 2 halves, 2 img, K rows, 1 plane}; rows above the image: zero fill
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
              "[%2];" ::"r"(smem_u32(sA + s * p.a_bytes + c * p.K * 2 * kRowBytes)),
              "l"(reinterpret_cast<uint64_t>(&tm_xb)), "r"(smem_u32(full + s)), "r"(0), "r"(0), "r"(0),
              "r"(oy * p.stride - p.pad), "r"(pair * p.Cin + c)
              : "memory");
#pragma unroll
        for (int pl = 0; pl < kNP; ++pl) {
          const CUtensorMap* tg = pl == 0 ? &tm_g : &tm_g1;
          tma_load_4d(sG + (s * kNP + pl) * kGBytes, tg, full + s, 0, 0, oy, pair * 2);
          tma_load_4d(sG + (s * kNP + pl) * kGBytes + 64 * 128, tg, full + s, 0, 0, oy, pair * 2 + 1);
        }
      }
    }
  } else if (warp == 1) {
    // D[(c,ky,kx), co] = sum_pos win[pos][(c,ky,kx)] * g[pos][co] with M = 128 = 16 (c,ky) rows x 8 kx:
    //   A (MN-major, no swizzle): the raw bf16 rows again -- 8-element MN groups are the rows (c,ky) of ONE image
    //     slot, 2048 B apart (SBO); the K index (position = img*64 + ox) advances 16 B, 8-position groups 128 B (LBO);
    //     kx 8..15 is the same operand started 16 B further;
    //   B (MN-major, 128B swizzle): the gradient tile [position][co], N = 64.
    // 2 row blocks x 2 kx chunks = 4 MMAs (128x64x16) per 16 positions instead of 30 MMAs (64x16x16).
    constexpr uint32_t idesc = make_idesc_bf16(128, kCo, 1, 1);
    if (elect_one_sync()) {
      int it = 0;
      for (int task = t_begin; task < t_end; ++task, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint64_t bd0 = make_sdesc_sw128(smem_u32(sG + s * kNP * kGBytes), 1024, 1024, 0);
        const uint32_t a_addr = smem_u32(sA + s * p.a_bytes);
#pragma unroll 1
        for (int ks = 0; ks < ((p.dbg & 1) ? 0 : 8); ++ks) {
          const uint64_t bd = bd0 + static_cast<uint64_t>(ks * 128);          // +2048 B: 16 positions
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int blk = a >> 1, chunk = a & 1;
            const uint64_t ad = make_sdesc_none(a_addr + blk * 16 * 2 * kRowBytes + chunk * 16 + ks * 256, 128,
                                                2 * kRowBytes);
#pragma unroll
            for (int pl = 0; pl < kNP; ++pl)       // + 16 KB (address field + 1024): the gradient tile of plane pl
              umma_bf16(tmem_base + (pl * 4 + a) * kCo, ad, bd + static_cast<uint64_t>(pl * (kGBytes >> 4)), idesc,
                        (it | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty + s);
        if (task + 1 == t_end) umma_commit(acc_full);
      }
    }
    __syncwarp();
  } else {
    const int et = threadIdx.x - 64;   // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum[kNP];
#pragma unroll
    for (int pl = 0; pl < kNP; ++pl) bsum[pl] = 0.f;
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      if (p.dbias && !(p.dbg & 2)) {
        const uint8_t* g = sG + s * kNP * kGBytes;
        for (int r = rpar; r < 128; r += 2) {
          const uint32_t off = r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1));
#pragma unroll
          for (int pl = 0; pl < kNP; ++pl)
            bsum[pl] += __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(g + pl * kGBytes + off));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    if (t_begin < t_end) {
      const int q = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // accumulator a = (row block, kx chunk); TMEM lane r = 8 * (row within block) + (kx within chunk)
      const int r = q * 32 + lane;
      const int KK = p.CK * p.K;
#pragma unroll 1
      for (int pa = 0; pa < 4 * kNP; ++pa) {
        const int a = pa & 3, pl = pa >> 2;
        const int cky = (a >> 1) * 16 + (r >> 3), kx = (a & 1) * 8 + (r & 7);
        const bool valid = cky < p.CK && kx < p.K;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + pa * kCo + half * 32, acc);
          tmem_ld_wait();
          if (valid && !(p.dbg & 4)) {
            float* dst = p.dw + static_cast<size_t>(pl * kCo + half * 32) * KK + cky * p.K + kx;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + static_cast<size_t>(j) * KK, __uint_as_float(acc[j]));
          }
        }
      }
      if (p.dbias) {
#pragma unroll
        for (int pl = 0; pl < kNP; ++pl) sBias[pl * 128 + et] = bsum[pl];
        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
        for (int pl = 0; pl < kNP; ++pl)
          if (et < 64) atomicAdd(p.dbias + pl * kCo + et, sBias[pl * 128 + et] + sBias[pl * 128 + et + 64]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------- forward from the bf16 copy
// The stem of a model wider than 64 filters runs once per 64-channel output plane.  The first plane's launch
// (stem_fwd_tc_kernel) reads the fp32 / uint8 images and leaves the bf16 copy behind; every further plane takes its A
// tile from that copy -- three TMA boxes per task, the loads of the bf16 weight-gradient kernel, and the 30 MMAs and the
// epilogue of the forward kernel: 2 B instead of 4 B per pixel from HBM, no converter warps.
// smem: [weights CK*2048][A stage 0][A stage 1][barriers].  Warps: 0 = TMA producer, 1 = MMA issuer + TMEM owner,
// 2..5 = epilogue.
__global__ void __launch_bounds__(kWg2Threads, 1)
stem_fwd_bf16_kernel(const __grid_constant__ CUtensorMap tm_xb, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = sW + p.CK * 2048;                        // 2 stages x a_bytes
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * p.a_bytes + 128);
  uint64_t* full = bars + 0;       // [2]
  uint64_t* empty = bars + 2;      // [2]
  uint64_t* acc_full = bars + 4;   // [2]
  uint64_t* acc_empty = bars + 6;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (uint32_t i = threadIdx.x * 16u; i < 2 * p.a_bytes + 128; i += kWg2Threads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  build_stem_weights(sW, p, threadIdx.x, kWg2Threads);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xb);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (warp == 0) {
    if (elect_one_sync()) {
      int it = 0;
      for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
        const int s = it & 1, ph = (it >> 1) & 1;
        const int pair = task / p.Ho, oy = task - pair * p.Ho;
        mbar_wait(empty + s, ph ^ 1);
        if (p.dbg & 8) { mbar_arrive(full + s); continue; }       // timing switch: no loads
        mbar_expect_tx(full + s, p.a_bytes);
        for (int c = 0; c < p.Cin; ++c)       // {256 e, 2 halves, 2 img, K rows, 1 plane}; rows outside the image: zero fill
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], "
              "[%2];" ::"r"(smem_u32(sA + s * p.a_bytes + c * p.K * 2 * kRowBytes)),
              "l"(reinterpret_cast<uint64_t>(&tm_xb)), "r"(smem_u32(full + s)), "r"(0), "r"(0), "r"(0),
              "r"(oy * p.stride - p.pad), "r"(pair * p.Cin + c)
              : "memory");
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kCo, 0, 0);
    const uint64_t bd0 = make_sdesc_none(smem_u32(sW), 1024, 128);
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc_empty + s, ph ^ 1);
      mbar_wait(full + s, ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ad0 = make_sdesc_none(smem_u32(sA + s * p.a_bytes), 16, 128);
#pragma unroll 2
        for (int cky = 0; cky < ((p.dbg & 1) ? 0 : p.CK); ++cky)    // +2048 B per (c,ky) on both operands = +128 in the address field
          umma_bf16(tmem_base + s * kCo, ad0 + static_cast<uint64_t>(cky * 128), bd0 + static_cast<uint64_t>(cky * 128),
                    idesc, cky != 0 ? 1u : 0u);
        umma_commit(empty + s);
        umma_commit(acc_full + s);
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      const int pair = task / p.Ho, oy = task - pair * p.Ho;
      mbar_wait(acc_full + s, ph);
      tc_fence_after();
      const int r = q * 32 + lane;
      const int img = r >> 6, ox = r & 63;
      const int n = pair * 2 + img;
      const bool valid = ox < p.Wo && n < p.B;
      __nv_bfloat16* dst = p.y + ((static_cast<size_t>(n) * p.Ho + oy) * p.Wo + ox) * kCo;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + s * kCo + half * 32, acc);
        tmem_ld_wait();
        if (valid && !(p.dbg & 4)) {
          uint4* d = reinterpret_cast<uint4*>(dst + half * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c0 = half * 32 + i * 8;
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(acc[8 * i + 0]) + __ldg(p.bias + c0 + 0),
                              __uint_as_float(acc[8 * i + 1]) + __ldg(p.bias + c0 + 1));
            u.y = pack_bf16x2(__uint_as_float(acc[8 * i + 2]) + __ldg(p.bias + c0 + 2),
                              __uint_as_float(acc[8 * i + 3]) + __ldg(p.bias + c0 + 3));
            u.z = pack_bf16x2(__uint_as_float(acc[8 * i + 4]) + __ldg(p.bias + c0 + 4),
                              __uint_as_float(acc[8 * i + 5]) + __ldg(p.bias + c0 + 5));
            u.w = pack_bf16x2(__uint_as_float(acc[8 * i + 6]) + __ldg(p.bias + c0 + 6),
                              __uint_as_float(acc[8 * i + 7]) + __ldg(p.bias + c0 + 7));
            d[i] = u;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}


bool tc_shape_ok(int Cin, int Win, int C, int K, int stride, int pad, int Wo) {
  static const bool force_generic = std::getenv("FD_STEM_GENERIC") != nullptr;   // A/B testing only
  if (force_generic) return false;
  return stride == 8 && K == 2 * kKH && Win == 8 * kQPR && C == kCo && Cin == 3 && Wo <= 64 && (Win % 32) == 0 &&
         Win / 2 <= 256 && (K / 2) * (Win / 2) * 4 <= kSlotBytes && Win + pad <= 510 && (pad % 2) == 0;
}

StemParams make_params(const void* x, int B, int Cin, int Hin, int Win, int K, int stride, int pad) {
  StemParams p{};
  p.B = B; p.Cin = Cin; p.Hin = Hin; p.Win = Win; p.K = K; p.stride = stride; p.pad = pad;
  p.Ho = (Hin + 2 * pad - K) / stride + 1;
  p.Wo = (Win + 2 * pad - K) / stride + 1;
  p.CK = Cin * K;
  p.npairs = (B + 1) / 2;
  p.ntask = p.npairs * p.Ho;
  p.a_bytes = static_cast<uint32_t>(p.CK) * 2 * kRowBytes;
  p.x = x;
  return p;
}

// [B*Cin, Hin, Win] image (fp32 or uint8) as a 3-D tiled map, box {Win/2, K/2, 1}, no swizzle, zero OOB fill
int make_tmap_image(CUtensorMap* m, const void* x, int is_u8, int planes, int Hin, int Win, int K) {
  return make_tmap_3d(m, x, is_u8 ? 1 : 4, is_u8, Win, Hin, planes, Win / 2, K / 2);
}

// Weight gradient of one (g1 == nullptr) or two 64-channel planes from the bf16 image copy.
int stem_wgrad_cached(const StemParams& p, const fd_bf16* xbf, const fd_bf16* g0, const fd_bf16* g1, cudaStream_t st) {
  CUtensorMap tm_g, tm_g1, tm_xb;
  int rc = make_tmap_nhwc_bf16(&tm_g, g0, p.B, p.Ho, p.Wo, kCo, p.Wo, 1);
  if (rc != FD_OK) return rc;
  rc = make_tmap_nhwc_bf16(&tm_g1, g1 ? g1 : g0, p.B, p.Ho, p.Wo, kCo, p.Wo, 1);
  if (rc != FD_OK) return rc;
  rc = make_tmap_xbf(&tm_xb, xbf, p.npairs * p.Cin, p.Hin, p.K);
  if (rc != FD_OK) return rc;
  const int np = g1 ? 2 : 1;
  const size_t smem = 2 * static_cast<size_t>(p.a_bytes) + 1024 + 2 * np * 128 * 128 + 2048 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(p.ntask, sm_count());
  cudaError_t e;
  if (np == 2) {
    e = set_max_dyn_smem(stem_wgrad_bf16_kernel<2>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_wgrad_bf16_kernel<2>, dim3(grid), dim3(kWg2Threads), smem, st, tm_xb, tm_g, tm_g1, p);
  } else {
    e = set_max_dyn_smem(stem_wgrad_bf16_kernel<1>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_wgrad_bf16_kernel<1>, dim3(grid), dim3(kWg2Threads), smem, st, tm_xb, tm_g, tm_g1, p);
  }
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

}  // namespace

// Elements of the bf16 image cache ([pairs][Cin][Hin][2][512]) for a shape the tensor-core stem handles, else 0.
long stem_cache_elems(int B, int Cin, int Hin, int Win, int C, int K, int stride, int pad) {
  const int Wo = (Win + 2 * pad - K) / stride + 1;
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, Wo)) return 0;
  return static_cast<long>((B + 1) / 2) * Cin * Hin * 2 * 512;
}

// Called from layers.cu's fd_stem_fwd / fd_stem_wgrad; returns FD_EUNSUPPORTED when the shape is
// not the stride-8 stem this kernel is built for (the caller then uses the generic kernel).
int stem_fwd_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win,
                int C, int K, int stride, int pad, fd_bf16* y, fd_bf16* xbf, cudaStream_t st) {
  StemParams p = make_params(x, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.w = w; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.xbf = reinterpret_cast<__nv_bfloat16*>(xbf);
  p.elem_bytes = x_is_u8 ? 1 : 4;
  { const char* d = getenv("FD_STEM_TIMING"); p.dbg = d ? atoi(d) : 0; }
  CUtensorMap tm_x;
  {
    int rc = make_tmap_image(&tm_x, x, x_is_u8, B * Cin, Hin, Win, K);
    if (rc != FD_OK) return rc;
  }
  const size_t smem = static_cast<size_t>(p.CK) * 2048 + p.a_bytes + 128 + kFwdSlots * kSlotBytes + 512 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(p.ntask, sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_fwd_tc_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_fwd_tc_kernel<uint8_t>, dim3(grid), dim3(kFwdThreads), smem, st, tm_x, p);
  } else {
    e = set_max_dyn_smem(stem_fwd_tc_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_fwd_tc_kernel<float>, dim3(grid), dim3(kFwdThreads), smem, st, tm_x, p);
  }
  count_launch();
  return launch_status();
}

int stem_wgrad_tc(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                  int stride, int pad, float* dw, float* dbias, const fd_bf16* xbf, cudaStream_t st) {
  StemParams p = make_params(x, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.dw = dw; p.dbias = dbias;
  { const char* d = getenv("FD_STEM_WG_DBG"); p.dbg = d ? atoi(d) : 0; }     // A/B switches of the bf16 kernel (tools/stem_wgrad_debug.py)
  if (xbf != nullptr) {
    // fast path: operand rows come from the bf16 copy written by the forward pass
    return stem_wgrad_cached(p, xbf, g, nullptr, st);
  }
  p.elem_bytes = x_is_u8 ? 1 : 4;
  CUtensorMap tm_g, tm_x;
  {
    int rc = make_tmap_image(&tm_x, x, x_is_u8, B * Cin, Hin, Win, K);
    if (rc != FD_OK) return rc;
  }
  // g: [B,Ho,Wo,64] bf16; box = one output row of one image
  {
    int rc = make_tmap_nhwc_bf16(&tm_g, g, B, p.Ho, p.Wo, C, p.Wo, 1);
    if (rc != FD_OK) return rc;
  }
  const size_t smem = 2 * static_cast<size_t>(p.a_bytes) + 1024 + 2 * 128 * 128 + 8 * kSlotBytes + 1024 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(p.ntask, sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_wgrad_tc_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_wgrad_tc_kernel<uint8_t>, dim3(grid), dim3(kThreads), smem, st, tm_x, tm_g, p);
  } else {
    e = set_max_dyn_smem(stem_wgrad_tc_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = launch_k(stem_wgrad_tc_kernel<float>, dim3(grid), dim3(kThreads), smem, st, tm_x, tm_g, p);
  }
  count_launch();
  return launch_status();
}

// Forward of one 64-channel plane from the bf16 image copy an earlier stem_fwd_tc launch of the same step wrote.
int stem_fwd_cached(const fd_bf16* xbf, const float* w, const float* bias, int B, int Cin, int Hin, int Win, int C, int K,
                    int stride, int pad, fd_bf16* y, cudaStream_t st) {
  StemParams p = make_params(nullptr, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.w = w; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  { const char* d = getenv("FD_STEM_WG_DBG"); p.dbg = d ? atoi(d) : 0; }     // timing switches (tools/stem_planes_ab.py)
  CUtensorMap tm_xb;
  const int rc = make_tmap_xbf(&tm_xb, xbf, p.npairs * Cin, Hin, K);
  if (rc != FD_OK) return rc;
  const size_t smem = static_cast<size_t>(p.CK) * 2048 + 2 * static_cast<size_t>(p.a_bytes) + 128 + 256 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(stem_fwd_bf16_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = launch_k(stem_fwd_bf16_kernel, dim3(min(p.ntask, sm_count())), dim3(kWg2Threads), smem, st, tm_xb, p);
  if (e != cudaSuccess) return (int)e;
  count_launch();
  return launch_status();
}

// Weight gradients of two 64-channel planes (dw [128][Cin][K][K], dbias [128], accumulated) in one pass over the copy.
int stem_wgrad_pair_cached(const fd_bf16* xbf, const fd_bf16* g0, const fd_bf16* g1, int B, int Cin, int Hin, int Win,
                           int K, int stride, int pad, float* dw, float* dbias, cudaStream_t st) {
  StemParams p = make_params(nullptr, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, kCo, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.dw = dw; p.dbias = dbias;
  { const char* d = getenv("FD_STEM_WG_DBG"); p.dbg = d ? atoi(d) : 0; }
  return stem_wgrad_cached(p, xbf, g0, g1, st);
}

}  // namespace fd

extern "C" FD_API int fd_debug_stem_timing(unsigned long long* out, int n) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, fd::g_stem_dbg, sizeof(unsigned long long) * n));
}
