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
constexpr int kLoaderWarps = 8;
constexpr int kThreads = (kLoaderWarps + 1 + 4) * 32;   // loaders | MMA | epilogue  = 416
constexpr int kMaxCK = 30;                 // Cin * K rows per image per task

struct StemParams {
  int B, Cin, Hin, Win, K, stride, pad, Ho, Wo;
  int CK;           // Cin * K
  int npairs;       // ceil(B / 2)
  int ntask;        // npairs * Ho
  uint32_t a_bytes; // CK * 2 * 1024
  const void* x;
  const float* w;        // [64][Cin][K][K] fp32 (forward)
  const float* bias;     // [64]
  __nv_bfloat16* y;      // [B,Ho,Wo,64]
  float* dw;             // [64][Cin][K][K] fp32, accumulated (wgrad)
  float* dbias;          // [64] accumulated (wgrad)
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
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 f = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
  }
};
template <>
struct Px4<uint8_t> {
  static __device__ __forceinline__ void load(const uint8_t* p, float (&v)[4]) {
    const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(p));
    v[0] = static_cast<float>(u.x) / 255.0f; v[1] = static_cast<float>(u.y) / 255.0f;   // PoolResnet.py:95
    v[2] = static_cast<float>(u.z) / 255.0f; v[3] = static_cast<float>(u.w) / 255.0f;
  }
};

// Loader warps: fill one A tile (CK*2 rows) for task (pair, oy).  Element e of a row holds input
// column e - pad; columns outside the image stay zero from the one-time clear.
template <typename TIn>
__device__ __forceinline__ void load_task(const StemParams& p, uint8_t* abuf, int pair, int oy, int ltid) {
  const TIn* x = static_cast<const TIn*>(p.x);
  const int q_per_row = p.Win >> 2;                 // 4-pixel chunks per row
  const int total = p.CK * 2 * q_per_row;
  constexpr int kBatch = 7;
  for (int base = ltid; base < total; base += kLoaderWarps * 32 * kBatch) {
    float v[kBatch][4];
    int dst[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int idx = base + u * kLoaderWarps * 32;
      dst[u] = -1;
      if (idx < total) {
        const int qx = idx % q_per_row;
        const int row = idx / q_per_row;              // (c*K + ky)*2 + img
        const int img = row & 1, cky = row >> 1;
        const int ky = cky % p.K, c = cky / p.K;
        const int n = pair * 2 + img;
        const int iy = oy * p.stride + ky - p.pad;
        dst[u] = row * kRowBytes + (qx * 4 + p.pad) * 2;
        if (n < p.B && iy >= 0 && iy < p.Hin) {
          Px4<TIn>::load(x + ((static_cast<size_t>(n) * p.Cin + c) * p.Hin + iy) * p.Win + qx * 4, v[u]);
        } else {
          v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      if (dst[u] >= 0) {
        uint32_t* d = reinterpret_cast<uint32_t*>(abuf + dst[u]);
        d[0] = pack_bf16x2(v[u][0], v[u][1]);
        d[1] = pack_bf16x2(v[u][2], v[u][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------- forward
// smem: [weights CK*2048][A stage 0][A stage 1][slack 64][barriers]
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1) stem_fwd_tc_kernel(const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sW = smem;
  uint8_t* sA = sW + p.CK * 2048;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * p.a_bytes + 64);
  uint64_t* a_full = bars + 0;     // [2] count = loader warps
  uint64_t* a_empty = bars + 2;    // [2]
  uint64_t* acc_full = bars + 4;   // [2]
  uint64_t* acc_empty = bars + 6;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one-time: zero the A stages (pad columns stay zero forever) and build the bf16 weight operand
  for (uint32_t i = threadIdx.x * 16u; i < 2 * p.a_bytes + 64; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  {
    // B operand per (c,ky): [k1 2][co 64][k0 8] bf16  (K-major, no swizzle: LBO = 1024, SBO = 128)
    __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(sW);
    const int KK = p.CK * p.K;
    for (int i = threadIdx.x; i < p.CK * 1024; i += kThreads) {
      const int k0 = i & 7, co = (i >> 3) & 63, k1 = (i >> 9) & 1, cky = i >> 10;
      const int kx = k1 * 8 + k0;
      const float v = kx < p.K ? p.w[static_cast<size_t>(co) * KK + cky * p.K + kx] : 0.f;
      w16[i] = __float2bfloat16(v);
    }
  }
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(a_full + s, kLoaderWarps);
      mbar_init(a_empty + s, 1);
      mbar_init(acc_full + s, 1);
      mbar_init(acc_empty + s, 4);
    }
    fence_barrier_init();
  }
  if (warp == kLoaderWarps) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < kLoaderWarps) {
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(a_empty + s, ph ^ 1);
      load_task<TIn>(p, sA + s * p.a_bytes, task / p.Ho, task % p.Ho, threadIdx.x);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + s);
    }
  } else if (warp == kLoaderWarps) {
    constexpr uint32_t idesc = make_idesc_bf16(128, kCo, 0, 0);
    const uint32_t w_addr = smem_u32(sW);
    int it = 0;
    for (int task = blockIdx.x; task < p.ntask; task += gridDim.x, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(acc_empty + s, ph ^ 1);
      mbar_wait(a_full + s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(sA + s * p.a_bytes);
        for (int cky = 0; cky < p.CK; ++cky) {
          const uint64_t ad = make_sdesc_none(a_addr + cky * 2048, 16, 128);
          const uint64_t bd = make_sdesc_none(w_addr + cky * 2048, 1024, 128);
          umma_bf16(tmem_base + s * kCo, ad, bd, idesc, cky != 0 ? 1u : 0u);
        }
        umma_commit(a_empty + s);
        umma_commit(acc_full + s);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kLoaderWarps) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------- weight gradient
// smem: [A stage 0][A stage 1][slack 1024][g stage 0 (16 KB)][g stage 1][barriers][bias scratch]
template <typename TIn>
__global__ void __launch_bounds__(kThreads, 1)
stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_g, const StemParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sG = sA + 2 * p.a_bytes + 1024;
  constexpr uint32_t kGBytes = 128 * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sG + 2 * kGBytes);
  uint64_t* full = bars + 0;     // [2] count = loader warps + 1 (TMA expect_tx)
  uint64_t* empty = bars + 2;    // [2] count = 1 (MMA commit) + 4 (bias warps)
  uint64_t* acc_full = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sBias = reinterpret_cast<float*>(bars + 6);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x * 16u; i < 2 * p.a_bytes + 1024 + 2 * kGBytes; i += kThreads * 16u)
    *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_g);
    for (int s = 0; s < 2; ++s) {
      mbar_init(full + s, kLoaderWarps + 1);
      mbar_init(empty + s, 1 + 4);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == kLoaderWarps) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // contiguous chunk of tasks per CTA
  const int per = (p.ntask + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per;
  const int t_end = min(p.ntask, t_begin + per);

  if (warp < kLoaderWarps) {
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(empty + s, ph ^ 1);
      load_task<TIn>(p, sA + s * p.a_bytes, task / p.Ho, task % p.Ho, threadIdx.x);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(full + s);
    }
  } else if (warp == kLoaderWarps) {
    constexpr uint32_t idesc = make_idesc_bf16(64, 16, 1, 1);
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      if (lane == 0) {
        mbar_wait(empty + s, ph ^ 1);
        mbar_expect_tx(full + s, 2u * p.Wo * 128u);
        // one box {64 ch, Wo, 1, 1} per image into 64-row slots; rows Wo..63 stay zero; n >= B is zero filled
        tma_load_4d(sG + s * kGBytes, &tm_g, full + s, 0, 0, task % p.Ho, (task / p.Ho) * 2);
        tma_load_4d(sG + s * kGBytes + 64 * 128, &tm_g, full + s, 0, 0, task % p.Ho, (task / p.Ho) * 2 + 1);
      }
      __syncwarp();
      mbar_wait(full + s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(sA + s * p.a_bytes);
        const uint32_t g_addr = smem_u32(sG + s * kGBytes);
        for (int ks = 0; ks < 8; ++ks) {
          // A = g^T: MN-major (M = co), 128B swizzle, 16 K-rows (positions) per MMA
          const uint64_t ad = make_sdesc_sw128(g_addr + ks * 2048, 1024, 1024, 0);
          for (int cky = 0; cky < p.CK; ++cky) {
            // B = input windows: MN-major (N = kx), no swizzle: kx chunks 16 B apart, 8-position groups 128 B apart
            const uint64_t bd = make_sdesc_none(a_addr + cky * 2048 + ks * 256, 128, 16);
            umma_bf16(tmem_base + cky * 16, ad, bd, idesc, (it | ks) != 0 ? 1u : 0u);
          }
        }
        umma_commit(empty + s);
        if (task + 1 == t_end) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int et = threadIdx.x - (kLoaderWarps + 1) * 32;   // 0..127
    const int c = et & 63, rpar = et >> 6;
    float bsum = 0.f;
    int it = 0;
    for (int task = t_begin; task < t_end; ++task, ++it) {
      const int s = it & 1, ph = (it >> 1) & 1;
      mbar_wait(full + s, ph);
      if (p.dbias) {
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
  if (warp == kLoaderWarps) tmem_dealloc(tmem_base, 512);
}

bool tc_shape_ok(int Cin, int Win, int C, int K, int stride, int pad, int Wo) {
  static const bool force_generic = std::getenv("FD_STEM_GENERIC") != nullptr;   // A/B testing only
  if (force_generic) return false;
  return stride == 8 && K <= 16 && C == kCo && Cin * K <= kMaxCK && Wo <= 64 && (Win % 4) == 0 &&
         Win + pad <= 510 && (pad % 2) == 0;
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

}  // namespace

// Called from layers.cu's fd_stem_fwd / fd_stem_wgrad; returns FD_EUNSUPPORTED when the shape is
// not the stride-8 stem this kernel is built for (the caller then uses the generic kernel).
int stem_fwd_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win,
                int C, int K, int stride, int pad, fd_bf16* y, cudaStream_t st) {
  StemParams p = make_params(x, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.w = w; p.bias = bias; p.y = reinterpret_cast<__nv_bfloat16*>(y);
  const size_t smem = static_cast<size_t>(p.CK) * 2048 + 2 * p.a_bytes + 64 + 256 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(p.ntask, sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = cudaFuncSetAttribute(stem_fwd_tc_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_fwd_tc_kernel<uint8_t><<<grid, kThreads, smem, st>>>(p);
  } else {
    e = cudaFuncSetAttribute(stem_fwd_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_fwd_tc_kernel<float><<<grid, kThreads, smem, st>>>(p);
  }
  count_launch();
  return launch_status();
}

int stem_wgrad_tc(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                  int stride, int pad, float* dw, float* dbias, cudaStream_t st) {
  StemParams p = make_params(x, B, Cin, Hin, Win, K, stride, pad);
  if (!tc_shape_ok(Cin, Win, C, K, stride, pad, p.Wo)) return FD_EUNSUPPORTED;
  p.dw = dw; p.dbias = dbias;
  CUtensorMap tm_g;
  // g: [B,Ho,Wo,64] bf16; box = one output row of one image
  {
    int rc = make_tmap_nhwc_bf16(&tm_g, g, B, p.Ho, p.Wo, C, p.Wo, 1);
    if (rc != FD_OK) return rc;
  }
  const size_t smem = 2 * static_cast<size_t>(p.a_bytes) + 1024 + 2 * 128 * 128 + 1024 + 1024;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(p.ntask, sm_count());
  cudaError_t e;
  if (x_is_u8) {
    e = cudaFuncSetAttribute(stem_wgrad_tc_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_tc_kernel<uint8_t><<<grid, kThreads, smem, st>>>(tm_g, p);
  } else {
    e = cudaFuncSetAttribute(stem_wgrad_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_tc_kernel<float><<<grid, kThreads, smem, st>>>(tm_g, p);
  }
  count_launch();
  return launch_status();
}

}  // namespace fd
