// CUDA-core kernels of the backbone that are not 64->64 3x3 convolutions: weight packing,
// the stem convolution (3 -> C, 10x10 stride 8), the 5-channel head (+ sigmoid) and MaxPool2d(2),
// each forward and backward.  All are memory- or latency-bound; the tensor-core work lives in
// conv3x3_tc.cu / wgrad3x3_tc.cu.
#include <cmath>
#include <cstdlib>

#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

// ============================================================================ weight packing
// torch [co][ci][ky][kx] fp32  ->  fwd [t][co][ci] bf16 ; dgrad [t'][ci][co] bf16 with t' = flipped tap.
// One CTA = (layer, kPackCo output channels): kPackCo x 576 contiguous floats in, transposed through shared
// memory.  C is a template constant so that all index arithmetic is shifts / multiplies by constants.
constexpr int kPackCo = 4;
template <int C>
__global__ void __launch_bounds__(256)
pack_conv3x3_kernel(const float* __restrict__ w, int n_layers, __nv_bfloat16* __restrict__ wf,
                    __nv_bfloat16* __restrict__ wd) {
  pdl_trigger();
  pdl_wait();
  constexpr int row = C * 9, pitch = row + 1, tiles = C / kPackCo;
  __shared__ float sm[kPackCo * pitch];
  const int l = blockIdx.x / tiles, co0 = (blockIdx.x % tiles) * kPackCo;
  const float* src = w + (static_cast<long>(l) * C + co0) * row;
  for (int i = threadIdx.x; i < kPackCo * row; i += 256) sm[(i / row) * pitch + i % row] = __ldg(src + i);
  __syncthreads();
  const long lbase = static_cast<long>(l) * 9 * C * C;
  if (wf) {                                          // wf[l][t][co][ci]: ci fastest
    for (int i = threadIdx.x; i < 9 * kPackCo * C; i += 256) {
      const int ci = i % C, co = (i / C) % kPackCo, t = i / (C * kPackCo);
      wf[lbase + (static_cast<long>(t) * C + co0 + co) * C + ci] = __float2bfloat16(sm[co * pitch + ci * 9 + t]);
    }
  }
  if (wd) {                                          // wd[l][8-t][ci][co]: co fastest
    for (int i = threadIdx.x; i < 9 * kPackCo * C; i += 256) {
      const int co = i % kPackCo, ci = (i / kPackCo) % C, t = i / (C * kPackCo);
      wd[lbase + (static_cast<long>(8 - t) * C + ci) * C + co0 + co] = __float2bfloat16(sm[co * pitch + ci * 9 + t]);
    }
  }
}

// packed gradient [t][ci][co] fp32 -> torch [co][ci][ky][kx] fp32 (same tiling, contiguous 576-float rows out)
template <int C>
__global__ void __launch_bounds__(256)
unpack_wgrad3x3_kernel(const float* __restrict__ dwp, int n_layers, float* __restrict__ dw) {
  pdl_trigger();
  pdl_wait();
  constexpr int row = C * 9, pitch = row + 1, tiles = C / kPackCo;
  __shared__ float sm[kPackCo * pitch];
  const int l = blockIdx.x / tiles, co0 = (blockIdx.x % tiles) * kPackCo;
  const long lbase = static_cast<long>(l) * 9 * C * C;
  for (int i = threadIdx.x; i < 9 * kPackCo * C; i += 256) {
    const int co = i % kPackCo, ci = (i / kPackCo) % C, t = i / (C * kPackCo);
    sm[co * pitch + ci * 9 + t] = __ldg(dwp + lbase + (static_cast<long>(t) * C + ci) * C + co0 + co);
  }
  __syncthreads();
  float* dst = dw + (static_cast<long>(l) * C + co0) * row;
  for (int i = threadIdx.x; i < kPackCo * row; i += 256) dst[i] = sm[(i / row) * pitch + i % row];
}

// The same for the channel-plane engines: packed 64 x 64 sub-blocks [layer][g][h][t][ci][co] -> the torch tensor of the WIDE
// layer [layer][64 G (co)][64 G (ci)][ky][kx] directly (row of 576 floats per (co, h) at its place in the wide row): no
// intermediate sub-block tensor and no permuting copy afterwards.
constexpr int kUnpT = 32;     // (co, ci) tile: 128-byte runs of the packed source, 1152-byte runs of the destination
__global__ void __launch_bounds__(256)
unpack_wgrad3x3_planes_kernel(const float* __restrict__ dwp, int G, float* __restrict__ dw) {
  pdl_trigger();
  pdl_wait();
  constexpr int C = 64, row = kUnpT * 9, pitch = row + 1, tiles = C / kUnpT;
  __shared__ float sm[kUnpT * pitch];
  const int sub = blockIdx.x / (tiles * tiles), rem = blockIdx.x % (tiles * tiles);
  const int co0 = (rem / tiles) * kUnpT, ci0 = (rem % tiles) * kUnpT;
  const int h = sub % G, g = (sub / G) % G, l = sub / (G * G);
  const long sbase = static_cast<long>(sub) * 9 * C * C;
  for (int i = threadIdx.x; i < 9 * kUnpT * kUnpT; i += 256) {
    const int co = i % kUnpT, ci = (i / kUnpT) % kUnpT, t = i / (kUnpT * kUnpT);
    sm[co * pitch + ci * 9 + t] = __ldg(dwp + sbase + (static_cast<long>(t) * C + ci0 + ci) * C + co0 + co);
  }
  __syncthreads();
  const long F = static_cast<long>(G) * C;
  for (int i = threadIdx.x; i < kUnpT * row; i += 256) {
    const int co = i / row, r = i % row;
    dw[((static_cast<long>(l) * F + g * C + co0 + co) * F + h * C + ci0) * 9 + r] = sm[co * pitch + r];
  }
}

// Adam on the flat parameter / gradient buffers (models/ModelMeta.py:104-112: the reference's SAMSGD never
// recomputes gradients, i.e. it is plain torch Adam).  Same update as torch.optim.Adam (no amsgrad, L2 weight decay):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// `state` (nullable, device): int32 {step count, CTA ticket} + fp32 {learning rate} -- with it the step count and
// the learning rate live on the device, so the launch can be replayed from a CUDA graph; the last CTA to have read
// the count publishes count + 1.
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long n, float lr, float b1, float b2, float eps, float wd,
                                 int step, int* __restrict__ state) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_hyper[2];
  if (threadIdx.x == 0) {
    if (state) {
      step = state[0] + 1;
      lr = __int_as_float(state[2]);
      __threadfence();
      if (atomicAdd(state + 1, 1) == static_cast<int>(gridDim.x) - 1) {      // every CTA has read state[0]
        state[1] = 0;
        state[0] = step;
      }
    }
    const double bc1 = 1.0 - pow(static_cast<double>(b1), static_cast<double>(step));
    const double bc2 = 1.0 - pow(static_cast<double>(b2), static_cast<double>(step));
    s_hyper[0] = static_cast<float>(static_cast<double>(lr) / bc1);
    s_hyper[1] = static_cast<float>(sqrt(bc2));
  }
  __syncthreads();
  const float step_size = s_hyper[0], bc2_sqrt = s_hyper[1];
  for (long i = (blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x) * 4; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x * 4) {
    float4 pv = *reinterpret_cast<float4*>(p + i), gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    float* pp = &pv.x; float* gp = &gv.x; float* mp = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = gp[e] + wd * pp[e];
      mp[e] = b1 * mp[e] + (1.f - b1) * gg;
      vp[e] = b2 * vp[e] + (1.f - b2) * gg * gg;
      const float denom = sqrtf(vp[e]) / bc2_sqrt + eps;
      pp[e] -= step_size * (mp[e] / denom);
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
  }
}

// Dropout2d multipliers from uniform randoms: rows [0, n_block_rows) use keep_b, the rest keep_h
// (models/PoolResnet.py:39 Dropout2d(0.25) per block, :100 Dropout2d(0.5) before the head).
__global__ void dropout_scale_kernel(const float* __restrict__ r, long n, long n_block, float keep_b, float keep_h,
                                     float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float keep = i < n_block ? keep_b : keep_h;
    out[i] = r[i] < keep ? 1.f / keep : 0.f;
  }
}

// ============================================================================ stem
// One CTA task = one output row (n, oy): Wo x C outputs, K = Cin*Kk*Kk.  256 threads:
// thread = (cog = tid & 15 -> 4 output channels, oxg = tid >> 4 -> 4 output columns).
// smem: weights transposed to [k][co] fp32 (persistent), the Kk input rows of every channel.
constexpr int kStemThreads = 256;

template <typename TIn>
__device__ __forceinline__ float stem_in(const TIn* p) { return static_cast<float>(*p); }
template <>
__device__ __forceinline__ float stem_in<uint8_t>(const uint8_t* p) { return static_cast<float>(*p) / 255.0f; }

template <typename TIn>
__device__ void stem_load_rows(const TIn* __restrict__ x, float* sIn, int n, int oy, int Cin, int Hin, int Win, int K,
                               int stride, int pad, int pitch) {
  // sIn[c][ky][pitch]: column j holds input column j - pad
  const int rows = Cin * K;
  for (int idx = threadIdx.x; idx < rows * pitch; idx += blockDim.x) {
    const int j = idx % pitch;
    const int r = idx / pitch;
    const int ky = r % K, c = r / K;
    const int iy = oy * stride + ky - pad;
    const int ix = j - pad;
    float v = 0.f;
    if (iy >= 0 && iy < Hin && ix >= 0 && ix < Win)
      v = stem_in<TIn>(x + ((static_cast<size_t>(n) * Cin + c) * Hin + iy) * Win + ix);
    sIn[idx] = v;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_fwd_kernel(const TIn* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int B, int Cin,
                int Hin, int Win, int C, int K, int stride, int pad, int Ho, int Wo, __nv_bfloat16* __restrict__ y) {
  extern __shared__ float sm[];
  const int KK = Cin * K * K;
  const int pitch = (Wo - 1) * stride + K + 1;
  float* sW = sm;               // [KK][C]
  float* sIn = sm + KK * C;     // [Cin*K][pitch]
  for (int i = threadIdx.x; i < KK * C; i += blockDim.x) {
    const int co = i % C, k = i / C;
    sW[i] = w[static_cast<size_t>(co) * KK + k];
  }
  const int cog = threadIdx.x & 15, oxg = threadIdx.x >> 4;
  const int ntask = B * Ho;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    const int n = task / Ho, oy = task % Ho;
    __syncthreads();
    stem_load_rows<TIn>(x, sIn, n, oy, Cin, Hin, Win, K, stride, pad, pitch);
    __syncthreads();
    for (int cb = 0; cb < C; cb += 64) {
      const int co0 = cb + cog * 4;
      for (int ob = 0; ob < Wo; ob += 64) {
        const int ox0 = ob + oxg * 4;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
        if (ox0 < Wo) {
          for (int r = 0; r < Cin * K; ++r) {
            const float* inrow = sIn + r * pitch;
            const float* wrow = sW + static_cast<size_t>(r) * K * C + co0;
            for (int kx = 0; kx < K; ++kx) {
              const float4 wv = *reinterpret_cast<const float4*>(wrow + kx * C);
#pragma unroll
              for (int a = 0; a < 4; ++a) {
                const int ox = min(ox0 + a, Wo - 1);
                const float iv = inrow[ox * stride + kx];
                acc[a][0] = fmaf(iv, wv.x, acc[a][0]);
                acc[a][1] = fmaf(iv, wv.y, acc[a][1]);
                acc[a][2] = fmaf(iv, wv.z, acc[a][2]);
                acc[a][3] = fmaf(iv, wv.w, acc[a][3]);
              }
            }
          }
          const float4 bv = *reinterpret_cast<const float4*>(bias + co0);
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int ox = ox0 + a;
            if (ox < Wo) {
              uint2 o;
              o.x = pack_bf16x2(acc[a][0] + bv.x, acc[a][1] + bv.y);
              o.y = pack_bf16x2(acc[a][2] + bv.z, acc[a][3] + bv.w);
              *reinterpret_cast<uint2*>(y + ((static_cast<size_t>(n) * Ho + oy) * Wo + ox) * C + co0) = o;
            }
          }
        }
      }
    }
  }
}

// wgrad: thread = (cog = tid & 15 -> 4 output channels, kg = tid >> 4 -> KPT consecutive k).
// Accumulators stay in registers over all row tasks of the persistent CTA, one atomic flush at the end.
constexpr int kStemKPT = 19;  // ceil(300 / 16)

template <typename TIn>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_wgrad_kernel(const TIn* __restrict__ x, const __nv_bfloat16* __restrict__ g, int B, int Cin, int Hin, int Win,
                  int C, int K, int stride, int pad, int Ho, int Wo, float* __restrict__ dw, float* __restrict__ dbias) {
  extern __shared__ float sm[];
  const int KK = Cin * K * K;
  const int pitch = (Wo - 1) * stride + K + 1;
  float* sG = sm;                 // [Wo][C]
  float* sIn = sm + Wo * C;       // [Cin*K][pitch]
  const int cog = threadIdx.x & 15, kg = threadIdx.x >> 4;
  const int ntask = B * Ho;
  for (int cb = 0; cb < C; cb += 64) {
    const int co0 = cb + cog * 4;
    float acc[kStemKPT][4];
    float bacc[4] = {0.f, 0.f, 0.f, 0.f};
    int koff[kStemKPT];
#pragma unroll
    for (int i = 0; i < kStemKPT; ++i) {
      acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      const int k = min(kg * kStemKPT + i, KK - 1);
      const int kx = k % K, r = k / K;  // r = c*K + ky
      koff[i] = r * pitch + kx;
    }
    for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
      const int n = task / Ho, oy = task % Ho;
      __syncthreads();
      stem_load_rows<TIn>(x, sIn, n, oy, Cin, Hin, Win, K, stride, pad, pitch);
      for (int i = threadIdx.x; i < Wo * C; i += blockDim.x)
        sG[i] = __bfloat162float(g[(static_cast<size_t>(n) * Ho + oy) * Wo * C + i]);
      __syncthreads();
      for (int ox = 0; ox < Wo; ++ox) {
        const float4 gv = *reinterpret_cast<const float4*>(sG + ox * C + co0);
        const float* inb = sIn + ox * stride;
        if (kg == 0) { bacc[0] += gv.x; bacc[1] += gv.y; bacc[2] += gv.z; bacc[3] += gv.w; }
#pragma unroll
        for (int i = 0; i < kStemKPT; ++i) {
          const float iv = inb[koff[i]];
          acc[i][0] = fmaf(iv, gv.x, acc[i][0]);
          acc[i][1] = fmaf(iv, gv.y, acc[i][1]);
          acc[i][2] = fmaf(iv, gv.z, acc[i][2]);
          acc[i][3] = fmaf(iv, gv.w, acc[i][3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kStemKPT; ++i) {
      const int k = kg * kStemKPT + i;
      if (k < KK) {
#pragma unroll
        for (int a = 0; a < 4; ++a) atomicAdd(dw + static_cast<size_t>(co0 + a) * KK + k, acc[i][a]);
      }
    }
    if (kg == 0 && dbias) {
#pragma unroll
      for (int a = 0; a < 4; ++a) atomicAdd(dbias + co0 + a, bacc[a]);
    }
  }
}

// ============================================================================ head
// grid = (B, kHeadSplit): each CTA stages one image (bf16, dropout multiplier applied on the fly)
// and computes a slice of the output pixels, one warp per pixel, lanes over channels.
constexpr int kHeadThreads = 256;
constexpr int kHeadSplit = 4;

__global__ void __launch_bounds__(kHeadThreads)
head_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ w,
                const float* __restrict__ bias, int B, int H, int W, int C, int K, int pad, int Ho, int Wo,
                float* __restrict__ y) {
  extern __shared__ float sm[];
  const int CP = C + 1;
  float* sW = sm;                                                      // [K*K][5][C+1]
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(sm + ((K * K * 5 * CP + 3) & ~3));  // [H*W][C]
  const int n = blockIdx.x;
  // coalesced read of w in its natural [o][c][t] order, transposing scatter into [t][o][c] rows of pitch C+1
  // (bank stride 5*(C+1) = 5 mod 32 for C = 64: conflict-free)
  for (int i = threadIdx.x; i < K * K * 5 * C; i += blockDim.x) {
    const int t = i % (K * K), c = (i / (K * K)) % C, o = i / (K * K * C);
    sW[(t * 5 + o) * CP + c] = __ldg(w + i);
  }
  const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(n) * H * W * C);
  for (int i = threadIdx.x; i < H * W * C / 8; i += blockDim.x) reinterpret_cast<uint4*>(sX)[i] = __ldg(xs + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int pix = blockIdx.y * nwarp + warp; pix < Ho * Wo; pix += gridDim.y * nwarp) {
    const int oy = pix / Wo, ox = pix % Wo;
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = lane; c < C; c += 32) {
      const float s = cs ? cs[n * C + c] : 1.f;
      for (int ky = 0; ky < K; ++ky) {
        const int iy = oy + ky - pad;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < K; ++kx) {
          const int ix = ox + kx - pad;
          if (ix < 0 || ix >= W) continue;
          const float xv = __bfloat162float(sX[(iy * W + ix) * C + c]) * s;
          const float* wp = sW + (ky * K + kx) * 5 * CP + c;
#pragma unroll
          for (int o = 0; o < 5; ++o) acc[o] = fmaf(xv, wp[o * CP], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 5; ++o) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
    }
    if (lane < 5) {
      float v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : lane == 3 ? acc[3] : acc[4];
      v += bias[lane];
      y[((static_cast<size_t>(n) * 5 + lane) * Ho + oy) * Wo + ox] = 1.f / (1.f + expf(-v));
    }
  }
}

// backward: one CTA per image (loops over images), 256 threads.
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ w,
                const float* __restrict__ y, const float* __restrict__ dy, int B, int H, int W, int C, int K, int pad,
                int Ho, int Wo, __nv_bfloat16* __restrict__ dx, const uint32_t* __restrict__ mask_bits,
                const float* __restrict__ cs2, float slope, __nv_bfloat16* __restrict__ dx2, float* __restrict__ dw,
                float* __restrict__ dbias) {
  extern __shared__ float sm[];
  const int KK = K * K, CP = C + 1;
  float* sW = sm;                       // [KK][5][C+1]
  float* sDz = sW + ((KK * 5 * CP + 3) & ~3);   // [5][Ho*Wo]
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(sDz + ((5 * Ho * Wo + 3) & ~3));  // [H*W][C]
  for (int i = threadIdx.x; i < KK * 5 * C; i += blockDim.x) {
    const int t = i % KK, c = (i / KK) % C, o = i / (KK * C);
    sW[(t * 5 + o) * CP + c] = __ldg(w + i);
  }
  const int c = threadIdx.x % C;        // requires blockDim % C == 0
  const int grp = threadIdx.x / C, ngrp = blockDim.x / C;
  for (int n = blockIdx.x; n < B; n += gridDim.x) {
    __syncthreads();
    const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(n) * H * W * C);
    for (int i = threadIdx.x; i < H * W * C / 8; i += blockDim.x) reinterpret_cast<uint4*>(sX)[i] = __ldg(xs + i);
    for (int i = threadIdx.x; i < 5 * Ho * Wo; i += blockDim.x) {
      const float yv = y[static_cast<size_t>(n) * 5 * Ho * Wo + i];
      sDz[i] = dy[static_cast<size_t>(n) * 5 * Ho * Wo + i] * yv * (1.f - yv);
    }
    __syncthreads();
    const float s = cs ? cs[n * C + c] : 1.f;
    // ---- dbias
    if (threadIdx.x < 5) {
      float t = 0.f;
      for (int i = 0; i < Ho * Wo; ++i) t += sDz[threadIdx.x * Ho * Wo + i];
      atomicAdd(dbias + threadIdx.x, t);
    }
    // ---- dx (and the masked copy that starts the last block's backward chain)
    const float s2 = cs2 ? cs2[n * C + c] : 1.f;
    for (int pix = grp; pix < H * W; pix += ngrp) {
      const int iy = pix / W, ix = pix % W;
      float acc = 0.f;
      for (int ky = 0; ky < K; ++ky) {
        const int oy = iy - ky + pad;
        if (oy < 0 || oy >= Ho) continue;
        for (int kx = 0; kx < K; ++kx) {
          const int ox = ix - kx + pad;
          if (ox < 0 || ox >= Wo) continue;
          const float* wp = sW + (ky * K + kx) * 5 * CP + c;
#pragma unroll
          for (int o = 0; o < 5; ++o) acc = fmaf(sDz[o * Ho * Wo + oy * Wo + ox], wp[o * CP], acc);
        }
      }
      acc *= s;
      const size_t gi = (static_cast<size_t>(n) * H * W + pix) * C + c;
      if (dx) dx[gi] = __float2bfloat16(acc);
      if (dx2) {
        const float m = ((mask_bits[gi >> 5] >> (gi & 31)) & 1u) ? 1.f : slope;
        dx2[gi] = __float2bfloat16(acc * m * s2);
      }
    }
    // ---- dw: thread (c, grp) owns taps grp, grp+ngrp, ...
    for (int t = grp; t < KK; t += ngrp) {
      const int ky = t / K, kx = t % K;
      float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
      for (int oy = 0; oy < Ho; ++oy) {
        const int iy = oy + ky - pad;
        if (iy < 0 || iy >= H) continue;
        for (int ox = 0; ox < Wo; ++ox) {
          const int ix = ox + kx - pad;
          if (ix < 0 || ix >= W) continue;
          const float xv = __bfloat162float(sX[(iy * W + ix) * C + c]) * s;
#pragma unroll
          for (int o = 0; o < 5; ++o) acc[o] = fmaf(xv, sDz[o * Ho * Wo + oy * Wo + ox], acc[o]);
        }
      }
#pragma unroll
      for (int o = 0; o < 5; ++o) atomicAdd(dw + (static_cast<size_t>(o) * C + c) * KK + t, acc[o]);
    }
  }
}

// Register-tiled head backward for small maps (W, Wo <= 16, C = 64): each thread owns a whole map
// row (dx) or a whole kernel row (dw) in registers, so the inner loops are FMA-bound instead of
// shared-memory-bound.  K and PAD are compile-time so that every register index is static.
template <int K, int PAD>
__global__ void __launch_bounds__(kHeadThreads)
head_bwd_small_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ w,
                      const float* __restrict__ y, const float* __restrict__ dy, int B, int H, int W, int Ho, int Wo,
                      __nv_bfloat16* __restrict__ dx, const uint32_t* __restrict__ mask_bits,
                      const float* __restrict__ cs2, float slope, __nv_bfloat16* __restrict__ dx2,
                      float* __restrict__ dw, float* __restrict__ dbias) {
  constexpr int C = 64, KK = K * K, MW = 16, CP = C + 1;
  extern __shared__ float sm[];
  float* sW = sm;                         // [KK][5][C+1]
  float* sDz = sW + ((KK * 5 * CP + 3) & ~3);   // [5][Ho][MW] (rows padded to 16, zero filled)
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(sDz + 5 * Ho * MW);  // [H*W][C]
  for (int i = threadIdx.x; i < KK * 5 * C; i += blockDim.x) {
    const int t = i % KK, c = (i / KK) % C, o = i / (KK * C);
    sW[(t * 5 + o) * CP + c] = __ldg(w + i);
  }
  for (int n = blockIdx.x; n < B; n += gridDim.x) {
    __syncthreads();
    const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(n) * H * W * C);
    for (int i = threadIdx.x; i < H * W * C / 8; i += blockDim.x) reinterpret_cast<uint4*>(sX)[i] = __ldg(xs + i);
    for (int i = threadIdx.x; i < 5 * Ho * MW; i += blockDim.x) {
      const int ox = i % MW, r = i / MW;     // r = o*Ho + oy
      float v = 0.f;
      if (ox < Wo) {
        const size_t gi = static_cast<size_t>(n) * 5 * Ho * Wo + static_cast<size_t>(r) * Wo + ox;
        const float yv = y[gi];
        v = dy[gi] * yv * (1.f - yv);
      }
      sDz[i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      float t = 0.f;
      for (int i = 0; i < Ho * MW; ++i) t += sDz[threadIdx.x * Ho * MW + i];
      atomicAdd(dbias + threadIdx.x, t);
    }
    // ---- dx: task = (c, iy), MW outputs in registers
    for (int task = threadIdx.x; task < C * H; task += blockDim.x) {
      const int c = task % C, iy = task / C;
      float acc[MW];
#pragma unroll
      for (int i = 0; i < MW; ++i) acc[i] = 0.f;
      for (int o = 0; o < 5; ++o) {
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int oy = iy - ky + PAD;
          if (oy < 0 || oy >= Ho) continue;
          float dz[MW];
          const float4* dzr = reinterpret_cast<const float4*>(sDz + (o * Ho + oy) * MW);
#pragma unroll
          for (int i = 0; i < MW / 4; ++i) {
            const float4 t = dzr[i];
            dz[4 * i] = t.x; dz[4 * i + 1] = t.y; dz[4 * i + 2] = t.z; dz[4 * i + 3] = t.w;
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float wv = sW[((ky * K + kx) * 5 + o) * CP + c];
#pragma unroll
            for (int ix = 0; ix < MW; ++ix) {
              constexpr int dummy = 0; (void)dummy;
              const int ox = ix - kx + PAD;
              if (ox >= 0 && ox < MW) acc[ix] = fmaf(dz[ox], wv, acc[ix]);   // dz is zero past Wo
            }
          }
        }
      }
      const float s = cs ? cs[n * C + c] : 1.f;
      const float s2 = cs2 ? cs2[n * C + c] : 1.f;
#pragma unroll
      for (int ix = 0; ix < MW; ++ix) {
        if (ix < W) {
          const size_t gi = (static_cast<size_t>(n) * H * W + iy * W + ix) * C + c;
          const float v = acc[ix] * s;
          if (dx) dx[gi] = __float2bfloat16(v);
          if (dx2) {
            const float m = ((mask_bits[gi >> 5] >> (gi & 31)) & 1u) ? 1.f : slope;
            dx2[gi] = __float2bfloat16(v * m * s2);
          }
        }
      }
    }
    // ---- dw: task = (c, ky), K x 5 accumulators in registers
    for (int task = threadIdx.x; task < C * K; task += blockDim.x) {
      const int c = task % C, ky = task / C;
      float acc[K][5];
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int o = 0; o < 5; ++o) acc[a][o] = 0.f;
      for (int oy = 0; oy < Ho; ++oy) {
        const int iy = oy + ky - PAD;
        if (iy < 0 || iy >= H) continue;
        float xr[MW + K];            // xr[j] = x[iy][j - PAD]
#pragma unroll
        for (int j = 0; j < MW + K; ++j) {
          const int ix = j - PAD;
          xr[j] = (ix >= 0 && ix < W) ? __bfloat162float(sX[(iy * W + ix) * C + c]) : 0.f;
        }
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          float dz[MW];
          const float4* dzr = reinterpret_cast<const float4*>(sDz + (o * Ho + oy) * MW);
#pragma unroll
          for (int i = 0; i < MW / 4; ++i) {
            const float4 t = dzr[i];
            dz[4 * i] = t.x; dz[4 * i + 1] = t.y; dz[4 * i + 2] = t.z; dz[4 * i + 3] = t.w;
          }
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < MW; ++ox) acc[kx][o] = fmaf(xr[ox + kx], dz[ox], acc[kx][o]);
        }
      }
      const float s = cs ? cs[n * C + c] : 1.f;
#pragma unroll
      for (int kx = 0; kx < K; ++kx)
#pragma unroll
        for (int o = 0; o < 5; ++o)
          atomicAdd(dw + (static_cast<size_t>(o) * C + c) * KK + ky * K + kx, acc[kx][o] * s);
    }
  }
}

// ============================================================================ maxpool 2x2
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
  return u;
}

// ============================================================================ head, C = 64 fast path
// grid = (B, kHeadParts), 256 threads.  Work item = (pixel, group of 8 channels): every weight fetch is one
// LDS.128 shared by the lanes of equal channel group, every activation fetch one LDS.128 of 8 bf16, so the
// inner loop is 40 FMA per 11-15 shared-memory instructions.  The transposed weight tile [tap][o][c] is
// staged with an XOR swizzle of the 8-channel groups (group ^ ((tap*5+o) & 7)): the transposing scatter of
// the coalesced [o][c][tap] read then hits 8 instead of 32 lanes per bank.
constexpr int kHeadParts = 4;
__device__ __forceinline__ int head_w_index(int to, int c) {      // to = tap*5 + o
  return to * 64 + ((((c >> 3) ^ to) & 7) << 3) + (c & 7);
}
__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}

// weights [5][64][K][K] fp32 -> w_t [K*K*5][64] fp32 in the swizzled layout of head_w_index (done once per
// step, so that every CTA stages its copy with plain coalesced float4 loads)
__global__ void head_pack_kernel(const float* __restrict__ w, int KK, float* __restrict__ wt) {
  pdl_trigger();
  pdl_wait();
  const int n = KK * 5 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i % KK, c = (i / KK) % 64, o = i / (KK * 64);
    wt[head_w_index(t * 5 + o, c)] = w[i];
  }
}
// channel of element j of the swizzled table
__device__ __forceinline__ int head_w_channel(int j) {
  const int to = j >> 6;
  return (((((j >> 3) & 7) ^ to) & 7) << 3) + (j & 7);
}

// Forward.  Register tiling: one thread = 5 consecutive output pixels x 5 outputs for one channel group (8 channels)
// and one tap residue, so each weight fetched from shared memory feeds 5 FMAs (a one-pixel-per-thread layout is
// bound by the shared-memory bandwidth of the weight reads, 1 FMA per 4 bytes).  One warp = (output row, strip of
// 5 pixels): lanes = (tap residue ts = lane/8, channel group g = lane%8); the 32 partial sums meet in a shuffle tree.
constexpr int kHeadStrip = 5;
template <int K, int PAD>
__global__ void __launch_bounds__(512)
head_fwd_c64_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ wt,
                    const float* __restrict__ bias, int H, int W, int Ho, int Wo, int rows_per_cta,
                    float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  constexpr int C = 64, KK = K * K, P = kHeadStrip;
  extern __shared__ float sm[];
  float* sW = sm;                                                       // [KK*5][64], pre-scaled by the dropout multiplier
  __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(sm + KK * 5 * C);  // [rows_per_cta + K - 1][W][64]
  const int n = blockIdx.x, oy0 = blockIdx.y * rows_per_cta;
  const int nrows_in = rows_per_cta + K - 1;
  for (int i = threadIdx.x; i < KK * 5 * C / 4; i += blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(wt) + i);
    if (cs) {
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(cs + n * C + head_w_channel(4 * i)));
      v.x *= s4.x; v.y *= s4.y; v.z *= s4.z; v.w *= s4.w;
    }
    reinterpret_cast<float4*>(sW)[i] = v;
  }
  for (int i = threadIdx.x; i < nrows_in * W * C / 8; i += blockDim.x) {
    const int r = i / (W * C / 8), rem = i - r * (W * C / 8);
    const int iy = oy0 + r - PAD;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H) v = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<size_t>(n) * H + iy) * W * C) + rem);
    reinterpret_cast<uint4*>(sX)[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & 7, ts = lane >> 3;
  const int strips = (Wo + P - 1) / P;
  const int ry = warp / strips, ox0 = (warp - ry * strips) * P;
  const int oy = oy0 + ry;
  if (ry >= rows_per_cta || oy >= Ho) return;
  float acc[P][5];
#pragma unroll
  for (int j = 0; j < P; ++j)
#pragma unroll
    for (int o = 0; o < 5; ++o) acc[j][o] = 0.f;
#pragma unroll 1
  for (int t = ts; t < KK; t += 4) {
    const int ky = t / K, kx = t - ky * K;
    float xv[P][8];
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int ix = ox0 + j + kx - PAD;
      uint4 u = make_uint4(0, 0, 0, 0);
      if (ix >= 0 && ix < W) u = *reinterpret_cast<const uint4*>(sX + ((ry + ky) * W + ix) * C + g * 8);
      bf16x8_to_f32(u, xv[j]);
    }
    const int to0 = t * 5;
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      const float4* wp = reinterpret_cast<const float4*>(sW + (to0 + o) * 64 + (((g ^ (to0 + o)) & 7) << 3));
      const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
      for (int j = 0; j < P; ++j) {
        float a = acc[j][o];
        a = fmaf(xv[j][0], w0.x, a); a = fmaf(xv[j][1], w0.y, a); a = fmaf(xv[j][2], w0.z, a); a = fmaf(xv[j][3], w0.w, a);
        a = fmaf(xv[j][4], w1.x, a); a = fmaf(xv[j][5], w1.y, a); a = fmaf(xv[j][6], w1.z, a); a = fmaf(xv[j][7], w1.w, a);
        acc[j][o] = a;
      }
    }
  }
  float mine = 0.f;
#pragma unroll
  for (int j = 0; j < P; ++j)
#pragma unroll
    for (int o = 0; o < 5; ++o) {
      float v = acc[j][o];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == j * 5 + o) mine = v;
    }
  if (lane < P * 5) {
    const int j = lane / 5, o = lane - j * 5;
    if (ox0 + j < Wo) {
      // bias == NULL: partial logits of one 64-channel plane of a wider head (no bias, no sigmoid; engine_planar sums them)
      y[(static_cast<size_t>(n) * 5 + o) * Ho * Wo + oy * Wo + ox0 + j] =
          bias ? 1.f / (1.f + expf(-(mine + __ldg(bias + o)))) : mine;
    }
  }
}

// Backward.  grid = (B, dx_parts + dw_parts).  CTAs with blockIdx.y < dx_parts produce `rows_per_cta` rows of dx
// with the register tiling of the forward (thread = 5 input pixels x 8 channels, lanes = (tap residue, channel
// group)); the others accumulate the weight gradient of kDwTaps taps each (one warp per (tap, half of the output
// pixels), thread = 5 outputs x 8 channels, lanes = (pixel residue, channel group)).
constexpr int kDwTaps = 18;
template <int K, int PAD>
__global__ void __launch_bounds__(512)
head_bwd_c64_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ cs, const float* __restrict__ wt,
                    const float* __restrict__ y, const float* __restrict__ dy, int H, int W, int Ho, int Wo,
                    int rows_per_cta, int dx_parts, __nv_bfloat16* __restrict__ dx,
                    const uint32_t* __restrict__ mask_bits, const float* __restrict__ cs2, float slope,
                    __nv_bfloat16* __restrict__ dx2, float* __restrict__ dw, float* __restrict__ dbias) {
  pdl_trigger();
  pdl_wait();
  constexpr int C = 64, KK = K * K, P = kHeadStrip;
  extern __shared__ float sm[];
  const int n = blockIdx.x;
  const int npo = Ho * Wo;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane & 7, ls = lane >> 3;
  float* sDz = sm;                                  // [5][npo] (+pad)
  const int dz_words = (5 * npo + 3) & ~3;
  for (int i = threadIdx.x; i < 5 * npo; i += blockDim.x) {
    const float yv = y[static_cast<size_t>(n) * 5 * npo + i];
    sDz[i] = dy[static_cast<size_t>(n) * 5 * npo + i] * yv * (1.f - yv);
  }
  if (static_cast<int>(blockIdx.y) < dx_parts) {
    // ------------------------------------------------------------------ dx rows
    float* sW = sm + dz_words;                      // [KK*5][64] swizzled, unscaled
    for (int i = threadIdx.x; i < KK * 5 * C / 4; i += blockDim.x)
      reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(wt) + i);
    __syncthreads();
    if (blockIdx.y == 0 && threadIdx.x < 5) {
      float t = 0.f;
      for (int i = 0; i < npo; ++i) t += sDz[threadIdx.x * npo + i];
      atomicAdd(dbias + threadIdx.x, t);
    }
    const int strips = (W + P - 1) / P;
    const int ry = warp / strips, ix0 = (warp - ry * strips) * P;
    const int iy = blockIdx.y * rows_per_cta + ry;
    if (ry >= rows_per_cta || iy >= H) return;
    float acc[P][8];
#pragma unroll
    for (int j = 0; j < P; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
#pragma unroll 1
    for (int t = ls; t < KK; t += 4) {
      const int ky = t / K, kx = t - ky * K;
      const int oy = iy - ky + PAD;
      if (oy < 0 || oy >= Ho) continue;
      const int to0 = t * 5;
#pragma unroll
      for (int o = 0; o < 5; ++o) {
        const float4* wp = reinterpret_cast<const float4*>(sW + (to0 + o) * 64 + (((g ^ (to0 + o)) & 7) << 3));
        const float4 w0 = wp[0], w1 = wp[1];
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const int ox = ix0 + j - kx + PAD;
          const float dz = (ox >= 0 && ox < Wo) ? sDz[o * npo + oy * Wo + ox] : 0.f;
          acc[j][0] = fmaf(dz, w0.x, acc[j][0]); acc[j][1] = fmaf(dz, w0.y, acc[j][1]);
          acc[j][2] = fmaf(dz, w0.z, acc[j][2]); acc[j][3] = fmaf(dz, w0.w, acc[j][3]);
          acc[j][4] = fmaf(dz, w1.x, acc[j][4]); acc[j][5] = fmaf(dz, w1.y, acc[j][5]);
          acc[j][6] = fmaf(dz, w1.z, acc[j][6]); acc[j][7] = fmaf(dz, w1.w, acc[j][7]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < P; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc[j][e] += __shfl_xor_sync(0xffffffffu, acc[j][e], 8);
        acc[j][e] += __shfl_xor_sync(0xffffffffu, acc[j][e], 16);
      }
    if (ls == 0) {
      float csv[8], cs2v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        csv[e] = cs ? __ldg(cs + n * C + g * 8 + e) : 1.f;
        cs2v[e] = cs2 ? __ldg(cs2 + n * C + g * 8 + e) : 1.f;
      }
#pragma unroll
      for (int j = 0; j < P; ++j) {
        const int ix = ix0 + j;
        if (ix >= W) break;
        const size_t gi = ((static_cast<size_t>(n) * H + iy) * W + ix) * C + g * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = acc[j][e] * csv[e];
        if (dx) *reinterpret_cast<uint4*>(dx + gi) = pack8(v);
        if (dx2) {
          const uint32_t mk = __ldg(mask_bits + (gi >> 5)) >> (gi & 31);
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = v[e] * (((mk >> e) & 1u) ? 1.f : slope) * cs2v[e];
          *reinterpret_cast<uint4*>(dx2 + gi) = pack8(v);
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ dw of kDwTaps taps
    const int t0 = (blockIdx.y - dx_parts) * kDwTaps;
    __nv_bfloat16* sX = reinterpret_cast<__nv_bfloat16*>(sm + dz_words);                 // [H*W][64]
    float* sAcc = reinterpret_cast<float*>(sX + static_cast<size_t>(H) * W * C);          // [kDwTaps][5][64]
    const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(n) * H * W * C);
    for (int i = threadIdx.x; i < H * W * C / 8; i += blockDim.x) reinterpret_cast<uint4*>(sX)[i] = __ldg(xs + i);
    for (int i = threadIdx.x; i < kDwTaps * 5 * C; i += blockDim.x) sAcc[i] = 0.f;
    __syncthreads();
    const int nwarps = blockDim.x >> 5;
    for (int wtask = warp; wtask < 2 * kDwTaps; wtask += nwarps) {      // warp task = (tap, half of the pixels)
      const int tl = wtask >> 1, half = wtask & 1;
      const int t = t0 + tl;
      if (t < KK) {
        const int ky = t / K, kx = t - ky * K;
        float acc[5][8];
#pragma unroll
        for (int o = 0; o < 5; ++o)
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[o][e] = 0.f;
        for (int po = half * 4 + ls; po < npo; po += 8) {
          const int oy = po / Wo, ox = po - oy * Wo;
          const int iy = oy + ky - PAD, ix = ox + kx - PAD;
          if (PAD > 0 && (iy < 0 || iy >= H || ix < 0 || ix >= W)) continue;
          float xv[8];
          bf16x8_to_f32(*reinterpret_cast<const uint4*>(sX + (iy * W + ix) * C + g * 8), xv);
#pragma unroll
          for (int o = 0; o < 5; ++o) {
            const float dz = sDz[o * npo + po];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[o][e] = fmaf(xv[e], dz, acc[o][e]);
          }
        }
#pragma unroll
        for (int o = 0; o < 5; ++o)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float v = acc[o][e];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (ls == 0) atomicAdd(sAcc + (tl * 5 + o) * C + g * 8 + e, v);
          }
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kDwTaps * 5 * C; i += blockDim.x) {
      const int c = i % C, o = (i / C) % 5, tl = i / (5 * C);
      if (t0 + tl < KK) {
        const float s = cs ? __ldg(cs + n * C + c) : 1.f;
        atomicAdd(dw + (static_cast<size_t>(o) * C + c) * KK + t0 + tl, sAcc[i] * s);
      }
    }
  }
}

// One CTA walks pooled rows (row = n * Ho + oy); its threads cover the (ox, 8-channel group) pairs of the row, so the
// only division per thread is the 32-bit row / Ho -- the former flat 64-bit index cost ~5 64-bit divisions per
// element and made these HBM-bound kernels instruction-bound.  C8 = C / 8 as a template constant (0 = runtime).
template <int kC8>
__global__ void maxpool2x2_fwd_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C,
                                      __nv_bfloat16* __restrict__ y, uint16_t* __restrict__ amax) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, C8 = kC8 ? kC8 : C / 8;
  const int rows = B * Ho, per_row = Wo * C8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, oy = row - n * Ho;
    const __nv_bfloat16* xrow = x + (static_cast<size_t>(n) * H + 2 * oy) * W * C;
    __nv_bfloat16* yrow = y + static_cast<size_t>(row) * Wo * C;
    for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
      const int ox = idx / C8, c8 = idx - ox * C8;
      const __nv_bfloat16* base = xrow + (2 * ox) * C + c8 * 8;
      float m[8], f[8];
      uint32_t arg = 0;                           // 2 bits per channel: position (dy*2 + dx) of the FIRST maximum
      unpack8(__ldg(reinterpret_cast<const uint4*>(base)), m);
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + C)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) if (f[j] > m[j]) { m[j] = f[j]; arg = (arg & ~(3u << (2 * j))) | (1u << (2 * j)); }
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(W) * C)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) if (f[j] > m[j]) { m[j] = f[j]; arg = (arg & ~(3u << (2 * j))) | (2u << (2 * j)); }
      unpack8(__ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(W) * C + C)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) if (f[j] > m[j]) { m[j] = f[j]; arg = (arg & ~(3u << (2 * j))) | (3u << (2 * j)); }
      if (amax) amax[(static_cast<size_t>(row) * Wo + ox) * C8 + c8] = static_cast<uint16_t>(arg);
      *reinterpret_cast<uint4*>(yrow + ox * C + c8 * 8) = pack8(m);
    }
  }
}

template <int kC8>
__global__ void maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gy, int B,
                                      int H, int W, int C, __nv_bfloat16* __restrict__ gs,
                                      const uint32_t* __restrict__ mask_bits, const float* __restrict__ cs,
                                      float slope, __nv_bfloat16* __restrict__ gs2, const uint16_t* __restrict__ amax) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, C8 = kC8 ? kC8 : C / 8;
  const int rows = B * Ho, per_row = Wo * C8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / Ho, oy = row - n * Ho;
    const size_t row_base = (static_cast<size_t>(n) * H + 2 * oy) * W * C;
    for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
      const int ox = idx / C8, c8 = idx - ox * C8;
      const size_t off[4] = {0, static_cast<size_t>(C), static_cast<size_t>(W) * C, static_cast<size_t>(W) * C + C};
      const size_t base = row_base + static_cast<size_t>(2 * ox) * C + c8 * 8;
      float g[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(gy + (static_cast<size_t>(row) * Wo + ox) * C + c8 * 8)), g);
      float sc[8];
      if (gs2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sc[j] = cs ? __ldg(cs + n * C + c8 * 8 + j) : 1.f;
      }
      int arg[8];
      if (amax) {                    // positions recorded by the forward: the pre-pool tensor is not read again
        const uint32_t a = __ldg(amax + (static_cast<size_t>(row) * Wo + ox) * C8 + c8);
#pragma unroll
        for (int j = 0; j < 8; ++j) arg[j] = (a >> (2 * j)) & 3;
      } else {
        float v[4][8];
#pragma unroll
        for (int k = 0; k < 4; ++k) unpack8(__ldg(reinterpret_cast<const uint4*>(x + base + off[k])), v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // first maximum in (dy,dx) row-major order, like ATen's max_pool2d
          int a = 0; float m = v[0][j];
#pragma unroll
          for (int k = 1; k < 4; ++k) if (v[k][j] > m) { m = v[k][j]; a = k; }
          arg[j] = a;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = arg[j] == k ? g[j] : 0.f;
        if (gs) *reinterpret_cast<uint4*>(gs + base + off[k]) = pack8(o);
        if (gs2) {
          const size_t e0 = base + off[k];               // element index; C % 32 == 0 keeps 8 channels in one word
          const uint32_t mk = __ldg(mask_bits + (e0 >> 5)) >> (e0 & 31);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = o[j] * (((mk >> j) & 1u) ? 1.f : slope) * sc[j];
          *reinterpret_cast<uint4*>(gs2 + base + off[k]) = pack8(o);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Elementwise halves of a residual block for backbones wider than the 64-channel tensor-core kernels (filters = 128:
// two 64-channel planes per tensor, engine.PlanarEngine).  A 128 -> 128 convolution is the SUM of two 64 -> 64
// convolutions per output plane (chained through the conv kernel's raw residual add), so the activation cannot ride in
// the conv epilogue and runs here:
//   act_mask  : out = lrelu(x) * chan_scale[n,c] + residual ; mask bit = sign bit of x clear   (PoolResnet.py:35-40)
//   grad_mask : out = g * (mask bit ? 1 : slope) * chan_scale[n,c]                              (its backward)
// x, g, residual, out: [B,HW,64] bf16 planes; mask: uint32 [B,HW,2]; 8 channels per thread.
__global__ void act_mask_kernel(const __nv_bfloat16* __restrict__ x, long n8, int hw, float slope,
                                const float* __restrict__ cs, const __nv_bfloat16* __restrict__ res,
                                uint8_t* __restrict__ mask_out, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
    const int c8 = static_cast<int>(i & 7);
    const long n = i / (8L * hw);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      bits |= (__float_as_uint(v[j]) >> 31 ? 0u : 1u) << j;
      v[j] = fmaxf(v[j], v[j] * slope);
      if (cs) v[j] *= __ldg(cs + n * 64 + c8 * 8 + j);
    }
    if (res) {
      float r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(res) + i), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (mask_out) mask_out[i] = static_cast<uint8_t>(bits);     // byte i = channels 8*c8.. of pixel i/8 (little endian words)
    reinterpret_cast<uint4*>(out)[i] = pack8(v);
  }
}

__global__ void grad_mask_kernel(const __nv_bfloat16* __restrict__ g, long n8, int hw, float slope,
                                 const uint8_t* __restrict__ mask, const float* __restrict__ cs,
                                 __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(g) + i), v);
    const int c8 = static_cast<int>(i & 7);
    const long n = i / (8L * hw);
    const uint32_t bits = mask ? mask[i] : 0xFFu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] *= ((bits >> j) & 1u) ? 1.f : slope;
      if (cs) v[j] *= __ldg(cs + n * 64 + c8 * 8 + j);
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(v);
  }
}

// Depthwise 3x3 (pad 1, no bias) + LeakyReLU on one 64-channel plane (models/SeparableCNN.py:20-27,45-46) for the
// separable backbone on channel planes (filters = 128); the 64-channel model runs the fused fd_sepblock_fwd instead.
// Thread = 8 channels of one pixel; the 9 x 64 tap-major fp32 weights sit in shared memory.
__global__ void __launch_bounds__(256)
dwconv3x3_lrelu_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, int B, int H, int W, float slope,
                       __nv_bfloat16* __restrict__ out) {
  __shared__ float sW[9 * 64];
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) sW[i] = __ldg(w + i);
  __syncthreads();
  const int rows = B * H, per_row = W * 8;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const int n = row / H, y = row - n * H;
    for (int idx = threadIdx.x; idx < per_row; idx += blockDim.x) {
      const int xx = idx >> 3, c8 = idx & 7;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xx + kx - 1;
          if (ix < 0 || ix >= W) continue;
          float v[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(x + ((static_cast<size_t>(n) * H + iy) * W + ix) * 64 + c8 * 8)), v);
          const float* wt = sW + (ky * 3 + kx) * 64 + c8 * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(wt[j], v[j], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], acc[j] * slope);
      *reinterpret_cast<uint4*>(out + (static_cast<size_t>(row) * W + xx) * 64 + c8 * 8) = pack8(acc);
    }
  }
}

inline int grid_for(long total, int block, int cap_mult = 8) {
  long g = (total + block - 1) / block;
  const long cap = static_cast<long>(sm_count()) * cap_mult;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" int fd_pack_conv3x3(const float* w, int n_layers, int C, fd_bf16* w_fwd, fd_bf16* w_dgrad, void* stream) {
  if (!w || n_layers <= 0 || C <= 0 || (!w_fwd && !w_dgrad)) return FD_EINVAL;
  const long total = static_cast<long>(n_layers) * 9 * C * C;
  (void)total;
  if (C == 64)
    launch_k(pack_conv3x3_kernel<64>, dim3(n_layers * (64 / kPackCo)), dim3(256), 0, static_cast<cudaStream_t>(stream), w,
             n_layers, reinterpret_cast<__nv_bfloat16*>(w_fwd), reinterpret_cast<__nv_bfloat16*>(w_dgrad));
  else if (C == 128)
    pack_conv3x3_kernel<128><<<n_layers * (128 / kPackCo), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        w, n_layers, reinterpret_cast<__nv_bfloat16*>(w_fwd), reinterpret_cast<__nv_bfloat16*>(w_dgrad));
  else
    return FD_EUNSUPPORTED;
  count_launch();
  return launch_status();
}

extern "C" int fd_unpack_wgrad3x3(const float* dw_packed, int n_layers, int C, float* dw, void* stream) {
  if (!dw_packed || !dw || n_layers <= 0 || C <= 0) return FD_EINVAL;
  const long total = static_cast<long>(n_layers) * 9 * C * C;
  (void)total;
  if (C == 64)
    launch_k(unpack_wgrad3x3_kernel<64>, dim3(n_layers * (64 / kPackCo)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             dw_packed, n_layers, dw);
  else if (C == 128)
    unpack_wgrad3x3_kernel<128><<<n_layers * (128 / kPackCo), 256, 0, static_cast<cudaStream_t>(stream)>>>(dw_packed, n_layers, dw);
  else
    return FD_EUNSUPPORTED;
  count_launch();
  return launch_status();
}

extern "C" int fd_unpack_wgrad3x3_planes(const float* dw_packed, int n_layers, int G, float* dw, void* stream) {
  if (!dw_packed || !dw || n_layers <= 0 || G <= 0) return FD_EINVAL;
  launch_k(unpack_wgrad3x3_planes_kernel, dim3(n_layers * G * G * (64 / kUnpT) * (64 / kUnpT)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), dw_packed, G, dw);
  count_launch();
  return launch_status();
}

extern "C" int fd_adam_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, int step, int32_t* state, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || (!state && step <= 0)) return FD_EINVAL;
  if (n % 4 != 0) return FD_EUNSUPPORTED;              // the flat buffers are padded to 16 bytes per section
  launch_k(adam_flat_kernel, dim3(grid_for(n / 4, 256, 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g, m, v, n,
           lr, beta1, beta2, eps, weight_decay, step, state);
  count_launch();
  return launch_status();
}

extern "C" int fd_dwconv3x3_lrelu(const fd_bf16* x, const float* w_dw, int B, int H, int W, int C, float slope, fd_bf16* out,
                                 void* stream) {
  if (!x || !w_dw || !out || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != 64 || !(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;
  {
    // strip kernel (4 pixels x 8 channels per thread, window rows loaded once); FD_DW_PIXEL_KERNEL=1 keeps the
    // thread-per-pixel kernel below for A/B runs
    static const bool pixel = getenv("FD_DW_PIXEL_KERNEL") != nullptr;
    if (!pixel) {
      const int rc = dwconv3x3_lrelu_strips(x, w_dw, B, H, W, slope, out, static_cast<cudaStream_t>(stream));
      if (rc != FD_EUNSUPPORTED) return rc;
    }
  }
  launch_k(dwconv3x3_lrelu_kernel, dim3(grid_for(static_cast<long>(B) * H, 1, 16)), dim3(256), 0,
           static_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x), w_dw, B, H, W, slope,
           reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_act_mask(const fd_bf16* x, int B, int HW, int C, float slope, const float* chan_scale,
                           const fd_bf16* residual, uint32_t* mask_out, fd_bf16* out, void* stream) {
  if (!x || !out || B <= 0 || HW <= 0) return FD_EINVAL;
  if (C != 64 || !(slope >= 0.f && slope <= 1.f)) return FD_EUNSUPPORTED;
  const long n8 = static_cast<long>(B) * HW * 8;
  launch_k(act_mask_kernel, dim3(grid_for(n8, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const __nv_bfloat16*>(x), n8, HW, slope, chan_scale,
           reinterpret_cast<const __nv_bfloat16*>(residual), reinterpret_cast<uint8_t*>(mask_out),
           reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_grad_mask(const fd_bf16* g, int B, int HW, int C, float slope, const uint32_t* mask_bits,
                            const float* chan_scale, fd_bf16* out, void* stream) {
  if (!g || !out || B <= 0 || HW <= 0) return FD_EINVAL;
  if (C != 64) return FD_EUNSUPPORTED;
  const long n8 = static_cast<long>(B) * HW * 8;
  launch_k(grad_mask_kernel, dim3(grid_for(n8, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const __nv_bfloat16*>(g), n8, HW, slope, reinterpret_cast<const uint8_t*>(mask_bits),
           chan_scale, reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_dropout_scale(const float* r, long n, long n_block, float keep_block, float keep_head, float* out,
                                void* stream) {
  if (!r || !out || n <= 0 || keep_block <= 0.f || keep_head <= 0.f) return FD_EINVAL;
  launch_k(dropout_scale_kernel, dim3(grid_for(n, 256, 4)), dim3(256), 0, static_cast<cudaStream_t>(stream), r, n, n_block,
           keep_block, keep_head, out);
  count_launch();
  return launch_status();
}

extern "C" long fd_stem_cache_elems(int B, int Cin, int Hin, int Win, int C, int K, int stride, int pad) {
  return stem_cache_elems(B, Cin, Hin, Win, C, K, stride, pad);
}

extern "C" int fd_stem_fwd(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin,
                           int Win, int C, int K, int stride, int pad, fd_bf16* y, fd_bf16* x_cache, void* stream) {
  if (!x || !w || !bias || !y || B <= 0) return FD_EINVAL;
  if (C % 64 != 0) return FD_EUNSUPPORTED;
  {
    const int rc = stem_fwd_tc(x, x_is_u8, w, bias, B, Cin, Hin, Win, C, K, stride, pad, y, x_cache,
                               static_cast<cudaStream_t>(stream));
    if (rc != FD_EUNSUPPORTED) return rc;      // tensor-core stem handled it (or failed for real)
  }
  {
    const int rc = stem_s2_fwd_tc(x, x_is_u8, w, bias, B, Cin, Hin, Win, C, K, stride, pad, y,
                                  static_cast<cudaStream_t>(stream));
    if (rc != FD_EUNSUPPORTED) return rc;      // the standard Resnet's 3x3 / stride-2 stem
  }
  const int Ho = (Hin + 2 * pad - K) / stride + 1, Wo = (Win + 2 * pad - K) / stride + 1;
  const int KK = Cin * K * K, pitch = (Wo - 1) * stride + K + 1;
  const size_t smem = (static_cast<size_t>(KK) * C + static_cast<size_t>(Cin) * K * pitch) * sizeof(float);
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(B * Ho, sm_count());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_fwd_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_fwd_kernel<uint8_t><<<grid, kStemThreads, smem, st>>>(static_cast<const uint8_t*>(x), w, bias, B, Cin, Hin, Win,
                                                               C, K, stride, pad, Ho, Wo,
                                                               reinterpret_cast<__nv_bfloat16*>(y));
  } else {
    e = set_max_dyn_smem(stem_fwd_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_fwd_kernel<float><<<grid, kStemThreads, smem, st>>>(static_cast<const float*>(x), w, bias, B, Cin, Hin, Win, C,
                                                             K, stride, pad, Ho, Wo,
                                                             reinterpret_cast<__nv_bfloat16*>(y));
  }
  count_launch();
  return launch_status();
}

extern "C" int fd_stem_fwd_cached(const fd_bf16* x_cache, const float* w, const float* bias, int B, int Cin, int Hin,
                                  int Win, int K, int stride, int pad, fd_bf16* y, void* stream) {
  if (!x_cache || !w || !bias || !y || B <= 0) return FD_EINVAL;
  return stem_fwd_cached(x_cache, w, bias, B, Cin, Hin, Win, 64, K, stride, pad, y, static_cast<cudaStream_t>(stream));
}

extern "C" int fd_stem_wgrad_pair(const fd_bf16* x_cache, const fd_bf16* g0, const fd_bf16* g1, int B, int Cin, int Hin,
                                  int Win, int K, int stride, int pad, float* dw, float* dbias, void* stream) {
  if (!x_cache || !g0 || !g1 || !dw || B <= 0) return FD_EINVAL;
  return stem_wgrad_pair_cached(x_cache, g0, g1, B, Cin, Hin, Win, K, stride, pad, dw, dbias,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int fd_stem_wgrad(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C,
                             int K, int stride, int pad, float* dw, float* dbias, const fd_bf16* x_cache, void* stream) {
  if (!x || !g || !dw || B <= 0) return FD_EINVAL;
  {
    const int rc = stem_wgrad_tc(x, x_is_u8, g, B, Cin, Hin, Win, C, K, stride, pad, dw, dbias, x_cache,
                                 static_cast<cudaStream_t>(stream));
    if (rc != FD_EUNSUPPORTED) return rc;
  }
  {
    const int rc = stem_s2_wgrad_tc(x, x_is_u8, g, B, Cin, Hin, Win, C, K, stride, pad, dw, dbias,
                                    static_cast<cudaStream_t>(stream));
    if (rc != FD_EUNSUPPORTED) return rc;
  }
  const int KK = Cin * K * K;
  if (C % 64 != 0 || KK > 16 * kStemKPT) return FD_EUNSUPPORTED;
  const int Ho = (Hin + 2 * pad - K) / stride + 1, Wo = (Win + 2 * pad - K) / stride + 1;
  const int pitch = (Wo - 1) * stride + K + 1;
  const size_t smem = (static_cast<size_t>(Wo) * C + static_cast<size_t>(Cin) * K * pitch) * sizeof(float);
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  const int grid = min(B * Ho, sm_count());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  if (x_is_u8) {
    e = set_max_dyn_smem(stem_wgrad_kernel<uint8_t>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_kernel<uint8_t><<<grid, kStemThreads, smem, st>>>(static_cast<const uint8_t*>(x),
                                                                 reinterpret_cast<const __nv_bfloat16*>(g), B, Cin, Hin,
                                                                 Win, C, K, stride, pad, Ho, Wo, dw, dbias);
  } else {
    e = set_max_dyn_smem(stem_wgrad_kernel<float>, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stem_wgrad_kernel<float><<<grid, kStemThreads, smem, st>>>(static_cast<const float*>(x),
                                                               reinterpret_cast<const __nv_bfloat16*>(g), B, Cin, Hin,
                                                               Win, C, K, stride, pad, Ho, Wo, dw, dbias);
  }
  count_launch();
  return launch_status();
}

extern "C" int fd_head_pack(const float* w, int C, int K, float* w_t, void* stream) {
  if (!w || !w_t || K <= 0) return FD_EINVAL;
  if (C != 64) return FD_EUNSUPPORTED;
  launch_k(head_pack_kernel, dim3((K * K * 5 * 64 + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), w, K * K,
           w_t);
  count_launch();
  return launch_status();
}

extern "C" int fd_head_fwd(const fd_bf16* x, const float* chan_scale, const float* w, const float* w_t,
                           const float* bias, int B, int H, int W, int C, int K, int pad, float* y, void* stream) {
  if (!x || !w || !y || B <= 0) return FD_EINVAL;
  if ((H * W * C) % 8 != 0) return FD_EUNSUPPORTED;
  const int Ho = H + 2 * pad - K + 1, Wo = W + 2 * pad - K + 1;
  if (Ho <= 0 || Wo <= 0) return FD_EINVAL;
  if (!bias && !(w_t && C == 64)) return FD_EUNSUPPORTED;      // partial-logit mode: 64-channel fast path only
  if (w_t && C == 64 && ((K == 6 && pad == 0) || (K == 3 && pad == 1))) {
    const int strips = (Wo + kHeadStrip - 1) / kHeadStrip;
    int rows_per_cta = 512 / 32 / strips;                    // one warp per (output row, strip of 5 pixels)
    if (rows_per_cta > (Ho + 1) / 2) rows_per_cta = (Ho + 1) / 2;
    const int parts = (Ho + rows_per_cta - 1) / rows_per_cta;
    const size_t sm2 = static_cast<size_t>(K) * K * 5 * C * 4 + static_cast<size_t>(rows_per_cta + K - 1) * W * C * 2;
    if (sm2 <= 227 * 1024 && rows_per_cta >= 1) {
      auto kern = (K == 6) ? head_fwd_c64_kernel<6, 0> : head_fwd_c64_kernel<3, 1>;
      cudaError_t e2 = set_max_dyn_smem(kern, (int)sm2);
      if (e2 != cudaSuccess) return (int)e2;
      launch_k(kern, dim3(B, parts), dim3(rows_per_cta * strips * 32), sm2, static_cast<cudaStream_t>(stream),
               reinterpret_cast<const __nv_bfloat16*>(x), chan_scale, w_t, bias, H, W, Ho, Wo, rows_per_cta, y);
      count_launch();
      return launch_status();
    }
  }
  const size_t smem = static_cast<size_t>((K * K * 5 * (C + 1) + 3) & ~3) * 4 + static_cast<size_t>(H) * W * C * 2;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(head_fwd_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  head_fwd_kernel<<<dim3(B, kHeadSplit), kHeadThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), chan_scale, w, bias, B, H, W, C, K, pad, Ho, Wo, y);
  count_launch();
  return launch_status();
}

extern "C" int fd_head_bwd(const fd_bf16* x, const float* chan_scale, const float* w, const float* w_t, const float* y,
                           const float* dy,
                           int B, int H, int W, int C, int K, int pad, fd_bf16* dx, const uint32_t* mask_bits,
                           const float* chan_scale2, float slope, fd_bf16* dx2, float* dw, float* dbias, void* stream) {
  if (!x || !w || !y || !dy || !dw || !dbias || B <= 0) return FD_EINVAL;
  if ((dx2 != nullptr) != (mask_bits != nullptr)) return FD_EINVAL;
  if (C % 32 != 0) return FD_EUNSUPPORTED;
  if ((H * W * C) % 8 != 0 || kHeadThreads % C != 0) return FD_EUNSUPPORTED;
  const int Ho = H + 2 * pad - K + 1, Wo = W + 2 * pad - K + 1;
  if (Ho <= 0 || Wo <= 0) return FD_EINVAL;
  {
    static const bool cuda_core_only = std::getenv("FD_HEAD_BWD_CUDA_CORES") != nullptr;     // A/B switch for tests
    if (!cuda_core_only) {
      const int rc = head_bwd_tc(x, chan_scale, w, y, dy, B, H, W, C, K, pad, dx, mask_bits, chan_scale2, slope, dx2, dw,
                                 dbias, static_cast<cudaStream_t>(stream));
      if (rc != FD_EUNSUPPORTED) return rc;      // the tensor-core kernel handled it (or failed for real)
    }
  }
  if (w_t && C == 64 && ((K == 6 && pad == 0) || (K == 3 && pad == 1))) {
    const int KK = K * K, dw_parts = (KK + kDwTaps - 1) / kDwTaps;
    const int strips = (W + kHeadStrip - 1) / kHeadStrip;
    int rows_per_cta = 512 / 32 / strips;
    if (rows_per_cta > (H + 2) / 3) rows_per_cta = (H + 2) / 3;
    const int dx_parts = (H + rows_per_cta - 1) / rows_per_cta;
    int threads = rows_per_cta * strips * 32;
    const size_t dz = static_cast<size_t>((5 * Ho * Wo + 3) & ~3) * 4;
    const size_t sm_dx = dz + static_cast<size_t>(KK) * 5 * C * 4;
    const size_t sm_dw = dz + static_cast<size_t>(H) * W * C * 2 + static_cast<size_t>(kDwTaps) * 5 * C * 4;
    const size_t sm2 = sm_dx > sm_dw ? sm_dx : sm_dw;
    if (sm2 <= 227 * 1024 && threads <= 512) {
      auto kern = (K == 6) ? head_bwd_c64_kernel<6, 0> : head_bwd_c64_kernel<3, 1>;
      cudaError_t e2 = set_max_dyn_smem(kern, (int)sm2);
      if (e2 != cudaSuccess) return (int)e2;
      launch_k(kern, dim3(B, dx_parts + dw_parts), dim3(threads), sm2, static_cast<cudaStream_t>(stream),
               reinterpret_cast<const __nv_bfloat16*>(x), chan_scale, w_t, y, dy, H, W, Ho, Wo, rows_per_cta, dx_parts,
               reinterpret_cast<__nv_bfloat16*>(dx), mask_bits, chan_scale2, slope, reinterpret_cast<__nv_bfloat16*>(dx2),
               dw, dbias);
      count_launch();
      return launch_status();
    }
  }
  const size_t smem = static_cast<size_t>((K * K * 5 * (C + 1) + 3) & ~3) * 4 + static_cast<size_t>((5 * Ho * Wo + 3) & ~3) * 4 +
                      static_cast<size_t>(H) * W * C * 2;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(head_bwd_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  head_bwd_kernel<<<min(B, sm_count()), kHeadThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), chan_scale, w, y, dy, B, H, W, C, K, pad, Ho, Wo,
      reinterpret_cast<__nv_bfloat16*>(dx), mask_bits, chan_scale2, slope,
      reinterpret_cast<__nv_bfloat16*>(dx2), dw, dbias);
  count_launch();
  return launch_status();
}

extern "C" int fd_maxpool2x2_fwd(const fd_bf16* x, int B, int H, int W, int C, fd_bf16* y, uint16_t* argmax,
                                 void* stream) {
  if (!x || !y || B <= 0) return FD_EINVAL;
  if (C % 8 != 0 || H < 2 || W < 2) return FD_EUNSUPPORTED;       // odd H / W: floor mode like nn.MaxPool2d(2) (SSD.py:80, 15 -> 7)
  const long rows = static_cast<long>(B) * (H / 2);
  if (rows > 0x7fffffffL) return FD_EUNSUPPORTED;
  const int per_row = (W / 2) * (C / 8);
  const int threads = per_row >= 256 ? 256 : (per_row + 31) / 32 * 32;
  launch_k(C == 64 ? maxpool2x2_fwd_kernel<8> : maxpool2x2_fwd_kernel<0>, dim3(grid_for(rows, 1, 16)), dim3(threads), 0,
           static_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x), B, H, W, C,
           reinterpret_cast<__nv_bfloat16*>(y), argmax);
  count_launch();
  return launch_status();
}

extern "C" int fd_maxpool2x2_bwd(const fd_bf16* x, const fd_bf16* gy, int B, int H, int W, int C, fd_bf16* gs,
                                 const uint32_t* mask_bits, const float* chan_scale, float slope, fd_bf16* gs2,
                                 const uint16_t* argmax, void* stream) {
  if ((!x && !argmax) || !gy || (!gs && !gs2) || B <= 0) return FD_EINVAL;
  if ((gs2 != nullptr) != (mask_bits != nullptr)) return FD_EINVAL;
  if (gs2 && C % 32 != 0) return FD_EUNSUPPORTED;
  if (C % 8 != 0 || H < 2 || W < 2) return FD_EUNSUPPORTED;
  if ((H | W) & 1) {      // floor mode: the last row / column belongs to no window -> zero gradient (the kernel writes windows only)
    const size_t bytes = static_cast<size_t>(B) * H * W * C * 2;
    cudaError_t e = cudaSuccess;
    if (gs) e = cudaMemsetAsync(gs, 0, bytes, static_cast<cudaStream_t>(stream));
    if (e == cudaSuccess && gs2) e = cudaMemsetAsync(gs2, 0, bytes, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) { (void)cudaGetLastError(); return static_cast<int>(e); }
  }
  const long rows = static_cast<long>(B) * (H / 2);
  if (rows > 0x7fffffffL) return FD_EUNSUPPORTED;
  const int per_row = (W / 2) * (C / 8);
  const int threads = per_row >= 256 ? 256 : (per_row + 31) / 32 * 32;
  launch_k(C == 64 ? maxpool2x2_bwd_kernel<8> : maxpool2x2_bwd_kernel<0>, dim3(grid_for(rows, 1, 16)), dim3(threads), 0,
           static_cast<cudaStream_t>(stream), reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(gy), B, H, W, C,
           reinterpret_cast<__nv_bfloat16*>(gs), mask_bits, chan_scale, slope, reinterpret_cast<__nv_bfloat16*>(gs2),
           argmax);
  count_launch();
  return launch_status();
}
