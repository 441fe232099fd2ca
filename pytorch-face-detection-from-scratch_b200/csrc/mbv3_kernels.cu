// CUDA-core kernels of the MobilenetV3 backbone (reference models/MobilenetV3Backbone.py:33-60: timm
// tf_mobilenetv3_small_100 minus its last five children + Conv2d(576 -> 5, 3x3, pad 1) + sigmoid), inference.
// The graph is the one stored in the official TorchScript archive (code/__torch__/timm/models/efficientnet_blocks.py):
// BatchNorm (eps 1e-3) is folded into the preceding convolution by the caller; what remains besides the pointwise
// GEMMs (pw_gemm_tc.cu) is memory-bound and has no tensor-core form:
//
//   mbv3_stem_kernel   Conv2dSame 3x3 stride 2, 3 -> 16, + bias + Hardswish; fp32 / uint8 NCHW in, NHWC bf16 out
//   mbv3_dw_kernel     depthwise 3x3 / 5x5, stride 1 / 2, TF "SAME" asymmetric padding, + bias + ReLU / Hardswish,
//                      optionally accumulating the SqueezeExcite channel sums of its own output
//   mbv3_se_kernel     SqueezeExcite gate: mean -> 1x1 reduce + ReLU -> 1x1 expand -> Hardsigmoid
//   mbv3_scale_kernel  x *= gate[n, c]
//   mbv3_head_kernel   3x3 pad 1 conv C -> 5 + sigmoid (C = 576 does not fit the shared-memory head of layers.cu)
#include <cstdlib>

#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v * fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);
  return v;
}

// ------------------------------------------------------------------------------------------------ stem
// thread = one output pixel x 16 channels; weights [27 taps (c,ky,kx)][16] fp32 in shared memory (broadcast reads)
template <bool kU8>
__global__ void __launch_bounds__(256)
mbv3_stem_kernel(const void* __restrict__ xin, const float* __restrict__ w, const float* __restrict__ bias, int B, int H,
                 int W, int Ho, int Wo, int pad_t, int pad_l, __nv_bfloat16* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float sw[27 * 16];
  __shared__ __align__(16) float sb[16];
  for (int i = threadIdx.x; i < 27 * 16; i += 256) {
    const int co = i & 15, t = i >> 4;             // w: [co][c][ky][kx] -> sw[t = c*9+ky*3+kx][co]
    sw[i] = __ldg(w + co * 27 + t);
  }
  if (threadIdx.x < 16) sb[threadIdx.x] = __ldg(bias + threadIdx.x);
  __syncthreads();
  const long total = static_cast<long>(B) * Ho * Wo;
  const long pix = blockIdx.x * 256L + threadIdx.x;
  if (pix >= total) return;
  const int ox = static_cast<int>(pix % Wo);
  const int oy = static_cast<int>((pix / Wo) % Ho);
  const long n = pix / (static_cast<long>(Wo) * Ho);
  uint64_t acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = pk2(sb[2 * i], sb[2 * i + 1]);
  const int iy0 = oy * 2 - pad_t, ix0 = ox * 2 - pad_l;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = iy0 + ky;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ix0 + kx;
        if (ix < 0 || ix >= W) continue;
        const long gi = ((n * 3 + c) * H + iy) * W + ix;
        float xv;
        if (kU8) {
          // x / 255 exactly as IEEE division (the reference divides, PoolResnet.py:95 / MobilenetV3Backbone.py:52)
          xv = __fdiv_rn(static_cast<float>(static_cast<const uint8_t*>(xin)[gi]), 255.f);
        } else {
          xv = __ldg(static_cast<const float*>(xin) + gi);
        }
        const uint64_t x2 = pk2(xv, xv);
        const float4* wp = reinterpret_cast<const float4*>(sw + (c * 9 + ky * 3 + kx) * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 w4 = wp[i];
          acc[2 * i] = fma2(x2, pk2(w4.x, w4.y), acc[2 * i]);
          acc[2 * i + 1] = fma2(x2, pk2(w4.z, w4.w), acc[2 * i + 1]);
        }
      }
    }
  }
  uint32_t u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a, b;
    upk2(acc[i], a, b);
    u[i] = pack_bf16x2(act_apply(a, 2), act_apply(b, 2));
  }
  uint4* dst = reinterpret_cast<uint4*>(out + pix * 16);
  dst[0] = make_uint4(u[0], u[1], u[2], u[3]);
  dst[1] = make_uint4(u[4], u[5], u[6], u[7]);
}

// ------------------------------------------------------------------------------------------------ depthwise
// thread = a strip of kDwT horizontally adjacent output pixels x 8 channels (one 16-byte vector per pixel); grid.y = image.
// Per kernel row the K weight vectors and the (T-1)*S+K input vectors of the strip are loaded ONCE and reused by all
// T outputs (a thread-per-pixel version re-loaded 3 vectors per tap and was L1/LSU bound at ~0.07 of HBM).
// w: [K*K][C] fp32 tap-major.
constexpr int kDwT = 4;
template <int K, int S>
__global__ void __launch_bounds__(256, 2)
mbv3_dw_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int H,
               int W, int C, int Ho, int Wo, int pad_t, int pad_l, int act, __nv_bfloat16* __restrict__ out,
               float* __restrict__ se_partial, float lrelu_slope) {
  pdl_trigger();
  pdl_wait();
  constexpr int T = kDwT, NX = (T - 1) * S + K;
  __shared__ float s_val[256 * 8];            // this block's bf16-rounded outputs (SqueezeExcite partial sums)
  const int n = blockIdx.y;
  const int C8 = C >> 3;
  const int strips = (Wo + T - 1) / T;
  const long items = static_cast<long>(Ho) * strips * C8;
  const long idx = blockIdx.x * 256L + threadIdx.x;
  float ssum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (idx < items) {
    const int cg = static_cast<int>(idx % C8);
    const long rest = idx / C8;
    const int strip = static_cast<int>(rest % strips), oy = static_cast<int>(rest / strips);
    const int ox0 = strip * T;
    const int c0 = cg * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    uint64_t acc[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) {
      acc[t][0] = pk2(b0.x, b0.y); acc[t][1] = pk2(b0.z, b0.w); acc[t][2] = pk2(b1.x, b1.y); acc[t][3] = pk2(b1.z, b1.w);
    }
    const __nv_bfloat16* xn = x + static_cast<long>(n) * H * W * C + c0;
    const int iy0 = oy * S - pad_t, ix0 = ox0 * S - pad_l;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int iy = iy0 + ky;
      if (iy < 0 || iy >= H) continue;
      uint64_t wv[K][4];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float* wp = w + (ky * K + kx) * C + c0;
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
        wv[kx][0] = pk2(w0.x, w0.y); wv[kx][1] = pk2(w0.z, w0.w); wv[kx][2] = pk2(w1.x, w1.y); wv[kx][3] = pk2(w1.z, w1.w);
      }
      const __nv_bfloat16* xr = xn + static_cast<long>(iy) * W * C;
      uint4 xv[NX];
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        const int ix = ix0 + j;
        xv[j] = (ix >= 0 && ix < W) ? __ldg(reinterpret_cast<const uint4*>(xr + static_cast<long>(ix) * C)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        const uint64_t x0 = pk2u(xv[j].x << 16, xv[j].x & 0xFFFF0000u), x1 = pk2u(xv[j].y << 16, xv[j].y & 0xFFFF0000u);
        const uint64_t x2 = pk2u(xv[j].z << 16, xv[j].z & 0xFFFF0000u), x3 = pk2u(xv[j].w << 16, xv[j].w & 0xFFFF0000u);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int kx = j - t * S;                       // compile-time after unrolling
          if (kx >= 0 && kx < K) {
            acc[t][0] = fma2(x0, wv[kx][0], acc[t][0]);
            acc[t][1] = fma2(x1, wv[kx][1], acc[t][1]);
            acc[t][2] = fma2(x2, wv[kx][2], acc[t][2]);
            acc[t][3] = fma2(x3, wv[kx][3], acc[t][3]);
          }
        }
      }
    }
    __nv_bfloat16* orow = out + ((static_cast<long>(n) * Ho + oy) * Wo) * C + c0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      if (ox0 + t < Wo) {
        float r[8];
        upk2(acc[t][0], r[0], r[1]); upk2(acc[t][1], r[2], r[3]); upk2(acc[t][2], r[4], r[5]); upk2(acc[t][3], r[6], r[7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = act == 3 ? fmaxf(r[i], r[i] * lrelu_slope) : act_apply(r[i], act);
        uint4 o;
        o.x = pack_bf16x2(r[0], r[1]); o.y = pack_bf16x2(r[2], r[3]);
        o.z = pack_bf16x2(r[4], r[5]); o.w = pack_bf16x2(r[6], r[7]);
        *reinterpret_cast<uint4*>(orow + static_cast<long>(ox0 + t) * C) = o;
        // SqueezeExcite averages the tensor the next layer READS, i.e. the bf16-rounded values
        ssum[0] += bf16lo(o.x); ssum[1] += bf16hi(o.x); ssum[2] += bf16lo(o.y); ssum[3] += bf16hi(o.y);
        ssum[4] += bf16lo(o.z); ssum[5] += bf16hi(o.z); ssum[6] += bf16lo(o.w); ssum[7] += bf16hi(o.w);
      }
    }
  }
  if (se_partial) {
    // DETERMINISTIC block partial (no atomics: inference must be bit-reproducible run to run): channel c is summed over
    // the block's threads that hold channel group c / 8, in thread order; se_partial[n][block][c] is a plain store and
    // fd_se_gate adds the blocks in block order.
#pragma unroll
    for (int e = 0; e < 8; ++e) s_val[threadIdx.x * 8 + e] = ssum[e];
    __syncthreads();
    const int first_cg = static_cast<int>((blockIdx.x * 256L) % C8);      // channel group of thread 0
    for (int c = threadIdx.x; c < C; c += 256) {
      const int cg = c >> 3, e = c & 7;
      int t = cg - first_cg;
      if (t < 0) t += C8;
      float a = 0.f;
      for (; t < 256; t += C8) a += s_val[t * 8 + e];
      se_partial[(static_cast<long>(n) * gridDim.x + blockIdx.x) * C + c] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------ SqueezeExcite gate
// one CTA per image: mean[c] = (sum over the depthwise kernel's block partials) / HW ; r = relu(W1 mean + b1) ;
// gate = hardsigmoid(W2 r + b2).
constexpr int kSeThreads = 512;
__global__ void __launch_bounds__(kSeThreads)
mbv3_se_kernel(const float* __restrict__ partial, int nblk, float inv_hw, const float* __restrict__ w1,
               const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2, int C, int R,
               float* __restrict__ gate) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];          // mean[C] | r[R]
  float* s_mean = sm;
  float* s_r = sm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += kSeThreads) {
    const float* pp = partial + static_cast<long>(n) * nblk * C + c;
    // 8 independent chains (8 loads in flight instead of one L2 round trip per block), combined in a FIXED order
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int k = 0;
    for (; k + 8 <= nblk; k += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += __ldg(pp + static_cast<long>(k + u) * C);
    }
    for (int u = 0; k < nblk; ++k, ++u) a[u] += __ldg(pp + static_cast<long>(k) * C);
    s_mean[c] = (((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))) * inv_hw;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int j = warp; j < R; j += kSeThreads / 32) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w1 + static_cast<long>(j) * C + c), s_mean[c], a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) s_r[j] = fmaxf(a + __ldg(b1 + j), 0.f);
  }
  __syncthreads();
  for (int c = warp; c < C; c += kSeThreads / 32) {
    float a = 0.f;
    for (int j = lane; j < R; j += 32) a = fmaf(__ldg(w2 + static_cast<long>(c) * R + j), s_r[j], a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if (lane == 0) {
      const float v = a + __ldg(b2 + c);
      gate[static_cast<long>(n) * C + c] = fminf(fmaxf(v + 3.f, 0.f), 6.f) * (1.f / 6.f);     // Hardsigmoid: relu6(x+3)/6
    }
  }
}

// x[n, p, c] *= gate[n, c] in place (bf16 x, fp32 gate), 8 channels per thread
__global__ void __launch_bounds__(256)
mbv3_scale_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ gate, long n8, int HW, int C) {
  pdl_trigger();
  pdl_wait();
  const int C8 = C >> 3;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * 256L) {
    const int cg = static_cast<int>(i % C8);
    const long n = i / (static_cast<long>(C8) * HW);
    const float* g = gate + n * C + cg * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g)), g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
    uint4 v = *reinterpret_cast<uint4*>(x + i * 8);
    v.x = pack_bf16x2(bf16lo(v.x) * g0.x, bf16hi(v.x) * g0.y);
    v.y = pack_bf16x2(bf16lo(v.y) * g0.z, bf16hi(v.y) * g0.w);
    v.z = pack_bf16x2(bf16lo(v.z) * g1.x, bf16hi(v.z) * g1.y);
    v.w = pack_bf16x2(bf16lo(v.w) * g1.z, bf16hi(v.w) * g1.w);
    *reinterpret_cast<uint4*>(x + i * 8) = v;
  }
}

// ------------------------------------------------------------------------------------------------ head
// 3x3 pad-1 conv C -> 5 + bias + sigmoid (MobilenetV3Backbone.py:40-46,57-58).  grid = (B, row groups); one warp per
// strip of 4 horizontally adjacent output pixels, lanes over channel pairs; weights bf16 [9][5][C] in shared memory:
// the five weight pairs of a (tap, channel pair) are fetched once and feed the four pixels (40 FMAs per 5 LDS + 4 LDG).
// y: [B,5,H,W] fp32.
constexpr int kHeadPix = 4;
__global__ void __launch_bounds__(256)
mbv3_head_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int H,
                 int W, int C, float* __restrict__ y) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __nv_bfloat16 s_w[];      // [9][5][C]
  for (int i = threadIdx.x; i < 9 * 5 * C; i += 256) {
    const int c = i % C, o = (i / C) % 5, t = i / (5 * C);       // w: [o][c][ky][kx]
    s_w[i] = __float2bfloat16_rn(__ldg(w + (static_cast<long>(o) * C + c) * 9 + t));
  }
  __syncthreads();
  const int n = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const __nv_bfloat16* xn = x + static_cast<long>(n) * H * W * C;
  const int strips = (W + kHeadPix - 1) / kHeadPix;
  for (int item = blockIdx.y * 8 + warp; item < H * strips; item += gridDim.y * 8) {
    const int oy = item / strips, ox0 = (item - oy * strips) * kHeadPix;
    float acc[kHeadPix][5];
#pragma unroll
    for (int p = 0; p < kHeadPix; ++p)
#pragma unroll
      for (int o = 0; o < 5; ++o) acc[p][o] = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy + ky - 1;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(s_w + (ky * 3 + kx) * 5 * C);
        const uint32_t* xrow = reinterpret_cast<const uint32_t*>(xn + static_cast<long>(iy) * W * C);
        bool ok[kHeadPix];
#pragma unroll
        for (int p = 0; p < kHeadPix; ++p) {
          const int ix = ox0 + p + kx - 1;
          ok[p] = ix >= 0 && ix < W;
        }
        for (int c2 = lane; c2 < (C >> 1); c2 += 32) {
          float w0[5], w1[5];
#pragma unroll
          for (int o = 0; o < 5; ++o) {
            const uint32_t wv = wp[o * (C >> 1) + c2];
            w0[o] = bf16lo(wv);
            w1[o] = bf16hi(wv);
          }
#pragma unroll
          for (int p = 0; p < kHeadPix; ++p) {
            if (ok[p]) {
              const uint32_t xv = __ldg(xrow + static_cast<long>(ox0 + p + kx - 1) * (C >> 1) + c2);
              const float x0 = bf16lo(xv), x1 = bf16hi(xv);
#pragma unroll
              for (int o = 0; o < 5; ++o) acc[p][o] = fmaf(x0, w0[o], fmaf(x1, w1[o], acc[p][o]));
            }
          }
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kHeadPix; ++p) {
#pragma unroll
      for (int o = 0; o < 5; ++o) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc[p][o] += __shfl_xor_sync(0xffffffffu, acc[p][o], d);
      }
      if (lane < 5 && ox0 + p < W) {
        float v = lane == 0 ? acc[p][0] : lane == 1 ? acc[p][1] : lane == 2 ? acc[p][2] : lane == 3 ? acc[p][3] : acc[p][4];
        v += __ldg(bias + lane);
        y[((static_cast<long>(n) * 5 + lane) * H + oy) * W + ox0 + p] = 1.f / (1.f + expf(-v));
      }
    }
  }
}

// depthwise weights fp32 [C][1][K][K] * scale[c] (folded BatchNorm, nullable) -> [K*K][C] fp32 tap-major
__global__ void mbv3_dw_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale, int C, int KK,
                                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * KK; i += gridDim.x * blockDim.x) {
    const int c = i % C, t = i / C;
    out[i] = w[c * KK + t] * (scale ? scale[c] : 1.f);
  }
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" int fd_mbv3_stem(const void* x, int x_is_u8, const float* w, const float* bias, int B, int H, int W, int pad_t,
                            int pad_l, int Ho, int Wo, fd_bf16* out, void* stream) {
  if (!x || !w || !bias || !out || B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return FD_EINVAL;
  {
    // tensor-core path (im2col patch tile in shared memory, mbv3_stem_tc.cu); the CUDA-core kernel below serves image
    // widths whose row pitch TMA cannot address (FD_MBV3_STEM_SIMT=1 forces it, for A/B runs)
    static const bool simt = getenv("FD_MBV3_STEM_SIMT") != nullptr;
    if (!simt) {
      const int rc = mbv3_stem_tc(x, x_is_u8, w, bias, B, H, W, pad_t, pad_l, Ho, Wo, out, static_cast<cudaStream_t>(stream));
      if (rc != FD_EUNSUPPORTED) return rc;
    }
  }
  const long total = static_cast<long>(B) * Ho * Wo;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (x_is_u8)
    launch_k(mbv3_stem_kernel<true>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, w, bias, B, H, W, Ho,
             Wo, pad_t, pad_l, reinterpret_cast<__nv_bfloat16*>(out));
  else
    launch_k(mbv3_stem_kernel<false>, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(stream), x, w, bias, B, H, W, Ho,
             Wo, pad_t, pad_l, reinterpret_cast<__nv_bfloat16*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_dw_pack(const float* w, const float* scale, int C, int K, float* out, void* stream) {
  if (!w || !out || C <= 0 || K <= 0) return FD_EINVAL;
  launch_k(mbv3_dw_pack_kernel, dim3((C * K * K + 255) / 256), dim3(256), 0, static_cast<cudaStream_t>(stream), w, scale,
           C, K * K, out);
  count_launch();
  return launch_status();
}

extern "C" int fd_dwconv_se_blocks(int Ho, int Wo, int C) {
  if (Ho <= 0 || Wo <= 0 || C <= 0 || C % 8) return -1;
  return static_cast<int>((static_cast<long>(Ho) * ((Wo + kDwT - 1) / kDwT) * (C / 8) + 255) / 256);
}

extern "C" int fd_dwconv(const fd_bf16* x, const float* w_packed, const float* bias, int B, int H, int W, int C, int K,
                         int stride, int pad_t, int pad_l, int Ho, int Wo, int act, fd_bf16* out, float* se_sum,
                         void* stream) {
  if (!x || !w_packed || !bias || !out || B <= 0 || H <= 0 || W <= 0 || Ho <= 0 || Wo <= 0) return FD_EINVAL;
  if (C % 8 || (K != 3 && K != 5) || (stride != 1 && stride != 2) || act < 0 || act > 2 || B > 65535) return FD_EUNSUPPORTED;
  const long items = static_cast<long>(Ho) * ((Wo + kDwT - 1) / kDwT) * (C / 8);
  const dim3 grid(static_cast<unsigned>((items + 255) / 256), B);
  const size_t smem = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(out);
#define FD_DW(KK, SS) launch_k(mbv3_dw_kernel<KK, SS>, grid, dim3(256), smem, st, xp, w_packed, bias, H, W, C, Ho, Wo, \
                               pad_t, pad_l, act, op, se_sum, 0.f)
  if (K == 3 && stride == 1) FD_DW(3, 1);
  else if (K == 3) FD_DW(3, 2);
  else if (stride == 1) FD_DW(5, 1);
  else FD_DW(5, 2);
#undef FD_DW
  count_launch();
  return launch_status();
}

// Depthwise 3x3 pad 1 + LeakyReLU (0 <= slope <= 1) on a 64-channel plane, no bias (fd_dwconv3x3_lrelu of layers.cu):
// the strip kernel above with act code 3.
__device__ float g_dw_zero_bias[64];
int fd::dwconv3x3_lrelu_strips(const fd_bf16* x, const float* w, int B, int H, int W, float slope, fd_bf16* out, cudaStream_t st) {
  if (B > 65535) return FD_EUNSUPPORTED;
  float* zb = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&zb), g_dw_zero_bias) != cudaSuccess) return FD_EUNSUPPORTED;
  const long items = static_cast<long>(H) * ((W + kDwT - 1) / kDwT) * 8;
  const dim3 grid(static_cast<unsigned>((items + 255) / 256), B);
  launch_k(mbv3_dw_kernel<3, 1>, grid, dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(x), w,
           static_cast<const float*>(zb), H, W, 64, H, W, 1, 1, 3, reinterpret_cast<__nv_bfloat16*>(out),
           static_cast<float*>(nullptr), slope);
  count_launch();
  return launch_status();
}

extern "C" int fd_se_gate(const float* se_partial, int nblk, int B, int HW, const float* w1, const float* b1, const float* w2,
                          const float* b2, int C, int R, float* gate, void* stream) {
  if (!se_partial || !w1 || !b1 || !w2 || !b2 || !gate || B <= 0 || HW <= 0 || C <= 0 || R <= 0 || nblk <= 0) return FD_EINVAL;
  launch_k(mbv3_se_kernel, dim3(B), dim3(kSeThreads), static_cast<size_t>(C + R) * 4, static_cast<cudaStream_t>(stream), se_partial,
           nblk, 1.f / static_cast<float>(HW), w1, b1, w2, b2, C, R, gate);
  count_launch();
  return launch_status();
}

extern "C" int fd_scale_channels(fd_bf16* x, const float* gate, int B, int HW, int C, void* stream) {
  if (!x || !gate || B <= 0 || HW <= 0 || C <= 0) return FD_EINVAL;
  if (C % 8) return FD_EUNSUPPORTED;
  const long n8 = static_cast<long>(B) * HW * (C / 8);
  long blocks = (n8 + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  launch_k(mbv3_scale_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<__nv_bfloat16*>(x), gate, n8, HW, C);
  count_launch();
  return launch_status();
}

extern "C" int fd_head3x3_fwd(const fd_bf16* x, const float* w, const float* bias, int B, int H, int W, int C, float* y,
                              void* stream) {
  if (!x || !w || !bias || !y || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C % 64 || B > 65535) return FD_EUNSUPPORTED;
  const size_t smem = static_cast<size_t>(9) * 5 * C * 2;
  if (smem > 200 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(mbv3_head_kernel, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int parts = 2;
  launch_k(mbv3_head_kernel, dim3(B, parts), dim3(256), smem, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const __nv_bfloat16*>(x), w, bias, H, W, C, y);
  count_launch();
  return launch_status();
}
