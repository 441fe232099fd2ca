// Small data-movement kernels around the backbone engines:
//   fd_resize_bilinear  transforms.Resize of models/BaseModel.py:64 / models/PoolResnet.py:94-95 (the `predict == 1`
//                       branch): bilinear, align_corners = False, no antialias; uint8 or fp32 NCHW
//   fd_index_copy_f32   dst[idx[i]] = src[i] / dst[i] = src[idx[i]]: parameters of models narrower than the 64-channel
//                       kernel planes are scattered into zero-padded planes, their gradients gathered back
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

// ATen upsample_bilinear2d semantics (align_corners = false): src = scale * (dst + 0.5) - 0.5 clamped at 0,
// scale = in / out; neighbours i0 = floor(src), i1 = i0 + (i0 < in - 1); weights (1 - l, l).
template <typename T>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const T* __restrict__ x, long planes, int h, int w, int H, int W, float sy, float sx,
                       T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const long total = planes * H * W;
  for (long i = blockIdx.x * 256L + threadIdx.x; i < total; i += static_cast<long>(gridDim.x) * 256L) {
    const int ox = static_cast<int>(i % W);
    const int oy = static_cast<int>((i / W) % H);
    const long pl = i / (static_cast<long>(W) * H);
    // separate multiply / subtract (no FMA contraction): the same roundings as ATen's CPU kernel, the parity oracle
    float fy = __fsub_rn(__fmul_rn(sy, static_cast<float>(oy) + 0.5f), 0.5f);
    float fx = __fsub_rn(__fmul_rn(sx, static_cast<float>(ox) + 0.5f), 0.5f);
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
    y0 = y0 > h - 1 ? h - 1 : y0;
    x0 = x0 > w - 1 ? w - 1 : x0;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = fy - static_cast<float>(y0), lx = fx - static_cast<float>(x0);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const T* p = x + pl * h * w;
    const float v00 = static_cast<float>(p[static_cast<long>(y0) * w + x0]), v01 = static_cast<float>(p[static_cast<long>(y0) * w + x1]);
    const float v10 = static_cast<float>(p[static_cast<long>(y1) * w + x0]), v11 = static_cast<float>(p[static_cast<long>(y1) * w + x1]);
    const float t0 = __fadd_rn(__fmul_rn(hx, v00), __fmul_rn(lx, v01));
    const float t1 = __fadd_rn(__fmul_rn(hx, v10), __fmul_rn(lx, v11));
    const float v = __fadd_rn(__fmul_rn(hy, t0), __fmul_rn(ly, t1));
    if (sizeof(T) == 1) {
      const float r = rintf(v);                                   // torch.round: half to even, then the uint8 cast
      out[i] = static_cast<T>(r < 0.f ? 0.f : (r > 255.f ? 255.f : r));
    } else {
      out[i] = static_cast<T>(v);
    }
  }
}

__global__ void __launch_bounds__(256)
index_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ idx, long n, int scatter) {
  pdl_trigger();
  pdl_wait();
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * 256L) {
    if (scatter) dst[idx[i]] = src[i];
    else dst[i] = src[idx[i]];
  }
}

// out = g * (ref >= +0 ? 1 : slope): the LeakyReLU' step of a backward chain, with the sign taken from the saved
// ACTIVATION itself (lrelu keeps the sign of its argument), 8 bf16 per thread.
__global__ void __launch_bounds__(256)
lrelu_bwd_kernel(const uint4* __restrict__ g, const uint4* __restrict__ ref, long n8, float slope, uint4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  for (long i = blockIdx.x * 256L + threadIdx.x; i < n8; i += static_cast<long>(gridDim.x) * 256L) {
    const uint4 a = __ldg(g + i), r = __ldg(ref + i);
    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, rv[4] = {r.x, r.y, r.z, r.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float lo = bf16lo(av[k]) * ((rv[k] & 0x8000u) ? slope : 1.f);
      const float hi = bf16hi(av[k]) * ((rv[k] & 0x80000000u) ? slope : 1.f);
      o[k] = pack_bf16x2(lo, hi);
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Weight gradient of a depthwise 3x3 (pad 1) convolution on one 64-channel NHWC plane:
//   dw[c][ky][kx] += sum_{n,y,x} g[n,y,x,c] * x[n,y+ky-1,x+kx-1,c]
// One warp walks output rows, lane = channel pair; 18 accumulators per lane, block reduction, 576 atomics per block.
__global__ void __launch_bounds__(256)
dw3x3_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ g, int B, int H, int W,
                   float* __restrict__ dw) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_acc[8][18 * 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[9][2];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t][0] = acc[t][1] = 0.f;
  const long rows = static_cast<long>(B) * H;
  for (long row = blockIdx.x * 8L + warp; row < rows; row += gridDim.x * 8L) {
    const int y = static_cast<int>(row % H);
    const long n = row / H;
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(g + (n * H + y) * static_cast<long>(W) * 64) + lane;
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + n * H * static_cast<long>(W) * 64) + lane;
    for (int xx = 0; xx < W; ++xx) {
      const uint32_t gv = __ldg(gp + xx * 32);
      const float g0 = bf16lo(gv), g1 = bf16hi(gv);
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = xx + kx - 1;
          if (ix < 0 || ix >= W) continue;
          const uint32_t xv = __ldg(xb + (static_cast<long>(iy) * W + ix) * 32);
          acc[ky * 3 + kx][0] = fmaf(g0, bf16lo(xv), acc[ky * 3 + kx][0]);
          acc[ky * 3 + kx][1] = fmaf(g1, bf16hi(xv), acc[ky * 3 + kx][1]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    s_acc[warp][(t * 2 + 0) * 32 + lane] = acc[t][0];
    s_acc[warp][(t * 2 + 1) * 32 + lane] = acc[t][1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 18 * 32; i += 256) {
    float v = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) v += s_acc[w8][i];
    const int l = i & 31, half = (i >> 5) & 1, t = i >> 6;
    const int c = 2 * l + half;
    if (v != 0.f) atomicAdd(dw + c * 9 + t, v);
  }
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" int fd_lrelu_bwd(const fd_bf16* g, const fd_bf16* ref, long n, float slope, fd_bf16* out, void* stream) {
  if (!g || !ref || !out || n <= 0) return FD_EINVAL;
  if (n % 8) return FD_EUNSUPPORTED;
  long blocks = (n / 8 + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  launch_k(lrelu_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const uint4*>(g), reinterpret_cast<const uint4*>(ref), n / 8, slope, reinterpret_cast<uint4*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_dwconv3x3_wgrad(const fd_bf16* x, const fd_bf16* g, int B, int H, int W, int C, float* dw, void* stream) {
  if (!x || !g || !dw || B <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  if (C != 64) return FD_EUNSUPPORTED;
  long blocks = (static_cast<long>(B) * H + 7) / 8;
  if (blocks > 148L * 4) blocks = 148L * 4;
  launch_k(dw3x3_wgrad_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
           reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(g), B, H, W, dw);
  count_launch();
  return launch_status();
}

extern "C" int fd_resize_bilinear(const void* x, int is_u8, long planes, int h, int w, int H, int W, void* out, void* stream) {
  if (!x || !out || planes <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return FD_EINVAL;
  const float sy = static_cast<float>(h) / static_cast<float>(H), sx = static_cast<float>(w) / static_cast<float>(W);
  const long total = planes * H * W;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 32) blocks = 148L * 32;
  if (is_u8)
    launch_k(resize_bilinear_kernel<uint8_t>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const uint8_t*>(x), planes, h, w, H, W, sy, sx, static_cast<uint8_t*>(out));
  else
    launch_k(resize_bilinear_kernel<float>, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream),
             static_cast<const float*>(x), planes, h, w, H, W, sy, sx, static_cast<float*>(out));
  count_launch();
  return launch_status();
}

extern "C" int fd_index_copy_f32(float* dst, const float* src, const int32_t* idx, long n, int scatter, void* stream) {
  if (!dst || !src || !idx || n <= 0) return FD_EINVAL;
  long blocks = (n + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  launch_k(index_copy_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), dst, src,
           idx, n, scatter);
  count_launch();
  return launch_status();
}
