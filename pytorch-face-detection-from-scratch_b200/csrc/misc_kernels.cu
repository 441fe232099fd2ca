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

}  // namespace
}  // namespace fd

using namespace fd;

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
