// Per-scale prediction heads of the SSD model (reference models/SSD.py:174-177,238-254): Linear(C -> 5) on every pixel
// of a feature map, sigmoid on the score column, prior scaling / offsets of apply_priors (:206-220), written straight
// into the concatenated [B, 4774, 5] output -- forward and backward.  8 MFLOP per image in total: latency / HBM bound,
// CUDA cores.  The feature map arrives as 64-channel NHWC bf16 planes (C = 64 * G), the layout of the convolution engine.
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kMaxPlanes = 8;
struct Planes {
  const __nv_bfloat16* p[kMaxPlanes];
};
struct PlanesOut {
  __nv_bfloat16* p[kMaxPlanes];
};

// one warp per pixel; lane l holds channels {2l, 2l+1} of every plane
__global__ void __launch_bounds__(256)
ssd_head_fwd_kernel(Planes x, int G, const float* __restrict__ w, const float* __restrict__ bias, long npix, int HW, int C,
                    const float* __restrict__ mult, const float* __restrict__ priors, int prior_off, int P,
                    float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * 256L + threadIdx.x) >> 5, nwarp = (gridDim.x * 256L) >> 5;
  for (long pix = warp0; pix < npix; pix += nwarp) {
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int g = 0; g < G; ++g) {
      const uint32_t xv = __ldg(reinterpret_cast<const uint32_t*>(x.p[g] + pix * 64) + lane);
      const float x0 = bf16lo(xv), x1 = bf16hi(xv);
      const int c = g * 64 + 2 * lane;
      if (c < C) {
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          const float2 wv = __ldg(reinterpret_cast<const float2*>(w + o * C + c));
          acc[o] = fmaf(x0, wv.x, fmaf(x1, wv.y, acc[o]));
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
      v += __ldg(bias + lane);
      const long b = pix / HW;
      const int r = prior_off + static_cast<int>(pix - b * HW);
      if (lane == 0) v = 1.f / (1.f + expf(-v));                                   // SSD.py:245 sigmoid on the scores only
      else {
        if (lane <= 2) v = __fmul_rn(v, __ldg(mult + r));                          // :210-215  x, y *= 1 / ps
        v = __fadd_rn(v, __ldg(priors + r * 4 + lane - 1));                        // :216      x[..., 1:5] += priors
      }
      out[(b * P + r) * 5 + lane] = v;
    }
  }
}

// backward: dz0 = dout0 * s (1 - s) ; dz1,2 = dout1,2 * mult ; dz3,4 = dout3,4
//   dx[c] = sum_o dz[o] w[o][c]  (bf16 planes, overwritten)   dw[o][c] += sum_pix dz[o] x[c]   db[o] += sum_pix dz[o]
__global__ void __launch_bounds__(256)
ssd_head_bwd_kernel(Planes x, PlanesOut dx, int G, const float* __restrict__ w, long npix, int HW, int C,
                    const float* __restrict__ mult, int prior_off, int P, const float* __restrict__ out,
                    const float* __restrict__ dout, float* __restrict__ dw, float* __restrict__ db) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const long warp0 = (blockIdx.x * 256L + threadIdx.x) >> 5, nwarp = (gridDim.x * 256L) >> 5;
  float wacc[kMaxPlanes][5][2];
  float bacc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int g = 0; g < kMaxPlanes; ++g)
#pragma unroll
    for (int o = 0; o < 5; ++o) wacc[g][o][0] = wacc[g][o][1] = 0.f;
  for (long pix = warp0; pix < npix; pix += nwarp) {
    const long b = pix / HW;
    const int r = prior_off + static_cast<int>(pix - b * HW);
    const float* dp = dout + (b * P + r) * 5;
    const float s = __ldg(out + (b * P + r) * 5);
    const float m = __ldg(mult + r);
    float dz[5];
    dz[0] = __ldg(dp) * s * (1.f - s);
    dz[1] = __ldg(dp + 1) * m;
    dz[2] = __ldg(dp + 2) * m;
    dz[3] = __ldg(dp + 3);
    dz[4] = __ldg(dp + 4);
#pragma unroll
    for (int o = 0; o < 5; ++o) bacc[o] += dz[o];
#pragma unroll
    for (int g = 0; g < kMaxPlanes; ++g) {
      if (g < G) {
        const int c = g * 64 + 2 * lane;
        const uint32_t xv = __ldg(reinterpret_cast<const uint32_t*>(x.p[g] + pix * 64) + lane);
        const float x0 = bf16lo(xv), x1 = bf16hi(xv);
        float d0 = 0.f, d1 = 0.f;
        if (c < C) {
#pragma unroll
          for (int o = 0; o < 5; ++o) {
            const float2 wv = __ldg(reinterpret_cast<const float2*>(w + o * C + c));
            d0 = fmaf(dz[o], wv.x, d0);
            d1 = fmaf(dz[o], wv.y, d1);
            wacc[g][o][0] = fmaf(dz[o], x0, wacc[g][o][0]);
            wacc[g][o][1] = fmaf(dz[o], x1, wacc[g][o][1]);
          }
        }
        reinterpret_cast<uint32_t*>(dx.p[g] + pix * 64)[lane] = pack_bf16x2(d0, d1);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < kMaxPlanes; ++g) {
    if (g < G) {
      const int c = g * 64 + 2 * lane;
      if (c < C) {
#pragma unroll
        for (int o = 0; o < 5; ++o) {
          if (wacc[g][o][0] != 0.f) atomicAdd(dw + o * C + c, wacc[g][o][0]);
          if (wacc[g][o][1] != 0.f) atomicAdd(dw + o * C + c + 1, wacc[g][o][1]);
        }
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int o = 0; o < 5; ++o)
      if (bacc[o] != 0.f) atomicAdd(db + o, bacc[o]);
  }
}

}  // namespace
}  // namespace fd

using namespace fd;

extern "C" int fd_ssd_head_fwd(const fd_bf16* const* x_planes, int G, const float* w, const float* bias, int B, int HW,
                               int C, const float* mult, const float* priors, int prior_off, int P, float* out,
                               void* stream) {
  if (!x_planes || !w || !bias || !mult || !priors || !out || B <= 0 || HW <= 0 || G <= 0) return FD_EINVAL;
  if (G > kMaxPlanes || C > G * 64 || C % 2) return FD_EUNSUPPORTED;
  Planes x = {};
  for (int g = 0; g < G; ++g) {
    if (!x_planes[g]) return FD_EINVAL;
    x.p[g] = reinterpret_cast<const __nv_bfloat16*>(x_planes[g]);
  }
  const long npix = static_cast<long>(B) * HW;
  long blocks = (npix + 7) / 8;
  if (blocks > 148L * 8) blocks = 148L * 8;
  launch_k(ssd_head_fwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, G,
           w, bias, npix, HW, C, mult, priors, prior_off, P, out);
  count_launch();
  return launch_status();
}

extern "C" int fd_ssd_head_bwd(const fd_bf16* const* x_planes, fd_bf16* const* dx_planes, int G, const float* w, int B,
                               int HW, int C, const float* mult, int prior_off, int P, const float* out,
                               const float* dout, float* dw, float* db, void* stream) {
  if (!x_planes || !dx_planes || !w || !mult || !out || !dout || !dw || !db || B <= 0 || HW <= 0 || G <= 0) return FD_EINVAL;
  if (G > kMaxPlanes || C > G * 64 || C % 2) return FD_EUNSUPPORTED;
  Planes x = {};
  PlanesOut dx = {};
  for (int g = 0; g < G; ++g) {
    if (!x_planes[g] || !dx_planes[g]) return FD_EINVAL;
    x.p[g] = reinterpret_cast<const __nv_bfloat16*>(x_planes[g]);
    dx.p[g] = reinterpret_cast<__nv_bfloat16*>(dx_planes[g]);
  }
  const long npix = static_cast<long>(B) * HW;
  long blocks = (npix + 7) / 8;
  if (blocks > 148L * 2) blocks = 148L * 2;
  launch_k(ssd_head_bwd_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, dx,
           G, w, npix, HW, C, mult, prior_off, P, out, dout, dw, db);
  count_launch();
  return launch_status();
}
