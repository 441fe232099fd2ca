// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/fd_b200.h"

namespace fd {

extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// cudaError_t -> ABI return code, consuming the sticky "last error" of a failed launch.
inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? FD_OK : static_cast<int>(e);
}

// Launch with programmatic dependent launch (PDL): the kernel may start (prologue: barrier init, TMEM alloc,
// shared-memory clears) while the tail of the previous kernel of the stream is still running; every kernel calls
// griddepcontrol.wait before its first global-memory access.  Captured into CUDA graphs as programmatic edges.
// FD_NO_PDL=1 falls back to ordinary stream order.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Opt-in dynamic shared memory of a kernel, as a HIGH-WATER MARK: the attribute belongs to the function, not to a launch, so
// lowering it for a later, smaller launch would invalidate kernel nodes already captured into CUDA graphs with the larger
// size (seen as LaunchFailed when a graph is replayed under a profiler).  Only ever raised; thread safe.
cudaError_t raise_dyn_smem(const void* func, int bytes);
template <typename K>
inline cudaError_t set_max_dyn_smem(K kern, int bytes) {
  return raise_dyn_smem(reinterpret_cast<const void*>(kern), bytes);
}

// [B,H,W,C] bf16 tensor as a 4-D tiled TMA map with box {C, boxW, boxH, 1}, 128B swizzle, zero OOB fill.
int make_tmap_nhwc_bf16(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int boxW, int boxH);
// a W-window [B,H,Wext,C] of a [B,H,Wfull,C] bf16 tensor (ptr = first pixel of the window): columns >= Wext are
// out of bounds (zero-filled on load) although they exist in memory
int make_tmap_nhwc_bf16_strided(CUtensorMap* m, const void* ptr, int B, int H, int Wext, int Wfull, int C, int boxW,
                                int boxH);
// [rows, cols] bf16 row-major as a 2-D tiled TMA map with box {boxCols, boxRows}, 128B swizzle.
int make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, int rows, int cols, int boxRows, int boxCols);

// [rows, cols] fp32 row-major as a 2-D tiled map with box {boxCols (<= 32), boxRows}, 128B swizzle
int make_tmap_2d_f32(CUtensorMap* m, const void* ptr, long rows, int cols, int boxRows, int boxCols);

// [d2, d1, d0] tensor of 1- or 4-byte elements (uint8 / fp32) as a 3-D tiled map, box {box0, box1, 1}, no swizzle
int make_tmap_3d(CUtensorMap* m, const void* ptr, int elem_bytes, int is_u8, int d0, int d1, int d2, int box0, int box1,
                 int box2 = 1);

// the stem's bf16 image copy [planes][Hin][2][512] as a 5-D map, box = K rows of one plane (both image slots)
int make_tmap_xbf(CUtensorMap* m, const void* ptr, int planes, int Hin, int K);

long stem_cache_elems(int B, int Cin, int Hin, int Win, int C, int K, int stride, int pad);

int sm_count();

// tcgen05 stem (stem_tc.cu); FD_EUNSUPPORTED means "not the stride-8 stem shape, use the generic kernel"
int stem_fwd_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win,
                int C, int K, int stride, int pad, fd_bf16* y, fd_bf16* xbf, cudaStream_t st);
int stem_wgrad_tc(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                  int stride, int pad, float* dw, float* dbias, const fd_bf16* xbf, cudaStream_t st);
int stem_fwd_cached(const fd_bf16* xbf, const float* w, const float* bias, int B, int Cin, int Hin, int Win, int C, int K,
                    int stride, int pad, fd_bf16* y, cudaStream_t st);
int stem_wgrad_pair_cached(const fd_bf16* xbf, const fd_bf16* g0, const fd_bf16* g1, int B, int Cin, int Hin, int Win,
                           int K, int stride, int pad, float* dw, float* dbias, cudaStream_t st);

// tcgen05 stem of the standard Resnet (3x3, stride 2, pad 1, 3 -> 64; stem_s2_tc.cu); FD_EUNSUPPORTED = other shape
int stem_s2_fwd_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win, int C,
                   int K, int stride, int pad, fd_bf16* y, cudaStream_t st);
int stem_s2_wgrad_tc(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                     int stride, int pad, float* dw, float* dbias, cudaStream_t st);

// tcgen05 stem of the MobilenetV3 backbone (3x3, stride 2, 3 -> 16, Hardswish; mbv3_stem_tc.cu); FD_EUNSUPPORTED = other shape
int mbv3_stem_tc(const void* x, int x_is_u8, const float* w, const float* bias, int B, int H, int W, int pad_t, int pad_l,
                 int Ho, int Wo, fd_bf16* out, cudaStream_t st);

// conv3x3_wide.cu: the cta_group::2 convolution for nout = 128 (fd_conv3x3_wide) or 64 output channels (one plane)
int conv3x3_pairs(int nout, const fd_bf16* const* x, int gin, const fd_bf16* w_packed, int B, int H, int W, const float* bias,
                  float slope, const float* const* chan_scale, const fd_bf16* const* residual, uint32_t* const* mask_out,
                  fd_bf16* const* out, const uint32_t* const* mask_in, const float* const* chan_scale2, fd_bf16* const* out2,
                  int flags, void* stream);

// depthwise 3x3 pad 1 + LeakyReLU on a 64-channel plane through the strip kernel of mbv3_kernels.cu (w: [9][64] tap-major)
int dwconv3x3_lrelu_strips(const fd_bf16* x, const float* w, int B, int H, int W, float slope, fd_bf16* out, cudaStream_t st);

// tcgen05 backward of the 64 -> 5 head (head_tc.cu); FD_EUNSUPPORTED = shape not instantiated
int head_bwd_tc(const fd_bf16* x, const float* chan_scale, const float* w, const float* y, const float* dy, int B, int H,
                int W, int C, int K, int pad, fd_bf16* dx, const uint32_t* mask_bits, const float* chan_scale2, float slope,
                fd_bf16* dx2, float* dw, float* dbias, cudaStream_t st);

}  // namespace fd
