// Host-side plumbing of libfd_b200.so: driver entry point for TMA descriptors, launch counter,
// error strings.
#include <mutex>
#include <unordered_map>

#include "fd_host.h"

#include <cstdlib>

namespace fd {

std::atomic<long long> g_launches{0};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

int make_tmap_nhwc_bf16(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int boxW, int boxH) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  if (C * 2 != 128 || boxW < 1 || boxW > 256 || boxH < 1 || boxH > 256) return FD_EUNSUPPORTED;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

int make_tmap_nhwc_bf16_strided(CUtensorMap* m, const void* ptr, int B, int H, int Wext, int Wfull, int C, int boxW,
                                int boxH) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  if (C * 2 != 128 || boxW < 1 || boxW > 256 || boxH < 1 || boxH > 256 || Wext < 1) return FD_EUNSUPPORTED;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wext, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Wfull * C * 2, (cuuint64_t)H * Wfull * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

int make_tmap_2d_bf16(CUtensorMap* m, const void* ptr, int rows, int cols, int boxRows, int boxCols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

int make_tmap_2d_f32(CUtensorMap* m, const void* ptr, long rows, int cols, int boxRows, int boxCols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  if (boxCols * 4 > 128 || boxRows > 256) return FD_EUNSUPPORTED;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)boxCols, (cuuint32_t)boxRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

int make_tmap_xbf(CUtensorMap* m, const void* ptr, int planes, int Hin, int K) {
  // [planes][Hin][img 2][512] bf16 viewed as 5-D {256, 2, 2, Hin, planes}; box {256, 2, 2, K, 1}, no swizzle, zero OOB fill
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  cuuint64_t dims[5] = {256, 2, 2, (cuuint64_t)Hin, (cuuint64_t)planes};
  cuuint64_t strides[4] = {512, 1024, 2048, (cuuint64_t)Hin * 2048};
  cuuint32_t box[5] = {256, 2, 2, (cuuint32_t)K, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

int make_tmap_3d(CUtensorMap* m, const void* ptr, int elem_bytes, int is_u8, int d0, int d1, int d2, int box0,
                 int box1, int box2) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return FD_EDRIVER;
  if (box0 > 256 || box1 > 256 || box2 < 1 || box2 > 256 || (box0 * elem_bytes) % 16 != 0) return FD_EUNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * elem_bytes, (cuuint64_t)d0 * d1 * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, (cuuint32_t)box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, is_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FD_OK : FD_EINVAL;
}

cudaError_t raise_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, int> high;
  std::lock_guard<std::mutex> lock(mu);
  int& h = high[func];
  if (bytes <= h) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) h = bytes;
  return e;
}

bool pdl_enabled() {
  static const bool on = std::getenv("FD_NO_PDL") == nullptr;
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace fd

extern "C" {

int fd_version(void) { return 100; }

long long fd_launch_count(void) { return fd::g_launches.load(std::memory_order_relaxed); }

const char* fd_error_string(int code) {
  switch (code) {
    case FD_OK: return "ok";
    case FD_EINVAL: return "invalid argument";
    case FD_EUNSUPPORTED: return "unsupported shape";
    case FD_EDRIVER: return "cuTensorMapEncodeTiled not available from the driver";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}
}
