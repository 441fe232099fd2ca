// The grid-head side of the hot path: YoloLoss forward+backward, decode + score threshold +
// NMS, and grid-cell assignment.  One CTA per image, everything in shared memory; results that
// the reference rounds or indexes (box corners, kept indices, cell assignment) are bit-exact:
// every float operation below that feeds a rounded value uses the explicit round-to-nearest
// intrinsics so that nvcc cannot contract a*b+c into an FMA.
#include "fd_host.h"
#include "fd_ptx.cuh"

namespace fd {
namespace {

constexpr int kYoloThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* scratch) {
  // deterministic: fixed shuffle tree, then warp partials added in warp order by every thread
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += scratch[w];
  return t;
}

// ---------------------------------------------------------------------------- YoloLoss
// losses/YoloLoss.py:4-44.  pred/gt [B,5,S1,S2].
__global__ void __launch_bounds__(kYoloThreads)
yolo_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int S1, int S2,
                 float* __restrict__ loss, const float* __restrict__ dscale, float* __restrict__ dpred) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ float scratch[kYoloThreads / 32];
  const int b = blockIdx.x;
  const int ncell = S1 * S2;
  const float* p = pred + static_cast<size_t>(b) * 5 * ncell;
  const float* g = gt + static_cast<size_t>(b) * 5 * ncell;
  float* d = dpred ? dpred + static_cast<size_t>(b) * 5 * ncell : nullptr;

  // YoloLoss.py:8-9: NaNs are replaced by 0.1 only when nansum(pred) != 0
  float ns = 0.f;
  for (int i = threadIdx.x; i < 5 * ncell; i += blockDim.x) {
    const float v = p[i];
    if (v == v) ns += v;
  }
  ns = block_sum(ns, scratch);
  const bool fix_nan = ns != 0.f;
  const float w_no = 1.0f / static_cast<float>(S1);  // YoloLoss.py:25
  const float ds = dscale ? dscale[b] : 1.f;

  float acc = 0.f;
  for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
    float pv[5];
    bool wasnan[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      float v = p[k * ncell + c];
      wasnan[k] = fix_nan && !(v == v);
      pv[k] = wasnan[k] ? 0.1f : v;
    }
    const float g0 = g[c], g1 = g[ncell + c], g2 = g[2 * ncell + c], g3 = g[3 * ncell + c], g4 = g[4 * ncell + c];
    // YoloLoss.py:17-18: gt_x,gt_y = gt[1],gt[2]; pred_y,pred_x = pred[1],pred[2]
    const float ex = g1 - pv[2], ey = g2 - pv[1];
    const float sp3 = sqrtf(pv[3]), sp4 = sqrtf(pv[4]);
    const float ew = sqrtf(g3) - sp3, eh = sqrtf(g4) - sp4;
    const float ec = g0 - pv[0];
    const float cw = g0 + (1.f - g0) * w_no;
    const float xy = 3.f * g0 * (ex * ex + ey * ey);
    const float wh = 3.f * g0 * (ew * ew + eh * eh);
    acc += xy + wh + cw * (ec * ec);
    if (d) {
      float dv[5];
      dv[0] = -2.f * cw * ec;
      dv[1] = -6.f * g0 * ey;
      dv[2] = -6.f * g0 * ex;
      dv[3] = -3.f * g0 * ew / sp3;   // p3 == 0 -> 0 * inf = NaN, exactly like autograd of p**0.5
      dv[4] = -3.f * g0 * eh / sp4;
#pragma unroll
      for (int k = 0; k < 5; ++k) d[k * ncell + c] = wasnan[k] ? 0.f : dv[k] * ds;
    }
  }
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) loss[b] = acc;
}

// ---------------------------------------------------------------------------- decode + NMS
// datasets/utils.py:157-170 + torchvision nms.  Dynamic smem layout (ncell = S1*S2, nw = ceil(ncell/32)):
//   float  sc[ncell], bx[4][ncell]   candidates in row-major cell order
//   int    cell[ncell], order[ncell] (sorted position -> candidate)
//   uint   mask[ncell][nw]           suppression bit-matrix in sorted order
__global__ void __launch_bounds__(kYoloThreads)
decode_nms_kernel(const float* __restrict__ pred, int S1, int S2, float p_thr, double iou_thr, float psx, float psy,
                  float width, float height, float* __restrict__ out_boxes, int* __restrict__ out_cell,
                  int* __restrict__ out_count) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ uint8_t smraw[];
  const int ncell = S1 * S2;
  const int nw = (ncell + 31) / 32;
  float* sc = reinterpret_cast<float*>(smraw);
  float* bx = sc + ncell;                       // [4][ncell]
  int* cell = reinterpret_cast<int*>(bx + 4 * ncell);
  int* order = cell + ncell;
  uint32_t* mask = reinterpret_cast<uint32_t*>(order + ncell);   // [ncell][nw]
  __shared__ int s_warp_cnt[kYoloThreads / 32];
  __shared__ int s_base, s_K, s_nkeep;
  __shared__ int s_keep[1024];

  const int b = blockIdx.x;
  const float* p = pred + static_cast<size_t>(b) * 5 * ncell;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // 1. ordered compaction of the cells with conf > thr (utils.py:112,119), chunk by chunk
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < ncell; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const float conf = c < ncell ? p[c] : 0.f;
    const bool pass = c < ncell && conf > p_thr;
    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
    if (pass) {
      const int k = off + __popc(bal & ((1u << lane) - 1u));
      const int i = c / S2, j = c - i * S2;
      // utils.py:122-125: X = x1*ps_x + i*ps_x (first spatial index pairs with x), W = x3*width
      const float X = __fadd_rn(__fmul_rn(p[ncell + c], psx), __fmul_rn(static_cast<float>(i), psx));
      const float Y = __fadd_rn(__fmul_rn(p[2 * ncell + c], psy), __fmul_rn(static_cast<float>(j), psy));
      const float Wd = __fmul_rn(p[3 * ncell + c], width);
      const float Hd = __fmul_rn(p[4 * ncell + c], height);
      sc[k] = conf;                                   // utils.py:163 un-rounded score
      bx[0 * ncell + k] = rintf(X);                   // utils.py:162 round half to even
      bx[1 * ncell + k] = rintf(Y);
      bx[2 * ncell + k] = rintf(__fadd_rn(Wd, X));    // utils.py:153-154 corners from un-rounded values
      bx[3 * ncell + k] = rintf(__fadd_rn(Hd, Y));
      cell[k] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = s_base;
      for (int w = 0; w < nwarps; ++w) t += s_warp_cnt[w];
      s_base = t;
    }
    __syncthreads();
  }
  const int K = s_base;
  if (K == 0) {
    if (threadIdx.x == 0) out_count[b] = 0;
    return;
  }

  // 2. stable descending sort by rank counting (ties: lower candidate index first)
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const float si = sc[i];
    int r = 0;
    for (int j = 0; j < K; ++j) {
      const float sj = sc[j];
      r += (sj > si) || (sj == si && j < i);
    }
    order[r] = i;
  }
  __syncthreads();

  // 3. suppression bit-matrix in sorted order: bit (a, b>a) set iff IoU(a,b) > thr
  const int Kw = (K + 31) / 32;
  for (int idx = threadIdx.x; idx < K * Kw; idx += blockDim.x) {
    const int a = idx / Kw, wq = idx - a * Kw;
    const int ia = order[a];
    const float ax1 = bx[ia], ay1 = bx[ncell + ia], ax2 = bx[2 * ncell + ia], ay2 = bx[3 * ncell + ia];
    const float aarea = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    uint32_t bits = 0;
    for (int t = 0; t < 32; ++t) {
      const int bpos = wq * 32 + t;
      if (bpos <= a || bpos >= K) continue;
      const int ib = order[bpos];
      const float bx1 = bx[ib], by1 = bx[ncell + ib], bx2 = bx[2 * ncell + ib], by2 = bx[3 * ncell + ib];
      const float barea = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
      const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
      const float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
      const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
      if (static_cast<double>(ovr) > iou_thr) bits |= 1u << t;   // NaN (0/0) compares false -> kept
    }
    mask[a * nw + wq] = bits;
  }
  __syncthreads();

  // 4. greedy scan by one warp: lane l owns words l, l+32, ... of the removed set
  if (warp == 0) {
    uint32_t removed = 0;  // word (lane) of the removed bitmap; Kw <= 32 for ncell <= 1024
    int nkeep = 0;
    for (int a = 0; a < K; ++a) {
      const uint32_t wa = __shfl_sync(0xffffffffu, removed, a >> 5);
      if (!((wa >> (a & 31)) & 1u)) {
        if (lane == 0) s_keep[nkeep] = a;
        ++nkeep;
        if (lane < Kw) removed |= mask[a * nw + lane];
      }
    }
    if (lane == 0) s_nkeep = nkeep;
  }
  __syncthreads();

  // 5. emit rows (score, X, Y, x2-X, y2-Y) from the ROUNDED corners (utils.py:165-166)
  const int nkeep = s_nkeep;
  for (int r = threadIdx.x; r < nkeep; r += blockDim.x) {
    const int i = order[s_keep[r]];
    float* o = out_boxes + (static_cast<size_t>(b) * ncell + r) * 5;
    const float x1 = bx[i], y1 = bx[ncell + i];
    o[0] = sc[i];
    o[1] = x1;
    o[2] = y1;
    o[3] = __fsub_rn(bx[2 * ncell + i], x1);
    o[4] = __fsub_rn(bx[3 * ncell + i], y1);
    if (out_cell) out_cell[static_cast<size_t>(b) * ncell + r] = cell[i];
  }
  if (threadIdx.x == 0) out_count[b] = nkeep;
}

// ---------------------------------------------------------------------------- grid encode
// datasets/WIDERFace/dataset.py:32-64.  Last box wins a cell: atomicMax of the box index.
__global__ void __launch_bounds__(kYoloThreads)
grid_encode_kernel(const float* __restrict__ boxes, const int* __restrict__ offsets, int S, double psx, double psy,
                   float width, float height, float* __restrict__ out) {
  extern __shared__ int owner[];  // [S*S]
  const int b = blockIdx.x;
  const int ncell = S * S;
  const int k0 = offsets[b], k1 = offsets[b + 1];
  const float fpsx = static_cast<float>(psx), fpsy = static_cast<float>(psy);
  for (int c = threadIdx.x; c < ncell; c += blockDim.x) owner[c] = -1;
  __syncthreads();
  for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
    const float* bx = boxes + static_cast<size_t>(k) * 5;
    int i = static_cast<int>(floorf(__fdiv_rn(bx[1], fpsx)));   // dataset.py:43
    int j = static_cast<int>(floorf(__fdiv_rn(bx[2], fpsy)));
    i = min(max(i, 0), S - 1);                                   // dataset.py:61-62
    j = min(max(j, 0), S - 1);
    atomicMax(&owner[i * S + j], k - k0);
  }
  __syncthreads();
  float* o = out + static_cast<size_t>(b) * 5 * ncell;
  for (int c = threadIdx.x; c < ncell; c += blockDim.x) {
    const int k = owner[c];
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (k >= 0) {
      const float* bx = boxes + static_cast<size_t>(k0 + k) * 5;
      const int i = static_cast<int>(floorf(__fdiv_rn(bx[1], fpsx)));   // UN-clamped (dataset.py:51-52)
      const int j = static_cast<int>(floorf(__fdiv_rn(bx[2], fpsy)));
      v[0] = bx[0];
      v[1] = __fdiv_rn(__fsub_rn(bx[1], __double2float_rn(static_cast<double>(i) * psx)), fpsx);
      v[2] = __fdiv_rn(__fsub_rn(bx[2], __double2float_rn(static_cast<double>(j) * psy)), fpsy);
      v[3] = __fdiv_rn(bx[3], width);
      v[4] = __fdiv_rn(bx[4], height);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) o[q * ncell + c] = v[q];
  }
}

// ---------------------------------------------------------------------------- step metrics
// models/ModelMeta.py:184-214: IoU matrix of the decoded ground-truth boxes against the decoded predictions
// (torchvision.ops.box_iou on xyxy built from the (x, y, w, h) rows), nan -> 0; per image the number of pairs with
// IoU > thr and the sum of all IoUs.  One CTA per image, no host synchronisation.
__global__ void __launch_bounds__(kYoloThreads)
box_metrics_kernel(const float* __restrict__ gt, const int* __restrict__ gt_count, const float* __restrict__ pred,
                   const int* __restrict__ pred_count, int cap, float thr, float* __restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __shared__ float s_red[kYoloThreads / 32];
  const int b = blockIdx.x;
  const int ng = gt_count[b], np = pred_count[b];
  const float* G = gt + static_cast<size_t>(b) * cap * 5;
  const float* P = pred + static_cast<size_t>(b) * cap * 5;
  float hits = 0.f, sum = 0.f;
  for (int idx = threadIdx.x; idx < ng * np; idx += blockDim.x) {
    const int i = idx / np, j = idx - i * np;
    const float ax1 = G[i * 5 + 1], ay1 = G[i * 5 + 2], ax2 = __fadd_rn(G[i * 5 + 3], ax1), ay2 = __fadd_rn(G[i * 5 + 4], ay1);
    const float bx1 = P[j * 5 + 1], by1 = P[j * 5 + 2], bx2 = __fadd_rn(P[j * 5 + 3], bx1), by2 = __fadd_rn(P[j * 5 + 4], by1);
    const float aa = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    const float ab = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
    const float w = fmaxf(0.f, __fsub_rn(fminf(ax2, bx2), fmaxf(ax1, bx1)));
    const float h = fmaxf(0.f, __fsub_rn(fminf(ay2, by2), fmaxf(ay1, by1)));
    const float inter = __fmul_rn(w, h);
    float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aa, ab), inter));
    if (iou != iou) iou = 0.f;                       // nan_to_num(.., 0)
    hits += iou > thr ? 1.f : 0.f;
    sum += iou;
  }
  const float th = block_sum(hits, s_red);
  const float ts = block_sum(sum, s_red);
  if (threadIdx.x == 0) {
    out[b * 4 + 0] = th;
    out[b * 4 + 1] = ts;
    out[b * 4 + 2] = static_cast<float>(ng);
    out[b * 4 + 3] = static_cast<float>(np);
  }
}

}  // namespace
}  // namespace fd

extern "C" int fd_box_metrics(const float* gt_boxes, const int32_t* gt_count, const float* pred_boxes,
                              const int32_t* pred_count, int B, int cap, float iou_thr, float* out, void* stream) {
  using namespace fd;
  if (!gt_boxes || !gt_count || !pred_boxes || !pred_count || !out || B <= 0 || cap <= 0) return FD_EINVAL;
  launch_k(box_metrics_kernel, dim3(B), dim3(kYoloThreads), 0, static_cast<cudaStream_t>(stream), gt_boxes, gt_count,
           pred_boxes, pred_count, cap, iou_thr, out);
  count_launch();
  return launch_status();
}

using namespace fd;

extern "C" int fd_yolo_loss(const float* pred, const float* gt, int B, int S1, int S2, float* loss,
                            const float* dloss_scale, float* dpred, void* stream) {
  if (!pred || !gt || !loss || B <= 0 || S1 <= 0 || S2 <= 0) return FD_EINVAL;
  launch_k(yolo_loss_kernel, dim3(B), dim3(kYoloThreads), 0, static_cast<cudaStream_t>(stream), pred, gt, S1, S2, loss,
           dloss_scale, dpred);
  count_launch();
  return launch_status();
}

extern "C" int fd_decode_nms(const float* pred, int B, int S1, int S2, float p_thr, double iou_thr, int width,
                             int height, int num_of_patches, float* out_boxes, int32_t* out_cell, int32_t* out_count,
                             void* stream) {
  if (!pred || !out_boxes || !out_count || B <= 0 || S1 <= 0 || S2 <= 0 || num_of_patches <= 0) return FD_EINVAL;
  const int ncell = S1 * S2;
  if (ncell > 1024) return FD_EUNSUPPORTED;
  const int nw = (ncell + 31) / 32;
  const size_t smem = static_cast<size_t>(ncell) * (5 * 4 + 2 * 4) + static_cast<size_t>(ncell) * nw * 4;
  cudaError_t e = set_max_dyn_smem(decode_nms_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // utils.py:108-109: python floats width/num_of_patches, rounded to f32 when they meet the f32 tensor
  const float psx = static_cast<float>(static_cast<double>(width) / num_of_patches);
  const float psy = static_cast<float>(static_cast<double>(height) / num_of_patches);
  launch_k(decode_nms_kernel, dim3(B), dim3(kYoloThreads), smem, static_cast<cudaStream_t>(stream), pred, S1, S2, p_thr,
           iou_thr, psx, psy, static_cast<float>(width), static_cast<float>(height), out_boxes, out_cell, out_count);
  count_launch();
  return launch_status();
}

extern "C" int fd_grid_encode(const float* boxes, const int32_t* box_offsets, int B, int S, int width, int height,
                              float* out, void* stream) {
  if (!box_offsets || !out || B <= 0 || S <= 0) return FD_EINVAL;
  const size_t smem = static_cast<size_t>(S) * S * 4;
  if (smem > 48 * 1024) return FD_EUNSUPPORTED;
  grid_encode_kernel<<<B, kYoloThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      boxes, box_offsets, S, static_cast<double>(width) / S, static_cast<double>(height) / S,
      static_cast<float>(width), static_cast<float>(height), out);
  count_launch();
  return launch_status();
}
