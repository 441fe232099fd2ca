// The SSD side of the hot path (SURVEY.md 8a rows 12, 13 and the grid assignment of dataset_ssd.py):
// multi-scale grid encoding, decode + score threshold + NMS over the 4774 priors, and ssd_loss with
// hard-negative mining (value + gradients).  One CTA per image / per row, everything in shared memory.
// Results that the reference rounds, ranks or indexes (box corners, kept order, mined set, cell
// assignment) are bit-exact: float operations feeding them use the explicit round-to-nearest
// intrinsics so that nvcc cannot contract a*b+c into an FMA.  HBM-bound / latency-bound integer and
// byte work: no tensor cores here.
#include "fd_host.h"

namespace fd {
namespace {

constexpr int kSsdThreads = 512;
constexpr int kMaxScales = 8;

struct SsdScales {
  int n;
  int ps[kMaxScales];      // cells per side
  int base[kMaxScales + 1];  // first prior row of each scale
};

__device__ __forceinline__ int scale_of(const SsdScales& sc, int row) {
  int s = 0;
  while (s + 1 < sc.n && row >= sc.base[s + 1]) ++s;
  return s;
}

// ---------------------------------------------------------------------------- grid encode
// datasets/WIDERFace/dataset_ssd.py:36-76 for every scale, rows concatenated like :134-139.
// Last box wins a cell: atomicMax of the box index per (scale, cell).
__global__ void __launch_bounds__(kSsdThreads)
ssd_grid_encode_kernel(const float* __restrict__ boxes, const int* __restrict__ offsets, SsdScales sc, float width,
                       float height, float* __restrict__ out) {
  extern __shared__ int owner[];   // [P]
  const int b = blockIdx.x;
  const int P = sc.base[sc.n];
  const int k0 = offsets[b], k1 = offsets[b + 1];
  for (int c = threadIdx.x; c < P; c += blockDim.x) owner[c] = -1;
  __syncthreads();
  for (int t = threadIdx.x; t < (k1 - k0) * sc.n; t += blockDim.x) {
    const int k = t / sc.n, s = t - k * sc.n;
    const int ps = sc.ps[s];
    const float pz = static_cast<float>(1.0 / ps);                 // dataset_ssd.py:46-49
    const float* bx = boxes + static_cast<size_t>(k0 + k) * 5;
    const float xn = __fdiv_rn(bx[1], width), yn = __fdiv_rn(bx[2], height);   // :41-43
    int i = static_cast<int>(floorf(__fdiv_rn(xn, pz)));           // :52
    int j = static_cast<int>(floorf(__fdiv_rn(yn, pz)));
    i = min(max(i, 0), ps - 1);                                    // :73-74
    j = min(max(j, 0), ps - 1);
    atomicMax(&owner[sc.base[s] + i * ps + j], k);
  }
  __syncthreads();
  float* o = out + static_cast<size_t>(b) * P * 5;
  for (int c = threadIdx.x; c < P; c += blockDim.x) {
    const int k = owner[c];
    float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (k >= 0) {
      const int s = scale_of(sc, c);
      const int ps = sc.ps[s];
      const double pzd = 1.0 / ps;
      const float pz = static_cast<float>(pzd);
      const float* bx = boxes + static_cast<size_t>(k0 + k) * 5;
      const float xn = __fdiv_rn(bx[1], width), yn = __fdiv_rn(bx[2], height);
      const int i = static_cast<int>(floorf(__fdiv_rn(xn, pz)));    // UN-clamped (:63-64)
      const int j = static_cast<int>(floorf(__fdiv_rn(yn, pz)));
      v[0] = __fsub_rn(bx[0], __double2float_rn(0.001 * ps));       // :59
      v[1] = __fdiv_rn(__fsub_rn(xn, __double2float_rn(static_cast<double>(i) * pzd)), pz);   // :63,67
      v[2] = __fdiv_rn(__fsub_rn(yn, __double2float_rn(static_cast<double>(j) * pzd)), pz);
      v[3] = __fdiv_rn(bx[3], width);
      v[4] = __fdiv_rn(bx[4], height);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) o[static_cast<size_t>(c) * 5 + q] = v[q];
  }
}

// ---------------------------------------------------------------------------- decode + NMS
// datasets/utils.py:56-92 + torchvision nms.  Dynamic smem: float sc[P], bx[4][P]; int order[P]; uint8 sup[P].
__global__ void __launch_bounds__(kSsdThreads)
ssd_decode_nms_kernel(const float* __restrict__ x, SsdScales scl, float p_thr, double iou_thr, float width,
                      float height, int with_priors, float* __restrict__ out_boxes, int* __restrict__ out_count) {
  extern __shared__ uint8_t smraw[];
  const int P = scl.base[scl.n];
  float* sc = reinterpret_cast<float*>(smraw);
  float* bx = sc + P;                            // [4][P]
  int* order = reinterpret_cast<int*>(bx + 4 * P);
  int* keep = order + P;
  uint8_t* sup = reinterpret_cast<uint8_t*>(keep + P);
  __shared__ int s_warp_cnt[kSsdThreads / 32];
  __shared__ int s_base, s_nkeep;

  const int b = blockIdx.x;
  const float* p = x + static_cast<size_t>(b) * P * 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // 1. scale (utils.py:56-67), threshold (:51), corners + rounding (:69-70,82): ordered compaction
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < P; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const float conf = c < P ? p[c * 5] : 0.f;
    const bool pass = c < P && conf > p_thr;
    const uint32_t bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp_cnt[w];
    if (pass) {
      const int k = off + __popc(bal & ((1u << lane) - 1u));
      float X = p[c * 5 + 1], Y = p[c * 5 + 2], Wd = p[c * 5 + 3], Hd = p[c * 5 + 4];
      if (with_priors) {
        const int s = scale_of(scl, c);
        const int ps = scl.ps[s];
        const float inv = static_cast<float>(1.0 / ps);
        const int cell = c - scl.base[s];
        const int i = cell / ps, j = cell - i * ps;
        X = __fadd_rn(__fmul_rn(X, inv), __fmul_rn(inv, static_cast<float>(i)));   // :60-64, priors of :35-48
        Y = __fadd_rn(__fmul_rn(Y, inv), __fmul_rn(inv, static_cast<float>(j)));
      }
      X = __fmul_rn(X, width); Wd = __fmul_rn(Wd, width);
      Y = __fmul_rn(Y, height); Hd = __fmul_rn(Hd, height);
      sc[k] = conf;
      bx[0 * P + k] = rintf(X);
      bx[1 * P + k] = rintf(Y);
      bx[2 * P + k] = rintf(__fadd_rn(Wd, X));
      bx[3 * P + k] = rintf(__fadd_rn(Hd, Y));
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
    sup[i] = 0;
  }
  __syncthreads();
  // 3. greedy NMS in sorted order: the current survivor suppresses everything behind it in parallel
  int nkeep = 0;
  for (int a = 0; a < K; ++a) {
    if (sup[a]) continue;                 // uniform: every thread reads the same flag after the barrier below
    if (threadIdx.x == 0) keep[nkeep] = a;
    ++nkeep;
    const int ia = order[a];
    const float ax1 = bx[ia], ay1 = bx[P + ia], ax2 = bx[2 * P + ia], ay2 = bx[3 * P + ia];
    const float aarea = __fmul_rn(__fsub_rn(ax2, ax1), __fsub_rn(ay2, ay1));
    for (int bpos = a + 1 + threadIdx.x; bpos < K; bpos += blockDim.x) {
      if (sup[bpos]) continue;
      const int ib = order[bpos];
      const float bx1 = bx[ib], by1 = bx[P + ib], bx2 = bx[2 * P + ib], by2 = bx[3 * P + ib];
      const float barea = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
      const float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1);
      const float xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
      const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
      const float inter = __fmul_rn(w, h);
      const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
      if (static_cast<double>(ovr) > iou_thr) sup[bpos] = 1;      // NaN (0/0) compares false -> kept
    }
    __syncthreads();
  }
  __syncthreads();
  // 4. emit rows (score, X, Y, x2-X, y2-Y) from the ROUNDED corners (utils.py:86-87)
  for (int r = threadIdx.x; r < nkeep; r += blockDim.x) {
    const int i = order[keep[r]];
    float* o = out_boxes + (static_cast<size_t>(b) * P + r) * 5;
    const float x1 = bx[i], y1 = bx[P + i];
    o[0] = sc[i];
    o[1] = x1;
    o[2] = y1;
    o[3] = __fsub_rn(bx[2 * P + i], x1);
    o[4] = __fsub_rn(bx[3 * P + i], y1);
  }
  if (threadIdx.x == 0) out_count[b] = nkeep;
}

// ---------------------------------------------------------------------------- ssd_loss
// losses/SSDLoss.py:27-86, one CTA per batch row.  Hard-negative mining keeps the ratio*num_pos negatives
// of largest -log(conf) = SMALLEST confidence (ties: lower prior index first, like the stable CPU sort the
// reference runs); the k-th smallest confidence is found with a 4-pass radix select over the float bits
// in shared memory instead of two full sorts.  Outputs per row: (sum BCE, sum smooth-L1), num_pos and the
// UN-normalised gradients; the caller divides by the batch-wide number of positives (SSDLoss.py:85-86).
__device__ __forceinline__ float block_sum(float v, float* s_red) {
  // deterministic: fixed shuffle tree, then warp partials added in warp order by every thread
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += s_red[w];
  return t;
}

__global__ void __launch_bounds__(kSsdThreads)
ssd_loss_kernel(const float* __restrict__ conf, const float* __restrict__ loc, const float* __restrict__ labels,
                const float* __restrict__ gt_loc, int P, int ratio, float lo, float hi, float* __restrict__ row_sums,
                int* __restrict__ num_pos_out, uint8_t* __restrict__ mask_out, float* __restrict__ dconf,
                float* __restrict__ dloc) {
  extern __shared__ uint32_t key[];          // [P] confidence bits of the negatives, 0xFFFFFFFF for positives
  __shared__ int hist[256];
  __shared__ float s_red[kSsdThreads / 32];
  __shared__ int s_cnt[kSsdThreads];
  __shared__ int s_npos, s_prefix_hi, s_k, s_take;
  const int b = blockIdx.x;
  const float* c = conf + static_cast<size_t>(b) * P;
  const float* lb = labels + static_cast<size_t>(b) * P;

  if (threadIdx.x == 0) s_npos = 0;
  __syncthreads();
  int local_pos = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const bool pos = lb[i] > 0.f;
    local_pos += pos;
    // ascending order of the float bits == ascending confidence for conf >= 0 (sigmoid outputs)
    key[i] = pos ? 0xFFFFFFFFu : __float_as_uint(c[i]);
  }
  atomicAdd(&s_npos, local_pos);
  __syncthreads();
  const int npos = s_npos;
  const int nneg = P - npos;
  const long want = static_cast<long>(npos) * ratio;
  // threshold key T and the number of negatives with key == T that are still taken (lowest indices first)
  uint32_t T = 0xFFFFFFFEu;
  int take_eq = 0x7fffffff;
  if (want < nneg) {
    // radix select of the k-th smallest key (k = want, 1-based) among the negatives
    if (threadIdx.x == 0) { s_prefix_hi = 0; s_k = static_cast<int>(want); }
    uint32_t prefix = 0, pmask = 0;
    for (int pass = 3; pass >= 0; --pass) {
      if (threadIdx.x < 256) hist[threadIdx.x] = 0;
      __syncthreads();
      const int shift = pass * 8;
      for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const uint32_t kk = key[i];
        if (kk != 0xFFFFFFFFu && (kk & pmask) == prefix) atomicAdd(&hist[(kk >> shift) & 255u], 1);
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int k = s_k, d = 0;
        for (; d < 256; ++d) {
          if (k <= hist[d]) break;
          k -= hist[d];
        }
        s_k = k;
        s_prefix_hi = d;
      }
      __syncthreads();
      prefix |= static_cast<uint32_t>(s_prefix_hi) << shift;
      pmask |= 0xFFu << shift;
    }
    T = prefix;
    take_eq = s_k;                  // how many keys equal to T belong to the mined set
    if (want == 0) { T = 0; take_eq = 0; }
  }
  __syncthreads();
  // ordered count of the ties at T: thread t owns the contiguous chunk [t*chunk, (t+1)*chunk)
  const int chunk = (P + blockDim.x - 1) / blockDim.x;
  const int i0 = threadIdx.x * chunk, i1 = min(P, i0 + chunk);
  int ties = 0;
  for (int i = i0; i < i1; ++i) ties += (key[i] == T);
  s_cnt[threadIdx.x] = ties;
  __syncthreads();
  int before = 0;
  for (int t = 0; t < threadIdx.x; ++t) before += s_cnt[t];
  // ---- loss and gradients over this thread's chunk
  float cls = 0.f, l1 = 0.f;
  int seen = before;
  for (int i = i0; i < i1; ++i) {
    const uint32_t kk = key[i];
    const bool pos = kk == 0xFFFFFFFFu;
    bool sel = pos;
    if (!pos) {
      if (want >= nneg) sel = true;
      else if (kk < T) sel = want > 0;
      else if (kk == T) { sel = seen < take_eq; ++seen; }
    }
    const size_t gi = static_cast<size_t>(b) * P + i;
    float dc = 0.f;
    if (sel) {
      const float cv = c[i];
      const float t = rintf(lb[i]);                                   // SSDLoss.py:73
      const float cc = fminf(fmaxf(cv, lo), hi);                      // :14
      cls += -(t * logf(cc) + (1.f - t) * logf(1.f - cc));            // :15-21
      if (cv >= lo && cv <= hi) dc = -(t / cc) + (1.f - t) / (1.f - cc);
    }
    if (dconf) dconf[gi] = dc;
    if (mask_out) mask_out[gi] = sel ? 1 : 0;
    const float4 pl = reinterpret_cast<const float4*>(loc)[gi];
    float4 dl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos) {                                                        // :78-83 smooth L1, beta = 1, summed
      const float4 gl = reinterpret_cast<const float4*>(gt_loc)[gi];
      const float d[4] = {pl.x - gl.x, pl.y - gl.y, pl.z - gl.z, pl.w - gl.w};
      float g[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float ad = fabsf(d[q]);
        l1 += ad < 1.f ? 0.5f * d[q] * d[q] : ad - 0.5f;
        g[q] = ad < 1.f ? d[q] : (d[q] > 0.f ? 1.f : -1.f);
      }
      dl = make_float4(g[0], g[1], g[2], g[3]);
    }
    if (dloc) reinterpret_cast<float4*>(dloc)[gi] = dl;
  }
  const float cls_sum = block_sum(cls, s_red);
  const float l1_sum = block_sum(l1, s_red);
  if (threadIdx.x == 0) {
    row_sums[2 * b] = cls_sum;
    row_sums[2 * b + 1] = l1_sum;
    num_pos_out[b] = npos;
  }
}

int fill_scales(SsdScales* sc, const int* patch_sizes, int n_scales) {
  if (!patch_sizes || n_scales <= 0 || n_scales > kMaxScales) return FD_EINVAL;
  sc->n = n_scales;
  sc->base[0] = 0;
  for (int s = 0; s < n_scales; ++s) {
    if (patch_sizes[s] <= 0) return FD_EINVAL;
    sc->ps[s] = patch_sizes[s];
    sc->base[s + 1] = sc->base[s] + patch_sizes[s] * patch_sizes[s];
  }
  return FD_OK;
}

}  // namespace
}  // namespace fd

extern "C" int fd_ssd_grid_encode(const float* boxes, const int32_t* box_offsets, int B, const int* patch_sizes,
                                  int n_scales, int width, int height, float* out, void* stream) {
  using namespace fd;
  if (!box_offsets || !out || B <= 0) return FD_EINVAL;
  SsdScales sc;
  int rc = fill_scales(&sc, patch_sizes, n_scales);
  if (rc != FD_OK) return rc;
  const size_t smem = static_cast<size_t>(sc.base[sc.n]) * sizeof(int);
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(ssd_grid_encode_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ssd_grid_encode_kernel<<<B, kSsdThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      boxes, box_offsets, sc, static_cast<float>(width), static_cast<float>(height), out);
  count_launch();
  return launch_status();
}

extern "C" int fd_ssd_decode_nms(const float* x, int B, const int* patch_sizes, int n_scales, float p_thr,
                                 double iou_thr, int width, int height, int with_priors, float* out_boxes,
                                 int32_t* out_count, void* stream) {
  using namespace fd;
  if (!x || !out_boxes || !out_count || B <= 0) return FD_EINVAL;
  SsdScales sc;
  int rc = fill_scales(&sc, patch_sizes, n_scales);
  if (rc != FD_OK) return rc;
  const size_t P = sc.base[sc.n];
  const size_t smem = P * (5 * sizeof(float) + 2 * sizeof(int) + 1) + 16;
  if (smem > 227 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(ssd_decode_nms_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  ssd_decode_nms_kernel<<<B, kSsdThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      x, sc, p_thr, iou_thr, static_cast<float>(width), static_cast<float>(height), with_priors, out_boxes, out_count);
  count_launch();
  return launch_status();
}

extern "C" int fd_ssd_loss(const float* conf, const float* loc, const float* labels, const float* gt_loc, int B, int P,
                           int neg_pos_ratio, float* row_sums, int32_t* num_pos, uint8_t* mask, float* dconf,
                           float* dloc, void* stream) {
  using namespace fd;
  if (!conf || !loc || !labels || !gt_loc || !row_sums || !num_pos || B <= 0 || P <= 0 || neg_pos_ratio < 0)
    return FD_EINVAL;
  const size_t smem = static_cast<size_t>(P) * sizeof(uint32_t);
  if (smem > 200 * 1024) return FD_EUNSUPPORTED;
  cudaError_t e = set_max_dyn_smem(ssd_loss_kernel, (int)smem);
  if (e != cudaSuccess) return (int)e;
  // SSDLoss.py:13-14: epsilon = 10**-7 (python float), clamp bounds rounded to f32 by torch
  const float lo = static_cast<float>(1e-7), hi = static_cast<float>(1.0 - 1e-7);
  ssd_loss_kernel<<<B, kSsdThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      conf, loc, labels, gt_loc, P, neg_pos_ratio, lo, hi, row_sums, num_pos, mask, dconf, dloc);
  count_launch();
  return launch_status();
}
