/* fd_b200.h -- C ABI of the B200-native detection hot path (libfd_b200.so).
 *
 * The reference (smpurkis/PyTorch-Face-Detection-from-Scratch) is pure Python and has no FFI;
 * every entry point below replaces a *PyTorch call site* of the reference (cited as
 * file:line under the reference root).  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns int: 0 = OK, > 0 = cudaError_t, < 0 = FD_E* argument error;
 *     nothing throws, nothing allocates device memory, nothing synchronises the device;
 *   - all data pointers are DEVICE pointers owned by the caller (16-byte aligned);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it (CUDA-graph capturable);
 *   - activations are NHWC bf16 (`fd_bf16` = uint16_t bit pattern), accumulators / losses /
 *     gradients of parameters are fp32, indices are int32;
 *   - functions are stateless and re-entrant.
 */
#ifndef FD_B200_H_
#define FD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t fd_bf16;

#if defined(__GNUC__)
#define FD_API __attribute__((visibility("default")))
#else
#define FD_API
#endif

#define FD_OK 0
#define FD_EINVAL (-1)      /* bad argument (shape / null pointer / alignment) */
#define FD_EUNSUPPORTED (-2) /* shape outside what the kernels are instantiated for */
#define FD_EDRIVER (-3)     /* could not resolve cuTensorMapEncodeTiled from the driver */

/* epilogue flags for fd_conv3x3 */
#define FD_EPI_LRELU 1   /* v = v > 0 ? v : slope * v            (PoolResnet.py:36,38) */
#define FD_CONV_1X1 2    /* fd_conv3x3: use only the centre tap of the packed weights = a 1x1 convolution
                          * (models/SeparableCNN.py:13-19,29-35 on 64-channel planes) */
#define FD_CONV_ONE_TAP 4 /* accepted and ignored (reserved): fd_conv3x3 always issues one tap per MMA.  A tap-row fused variant
                          * (N = 192 MMAs, the three taps of a kernel row sharing one A operand) was measured in round 2 and
                          * dropped: it triples the TMEM read traffic and the epilogue becomes the bound. */

FD_API int fd_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
FD_API long long fd_launch_count(void);
FD_API const char* fd_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * 3x3, stride 1, pad 1 convolution, C -> C channels (C = 64), implicit GEMM on tcgen05.
 * Replaces aten::conv2d + leaky_relu + dropout2d + residual add at models/PoolResnet.py:35-40 /
 * models/Resnet.py:29-36 (forward) and, called with dgrad-packed weights, the input-gradient of the
 * same conv.
 *
 *   acc = conv3x3(x, w)                                   fp32 accumulate in TMEM
 *   v   = acc + bias[c]            (bias != NULL)
 *   v   = lrelu(v)                 (flags & FD_EPI_LRELU; requires 0 <= slope <= 1)
 *   v  *= chan_scale[n, c]         (chan_scale != NULL; Dropout2d multiplier)
 *   mask_out = sign bits of v      (mask_out != NULL; LeakyReLU' mask saved for the backward pass)
 *   v  += residual                 (residual != NULL)
 *   out  = bf16(v)                 (out != NULL)
 *   out2 = bf16(v * (mask_in bit ? 1 : slope) * chan_scale2[n, c])   (out2 != NULL; the backward chain;
 *                                   mask_in == NULL means all ones, chan_scale2 == NULL means 1)
 *
 * x, residual, out, out2: [B,H,W,C] bf16.  w_packed: [9][C][C] bf16 from fd_pack_conv3x3 (forward or
 * dgrad packing).  bias: [C] fp32.  chan_scale*: [B,C] fp32.  Masks: uint32 [B,H,W,C/32], bit (c % 32)
 * of word c / 32 is set iff the sign bit of the value is clear (v >= +0).  Any H, W (wide images are tiled in 62-column strips).
 * One of out / out2 leaves through a TMA tensor store (the fast path); if both are given, out2 is
 * written with plain stores.
 */
FD_API int fd_conv3x3(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C,
               const float* bias, float slope, const float* chan_scale,
               const fd_bf16* residual, uint32_t* mask_out, fd_bf16* out,
               const uint32_t* mask_in, const float* chan_scale2, fd_bf16* out2,
               int flags, void* stream);

/* The same convolution with the block's MaxPool2d(2) fused into the epilogue (models/PoolResnet.py:37-42: conv2 ->
 * LeakyReLU -> Dropout2d -> + skip -> pool): the un-pooled sum never reaches HBM.  pooled: [B,H/2,W/2,C] bf16;
 * argmax (nullable): uint16 [B,H/2,W/2,C/8], the window positions of fd_maxpool2x2_fwd (bit-identical to fd_conv3x3
 * followed by fd_maxpool2x2_fwd).  H and W must be even (odd maps: the two separate calls). */
FD_API int fd_conv3x3_pool(const fd_bf16* x, const fd_bf16* w_packed, int B, int H, int W, int C, const float* bias,
                    float slope, const float* chan_scale, const fd_bf16* residual, uint32_t* mask_out,
                    fd_bf16* pooled, uint16_t* argmax, int flags, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The same convolution for the WIDE models (filters = 128: the width train_model.py:17 trains; the 128- / 256-channel
 * blocks of models/SSD.py:164-189) on channel PLANES: every activation tensor is a set of [B,H,W,64] bf16 planes.
 * One launch computes 128 output channels (two planes) from 64*gin input channels (gin <= 4 planes) with
 * tcgen05.mma.cta_group::2 (the two SMs of a TPC on one M=256, N=128 tile, B operand split between their shared
 * memories), partial sums over input planes in TMEM, weights streamed from L2.  Epilogue exactly as fd_conv3x3, per
 * output plane: x[gin], residual[2], mask_out[2], out[2], mask_in[2], chan_scale[2] ([B,64] each), chan_scale2[2],
 * out2[2] are arrays of per-plane device pointers (NULL array = absent); bias is [128].  One of out / out2 (both: see
 * fd_conv3x3_wide_shared_tile).
 * w_packed: bf16 [gin][9][128][64] (tap-major chunks of 128 couts x 64 cins) as written by fd_pack_conv3x3_wide for
 * this launch's group of 128 output channels.  FD_CONV_1X1: centre tap only (pointwise convolution). */
FD_API int fd_conv3x3_wide(const fd_bf16* const* x, int gin, const fd_bf16* w_packed, int B, int H, int W,
                    const float* bias, float slope, const float* const* chan_scale,
                    const fd_bf16* const* residual, uint32_t* const* mask_out, fd_bf16* const* out,
                    const uint32_t* const* mask_in, const float* const* chan_scale2, fd_bf16* const* out2,
                    int flags, void* stream);
/* 1 when fd_conv3x3_wide runs this map in SHARED-TILE mode (fewer two-block tiles than SM pairs: a CTA pair works on ONE
 * tile, outputs through plain stores).  Only in that mode may out AND out2 both be given: out = the (residual-added) sum,
 * out2 = its LeakyReLU'-masked, chan_scale2-scaled copy -- the input gradient of a block and the gradient entering the
 * previous block's conv2 in one launch (models/PoolResnet.py:35-40 backward).  Otherwise both -> FD_EUNSUPPORTED. */
FD_API int fd_conv3x3_wide_shared_tile(int B, int H, int W, int flags);

/* A RUN of wide 3x3 convolutions in ONE launch, for maps small enough that a CTA pair holds a whole image
 * (fd_conv3x3_wide_chain_ok: H * (W + 2) <= 256 and B <= #SM / 2 -- the 15x15 maps of PoolResnet(filters = 128), where
 * models/PoolResnet.py:78-81 stacks 8 residual blocks = 16 convolutions).  Layer l reads slab `in_index` of the two stacked
 * input planes x_stack[g] = [n_stack][B][H][W][64] (bf16) and layer `w_index` of w_packed ([n_w_layers][2][9][128][64], the
 * fd_pack_conv3x3_wide packing); its outputs / skip / masks / dropout multipliers are those of fd_conv3x3_wide (per-plane
 * pointers; `out` and `out2` may both be given).  A layer may read what an earlier layer of the same call wrote (that is the
 * point): layer l + 1 starts when layer l's outputs are complete for the image.  Results are bit-identical to the same
 * sequence of fd_conv3x3_wide calls. */
typedef struct fd_wide_chain_layer {
  int in_index, w_index, flags, reserved;   /* flags: FD_EPI_LRELU */
  const float* bias;                         /* [128] or NULL */
  const fd_bf16* residual[2];
  fd_bf16* out[2];
  fd_bf16* out2[2];
  const float* chan_scale[2];
  const float* chan_scale2[2];
  const uint32_t* mask_in[2];
  uint32_t* mask_out[2];
} fd_wide_chain_layer;
FD_API int fd_conv3x3_wide_chain_ok(int B, int H, int W);
FD_API int fd_conv3x3_wide_chain(const fd_bf16* const* x_stack, int n_stack, const fd_bf16* w_packed, int n_w_layers, int B,
                          int H, int W, float slope, const fd_wide_chain_layer* layers, int n_layers, void* stream);
/* w: [n_layers][Cout][Cin][3][3] fp32 (torch layout) -> w_fwd [n_layers][Cout/128][Cin/64][9][128][64] bf16 (forward)
 * and w_dgrad [n_layers][Cin/128][Cout/64][9][128][64] bf16 (input gradient: taps flipped, channel roles swapped).
 * Either output may be NULL.  Cout (w_fwd) / Cin (w_dgrad) must be multiples of 128, the other a multiple of 64. */
FD_API int fd_pack_conv3x3_wide(const float* w, int n_layers, int Cout, int Cin, fd_bf16* w_fwd, fd_bf16* w_dgrad,
                         void* stream);
/* The same for 1x1 convolutions, w: [n_layers][Cout][Cin][1][1] (models/SSD.py:23-30 skip, models/SeparableCNN.py:13-19):
 * only the centre tap (tap 4) of the packed layouts is written, the others are left as they are -- for FD_CONV_1X1. */
FD_API int fd_pack_conv1x1_wide(const float* w, int n_layers, int Cout, int Cin, fd_bf16* w_fwd, fd_bf16* w_dgrad,
                         void* stream);

/* Weight gradient of the wide convolution: ONE 128 x 128 channel block of dW per call -- input planes x0, x1 (channels
 * 128hh .. 128hh+127) against gradient planes g0, g1 (output channels 128gg .. 128gg+127), each holding `nprob` stacked
 * [B,H,W,64] bf16 tensors (problem q = layer q of a run of equal-shape layers).  tcgen05.mma.cta_group::2: a CTA pair loads
 * one x plane and one g plane each and issues M=256, N=128 instructions (two passes: taps 0..7, then tap 8 -- TMEM holds
 * eight taps of a 128 x 64 block).  The four 64 x 64 sub-blocks are ACCUMULATED into the packed layout of fd_conv3x3_wgrad:
 * dw_packed + sub_off[2*r + c] (+ q * dw_stride) is the [9][64 ci][64 co] fp32 block of (x plane r, g plane c); sub_off is a
 * HOST array of 4 element offsets (multiples of 64).  dbias0 / dbias1 (nullable): [64] fp32 bias gradients of g0 / g1
 * (+ q * dbias_stride), accumulated. */
FD_API int fd_conv3x3_wgrad_wide(const fd_bf16* x0, const fd_bf16* x1, const fd_bf16* g0, const fd_bf16* g1, int nprob, int B,
                          int H, int W, float* dw_packed, const long* sub_off, long dw_stride, float* dbias0,
                          float* dbias1, long dbias_stride, int flags, void* stream);

/* Weight gradient of the same convolution (replaces the wgrad half of autograd's
 * conv2d backward for models/PoolResnet.py:35,37).
 *   dw_packed[t][ci][co] += sum_{n,y,x} g[n,y,x,co] * xpad[n,y+ky-1,x+kx-1,ci],  t = ky*3+kx
 *   dbias[co]            += sum_{n,y,x} g[n,y,x,co]
 * x, g: [B,H,W,C] bf16; dw_packed: [9][C][C] fp32 and dbias: [C] fp32 are ACCUMULATED into
 * (zero them first).  fd_unpack_wgrad3x3 converts to the torch layout [co][ci][3][3]. */
FD_API int fd_conv3x3_wgrad(const fd_bf16* x, const fd_bf16* g, int B, int H, int W, int C,
                     float* dw_packed, float* dbias, int flags, void* stream);

/* The same for `nprob` independent convolutions of identical shape in ONE launch: x and g hold
 * nprob stacked [B,H,W,C] tensors; problem q accumulates into dw_packed + q*dw_stride and
 * dbias + q*dbias_stride (strides in elements).  Used for the run of equal-shape residual blocks,
 * whose weight gradients are mutually independent once the dgrad chain has finished. */
FD_API int fd_conv3x3_wgrad_multi(const fd_bf16* x, const fd_bf16* g, int nprob, int B, int H, int W, int C,
                           float* dw_packed, long dw_stride, float* dbias, long dbias_stride, int flags,
                           void* stream);

/* w: [n_layers][C][C][3][3] fp32 (torch layout, PoolResnet.py:15-28) ->
 *   w_fwd  [n_layers][9][co][ci] bf16,  w_dgrad [n_layers][9][ci][co] bf16 with flipped taps
 * (either output may be NULL). */
FD_API int fd_pack_conv3x3(const float* w, int n_layers, int C, fd_bf16* w_fwd, fd_bf16* w_dgrad, void* stream);
/* dw_packed [n_layers][9][ci][co] fp32 -> dw [n_layers][co][ci][3][3] fp32 (overwrites). */
FD_API int fd_unpack_wgrad3x3(const float* dw_packed, int n_layers, int C, float* dw, void* stream);
/* Channel-plane engines: packed sub-blocks [n_layers][G (g)][G (h)][9][64 ci][64 co] fp32 -> the wide torch tensor
 * dw [n_layers][64 G][64 G][3][3] fp32 (overwrites), in one pass. */
FD_API int fd_unpack_wgrad3x3_planes(const float* dw_packed, int n_layers, int G, float* dw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * A run of `n_blocks` residual blocks of ONE spatial shape without pooling (blocks 2..9 of the
 * reference PoolResnet, models/PoolResnet.py:33-43; the 15x15 blocks of models/Resnet.py:27-40) as one
 * persistent kernel: one CTA per image, activations resident in shared memory across all 2*n_blocks
 * convolutions, weights streamed from L2.  Arithmetic is identical to the equivalent sequence of
 * fd_conv3x3 calls (same MMA order, same bf16 rounding points).
 *
 * LeakyReLU' masks are exchanged as sign bits: mask[B,H,W,C/32] uint32, bit (c % 32) of word c / 32
 * is set iff the sign bit of the activation is clear (v >= +0).
 * fd_resblock_chain_shape_ok returns 1 when (H, W, C) fits the kernel (whole padded image in smem). */
typedef struct fd_chain_fwd_block {
  const float* bias1;       /* [C] conv1 bias                                  (PoolResnet.py:35) */
  const float* bias2;       /* [C] conv2 bias                                  (PoolResnet.py:37) */
  const float* chan_scale;  /* [B,C] Dropout2d multiplier of the block or NULL (PoolResnet.py:39) */
  fd_bf16* a;               /* out [B,H,W,C]: lrelu(conv1), saved for the weight gradient; NULL = not stored */
  uint32_t* mask_a;         /* out: sign bits of a; NULL = not stored */
  uint32_t* mask_b;         /* out: sign bits of b = dropout(lrelu(conv2)) before the skip add; NULL = not stored */
  fd_bf16* out;             /* out [B,H,W,C]: block output b + input; may be NULL except for the last block */
} fd_chain_fwd_block;
/* w_fwd: [2*n_blocks][9][C][C] forward-packed weights of the run (conv1, conv2 of block 0, ...). */
FD_API int fd_resblock_chain_shape_ok(int H, int W, int C);
FD_API int fd_resblock_chain_fwd(const fd_bf16* x, const fd_bf16* w_fwd, const fd_chain_fwd_block* blocks, int n_blocks,
                          int B, int H, int W, int C, float slope, void* stream);

/* Input-gradient chain of the same run (autograd backward of the blocks, without the weight gradients):
 * for block k = n_blocks-1 .. 0, with G = gradient w.r.t. the block output and gp2 = G * drop * lrelu'(b):
 *   gp1      = dgrad_conv2(gp2) * lrelu'(a_k)
 *   G        = dgrad_conv1(gp1) + G                      (skip connection)      -> g_in (if non-NULL)
 *   gp2_prev = G * chan_scale_prev * lrelu'(b_{k-1})     (k > 0)
 * blocks[] is in FORWARD order.  g_out, gp2_last: [B,H,W,C] gradient w.r.t. the last block's output and
 * its gp2 (both produced by fd_head_bwd / fd_maxpool2x2_bwd). */
typedef struct fd_chain_bwd_block {
  const uint32_t* mask_a;        /* sign bits of this block's a */
  fd_bf16* gp1;                  /* out [B,H,W,C] or NULL */
  fd_bf16* g_in;                 /* out [B,H,W,C]: gradient w.r.t. the block input, or NULL */
  const uint32_t* mask_b_prev;   /* sign bits of the previous block's b (NULL for block 0) */
  const float* chan_scale_prev;  /* [B,C] Dropout2d multiplier of the previous block or NULL */
  fd_bf16* gp2_prev;             /* out [B,H,W,C]: gp2 of the previous block (NULL for block 0) */
} fd_chain_bwd_block;
/* w_dgrad: [2*n_blocks][9][C][C] dgrad-packed weights of the run, in FORWARD layer order. */
FD_API int fd_resblock_chain_bwd(const fd_bf16* g_out, const fd_bf16* gp2_last, const fd_bf16* w_dgrad,
                          const fd_chain_bwd_block* blocks, int n_blocks, int B, int H, int W, int C, float slope,
                          void* stream);

/* Adam step on flat fp32 buffers of n elements (n % 4 == 0), torch.optim.Adam semantics without amsgrad
 * (models/ModelMeta.py:104-112: the reference's SAMSGD is numerically plain Adam).  One launch for all parameters
 * (the reference's _multi_tensor Adam issues ~10 foreach kernels over 44 tensors).
 * state == NULL: `step` is the 1-based step count used for the bias corrections and `lr` the learning rate.
 * state != NULL (device, int32[4], zero-initialised): the step count (state[0]) and the learning rate (state[2], fp32
 * bits) live on the device and the kernel advances the count itself, so the launch can be replayed from a CUDA
 * graph; `step` and `lr` are ignored. */
FD_API int fd_adam_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, int32_t* state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory (csrc/comm.cu).  The reference is single-GPU
 * (train_model.py:47-53); its loss is a SUM over the batch (models/ModelMeta.py:173-176,215), so the exchanged
 * quantity is the plain sum of the per-rank flat gradient buffers.  One process per GPU; every rank allocates a
 * window (fd_comm_alloc, a cudaMalloc allocation, zero-filled), exports it (64-byte CUDA IPC handle), exchanges the
 * handles through its host-side plumbing (torch.distributed all_gather), imports the peers' windows, synchronises
 * once on the host, and from then on calls fd_allreduce_sum_f32 -- ONE kernel per call, no host involvement,
 * CUDA-graph capturable, bit-identical results on every rank.  These are the only entry points that allocate. */
FD_API long fd_comm_window_bytes(long n, int world);              /* window size for n fp32 elements (n % 4 == 0), world <= 8 */
FD_API int fd_comm_alloc(long bytes, void** ptr);
FD_API int fd_comm_free(void* ptr);
FD_API int fd_comm_export(void* ptr, unsigned char* handle64);
FD_API int fd_comm_import(const unsigned char* handle64, void** peer_ptr);
FD_API int fd_comm_release(void* peer_ptr);
FD_API int fd_comm_error_offset(void);                            /* byte offset of the u32 status word in a window (0 = ok) */
/* status of the LOCAL window: 0 = ok, 1 + r = a barrier wait on rank r timed out (20 s) and that call's result is invalid.
 * Synchronises the device: a host-side health check, not part of the data path. */
FD_API int fd_comm_status(void* window, int* status);
/* windows: HOST array of `world` device pointers (entry `rank` = the local window, the rest imported);
 * data: the local flat fp32 buffer of n elements, summed over ranks in place. */
FD_API int fd_allreduce_sum_f32(void* const* windows, int rank, int world, float* data, long n, void* stream);
/* the same on `blocks` (1..64) thread blocks -- an exchange overlapped with compute kernels uses few (every rank the same) */
FD_API int fd_allreduce_sum_f32_blocks(void* const* windows, int rank, int world, float* data, long n, int blocks, void* stream);

/* Elementwise halves of a residual block for backbones wider than the 64-channel tensor-core kernels (filters = 128
 * = two 64-channel planes per tensor, engine.PlanarEngine).  A 128 -> 128 convolution is evaluated as the sum of two
 * 64 -> 64 convolutions per output plane (the second fd_conv3x3 call takes the first one's raw output as `residual`,
 * without FD_EPI_LRELU), so the activation of models/PoolResnet.py:35-40 runs here on 64-channel planes [B,HW,64]:
 *   fd_act_mask : out = lrelu(x) * chan_scale[n,c] + residual ; mask_out bit = sign bit of x clear (nullable operands)
 *   fd_grad_mask: out = g * (mask bit ? 1 : slope) * chan_scale[n,c]   (mask_bits NULL = all ones) */
/* Depthwise 3x3 (pad 1, no bias) + LeakyReLU on one 64-channel plane [B,H,W,64] (models/SeparableCNN.py:20-27,45-46);
 * w_dw: [9 taps][64] fp32 as written by fd_sep_pack.  Used by the separable backbone on channel planes (filters = 128);
 * the 64-channel model runs the fused fd_sepblock_fwd. */
FD_API int fd_dwconv3x3_lrelu(const fd_bf16* x, const float* w_dw, int B, int H, int W, int C, float slope, fd_bf16* out,
                       void* stream);
FD_API int fd_act_mask(const fd_bf16* x, int B, int HW, int C, float slope, const float* chan_scale, const fd_bf16* residual,
                uint32_t* mask_out, fd_bf16* out, void* stream);
FD_API int fd_grad_mask(const fd_bf16* g, int B, int HW, int C, float slope, const uint32_t* mask_bits, const float* chan_scale,
                 fd_bf16* out, void* stream);

/* Dropout2d multipliers (models/PoolResnet.py:39,100; nn.Dropout2d zeroes whole channels and rescales):
 * out[i] = r[i] < keep ? 1/keep : 0 with keep = keep_block for i < n_block and keep_head otherwise.
 * r: n uniform randoms in [0,1) (one per (layer, image, channel)), produced by the caller's generator. */
FD_API int fd_dropout_scale(const float* r, long n, long n_block, float keep_block, float keep_head, float* out,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Depthwise-separable residual block of SeparableCNN, forward (models/SeparableCNN.py:10-51), eval mode:
 *   out = pw2( lrelu( dw3x3( lrelu( pw1(x) ) ) ) ) + x        (all three convolutions without bias, :10,19,28,35)
 * as ONE kernel (csrc/sepblock.cu): the two 1x1 convolutions on the tensor cores, the depthwise 3x3 (pad 1) on the
 * CUDA cores out of shared memory; the intermediates never touch HBM.  x, out: [B,H,W,64] bf16 NHWC.
 *   w_pw1, w_pw2: [64 cout][64 cin] bf16;  w_dw: [9 taps][64] fp32 -- all three as written by fd_sep_pack.
 * pool != 0 fuses the MaxPool2d(2) that follows while H > num_of_patches (:49-50): out is then [B,H/2,W/2,64] and
 * the un-pooled sum never reaches HBM (bit-identical to fd_sepblock_fwd + fd_maxpool2x2_fwd). */
FD_API int fd_sepblock_fwd(const fd_bf16* x, const fd_bf16* w_pw1, const float* w_dw, const fd_bf16* w_pw2, int B, int H,
                    int W, int C, float slope, int pool, fd_bf16* out, void* stream);
/* pw: n_pw fp32 pointwise weights (any number of [64][64] matrices, nn.Conv2d layout [cout][cin][1][1]) -> bf16, same
 * order; dw: n_dw_layers depthwise weights [64][1][3][3] fp32 -> [layer][9][64] fp32. */
FD_API int fd_sep_pack(const float* pw, long n_pw, fd_bf16* pw_out, const float* dw, int n_dw_layers, float* dw_out,
                void* stream);

/* ---------------------------------------------------------------------------------------------
 * Stem convolution (models/PoolResnet.py:70-76,98): KxK stride s pad p, Cin(3) -> C, input fp32 NCHW
 * (or uint8 NCHW with the /255 of PoolResnet.py:95 fused: x_is_u8 = 1), output NHWC bf16, bias added.
 * w: [C][Cin][K][K] fp32.
 * x_cache (nullable): a ZERO-INITIALISED buffer of fd_stem_cache_elems(...) bf16 elements.  The forward fills it
 * with the bf16-converted images in the operand layout of its tensor-core tile; passing the same buffer to
 * fd_stem_wgrad makes the weight gradient read 2 B/pixel through TMA instead of re-reading and re-converting the
 * fp32 images.  fd_stem_cache_elems returns 0 for shapes that do not use the cache (generic stem kernel). */
FD_API long fd_stem_cache_elems(int B, int Cin, int Hin, int Win, int C, int K, int stride, int pad);
FD_API int fd_stem_fwd(const void* x, int x_is_u8, const float* w, const float* bias, int B, int Cin, int Hin, int Win,
                int C, int K, int stride, int pad, fd_bf16* y, fd_bf16* x_cache, void* stream);
/* dw[C][Cin][K][K] += x (*) g ; dbias[C] += sum g.  g: [B,Ho,Wo,C] bf16. */
FD_API int fd_stem_wgrad(const void* x, int x_is_u8, const fd_bf16* g, int B, int Cin, int Hin, int Win, int C, int K,
                  int stride, int pad, float* dw, float* dbias, const fd_bf16* x_cache, void* stream);
/* Models wider than 64 filters (train_model.py:17 trains filters = 128) run the stem once per 64-channel output plane.
 * The first plane's fd_stem_fwd (with x_cache) reads the images; the other planes read the bf16 copy it left behind:
 *   fd_stem_fwd_cached      y[B,Ho,Wo,64] = conv(x_cache, w[64][Cin][K][K]) + bias, same arithmetic as fd_stem_fwd;
 *   fd_stem_wgrad_pair      the weight gradients of TWO planes from one pass over the copy:
 *                           dw[128][Cin][K][K] += x (*) {g0, g1}, dbias[128] += sum {g0, g1}.
 * Both return FD_EUNSUPPORTED for shapes fd_stem_cache_elems reports 0 for. */
FD_API int fd_stem_fwd_cached(const fd_bf16* x_cache, const float* w, const float* bias, int B, int Cin, int Hin, int Win,
                       int K, int stride, int pad, fd_bf16* y, void* stream);
FD_API int fd_stem_wgrad_pair(const fd_bf16* x_cache, const fd_bf16* g0, const fd_bf16* g1, int B, int Cin, int Hin, int Win,
                       int K, int stride, int pad, float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Head (models/PoolResnet.py:83-89,100-102): Dropout2d multiplier, KxK stride-1 conv C -> 5 with
 * padding `pad`, + bias, sigmoid.  x: [B,H,W,C] bf16, w: [5][C][K][K] fp32, y: [B,5,Ho,Wo] fp32. */
/* w_t (nullable): the same weights pre-arranged by fd_head_pack ([K*K*5][C] fp32, tap-major, swizzled channel
 * groups); with it the C = 64 fast-path kernels run, without it a generic kernel transposes w itself. */
/* bias == NULL (64-channel fast path only): y receives the PARTIAL LOGITS of this 64-channel plane -- no bias, no
 * sigmoid -- for heads wider than 64 channels evaluated plane by plane (engine_planar.PlanarEngine). */
FD_API int fd_head_pack(const float* w, int C, int K, float* w_t, void* stream);
FD_API int fd_head_fwd(const fd_bf16* x, const float* chan_scale, const float* w, const float* w_t, const float* bias,
                int B, int H, int W, int C, int K, int pad, float* y, void* stream);
/* dy: [B,5,Ho,Wo] fp32 gradient w.r.t. the sigmoid OUTPUT y.  Produces
 *   dx [B,H,W,C] bf16 (gradient w.r.t. the block output, dropout multiplier applied; overwritten) and,
 *   when dx2 != NULL, dx2 = dx * chan_scale2 * (mask_bits bit ? 1 : slope)   (mask_bits: uint32 [B,H,W,C/32]),
 *   dw [5][C][K][K] fp32 (+=), dbias [5] fp32 (+=). */
FD_API int fd_head_bwd(const fd_bf16* x, const float* chan_scale, const float* w, const float* w_t, const float* y,
                const float* dy, int B, int H, int W, int C, int K, int pad, fd_bf16* dx, const uint32_t* mask_bits,
                const float* chan_scale2, float slope, fd_bf16* dx2, float* dw, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MaxPool2d(2) (models/PoolResnet.py:41-42; floor mode for odd H / W, models/SSD.py:80 pools 15 -> 7).
 * x: [B,H,W,C] bf16 -> y: [B,H/2,W/2,C] bf16.
 * argmax (nullable, training): uint16 [B,H/2,W/2,C/8], 2 bits per channel = position dy*2+dx of the FIRST maximum of
 * the window (torch's tie rule); with it the backward does not re-read x. */
FD_API int fd_maxpool2x2_fwd(const fd_bf16* x, int B, int H, int W, int C, fd_bf16* y, uint16_t* argmax, void* stream);
/* Backward of the pool fused with the start of the block's backward chain:
 *   gs  = unpool(gy) routed to the FIRST maximum of each 2x2 window of x (torch tie rule),
 *   gs2 = gs * chan_scale[n,c] * (mask_bits bit ? 1 : slope)     (gs2 != NULL)
 * x, gs, gs2: [B,H,W,C] bf16; gy: [B,H/2,W/2,C] bf16; mask_bits: uint32 [B,H,W,C/32] sign bits.
 * argmax (nullable): the positions written by fd_maxpool2x2_fwd; when given, x is not read (and may be NULL). */
FD_API int fd_maxpool2x2_bwd(const fd_bf16* x, const fd_bf16* gy, int B, int H, int W, int C, fd_bf16* gs,
                      const uint32_t* mask_bits, const float* chan_scale, float slope, fd_bf16* gs2,
                      const uint16_t* argmax, void* stream);

/* ---------------------------------------------------------------------------------------------
 * losses/YoloLoss.py:4-44 for a batch: loss[b] = yolo_loss(pred[b], gt[b]) and, when dpred != NULL,
 * dpred[b] = dloss_scale[b] * d loss[b] / d pred[b]  (dloss_scale == NULL means 1).
 * pred, gt, dpred: [B,5,S1,S2] fp32.  The no-object weight is 1/S1 (YoloLoss.py:6,25). */
FD_API int fd_yolo_loss(const float* pred, const float* gt, int B, int S1, int S2, float* loss, const float* dloss_scale,
                 float* dpred, void* stream);

/* ---------------------------------------------------------------------------------------------
 * datasets/utils.py:95-170 ReduceBoundingBoxes.forward for a batch (decode, score threshold,
 * round, torchvision-compatible NMS).  pred: [B,5,S1,S2] fp32.
 *   out_boxes [B, S1*S2, 5] fp32: rows (score, x, y, w, h) in keep order (descending score)
 *   out_cell  [B, S1*S2] int32 (nullable): flat cell index i*S2+j of each kept row
 *   out_count [B] int32: number of kept rows
 * patch sizes are width/num_of_patches and height/num_of_patches (utils.py:108-109), NOT derived
 * from S1/S2.  iou_thr is compared in double like torchvision's CPU kernel. */
FD_API int fd_decode_nms(const float* pred, int B, int S1, int S2, float p_thr, double iou_thr, int width, int height,
                  int num_of_patches, float* out_boxes, int32_t* out_cell, int32_t* out_count, void* stream);

/* models/ModelMeta.py:184-214 step metrics for a batch, without host synchronisation: for every image the IoU
 * matrix (torchvision.ops.box_iou, nan -> 0) of the decoded ground-truth rows against the decoded prediction rows.
 * gt_boxes / pred_boxes: [B, cap, 5] fp32 rows (score, x, y, w, h) as written by fd_decode_nms, with row counts
 * gt_count / pred_count [B] int32.  out: [B,4] fp32 = (#pairs with IoU > iou_thr, sum of IoUs, n_gt, n_pred). */
FD_API int fd_box_metrics(const float* gt_boxes, const int32_t* gt_count, const float* pred_boxes, const int32_t* pred_count,
                   int B, int cap, float iou_thr, float* out, void* stream);

/* datasets/WIDERFace/dataset.py:32-64 convert_bbx_to_feature_map for a ragged batch.
 * boxes: [total,5] fp32 rows (1,x,y,w,h); box_offsets: [B+1] int32; out: [B,5,S,S] fp32 (overwritten).
 * Later boxes overwrite earlier ones that fall in the same cell. */
FD_API int fd_grid_encode(const float* boxes, const int32_t* box_offsets, int B, int S, int width, int height, float* out,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * SSD head path.  P = sum(ps*ps) priors over `n_scales` grids of ps x ps cells
 * (patch_sizes is a HOST array; the reference uses (60, 30, 15, 7) -> 4774, datasets/utils.py:15).
 *
 * fd_ssd_grid_encode: datasets/WIDERFace/dataset_ssd.py:36-76,134-139 for a ragged batch.
 *   boxes [total,5] fp32 rows (1,x,y,w,h) in pixels; box_offsets [B+1] int32; out [B,P,5] fp32 (overwritten),
 *   rows = scale-major, then cell (i,j) row-major; later boxes overwrite earlier ones of the same cell. */
FD_API int fd_ssd_grid_encode(const float* boxes, const int32_t* box_offsets, int B, const int* patch_sizes, int n_scales,
                       int width, int height, float* out, void* stream);

/* fd_ssd_decode_nms: datasets/utils.py:56-92 ReduceSSDBoundingBoxes.forward for a batch.
 *   x [B,P,5] fp32 rows (score,x,y,w,h); with_priors applies the prior scaling/offsets of utils.py:59-64.
 *   out_boxes [B,P,5] rows (score,x,y,w,h) in keep order (descending score); out_count [B] kept rows. */
FD_API int fd_ssd_decode_nms(const float* x, int B, const int* patch_sizes, int n_scales, float p_thr, double iou_thr,
                      int width, int height, int with_priors, float* out_boxes, int32_t* out_count, void* stream);

/* fd_ssd_loss: losses/SSDLoss.py:27-86 (hard-negative mining + clamped BCE + smooth L1), per batch row.
 *   conf, labels [B,P]; loc, gt_loc [B,P,4] fp32.
 *   row_sums [B,2]: (sum of BCE over the mined set, sum of smooth-L1 over the positive priors) of each row;
 *   num_pos [B] int32: positive priors of each row.  loss = (sum row_sums) / (sum num_pos)  (SSDLoss.py:85-86).
 *   mask [B,P] uint8 (nullable): the mined set.  dconf [B,P], dloc [B,P,4] (nullable): gradients of the
 *   UN-normalised sums; the caller scales them by 1 / sum(num_pos) -- across data-parallel ranks that sum is
 *   one extra integer all-reduce. */
FD_API int fd_ssd_loss(const float* conf, const float* loc, const float* labels, const float* gt_loc, int B, int P,
                int neg_pos_ratio, float* row_sums, int32_t* num_pos, uint8_t* mask, float* dconf, float* dloc,
                void* stream);


/* ---------------------------------------------------------------------------------------------
 * MobilenetV3 backbone, inference (models/MobilenetV3Backbone.py:33-60; the timm tf_mobilenetv3_small_100 graph as
 * stored in the official TorchScript archive).  BatchNorm (eval, eps 1e-3) is folded into the preceding convolution by
 * the caller: every entry point takes a per-output-channel `scale` at pack time and a fused bias.  Activations are
 * NHWC bf16 with the network's own channel counts (multiples of 8, not padded).  act: 0 none, 1 ReLU, 2 Hardswish.
 *
 * fd_pw_conv: 1x1 convolution as a tcgen05 GEMM over pixels (conv_pw / conv_pwl / ConvBnAct / SE-free 1x1 layers):
 *   out[m, n] = act(sum_k x[m, k] * w[n, k] + bias[n]) (+ residual[m, n]);  x [M, K], out / residual [M, N] bf16.
 *   w_packed: fd_pw_packed_elems(N, K) bf16 written by fd_pw_pack from w [N][K] fp32 (* scale[n], nullable);
 *   bias_padded: fd_pw_padded_n(N) fp32 (entries >= N zero). */
FD_API long fd_pw_packed_elems(int N, int K);
FD_API int fd_pw_padded_n(int N);
FD_API int fd_pw_pack(const float* w, const float* scale, int N, int K, fd_bf16* out, void* stream);
FD_API int fd_pw_conv(const fd_bf16* x, const fd_bf16* w_packed, const float* bias_padded, long M, int K, int N, int act,
               const fd_bf16* residual, fd_bf16* out, void* stream);
/* Conv2dSame 3x3 stride 2, 3 -> 16, + bias + Hardswish (feature_extractor.0-2 of the archive).  x: fp32 (or uint8 with
 * the /255 of MobilenetV3Backbone.py:52 fused) NCHW [B,3,H,W]; w [16][3][3][3] fp32 (BatchNorm folded); out NHWC bf16
 * [B,Ho,Wo,16].  pad_t / pad_l: the TF "SAME" leading padding (0 for even sizes). */
FD_API int fd_mbv3_stem(const void* x, int x_is_u8, const float* w, const float* bias, int B, int H, int W, int pad_t, int pad_l,
                 int Ho, int Wo, fd_bf16* out, void* stream);
/* Depthwise KxK (K = 3 | 5) stride 1 | 2 convolution + bias + act.  w_packed [K*K][C] fp32 from fd_dw_pack (w [C][1][K][K]
 * * scale[c]).  se_partial (nullable): fp32 [B][fd_dwconv_se_blocks(Ho,Wo,C)][C], receives per-thread-block channel sums
 * of the bf16 output (SqueezeExcite's mean; plain stores in a fixed order -- inference is bit-reproducible). */
FD_API int fd_dw_pack(const float* w, const float* scale, int C, int K, float* out, void* stream);
FD_API int fd_dwconv_se_blocks(int Ho, int Wo, int C);
FD_API int fd_dwconv(const fd_bf16* x, const float* w_packed, const float* bias, int B, int H, int W, int C, int K, int stride,
              int pad_t, int pad_l, int Ho, int Wo, int act, fd_bf16* out, float* se_partial, void* stream);
/* SqueezeExcite gate (efficientnet_blocks.py SqueezeExcite): gate[n,c] = hardsigmoid(W2 relu(W1 mean + b1) + b2) with
 * mean = (sum of the nblk block partials written by fd_dwconv) / HW.  w1 [R][C], w2 [C][R] fp32. */
FD_API int fd_se_gate(const float* se_partial, int nblk, int B, int HW, const float* w1, const float* b1, const float* w2,
               const float* b2, int C, int R, float* gate, void* stream);
/* x[n,p,c] *= gate[n,c] in place. */
FD_API int fd_scale_channels(fd_bf16* x, const float* gate, int B, int HW, int C, void* stream);
/* 3x3 pad-1 convolution C -> 5 + bias + sigmoid (MobilenetV3Backbone.py:40-46,57-58).  y [B,5,H,W] fp32. */
FD_API int fd_head3x3_fwd(const fd_bf16* x, const float* w, const float* bias, int B, int H, int W, int C, float* y, void* stream);

/* transforms.Resize of the `predict == 1` branch (models/PoolResnet.py:94-95, models/BaseModel.py:64): bilinear,
 * align_corners = False, no antialias.  x, out: `planes` images of h x w / H x W, uint8 (rounded half to even) or fp32. */
FD_API int fd_resize_bilinear(const void* x, int is_u8, long planes, int h, int w, int H, int W, void* out, void* stream);
/* scatter != 0: dst[idx[i]] = src[i];  else dst[i] = src[idx[i]]   (i < n).  Parameters of models narrower than the
 * 64-channel kernel planes are scattered into zero-padded planes and their gradients gathered back. */
FD_API int fd_index_copy_f32(float* dst, const float* src, const int32_t* idx, long n, int scatter, void* stream);


/* ---------------------------------------------------------------------------------------------
 * Prediction heads of the SSD model (models/SSD.py:174-177,238-254 + apply_priors :206-220): Linear(C -> 5) on every
 * pixel of one scale's feature map, sigmoid on column 0, columns 1-2 times `mult` (1 / ps), columns 1-4 plus `priors`,
 * written into rows [prior_off, prior_off + HW) of out [B,P,5] fp32.  The map is given as G 64-channel NHWC bf16 planes
 * (HOST array of G device pointers; C <= 64 * G real channels); w [5][C], bias [5] fp32; mult [P], priors [P,4] fp32.
 * Backward: dout [B,P,5] = gradient w.r.t. out; dx planes are overwritten, dw [5][C] and db [5] accumulated into. */
FD_API int fd_ssd_head_fwd(const fd_bf16* const* x_planes, int G, const float* w, const float* bias, int B, int HW, int C,
                    const float* mult, const float* priors, int prior_off, int P, float* out, void* stream);
FD_API int fd_ssd_head_bwd(const fd_bf16* const* x_planes, fd_bf16* const* dx_planes, int G, const float* w, int B, int HW,
                    int C, const float* mult, int prior_off, int P, const float* out, const float* dout, float* dw,
                    float* db, void* stream);


/* Backward pieces of the depthwise-separable block (models/SeparableCNN.py:40-51) on 64-channel planes:
 *   fd_lrelu_bwd: out = g * (ref >= +0 ? 1 : slope), the LeakyReLU' factor taken from the sign of the saved activation
 *                 (n bf16 elements, n % 8 == 0);
 *   fd_dwconv3x3_wgrad: dw[c][ky][kx] += sum g[n,y,x,c] * x[n,y+ky-1,x+kx-1,c]  (dw: [64][1][3][3] fp32, accumulated).
 * The depthwise input gradient is fd_dwconv with the taps flipped; the pointwise gradients are fd_conv3x3 (FD_CONV_1X1,
 * dgrad packing) and the centre tap of fd_conv3x3_wgrad. */
FD_API int fd_lrelu_bwd(const fd_bf16* g, const fd_bf16* ref, long n, float slope, fd_bf16* out, void* stream);
FD_API int fd_dwconv3x3_wgrad(const fd_bf16* x, const fd_bf16* g, int B, int H, int W, int C, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FD_B200_H_ */
