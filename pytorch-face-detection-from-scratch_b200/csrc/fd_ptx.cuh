// Thin inline-PTX wrappers for the sm_100a features the convolution kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the
// shared-memory matrix descriptors.  Device-only; include from .cu files.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a converged warp.  Branching on elect.sync (instead of `lane == 0`) tells ptxas that the
// region is executed by exactly one thread, so descriptors stay in uniform registers and every
// tcgen05.mma / TMA issue is a single UTCHMMA / UTMALDG without a divergence ("waterfall") loop.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- programmatic dependent launch
// pdl_trigger: the next kernel of the stream may be scheduled (its CTAs start as SMs free up);
// pdl_wait   : returns once the previous kernel has completed and its memory is visible.  Both are no-ops for a
// kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure (trap), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}

// Same, with a suspend-time hint: the waiting warp sleeps in hardware (up to `ns`) instead of re-issuing
// try_wait every few cycles and stealing issue slots from the warps that have work.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t ns = 4000) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 22)) __trap();
  }
}

// ----------------------------------------------------------------------------- thread-block clusters / DSMEM
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_v4(uint32_t cluster_addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
// Asynchronous 16-byte store into the shared memory of a cluster peer; completion is reported to an mbarrier of
// the DESTINATION CTA as 16 transaction bytes (like TMA): no release fence on the writer, and the data is
// visible to whoever observes that barrier's phase, including the async proxy (tcgen05.mma operands).
__device__ __forceinline__ void st_async_cluster_v4(uint32_t cluster_addr, const uint4& v, uint32_t cluster_bar_addr) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   cluster_addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_bar_addr)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
// wait with cluster-scope acquire: pairs with mbar_arrive_remote of a thread in the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_cluster() {
  asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (tcgen05), 128-byte swizzle.
//   bits [0,14)  start address >> 4        bits [16,30) leading-dim byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) version = 1 (Blackwell)
//   bits [49,52) base offset               bits [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand  : rows of 128 B (64 bf16 of K), 8-row groups SBO bytes apart; LBO unused.
// MN-major operand : rows of 128 B (64 bf16 of M/N), one row per K index; 8-K-row groups SBO bytes
//                    apart; successive 64-element M/N atoms LBO bytes apart.
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                     uint32_t base_offset = 0) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(base_offset & 7u) << 49;
  d |= 2ull << 61;
  return d;
}
// Incremental form for the issue loops: constant high word (SBO = 1024 B, version 1, 128B swizzle) and
// a low word = (address >> 4) | (LBO >> 4) << 16 that is advanced with plain 32-bit adds.
constexpr uint64_t kSdescHiSw128 = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
__device__ __forceinline__ uint32_t sdesc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t sdesc_sw128(uint32_t lo) { return kSdescHiSw128 | static_cast<uint64_t>(lo); }

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// The 36 MMAs (9 taps x 4 K-steps, M=128, N=64, K=16) of one 128-row block of a 3x3 convolution in the
// halo-tile formulation: tap (ky,kx) reads the A tile shifted by (ky*Wp + kx) rows (128 B = 8 descriptor
// units each), B = tap t of the packed weights (8 KB = 512 units apart).  Loops over the taps stay ROLLED
// and the descriptors advance incrementally: ptxas otherwise precomputes all 72 descriptors, runs out of
// uniform registers and spills (MOV.SPILL / R2UR), and the single issuing thread - not the tensor pipe -
// becomes the bottleneck.  `after_tap(t)` runs after the MMAs of tap t (commit hooks).
template <typename F>
__device__ __forceinline__ void issue_conv3x3_block(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t wp_units,
                                                    uint32_t idesc, F&& after_tap) {
  uint32_t acc = 0;
  int t = 0;
#pragma unroll 1
  for (int ky = 0; ky < 3; ++ky) {
#pragma unroll 1
    for (int kx = 0; kx < 3; ++kx, ++t) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_bf16(d_tmem, sdesc_sw128(a_lo + 2 * k), sdesc_sw128(b_lo + 2 * k), idesc, acc);
        acc = 1;
      }
      after_tap(t);
      a_lo += 8;
      b_lo += 512;
    }
    a_lo += wp_units - 24;      // next tap row: + Wp rows, minus the 3 columns already walked
  }
}

// ----------------------------------------------------------------------------- packed fp32x2 math (FADD2 / FMUL2)
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ----------------------------------------------------------------------------- epilogue building blocks
// One thread = one output pixel x 16 channels (8 packed pairs).  Shared by conv3x3_tc.cu and
// resblock_chain.cu so that both paths round identically.
//   v = acc + bias ; v = max(v, slope*v) (LeakyReLU, 0 <= slope <= 1) ; v *= chan_scale
__device__ __forceinline__ void epi_bias_act16(const uint32_t (&acc)[16], const float* s_bias, const float* s_cs,
                                               bool lrelu, bool has_cs, uint64_t slope2, uint64_t (&v2)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + 4 * i);
    v2[2 * i] = add2(pk2u(acc[4 * i], acc[4 * i + 1]), pk2(b4.x, b4.y));
    v2[2 * i + 1] = add2(pk2u(acc[4 * i + 2], acc[4 * i + 3]), pk2(b4.z, b4.w));
  }
  if (lrelu) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float a0, a1, t0, t1;
      upk2(v2[i], a0, a1);
      upk2(mul2(v2[i], slope2), t0, t1);
      v2[i] = pk2(fmaxf(a0, t0), fmaxf(a1, t1));
    }
  }
  if (has_cs) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 s4 = *reinterpret_cast<const float4*>(s_cs + 4 * i);
      v2[2 * i] = mul2(v2[2 * i], pk2(s4.x, s4.y));
      v2[2 * i + 1] = mul2(v2[2 * i + 1], pk2(s4.z, s4.w));
    }
  }
}
// 16 sign bits, bit j set iff the sign bit of channel j is clear (v >= +0; -0 and negatives -> 0).
// One funnel shift per element: m = (m << 1) | (bits(v) >> 31), walked from channel 15 down to 0.
__device__ __forceinline__ uint32_t epi_sign_bits16(const uint64_t (&v2)[8]) {
  uint32_t m = 0;
#pragma unroll
  for (int i = 7; i >= 0; --i) {
    const uint32_t lo = static_cast<uint32_t>(v2[i]), hi = static_cast<uint32_t>(v2[i] >> 32);
    m = __funnelshift_l(hi, m, 1);
    m = __funnelshift_l(lo, m, 1);
  }
  return (~m) & 0xFFFFu;
}
// v += bf16 residual (two 16-byte chunks = 16 channels)
__device__ __forceinline__ void epi_add_bf16x16(uint64_t (&v2)[8], const uint4& u0, const uint4& u1) {
  v2[0] = add2(v2[0], pk2u(u0.x << 16, u0.x & 0xFFFF0000u));
  v2[1] = add2(v2[1], pk2u(u0.y << 16, u0.y & 0xFFFF0000u));
  v2[2] = add2(v2[2], pk2u(u0.z << 16, u0.z & 0xFFFF0000u));
  v2[3] = add2(v2[3], pk2u(u0.w << 16, u0.w & 0xFFFF0000u));
  v2[4] = add2(v2[4], pk2u(u1.x << 16, u1.x & 0xFFFF0000u));
  v2[5] = add2(v2[5], pk2u(u1.y << 16, u1.y & 0xFFFF0000u));
  v2[6] = add2(v2[6], pk2u(u1.z << 16, u1.z & 0xFFFF0000u));
  v2[7] = add2(v2[7], pk2u(u1.w << 16, u1.w & 0xFFFF0000u));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi);
__device__ __forceinline__ void epi_pack16(const uint64_t (&v2)[8], uint4& u0, uint4& u1) {
  float a0, a1;
  upk2(v2[0], a0, a1); u0.x = pack_bf16x2(a0, a1);
  upk2(v2[1], a0, a1); u0.y = pack_bf16x2(a0, a1);
  upk2(v2[2], a0, a1); u0.z = pack_bf16x2(a0, a1);
  upk2(v2[3], a0, a1); u0.w = pack_bf16x2(a0, a1);
  upk2(v2[4], a0, a1); u1.x = pack_bf16x2(a0, a1);
  upk2(v2[5], a0, a1); u1.y = pack_bf16x2(a0, a1);
  upk2(v2[6], a0, a1); u1.z = pack_bf16x2(a0, a1);
  upk2(v2[7], a0, a1); u1.w = pack_bf16x2(a0, a1);
}
// o = v * (mask bit ? 1 : slope) * chan_scale2   (the LeakyReLU' + Dropout2d step of the backward chain)
__device__ __forceinline__ void epi_masked16(const uint64_t (&v2)[8], uint32_t mbits, float slope, const float* s_cs2,
                                             bool has_cs2, uint64_t (&o2)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a0, a1;
    upk2(v2[i], a0, a1);
    a0 = ((mbits >> (2 * i)) & 1u) ? a0 : a0 * slope;
    a1 = ((mbits >> (2 * i + 1)) & 1u) ? a1 : a1 * slope;
    o2[i] = pk2(a0, a1);
    if (has_cs2) {
      const float2 s2 = *reinterpret_cast<const float2*>(s_cs2 + 2 * i);
      o2[i] = mul2(o2[i], pk2(s2.x, s2.y));
    }
  }
}

// ----------------------------------------------------------------------------- small helpers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace fd
