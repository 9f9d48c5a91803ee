// Shared device/host helpers for the sm_100a kernels of the singleshotpose hot path.
//
// Data layout ("padded-flat NHWC"): a feature map (N, C, H, W) is stored as a row-major matrix
// [rows, C] whose row index is
//     row(n, h, w) = n*(H+1)*(W+1) + (h+1)*(W+1) + (w+1)
// i.e. every image row is preceded by ONE zero pad pixel (which is also the right pad of the previous
// row) and every image by ONE zero pad row (also the bottom pad of the previous image).  A 3x3 / pad 1
// convolution then is, for every tap (kh, kw), the SAME matrix shifted by the constant row offset
// (kh-1)*(W+1) + (kw-1): each tap of the implicit GEMM is a plain 2-D TMA tile load, no im2col.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/ssp_b200.h"   // SSP_OK / SSP_ERR_* / SSP_* enums shared with the C ABI

#define SSP_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    if (e__ != cudaSuccess) return ssp::fail_cuda(e__, __FILE__, __LINE__); \
  } while (0)

namespace ssp {

int fail_cuda(cudaError_t e, const char* file, int line);   // abi.cu: records message, returns SSP_ERR_CUDA
int fail_msg(int code, const char* msg);

// SMs the persistent kernels size their grids for.  SSP_SM_LIMIT=n (even, < the device's count) leaves SMs free, e.g. for NCCL's kernels in
// the data-parallel step: a collective cannot co-reside with a 227-KB GEMM CTA, it needs SMs of its own to overlap with the backward pass.
inline int ssp_sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    const char* e = getenv("SSP_SM_LIMIT");
    if (e) { const int l = atoi(e) & ~1; if (l >= 2 && l < n) n = l; }
  }
  return n;
}
enum Fmt16 { FMT_F16 = 0, FMT_BF16 = 1 };
// GEMM epilogues: plain fp32 store | store + per-channel sum / sum-of-squares over valid pixels | + bias
enum { EPI_F32 = 0, EPI_STATS = 1, EPI_BIAS = 2, EPI_BNACT = 4, EPI_F16 = 8 };
// inference-mode fusion: z = leaky(acc*scale[c] + shift[c]) written straight into the consumer's fp16 hi/lo operand planes
struct FusedAct {
  const float* scale; const float* shift; float slope;
  uint16_t* d_hi; uint16_t* d_lo; long long d_ld; int d_c0;
};
struct Geom {
  int N, H, W;
  __host__ __device__ int Wp() const { return W + 1; }
  __host__ __device__ int HpWp() const { return (H + 1) * (W + 1); }
  __host__ __device__ long long m_rows() const { return (long long)N * (H + 1) * (W + 1); }
  __host__ __device__ long long row(int n, int h, int w) const {
    return (long long)n * HpWp() + (long long)(h + 1) * Wp() + (w + 1);
  }
};

__host__ __device__ inline long long flat_alloc_rows(int N, int H, int W) {
  long long m = (long long)N * (H + 1) * (W + 1) + (W + 1) + 2;
  return ((m + 127) / 128) * 128 + 128;
}

// ---------------------------------------------------------------------------------------------
// 16-bit conversions with runtime format
__device__ __forceinline__ float cvt16_to_f32(uint16_t v, int fmt) {
  return fmt == FMT_F16 ? __half2float(__ushort_as_half(v)) : __bfloat162float(__ushort_as_bfloat16(v));
}
__device__ __forceinline__ uint16_t cvt_f32_to_16(float f, int fmt) {
  // fp16 saturates instead of overflowing to inf (loss-scaled gradients)
  return fmt == FMT_F16 ? __half_as_ushort(__float2half_rn(fminf(fmaxf(f, -65504.f), 65504.f)))
                        : __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
// split an fp32 value into fp16 hi + fp16 lo (hi + lo carries ~22 mantissa bits)
__device__ __forceinline__ void split_f16(float f, uint16_t& hi, uint16_t& lo) {
  __half h = __float2half_rn(f);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn(f - __half2float(h)));
}

// ---------------------------------------------------------------------------------------------
// PTX wrappers (mbarrier, TMA, tcgen05)
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
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
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// Bounded wait: a protocol bug must surface as a trapped kernel (cudaErrorLaunchFailure), never as a hang.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { __trap(); }
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// One lane of a CONVERGED warp.  The role loops run with the whole warp converged and only the MMA / TMA issue itself
// under this predicate: operands then live in uniform registers (UTCHMMA / UTMALDG take UR operands); issuing from a
// lane-divergent region makes the compiler wrap every instruction in an ELECT / R2UR.BROADCAST retry loop (~10 extra
// instructions per MMA, which made N <= 128 layers issue-bound).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
// SWIZZLE_128B K-major descriptor from its two 32-bit halves: hi is constant (SBO = 1024 B, version 1, layout 2)
#define SSP_DESC_HI_SW128 0x40004040u
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return ((uint64_t)SSP_DESC_HI_SW128 << 32) | (uint64_t)(((smem_addr >> 4) & 0x3FFFu) | 0x10000u);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16/bf16 operands, fp32 accumulate), issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, SWIZZLE_128B (sm_100 encoding: version 1 at bit 46, layout 2 at bits 61-63).
//   K-major tile  [rows][64 x 16-bit] (one 128-byte swizzled line per row): LBO unused (1), SBO = 8 rows = 1024 B.
//   MN-major tile [k rows][64 x 16-bit]: LBO = byte distance between 64-element MN groups, SBO = 8 k-rows = 1024 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 with fp32 accumulation, M = 128.
__host__ __device__ inline uint32_t umma_idesc_f16(int a_fmt, int b_fmt, int a_mn_major, int b_mn_major, int n) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// optim.SGD(momentum, dampening = 0, weight_decay), train.py:388:  g += wd*p ; v = mu*v + g ; p -= lr*v  (torch's first step,
// v = g, is the same with v0 = 0).  Explicit FMAs: every kernel that applies the update produces the same bits.
__device__ __forceinline__ void sgd_update(float& p, float g, float& v, float lr, float mu, float wd, float gscale) {
  const float vn = fmaf(mu, v, fmaf(wd, p, g * gscale));
  v = vn;
  p = fmaf(-lr, vn, p);
}

// Sum over the 32 lanes of v[j], delivered to lane j (31 shuffles instead of 32*5).
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; i++) {
      float send = up ? v[i] : v[i + off];
      float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

}  // namespace ssp
