// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM accumulators, TMA-fed).
//
//   out[m, n] = sum_{tap} sum_{c} A[m + shift(tap), c] * B[n, tap*cin + c]
//
// A  = activation matrix in the padded-flat NHWC layout (ssp_common.cuh), 16-bit, optionally as a hi/lo
//      pair (x = hi + lo) so that three MMAs  Ahi*Bhi + Alo*Bhi + Ahi*Blo  reproduce an fp32 product to
//      ~2^-22: the yolo-pose stack amplifies operand rounding ~50x, a single fp16/tf32 pass misses the
//      reference's logits by 3e-2 ("numerics").
// B  = weights [cout][taps*cin] (K contiguous), same hi/lo convention.
// Every tap is one plain 2-D TMA tile at a shifted row coordinate (negative / past-the-end rows are
// zero-filled by TMA), so 3x3 convs need no im2col buffer; 1x1 convs and the im2col'ed first layer are
// the taps==1 case.  Replaces nn.Conv2d in reference darknet.py:156-160 (forward) and, with re-packed
// weights, its data gradient (train.py:103 autograd).
//
// CTA = 8 warps, persistent over (m-tile, n-tile) pairs:
//   warp 0   TMA producer (one lane)       smem ring of `stages` x {A_hi, A_lo, B_hi, B_lo}
//   warp 1   MMA issuer (one lane)         tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue                     tcgen05.ld -> registers -> (bias | BN statistics) -> global fp32
#include "ssp_common.cuh"
#include "tmap.cuh"

namespace ssp {

struct ConvTcParams {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];
  long long m_rows;       // rows of the output matrix that exist (N*(H+1)*(W+1))
  long long store_rows;   // rows that may be written (allocation bound)
  int m_tiles, n_tiles;
  int kc_per_tap, cin, taps;
  int shifts[9];
  int Wp, HpWp;
  int cout, bn, n_terms;
  uint32_t idesc;
  int stages, stage_bytes, b_bytes;
  float* out;
  long long out_ld;
  const float* bias;
  double* stat_sum;
  double* stat_sq;
  int epi;
  FusedAct fa;            // EPI_BNACT only
};

static constexpr int kABytes = 128 * 128;     // 128 rows x 64 x 2 B
static constexpr int kMaxStages = 8;
static constexpr int kAccCols = 1024;         // per-CTA statistics accumulators (channels)
static constexpr int kThreads = 256;

template <bool FUSED>   // FUSED: the inference epilogue (EPI_BNACT) -- a separate instantiation keeps the training kernel's code unchanged
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-B aligned: round up to 1024 (swizzle-128B atoms)
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  double* acc_sum = (double*)(smem + (size_t)p.stages * p.stage_bytes);
  double* acc_sq = acc_sum + kAccCols;
  uint64_t* full_bar = (uint64_t*)(acc_sq + kAccCols);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;
  const int kblocks = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
    if (p.n_terms == 3) { tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmB[1]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  if (p.epi == EPI_STATS)
    for (int i = threadIdx.x; i < 2 * kAccCols; i += kThreads) acc_sum[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)(p.n_terms == 3 ? 2 : 1) * (uint32_t)(kABytes + p.b_bytes);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t / p.n_tiles, nt = t % p.n_tiles;
        const int m0 = mt * 128, n0 = nt * p.bn;
        for (int tap = 0; tap < p.taps; tap++) {
          const int arow = m0 + p.shifts[tap];
          for (int kc = 0; kc < p.kc_per_tap; kc++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
            mbar_expect_tx(&full_bar[stage], tx);
            const int kcol_b = tap * p.cin + kc * 64;
            tma_load_2d(sa, &p.tmA[0], &full_bar[stage], kc * 64, arow);
            if (p.n_terms == 3) {
              tma_load_2d(sa + kABytes, &p.tmA[1], &full_bar[stage], kc * 64, arow);
              tma_load_2d(sa + 2 * kABytes, &p.tmB[0], &full_bar[stage], kcol_b, n0);
              tma_load_2d(sa + 2 * kABytes + p.b_bytes, &p.tmB[1], &full_bar[stage], kcol_b, n0);
            } else {
              tma_load_2d(sa + kABytes, &p.tmB[0], &full_bar[stage], kcol_b, n0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.bn);
        uint32_t acc = 0;
        for (int kb = 0; kb < kblocks; kb++) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
          const uint32_t a_hi = sa, a_lo = sa + kABytes;
          const uint32_t b_hi = (p.n_terms == 3) ? sa + 2 * kABytes : sa + kABytes;
          const uint32_t b_lo = b_hi + p.b_bytes;
#pragma unroll
          for (int k = 0; k < 4; k++) {     // 4 x UMMA_K(16) = 64 K elements; +32 B inside the swizzle atom
            const uint64_t dah = umma_desc_sw128(a_hi + k * 32, 16, 1024);
            const uint64_t dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
            if (p.n_terms == 3) {
              const uint64_t dal = umma_desc_sw128(a_lo + k * 32, 16, 1024);
              const uint64_t dbl = umma_desc_sw128(b_lo + k * 32, 16, 1024);
              umma_f16(d_tmem, dal, dbh, p.idesc, acc);  acc = 1;   // small cross terms first
              umma_f16(d_tmem, dah, dbl, p.idesc, 1);
            }
            umma_f16(d_tmem, dah, dbh, p.idesc, acc);  acc = 1;
          }
          umma_commit(&empty_bar[stage]);                 // smem slot free once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[buf]);                     // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (128 threads = 128 TMEM lanes)
    const int q = warp - 4;
    int it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      const int mt = t / p.n_tiles, nt = t % p.n_tiles;
      const long long m = (long long)mt * 128 + q * 32 + lane;
      const int n0 = nt * p.bn;
      bool valid = false;
      if (m < p.m_rows) {
        const int rem = (int)(m % p.HpWp);
        valid = (rem / p.Wp >= 1) && (rem % p.Wp >= 1);
      }
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.bn);
      float* orow = p.out ? p.out + m * p.out_ld : nullptr;
      const bool can_store = m < p.store_rows;
      for (int ch = 0; ch < p.bn / 32; ch++) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + ch * 32, r);
        tmem_ld_wait();
        const int c0 = n0 + ch * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
        if constexpr (FUSED) {
          // folded BatchNorm (running statistics) + LeakyReLU in the epilogue; only valid rows are written so that the
          // consumer's pad rows stay zero; the fp32 Y tensor is never materialised in this mode
          if (valid && c0 + 32 <= p.cout) {
            uint32_t ph[16], pl[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              uint16_t h0, l0, h1, l1;
              float z0 = fmaf(v[j], __ldg(p.fa.scale + c0 + j), __ldg(p.fa.shift + c0 + j));
              float z1 = fmaf(v[j + 1], __ldg(p.fa.scale + c0 + j + 1), __ldg(p.fa.shift + c0 + j + 1));
              z0 = z0 > 0.f ? z0 : z0 * p.fa.slope; z1 = z1 > 0.f ? z1 : z1 * p.fa.slope;
              split_f16(z0, h0, l0); split_f16(z1, h1, l1);
              ph[j >> 1] = h0 | ((uint32_t)h1 << 16); pl[j >> 1] = l0 | ((uint32_t)l1 << 16);
            }
            uint4* dh = reinterpret_cast<uint4*>(p.fa.d_hi + m * p.fa.d_ld + p.fa.d_c0 + c0);
            uint4* dl = reinterpret_cast<uint4*>(p.fa.d_lo + m * p.fa.d_ld + p.fa.d_c0 + c0);
#pragma unroll
            for (int j = 0; j < 4; j++) {
              dh[j] = make_uint4(ph[4 * j], ph[4 * j + 1], ph[4 * j + 2], ph[4 * j + 3]);
              dl[j] = make_uint4(pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
            }
          }
          continue;
        }
        if (p.epi == EPI_BIAS) {
#pragma unroll
          for (int j = 0; j < 32; j++) if (c0 + j < p.cout) v[j] += __ldg(p.bias + c0 + j);
        }
        if (can_store && p.epi != 3) {     // epi 3: diagnostic mode without the global store
          if (c0 + 32 <= p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (c0 + j < p.cout) orow[c0 + j] = v[j];
          }
        }
        if (p.epi == EPI_STATS) {
          float s1[32], s2[32];
#pragma unroll
          for (int j = 0; j < 32; j++) { const float x = valid ? v[j] : 0.f; s1[j] = x; s2[j] = x * x; }
          const float cs = warp_transpose_sum32(s1, lane);
          const float cq = warp_transpose_sum32(s2, lane);
          if (c0 + lane < p.cout) {
            atomicAdd(&acc_sum[c0 + lane], (double)cs);
            atomicAdd(&acc_sq[c0 + lane], (double)cq);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if (p.epi == EPI_STATS) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = threadIdx.x - 128; c < p.cout; c += 128) {
        const double a = acc_sum[c], b = acc_sq[c];
        if (a != 0.0 || b != 0.0) { atomicAdd(p.stat_sum + c, a); atomicAdd(p.stat_sq + c, b); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int g_num_sms = 0;

int conv_gemm_tc(const void* a_hi, const void* a_lo, long long a_rows, int a_ld, int cin,
                 const void* b_hi, const void* b_lo, int b_rows, int b_ld, int a_fmt, int b_fmt,
                 int N, int H, int W, int taps, int cout, float* out, int out_ld, long long out_rows,
                 int epi, const float* bias, double* stat_sum, double* stat_sq, cudaStream_t stream, const FusedAct* fa) {
  if (fa) {
    if (!fa->scale || !fa->shift || !fa->d_hi || !fa->d_lo || (cout % 32) || (fa->d_ld % 8) || (fa->d_c0 % 8))
      return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: fused BN+activation epilogue needs cout % 32 == 0 and 16-B aligned destination rows");
    epi = EPI_BNACT;
  }
  if (!a_hi || !b_hi || (!out && !fa) || (taps != 1 && taps != 9) || cin <= 0 || cout <= 0) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: bad argument");
  if ((a_ld % 8) || (b_ld % 8)) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: leading dimensions must be multiples of 8 elements (16 B)");
  if (epi == EPI_STATS && (cout > kAccCols || !stat_sum || !stat_sq)) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: statistics need cout <= 1024 and buffers");
  if (epi == EPI_BIAS && !bias) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: bias missing");
  if (!fa && ((out_ld % 4) || ((uintptr_t)out % 16))) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: output must be 16-B aligned with ld % 4 == 0");
  if (!g_num_sms) {
    g_num_sms = ssp_sm_count();
  }
  ConvTcParams p;
  Geom g{N, H, W};
  p.n_terms = (a_lo && b_lo) ? 3 : 1;
  int bn = ((cout + 31) / 32) * 32;
  if (bn > 256) bn = 256;
  if (bn > 128 && bn < 256) bn = 256;
  if (bn > 64 && bn < 128) bn = 128;
  p.bn = bn;
  p.b_bytes = bn * 128;
  p.m_rows = g.m_rows();
  p.store_rows = out_rows;
  p.m_tiles = (int)((p.m_rows + 127) / 128);
  p.n_tiles = (cout + bn - 1) / bn;
  p.kc_per_tap = (cin + 63) / 64;
  p.cin = cin;
  p.taps = taps;
  for (int t = 0; t < 9; t++) p.shifts[t] = (taps == 9) ? ((t / 3) - 1) * g.Wp() + ((t % 3) - 1) : 0;
  p.Wp = g.Wp(); p.HpWp = g.HpWp();
  p.cout = cout;
  p.idesc = umma_idesc_f16(a_fmt, b_fmt, 0, 0, bn);
  p.stage_bytes = (p.n_terms == 3 ? 2 : 1) * (kABytes + p.b_bytes);
  const int fixed = 2 * kAccCols * 8 + (2 * kMaxStages + 4) * 8 + 16 + 1024;
  int stages = (227 * 1024 - fixed) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc: tile does not fit shared memory");
  p.stages = stages;
  p.out = out; p.out_ld = out_ld; p.bias = bias; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.epi = epi;
  if (fa) p.fa = *fa; else p.fa = FusedAct{nullptr, nullptr, 1.f, nullptr, nullptr, 0, 0};
  int rc = 0;
  rc |= tmap_2d_16bit(&p.tmA[0], a_hi, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 128, a_fmt == FMT_BF16);
  rc |= tmap_2d_16bit(&p.tmB[0], b_hi, (uint64_t)taps * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, bn, b_fmt == FMT_BF16);
  if (p.n_terms == 3) {
    rc |= tmap_2d_16bit(&p.tmA[1], a_lo, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 128, a_fmt == FMT_BF16);
    rc |= tmap_2d_16bit(&p.tmB[1], b_lo, (uint64_t)taps * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, bn, b_fmt == FMT_BF16);
  }
  if (rc) return fail_msg(SSP_ERR_DRIVER, "conv_gemm_tc: cuTensorMapEncodeTiled failed (no driver, or misaligned operand)");
  const int smem_bytes = stages * p.stage_bytes + fixed;
  static int configured = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  const int total = p.m_tiles * p.n_tiles;
  const int grid = total < g_num_sms ? total : g_num_sms;
  if (fa) conv_tc_kernel<true><<<grid, kThreads, smem_bytes, stream>>>(p);
  else conv_tc_kernel<false><<<grid, kThreads, smem_bytes, stream>>>(p);
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
