// Implicit-GEMM 3x3 convolution for NARROW layers (few channels, huge images): "band" loads + resident weights.
//
// conv_tc.cu issues one TMA tile per tap; with <= 64 channels every 128-B row request carries little work and the
// kernel is bound by the L2->SM request rate (~0.3 row requests / clk / SM, measured), not by the tensor pipe.
// Here, per kernel row kh, ONE band of 136 activation rows [m0 + (kh-1)(W+1) - 1, +136) is loaded and the three
// horizontal taps kw = 0,1,2 are read from it by starting the UMMA descriptor 0 / 128 / 256 bytes into the band
// (rows are 128-B lines; measured: the 128-B swizzle phase follows the absolute shared-memory address, so a row offset
// keeps TMA's and UMMA's swizzles consistent -- the descriptor's base_offset field must stay 0, setting it breaks parity).  The weight tiles of all taps are loaded ONCE per CTA and stay resident in
// shared memory (they fit because the layer is narrow).  Row requests per tile drop from 9*(256+2*BN) to 3*272.
// Same operands / epilogue / outputs as conv_tc.cu; replaces it for block-2-like layers and their data gradients.
#include "ssp_common.cuh"
#include "tmap.cuh"

namespace ssp {

struct ConvBandParams {
  CUtensorMap tmA[2];     // box {64, 136}
  CUtensorMap tmB[2];     // box {64, bn}
  long long m_rows, store_rows;
  int m_tiles;
  int kc_per_tap, cin;
  int Wp, HpWp;
  int cout, bn, n_terms;
  uint32_t idesc;
  int stages, stage_bytes, b_bytes, res_bytes, acc_cols;
  float* out; long long out_ld;
  const float* bias; double* stat_sum; double* stat_sq; int epi;
};

namespace {
constexpr int kBandRows = 136;
constexpr int kBandBytes = kBandRows * 128;      // 17408 = 17 swizzle atoms
constexpr int kMaxStagesB = 8;
constexpr int kThreadsB = 256;
}

__global__ void __launch_bounds__(kThreadsB, 1) conv_band_kernel(const __grid_constant__ ConvBandParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* res_base = smem;                                      // resident weights: [(tap, kc)][plane][bn x 128 B]
  uint8_t* stage_base = smem + p.res_bytes;                      // ring of A bands: [plane][136 x 128 B]
  double* acc_sum = (double*)(stage_base + (size_t)p.stages * p.stage_bytes);
  double* acc_sq = acc_sum + p.acc_cols;
  uint64_t* full_bar = (uint64_t*)(acc_sq + p.acc_cols);
  uint64_t* empty_bar = full_bar + kMaxStagesB;
  uint64_t* tfull_bar = empty_bar + kMaxStagesB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(res_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = p.n_terms == 3 ? 2 : 1;
  const int units = 3 * p.kc_per_tap;                            // (kh, kc) per tile

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmA[0]); tma_prefetch_desc(&p.tmB[0]); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  if (p.epi == EPI_STATS)
    for (int i = threadIdx.x; i < 2 * p.acc_cols; i += kThreadsB) acc_sum[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weights: all 9 taps x kc tiles, once ----
      mbar_expect_tx(res_bar, (uint32_t)p.res_bytes);
      for (int tap = 0; tap < 9; tap++)
        for (int kc = 0; kc < p.kc_per_tap; kc++)
          for (int pl = 0; pl < planes; pl++)
            tma_load_2d(res_base + (size_t)((tap * p.kc_per_tap + kc) * planes + pl) * p.b_bytes, &p.tmB[pl], res_bar,
                        tap * p.cin + kc * 64, 0);
      // ---- activation bands ----
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)planes * kBandBytes;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
        const int m0 = t * 128;
        for (int kh = 0; kh < 3; kh++) {
          const int arow = m0 + (kh - 1) * p.Wp - 1;
          for (int kc = 0; kc < p.kc_per_tap; kc++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
            mbar_expect_tx(&full_bar[stage], tx);
            tma_load_2d(sa, &p.tmA[0], &full_bar[stage], kc * 64, arow);
            if (planes == 2) tma_load_2d(sa + kBandBytes, &p.tmA[1], &full_bar[stage], kc * 64, arow);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp converged, one elected lane issues (see elect_one_sync) ----
    mbar_wait(res_bar, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0; int it = 0;
    const uint32_t rb = smem_u32(res_base);
    const int rem_k = p.cin - (p.kc_per_tap - 1) * 64;            // channels in the last 64-wide chunk
    const int ksteps_last = rem_k >= 64 ? 4 : (rem_k + 15) / 16;   // all-zero K steps (TMA zero fill) are skipped
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.bn);
      uint32_t acc = 0;
      for (int u = 0; u < units; u++) {
        const int kh = u / p.kc_per_tap, kc = u % p.kc_per_tap;
        const int ksteps = (kc == p.kc_per_tap - 1) ? ksteps_last : 4;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
        const uint32_t b0 = rb + (uint32_t)((kh * 3 * p.kc_per_tap + kc) * planes) * p.b_bytes;
        const uint32_t b_tap = (uint32_t)(p.kc_per_tap * planes) * p.b_bytes;
        if (elect_one_sync()) {
#pragma unroll
          for (int kw = 0; kw < 3; kw++) {
            const uint32_t a_hi = sa + kw * 128, a_lo = a_hi + kBandBytes;     // tap kw starts kw rows into the band
            const uint32_t b_hi = b0 + kw * b_tap, b_lo = b_hi + p.b_bytes;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (k < ksteps) {
                const uint64_t dah = umma_desc_k_sw128(a_hi + k * 32);
                const uint64_t dbh = umma_desc_k_sw128(b_hi + k * 32);
                if (planes == 2) {
                  umma_f16(d_tmem, umma_desc_k_sw128(a_lo + k * 32), dbh, p.idesc, acc); acc = 1;
                  umma_f16(d_tmem, dah, umma_desc_k_sw128(b_lo + k * 32), p.idesc, 1);
                }
                umma_f16(d_tmem, dah, dbh, p.idesc, acc); acc = 1;
              }
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        acc = 1;
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (elect_one_sync()) umma_commit(&tfull_bar[buf]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      const long long m = (long long)t * 128 + q * 32 + lane;
      bool valid = false;
      if (m < p.m_rows) { const int rem = (int)(m % p.HpWp); valid = (rem / p.Wp >= 1) && (rem % p.Wp >= 1); }
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.bn);
      float* orow = p.out + m * p.out_ld;
      const bool can_store = m < p.store_rows;
      for (int ch = 0; ch < p.bn / 32; ch++) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + ch * 32, r);
        tmem_ld_wait();
        const int c0 = ch * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
        if (p.epi == EPI_BIAS) {
#pragma unroll
          for (int j = 0; j < 32; j++) if (c0 + j < p.cout) v[j] += __ldg(p.bias + c0 + j);
        }
        if (can_store) {
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
          if (c0 + lane < p.cout) { atomicAdd(&acc_sum[c0 + lane], (double)cs); atomicAdd(&acc_sq[c0 + lane], (double)cq); }
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

// returns SSP_OK, or 1 when the layer is not eligible (caller falls back to the per-tap kernel)
int conv_gemm_band(const void* a_hi, const void* a_lo, long long a_rows, int a_ld, int cin,
                   const void* b_hi, const void* b_lo, int b_rows, int b_ld, int a_fmt, int b_fmt,
                   int N, int H, int W, int taps, int cout, float* out, int out_ld, long long out_rows,
                   int epi, const float* bias, double* stat_sum, double* stat_sq, cudaStream_t stream) {
  if (taps != 9 || cout > 256) return 1;
  if (!a_hi || !b_hi || !out || (a_ld % 8) || (b_ld % 8) || (out_ld % 4)) return fail_msg(SSP_ERR_ARG, "conv_gemm_band: bad argument");
  ConvBandParams p;
  Geom g{N, H, W};
  p.n_terms = (a_lo && b_lo) ? 3 : 1;
  const int planes = p.n_terms == 3 ? 2 : 1;
  int bn = ((cout + 31) / 32) * 32;
  if (bn > 128 && bn < 256) bn = 256;
  if (bn > 64 && bn < 128) bn = 128;
  p.bn = bn; p.b_bytes = bn * 128;
  p.kc_per_tap = (cin + 63) / 64; p.cin = cin;
  p.res_bytes = 9 * p.kc_per_tap * planes * p.b_bytes;
  p.stage_bytes = planes * kBandBytes;
  p.acc_cols = ((cout + 31) / 32) * 32;
  const int fixed = 2 * p.acc_cols * 8 + (2 * kMaxStagesB + 5) * 8 + 16 + 1024;
  int stages = (227 * 1024 - fixed - p.res_bytes) / p.stage_bytes;
  if (stages > kMaxStagesB) stages = kMaxStagesB;
  if (stages < 2) return 1;                       // weights do not fit next to two bands: not a narrow layer
  p.stages = stages;
  p.m_rows = g.m_rows(); p.store_rows = out_rows;
  p.m_tiles = (int)((p.m_rows + 127) / 128);
  p.Wp = g.Wp(); p.HpWp = g.HpWp(); p.cout = cout;
  p.idesc = umma_idesc_f16(a_fmt, b_fmt, 0, 0, bn);
  p.out = out; p.out_ld = out_ld; p.bias = bias; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.epi = epi;
  if (epi == EPI_STATS && (!stat_sum || !stat_sq)) return fail_msg(SSP_ERR_ARG, "conv_gemm_band: statistics buffers missing");
  int rc = 0;
  rc |= tmap_2d_16bit(&p.tmA[0], a_hi, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, kBandRows, a_fmt == FMT_BF16);
  rc |= tmap_2d_16bit(&p.tmB[0], b_hi, (uint64_t)9 * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, bn, b_fmt == FMT_BF16);
  if (planes == 2) {
    rc |= tmap_2d_16bit(&p.tmA[1], a_lo, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, kBandRows, a_fmt == FMT_BF16);
    rc |= tmap_2d_16bit(&p.tmB[1], b_lo, (uint64_t)9 * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, bn, b_fmt == FMT_BF16);
  }
  if (rc) return fail_msg(SSP_ERR_DRIVER, "conv_gemm_band: cuTensorMapEncodeTiled failed");
  static int sms = 0, configured = 0;
  if (!sms) sms = ssp_sm_count();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  conv_band_kernel<<<grid, kThreadsB, p.res_bytes + stages * p.stage_bytes + fixed, stream>>>(p);
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
