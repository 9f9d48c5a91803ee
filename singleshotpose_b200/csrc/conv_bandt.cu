// Implicit-GEMM convolution for layers with FEW OUTPUT CHANNELS, operands swapped: the tensor cores compute the TRANSPOSED tile
//
//     D^T[n, m] = sum_tap sum_c  Wt[n, tap*cin + c] * X[m + shift(tap), c]          (n = output channel, m = pixel row)
//
// i.e. the weights are the M side (128 TMEM lanes) and a band of 128 / 256 PIXELS is the N side of every tcgen05.mma.
// Why: with cout <= 64 the forward GEMM of conv_band.cu issues 128 x 64 x 16 instructions, and an MMA with N <= 64 costs ~95 clk
// whatever its size (measured) -- block 2 of yolo-pose.cfg (32 -> 64 channels on 208 x 208) ran at 0.64 ms
// forward / 0.49 ms data gradient / batch 64, MMA-issue-bound at a sixth of the pipe.  Swapped, every instruction is 128 x 256 x 16
// (or 128 x 128 x 16) whatever the layer's width.  For the split-fp16 forward (x = hi + lo, three products) the unused half of
// the M side carries the second weight plane: rows 0-63 = W_hi, rows 64-127 = W_lo, so ONE instruction against X_hi yields
// W_hi X_hi and W_lo X_hi, one against X_lo yields W_hi X_lo (and W_lo X_lo, 2^-22, harmless): 2 full-width MMAs per K step
// instead of 3 narrow ones, and the epilogue adds TMEM lanes n and n + 64 through shared memory.
//   * activations: per kernel row kh ONE band of np + 8 rows (conv_band.cu's trick): the three horizontal taps are the same band
//     read 0 / 128 / 256 bytes further in (the B descriptor's start address; the 128-B swizzle follows the absolute address);
//   * weights: all taps resident in shared memory, loaded once per CTA;
//   * epilogue: thread = output channel (TMEM lane), 32 consecutive pixels per tcgen05.ld; for every pixel the warp stores 32
//     consecutive channels = one 128-B line of the row-major output; BN statistics per thread (its channel) in fp64, pixel
//     validity (pad rows of the padded-flat layout) from a per-tile ballot mask.
// Same operands / outputs as conv_band.cu (drop-in behind ssp_conv_gemm, SSP_IMPL_BANDT); not eligible (returns 1): cout > 64 with
// split operands, cout > 128 single-term, bias epilogue, weights that do not fit next to two bands.
// Replaces nn.Conv2d of reference darknet.py:156-160 for the narrow blocks, and their data gradients (train.py:103).
#include "ssp_common.cuh"
#include "tmap.cuh"

namespace ssp {

struct ConvBandTParams {
  CUtensorMap tmX[2][2];   // [plane][part]: band boxes {64, rows0} and {64, rows1}
  CUtensorMap tmW[2];      // [plane]: weight box {64, wbox_rows}
  long long m_rows, store_rows;
  int m_tiles, np;         // pixels per tile = MMA N (128 or 256)
  int taps, nk;            // nk = 3 (3x3) or 1 (1x1)
  int kc_per_tap, cin;
  int Wp, HpWp;
  int cout, stacked;
  int ovl;                 // 32-channel 3x3 layer through the overlapping-row tensor map: a band row = [pixel | pixel + 1], K chunk 0 = taps kw 0,1, chunk 1 = tap kw 2
  uint32_t idesc;
  int stages, stage_bytes, plane_bytes, rows0, rows1;
  int w_tile_bytes, res_bytes;
  float* out; long long out_ld;
  double* stat_sum; double* stat_sq; int epi;
};

namespace {
constexpr int kMaxStagesT = 6;
constexpr int kThreadsT = 256;
constexpr int kXchBytes = 2 * 2 * 16 * 32 * 4;     // [half][giver warp][16 columns][32 lanes] fp32
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(kThreadsT, 1) conv_bandt_kernel(const __grid_constant__ ConvBandTParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* res_base = smem;                                       // resident weight tiles [(tap, kc)][w_tile_bytes]
  uint8_t* stage_base = smem + p.res_bytes;                       // ring of bands: [plane][(rows0 + rows1) x 128 B]
  float* xch = (float*)(stage_base + (size_t)p.stages * p.stage_bytes);
  uint32_t* vmask = (uint32_t*)((uint8_t*)xch + kXchBytes);       // [2 tile buffers][8 words]
  uint64_t* full_bar = (uint64_t*)(vmask + 16);
  uint64_t* empty_bar = full_bar + kMaxStagesT;
  uint64_t* tfull_bar = empty_bar + kMaxStagesT;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(res_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int planes = p.stacked ? 2 : 1;
  const int units = p.nk * p.kc_per_tap;                          // (kh, kc) per tile

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmX[0][0]); tma_prefetch_desc(&p.tmW[0]); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    mbar_init(res_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ---- weights: every (tap, kc) tile, once.  stacked: rows 0-63 <- W_hi, rows 64-127 <- W_lo of the same 16 KB tile ----
      const int wbox_rows = p.stacked ? 64 : p.w_tile_bytes / 128;
      mbar_expect_tx(res_bar, (uint32_t)((p.ovl ? 6 : p.taps * p.kc_per_tap) * planes * wbox_rows * 128));
      const int ntile = p.ovl ? 6 : p.taps * p.kc_per_tap;
      for (int ti = 0; ti < ntile; ti++) {
        // ovl: tile 2 kh = the 64 K columns of taps (kh, 0) and (kh, 1), tile 2 kh + 1 = tap (kh, 2) (its upper 32 columns are never multiplied)
        const int kcol = p.ovl ? (ti >> 1) * 3 * p.cin + (ti & 1) * 64 : (ti / p.kc_per_tap) * p.cin + (ti % p.kc_per_tap) * 64;
        uint8_t* wt = res_base + (size_t)ti * p.w_tile_bytes;
        tma_load_2d(wt, &p.tmW[0], res_bar, kcol, 0);
        if (p.stacked) tma_load_2d(wt + 64 * 128, &p.tmW[1], res_bar, kcol, 0);
      }
      // ---- activation bands ----
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = (uint32_t)planes * (uint32_t)p.plane_bytes;
      for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x) {
        const int m0 = t * p.np;
        for (int kh = 0; kh < p.nk; kh++) {
          const int arow = p.nk == 3 ? m0 + (kh - 1) * p.Wp - 1 : m0;
          for (int kc = 0; kc < p.kc_per_tap; kc++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
            mbar_expect_tx(&full_bar[stage], tx);
            for (int pl = 0; pl < planes; pl++) {
              tma_load_2d(sa + (size_t)pl * p.plane_bytes, &p.tmX[pl][0], &full_bar[stage], kc * 64, arow);
              if (p.rows1) tma_load_2d(sa + (size_t)pl * p.plane_bytes + (size_t)p.rows0 * 128, &p.tmX[pl][1], &full_bar[stage], kc * 64, arow + p.rows0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer: whole warp converged, one elected lane issues ----
    mbar_wait(res_bar, 0);
    tc_fence_after();
    int stage = 0; uint32_t phase = 0; int it = 0;
    const uint32_t rb = smem_u32(res_base);
    const int rem_k = p.cin - (p.kc_per_tap - 1) * 64;
    const int ksteps_last = rem_k >= 64 ? 4 : (rem_k + 15) / 16;   // all-zero K steps (TMA zero fill) are skipped
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.np);
      uint32_t acc = 0;
      for (int u = 0; u < units; u++) {
        const int kh = u / p.kc_per_tap, kc = u % p.kc_per_tap;
        const int ksteps = (kc == p.kc_per_tap - 1) ? ksteps_last : 4;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
        if (elect_one_sync()) {
          const int ngrp = p.ovl ? 2 : p.nk;
          for (int kw = 0; kw < ngrp; kw++) {
            // ovl: group 0 = taps kw 0,1 (64 K of a [pixel | pixel + 1] row), group 1 = tap kw 2 (32 K, band read two rows further in)
            const uint32_t wt = rb + (uint32_t)((p.ovl ? kh * 2 + kw : (kh * p.nk + kw) * p.kc_per_tap + kc) * p.w_tile_bytes);
            const uint32_t x_hi = sa + (p.ovl ? 2 * kw : kw) * 128, x_lo = x_hi + (uint32_t)p.plane_bytes;     // tap kw starts kw rows into the band
            const int ks = p.ovl ? (kw == 0 ? 4 : 2) : ksteps;
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (k < ks) {
                const uint64_t da = umma_desc_k_sw128(wt + k * 32);
                if (p.stacked) { umma_f16(d_tmem, da, umma_desc_k_sw128(x_lo + k * 32), p.idesc, acc); acc = 1; }   // small terms first
                umma_f16(d_tmem, da, umma_desc_k_sw128(x_hi + k * 32), p.idesc, acc); acc = 1;
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
    // ---- epilogue: thread = TMEM lane = weight row; columns = pixels ----
    const int q = warp - 4, et = threadIdx.x - 128;
    const bool stats = p.epi == EPI_STATS;
    const int n_own = p.stacked ? ((q & 1) * 32 + lane) : (q * 32 + lane);          // output channel of this thread
    const bool warp_live = p.stacked || (q * 32 < p.cout);                             // non-stacked: warps above cout have nothing to do
    double d1 = 0.0, d2 = 0.0;
    int it = 0;
    for (int t = blockIdx.x; t < p.m_tiles; t += gridDim.x, it++) {
      const int buf = it & 1;
      const long long m0 = (long long)t * p.np;
      if (stats) {
        for (int i = et; i < p.np; i += 128) {
          const long long m = m0 + i;
          bool valid = false;
          if (m < p.m_rows) { const int rem = (int)(m % p.HpWp); valid = (rem / p.Wp >= 1) && (rem % p.Wp >= 1); }
          const uint32_t b = __ballot_sync(0xffffffffu, valid);
          if (lane == 0) vmask[buf * 8 + (i >> 5)] = b;
        }
        epi_bar();
      }
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.np);
      float ts1 = 0.f, ts2 = 0.f;
      if (!p.stacked) {
        if (warp_live) {
          for (int ch = 0; ch < p.np / 32; ch++) {
            uint32_t r[32];
            tmem_ld_32x32(t_row + ch * 32, r);
            tmem_ld_wait();
            const uint32_t vm = stats ? vmask[buf * 8 + ch] : 0u;
            if (n_own < p.cout) {
              // branch-free inner loops: a predicated store per element compiles to a BSSY / BRA / BSYNC triple per column and the
              // epilogue, not the tensor pipe, bounded the kernel (round 2, ncu: 2k clk per 32-column chunk, 28 % tensor active)
              float* o = p.out + (m0 + ch * 32) * p.out_ld + n_own;
              if (p.epi == EPI_F16) {                  // fp16 data gradient: the warp writes 64 B of every pixel row
                uint16_t* o16 = reinterpret_cast<uint16_t*>(p.out) + (m0 + ch * 32) * p.out_ld + n_own;
                if (m0 + ch * 32 + 32 <= p.store_rows) {
#pragma unroll
                  for (int j = 0; j < 32; j++) { *o16 = cvt_f32_to_16(__uint_as_float(r[j]), FMT_F16); o16 += p.out_ld; }
                } else {
                  for (int j = 0; j < 32; j++) if (m0 + ch * 32 + j < p.store_rows) o16[(long long)j * p.out_ld] = cvt_f32_to_16(__uint_as_float(r[j]), FMT_F16);
                }
              } else if (m0 + ch * 32 + 32 <= p.store_rows) {
#pragma unroll
                for (int j = 0; j < 32; j++) { *o = __uint_as_float(r[j]); o += p.out_ld; }
              } else {
                for (int j = 0; j < 32; j++) if (m0 + ch * 32 + j < p.store_rows) o[(long long)j * p.out_ld] = __uint_as_float(r[j]);
              }
              if (stats) {
#pragma unroll
                for (int j = 0; j < 32; j++) { const float v = ((vm >> j) & 1u) ? __uint_as_float(r[j]) : 0.f; ts1 += v; ts2 = fmaf(v, v, ts2); }
              }
            }
          }
        }
      } else {
        // lanes n and n + 64 hold the two halves of the sum: warps 2,3 give columns [0,16) of every 32-column group to warps 0,1,
        // warps 0,1 give columns [16,32) to warps 2,3; each side finishes, stores and counts its 16 columns
        const int side = q >> 1;                                       // 0: finishes columns [0,16), 1: columns [16,32)
        float* give = xch + ((size_t)((1 - side) * 2 + (q & 1)) * 16) * 32;       // buffer of the half this warp gives away
        const float* take = xch + ((size_t)(side * 2 + (q & 1)) * 16) * 32;       // the half it finishes, written by warp q ^ 2
        for (int g = 0; g < p.np / 32; g++) {
          uint32_t r0[16], r1[16];
          tmem_ld_32x16(t_row + g * 32, r0);
          tmem_ld_32x16(t_row + g * 32 + 16, r1);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; j++) give[j * 32 + lane] = __uint_as_float(side == 0 ? r1[j] : r0[j]);
          epi_bar();
          const uint32_t vm = stats ? (vmask[buf * 8 + g] >> (side * 16)) : 0u;
          if (n_own < p.cout) {
            const long long mb = m0 + g * 32 + side * 16;
            float* o = p.out + mb * p.out_ld + n_own;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = __uint_as_float(side == 0 ? r0[j] : r1[j]) + take[j * 32 + lane];
            if (mb + 16 <= p.store_rows) {
#pragma unroll
              for (int j = 0; j < 16; j++) { *o = v[j]; o += p.out_ld; }
            } else {
              for (int j = 0; j < 16; j++) if (mb + j < p.store_rows) o[(long long)j * p.out_ld] = v[j];
            }
            if (stats) {
#pragma unroll
              for (int j = 0; j < 16; j++) { const float u = ((vm >> j) & 1u) ? v[j] : 0.f; ts1 += u; ts2 = fmaf(u, u, ts2); }
            }
          }
          epi_bar();                                                   // the exchange buffer is rewritten by the next group
        }
      }
      d1 += (double)ts1; d2 += (double)ts2;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
    if (stats && n_own < p.cout && warp_live && (d1 != 0.0 || d2 != 0.0)) {
      atomicAdd(p.stat_sum + n_own, d1);
      atomicAdd(p.stat_sq + n_own, d2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int g_bandt_launches = 0;
int conv_bandt_launch_count() { return g_bandt_launches; }

// returns SSP_OK, or 1 when the layer is not eligible (caller falls back to conv_band / the per-tap kernels)
int conv_gemm_bandt(const void* a_hi, const void* a_lo, long long a_rows, int a_ld, int cin,
                    const void* b_hi, const void* b_lo, int b_rows, int b_ld, int a_fmt, int b_fmt,
                    int N, int H, int W, int taps, int cout, float* out, int out_ld, long long out_rows,
                    int epi, const float* bias, double* stat_sum, double* stat_sq, cudaStream_t stream) {
  (void)bias;
  if ((taps != 9 && taps != 1) || epi == EPI_BIAS || epi == EPI_BNACT) return 1;
  if (!a_hi || !b_hi || !out || (a_ld % 8) || (b_ld % 8)) return fail_msg(SSP_ERR_ARG, "conv_gemm_bandt: bad argument");
  ConvBandTParams p;
  Geom g{N, H, W};
  p.stacked = (a_lo && b_lo) ? 1 : 0;
  if (p.stacked ? cout > 64 : cout > 128) return 1;
  if (epi == EPI_F16 && p.stacked) return 1;
  if (a_fmt != b_fmt) return 1;
  const int planes = p.stacked ? 2 : 1;
  p.taps = taps; p.nk = taps == 9 ? 3 : 1;
  p.kc_per_tap = (cin + 63) / 64; p.cin = cin; p.cout = cout;
  // 32-channel 3x3 layer with a row pitch of exactly 32 channels: overlapping-row tensor map (tools/probes/tmap_overlap_probe.cu) -- a band row
  // is [pixel | pixel + 1], dense 128 B instead of 64 B + zero fill, and the taps (kh, 0), (kh, 1) share one 64-wide K chunk: six resident
  // weight tiles instead of nine, which is what makes room for a third ring stage (block 2 forward: 477 us at 42 % tensor-active with two)
  static const int ovl_on = []() { const char* e = getenv("SSP_BANDT_OVL"); return e ? atoi(e) : 1; }();
  p.ovl = (ovl_on && taps == 9 && cin == 32 && a_ld == 32) ? 1 : 0;
  const int mrows = p.stacked ? 128 : ((cout + 7) / 8) * 8;
  p.w_tile_bytes = mrows * 128;
  long long res = (long long)(p.ovl ? 6 : taps * p.kc_per_tap) * p.w_tile_bytes + (128 - mrows) * 128;      // every MMA reads 128 rows from its tile's start
  res = (res + 1023) / 1024 * 1024;
  const int fixed = kXchBytes + 64 + (2 * kMaxStagesT + 5) * 8 + 16 + 1024;
  int np = 0, stages = 0;
  for (int cand = 256; cand >= 128; cand -= 128) {
    const int band_rows = taps == 9 ? cand + 8 : cand;
    const long long sb = (long long)planes * band_rows * 128;
    const long long st = (227 * 1024 - fixed - res) / sb;
    if (res < 200 * 1024 && st >= 2) { np = cand; stages = (int)(st > kMaxStagesT ? kMaxStagesT : st); break; }
  }
  if (!np) return 1;
  p.np = np; p.stages = stages; p.res_bytes = (int)res;
  if (taps == 9) { p.rows0 = 136; p.rows1 = np == 256 ? 128 : 0; } else { p.rows0 = np; p.rows1 = 0; }
  p.plane_bytes = (p.rows0 + p.rows1) * 128;
  p.stage_bytes = planes * p.plane_bytes;
  p.m_rows = g.m_rows(); p.store_rows = out_rows;
  p.m_tiles = (int)((p.m_rows + np - 1) / np);
  p.Wp = g.Wp(); p.HpWp = g.HpWp();
  p.idesc = umma_idesc_f16(b_fmt, a_fmt, 0, 0, np);            // A operand = weights, B operand = activations
  p.out = out; p.out_ld = out_ld; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.epi = epi;
  if (epi == EPI_STATS && (!stat_sum || !stat_sq)) return fail_msg(SSP_ERR_ARG, "conv_gemm_bandt: statistics buffers missing");
  int rc = 0;
  for (int pl = 0; pl < planes; pl++) {
    const void* xa = pl ? a_lo : a_hi; const void* wa = pl ? b_lo : b_hi;
    const uint64_t xin = p.ovl ? 64 : (uint64_t)cin, xrows = p.ovl ? (uint64_t)a_rows - 1 : (uint64_t)a_rows;
    rc |= tmap_2d_16bit(&p.tmX[pl][0], xa, xin, xrows, (uint64_t)a_ld, 64, p.rows0, a_fmt == FMT_BF16);
    if (p.rows1) rc |= tmap_2d_16bit(&p.tmX[pl][1], xa, xin, xrows, (uint64_t)a_ld, 64, p.rows1, a_fmt == FMT_BF16);
    rc |= tmap_2d_16bit(&p.tmW[pl], wa, (uint64_t)taps * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, p.stacked ? 64 : mrows, b_fmt == FMT_BF16);
  }
  if (rc) return fail_msg(SSP_ERR_DRIVER, "conv_gemm_bandt: cuTensorMapEncodeTiled failed");
  static int sms = 0, configured = 0;
  if (!sms) sms = ssp_sm_count();
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_bandt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  conv_bandt_kernel<<<grid, kThreadsT, p.res_bytes + stages * p.stage_bytes + fixed, stream>>>(p);
  SSP_CHECK_LAUNCH();
  g_bandt_launches++;
  return SSP_OK;
}

}  // namespace ssp
