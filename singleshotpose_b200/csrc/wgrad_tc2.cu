// CTA-pair version of the weight-gradient GEMM (default for layers with cout, cin multiples of 256; round 2: 4.0 -> 3.5 ms/step):
//     dW[co, tap, ci] += sum_m dY[m, co] * X[m + shift(tap), ci]
// Why: wgrad_tc.cu is bound by the L2->SMEM fill rate, not by the tensor pipe (profiles/r01_experiments.md: 43 % tensor).
// Per 64-row k-block a 1-CTA item stages dY 128 ch (16 KB) + X 256 ch (32 KB) = 48 KB for 128x256x64 MACs = 87 FLOP/B.
// With tcgen05.mma.cta_group::2 the pair computes a 256(co) x 256(ci) tile: each CTA stages ITS 128 output channels of dY
// (16 KB) and ITS half of the 256 input channels of X (16 KB) = 32 KB for the same MACs per CTA = 131 FLOP/B (x1.5).
// Both operands stay MN-major (channels contiguous, K = pixel rows), boxes [64 rows][64 ch] in SWIZZLE_128B as in wgrad_tc.cu.
// One tap per work item, double-buffered 2 x 256-column accumulators; pair mechanics (leader-credited TMA, multicast commits,
// remote tempty arrives) are those of conv_tc2.cu.  Eligible layers: cout % 256 == 0 and cin % 256 == 0 (all 13x13 and 26x26
// 3x3 / 1x1 layers of yolo-pose.cfg); anything else returns 1 and the caller falls back to wgrad_tc.cu.
// Clusters of TWO pairs (CSZ = 4, cout % 512 == 0): the pairs take neighbouring 256-channel blocks of dY and the SAME X tile, every CTA
// fetches one of its two X boxes and multicasts it to its counterpart in the other pair -- 24 KB instead of 32 KB requested from
// L2 per CTA and k-block (the fill rate is bounded by what a CTA requests, tools/probes/mc_probe2.cu; see conv_tc2.cu).
#include "ssp_common.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace ssp {

struct WgradTc2Params {
  CUtensorMap tmDy;     // [rows][cout]   box {64, 64}
  CUtensorMap tmX;      // [rows][cin]    box {64, 64}
  long long m_rows;
  int co_pairs, ci_tiles, taps, splits;
  int kblocks_total;    // ceil(m_rows / 64)
  int shifts[9];
  int cout, cin;
  uint32_t idesc;
  int stages;
  float* dw;
  int dw_ld, cin_store;
  float scale;
};

namespace {
constexpr int kBox2 = 64 * 128;          // 64 rows x 64 ch x 2 B
constexpr int kStage2 = 4 * kBox2;       // dY: 2 boxes (this CTA's 128 co), X: 2 boxes (this CTA's 128 ci)
constexpr int kMaxStages2 = 8;
constexpr int kThreads2 = 256;

__device__ __forceinline__ uint32_t w2_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void w2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void w2_tma_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void w2_umma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void w2_tma_pair_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void w2_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void w2_arrive_rank(uint64_t* bar, uint32_t target) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(smem_u32(bar)), "r"(target) : "memory");
}
__device__ __forceinline__ void w2_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void w2_tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void w2_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
}  // namespace

template <int CSZ>
__global__ void __launch_bounds__(kThreads2, 1) __cluster_dims__(CSZ, 1, 1) wgrad_tc2_kernel(const __grid_constant__ WgradTc2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * kStage2);
  uint64_t* empty_bar = full_bar + kMaxStages2;
  uint64_t* tfull_bar = empty_bar + kMaxStages2;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NP = CSZ / 2;                                     // CTA pairs per cluster
  const uint32_t crank = w2_ctarank();
  const uint32_t rank = crank & 1, pr = crank >> 1;               // rank inside the pair (0 = leader), pair inside the cluster
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pr)), all_mask = (uint16_t)((1u << CSZ) - 1);
  const int cid = blockIdx.x / CSZ, ncl = gridDim.x / CSZ;
  const int items = (p.co_pairs / NP) * p.ci_tiles * p.taps * p.splits;      // work items of a cluster
  const int kb_per_split = (p.kblocks_total + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmDy); tma_prefetch_desc(&p.tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NP); }   // free when every pair reading the slot is done
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }   // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 2) { w2_tmem_alloc(tmem_ptr, 512); w2_tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  w2_cluster_sync();              // peer barriers initialised, pair TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // item -> (split, tap, ci tile, co pair); co fastest so that concurrently running pairs share X tiles in L2
  auto decode = [&](int it, int& co_p, int& ci_t, int& tap, int& kb0, int& kb1) {
    co_p = (it % (p.co_pairs / NP)) * NP + (int)pr; it /= (p.co_pairs / NP);
    ci_t = it % p.ci_tiles; it /= p.ci_tiles;
    tap = it % p.taps; it /= p.taps;
    kb0 = it * kb_per_split;
    kb1 = kb0 + kb_per_split; if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
  };

  if (warp == 0) {
    if (lane == 0) {            // TMA producer, one per CTA: this CTA's halves of both operands
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = 2u * (uint32_t)kStage2;                     // both CTAs' bytes land on the leader's barrier
      for (int it = cid; it < items; it += ncl) {
        int co_p, ci_t, tap, kb0, kb1; decode(it, co_p, ci_t, tap, kb0, kb1);
        const int co0 = co_p * 256 + (int)rank * 128, ci0 = ci_t * 256 + (int)rank * 128;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + (size_t)stage * kStage2;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
          const int row = kb * 64;
          w2_tma_pair(s, &p.tmDy, &full_bar[stage], co0, row);
          w2_tma_pair(s + kBox2, &p.tmDy, &full_bar[stage], co0 + 64, row);
          if (CSZ == 2) {
            w2_tma_pair(s + 2 * kBox2, &p.tmX, &full_bar[stage], ci0, row + p.shifts[tap]);
            w2_tma_pair(s + 3 * kBox2, &p.tmX, &full_bar[stage], ci0 + 64, row + p.shifts[tap]);
          } else {     // one of this CTA's two X boxes, delivered to the same half of both pairs
            w2_tma_pair_mc(s + (2 + (int)pr) * kBox2, &p.tmX, &full_bar[stage], ci0 + (int)pr * 64, row + p.shifts[tap],
                           (uint16_t)((1u << rank) | (1u << (rank + 2))));
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {   // MMA issuer: leader CTA only
      int stage = 0; uint32_t phase = 0; int n = 0;
      for (int it = cid; it < items; it += ncl, n++) {
        int co_p, ci_t, tap, kb0, kb1; decode(it, co_p, ci_t, tap, kb0, kb1);
        const int buf = n & 1;
        mbar_wait(&tempty_bar[buf], ((n >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
        uint32_t acc = 0;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t s = smem_u32(smem + (size_t)stage * kStage2);
#pragma unroll
          for (int k = 0; k < 4; k++) {   // 16 K rows per MMA = 2 swizzle groups of 8 rows = 2048 B
            const uint64_t da = umma_desc_sw128(s + k * 2048, kBox2, 1024);
            const uint64_t db = umma_desc_sw128(s + 2 * kBox2 + k * 2048, kBox2, 1024);
            w2_umma_pair(d_tmem, da, db, p.idesc, (k == 0) ? acc : 1u);
          }
          acc = 1;
          w2_commit_pair(&empty_bar[stage], all_mask);      // this pair is done with the slot (every producer of the cluster hears it)
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        w2_commit_pair(&tfull_bar[buf], pair_mask);       // accumulator complete -> both epilogues of this pair
      }
    }
  } else if (warp >= 4) {
    const int q = warp - 4;
    int n = 0;
    for (int it = cid; it < items; it += ncl, n++) {
      int co_p, ci_t, tap, kb0, kb1; decode(it, co_p, ci_t, tap, kb0, kb1);
      const int buf = n & 1;
      mbar_wait(&tfull_bar[buf], (n >> 1) & 1);
      tc_fence_after();
      const int co = co_p * 256 + (int)rank * 128 + q * 32 + lane;      // this CTA's TMEM lanes = its 128 output channels
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 256);
      float* drow = p.dw + ((long long)co * p.taps + tap) * p.dw_ld;
      for (int ch = 0; ch < 8; ch++) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + ch * 32, r);
        tmem_ld_wait();
        const int c0 = ci_t * 256 + ch * 32;
        if (co < p.cout && kb1 > kb0) {
#pragma unroll
          for (int j = 0; j < 32; j++)
            if (c0 + j < p.cin_store) atomicAdd(drow + c0 + j, __uint_as_float(r[j]) * p.scale);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) w2_arrive_rank(&tempty_bar[buf], 2 * pr);
    }
  }
  tc_fence_before();
  __syncthreads();
  w2_cluster_sync();              // the peer may still read this CTA's smem / arrive on its barriers
  if (warp == 2) { tc_fence_after(); w2_tmem_dealloc(tmem_base, 512); }
}

static int g_num_sms_w2 = 0;

// returns 1 when the layer is not eligible (caller falls back to wgrad_gemm_tc)
int wgrad_gemm_tc2(const void* dy, long long dy_rows, int dy_ld, int cout, int dy_fmt,
                   const void* x, long long x_rows, int x_ld, int cin, int x_fmt,
                   int N, int H, int W, int taps, float* dw, int dw_ld, int cin_store, float scale, cudaStream_t stream) {
  if (!dy || !x || !dw || (taps != 1 && taps != 9)) return fail_msg(SSP_ERR_ARG, "wgrad_gemm_tc2: bad argument");
  if ((dy_ld % 8) || (x_ld % 8)) return fail_msg(SSP_ERR_ARG, "wgrad_gemm_tc2: leading dimensions must be multiples of 8");
  if ((cout % 256) || (cin % 256)) return 1;
  if (!g_num_sms_w2) {
    g_num_sms_w2 = ssp_sm_count();
  }
  WgradTc2Params p;
  Geom g{N, H, W};
  p.m_rows = g.m_rows();
  p.kblocks_total = (int)((p.m_rows + 63) / 64);
  p.co_pairs = cout / 256;
  p.ci_tiles = cin / 256;
  p.taps = taps;
  for (int t = 0; t < 9; t++) p.shifts[t] = (taps == 9) ? ((t / 3) - 1) * g.Wp() + ((t % 3) - 1) : 0;
  p.cout = cout; p.cin = cin;
  static const int csz_env = []() { const char* e = getenv("SSP_TC2_CLUSTER"); return e ? atoi(e) : 2; }();      // 4 = opt-in (measured slower, see the kernel comment)
  const int csz = (csz_env == 4 && (p.co_pairs % 2) == 0) ? 4 : 2;
  int pairs = g_num_sms_w2 / csz;                                         // clusters that fit the chip ...
  if (csz == 4) {                                                         // ... and are co-resident (see conv_tc2.cu: GPCs are not multiples of 4 SMs)
    static int mc = 0;
    if (!mc) {
      cudaFuncSetAttribute(wgrad_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(4 * (g_num_sms_w2 / 4)); cfg.blockDim = dim3(kThreads2); cfg.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, wgrad_tc2_kernel<4>, &cfg);
      mc = (e == cudaSuccess && n > 0) ? n : g_num_sms_w2 / 8;
    }
    if (pairs > mc) pairs = mc;
  }
  const int base_items = (p.co_pairs / (csz / 2)) * p.ci_tiles * taps;    // work items of a cluster
  // split K only when one pass leaves most pairs idle or the last wave mostly empty; keep >= 32 k-blocks per item
  int max_splits = p.kblocks_total / 32; if (max_splits < 1) max_splits = 1;
  int best = 1; double best_eff = 0.0;
  for (int s = 1; s <= max_splits && s <= 8; s++) {
    const int items = base_items * s;
    const int waves = (items + pairs - 1) / pairs;
    const double eff = (double)items / ((double)waves * pairs) - 0.02 * (s - 1);     // small penalty per extra RED pass
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
  }
  p.splits = best;
  p.idesc = (umma_idesc_f16(dy_fmt, x_fmt, 1, 1, 256) & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);   // M = 256 across the pair
  const int fixed = (2 * kMaxStages2 + 4) * 8 + 16 + 1024;
  int stages = (227 * 1024 - fixed) / kStage2;
  if (stages > kMaxStages2) stages = kMaxStages2;
  p.stages = stages;
  p.dw = dw; p.dw_ld = dw_ld; p.cin_store = cin_store; p.scale = scale;
  int rc = 0;
  rc |= tmap_2d_16bit(&p.tmDy, dy, (uint64_t)cout, (uint64_t)dy_rows, (uint64_t)dy_ld, 64, 64, dy_fmt == FMT_BF16);
  rc |= tmap_2d_16bit(&p.tmX, x, (uint64_t)cin, (uint64_t)x_rows, (uint64_t)x_ld, 64, 64, x_fmt == FMT_BF16);
  if (rc) return fail_msg(SSP_ERR_DRIVER, "wgrad_gemm_tc2: cuTensorMapEncodeTiled failed");
  static int configured = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(wgrad_tc2_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  const int items = base_items * p.splits;
  int use = items < pairs ? items : pairs;
  if (csz == 4) wgrad_tc2_kernel<4><<<4 * use, kThreads2, stages * kStage2 + fixed, stream>>>(p);
  else wgrad_tc2_kernel<2><<<2 * use, kThreads2, stages * kStage2 + fixed, stream>>>(p);     // __cluster_dims__(2,1,1): CTA pairs on one TPC
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
