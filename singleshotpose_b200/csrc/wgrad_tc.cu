// Weight-gradient GEMM on tcgen05:   dW[co, tap, ci] += sum_m dY[m, co] * X[m + shift(tap), ci]
//
// The contraction runs over pixels (rows of the padded-flat matrices), so BOTH operands are "MN-major"
// for the tensor core (channels contiguous, K = rows): the TMA boxes [64 rows][64 channels] land in
// shared memory exactly in the canonical MN-major SWIZZLE_128B layout, no transposed copies needed.
// dY has zero pad rows, X has zero pad rows, so the shifted product never picks up wrap-around terms.
// Work item = (co tile of 128, ci tile of BN, GROUP of taps, K split): the dY tile is loaded once per k-block and
// multiplied against the X tiles of all taps of the group, one TMEM accumulator (BN columns) per tap (up to
// 512/BN taps) -- the kernel is L2->SMEM bandwidth bound, sharing dY across taps is what matters.  Partial sums
// are reduced with fp32 RED atomics into dW, which the caller zeroes once per step.  Output layout [co][tap][ci] is the layout the
// master weights are kept in (engine.py), i.e. the gradient of nn.Conv2d.weight seen through a permuted view.
// Round 2: the X boxes of consecutive taps sit at a constant distance in shared memory and their accumulators in adjacent TMEM
// columns, which is exactly the MN-major operand's "next 64-element N group at LBO" rule -- so up to 256 / 64 taps go into ONE
// tcgen05.mma of N = 256 instead of one N = 64 instruction per tap (an MMA with N <= 64 costs ~95 clk whatever its size: the narrow
// layers were issue-bound, block 2's weight gradient ran 545 us at batch 64).  Narrow layers therefore use a 64-column accumulator
// stride per tap (cin <= 32 leaves the upper half of each group zero).
// Replaces the conv weight-gradient autograd computes for reference train.py:103.
#include "ssp_common.cuh"
#include "tmap.cuh"
#include <stdlib.h>

namespace ssp {

struct WgradTcParams {
  CUtensorMap tmDy;     // [rows][cout]   box {64, 64}
  CUtensorMap tmX;      // [rows][cin]    box {64, 64}
  long long m_rows;
  int co_tiles, ci_tiles, taps, splits;
  int groups, taps_per_group, nbuf;   // tap groups per (co, ci) tile; accumulator buffers in TMEM (2 if they fit twice)
  int kblocks_total;    // ceil(m_rows / 64)
  int shifts[9];
  int cout, cin, bn;
  int acc_stride;       // TMEM columns per tap accumulator (>= bn; 64 for bn = 32 so that taps merge into one N-wide MMA)
  int merge;            // taps per MMA (1 = one instruction per tap)
  int xbox0;            // index of the first X box inside a stage: 2 (two dY boxes), or 1 when cout <= 64 (the second dY box would be all zero fill)
  int ovl, real_taps;   // ovl: 32-channel layer through the overlapping-row tensor map -- every X box holds TWO horizontally adjacent taps (see the launcher)
  uint32_t idesc;       // N field left zero: filled per instruction
  int stages, stage_bytes;
  float* dw;
  int dw_ld, cin_store;  // row pitch of dW per (co, tap) and number of real input channels
  float scale;          // applied to the partial sums before accumulation (loss-scale undo)
};

static constexpr int kBox = 64 * 128;   // 64 rows x 64 ch x 2 B
static constexpr int kMaxStagesW = 8;
static constexpr int kThreadsW = 256;

__global__ void __launch_bounds__(kThreadsW, 1) wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(smem + (size_t)p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + kMaxStagesW;
  uint64_t* tfull_bar = empty_bar + kMaxStagesW;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_ptr = (uint32_t*)(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.co_tiles * p.ci_tiles * p.groups * p.splits;
  const int nb = p.bn <= 64 ? 1 : p.bn / 64;        // X boxes per tap and stage
  const int kb_per_split = (p.kblocks_total + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&p.tmDy); tma_prefetch_desc(&p.tmX); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // item -> (split, tap, ci tile, co tile); co fastest so that concurrently running CTAs share X tiles in L2
  auto decode = [&](int it, int& co_t, int& ci_t, int& tap0, int& ntap, int& kb0, int& kb1) {
    co_t = it % p.co_tiles; it /= p.co_tiles;
    ci_t = it % p.ci_tiles; it /= p.ci_tiles;
    const int grp = it % p.groups; it /= p.groups;
    tap0 = grp * p.taps_per_group;
    ntap = p.taps - tap0 < p.taps_per_group ? p.taps - tap0 : p.taps_per_group;
    kb0 = it * kb_per_split;
    kb1 = kb0 + kb_per_split; if (kb1 > p.kblocks_total) kb1 = p.kblocks_total;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        int co_t, ci_t, tap0, ntap, kb0, kb1; decode(it, co_t, ci_t, tap0, ntap, kb0, kb1);
        const bool two_dy = co_t * 128 + 64 < p.cout;
        const uint32_t tx = (uint32_t)((two_dy ? 2 : 1) + ntap * nb) * kBox;
        for (int kb = kb0; kb < kb1; kb++) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + (size_t)stage * p.stage_bytes;
          mbar_expect_tx(&full_bar[stage], tx);
          const int row = kb * 64;
          tma_load_2d(s, &p.tmDy, &full_bar[stage], co_t * 128, row);
          if (two_dy) tma_load_2d(s + kBox, &p.tmDy, &full_bar[stage], co_t * 128 + 64, row);     // else: TMEM lanes 64-127 accumulate stale smem, never read
          for (int t = 0; t < ntap; t++)
            for (int j = 0; j < nb; j++)
              tma_load_2d(s + (p.xbox0 + t * nb + j) * kBox, &p.tmX, &full_bar[stage], ci_t * p.bn + j * 64, row + p.shifts[tap0 + t]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: the whole warp runs the loops converged and ONE elected lane issues (operands then live in uniform registers;
    // from an `if (lane == 0)` region every tcgen05.mma is wrapped in an ELECT / retry loop -- round 2, ncu: the narrow layers'
    // weight gradients spent 70 % of the issuing warp's time in that scalar code at 20-33 % tensor-active)
    int stage = 0; uint32_t phase = 0; int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, n++) {
      int co_t, ci_t, tap0, ntap, kb0, kb1; decode(it, co_t, ci_t, tap0, ntap, kb0, kb1);
      const int buf = n % p.nbuf, use = n / p.nbuf;
      mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.taps_per_group * p.acc_stride);
      uint32_t acc = 0;
      for (int kb = kb0; kb < kb1; kb++) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t s = smem_u32(smem + (size_t)stage * p.stage_bytes);
        if (elect_one_sync()) {
          for (int t = 0; t < ntap; t += p.merge) {
            const int g = ntap - t < p.merge ? ntap - t : p.merge;                       // taps in this instruction
            const uint32_t idesc = p.idesc | ((uint32_t)((p.merge > 1 ? g * p.acc_stride : p.bn) >> 3) << 17);
#pragma unroll
            for (int k = 0; k < 4; k++) {   // 16 K rows per MMA = 2 swizzle groups of 8 rows = 2048 B
              const uint64_t da = umma_desc_sw128(s + k * 2048, kBox, 1024);
              const uint64_t db = umma_desc_sw128(s + (p.xbox0 + t * nb) * kBox + k * 2048, kBox, 1024);
              umma_f16(d_tmem + (uint32_t)(t * p.acc_stride), da, db, idesc, (k == 0) ? acc : 1u);
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
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, n++) {
      int co_t, ci_t, tap0, ntap, kb0, kb1; decode(it, co_t, ci_t, tap0, ntap, kb0, kb1);
      const int buf = n % p.nbuf, use = n / p.nbuf;
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const int co = co_t * 128 + q * 32 + lane;
      if (p.ovl) {
        // pseudo-tap t = kernel row t / 2, box t % 2: columns [0,32) are tap kw = 2 (t % 2), columns [32,64) tap kw + 1 (kw = 3 does not exist)
        for (int t = 0; t < ntap; t++)
          for (int hf = 0; hf < 2; hf++) {
            const int kw = 2 * ((tap0 + t) & 1) + hf;
            if (kw > 2) continue;
            const int tap = ((tap0 + t) >> 1) * 3 + kw;
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.taps_per_group + t) * p.acc_stride + hf * 32), r);
            tmem_ld_wait();
            float* drow = p.dw + ((long long)co * p.real_taps + tap) * p.dw_ld;
            if (co < p.cout && kb1 > kb0) {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (j < p.cin_store) atomicAdd(drow + j, __uint_as_float(r[j]) * p.scale);
            }
          }
      } else
      for (int t = 0; t < ntap; t++) {
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.taps_per_group + t) * p.acc_stride);
        float* drow = p.dw + ((long long)co * p.taps + tap0 + t) * p.dw_ld;
        for (int ch = 0; ch < p.bn / 32; ch++) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + ch * 32, r);
          tmem_ld_wait();
          const int c0 = ci_t * p.bn + ch * 32;
          if (co < p.cout && kb1 > kb0) {
#pragma unroll
            for (int j = 0; j < 32; j++)
              if (c0 + j < p.cin_store) atomicAdd(drow + c0 + j, __uint_as_float(r[j]) * p.scale);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

static int g_num_sms_w = 0;

int wgrad_gemm_tc(const void* dy, long long dy_rows, int dy_ld, int cout, int dy_fmt,
                  const void* x, long long x_rows, int x_ld, int cin, int x_fmt,
                  int N, int H, int W, int taps, float* dw, int dw_ld, int cin_store, float scale, cudaStream_t stream) {
  if (!dy || !x || !dw || (taps != 1 && taps != 9)) return fail_msg(SSP_ERR_ARG, "wgrad_gemm_tc: bad argument");
  if ((dy_ld % 8) || (x_ld % 8)) return fail_msg(SSP_ERR_ARG, "wgrad_gemm_tc: leading dimensions must be multiples of 8");
  if (!g_num_sms_w) {
    g_num_sms_w = ssp_sm_count();
  }
  WgradTcParams p;
  Geom g{N, H, W};
  int bn = ((cin + 63) / 64) * 64;
  if (bn > 256) bn = 256;
  if (bn > 128 && bn < 256) bn = 256;
  if (cin <= 32) bn = 32;             // N = 32 MMAs on the first half of the 64-channel box (rest is TMA zero fill)
  p.bn = bn;
  const int nb = bn <= 64 ? 1 : bn / 64;
  // sharing dY across taps pays for narrow tiles (L2-bound, deep pipeline still fits); for BN >= 128 the shallower
  // smem ring and the single TMEM buffer cost more than the saved traffic (measured), so one tap per item there
  static const int merge_on = []() { const char* e = getenv("SSP_WGRAD_MERGE"); return e ? atoi(e) : 1; }();
  // 32-channel 3x3 layer stored with a row pitch of exactly 32 channels (block 2 of yolo-pose.cfg): a tensor map whose rows OVERLAP
  // (inner extent 64 elements, row pitch 64 B; legal and delivered correctly, tools/probes/tmap_overlap_probe.cu) makes row r of a box
  // the channels of pixel r followed by those of pixel r + 1, i.e. the operands of TWO horizontally adjacent taps.  Six boxes instead of
  // nine half-empty ones, one accumulator group (384 TMEM columns) instead of two, so dY is staged once per k-block: 56 KB instead of
  // 88-104 KB requested from L2 per 64 pixel rows (the kernel was fill-bound: 572 us at batch 64, 33 % tensor-active).
  static const int ovl_on = []() { const char* e = getenv("SSP_WGRAD_OVL"); return e ? atoi(e) : 1; }();
  const bool ovl = merge_on && ovl_on && taps == 9 && cin == 32 && x_ld == 32 && cin_store <= 32;
  p.ovl = ovl ? 1 : 0; p.real_taps = taps;
  const int ptaps = ovl ? 6 : taps;                 // ovl: six pseudo-taps = (kernel row, box 0 | 1), 64 accumulator columns each
  p.acc_stride = (merge_on && bn < 64) ? 64 : bn;
  p.merge = (merge_on && taps > 1 && p.acc_stride <= 128) ? 256 / p.acc_stride : 1;
  int tmax = (bn <= 64) ? 512 / p.acc_stride : (p.merge > 1 ? p.merge : 1); if (tmax > ptaps) tmax = ptaps;
  p.groups = (ptaps + tmax - 1) / tmax;
  p.taps_per_group = (ptaps + p.groups - 1) / p.groups;
  p.nbuf = (2 * p.taps_per_group * p.acc_stride <= 512) ? 2 : 1;
  p.m_rows = g.m_rows();
  p.kblocks_total = (int)((p.m_rows + 63) / 64);
  p.co_tiles = (cout + 127) / 128;
  p.ci_tiles = (cin + bn - 1) / bn;
  p.taps = ptaps;
  for (int t = 0; t < 9; t++) p.shifts[t] = (taps == 9) ? ((t / 3) - 1) * g.Wp() + ((t % 3) - 1) : 0;
  if (ovl) for (int t = 0; t < 6; t++) p.shifts[t] = ((t / 2) - 1) * g.Wp() - 1 + 2 * (t % 2);     // box 0: taps kw = 0, 1; box 1: kw = 2 (and a non-existent kw = 3)
  p.cout = cout; p.cin = cin;
  const int base_items = p.co_tiles * p.ci_tiles * p.groups;
  // split K until there are ~2 waves of work items, keeping at least 32 k-blocks (2048 rows) per item
  int splits = (2 * g_num_sms_w + base_items - 1) / base_items;
  int max_splits = p.kblocks_total / 32; if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.idesc = umma_idesc_f16(dy_fmt, x_fmt, 1, 1, 0);
  p.xbox0 = cout > 64 ? 2 : 1;          // cout <= 64: TMEM lanes 64-127 accumulate whatever follows in the stage (the first X box); never read
  p.stage_bytes = (p.xbox0 + p.taps_per_group * nb) * kBox;
  const int fixed = (2 * kMaxStagesW + 4) * 8 + 16 + 1024;
  int stages = (227 * 1024 - fixed) / p.stage_bytes;
  if (stages > kMaxStagesW) stages = kMaxStagesW;
  p.stages = stages;
  p.dw = dw; p.dw_ld = dw_ld; p.cin_store = cin_store; p.scale = scale;
  int rc = 0;
  rc |= tmap_2d_16bit(&p.tmDy, dy, (uint64_t)cout, (uint64_t)dy_rows, (uint64_t)dy_ld, 64, 64, dy_fmt == FMT_BF16);
  if (ovl) rc |= tmap_2d_16bit(&p.tmX, x, 64, (uint64_t)x_rows - 1, 32, 64, 64, x_fmt == FMT_BF16);      // overlapping rows: [pixel r | pixel r + 1]
  else rc |= tmap_2d_16bit(&p.tmX, x, (uint64_t)cin, (uint64_t)x_rows, (uint64_t)x_ld, 64, 64, x_fmt == FMT_BF16);
  if (rc) return fail_msg(SSP_ERR_DRIVER, "wgrad_gemm_tc: cuTensorMapEncodeTiled failed");
  static int configured = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  const int items = base_items * splits;
  const int grid = items < g_num_sms_w ? items : g_num_sms_w;
  wgrad_tc_kernel<<<grid, kThreadsW, stages * p.stage_bytes + fixed, stream>>>(p);
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
