// Implicit-GEMM convolution, CTA-PAIR version: tcgen05.mma.cta_group::2 (256 x BN x 16 per instruction).
// Same math, operands and epilogue as conv_tc.cu; two CTAs of a cluster (one TPC) share every MMA: each CTA stages
// its own 128 activation rows but only HALF of the weight tile (BN/2 rows), the tensor cores read the other half
// from the peer's shared memory.  The 1-CTA kernel is L2->SM bandwidth bound on the 13x13 layers (measured ~5.3 kB/clk
// chip-wide); halving the weight bytes per SM is what lifts it.  Protocol (CUTLASS 2-SM pipeline restated):
//   full[s]   lives in the leader (rank 0): count 1, expect_tx = both CTAs' bytes; both producers' TMA complete_tx on it
//   empty[s]  one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[b]  one per CTA (multicast commit);  tempty[b] in the leader, count 8 = 4 epilogue warps x 2 CTAs (remote arrive)
// ---- text of the 1-CTA kernel's header follows ----
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
#include <stdlib.h>

namespace ssp {

struct ConvTc2Params {
  CUtensorMap tmA[2];
  CUtensorMap tmB[2];     // box rows = bn / 2 (clusters of one pair) or bn / 4 (clusters of two pairs: the weight tile is multicast)
  CUtensorMap tmAband[2]; // band mode: box {64, 136 rows}
  int band;               // 3x3 layer with <= 64 input channels, split operands: ONE 136-row activation band per kernel row serves the three horizontal taps
  long long m_rows;       // rows of the output matrix that exist (N*(H+1)*(W+1))
  long long store_rows;   // rows that may be written (allocation bound)
  int m_tiles, n_tiles;
  int ksplit;             // > 1: every tile's K loop is cut into ksplit work items whose partial sums meet in `out` through fp32 atomics (EPI_F32 only)
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

namespace {
static constexpr int kABytes = 128 * 128;     // 128 rows x 64 x 2 B
static constexpr int kMaxStages = 8;
static constexpr int kAccCols = 1024;         // per-CTA statistics accumulators (channels)
static constexpr int kThreads = 256;
static constexpr int kBandBytes2 = 136 * 128;  // band mode: 136 activation rows (17 swizzle atoms)
}

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA tile load whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit 24 cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
// the same, written into this CTA and the CTAs of `mask` at the same offset; every destination's pair leader is credited
__device__ __forceinline__ void tma_load_2d_pair_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all previously issued MMAs arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t target) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(smem_u32(bar)), "r"(target) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// MODE 0: training / generic epilogues.  MODE 1: the inference epilogue (EPI_BNACT).  Separate instantiations keep MODE 0's
// register budget.  (A MODE 2 that folded the BN-backward reduction of the producer into the data-gradient epilogue was measured
// in round 2: 18.29 vs 18.13 ms/step, the extra Y reads sit on the dgrad critical path -- removed.)
// CSZ = CTAs per cluster: 2 = one CTA pair; 4 = two pairs on consecutive M tiles and the SAME N tile -- every CTA fetches half of its
// half of the weight tile and multicasts it to its counterpart in the other pair.  Round 2, tools/probes/mc_probe2.cu: the operand
// fill rate is bounded by what a CTA REQUESTS from L2 (64 B/clk/SM unicast with a deep queue, ~36-40 with the 224 KB a GEMM ring
// can keep in flight) while an SM can take in 87-95 B/clk when the tiles arrive by multicast; a 256 x 256 single-term pair tile
// needs 64 B/clk/SM at the full MMA rate, so the data / weight gradients sat at 62 % of the tensor pipe.  With the weight tile
// shared the requested bytes per k-block drop from 32 KB to 24 KB per CTA.
// MEASURED (round 2, B200, batch 64, same box): correct (every parity test passes with SSP_TC2_CLUSTER=4) but SLOWER -- forward
// 4.61 -> 5.22 ms, data gradient 2.56 -> 2.70 ms, weight gradient (wgrad_tc2.cu) 3.54 -> 4.38 ms per step; 13x13 forward 240 -> 284 us
// at 63 % instead of 77 % tensor-active.  Two reasons seen: only 33 clusters of 4 are co-resident (cudaOccupancyMaxActiveClusters:
// 132 of the 148 SMs), and a ring slot is now released only when BOTH pairs have consumed it while a pair can start a k-block only
// when all four producers have delivered -- the pairs run in lockstep and every hiccup of one stalls the other.  Default stays
// CSZ = 2; SSP_TC2_CLUSTER=4 keeps the path reachable for the next attempt (2 x 2 clusters of 8 with both operands shared).
template <int MODE, int CSZ>
__global__ void __launch_bounds__(kThreads, 1) __cluster_dims__(CSZ, 1, 1) conv_tc2_kernel(const __grid_constant__ ConvTc2Params p) {
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
  constexpr int NP = CSZ / 2;                                     // CTA pairs per cluster
  const int total_tiles = ((p.m_tiles + NP - 1) / NP) * p.n_tiles * p.ksplit;      // work items of a CLUSTER; m_tiles counts 256-row pair tiles
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;                                // rank inside the CTA pair (0 = leader: issues the MMAs, owns full / tempty)
  const uint32_t pr = crank >> 1;                                 // pair inside the cluster
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pr)), all_mask = (uint16_t)((1u << CSZ) - 1);
  const int cid = blockIdx.x / CSZ, ncl = gridDim.x / CSZ;
  const int kblocks = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB[0]);
    if (p.n_terms == 3) { tma_prefetch_desc(&p.tmA[1]); tma_prefetch_desc(&p.tmB[1]); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NP); }    // a slot is free when EVERY pair that reads what lands in it has consumed it
    for (int b = 0; b < 2; b++) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 8); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc_pair(tmem_ptr, 512); tmem_relinquish_pair(); }
  if (p.epi == EPI_STATS)
    for (int i = threadIdx.x; i < 2 * kAccCols; i += kThreads) acc_sum[i] = 0.0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();             // peer barriers initialised, pair TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx = 2u * (uint32_t)(p.n_terms == 3 ? 2 : 1) * (uint32_t)(kABytes + p.b_bytes);   // both CTAs' bytes
      for (int w = cid; w < total_tiles; w += ncl) {
        const int t = w / p.ksplit, ks = w - t * p.ksplit;
        const int mt = (t / p.n_tiles) * NP + (int)pr, nt = t % p.n_tiles;
        const int m0 = mt * 256 + (int)rank * 128, n0 = nt * p.bn + (int)rank * (p.bn / 2);
        const int kb0 = ks * kblocks / p.ksplit, kb1 = (ks + 1) * kblocks / p.ksplit;
        const int nq = n0 + (int)pr * (p.bn / 4), qoff = (int)pr * (p.b_bytes / 2);      // CSZ == 4: this CTA's quarter of the weight tile
        const uint16_t bmask = (uint16_t)((1u << rank) | (1u << (rank + 2)));             // ... goes to the same half of both pairs
        if (p.band) {
          // per kernel row: the 136-row band [m0 + (kh-1)(W+1) - 1, +136) in both planes, then the three taps' weight tiles
          const uint32_t txb = 2u * (uint32_t)p.stage_bytes;
          for (int kh = 0; kh < 3; kh++) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], txb);
            const int arow = m0 + (kh - 1) * p.Wp - 1;
            tma_load_2d_pair(sa, &p.tmAband[0], &full_bar[stage], 0, arow);
            tma_load_2d_pair(sa + kBandBytes2, &p.tmAband[1], &full_bar[stage], 0, arow);
            for (int kw = 0; kw < 3; kw++) {
              uint8_t* sb = sa + 2 * kBandBytes2 + (size_t)kw * 2 * p.b_bytes;
              tma_load_2d_pair(sb, &p.tmB[0], &full_bar[stage], (kh * 3 + kw) * p.cin, n0);
              tma_load_2d_pair(sb + p.b_bytes, &p.tmB[1], &full_bar[stage], (kh * 3 + kw) * p.cin, n0);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        } else {
          for (int kb = kb0; kb < kb1; kb++) {
            const int tap = kb / p.kc_per_tap, kc = kb - tap * p.kc_per_tap;
            const int arow = m0 + p.shifts[tap];
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + (size_t)stage * p.stage_bytes;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], tx);
            const int kcol_b = tap * p.cin + kc * 64;
            tma_load_2d_pair(sa, &p.tmA[0], &full_bar[stage], kc * 64, arow);
            if (p.n_terms == 3) {
              tma_load_2d_pair(sa + kABytes, &p.tmA[1], &full_bar[stage], kc * 64, arow);
              if (CSZ == 2) {
                tma_load_2d_pair(sa + 2 * kABytes, &p.tmB[0], &full_bar[stage], kcol_b, n0);
                tma_load_2d_pair(sa + 2 * kABytes + p.b_bytes, &p.tmB[1], &full_bar[stage], kcol_b, n0);
              } else {
                tma_load_2d_pair_mc(sa + 2 * kABytes + qoff, &p.tmB[0], &full_bar[stage], kcol_b, nq, bmask);
                tma_load_2d_pair_mc(sa + 2 * kABytes + p.b_bytes + qoff, &p.tmB[1], &full_bar[stage], kcol_b, nq, bmask);
              }
            } else {
              if (CSZ == 2) tma_load_2d_pair(sa + kABytes, &p.tmB[0], &full_bar[stage], kcol_b, n0);
              else tma_load_2d_pair_mc(sa + kABytes + qoff, &p.tmB[0], &full_bar[stage], kcol_b, nq, bmask);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int w = cid; w < total_tiles; w += ncl, it++) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.bn);
        uint32_t acc = 0;
        const int ks = w % p.ksplit;
        const int nkb = (ks + 1) * kblocks / p.ksplit - ks * kblocks / p.ksplit;
        if (p.band) {
          const int ksteps = p.cin >= 64 ? 4 : (p.cin + 15) / 16;         // all-zero K steps (TMA zero fill) are skipped
          for (int kh = 0; kh < 3; kh++) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + (size_t)stage * p.stage_bytes);
            for (int kw = 0; kw < 3; kw++) {
              const uint32_t a_hi = sa + kw * 128, a_lo = a_hi + kBandBytes2;      // tap kw starts kw rows into the band
              const uint32_t b_hi = sa + 2 * kBandBytes2 + (uint32_t)(kw * 2 * p.b_bytes), b_lo = b_hi + p.b_bytes;
#pragma unroll
              for (int k = 0; k < 4; k++) {
                if (k < ksteps) {
                  const uint64_t dah = umma_desc_sw128(a_hi + k * 32, 16, 1024), dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
                  umma_f16_pair(d_tmem, umma_desc_sw128(a_lo + k * 32, 16, 1024), dbh, p.idesc, acc);  acc = 1;
                  umma_f16_pair(d_tmem, dah, umma_desc_sw128(b_lo + k * 32, 16, 1024), p.idesc, 1);
                  umma_f16_pair(d_tmem, dah, dbh, p.idesc, 1);
                }
              }
            }
            umma_commit_pair(&empty_bar[stage], all_mask);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        } else
        for (int kb = 0; kb < nkb; kb++) {
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
              umma_f16_pair(d_tmem, dal, dbh, p.idesc, acc);  acc = 1;   // small cross terms first
              umma_f16_pair(d_tmem, dah, dbl, p.idesc, 1);
            }
            umma_f16_pair(d_tmem, dah, dbh, p.idesc, acc);  acc = 1;
          }
          umma_commit_pair(&empty_bar[stage], all_mask);  // this pair is done with the slot (every producer of the cluster hears it)
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tfull_bar[buf], pair_mask);     // accumulator complete -> both epilogues of this pair
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (128 threads = 128 TMEM lanes)
    const int q = warp - 4;
    int it = 0;
    for (int w = cid; w < total_tiles; w += ncl, it++) {
      const int buf = it & 1;
      const int t = w / p.ksplit;
      const int mt = (t / p.n_tiles) * NP + (int)pr, nt = t % p.n_tiles;
      const long long m = (long long)mt * 256 + rank * 128 + q * 32 + lane;
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
        if constexpr (MODE == 1) {
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
        if (p.epi == EPI_F16) {                     // data gradient kept in fp16 (the BN backward reads it twice): 64 B per row chunk
          if (can_store) {
            uint16_t* o16 = reinterpret_cast<uint16_t*>(p.out) + m * p.out_ld + c0;
            if (c0 + 32 <= p.cout) {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 16; j++) pk[j] = cvt_f32_to_16(v[2 * j], FMT_F16) | ((uint32_t)cvt_f32_to_16(v[2 * j + 1], FMT_F16) << 16);
#pragma unroll
              for (int j = 0; j < 4; j++) reinterpret_cast<uint4*>(o16)[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++) if (c0 + j < p.cout) o16[j] = cvt_f32_to_16(v[j], FMT_F16);
            }
          }
        } else if (can_store && p.ksplit > 1) {            // split K: partial sums of the ksplit work items meet in the (pre-zeroed) output
          if (c0 + 32 <= p.cout) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              atomicAdd(reinterpret_cast<float4*>(orow + c0 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (c0 + j < p.cout) atomicAdd(orow + c0 + j, v[j]);
          }
        } else if (can_store) {
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
      if (lane == 0) mbar_arrive_rank(&tempty_bar[buf], 2 * pr);
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
  cluster_sync_all();             // the peer may still read this CTA's smem / arrive on its barriers
  if (warp == 2) { tc_fence_after(); tmem_dealloc_pair(tmem_base, 512); }
}

static int g_num_sms2 = 0;

int conv_gemm_tc2(const void* a_hi, const void* a_lo, long long a_rows, int a_ld, int cin,
                 const void* b_hi, const void* b_lo, int b_rows, int b_ld, int a_fmt, int b_fmt,
                 int N, int H, int W, int taps, int cout, float* out, int out_ld, long long out_rows,
                 int epi, const float* bias, double* stat_sum, double* stat_sq, cudaStream_t stream, const FusedAct* fa) {
  if (fa) {
    if (!fa->scale || !fa->shift || !fa->d_hi || !fa->d_lo || (cout % 32) || (fa->d_ld % 8) || (fa->d_c0 % 8))
      return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: fused BN+activation epilogue needs cout % 32 == 0 and 16-B aligned destination rows");
    epi = EPI_BNACT;
  }
  if (!a_hi || !b_hi || (!out && !fa) || (taps != 1 && taps != 9) || cin <= 0 || cout <= 0) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: bad argument");
  if ((a_ld % 8) || (b_ld % 8)) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: leading dimensions must be multiples of 8 elements (16 B)");
  if (epi == EPI_STATS && (cout > kAccCols || !stat_sum || !stat_sq)) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: statistics need cout <= 1024 and buffers");
  if (epi == EPI_BIAS && !bias) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: bias missing");
  if (!fa && epi != EPI_F16 && ((out_ld % 4) || ((uintptr_t)out % 16))) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: output must be 16-B aligned with ld % 4 == 0");
  if (epi == EPI_F16 && ((out_ld % 8) || ((uintptr_t)out % 16))) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: fp16 output must be 16-B aligned with ld % 8 == 0");
  if (!g_num_sms2) {
    g_num_sms2 = ssp_sm_count();
  }
  ConvTc2Params p;
  Geom g{N, H, W};
  p.n_terms = (a_lo && b_lo) ? 3 : 1;
  int bn = ((cout + 31) / 32) * 32;
  if (bn < 64) bn = 64;             // each CTA stages bn/2 >= 32 weight rows
  if (bn > 256) bn = 256;
  if (bn > 128 && bn < 256) bn = 256;
  if (bn > 64 && bn < 128) bn = 128;
  // (a narrower N tile would fill the last wave better on the 13x13 maps, but halves the arithmetic intensity per
  //  activation byte: measured 1.4x slower, so the tile stays 256 wide wherever the layer allows it)
  p.bn = bn;
  p.b_bytes = (bn / 2) * 128;       // per CTA: half of the weight tile
  p.m_rows = g.m_rows();
  p.store_rows = out_rows;
  p.m_tiles = (int)((p.m_rows + 255) / 256);
  p.n_tiles = (cout + bn - 1) / bn;
  p.kc_per_tap = (cin + 63) / 64;
  p.cin = cin;
  p.taps = taps;
  for (int t = 0; t < 9; t++) p.shifts[t] = (taps == 9) ? ((t / 3) - 1) * g.Wp() + ((t % 3) - 1) : 0;
  p.Wp = g.Wp(); p.HpWp = g.HpWp();
  p.cout = cout;
  p.idesc = (umma_idesc_f16(a_fmt, b_fmt, 0, 0, bn) & ~(0x1Fu << 24)) | ((uint32_t)(256 >> 4) << 24);   // M = 256
  // Band mode (round 2): 3x3, <= 64 input channels, split operands, one pair per cluster.  The per-tap kernel re-fetches the activation
  // tile nine times (blocks 3 / 5 forward: 2.44 GB staged per launch, fill-bound at 48 % tensor-active); a 136-row band per kernel row serves the
  // three horizontal taps through the descriptor start address (conv_band.cu's trick, here for cta_group::2): 252 KB instead of 432 KB per tile.
  static const int band_on = []() { const char* e = getenv("SSP_TC2_BAND"); return e ? atoi(e) : 1; }();
  p.band = (band_on && taps == 9 && cin <= 64 && p.n_terms == 3 && !fa) ? 1 : 0;
  const int fixed = 2 * kAccCols * 8 + (2 * kMaxStages + 4) * 8 + 16 + 1024;
  if (p.band && (227 * 1024 - fixed) / (2 * kBandBytes2 + 3 * 2 * p.b_bytes) < 2) p.band = 0;      // wide N tiles: two band stages do not fit
  p.stage_bytes = p.band ? 2 * kBandBytes2 + 3 * 2 * p.b_bytes : (p.n_terms == 3 ? 2 : 1) * (kABytes + p.b_bytes);
  int stages = (227 * 1024 - fixed) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) return fail_msg(SSP_ERR_ARG, "conv_gemm_tc2: tile does not fit shared memory");
  p.stages = stages;
  p.out = out; p.out_ld = out_ld; p.bias = bias; p.stat_sum = stat_sum; p.stat_sq = stat_sq; p.epi = epi;
  if (fa) p.fa = *fa; else p.fa = FusedAct{nullptr, nullptr, 1.f, nullptr, nullptr, 0, 0};
  int rc = 0;
  rc |= tmap_2d_16bit(&p.tmA[0], a_hi, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 128, a_fmt == FMT_BF16);
  // clusters of two pairs (weight tile multicast) unless switched off, a one-tile layer, or split K (its work items are not paired)
  static const int csz_env = []() { const char* e = getenv("SSP_TC2_CLUSTER"); return e ? atoi(e) : 2; }();      // 4 = opt-in (measured slower, see the kernel comment)
  const int csz = (csz_env == 4 && p.m_tiles >= 2 && !p.band) ? 4 : 2;
  if (p.band) {
    rc |= tmap_2d_16bit(&p.tmAband[0], a_hi, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 136, a_fmt == FMT_BF16);
    rc |= tmap_2d_16bit(&p.tmAband[1], a_lo, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 136, a_fmt == FMT_BF16);
  }
  const int brows = csz == 4 ? bn / 4 : bn / 2;
  rc |= tmap_2d_16bit(&p.tmB[0], b_hi, (uint64_t)taps * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, brows, b_fmt == FMT_BF16);
  if (p.n_terms == 3) {
    rc |= tmap_2d_16bit(&p.tmA[1], a_lo, (uint64_t)cin, (uint64_t)a_rows, (uint64_t)a_ld, 64, 128, a_fmt == FMT_BF16);
    rc |= tmap_2d_16bit(&p.tmB[1], b_lo, (uint64_t)taps * cin, (uint64_t)b_rows, (uint64_t)b_ld, 64, brows, b_fmt == FMT_BF16);
  }
  if (rc) return fail_msg(SSP_ERR_DRIVER, "conv_gemm_tc2: cuTensorMapEncodeTiled failed (no driver, or misaligned operand)");
  const int smem_bytes = stages * p.stage_bytes + fixed;
  static int configured = 0;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc2_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc2_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc2_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    configured = 1;
  }
  // Wave quantisation: the 13x13 / 26x26 maps give 49 / 183 pair tiles per N tile, i.e. 98 ... 245 tiles for 74 CTA pairs (66 ... 88 %
  // of the last wave busy).  For the plain-store epilogue (data gradient) the K loop is cut so that tiles x ksplit fills whole waves;
  // the partial sums are added with fp32 vector atomics into the zeroed output.  (49 x 3 = 147 ~ 2 x 74: three cuts fit almost exactly.)
  p.ksplit = 1;
  static const int splitk_on = []() { const char* e = getenv("SSP_DGRAD_SPLITK"); return e ? atoi(e) : 0; }();
  if (splitk_on && csz == 2 && epi == EPI_F32 && !fa && out) {
    const int npairs = g_num_sms2 / 2, tiles = p.m_tiles * p.n_tiles, kb = taps * p.kc_per_tap;
    double best = (double)tiles / ((double)((tiles + npairs - 1) / npairs) * npairs);
    for (int sp = 2; sp <= 4; sp++) {
      if (kb / sp < 8) break;                                  // keep >= 8 k-blocks (512 K elements) per item
      const int items = tiles * sp;
      const double eff = (double)items / ((double)((items + npairs - 1) / npairs) * npairs) - 0.03 * (sp - 1);   // atomics are not free
      if (eff > best + 0.05) { best = eff; p.ksplit = sp; }
    }
    if (p.ksplit > 1) {
      cudaError_t e = cudaMemsetAsync(out, 0, (size_t)out_rows * out_ld * sizeof(float), stream);
      if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
    }
  }
  const int np = csz / 2;
  const int total = ((p.m_tiles + np - 1) / np) * p.n_tiles * p.ksplit;      // work items of a cluster
  // A persistent kernel must not launch more clusters than are co-resident: a cluster of 4 needs 4 free SMs in ONE GPC, and the
  // GPCs of a B200 do not all hold a multiple of 4 (round 2: with 148 / 4 = 37 clusters the stragglers ran as a second wave after
  // the first had finished ALL the work items of their stride -- forward 4.8 -> 6.7 ms).  Ask the driver once per kernel.
  static int max_clusters4[2] = {0, 0};
  int clusters = g_num_sms2 / csz;
  if (csz == 4) {
    int& mc = max_clusters4[fa ? 1 : 0];
    if (!mc) {
      cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(4 * (g_num_sms2 / 4)); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      cudaError_t e = fa ? cudaOccupancyMaxActiveClusters(&n, conv_tc2_kernel<1, 4>, &cfg) : cudaOccupancyMaxActiveClusters(&n, conv_tc2_kernel<0, 4>, &cfg);
      mc = (e == cudaSuccess && n > 0) ? n : g_num_sms2 / 8;                // conservative if the query is not available
    }
    if (clusters > mc) clusters = mc;
  }
  if (total < clusters) clusters = total;
  const unsigned grid = (unsigned)(csz * clusters);
  if (csz == 4) {
    if (fa) conv_tc2_kernel<1, 4><<<grid, kThreads, smem_bytes, stream>>>(p);
    else conv_tc2_kernel<0, 4><<<grid, kThreads, smem_bytes, stream>>>(p);
  } else {
    if (fa) conv_tc2_kernel<1, 2><<<grid, kThreads, smem_bytes, stream>>>(p);
    else conv_tc2_kernel<0, 2><<<grid, kThreads, smem_bytes, stream>>>(p);   // __cluster_dims__(2,1,1): CTA pairs on one TPC
  }
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
