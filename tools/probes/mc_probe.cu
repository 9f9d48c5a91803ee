// Round-2 probe (standalone; not part of libssp_b200.so): is the 5.3-6.2 kB/clk operand-fill cap measured in round 1 on the
// L2 side or on the SM side?  Every CTA streams 16 KB TMA tiles ([128 rows][64 fp16], SWIZZLE_128B) from an L2-resident
// buffer through a 6-slot mbarrier ring (4 tiles in flight) and does nothing else.
//   mode 0: 1 CTA per SM, every CTA reads its OWN tiles                      (baseline fill rate)
//   mode 1: clusters of 2, both CTAs need the SAME tile, each reads all of it (same SM-side bytes, same L2-side bytes)
//   mode 2: clusters of 2, each CTA reads HALF of the tile and multicasts it  (same SM-side bytes, HALF the L2-side reads)
// If mode 2 delivers ~2x the bytes/clk/SM of mode 1, the cap is L2-side and weight/activation multicast across a cluster
// will lift the conv GEMMs; if the three modes agree, it is the SM's fill port and multicast buys nothing.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I singleshotpose_b200/csrc -o /tmp/mc_probe tools/probes/mc_probe.cu && /tmp/mc_probe
#include "ssp_common.cuh"
#include "tmap.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace ssp {   // ssp_common.cuh declares these; the probe does not link abi.cu
int fail_cuda(cudaError_t e, const char* file, int line) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line); return -2; }
int fail_msg(int code, const char* msg) { fprintf(stderr, "%s\n", msg); return code; }
}
using namespace ssp;

static constexpr int kTile = 128 * 128;   // 128 rows x 64 x 2 B
static constexpr int kStagesP = 6;

__device__ __forceinline__ uint32_t p_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void p_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// half tile (64 rows) loaded by this CTA and written to the same smem offset in BOTH CTAs; each CTA's own barrier gets the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__device__ __forceinline__ void p_arrive_rank(uint64_t* bar, uint32_t target) {       // arrive on the barrier at this offset in CTA `target`
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
               ::"r"(smem_u32(bar)), "r"(target) : "memory");
}

// One thread per CTA runs a software pipeline over a ring of kStagesP slots:
//   iteration i:  (1) wait for the bytes of iteration i - S (slot free again)   (2) arm the slot's barrier for iteration i and, in
//   mode 2, tell BOTH CTAs of the cluster so (a multicast write credits the barrier of every destination CTA, so nobody may issue
//   into a slot before both barriers are armed)   (3) issue the load of iteration i - D, D = 2 iterations later, by which time the
//   peer's "armed" arrive has crossed the cluster.  S - D = 4 tiles in flight per CTA in every mode.
template <int MODE>
__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmFull, const __grid_constant__ CUtensorMap tmHalf,
                                                       int iters, int tiles_total, long long* clocks) {
  constexpr int S = kStagesP, D = 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S * kTile);
  uint64_t* armed = full + S;
  const uint32_t rank = MODE == 0 ? 0 : p_ctarank();
  const int unit = MODE == 0 ? blockIdx.x : (blockIdx.x >> 1);          // CTAs of a cluster walk the same tiles
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&armed[s], 2); }
    fence_barrier_init();
  }
  __syncthreads();
  if (MODE != 0) p_cluster_sync();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int i = 0; i < iters + S; i++) {
      const int s = i % S;
      if (i >= S) mbar_wait(&full[s], ((i / S) - 1) & 1);                // bytes of iteration i - S have landed
      if (i < iters) {
        mbar_expect_tx(&full[s], kTile);
        if (MODE == 2) { p_arrive_rank(&armed[s], 0); p_arrive_rank(&armed[s], 1); }
      }
      const int j = i - D;
      if (j >= 0 && j < iters) {
        const int sj = j % S;
        const int tile = (unit * 7 + j * 31) % tiles_total;              // scattered tiles of the L2-resident buffer
        if (MODE == 2) {
          mbar_wait(&armed[sj], (j / S) & 1);                            // both CTAs armed slot sj for iteration j
          tma_load_2d_mc(smem + sj * kTile + rank * (kTile / 2), &tmHalf, &full[sj], 0, tile * 128 + (int)rank * 64, (uint16_t)3);
        } else {
          tma_load_2d(smem + sj * kTile, &tmFull, &full[sj], 0, tile * 128);
        }
      }
    }
    clocks[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (MODE != 0) p_cluster_sync();
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int tiles_total = 2048;                              // 2048 x 16 KB = 32 MB: resident in the 126 MB L2 after the warm-up
  uint16_t* buf; cudaMalloc(&buf, (size_t)tiles_total * kTile); cudaMemset(buf, 1, (size_t)tiles_total * kTile);
  long long* clk; cudaMallocManaged(&clk, sizeof(long long) * 1024);
  CUtensorMap tmFull, tmHalf;
  if (tmap_2d_16bit(&tmFull, buf, 64, (uint64_t)tiles_total * 128, 64, 64, 128, false) ||
      tmap_2d_16bit(&tmHalf, buf, 64, (uint64_t)tiles_total * 128, 64, 64, 64, false)) { fprintf(stderr, "tensor map failed\n"); return 1; }
  const int smem_bytes = kStagesP * kTile + 128 + 1024;
  cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int iters = 4000, grid = sms & ~1;
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 2; rep++) {                      // rep 0 warms L2
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      if (mode == 0) probe_kernel<0><<<grid, 128, smem_bytes>>>(tmFull, tmHalf, iters, tiles_total, clk);
      else {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_bytes;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (mode == 1) cudaLaunchKernelEx(&cfg, probe_kernel<1>, tmFull, tmHalf, iters, tiles_total, clk);
        else cudaLaunchKernelEx(&cfg, probe_kernel<2>, tmFull, tmHalf, iters, tiles_total, clk);
      }
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { fprintf(stderr, "mode %d: %s\n", mode, cudaGetErrorString(e)); return 2; }
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      if (rep == 1) {
        long long mx = 0; for (int i = 0; i < grid; i++) if (clk[i] > mx) mx = clk[i];
        const double bytes_sm = (double)iters * kTile;                        // delivered into EACH CTA's smem
        printf("mode %d: %.3f ms  %.1f B/clk/SM delivered (%.2f kB/clk chip-wide, %.0f GB/s)  L2-side reads %s\n", mode, ms,
               bytes_sm / (double)mx, bytes_sm * grid / (double)mx / 1e3, bytes_sm * grid / (ms * 1e-3) / 1e9,
               mode == 2 ? "HALF of delivered" : "equal to delivered");
      }
    }
  }
  return 0;
}
