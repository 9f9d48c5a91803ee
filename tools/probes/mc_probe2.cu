// Round-2 probe #2 (standalone): does TMA multicast raise the operand-fill rate per SM?  mc_probe.cu could not tell: its multicast mode
// re-armed every slot through a cluster handshake with only 4 tiles in flight and came out latency-bound.  Here every CTA arms its
// barrier ONCE per burst of B tiles (expect_tx = B * 16 KB), all CTAs of the cluster sync, then every CTA issues its share of the
// burst back to back into a ring of smem slots (data is never read, slots are overwritten freely), waits for the B * 16 KB and
// repeats.  Per burst every CTA RECEIVES B tiles in every mode; what changes is who asks L2 for them:
//   csz = 1 : every CTA loads its own B tiles                                  (L2 reads = delivered bytes)
//   csz = 2, 4, 8 : each CTA loads 1/csz of every tile ([128/csz rows][64]) and multicasts it to the whole cluster (L2 reads = delivered / csz)
// If delivered bytes / clk / SM rise with csz, the ~6 kB/clk cap of the conv GEMMs is on the L2 side and operand multicast across
// CTA pairs is worth building; if they stay flat, the cap is the SM's fill port.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I singleshotpose_b200/csrc -o /tmp/mc_probe2 tools/probes/mc_probe2.cu && /tmp/mc_probe2
#include "ssp_common.cuh"
#include "tmap.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace ssp {
int fail_cuda(cudaError_t e, const char* file, int line) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line); return -2; }
int fail_msg(int code, const char* msg) { fprintf(stderr, "%s\n", msg); return code; }
}
using namespace ssp;

static constexpr int kTile = 128 * 128;   // 128 rows x 64 x 2 B
static constexpr int kSlots = 12;         // 192 KB ring
static constexpr int kBurst = 48;         // tiles per burst (768 KB per CTA, expect_tx limit is 1 MB)

__device__ __forceinline__ uint32_t q_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void q_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CSZ>
__global__ void __launch_bounds__(128, 1) probe2_kernel(const __grid_constant__ CUtensorMap tm, int bursts, int tiles_total, long long* clocks) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + kSlots * kTile);
  const uint32_t rank = CSZ == 1 ? 0 : q_ctarank();
  const int unit = blockIdx.x / CSZ;                                   // CTAs of a cluster walk the same tiles
  constexpr int part = 128 / CSZ;                                      // rows of every tile loaded by one CTA
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (CSZ > 1) q_cluster_sync();
  long long t0 = 0;
  for (int b = 0; b < bursts; b++) {
    if (threadIdx.x == 0) mbar_expect_tx(bar, kBurst * kTile);
    if (CSZ > 1) q_cluster_sync();                                     // every barrier of the cluster is armed before anybody multicasts
    if (b == 1 && threadIdx.x == 0) t0 = clock64();                    // burst 0 warms up
    if (threadIdx.x == 0) {
      for (int j = 0; j < kBurst; j++) {
        const int tile = (unit * 7 + (b * kBurst + j) * 31) % tiles_total;
        uint8_t* dst = smem + (j % kSlots) * kTile + rank * part * 128;
        if (CSZ == 1) tma_load_2d(dst, &tm, bar, 0, tile * 128);
        else tma_load_2d_mc(dst, &tm, bar, 0, tile * 128 + (int)rank * part, (uint16_t)((1u << CSZ) - 1));
      }
      mbar_wait(bar, b & 1);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) clocks[blockIdx.x] = clock64() - t0;
  if (CSZ > 1) q_cluster_sync();
}

template <int CSZ>
static int run(const char* what, const void* buf, int tiles_total, int sms, long long* clk) {
  CUtensorMap tm;
  if (tmap_2d_16bit(&tm, buf, 64, (uint64_t)tiles_total * 128, 64, 64, 128 / CSZ, false)) { fprintf(stderr, "tensor map failed\n"); return 1; }
  const int smem_bytes = kSlots * kTile + 64 + 1024;
  cudaFuncSetAttribute(probe2_kernel<CSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (CSZ > 8) cudaFuncSetAttribute(probe2_kernel<CSZ>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int bursts = 41, grid = (sms / CSZ) * CSZ;
  for (int rep = 0; rep < 2; rep++) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CSZ; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe2_kernel<CSZ>, tm, bursts, tiles_total, clk);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return 2; }
  }
  long long mx = 0; for (int i = 0; i < grid; i++) if (clk[i] > mx) mx = clk[i];
  const double bytes_sm = (double)(bursts - 1) * kBurst * kTile;
  printf("%-34s grid %3d  %.1f B/clk/SM delivered  (%.2f kB/clk chip-wide delivered, %.2f kB/clk read from L2)\n", what, grid,
         bytes_sm / (double)mx, bytes_sm * grid / (double)mx / 1e3, bytes_sm * grid / CSZ / (double)mx / 1e3);
  return 0;
}

int main() {
  int dev = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_total = 2048;                              // 32 MB: L2-resident after the warm-up launch
  uint16_t* buf; cudaMalloc(&buf, (size_t)tiles_total * kTile); cudaMemset(buf, 1, (size_t)tiles_total * kTile);
  long long* clk; cudaMallocManaged(&clk, sizeof(long long) * 1024);
  int rc = 0;
  rc |= run<1>("unicast (own tiles)", buf, tiles_total, sms, clk);
  rc |= run<2>("multicast, cluster of 2", buf, tiles_total, sms, clk);
  rc |= run<4>("multicast, cluster of 4", buf, tiles_total, sms, clk);
  rc |= run<8>("multicast, cluster of 8", buf, tiles_total, sms, clk);
  return rc;
}
