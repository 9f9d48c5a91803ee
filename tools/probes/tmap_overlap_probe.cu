// Is a tensor map whose rows OVERLAP legal?  dims {64 elements, R rows}, row stride 64 B (= 32 fp16): row r = channels of pixel r followed by
// the channels of pixel r+1 of a [pixels][32] fp16 plane.  If cuTensorMapEncodeTiled accepts it and TMA delivers it, one 64-wide box holds the
// operands of TWO horizontally adjacent conv taps of a 32-channel layer (no zero-fill half).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I singleshotpose_b200/csrc -o /tmp/tmap_probe tools/probes/tmap_overlap_probe.cu && /tmp/tmap_probe
#include "ssp_common.cuh"
#include "tmap.cuh"
#include <stdio.h>
#include <vector>
namespace ssp {
int fail_cuda(cudaError_t e, const char* file, int line) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), file, line); return -2; }
int fail_msg(int code, const char* msg) { fprintf(stderr, "%s\n", msg); return code; }
}
using namespace ssp;
__global__ void k(const __grid_constant__ CUtensorMap tm, uint16_t* out, int row0) {
  __shared__ __align__(1024) uint8_t tile[64 * 128];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(&bar, 64 * 128); tma_load_2d(tile, &tm, &bar, 0, row0); mbar_wait(&bar, 0); }
  __syncthreads();
  // undo the 128-B swizzle: 16-B chunk c of row r sits at chunk c ^ (r & 7)
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, e = i % 64, c = e / 8;
    out[i] = ((const uint16_t*)tile)[r * 64 + ((c ^ (r & 7)) * 8) + (e % 8)];
  }
}
int main() {
  const int R = 1000, C = 32;
  std::vector<uint16_t> h(R * C); for (int i = 0; i < R * C; i++) h[i] = (uint16_t)(i & 0xffff);
  uint16_t *d, *o; cudaMalloc(&d, h.size() * 2); cudaMalloc(&o, 64 * 64 * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  // inner extent 64 elements, R - 1 rows, row pitch 32 elements (64 B)
  int rc = tmap_2d_16bit(&tm, d, 64, R - 1, 32, 64, 64, false);
  printf("cuTensorMapEncodeTiled (overlapping rows, stride 64 B < inner 128 B): %s\n", rc ? "REJECTED" : "accepted");
  if (rc) return 0;
  k<<<1, 128>>>(tm, o, 5);
  cudaError_t e = cudaDeviceSynchronize(); printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<uint16_t> g(64 * 64); cudaMemcpy(g.data(), o, g.size() * 2, cudaMemcpyDeviceToHost);
  int bad = 0; for (int r = 0; r < 64; r++) for (int c = 0; c < 64; c++) if (g[r * 64 + c] != h[(5 + r) * 32 + c]) bad++;
  printf("box rows 5..68: %d of 4096 elements differ from [pixel r | pixel r+1]\n", bad);
  return 0;
}
