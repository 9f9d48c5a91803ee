// Device-vs-host probe of pnp_core.h: the same pnp_solve_one() runs on the GPU (one thread per problem) and on the CPU inside this
// executable; prints the first problems whose DLT initialisation or final pose differ.  Build variants decide whether a difference
// comes from FMA contraction or from the optimiser:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 [-fmad=false] -o /tmp/pnp_probe tools/probes/pnp_probe.cu && /tmp/pnp_probe uv.bin
// uv.bin: int32 n, int32 np, float32 K[9], float32 P3[np*3], float32 uv[n*np*2]  (tools/probes/pnp_probe_data.py)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../singleshotpose_b200/csrc/pnp_core.h"

__global__ void k(const float* P3, const float* uv, const float* K, int np, int n, double* R, double* t, int* work, double* dbg) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n) ssp_pnp::pnp_solve_one(P3, uv + id * 2 * np, K, np, 20, R + id * 9, t + id * 3, work + id * 3, dbg + id * 20);
}
static double ang(const double* a, const double* b) {
  double tr = 0; for (int i = 0; i < 9; i++) tr += a[i] * b[i];
  double c = (tr - 1) / 2; c = c > 1 ? 1 : (c < -1 ? -1 : c);
  return acos(c) * 180.0 / M_PI;
}
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); if (!f) { printf("no input\n"); return 2; }
  int n, np; fread(&n, 4, 1, f); fread(&np, 4, 1, f);
  std::vector<float> K(9), P3(np * 3), uv((size_t)n * np * 2);
  fread(K.data(), 4, 9, f); fread(P3.data(), 4, np * 3, f); fread(uv.data(), 4, uv.size(), f); fclose(f);
  std::vector<double> Rh(n * 9), th(n * 3), dh(n * 20), Rd(n * 9), td(n * 3), dd(n * 20);
  std::vector<int> wh(n * 3), wd(n * 3);
  for (int i = 0; i < n; i++) ssp_pnp::pnp_solve_one(P3.data(), uv.data() + (size_t)i * 2 * np, K.data(), np, 20, &Rh[i * 9], &th[i * 3], &wh[i * 3], &dh[i * 20]);
  float *dP3, *duv, *dK; double *dR, *dt, *ddbg; int* dw;
  cudaMalloc(&dP3, P3.size() * 4); cudaMalloc(&duv, uv.size() * 4); cudaMalloc(&dK, 36);
  cudaMalloc(&dR, n * 72); cudaMalloc(&dt, n * 24); cudaMalloc(&ddbg, n * 160); cudaMalloc(&dw, n * 12);
  cudaMemcpy(dP3, P3.data(), P3.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(duv, uv.data(), uv.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dK, K.data(), 36, cudaMemcpyHostToDevice);
  k<<<(n + 63) / 64, 64>>>(dP3, duv, dK, np, n, dR, dt, dw, ddbg);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(Rd.data(), dR, n * 72, cudaMemcpyDeviceToHost); cudaMemcpy(td.data(), dt, n * 24, cudaMemcpyDeviceToHost);
  cudaMemcpy(dd.data(), ddbg, n * 160, cudaMemcpyDeviceToHost); cudaMemcpy(wd.data(), dw, n * 12, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < n; i++) {
    const double a = ang(&Rh[i * 9], &Rd[i * 9]);
    double dv = 0; for (int j = 0; j < 12; j++) dv = fmax(dv, fabs(fabs(dh[i * 20 + 1 + j]) - fabs(dd[i * 20 + 1 + j])));
    if (a > 1e-3 || dv > 1e-6) {
      if (bad < 6) {
        printf("problem %d: final angle host-vs-device %.4g deg | eig host %.3e dev %.3e | det host %.3e dev %.3e | sweeps %d/%d iters %d/%d | eigvec maxdiff %.3e\n", i, a,
               dh[i * 20], dd[i * 20], dh[i * 20 + 13], dd[i * 20 + 13], wh[i * 3], wd[i * 3], wh[i * 3 + 1], wd[i * 3 + 1], dv);
        printf("   init host:"); for (int j = 0; j < 6; j++) printf(" %.5f", dh[i * 20 + 14 + j]); printf("\n   init dev :"); for (int j = 0; j < 6; j++) printf(" %.5f", dd[i * 20 + 14 + j]);
        printf("\n   vec host:"); for (int j = 0; j < 12; j++) printf(" %.4f", dh[i * 20 + 1 + j]); printf("\n   vec dev :"); for (int j = 0; j < 12; j++) printf(" %.4f", dd[i * 20 + 1 + j]); printf("\n");
      }
      bad++;
    }
  }
  printf("pnp_probe: %d of %d problems differ between host and device\n", bad, n);
  return 0;
}
