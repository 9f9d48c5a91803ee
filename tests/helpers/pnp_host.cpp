// Host harness: compiles singleshotpose_b200/csrc/pnp_core.h (the arithmetic of pnp_kernel) with g++ so that the CPU suite can check
// it against the reference-generated goldens (tests/golden/pnp*.npz).  Test infrastructure; never loaded by the product.
#include "../../singleshotpose_b200/csrc/pnp_core.h"

extern "C" int h_pnp(const float* p3, int shared, const float* uv, const float* K, int np, long long n, int max_iter, double* R, double* t, int* work) {
  if (np < 6 || np > PNP_MAXP) return -1;
  for (long long i = 0; i < n; i++)
    ssp_pnp::pnp_solve_one(p3 + (shared ? 0 : i * 3 * np), uv + i * 2 * np, K, np, max_iter, R + i * 9, t + i * 3, work + i * 3);
  return 0;
}
