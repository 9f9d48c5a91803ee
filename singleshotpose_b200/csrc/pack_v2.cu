// EXPERIMENTAL (opt-in: SSP_PACK=v2; compiled, NOT yet run on hardware -- written after the round-1 GPU budget was spent;
// the default path is pack_weights_kernel in elementwise.cu).  Same outputs, bit for bit, as pack_weights_kernel:
//   fwd  : hi/lo fp16 [cout][ld_f]   k = tap*cin + ci
//   dgrad: 16-bit     [cin][ld_d]    k = tap'*cout + co, tap' = taps-1-tap
// Why: the default kernel writes the transposed dgrad copy with a 2-byte store per thread and a row stride between
// neighbouring threads (1.06 TB/s on the largest layer, 0.53 ms per step in total, profiles/r01_launches_bench_b64.txt);
// this one moves a 64(co) x 64(ci) tile through shared memory so that BOTH sides are written in 128-B rows.
#include "ssp_common.cuh"

namespace ssp {

__global__ void __launch_bounds__(256) pack_weights_tiled_kernel(const float* __restrict__ w, int cout, int taps, int cin,
                                                                 uint16_t* __restrict__ f_hi, uint16_t* __restrict__ f_lo, int ld_f,
                                                                 uint16_t* __restrict__ d, int ld_d, int d_fmt) {
  __shared__ uint16_t tile[64][66];                 // [ci][co], +2 pad: 33-word row pitch, conflict-free both ways
  const int ci0 = blockIdx.x * 64, co0 = blockIdx.y * 64, tap = blockIdx.z;
  const int lane64 = threadIdx.x & 63, grp = threadIdx.x >> 6;      // 4 groups of 64 threads
#pragma unroll 4
  for (int r = 0; r < 16; r++) {
    const int co = co0 + r * 4 + grp, ci = ci0 + lane64;            // consecutive threads -> consecutive ci (coalesced fp32 reads)
    uint16_t t = 0;
    if (co < cout && ci < cin) {
      const float v = w[((long long)co * taps + tap) * cin + ci];
      if (f_hi) {
        uint16_t a, b; split_f16(v, a, b);
        const long long o = (long long)co * ld_f + tap * cin + ci;
        f_hi[o] = a; if (f_lo) f_lo[o] = b;
      }
      t = cvt_f32_to_16(v, d_fmt);
    }
    tile[lane64][r * 4 + grp] = t;
  }
  if (!d) return;
  __syncthreads();
#pragma unroll 4
  for (int r = 0; r < 16; r++) {
    const int ci = ci0 + r * 4 + grp, co = co0 + lane64;            // consecutive threads -> consecutive co (coalesced 16-bit writes)
    if (ci < cin && co < cout) d[(long long)ci * ld_d + (long long)(taps - 1 - tap) * cout + co] = tile[r * 4 + grp][lane64];
  }
}

int pack_weights_v2(const float* w, int cout, int taps, int cin, void* f_hi, void* f_lo, int ld_f, void* d, int ld_d, int d_fmt, cudaStream_t s) {
  if (!w || cout <= 0 || taps <= 0 || cin <= 0 || taps > 65535) return fail_msg(SSP_ERR_ARG, "pack_weights_v2: bad argument");
  dim3 grid((cin + 63) / 64, (cout + 63) / 64, taps);
  if (grid.y > 65535) return fail_msg(SSP_ERR_ARG, "pack_weights_v2: cout too large");
  pack_weights_tiled_kernel<<<grid, 256, 0, s>>>(w, cout, taps, cin, (uint16_t*)f_hi, (uint16_t*)f_lo, ld_f, (uint16_t*)d, ld_d, d_fmt);
  SSP_CHECK_LAUNCH();
  return SSP_OK;
}

}  // namespace ssp
