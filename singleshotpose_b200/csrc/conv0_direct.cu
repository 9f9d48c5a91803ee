// First layer (3 -> 32 channels, 3x3, K = 27): direct fp32 convolution on the CUDA cores, register-tiled.
// With K = 27 the layer cannot feed the tensor pipe usefully (12 flop per byte of Y); reading the raw NCHW image (12 B/pixel)
// instead of an im2col'ed operand plane (128 B/pixel) and writing Y once is the HBM floor (0.24 ms at batch 64), and
// 864 FMA per pixel is 0.27 ms of the chip's FFMA rate: the kernel has to be FMA-issue-bound, not LDS-bound.
//   tile   : 16 rows x 32 columns of one image per block iteration (256 threads), input halo tile (3 x 18 x 34) in shared memory
//   thread : 4 consecutive pixels x 16 output channels = 64 accumulators; per k: 4 broadcast LDS.128 of weights feed 64 FFMA,
//            the 6 input values of a kernel row feed 3 taps (1 LDS : 10.7 FFMA; round 1's 1 pixel x 32 channels was 1 : 4)
//   stores : the two threads of a pixel group own channels 8q + 4h + j, so every st.128 pair fills one 32-B sector
//   BN     : per-tile fp32 partial sums, folded across the 16 pixel-group lanes by a transposing butterfly (30 shuffles),
//            accumulated in fp64 per lane over the block's tiles, one fp64 atomic per channel per block at the end
// Replaces nn.Conv2d(3, 32, 3, 1, 1) of reference darknet.py:156 for block 0 (+ the BN batch statistics epilogue).
#include "ssp_common.cuh"

namespace ssp {

static constexpr int kC0 = 32;          // output channels
static constexpr int kTH = 16, kTW = 32;
static constexpr int kInW = 37;         // smem row pitch of the halo tile (odd: the two tile rows of a warp hit disjoint banks)

__global__ void __launch_bounds__(256, 2) conv0_direct_kernel(const float* __restrict__ x, const float* __restrict__ wgt /*[32][27]*/,
                                                              const float* __restrict__ bias, float* __restrict__ y, int y_ld,
                                                              double* __restrict__ ssum, double* __restrict__ ssq, int N, int H, int W) {
  __shared__ __align__(16) float sw[27][kC0];            // [k][co], k = tap*3 + c
  __shared__ float sin[3][kTH + 2][kInW];
  __shared__ double sred[2][kC0];
  for (int i = threadIdx.x; i < 27 * kC0; i += 256) sw[i / kC0][i % kC0] = wgt[(i % kC0) * 27 + (i / kC0)];
  if (threadIdx.x < 2 * kC0) sred[threadIdx.x / kC0][threadIdx.x % kC0] = 0.0;
  const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
  const int ntiles = N * tiles_h * tiles_w;
  const int half = threadIdx.x & 1, pg = threadIdx.x >> 1;          // channel interleave, pixel group
  const int prow = pg >> 3, pcol = (pg & 7) * 4;
  const int lane = threadIdx.x & 31;
  Geom g{N, H, W};
  const int HW = H * W;
  double d0 = 0.0, d1 = 0.0;                                         // this lane's two statistics (see the butterfly below)
  float bv[16];
#pragma unroll
  for (int j = 0; j < 16; j++) bv[j] = bias ? bias[8 * (j >> 2) + 4 * half + (j & 3)] : 0.f;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * kTH, w0 = tw * kTW;
    __syncthreads();                                                 // previous tile's readers are done (and sw / sred are set)
    const float* xi = x + (long long)n * 3 * HW;
    for (int i = threadIdx.x; i < 3 * (kTH + 2) * (kTW + 2); i += 256) {
      const int cc = i % (kTW + 2), rr = (i / (kTW + 2)) % (kTH + 2), c = i / ((kTW + 2) * (kTH + 2));
      const int hh = h0 + rr - 1, ww = w0 + cc - 1;
      sin[c][rr][cc] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(xi + c * HW + hh * W + ww) : 0.f;
    }
    __syncthreads();
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; p++)
#pragma unroll
      for (int j = 0; j < 16; j++) acc[p][j] = bv[j];
#pragma unroll
    for (int c = 0; c < 3; c++) {
#pragma unroll
      for (int dy = 0; dy < 3; dy++) {
        float in[6];
#pragma unroll
        for (int i = 0; i < 6; i++) in[i] = sin[c][prow + dy][pcol + i];
#pragma unroll
        for (int dx = 0; dx < 3; dx++) {
          const float* wr = &sw[(dy * 3 + dx) * 3 + c][4 * half];
          float wv[16];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const float4 t = *reinterpret_cast<const float4*>(wr + 8 * q);
            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
          }
#pragma unroll
          for (int p = 0; p < 4; p++)
#pragma unroll
            for (int j = 0; j < 16; j++) acc[p][j] = fmaf(in[p + dx], wv[j], acc[p][j]);
        }
      }
    }
    const int h = h0 + prow;
    const bool row_ok = h < H;
    float st[32];                                                    // [0,16): sum, [16,32): sum of squares over this thread's valid pixels
#pragma unroll
    for (int j = 0; j < 32; j++) st[j] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const int w = w0 + pcol + p;
      if (row_ok && w < W) {
        float* yr = y + g.row(n, h, w) * y_ld + 4 * half;
#pragma unroll
        for (int q = 0; q < 4; q++)
          *reinterpret_cast<float4*>(yr + 8 * q) = make_float4(acc[p][4 * q], acc[p][4 * q + 1], acc[p][4 * q + 2], acc[p][4 * q + 3]);
#pragma unroll
        for (int j = 0; j < 16; j++) { st[j] += acc[p][j]; st[16 + j] = fmaf(acc[p][j], acc[p][j], st[16 + j]); }
      }
    }
    if (ssum) {
      // sum over the 16 lanes that share `half` (lane bits 1..4), halving the value count at every step:
      // afterwards this lane holds the totals of st[lane & 30] and st[(lane & 30) + 1]
      int cnt = 32;
#pragma unroll
      for (int b = 16; b >= 2; b >>= 1) {
        cnt >>= 1;
        const bool up = (lane & b) != 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
          if (i < cnt) {
            const float send = up ? st[i] : st[i + cnt], keep = up ? st[i + cnt] : st[i];
            st[i] = keep + __shfl_xor_sync(0xffffffffu, send, b);
          }
        }
      }
      d0 += (double)st[0]; d1 += (double)st[1];
    }
  }
  if (ssum) {
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int idx = (lane & 30) + e, j = idx & 15;
      atomicAdd(&sred[idx >> 4][8 * (j >> 2) + 4 * half + (j & 3)], e ? d1 : d0);
    }
    __syncthreads();
    if (threadIdx.x < 2 * kC0) {
      const int c = threadIdx.x % kC0, which = threadIdx.x / kC0;
      atomicAdd((which ? ssq : ssum) + c, sred[which][c]);
    }
  }
}

int conv0_direct(const float* x, const float* w, const float* bias, float* y, int y_ld, double* ssum, double* ssq, int N, int H, int W,
                 cudaStream_t s) {
  if (!x || !w || !y || (y_ld % 4) || y_ld < kC0 || (ssum && !ssq)) return fail_msg(SSP_ERR_ARG, "conv0_direct: bad argument");
  static int sms = 0;
  if (!sms) sms = ssp_sm_count();
  const long long ntiles = (long long)N * ((H + kTH - 1) / kTH) * ((W + kTW - 1) / kTW);
  if (ntiles <= 0 || ntiles > 0x7fffffffLL) return fail_msg(SSP_ERR_ARG, "conv0_direct: bad shape");
  long long grid = (long long)sms * 2; if (grid > ntiles) grid = ntiles;      // persistent: 2 resident blocks per SM
  conv0_direct_kernel<<<(unsigned)grid, 256, 0, s>>>(x, w, bias, y, y_ld, ssum, ssq, N, H, W);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

}  // namespace ssp
