// First layer (3 -> 32 channels, 3x3, K = 27): direct fp32 convolution on the CUDA cores.
// With K = 27 the layer is HBM-bound (12 flop/B): reading the raw NCHW image (12 B/pixel) instead of an im2col'ed
// operand plane (128 B/pixel) and writing Y once is the roofline; exact fp32 FMA, no operand splitting needed.
// Replaces nn.Conv2d(3, 32, 3, 1, 1) of reference darknet.py:156 for block 0 (+ the BN batch statistics epilogue).
#include "ssp_common.cuh"

namespace ssp {

static constexpr int kC0 = 32;          // output channels
static constexpr int kPix = 256;        // pixels per tile (= threads per block)

__global__ void __launch_bounds__(kPix, 2) conv0_direct_kernel(const float* __restrict__ x, const float* __restrict__ wgt /*[32][27]*/,
                                                               const float* __restrict__ bias, float* __restrict__ y, int y_ld,
                                                               double* __restrict__ ssum, double* __restrict__ ssq, int N, int H, int W) {
  __shared__ __align__(16) float sw[27][kC0];            // [k][co]: a thread reads its 32 weights of tap k as 8 broadcast float4
  __shared__ float stile[kPix][kC0 + 1];                 // output tile, padded against bank conflicts on the transposed read
  __shared__ double sred[8][kC0][2];
  __shared__ long long srow[kPix];                      // output row of every pixel of the tile (-1: past the end)
  for (int i = threadIdx.x; i < 27 * kC0; i += kPix) sw[i / kC0][i % kC0] = wgt[(i % kC0) * 27 + (i / kC0)];
  __syncthreads();
  const long long npix = (long long)N * H * W;
  const long long ntiles = (npix + kPix - 1) / kPix;
  Geom g{N, H, W};
  double acc1 = 0.0, acc2 = 0.0;                          // statistics of channel (tid % 32) over this thread's row subset
  const int sc = threadIdx.x % kC0, sr = threadIdx.x / kC0;    // 8 row groups
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long pix = tile * kPix + threadIdx.x;
    const bool ok = pix < npix;
    float acc[kC0];
#pragma unroll
    for (int c = 0; c < kC0; c++) acc[c] = bias ? bias[c] : 0.f;
    long long orow = 0;
    if (ok) {
      const unsigned up = (unsigned)pix, tq = up / (unsigned)W;       // 32-bit index math (64-bit div/mod is ~10x the cost)
      const unsigned w = up - tq * (unsigned)W, n = tq / (unsigned)H, h = tq - n * (unsigned)H;
      orow = g.row((int)n, (int)h, (int)w);
      const int HW = H * W;
      const float* px = x + (long long)n * 3 * HW + (int)(h * W + w);     // this pixel, channel 0: one 64-bit address per pixel,
#pragma unroll                                                            // the 27 taps are small 32-bit offsets from it
      for (int tap = 0; tap < 9; tap++) {
        const int hh = (int)h + tap / 3 - 1, ww = (int)w + tap % 3 - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
        const int off = (tap / 3 - 1) * W + (tap % 3 - 1);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const float v = in ? __ldg(px + (c * HW + off)) : 0.f;
          const float4* wr = reinterpret_cast<const float4*>(&sw[tap * 3 + c][0]);
#pragma unroll
          for (int q = 0; q < kC0 / 4; q++) {
            const float4 wv = wr[q];
            acc[4 * q] = fmaf(v, wv.x, acc[4 * q]); acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < kC0; c++) stile[threadIdx.x][c] = ok ? acc[c] : 0.f;
    srow[threadIdx.x] = ok ? orow : -1;
    __syncthreads();
    // coalesced store: pixels of one image row are consecutive rows of Y; 8 threads write one 128-B pixel row
    {
#pragma unroll
      for (int it = 0; it < kC0 / 4; it++) {
        const int e = it * kPix + threadIdx.x;           // float4 index within the tile
        const int prow = e / (kC0 / 4), q4 = e % (kC0 / 4);
        const long long r = srow[prow];
        if (r >= 0) {
          float4 v = make_float4(stile[prow][4 * q4], stile[prow][4 * q4 + 1], stile[prow][4 * q4 + 2], stile[prow][4 * q4 + 3]);
          *reinterpret_cast<float4*>(y + r * y_ld + 4 * q4) = v;
        }
      }
    }
    if (ssum) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
      for (int r = sr; r < kPix; r += kPix / kC0) { const float v = stile[r][sc]; s1 += v; s2 = fmaf(v, v, s2); }
      acc1 += (double)s1; acc2 += (double)s2;
    }
    __syncthreads();
  }
  if (ssum) {
    sred[sr][sc][0] = acc1; sred[sr][sc][1] = acc2;
    __syncthreads();
    if (threadIdx.x < 2 * kC0) {
      const int c = threadIdx.x % kC0, which = threadIdx.x / kC0;
      double s = 0.0;
      for (int r = 0; r < kPix / kC0; r++) s += sred[r][c][which];
      atomicAdd((which ? ssq : ssum) + c, s);
    }
  }
}

int conv0_direct(const float* x, const float* w, const float* bias, float* y, int y_ld, double* ssum, double* ssq, int N, int H, int W,
                 cudaStream_t s) {
  if (!x || !w || !y || (y_ld % 4) || y_ld < kC0 || (ssum && !ssq)) return fail_msg(SSP_ERR_ARG, "conv0_direct: bad argument");
  static int sms = 0;
  if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const long long ntiles = ((long long)N * H * W + kPix - 1) / kPix;
  long long grid = (long long)sms * 4; if (grid > ntiles) grid = ntiles;
  conv0_direct_kernel<<<(unsigned)grid, kPix, 0, s>>>(x, w, bias, y, y_ld, ssum, ssq, N, H, W);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

}  // namespace ssp
