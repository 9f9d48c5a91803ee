// HBM-bound kernels of the conv stack: layout packing, BatchNorm (train statistics, apply + LeakyReLU fused
// with 2x2 max-pool / reorg / concat placement), BatchNorm+LeakyReLU+pool backward, weight re-packing, SGD.
// They replace nn.BatchNorm2d / nn.LeakyReLU / nn.MaxPool2d / Reorg / torch.cat of reference
// darknet.py:16-35,96-106,156-176 and their autograd, and optim.SGD of train.py:388.
// All activations live in the padded-flat NHWC layout (ssp_common.cuh); 4 channels per thread (16-B fp32 /
// 8-B fp16 vectors), consecutive threads on consecutive channels -> fully coalesced rows.
#include "ssp_common.cuh"

namespace ssp {

// ------------------------------------------------------------------------------------------------
// Layer-0 input: NCHW fp32 image -> im2col'ed rows [row(n,h,w)][32] (k = (kh*3+kw)*3 + c, k >= 27 zero), hi/lo fp16.
__global__ void pack_input_im2col_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                         int N, int H, int W) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W * 32;
  if (idx >= total) return;
  const int k = (int)(idx & 31);
  long long pix = idx >> 5;
  const int w = (int)(pix % W); pix /= W;
  const int h = (int)(pix % H);
  const int n = (int)(pix / H);
  float v = 0.f;
  if (k < 27) {
    const int c = k % 3, tap = k / 3;
    const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(x + (((long long)n * 3 + c) * H + hh) * W + ww);
  }
  Geom g{N, H, W};
  uint16_t a, b; split_f16(v, a, b);
  const long long o = g.row(n, h, w) * 32 + k;
  hi[o] = a; lo[o] = b;
}

// generic NCHW fp32 -> padded-flat rows (hi/lo fp16, or a single 16-bit plane in `fmt` when lo == nullptr)
__global__ void pack_nchw_kernel(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                 int N, int C, int H, int W, int ld, int c0, int fmt, float scale) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * H * W * C;
  if (idx >= total) return;
  const int c = (int)(idx % C);
  long long pix = idx / C;
  const int w = (int)(pix % W); pix /= W;
  const int h = (int)(pix % H);
  const int n = (int)(pix / H);
  const float v = __ldg(x + (((long long)n * C + c) * H + h) * W + w) * scale;
  Geom g{N, H, W};
  const long long o = g.row(n, h, w) * ld + c0 + c;
  if (lo) { uint16_t a, b; split_f16(v, a, b); hi[o] = a; lo[o] = b; }
  else hi[o] = cvt_f32_to_16(v, fmt);
}

// padded-flat fp32 rows -> NCHW fp32 (network output / tests)
__global__ void unpack_nchw_kernel(const float* __restrict__ y, float* __restrict__ out, int N, int C, int H, int W, int ld, int c0) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * C * H * W;
  if (idx >= total) return;
  const int w = (int)(idx % W);
  long long r = idx / W;
  const int h = (int)(r % H); r /= H;
  const int c = (int)(r % C);
  const int n = (int)(r / C);
  Geom g{N, H, W};
  out[idx] = y[g.row(n, h, w) * ld + c0 + c];
}

// padded-flat 16-bit plane(s) -> NCHW fp32 (tests: read activations back)
__global__ void unpack16_nchw_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, float* __restrict__ out,
                                     int N, int C, int H, int W, int ld, int c0, int fmt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)N * C * H * W;
  if (idx >= total) return;
  const int w = (int)(idx % W);
  long long r = idx / W;
  const int h = (int)(r % H); r /= H;
  const int c = (int)(r % C);
  const int n = (int)(r / C);
  Geom g{N, H, W};
  const long long o = g.row(n, h, w) * ld + c0 + c;
  float v = cvt16_to_f32(hi[o], fmt);
  if (lo) v += cvt16_to_f32(lo[o], fmt);
  out[idx] = v;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics -> per-channel affine.  train: batch statistics from the conv epilogue's fp64 sums
// (biased variance for normalisation, unbiased for running_var, momentum, eps as nn.BatchNorm2d(eps=1e-4),
// darknet.py:157); eval: running statistics.  Zeroes the sum buffers for the next step.
__global__ void bn_finalize_kernel(double* __restrict__ ssum, double* __restrict__ ssq, double count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float momentum, float eps, int train,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                   float* __restrict__ scale_out, float* __restrict__ shift_out, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (train) {
    const double m = ssum[c] / count;
    double var = ssq[c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
    ssum[c] = 0.0; ssq[c] = 0.0;
  } else {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  }
  const float sc = gamma[c] * invstd;
  mean_out[c] = mean; invstd_out[c] = invstd;
  scale_out[c] = sc; shift_out[c] = beta[c] - mean * sc;
}

// ------------------------------------------------------------------------------------------------
// z = leaky(y*scale + shift) written to up to two destinations.
enum { DST_NONE = 0, DST_DIRECT = 1, DST_POOL = 2, DST_REORG = 3 };
struct ActDst {
  uint16_t* hi; uint16_t* lo;   // lo may be null (single plane in fmt)
  int ld, c0, kind;
};
struct BnApplyParams {
  const float* y; int y_ld;
  const float* scale; const float* shift;
  int N, C, H, W;
  float slope;          // 0.1 leaky, 1.0 linear
  ActDst dst[2];
};

__device__ __forceinline__ float leaky(float v, float slope) { return v > 0.f ? v : v * slope; }

__device__ __forceinline__ void store4(const ActDst& d, long long row, int c, const float (&z)[4]) {
  const long long o = row * d.ld + d.c0 + c;
  uint16_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; j++) split_f16(z[j], h[j], l[j]);
  *reinterpret_cast<uint2*>(d.hi + o) = make_uint2(h[0] | ((uint32_t)h[1] << 16), h[2] | ((uint32_t)h[3] << 16));
  if (d.lo) *reinterpret_cast<uint2*>(d.lo + o) = make_uint2(l[0] | ((uint32_t)l[1] << 16), l[2] | ((uint32_t)l[3] << 16));
}

// POOLED = true : one thread = one 2x2 window x 4 channels (needed when any destination is DST_POOL)
template <bool POOLED>
__global__ void __launch_bounds__(256) bn_apply_kernel(const BnApplyParams p) {
  const int cg = p.C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Hs = POOLED ? p.H / 2 : p.H, Ws = POOLED ? p.W / 2 : p.W;
  const long long total = (long long)p.N * Hs * Ws * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 4;
  long long pix = idx / cg;
  const int ws = (int)(pix % Ws); pix /= Ws;
  const int hs = (int)(pix % Hs);
  const int n = (int)(pix / Hs);
  const float4 sc = *reinterpret_cast<const float4*>(p.scale + c);
  const float4 sh = *reinterpret_cast<const float4*>(p.shift + c);
  Geom g{p.N, p.H, p.W};
  Geom gh{p.N, p.H / 2, p.W / 2};
  float zmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  constexpr int NP = POOLED ? 4 : 1;
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const int h = POOLED ? hs * 2 + (q >> 1) : hs, w = POOLED ? ws * 2 + (q & 1) : ws;
    const long long row = g.row(n, h, w);
    const float4 yv = *reinterpret_cast<const float4*>(p.y + row * p.y_ld + c);
    float z[4] = {leaky(fmaf(yv.x, sc.x, sh.x), p.slope), leaky(fmaf(yv.y, sc.y, sh.y), p.slope),
                  leaky(fmaf(yv.z, sc.z, sh.z), p.slope), leaky(fmaf(yv.w, sc.w, sh.w), p.slope)};
#pragma unroll
    for (int j = 0; j < 4; j++) zmax[j] = fmaxf(zmax[j], z[j]);
#pragma unroll
    for (int d = 0; d < 2; d++) {
      if (p.dst[d].kind == DST_DIRECT) store4(p.dst[d], row, c, z);
      else if (p.dst[d].kind == DST_REORG)       // marvis ordering, darknet.py:31-34: ch = ((h%2)*2 + w%2)*C + c
        store4(p.dst[d], gh.row(n, h >> 1, w >> 1), ((h & 1) * 2 + (w & 1)) * p.C + c, z);
    }
  }
  if (POOLED) {
#pragma unroll
    for (int d = 0; d < 2; d++)
      if (p.dst[d].kind == DST_POOL) store4(p.dst[d], gh.row(n, hs, ws), c, zmax);
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of (BN -> leaky -> {direct | 2x2 max-pool | reorg} consumers).  Upstream gradients are fp32
// matrices in the consumers' geometries; dz = (sum of routed grads) * leaky'(z).
//   pass 1 (reduce): per-channel  S1 = sum dz,  S2 = sum dz * xhat          (fp64 atomics, one per block/channel)
//   pass 2 (apply):  dY = gamma*invstd * (dz - S1/cnt - xhat*S2/cnt)  -> 16-bit plane (operand of dgrad/wgrad)
enum { SRC_NONE = 0, SRC_DIRECT = 1, SRC_POOL = 2, SRC_REORG = 3 };
struct GradSrc { const float* g; int ld, c0, kind; };
struct BnBwdParams {
  const float* y; int y_ld;
  const float* scale; const float* shift; const float* mean; const float* invstd; const float* gamma;
  int N, C, H, W; float slope; int has_bn;
  GradSrc src[2];
  double* s1; double* s2; double count;
  uint16_t* dy; int dy_ld, dy_fmt; float dy_scale;
};

template <bool POOLED>
__device__ __forceinline__ void bn_bwd_gather(const BnBwdParams& p, int n, int hs, int ws, int c,
                                              float (&dz)[POOLED ? 4 : 1][4], float (&xh)[POOLED ? 4 : 1][4],
                                              long long (&rows)[POOLED ? 4 : 1]) {
  constexpr int NP = POOLED ? 4 : 1;
  Geom g{p.N, p.H, p.W};
  Geom gh{p.N, p.H / 2, p.W / 2};
  float4 sc = make_float4(1, 1, 1, 1), sh = make_float4(0, 0, 0, 0), mu = sh, is = sc;
  if (p.has_bn) {
    sc = *reinterpret_cast<const float4*>(p.scale + c); sh = *reinterpret_cast<const float4*>(p.shift + c);
    mu = *reinterpret_cast<const float4*>(p.mean + c); is = *reinterpret_cast<const float4*>(p.invstd + c);
  }
  float z[NP][4];
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const int h = POOLED ? hs * 2 + (q >> 1) : hs, w = POOLED ? ws * 2 + (q & 1) : ws;
    rows[q] = g.row(n, h, w);
    const float4 yv = *reinterpret_cast<const float4*>(p.y + rows[q] * p.y_ld + c);
    z[q][0] = fmaf(yv.x, sc.x, sh.x); z[q][1] = fmaf(yv.y, sc.y, sh.y);
    z[q][2] = fmaf(yv.z, sc.z, sh.z); z[q][3] = fmaf(yv.w, sc.w, sh.w);
    xh[q][0] = (yv.x - mu.x) * is.x; xh[q][1] = (yv.y - mu.y) * is.y;
    xh[q][2] = (yv.z - mu.z) * is.z; xh[q][3] = (yv.w - mu.w) * is.w;
#pragma unroll
    for (int j = 0; j < 4; j++) dz[q][j] = 0.f;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const GradSrc& gs = p.src[s];
      if (gs.kind == SRC_DIRECT) {
        const float4 gv = *reinterpret_cast<const float4*>(gs.g + rows[q] * gs.ld + gs.c0 + c);
        dz[q][0] += gv.x; dz[q][1] += gv.y; dz[q][2] += gv.z; dz[q][3] += gv.w;
      } else if (gs.kind == SRC_REORG) {
        const float4 gv = *reinterpret_cast<const float4*>(gs.g + gh.row(n, h >> 1, w >> 1) * gs.ld + gs.c0 +
                                                            ((h & 1) * 2 + (w & 1)) * p.C + c);
        dz[q][0] += gv.x; dz[q][1] += gv.y; dz[q][2] += gv.z; dz[q][3] += gv.w;
      }
    }
  }
  if (POOLED) {
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const GradSrc& gs = p.src[s];
      if (gs.kind != SRC_POOL) continue;
      const float4 gv4 = *reinterpret_cast<const float4*>(gs.g + gh.row(n, hs, ws) * gs.ld + gs.c0 + c);
      const float gv[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        // argmax of the ACTIVATED values; first maximum in (h, w) scan order wins (max_pool2d semantics)
        int best = 0; float bv = leaky(z[0][j], p.slope);
#pragma unroll
        for (int q = 1; q < NP; q++) { const float a = leaky(z[q][j], p.slope); if (a > bv) { bv = a; best = q; } }
#pragma unroll
        for (int q = 0; q < NP; q++) if (q == best) dz[q][j] += gv[j];
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NP; q++)
#pragma unroll
    for (int j = 0; j < 4; j++) dz[q][j] *= (z[q][j] > 0.f ? 1.f : p.slope);
}

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdParams p) {
  // thread layout: channel group = tid % CG (4 channels each), pixel lane = tid / CG
  extern __shared__ float red[];           // [2][PL][CG*4]
  const int cgs = p.C >> 2;
  const int CG = cgs < 256 ? cgs : 256;
  const int PL = 256 / CG;
  const int cgi = threadIdx.x % CG, pl = threadIdx.x / CG;
  const int Hs = POOLED ? p.H / 2 : p.H, Ws = POOLED ? p.W / 2 : p.W;
  const long long npix = (long long)p.N * Hs * Ws;
  const int cblocks = (cgs + CG - 1) / CG;
  const int cb = blockIdx.x % cblocks;
  const int pb = blockIdx.x / cblocks, npb = gridDim.x / cblocks;
  const int c = (cb * CG + cgi) * 4;
  float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  if (c < p.C && pl < PL) {
    const long long per = (npix + npb - 1) / npb;
    long long e = (long long)(pb + 1) * per; if (e > npix) e = npix;
    for (long long pix = (long long)pb * per + pl; pix < e; pix += PL) {
      const int ws = (int)(pix % Ws); const long long t = pix / Ws;
      const int hs = (int)(t % Hs); const int n = (int)(t / Hs);
      float dz[POOLED ? 4 : 1][4], xh[POOLED ? 4 : 1][4]; long long rows[POOLED ? 4 : 1];
      bn_bwd_gather<POOLED>(p, n, hs, ws, c, dz, xh, rows);
#pragma unroll
      for (int q = 0; q < (POOLED ? 4 : 1); q++)
#pragma unroll
        for (int j = 0; j < 4; j++) { a1[j] += dz[q][j]; a2[j] += dz[q][j] * xh[q][j]; }
    }
  }
  const int CW = CG * 4;
  if (pl < PL) {
#pragma unroll
    for (int j = 0; j < 4; j++) { red[(0 * PL + pl) * CW + cgi * 4 + j] = a1[j]; red[(1 * PL + pl) * CW + cgi * 4 + j] = a2[j]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * CW; i += 256) {
    const int which = i / CW, cc = i % CW;
    double s = 0.0;
    for (int r = 0; r < PL; r++) s += (double)red[(which * PL + r) * CW + cc];
    const int ch = cb * CW + cc;
    if (ch < p.C) atomicAdd((which ? p.s2 : p.s1) + ch, s);
  }
}

template <bool POOLED>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdParams p) {
  const int cg = p.C >> 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Hs = POOLED ? p.H / 2 : p.H, Ws = POOLED ? p.W / 2 : p.W;
  const long long total = (long long)p.N * Hs * Ws * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 4;
  long long pix = idx / cg;
  const int ws = (int)(pix % Ws); pix /= Ws;
  const int hs = (int)(pix % Hs);
  const int n = (int)(pix / Hs);
  float dz[POOLED ? 4 : 1][4], xh[POOLED ? 4 : 1][4]; long long rows[POOLED ? 4 : 1];
  bn_bwd_gather<POOLED>(p, n, hs, ws, c, dz, xh, rows);
  float k1[4] = {0, 0, 0, 0}, k2[4] = {0, 0, 0, 0}, gs[4] = {1, 1, 1, 1};
  if (p.has_bn) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      k1[j] = (float)(p.s1[c + j] / p.count); k2[j] = (float)(p.s2[c + j] / p.count);
      gs[j] = p.gamma[c + j] * p.invstd[c + j];
    }
  }
#pragma unroll
  for (int q = 0; q < (POOLED ? 4 : 1); q++) {
    uint16_t o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) o[j] = cvt_f32_to_16(gs[j] * (dz[q][j] - k1[j] - xh[q][j] * k2[j]) * p.dy_scale, p.dy_fmt);
    *reinterpret_cast<uint2*>(p.dy + rows[q] * p.dy_ld + c) = make_uint2(o[0] | ((uint32_t)o[1] << 16), o[2] | ((uint32_t)o[3] << 16));
  }
}

// dgamma = S2, dbeta = S1 (accumulate into the gradient buffers), then clear S1/S2.  Launch AFTER bn_bwd_apply.
__global__ void bn_bwd_finalize_kernel(double* __restrict__ s1, double* __restrict__ s2, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int C, int accumulate, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = (float)(s2[c] * (double)scale), b = (float)(s1[c] * (double)scale);
  if (accumulate) { dgamma[c] += a; dbeta[c] += b; }
  else { dgamma[c] = a; dbeta[c] = b; }
  s1[c] = 0.0; s2[c] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// column sums of an NCHW fp32 tensor over (n, h, w): bias gradient of the linear head (conv 30)
__global__ void bias_grad_nchw_kernel(const float* __restrict__ g, float* __restrict__ db, int N, int C, int HW, int accumulate, float scale) {
  const int c = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < N * HW; i += blockDim.x) s += (double)g[((long long)(i / HW) * C + c) * HW + (i % HW)];
  __shared__ double sm[256];
  sm[threadIdx.x] = s; __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) { const float v = (float)(sm[0] * (double)scale); if (accumulate) db[c] += v; else db[c] = v; }
}

// ------------------------------------------------------------------------------------------------
// Weight re-packing.  Master weights: fp32 [cout][taps][cin] (the memory behind the permuted nn.Conv2d.weight view).
//   fwd  : hi/lo fp16 [cout][ld_f]      k = tap*cin + ci                 (B operand of the forward GEMM)
//   dgrad: 16-bit     [cin][ld_d]       k = tap'*cout + co, tap' = taps-1-tap   (B operand of the data-gradient GEMM)
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int taps, int cin,
                                    uint16_t* __restrict__ f_hi, uint16_t* __restrict__ f_lo, int ld_f,
                                    uint16_t* __restrict__ d, int ld_d, int d_fmt) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)cout * taps * cin;
  if (idx >= total) return;
  const int ci = (int)(idx % cin);
  const int tap = (int)((idx / cin) % taps);
  const int co = (int)(idx / ((long long)cin * taps));
  const float v = w[idx];
  if (f_hi) {
    uint16_t a, b; split_f16(v, a, b);
    const long long o = (long long)co * ld_f + tap * cin + ci;
    f_hi[o] = a; if (f_lo) f_lo[o] = b;
  }
  if (d) d[(long long)ci * ld_d + (long long)(taps - 1 - tap) * cout + co] = cvt_f32_to_16(v, d_fmt);
}

// ------------------------------------------------------------------------------------------------
// optim.SGD(momentum, dampening=0, weight_decay) over one flat buffer (train.py:388):
//   g += wd*p ; v = mu*v + g ; p -= lr*v      (first step of torch: v = g, identical with v0 = 0)
__global__ void sgd_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, long long n,
                                float lr, float mu, float wd, float gscale) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 + 3 < n) {
    float4 pp = *reinterpret_cast<float4*>(p + i4);
    const float4 gg = *reinterpret_cast<const float4*>(g + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    vv.x = mu * vv.x + (gg.x * gscale + wd * pp.x); vv.y = mu * vv.y + (gg.y * gscale + wd * pp.y);
    vv.z = mu * vv.z + (gg.z * gscale + wd * pp.z); vv.w = mu * vv.w + (gg.w * gscale + wd * pp.w);
    pp.x -= lr * vv.x; pp.y -= lr * vv.y; pp.z -= lr * vv.z; pp.w -= lr * vv.w;
    *reinterpret_cast<float4*>(v + i4) = vv;
    *reinterpret_cast<float4*>(p + i4) = pp;
  } else {
    for (long long i = i4; i < n; i++) {
      const float vn = mu * v[i] + (g[i] * gscale + wd * p[i]);
      v[i] = vn; p[i] -= lr * vn;
    }
  }
}

// ================================================================================================ host launchers
static inline unsigned nblk(long long total, int bs) { return (unsigned)((total + bs - 1) / bs); }

int pack_input_im2col(const float* x, void* hi, void* lo, int N, int H, int W, cudaStream_t s) {
  if (!x || !hi || !lo) return fail_msg(SSP_ERR_ARG, "pack_input_im2col: null pointer");
  const long long total = (long long)N * H * W * 32;
  pack_input_im2col_kernel<<<nblk(total, 256), 256, 0, s>>>(x, (uint16_t*)hi, (uint16_t*)lo, N, H, W);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int pack_nchw(const float* x, void* hi, void* lo, int N, int C, int H, int W, int ld, int c0, int fmt, float scale, cudaStream_t s) {
  if (!x || !hi) return fail_msg(SSP_ERR_ARG, "pack_nchw: null pointer");
  const long long total = (long long)N * C * H * W;
  pack_nchw_kernel<<<nblk(total, 256), 256, 0, s>>>(x, (uint16_t*)hi, (uint16_t*)lo, N, C, H, W, ld, c0, fmt, scale);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int unpack_nchw(const float* y, float* out, int N, int C, int H, int W, int ld, int c0, cudaStream_t s) {
  if (!y || !out) return fail_msg(SSP_ERR_ARG, "unpack_nchw: null pointer");
  const long long total = (long long)N * C * H * W;
  unpack_nchw_kernel<<<nblk(total, 256), 256, 0, s>>>(y, out, N, C, H, W, ld, c0);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int unpack16_nchw(const void* hi, const void* lo, float* out, int N, int C, int H, int W, int ld, int c0, int fmt, cudaStream_t s) {
  if (!hi || !out) return fail_msg(SSP_ERR_ARG, "unpack16_nchw: null pointer");
  const long long total = (long long)N * C * H * W;
  unpack16_nchw_kernel<<<nblk(total, 256), 256, 0, s>>>((const uint16_t*)hi, (const uint16_t*)lo, out, N, C, H, W, ld, c0, fmt);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int bn_finalize(double* ssum, double* ssq, double count, const float* gamma, const float* beta, float* rm, float* rv,
                float momentum, float eps, int train, float* mean, float* invstd, float* scale, float* shift, int C, cudaStream_t s) {
  if (!gamma || !beta || !mean || !invstd || !scale || !shift || (train && (!ssum || !ssq)) || (!train && (!rm || !rv)))
    return fail_msg(SSP_ERR_ARG, "bn_finalize: null pointer");
  bn_finalize_kernel<<<nblk(C, 128), 128, 0, s>>>(ssum, ssq, count, gamma, beta, rm, rv, momentum, eps, train, mean, invstd, scale, shift, C);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int bn_apply(const float* y, int y_ld, const float* scale, const float* shift, int N, int C, int H, int W, float slope,
             void* d0_hi, void* d0_lo, int d0_ld, int d0_c0, int d0_kind,
             void* d1_hi, void* d1_lo, int d1_ld, int d1_c0, int d1_kind, cudaStream_t s) {
  if (!y || !scale || !shift || (C % 4)) return fail_msg(SSP_ERR_ARG, "bn_apply: bad argument (C must be a multiple of 4)");
  BnApplyParams p;
  p.y = y; p.y_ld = y_ld; p.scale = scale; p.shift = shift; p.N = N; p.C = C; p.H = H; p.W = W; p.slope = slope;
  p.dst[0] = ActDst{(uint16_t*)d0_hi, (uint16_t*)d0_lo, d0_ld, d0_c0, d0_hi ? d0_kind : DST_NONE};
  p.dst[1] = ActDst{(uint16_t*)d1_hi, (uint16_t*)d1_lo, d1_ld, d1_c0, d1_hi ? d1_kind : DST_NONE};
  const bool pooled = p.dst[0].kind == DST_POOL || p.dst[1].kind == DST_POOL;
  const bool halves = pooled || p.dst[0].kind == DST_REORG || p.dst[1].kind == DST_REORG;
  if (halves && ((H | W) & 1)) return fail_msg(SSP_ERR_ARG, "bn_apply: pool/reorg need even H and W");
  if (pooled) {
    const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
    bn_apply_kernel<true><<<nblk(total, 256), 256, 0, s>>>(p);
  } else {
    const long long total = (long long)N * H * W * (C / 4);
    bn_apply_kernel<false><<<nblk(total, 256), 256, 0, s>>>(p);
  }
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

static int fill_bwd(BnBwdParams& p, const float* y, int y_ld, const float* scale, const float* shift, const float* mean,
                    const float* invstd, const float* gamma, int N, int C, int H, int W, float slope,
                    const float* g0, int g0_ld, int g0_c0, int g0_kind, const float* g1, int g1_ld, int g1_c0, int g1_kind,
                    double* s1, double* s2) {
  if (!y || (C % 4) || !g0) return SSP_ERR_ARG;
  p.y = y; p.y_ld = y_ld; p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd; p.gamma = gamma;
  p.has_bn = (scale && shift && mean && invstd && gamma) ? 1 : 0;
  p.N = N; p.C = C; p.H = H; p.W = W; p.slope = slope;
  p.src[0] = GradSrc{g0, g0_ld, g0_c0, g0_kind};
  p.src[1] = GradSrc{g1, g1_ld, g1_c0, g1 ? g1_kind : SRC_NONE};
  p.s1 = s1; p.s2 = s2; p.count = (double)N * H * W;
  p.dy = nullptr; p.dy_ld = 0; p.dy_fmt = 0; p.dy_scale = 1.f;
  return SSP_OK;
}
int bn_bwd_reduce(const float* y, int y_ld, const float* scale, const float* shift, const float* mean, const float* invstd,
                  const float* gamma, int N, int C, int H, int W, float slope,
                  const float* g0, int g0_ld, int g0_c0, int g0_kind, const float* g1, int g1_ld, int g1_c0, int g1_kind,
                  double* s1, double* s2, cudaStream_t s) {
  BnBwdParams p;
  if (fill_bwd(p, y, y_ld, scale, shift, mean, invstd, gamma, N, C, H, W, slope, g0, g0_ld, g0_c0, g0_kind, g1, g1_ld, g1_c0, g1_kind, s1, s2) || !s1 || !s2 || !p.has_bn)
    return fail_msg(SSP_ERR_ARG, "bn_bwd_reduce: bad argument");
  const bool pooled = p.src[0].kind == SRC_POOL || p.src[1].kind == SRC_POOL;
  const int cgs = C / 4, CG = cgs < 256 ? cgs : 256, PL = 256 / CG;
  const int cblocks = (cgs + CG - 1) / CG;
  const long long npix = (long long)N * (pooled ? H / 2 : H) * (pooled ? W / 2 : W);
  long long npb = (npix + (long long)PL * 16 - 1) / ((long long)PL * 16);
  const long long cap = 148 * 8 / cblocks; if (npb > cap) npb = cap; if (npb < 1) npb = 1;
  const size_t sm = (size_t)2 * PL * CG * 4 * sizeof(float);
  if (pooled) bn_bwd_reduce_kernel<true><<<(unsigned)(npb * cblocks), 256, sm, s>>>(p);
  else bn_bwd_reduce_kernel<false><<<(unsigned)(npb * cblocks), 256, sm, s>>>(p);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int bn_bwd_apply(const float* y, int y_ld, const float* scale, const float* shift, const float* mean, const float* invstd,
                 const float* gamma, int N, int C, int H, int W, float slope,
                 const float* g0, int g0_ld, int g0_c0, int g0_kind, const float* g1, int g1_ld, int g1_c0, int g1_kind,
                 double* s1, double* s2, void* dy, int dy_ld, int dy_fmt, float dy_scale, cudaStream_t s) {
  BnBwdParams p;
  if (fill_bwd(p, y, y_ld, scale, shift, mean, invstd, gamma, N, C, H, W, slope, g0, g0_ld, g0_c0, g0_kind, g1, g1_ld, g1_c0, g1_kind, s1, s2) || !dy)
    return fail_msg(SSP_ERR_ARG, "bn_bwd_apply: bad argument");
  if (p.has_bn && (!s1 || !s2)) return fail_msg(SSP_ERR_ARG, "bn_bwd_apply: statistics buffers missing");
  p.dy = (uint16_t*)dy; p.dy_ld = dy_ld; p.dy_fmt = dy_fmt; p.dy_scale = dy_scale;
  const bool pooled = p.src[0].kind == SRC_POOL || p.src[1].kind == SRC_POOL;
  const long long total = (long long)N * (pooled ? H / 2 : H) * (pooled ? W / 2 : W) * (C / 4);
  if (pooled) bn_bwd_apply_kernel<true><<<nblk(total, 256), 256, 0, s>>>(p);
  else bn_bwd_apply_kernel<false><<<nblk(total, 256), 256, 0, s>>>(p);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int bn_bwd_finalize(double* s1, double* s2, float* dgamma, float* dbeta, int C, int accumulate, float scale, cudaStream_t s) {
  if (!s1 || !s2 || !dgamma || !dbeta) return fail_msg(SSP_ERR_ARG, "bn_bwd_finalize: null pointer");
  bn_bwd_finalize_kernel<<<nblk(C, 128), 128, 0, s>>>(s1, s2, dgamma, dbeta, C, accumulate, scale);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int bias_grad_nchw(const float* g, float* db, int N, int C, int HW, int accumulate, float scale, cudaStream_t s) {
  if (!g || !db) return fail_msg(SSP_ERR_ARG, "bias_grad_nchw: null pointer");
  bias_grad_nchw_kernel<<<C, 256, 0, s>>>(g, db, N, C, HW, accumulate, scale);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int pack_weights(const float* w, int cout, int taps, int cin, void* f_hi, void* f_lo, int ld_f, void* d, int ld_d, int d_fmt, cudaStream_t s) {
  if (!w) return fail_msg(SSP_ERR_ARG, "pack_weights: null pointer");
  const long long total = (long long)cout * taps * cin;
  pack_weights_kernel<<<nblk(total, 256), 256, 0, s>>>(w, cout, taps, cin, (uint16_t*)f_hi, (uint16_t*)f_lo, ld_f, (uint16_t*)d, ld_d, d_fmt);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}
int sgd_step_flat(float* p, const float* g, float* v, long long n, float lr, float mu, float wd, float gscale, cudaStream_t s) {
  if (!p || !g || !v) return fail_msg(SSP_ERR_ARG, "sgd_step_flat: null pointer");
  sgd_flat_kernel<<<nblk((n + 3) / 4, 256), 256, 0, s>>>(p, g, v, n, lr, mu, wd, gscale);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

}  // namespace ssp
