// Multi-object RegionLoss head and decode (yolo-pose-multi.cfg: 5 anchors, 13 classes).
// Restates reference multi_obj_pose_estimation/region_loss_multi.py:9-189 (build_targets with IoU anchor choice, the
// masked MSE terms, CrossEntropy(sum) on the class logits) and utils_multi.py:266-382 (get_multi_region_boxes).
// The reference's "best_n = -1" read (region_loss_multi.py:51,63: tconf is computed from the LAST anchor of the
// PREVIOUS image at the ground-truth cell, wrapping to the last image for b = 0) is reproduced on purpose.
#include "ssp_common.cuh"

namespace ssp {

#define SSPM_MAX_KP 16
#define SSPM_MAX_GT 50
#define SSPM_MAX_ANCHORS 16

__device__ __forceinline__ float sigmoidm_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float corner_conf_m(const float* gt, const float* px, const float* py, int K, float eps) {
  const float conf0 = expf(2.f) - 1.f + eps;
  float s = 0.f;
  for (int k = 0; k < K; k++) {
    const float dx = (gt[2 * k] - px[k]) * 640.f, dy = (gt[2 * k + 1] - py[k]) * 480.f;
    const float d = sqrtf(dx * dx + dy * dy);
    if (d < 80.f) s += (expf(2.f * (1.f - d / 80.f)) - 1.f) / conf0;
  }
  return s / (float)K;
}

struct RegionMultiParams {
  const float* out; const float* target; float* grad; double* acc;
  int B, K, nC, nA, H, W;
  float anchors[2 * SSPM_MAX_ANCHORS]; int anchor_step;
  float coord_scale, noobject_scale, object_scale, class_scale, thresh;
  int use_conf; float grad_scale;
};

__global__ void __launch_bounds__(256) region_loss_multi_kernel(const RegionMultiParams p) {
  const int b = blockIdx.x, K = p.K, HW = p.H * p.W, nch = 2 * K + 1 + p.nC, nl = 2 * K + 3;
  __shared__ float s_gt[SSPM_MAX_GT][2 * SSPM_MAX_KP];
  __shared__ float s_tconf[SSPM_MAX_GT];
  __shared__ int s_cell[SSPM_MAX_GT], s_anchor[SSPM_MAX_GT], s_cls[SSPM_MAX_GT], s_valid[SSPM_MAX_GT];
  __shared__ int s_n;
  __shared__ double sred[8][8];
  const float* t = p.target + (long long)b * SSPM_MAX_GT * nl;
  if (threadIdx.x == 0) {
    int n = 0;
    while (n < SSPM_MAX_GT && t[n * nl + 1] != 0.f) n++;       // the list ends at the first x0 == 0 (region_loss_multi.py:31,47)
    s_n = n;
  }
  __syncthreads();
  const int nG = s_n;
  const float* o = p.out + (long long)b * p.nA * nch * HW;
  if (threadIdx.x < nG) {
    const int g = threadIdx.x;
    const float* tg = t + g * nl;
    for (int j = 0; j < 2 * K; j++) s_gt[g][j] = tg[1 + j];
    const int gi0 = (int)(tg[1] * p.W), gj0 = (int)(tg[2] * p.H);
    const bool ok = gi0 >= 0 && gi0 < p.W && gj0 >= 0 && gj0 < p.H;
    s_valid[g] = ok; s_cell[g] = gj0 * p.W + gi0; s_cls[g] = (int)tg[0];
    // anchor by IoU of (gw, gh) against the anchor boxes, both centred at the origin
    const float gw = tg[nl - 2] * p.W, gh = tg[nl - 1] * p.H;
    float best = 0.f; int bn = -1;
    for (int n = 0; n < p.nA; n++) {
      const float aw = p.anchors[p.anchor_step * n], ah = p.anchors[p.anchor_step * n + 1];
      const float mx = fminf(-aw / 2.f, -gw / 2.f), Mx = fmaxf(aw / 2.f, gw / 2.f);
      const float my = fminf(-ah / 2.f, -gh / 2.f), My = fmaxf(ah / 2.f, gh / 2.f);
      const float cw = aw + gw - (Mx - mx), ch = ah + gh - (My - my);
      float iou = 0.f;
      if (cw > 0.f && ch > 0.f) { const float ca = cw * ch; iou = ca / (aw * ah + gw * gh - ca); }
      if (iou > best) { best = iou; bn = n; }
    }
    s_anchor[g] = bn < 0 ? p.nA - 1 : bn;                        // python's [-1] indexing when no anchor overlaps
    float tc = 0.f;
    if (ok) {                                                    // prediction of image b-1 (wrapping), last anchor, same cell
      const int pb = (b + p.B - 1) % p.B;
      const float* op = p.out + ((long long)pb * p.nA + (p.nA - 1)) * nch * HW + s_cell[g];
      float px[SSPM_MAX_KP], py[SSPM_MAX_KP];
      for (int k = 0; k < K; k++) {
        float vx = op[(2 * k) * HW], vy = op[(2 * k + 1) * HW];
        if (k == 0) { vx = sigmoidm_(vx); vy = sigmoidm_(vy); }
        px[k] = (vx + (float)gi0) / (float)p.W; py[k] = (vy + (float)gj0) / (float)p.H;
      }
      tc = corner_conf_m(s_gt[g], px, py, K, 1e-5f);
    }
    s_tconf[g] = tc;
  }
  __syncthreads();
  double part[7] = {0, 0, 0, 0, 0, 0, 0};
  if (threadIdx.x < nG && s_valid[threadIdx.x]) { part[3] = 1.0; if (s_tconf[threadIdx.x] > 0.5f) part[4] = 1.0; }
  float* g = p.grad ? p.grad + (long long)b * p.nA * nch * HW : nullptr;
  for (int i = threadIdx.x; i < p.nA * HW; i += blockDim.x) {
    const int a = i / HW, cell = i % HW, cy = cell / p.W, cx = cell % p.W;
    const float* oa = o + (long long)a * nch * HW + cell;
    float* ga = g ? g + (long long)a * nch * HW + cell : nullptr;
    float xs[SSPM_MAX_KP], ys[SSPM_MAX_KP], px[SSPM_MAX_KP], py[SSPM_MAX_KP];
    for (int k = 0; k < K; k++) {
      float vx = oa[(2 * k) * HW], vy = oa[(2 * k + 1) * HW];
      if (k == 0) { vx = sigmoidm_(vx); vy = sigmoidm_(vy); }
      xs[k] = vx; ys[k] = vy;
      px[k] = (vx + (float)cx) / (float)p.W; py[k] = (vy + (float)cy) / (float)p.H;
    }
    const float conf = sigmoidm_(oa[(2 * K) * HW]);
    if (conf > 0.25f) part[5] += 1.0;
    float conf_mask = p.noobject_scale, tconf = 0.f;
    int sel = -1;
    for (int gidx = 0; gidx < nG; gidx++) {
      if (corner_conf_m(s_gt[gidx], px, py, K, 0.f) > p.thresh) conf_mask = 0.f;
      if (s_valid[gidx] && s_anchor[gidx] == a && s_cell[gidx] == cell) sel = gidx;      // the LAST ground truth on this slot wins
    }
    if (sel >= 0) { conf_mask = p.object_scale; tconf = s_tconf[sel]; }
    for (int k = 0; k < K; k++) {
      float gx = 0.f, gy = 0.f;
      if (sel >= 0) {
        const float tx = s_gt[sel][2 * k] * (float)p.W - (float)cx, ty = s_gt[sel][2 * k + 1] * (float)p.H - (float)cy;
        const float ex = xs[k] - tx, ey = ys[k] - ty;
        part[0] += 0.5 * (double)p.coord_scale * (double)(ex * ex);
        part[1] += 0.5 * (double)p.coord_scale * (double)(ey * ey);
        gx = p.coord_scale * ex; gy = p.coord_scale * ey;
        if (k == 0) { gx *= xs[0] * (1.f - xs[0]); gy *= ys[0] * (1.f - ys[0]); }
      }
      if (ga) { ga[(2 * k) * HW] = gx * p.grad_scale; ga[(2 * k + 1) * HW] = gy * p.grad_scale; }
    }
    const float ec = conf - tconf;
    part[2] += 0.5 * (double)conf_mask * (double)(ec * ec);
    if (ga) ga[(2 * K) * HW] = p.use_conf ? conf_mask * ec * conf * (1.f - conf) * p.grad_scale : 0.f;
    // class term: CrossEntropyLoss(sum) over the selected slots
    if (sel >= 0) {
      float mx = -INFINITY;
      for (int c = 0; c < p.nC; c++) mx = fmaxf(mx, oa[(2 * K + 1 + c) * HW]);
      float den = 0.f;
      for (int c = 0; c < p.nC; c++) den += expf(oa[(2 * K + 1 + c) * HW] - mx);
      const int tc = s_cls[sel];
      const float lt = (tc >= 0 && tc < p.nC) ? oa[(2 * K + 1 + tc) * HW] : 0.f;
      part[6] += (double)p.class_scale * (double)(logf(den) + mx - lt);
      if (ga) for (int c = 0; c < p.nC; c++)
        ga[(2 * K + 1 + c) * HW] = p.class_scale * (expf(oa[(2 * K + 1 + c) * HW] - mx) / den - (c == tc ? 1.f : 0.f)) * p.grad_scale;
    } else if (ga) {
      for (int c = 0; c < p.nC; c++) ga[(2 * K + 1 + c) * HW] = 0.f;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 7; j++) {
    double v = part[j];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) sred[j][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += sred[threadIdx.x][w];
    if (v != 0.0) atomicAdd(p.acc + threadIdx.x, v);
  }
}

// ------------------------------------------------------------------------------------------------ decode
// dense pass: for every (image, cell, anchor) -- cell-major, anchor fastest, the reference's visiting order -- the box
// [x0/w, y0/h, ..., det_conf, cls_max_conf, cls_max_id], the selection confidence and softmax[correspondingclass]
__global__ void __launch_bounds__(256) region_decode_multi_kernel(const float* __restrict__ out, int B, int K, int nC, int nA, int H, int W,
                                                                  int only_objectness, int corr, float* __restrict__ boxes,
                                                                  float* __restrict__ conf_sel, float* __restrict__ det, float* __restrict__ cls_corr) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W, nch = 2 * K + 1 + nC;
  if (idx >= (long long)B * HW * nA) return;
  const int a = (int)(idx % nA); const int cell = (int)((idx / nA) % HW); const int b = (int)(idx / ((long long)nA * HW));
  const int cy = cell / W, cx = cell % W;
  const float* o = out + ((long long)b * nA + a) * nch * HW + cell;
  float* bx = boxes + idx * (2 * K + 3);
  for (int k = 0; k < K; k++) {
    float vx = o[(2 * k) * HW], vy = o[(2 * k + 1) * HW];
    if (k == 0) { vx = sigmoidm_(vx); vy = sigmoidm_(vy); }
    bx[2 * k] = (vx + (float)cx) / (float)W; bx[2 * k + 1] = (vy + (float)cy) / (float)H;
  }
  const float dc = sigmoidm_(o[(2 * K) * HW]);
  float mx = -INFINITY; int id = 0;
  for (int c = 0; c < nC; c++) { const float v = o[(2 * K + 1 + c) * HW]; if (v > mx) { mx = v; id = c; } }
  float den = 0.f;
  for (int c = 0; c < nC; c++) den += expf(o[(2 * K + 1 + c) * HW] - mx);
  const float cmax = 1.f / den;
  bx[2 * K] = dc; bx[2 * K + 1] = cmax; bx[2 * K + 2] = (float)id;
  conf_sel[idx] = only_objectness ? dc : dc * cmax;
  det[idx] = dc;
  cls_corr[idx] = (corr >= 0 && corr < nC) ? expf(o[(2 * K + 1 + corr) * HW] - mx) / den : 0.f;
}

// the reference's running maxima (max_conf reset per image, max_cls_conf and max_ind never reset): inherently sequential
__global__ void region_decode_multi_fallback_kernel(const float* __restrict__ det, const float* __restrict__ cls_corr, int B, int per_image,
                                                    long long* __restrict__ max_ind, float* __restrict__ max_conf, float* __restrict__ max_cls) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float mcls = -INFINITY; long long mind = -1;
  for (int b = 0; b < B; b++) {
    float mconf = -1.f;
    for (int i = 0; i < per_image; i++) {
      const long long idx = (long long)b * per_image + i;
      if (det[idx] > mconf && cls_corr[idx] > mcls) { mconf = det[idx]; mcls = cls_corr[idx]; mind = idx; }
    }
    max_ind[b] = mind; max_conf[b] = mconf; max_cls[b] = mcls;
  }
}

int region_loss_multi_fwd_bwd(const float* out, const float* target, float* grad, double* acc, int B, int K, int nC, int nA, int H, int W,
                              const float* anchors_host, int anchor_step, float coord_scale, float noobject_scale, float object_scale,
                              float class_scale, float thresh, int use_conf, float grad_scale, cudaStream_t s) {
  if (!out || !target || !acc || !anchors_host || K < 1 || K > SSPM_MAX_KP || nA < 1 || nA > SSPM_MAX_ANCHORS || anchor_step < 2)
    return fail_msg(SSP_ERR_ARG, "region_loss_multi_fwd_bwd: bad argument");
  cudaError_t e = cudaMemsetAsync(acc, 0, 8 * sizeof(double), s);
  if (e != cudaSuccess) return fail_cuda(e, __FILE__, __LINE__);
  RegionMultiParams p;
  p.out = out; p.target = target; p.grad = grad; p.acc = acc; p.B = B; p.K = K; p.nC = nC; p.nA = nA; p.H = H; p.W = W;
  for (int i = 0; i < 2 * SSPM_MAX_ANCHORS; i++) p.anchors[i] = (i < nA * anchor_step && i < 2 * SSPM_MAX_ANCHORS) ? anchors_host[i] : 0.f;
  p.anchor_step = anchor_step; p.coord_scale = coord_scale; p.noobject_scale = noobject_scale; p.object_scale = object_scale;
  p.class_scale = class_scale; p.thresh = thresh; p.use_conf = use_conf; p.grad_scale = grad_scale;
  region_loss_multi_kernel<<<B, 256, 0, s>>>(p);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

int region_decode_multi(const float* out, int B, int K, int nC, int nA, int H, int W, int only_objectness, int corr, float* boxes,
                        float* conf_sel, float* det, float* cls_corr, long long* max_ind, float* max_conf, float* max_cls, cudaStream_t s) {
  if (!out || !boxes || !conf_sel || !det || !cls_corr || !max_ind || !max_conf || !max_cls || K < 1 || K > SSPM_MAX_KP)
    return fail_msg(SSP_ERR_ARG, "region_decode_multi: bad argument");
  const long long total = (long long)B * H * W * nA;
  region_decode_multi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(out, B, K, nC, nA, H, W, only_objectness, corr, boxes, conf_sel, det, cls_corr);
  SSP_CHECK_LAUNCH();
  region_decode_multi_fallback_kernel<<<1, 32, 0, s>>>(det, cls_corr, B, H * W * nA, max_ind, max_conf, max_cls);
  SSP_CHECK_LAUNCH(); return SSP_OK;
}

}  // namespace ssp
