// Host harness: compiles singleshotpose_b200/csrc/augment_core.h (the arithmetic AND the drivers the CUDA kernels use) with
// g++ and runs it with plain loops, so that the CPU test-suite can check it bit-exactly against Pillow and the oracle.
// Test infrastructure: built by tests/test_augment_host.py into a temporary .so; never loaded by the product.
#include "../../singleshotpose_b200/csrc/augment_core.h"
#include <string.h>

using namespace ssp_aug;

namespace {
struct HostBackend {
  void coeffs(int in_size, int in0, int in1, int out_size, int resample, int ksize, int* bounds, int* kk) {
    for (int xx = 0; xx < out_size; xx++) coeff_row(in_size, in0, in1, out_size, resample, ksize, xx, bounds + 2 * xx, kk + (long long)xx * ksize);
  }
  void pass(const PassArgs& a) {
    for (int y = 0; y < a.dst_h; y++) for (int x = 0; x < a.dst_w; x++) resample_pass_px(a, x, y);
  }
  void nearest(const PassArgs& a) {
    for (int y = 0; y < a.dst_h; y++) for (int x = 0; x < a.dst_w; x++) nearest_px(a, x, y);
  }
  void composite(const uint8_t* img, const uint8_t* bg, const uint8_t* mask, const uint8_t* lp, const uint8_t* ln, long long n, uint8_t* out) {
    for (long long i = 0; i < n; i++) out[i] = composite_px(img[i], bg[i], mask[i], lp, ln);
  }
  void distort(const uint8_t* src, int w, int h, const uint8_t* luts, uint8_t* out_u8, float* out_chw) {
    const long long n = (long long)w * h;
    for (long long i = 0; i < n; i++) {
      uint8_t o[3];
      distort_px(src + 3 * i, luts, luts + 256, luts + 512, o);
      if (out_u8) { out_u8[3 * i] = o[0]; out_u8[3 * i + 1] = o[1]; out_u8[3 * i + 2] = o[2]; }
      if (out_chw) for (int c = 0; c < 3; c++) out_chw[c * n + i] = (float)o[c] / 255.0f;
    }
  }
};
}  // namespace

extern "C" {
void h_rgb2hsv(const uint8_t* in, uint8_t* out, long long n) { for (long long i = 0; i < n; i++) rgb2hsv_px(in[3 * i], in[3 * i + 1], in[3 * i + 2], out + 3 * i); }
void h_hsv2rgb(const uint8_t* in, uint8_t* out, long long n) { for (long long i = 0; i < n; i++) hsv2rgb_px(in[3 * i], in[3 * i + 1], in[3 * i + 2], out + 3 * i); }
long long h_resize_work_bytes(int in_w, int in_h, int out_w, int out_h, int resample) { return resize_work_bytes(in_w, in_h, out_w, out_h, resample); }
int h_resize(const uint8_t* src, int src_w, int src_h, int x0, int y0, int in_w, int in_h, uint8_t* dst, int out_w, int out_h, int resample,
             uint8_t* work, long long work_bytes) {
  HostBackend be;
  return resize_u8_driver(be, src, src_w, src_h, x0, y0, in_w, in_h, dst, out_w, out_h, resample, work, work_bytes);
}
long long h_augment_work_bytes(int ow, int oh, int bw, int bh, int cw, int ch, int out_w, int out_h, int resample) {
  return augment_work_bytes(ow, oh, bw, bh, cw, ch, out_w, out_h, resample);
}
int h_augment_sample(const uint8_t* img, const uint8_t* mask, int ow, int oh, const uint8_t* bg, int bw, int bh, const uint8_t* luts,
                     int pleft, int ptop, int cw, int ch, int out_w, int out_h, int resample, uint8_t* work, long long work_bytes,
                     uint8_t* out_u8, float* out_chw) {
  HostBackend be;
  return augment_sample_driver(be, img, mask, ow, oh, bg, bw, bh, luts, pleft, ptop, cw, ch, out_w, out_h, resample, work, work_bytes, out_u8, out_chw);
}
// batched path: plan with the recording back end, then execute the op table stage by stage with plain loops
long long h_aug_op_bytes() { return (long long)sizeof(AugOp); }
long long h_aug_item_bytes() { return (long long)sizeof(AugItem); }
int h_augment_batch(const AugItem* items, int n, int out_w, int out_h, int resample, AugOp* table, int* stage_dims) {
  const int rc = augment_batch_plan(items, n, out_w, out_h, resample, table, stage_dims);
  if (rc) return rc;
  for (int s = 0; s < kMaxStages; s++)
    for (int i = 0; i < n; i++) {
      const AugOp& o = table[(long long)s * n + i];
      for (int y = 0; y < o.ny; y++) for (int x = 0; x < o.nx; x++) op_element(o, x, y);
    }
  return 0;
}
}
