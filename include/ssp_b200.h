/* ssp_b200.h -- C ABI of libssp_b200.so: the sm_100a kernels behind the singleshotpose hot path.
 *
 * The reference (microsoft/singleshotpose) has no FFI: its hot path is the Python surface
 * Darknet.forward / RegionLoss.forward / get_region_boxes / pnp.  Each entry point below names the
 * reference code it replaces (file:line under /root/reference); singleshotpose_b200/*.py are the
 * thin Python mirrors of that surface that bind these symbols with ctypes.
 *
 * Conventions: plain pointers and sizes only; every pointer is a DEVICE pointer unless stated; all
 * functions are asynchronous on `stream` (a cudaStream_t passed as void*), never allocate, never
 * synchronise, and return 0 on success or a negative SSP_ERR_* code (ssp_last_error() gives the text).
 * Thread-compatible: distinct host threads may call concurrently on distinct streams.
 *
 * Activation layout ("padded-flat NHWC"): a (N,C,H,W) feature map is a row-major matrix [rows][ld] with
 *   row(n,h,w) = n*(H+1)*(W+1) + (h+1)*(W+1) + (w+1)
 * (one shared zero pad pixel per image row, one shared zero pad row per image); buffers hold
 * ssp_flat_alloc_rows(N,H,W) rows and must be zero-initialised once: kernels only ever write valid rows
 * of operand planes, the zero pads are what makes a 3x3 tap a constant row shift.
 * 16-bit operand planes come as fp16 "hi" (+ optional fp16 "lo", value = hi + lo) -- see DESIGN.md numerics.
 */
#ifndef SSP_B200_H
#define SSP_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define SSP_OK 0
#define SSP_ERR_ARG (-1)
#define SSP_ERR_CUDA (-2)
#define SSP_ERR_DRIVER (-3)

#define SSP_FMT_F16 0
#define SSP_FMT_BF16 1
#define SSP_IMPL_TC 0    /* tcgen05 tensor-core kernel */
#define SSP_IMPL_SIMT 1  /* fp32 CUDA-core kernel (cross-check / bring-up) */
#define SSP_IMPL_BAND 3  /* narrow 3x3 layers: one activation band per kernel row + resident weights (falls back to TC) */
#define SSP_IMPL_TC2 2   /* tcgen05 cta_group::2 kernel: CTA pairs share the weight tile (ssp_conv_gemm only) */
#define SSP_IMPL_BANDT 4 /* few output channels (<= 64 split-fp16, <= 128 single-term): operands swapped, weights on the M side, 128/256 pixels as N (csrc/conv_bandt.cu; falls back to BAND / TC2 / TC when not eligible) */
#define SSP_EPI_F32 0    /* store fp32 */
#define SSP_EPI_STATS 1  /* store fp32 + per-channel sum / sum of squares over valid pixels (fp64) */
#define SSP_EPI_BIAS 2   /* add bias, store fp32 */
#define SSP_EPI_F16 8    /* store fp16 (saturating): `out` points to 16-bit elements, out_ld in elements; data gradients only (TC2 / BANDT kernels) */
#define SSP_ROUTE_NONE 0
#define SSP_ROUTE_DIRECT 1 /* consumer has the same geometry */
#define SSP_ROUTE_POOL 2   /* consumer is behind MaxPool2d(2,2)          (darknet.py:168-176) */
#define SSP_ROUTE_REORG 3  /* consumer is behind Reorg(2), marvis order  (darknet.py:16-35)   */
#define SSP_ROUTE_F16 16   /* OR-ed into a g*_route of ssp_bn_bwd_*: that upstream gradient plane holds fp16 (written with SSP_EPI_F16) instead of fp32 */

int ssp_version(void);
const char* ssp_last_error(void);              /* host pointer, thread-local text of the last failure */
long long ssp_flat_alloc_rows(int N, int H, int W);
long long ssp_flat_row(int n, int h, int w, int H, int W);

/* ---- layout: train.py:83 `data.cuda()` hands NCHW fp32; the conv stack runs on padded-flat rows ---- */
int ssp_pack_input_im2col(const float* x_nchw, void* hi, void* lo, int N, int H, int W, void* stream);
int ssp_pack_nchw(const float* x_nchw, void* hi, void* lo_or_null, int N, int C, int H, int W, int ld, int c0,
                  int fmt, float scale, void* stream);
int ssp_unpack_nchw(const float* y_flat, float* out_nchw, int N, int C, int H, int W, int ld, int c0, void* stream);
int ssp_unpack16_nchw(const void* hi, const void* lo_or_null, float* out_nchw, int N, int C, int H, int W, int ld,
                      int c0, int fmt, void* stream);

/* ---- nn.Conv2d forward and data gradient (darknet.py:156-160; autograd of train.py:103) ----
 * out[m][n] = sum_tap sum_c A[m + shift(tap)][c] * B[n][tap*cin + c],  taps in {1, 9} (3x3 pad 1 / 1x1). */
int ssp_conv_gemm(int impl, const void* a_hi, const void* a_lo_or_null, long long a_rows, int a_ld, int cin,
                  const void* b_hi, const void* b_lo_or_null, int b_rows, int b_ld, int a_fmt, int b_fmt,
                  int N, int H, int W, int taps, int cout, float* out, int out_ld, long long out_rows, int epi,
                  const float* bias, double* stat_sum, double* stat_sq, void* stream);
/* number of launches of the operand-swapped kernel (SSP_IMPL_BANDT) so far in this process: lets a test tell the kernel from its fall-backs */
int ssp_conv_bandt_launches(void);
/* ---- inference: nn.Conv2d + nn.BatchNorm2d(eval) + nn.LeakyReLU as ONE kernel (darknet.py:154-164 under model.eval()):
 *      z = leaky(conv * scale[c] + shift[c]) is written by the GEMM epilogue straight into the consumer's fp16 hi/lo operand
 *      planes (rows [row][d_ld], channel offset d_c0); scale/shift from ssp_bn_finalize(train=0).  fp16 hi/lo operands. ---- */
int ssp_conv_gemm_bnact(int impl, const void* a_hi, const void* a_lo, long long a_rows, int a_ld, int cin,
                        const void* b_hi, const void* b_lo, int b_rows, int b_ld, int N, int H, int W, int taps, int cout,
                        const float* scale, const float* shift, float slope, void* d_hi, void* d_lo, int d_ld, int d_c0,
                        void* stream);
/* ---- first layer nn.Conv2d(3, 32, 3, 1, 1) (darknet.py:156, block 0): direct fp32 convolution of the NCHW image with the
 *      fp32 master weights [32][3][3][3] (k = (kh*3+kw)*3 + ci), output rows [row(n,h,w)][y_ld], optional fp64 BN statistics ---- */
int ssp_conv0_direct(const float* x_nchw, const float* w, const float* bias_or_null, float* y, int y_ld,
                     double* stat_sum_or_null, double* stat_sq_or_null, int N, int H, int W, void* stream);
/* ---- blocks 0-1 of cfg/yolo-pose.cfg as a unit -- nn.Conv2d(3,32,3,1,1) + BatchNorm2d + LeakyReLU + MaxPool2d(2,2) (darknet.py:154-167)
 *      and their autograd (train.py:103) -- without materialising the full-resolution conv output (csrc/l0_fused.cu).
 *      gram: double[SSP_L0_GRAM_DOUBLES]; the first 28*28 hold the matrix (upper triangle: sums of q q^T over all pixels, q = (27 patch
 *      values, 1)), kept from forward to backward; the rest is scratch of ssp_l0_gram (shift correlations, border sums);
 *      code: uint8 [pooled rows][32] (bits 0-1 arg-max position of the 2x2 window, bit 2 pre-activation > 0);
 *      t1: double[28*32] scratch (27 x 32 patch-weighted gradient sums + the 32 plain sums).  w = fp32 master weights [32][27].
 *      ssp_l0_stats writes the per-channel sum / sum of squares that ssp_bn_finalize(count = N*H*W) expects. ---- */
#define SSP_L0_GRAM_DOUBLES 2816
int ssp_l0_gram(const float* x_nchw, int N, int H, int W, double* gram, void* stream);
int ssp_l0_stats(const double* gram, const float* w, double* stat_sum, double* stat_sq, void* stream);
int ssp_l0_fused_fwd(const float* x_nchw, const float* w, const float* scale, const float* shift, float slope, int N, int H, int W,
                     void* d_hi, void* d_lo, int d_ld, int d_c0, unsigned char* code_or_null, void* stream);
int ssp_l0_bwd(const float* x_nchw, const void* g_pooled, int g_f16, int g_ld, int g_c0, const unsigned char* code, float slope, int N,
               int H, int W, double* t1, void* stream);   /* g_f16: the pooled gradient plane holds fp16 (SSP_EPI_F16) instead of fp32 */
int ssp_l0_bwd_finalize(const double* t1, const double* gram, const float* w, const float* gamma, const float* mean,
                        const float* invstd, double count, float grad_scale, float* dw, float* dgamma, float* dbeta, void* stream);
/* ---- nn.Conv2d weight gradient: dW[co][tap][ci] += scale * sum_m dY[m][co] * X[m + shift(tap)][ci] ---- */
int ssp_wgrad_gemm(int impl, const void* dy, long long dy_rows, int dy_ld, int cout, int dy_fmt, const void* x,
                   long long x_rows, int x_ld, int cin, int x_fmt, int N, int H, int W, int taps, float* dw,
                   int dw_ld, int cin_store, float scale, void* stream);

/* ---- nn.BatchNorm2d(eps=1e-4) + nn.LeakyReLU(0.1) + MaxPool2d/Reorg/route placement (darknet.py:96-106,157-176) ---- */
int ssp_bn_finalize(double* stat_sum, double* stat_sq, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, int train, float* mean,
                    float* invstd, float* scale, float* shift, int C, void* stream);
/* ypool (optional, max-pool destinations only): fp32 plane in the POOLED geometry receiving the conv output y at the arg-max
 * position of every 2x2 window.  With it the first pass of the BN backward of a pooled layer runs at a quarter of the
 * resolution: ssp_bn_bwd_reduce(y = ypool, H/2, W/2, g0 = pooled upstream gradient, SSP_ROUTE_DIRECT) accumulates the same
 * S1 / S2 as the full-resolution call (only arg-max positions receive gradient). */
int ssp_bn_apply(const float* y, int y_ld, const float* scale, const float* shift, int N, int C, int H, int W,
                 float slope, void* d0_hi, void* d0_lo, int d0_ld, int d0_c0, int d0_route, void* d1_hi, void* d1_lo,
                 int d1_ld, int d1_c0, int d1_route, float* ypool_or_null, int ypool_ld, void* stream);
int ssp_bn_bwd_reduce(const float* y, int y_ld, const float* scale, const float* shift, const float* mean,
                      const float* invstd, const float* gamma, int N, int C, int H, int W, float slope,
                      const float* g0, int g0_ld, int g0_c0, int g0_route, const float* g1, int g1_ld, int g1_c0,
                      int g1_route, double* s1, double* s2, void* stream);
int ssp_bn_bwd_apply(const float* y, int y_ld, const float* scale, const float* shift, const float* mean,
                     const float* invstd, const float* gamma, int N, int C, int H, int W, float slope,
                     const float* g0, int g0_ld, int g0_c0, int g0_route, const float* g1, int g1_ld, int g1_c0,
                     int g1_route, double* s1, double* s2, void* dy, int dy_ld, int dy_fmt, float dy_scale,
                     void* stream);
int ssp_bn_bwd_finalize(double* s1, double* s2, float* dgamma, float* dbeta, int C, int accumulate, float scale,
                        void* stream);
int ssp_bias_grad_nchw(const float* g_nchw, float* dbias, int N, int C, int HW, int accumulate, float scale,
                       void* stream);

/* ---- parameters: weight re-pack from the fp32 master [cout][taps][cin]; optim.SGD (train.py:388) ---- */
int ssp_pack_weights(const float* w, int cout, int taps, int cin, void* fwd_hi, void* fwd_lo, int fwd_ld,
                     void* dgrad, int dgrad_ld, int dgrad_fmt, void* stream);
int ssp_sgd_step_flat(float* p, const float* g, float* v, long long n, float lr, float momentum,
                      float weight_decay, float grad_scale, void* stream);
/* SGD and the operand-plane re-pack in ONE pass over the flat buffers (csrc/sgd_pack.cu).  `segments` is a DEVICE array with one
 * entry per parameter tensor in flat-buffer order; taps == 0 marks a tensor without operand planes (BN affine, bias).  A conv
 * weight [cout][taps][cin] owns ssp_sgd_segment_blocks() consecutive blocks starting at block0 (prefix sum, filled by the
 * caller); a launch covers blocks [block_begin, block_end), i.e. any run of whole segments: the data-parallel path updates one
 * gradient bucket at a time.  Writes p, v and -- where the pointers are non-null -- W_hi / W_lo [cout][ld_f] (k = tap*cin + ci)
 * and W_d [cin][ld_d] (k = (taps-1-tap)*cout + co), exactly the bytes ssp_pack_weights would produce from the updated p. */
typedef struct ssp_sgd_segment {
  long long off, n;                 /* element offset and count inside the flat buffers */
  int cout, taps, cin;              /* conv weight geometry; taps == 0: plain tensor */
  int ld_f, ld_d, d_fmt;            /* row pitches (elements) of the planes; SSP_FMT_* of W_d */
  void* f_hi; void* f_lo; void* d;  /* device planes, each may be NULL */
  int block0, reserved;
} ssp_sgd_segment;
int ssp_sgd_segment_blocks(int cout, int taps, int cin, long long n);
int ssp_sgd_pack_step(const ssp_sgd_segment* segments_dev, int n_segments, int block_begin, int block_end, float* p,
                      const float* g, float* v, float lr, float momentum, float weight_decay, float grad_scale,
                      void* stream);

/* ---- RegionLoss.forward + build_targets + gradient (region_loss.py:9-175); acc = 8 doubles:
 *      loss_x, loss_y, loss_conf, nGT, nCorrect, nProposals ---- */
int ssp_region_loss_fwd_bwd(const float* out_nchw, const float* target, float* grad_nchw_or_null, double* acc,
                            int B, int num_keypoints, int num_classes, int H, int W, float coord_scale,
                            float noobject_scale, float object_scale, float thresh, int use_conf,
                            float grad_scale, void* stream);
/* ---- get_region_boxes (utils.py:216-296): boxes[B][2K+3] per image, box_global[2K+3] = reference semantics ---- */
int ssp_region_decode_argmax(const float* out_nchw, int B, int num_keypoints, int num_classes, int H, int W,
                             int only_objectness, float* boxes, float* best_conf, float* box_global_or_null,
                             void* stream);

/* ---- multi-object head (multi_obj_pose_estimation/region_loss_multi.py:9-189, utils_multi.py:266-382).
 *      `anchors` is a HOST array of num_anchors*anchor_step floats; acc[6] additionally holds loss_cls.
 *      decode: dense per (image, cell, anchor) arrays in the reference's visiting order + its sequential fallback maxima ---- */
int ssp_region_loss_multi_fwd_bwd(const float* out_nchw, const float* target, float* grad_nchw_or_null, double* acc,
                                  int B, int num_keypoints, int num_classes, int num_anchors, int H, int W,
                                  const float* anchors_host, int anchor_step, float coord_scale, float noobject_scale,
                                  float object_scale, float class_scale, float thresh, int use_conf, float grad_scale,
                                  void* stream);
int ssp_region_decode_multi(const float* out_nchw, int B, int num_keypoints, int num_classes, int num_anchors, int H,
                            int W, int only_objectness, int correspondingclass, float* boxes, float* conf_sel,
                            float* det_conf, float* cls_corr, long long* max_ind, float* max_conf, float* max_cls,
                            void* stream);

/* ---- pnp (utils.py:86-100 -> cv2.solvePnP ITERATIVE + Rodrigues), compute_projection (utils.py:40-45) ---- */
int ssp_pnp_batched(const float* points3d, int points3d_shared, const float* points2d, const float* K3x3,
                    int num_points, long long n, int max_iter, double* R_out, double* t_out,
                    int* iters_out_or_null, void* stream);
/* same solve; work_out [n][3] int = {Jacobi sweeps of the 12x12 DLT, accepted LM iterations, LM linear solves} per problem: what the
 * bench's achieved-FLOP/s figure is computed from */
int ssp_pnp_batched_work(const float* points3d, int points3d_shared, const float* points2d, const float* K3x3,
                         int num_points, long long n, int max_iter, double* R_out, double* t_out, int* work_out,
                         void* stream);
int ssp_project_points(const float* X, int rows, int nv, const double* Rt, const double* K3x3, long long n,
                       float* out, void* stream);

/* ---- training-image pipeline: byte-exact with the Pillow routines image.py calls.  Images are device
 *      uint8 HWC RGB, dense.  resample = PIL.Image.Resampling value (0 NEAREST, 2 BILINEAR, 3 BICUBIC = resize()'s default
 *      in Pillow >= 7).  `work` is caller-provided device scratch (16-B aligned) of at least *_work_bytes() bytes.
 *  ssp_aug_resize_u8: Image.crop((x0, y0, x0+in_w, y0+in_h)).resize((out_w, out_h), resample) (image.py:64,69; dataset.py:103);
 *      the crop window may stick out of the source (zero fill), pass (0, 0, src_w, src_h) for a plain resize.
 *  ssp_aug_rgb2hsv_u8 / hsv2rgb_u8: Image.convert('HSV') / ('RGB') (image.py:15,30).
 *  ssp_aug_to_tensor_u8: torchvision ToTensor (the `transform` of dataset.py:103-118) of a dense uint8 HWC image: float32 CHW,
 *      byte / 255 as an IEEE division (what the CPU reference computes; a reciprocal multiply differs by 1 ulp).
 *  ssp_aug_sample: change_background (image.py:110-127) -> crop -> resize -> distort_image (image.py:14-32) -> ToTensor
 *      (dataset.py transform) for one sample.  luts = 5 x 256 bytes: posmask, negmask (image.py:121-122), hue, saturation,
 *      value (image.py:17-27) point() tables, built by the host exactly as Image.point() builds them.  Crop window
 *      (pleft, ptop, cw, ch) with cw = swidth - 1, ch = sheight - 1 (image.py:64).  out_u8 (HWC) and out_chw (float32
 *      CHW in [0,1]) are both optional, at least one required. ---- */
long long ssp_aug_resize_work_bytes(int in_w, int in_h, int out_w, int out_h, int resample);
int ssp_aug_resize_u8(const void* src, int src_w, int src_h, int x0, int y0, int in_w, int in_h, void* dst, int out_w,
                      int out_h, int resample, void* work, long long work_bytes, void* stream);
int ssp_aug_rgb2hsv_u8(const void* rgb, void* hsv, long long n_pixels, void* stream);
int ssp_aug_hsv2rgb_u8(const void* hsv, void* rgb, long long n_pixels, void* stream);
int ssp_aug_to_tensor_u8(const void* hwc_u8, long long n_pixels, float* out_chw, void* stream);
long long ssp_aug_sample_work_bytes(int ow, int oh, int bw, int bh, int cw, int ch, int out_w, int out_h, int resample);
int ssp_aug_sample(const void* img, const void* mask, int ow, int oh, const void* bg, int bw, int bh, const void* luts,
                   int pleft, int ptop, int cw, int ch, int out_w, int out_h, int resample, void* work,
                   long long work_bytes, void* out_u8_or_null, float* out_chw_or_null, void* stream);

/* Batched form of ssp_aug_sample: ONE launch per pipeline stage for the whole batch (<= 10 launches instead of ~10 per sample).
 *   ssp_aug_batch_plan (host only, no device access): items[n] hold the per-sample arguments of ssp_aug_sample with DEVICE
 *     pointers (each sample its own work buffer of ssp_aug_sample_work_bytes()); writes the op table (table_bytes >=
 *     ssp_aug_batch_table_bytes(n)) into HOST memory -- typically the tail of the pinned staging buffer, so that it travels in the
 *     batch's single host->device copy -- and stage_dims[20].
 *   ssp_aug_batch_run: table_dev = the device copy of that table; launches the stages on `stream`. */
typedef struct ssp_aug_item {
  const void* img; const void* mask; int ow, oh; const void* bg; int bw, bh; const void* luts; int pleft, ptop, cw, ch;
  void* work; long long work_bytes; void* out_u8; float* out_chw;
} ssp_aug_item;
long long ssp_aug_batch_table_bytes(int n);
int ssp_aug_batch_plan(const ssp_aug_item* items_host, int n, int out_w, int out_h, int resample, void* table_host,
                       long long table_bytes, int* stage_dims_host20);
int ssp_aug_batch_run(const void* table_dev, int n, const int* stage_dims_host20, void* stream);

#ifdef __cplusplus
}
#endif
#endif
