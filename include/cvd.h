/*
 * cvd.h — C-ABI of libcvd_sm100.so: the B200 (sm_100a) kernels behind the
 * consistent_depth test-time fine-tuning hot path.
 *
 * The reference (facebookresearch/consistent_depth) is pure Python/PyTorch and
 * has NO FFI of its own; every entry point below replaces a group of ATen ops
 * the reference dispatches from Python.  Each declaration cites the reference
 * interface (file:line under /root/reference) it stands in for.  The Python
 * host layer (the consistent_depth_b200 package) binds these with ctypes and mirrors
 * the reference's module / class / argument names.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - all data pointers are DEVICE pointers unless the name ends in _host.
 *   - every call enqueues work on `stream` (a cudaStream_t passed as void*),
 *     never synchronises it, never allocates device memory.
 *   - return value: 0 = ok, non-zero = error; cvd_last_error() gives a
 *     thread-local message.  Nothing throws.
 *   - images/activations handed over by the Python boundary are NCHW fp32 as
 *     in the reference; the engine's internal activation layout is NHWC fp32.
 */
#ifndef CVD_H_
#define CVD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVD_VERSION 100

/* ---- misc ---------------------------------------------------------------- */
int         cvd_version(void);
const char* cvd_last_error(void);
/* number of kernels this library has launched since load (bench.py's
 * gpu_launches evidence). */
long long   cvd_launch_count(void);

/* ---- fused geometric-consistency loss ------------------------------------
 * Replaces loss/consistency_loss.py:98-253 (ConsistencyLoss.__call__ +
 * geometry_consistency_loss + weighted_mean_loss:73-89) and the
 * utils/geometry.py primitives it calls (pixel_grid:9, pixels_to_rays:38,
 * pixels_to_points:86, reproject_points:103, project:64, sample:201),
 * forward AND backward (what loss.backward() does for this sub-graph,
 * depth_fine_tuning.py:282), in one memory-bound kernel.
 *
 *   depth      (B,2,H,W)   predicted depths (DepthModel.forward output)
 *   flow0/1    (B,2,H,W)   metadata["geometry_consistency"]["flows"][k]
 *   mask0/1    (B,1,H,W)   ...["masks"][k], values 0/1
 *   extr       (B,2,3,4)   [R|t] camera->world ; intr (B,2,4) = fx,fy,cx,cy
 *   msum       (B,2)       sum(mask_k[b]) from cvd_mask_sums; the kernel applies
 *                          1/max(msum,1e-6) (weighted_mean_loss, :73-89)
 *   f_dir_host [2] or NULL mean focal length over the GLOBAL batch for
 *                          direction k (consistency_loss.py:178); NULL =>
 *                          computed on device from this call's intr.
 *   f_dir_dev  [2] or NULL the same two scalars in DEVICE memory (read by the
 *                          kernels at run time, so a captured CUDA graph stays
 *                          valid when the global batch's intrinsics change);
 *                          takes precedence over f_dir_host.
 *   B_global               divisor of the final torch.mean over pairs (:208)
 *   workspace  cvd_consistency_workspace_bytes(B) bytes of device scratch, 16-byte aligned
 *   out_pair   (2,B)  f32  [0]=lambda_r*reprojection[b], [1]=lambda_b*disparity[b]
 *                          (the reference's batch_losses dict)
 *   out_loss   (1)    f32  sum_b(out_pair)/B_global
 *   grad_depth (B,2,H,W) or NULL. d out_loss / d depth; zeroed by the call.
 */
int cvd_mask_sums(const float* mask0, const float* mask1, int B, int H, int W,
                  float* msum, void* stream);

size_t cvd_consistency_workspace_bytes(int B);

int cvd_consistency_fwd_bwd(const float* depth,
                            const float* flow0, const float* flow1,
                            const float* mask0, const float* mask1,
                            const float* extr, const float* intr,
                            const float* msum,
                            const float* f_dir_host,
                            const float* f_dir_dev,
                            float lambda_reprojection, float lambda_view_baseline,
                            int B, int B_global, int H, int W,
                            void* workspace, float* out_pair, float* out_loss,
                            float* grad_depth, void* stream);

/* ---- fused Adam -----------------------------------------------------------
 * Replaces optimizer/__init__.py:16 (torch.optim.Adam) .step() as called at
 * depth_fine_tuning.py:283, including the NaN guard of :278-280 evaluated on
 * the device: if *loss_flag is NaN the step is skipped (no state change).
 * step_state: 16-byte device record {int step; int skip; float step_size;
 * float bc2_sqrt}, zero-initialised by the caller once, then owned by the kernels.
 *   grad_scale multiplies g before use (1.0, or 1/world for rank-summed grads).
 */
int cvd_adam_flat(float* p, const float* g, float* m, float* v, long long n,
                  float lr, float beta1, float beta2, float eps, float grad_scale,
                  int* step_state /* device int[4], zero-initialised once */,
                  const float* loss_flag /* device float[1] or NULL */,
                  void* stream);

/* lambda_parameter regulariser (loss/parameter_loss.py:13-19):
 * out_loss[0] += lambda * sum|p - p0| ; g += lambda*sign(p-p0). */
int cvd_param_l1(const float* p, const float* p0, long long n, float lambda,
                 float* g_accum, float* out_loss_accum, void* stream);

/* ---- conv engine (tcgen05 implicit GEMM), csrc/conv_tc.cu ------------------
 * Activations are NHWC fp32.  A conv reads / writes a channel VIEW of a wider
 * buffer: logical channel c lives at physical channel
 *      c_off + c + (c >= n0 ? gap : 0)          (n0 <= 0: no gap)
 * so torch.cat (hourglass.py:55) costs nothing.
 *
 * Per-channel input transform applied while the activation tile is staged:
 *   CVD_XF_AFFINE : v = a[c]*x + b[c] (a == NULL: identity); if relu: v = max(v,0)
 *                   == BatchNorm2d(train) + ReLU of the PRODUCER layer (hourglass.py:28-29)
 *   CVD_XF_BNBWD  : y = a[c]*x + b[c]; g = (!relu || y > 0) ? dy : 0;
 *                   v = bw[c].x*g - bw[c].y - bw[c].z*y
 *                   == backward of that BatchNorm+ReLU (autograd, depth_fine_tuning.py:282)
 * a / b / bw are indexed by PHYSICAL channel of x.
 */
#define CVD_XF_AFFINE 0
#define CVD_XF_BNBWD  1

typedef struct {
  const float* x;        /* NHWC fp32 tensor, c_total channels per pixel            */
  const float* dy;       /* BNBWD only: gradient wrt the post-activation tensor     */
  const float* a;        /* per-channel scale or NULL                               */
  const float* b;        /* per-channel shift or NULL                               */
  const float* bw;       /* BNBWD only: [C][4] floats (c0, c1, c2, unused)          */
  int c_total, c_off, n0, gap;
  int dy_ctotal, dy_coff, dy_n0, dy_gap;
  int relu;
  int mode;              /* CVD_XF_*                                                */
} cvd_src_t;

typedef struct {
  float* y;              /* NHWC fp32 destination                                   */
  int c_total, c_off, n0, gap;
} cvd_dst_t;

/* Packs fp32 OIHW conv weights (torch layout, hourglass.py:27,39,42) into bf16
 * hi(/lo) UMMA core-matrix blobs, one per (tap, 16-channel k-block), that the
 * conv kernel streams with cp.async.bulk.  transpose_flip != 0 builds the dgrad
 * operand (Cin<->Cout swapped, taps rotated 180 degrees).
 * precision: 1 = bf16, 3 = bf16x3 split (hi+lo, fp32-class).
 * Grouped convolutions (ResNeXt's 3x3, groups=32): pass group_size << 8 in the upper bits of transpose_flip and
 * cin == cout == the chunk width (a multiple of group_size); w then points at rows [chunk .. chunk+cout) of the
 * (Cout, group_size, k, k) weight and is expanded block-diagonally, so the dense kernel computes the grouped conv
 * of that channel chunk. */
size_t cvd_conv_packed_bytes(int cin, int cout, int k, int precision);
int cvd_conv_pack_weights(const float* w_oihw, int cin, int cout, int k, int transpose_flip,
                          int precision, void* packed, void* stream);

/* The same packing for n convolutions in one launch.  descs_dev: device array of
 * { const float* w_oihw; void* packed; int cin; int cout; int k; int transpose_flip; } (32 bytes each). */
int cvd_conv_pack_batch(const void* descs_dev, int n, int precision, void* stream);

/* Convolution, stride 1, "same" zero padding, over N images of HxW: replaces
 * nn.Conv2d forward (hourglass.py:27,39,42,164,173).  With transpose_flip
 * weights (cin/cout given in GEMM terms: cin = channels of the tensor being
 * read) and a CVD_XF_BNBWD source it is the conv input-gradient.
 *   flags bit0: accumulate into dst (+=) ; bit1: exp() epilogue
 *         (mannequin_challenge_model.py:66)
 */
int cvd_conv_fwd(const cvd_src_t* src, const void* packed_w, const float* bias,
                 const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                 int precision, int flags, void* stream);

/* The same convolution with the BatchNorm2d(train) batch statistics of its OUTPUT fused into the epilogue
 * (what cvd_bn_stats computes in a separate pass): per-channel sum / sum of squares -> a, b, rstd, mean at
 * physical channels [dst->c_off, +cout) of the arrays, running statistics of this conv's [cout] arrays updated.
 * scratch: cvd_bn_scratch_bytes(cout) bytes (one block per 256 output channels: the 256-column chunks of a wide
 * conv run concurrently in one launch), zeroed once by the caller (self-cleaning; may be shared with
 * cvd_bn_stats / cvd_bn_bwd_reduce as long as the calls are stream-ordered). */
typedef struct {
  void* scratch;
  const float* gamma; const float* beta;        /* [cout] or NULL (affine=False) */
  float* running_mean; float* running_var;      /* [cout] or NULL */
  float* a; float* b; float* rstd; float* mean; /* per physical channel of the destination buffer */
  float eps, momentum;
} cvd_bn_t;
int cvd_conv_fwd_bn(const cvd_src_t* src, const void* packed_w, const float* bias,
                    const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                    int precision, int flags, const cvd_bn_t* bn, void* stream);

/* nchunks independent convolutions of one shape in ONE launch (blockIdx.y = chunk): chunk j reads the source view
 * shifted by j * src_shift channels, writes the destination view (bias, BatchNorm arrays) shifted by j * dst_shift
 * channels and uses the packed weights at packed_w + j * packed_stride bytes.  bn may be NULL; if not, scratch holds
 * nchunks blocks of cvd_bn_scratch_bytes(256).  Views must be gap-free.  Used for the 64-channel chunks of the grouped
 * ResNeXt conv2 (midas_v2: blocks.py:23 resnext101_32x8d) -- same results as nchunks cvd_conv_fwd[_bn] calls. */
int cvd_conv_fwd_chunks(const cvd_src_t* src, const void* packed_w, const float* bias,
                        const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                        int precision, int flags, const cvd_bn_t* bn,
                        int nchunks, int src_shift, int dst_shift, long long packed_stride, void* stream);

/* ---- second-generation conv path: pre-split operands + TMA-fed, kx-fused tcgen05 conv (csrc/prep.cu, conv2.cu) ----
 * cvd_prep_operand applies the per-channel transform of `src` (CVD_XF_AFFINE: BatchNorm+ReLU of the producer,
 * hourglass.py:28-29; CVD_XF_BNBWD: its backward) ONCE and writes the result as two bf16 planes (v = hi + lo) in the
 * chunk-planar layout  z[plane][n][c/8][y*W+x][c%8]  -- the layout a TMA box load turns into the UMMA shared-memory
 * operand.  The C logical channels of the view become dense channels [8*zc8_off, 8*zc8_off + ceil16(C)) of z, whose
 * planes hold zc8 chunks per image.  z bytes: 2 * N * zc8 * HW * 16. */
int cvd_prep_operand(const cvd_src_t* src, int C, long long N, long long HW, void* z, int zc8, int zc8_off,
                     int precision, void* stream);

/* Weight tiles of the kx-fused conv.  G horizontal taps share one GEMM-N block (N = G * ceil16(cout_gemm) <= 256);
 * cvd_conv2_tap_groups reports (G, number of groups).  cvd_conv2_pack_batch packs n convolutions in one launch from a
 * device table of { const float* w_oihw; void* packed; int cin, cout, k, flip, G, ng } (40 bytes; cin/cout = the OIHW
 * extents, flip = 1 builds the dgrad operand: channels swapped, taps rotated 180 degrees). */
int cvd_conv2_tap_groups(int cout_gemm, int k, int* G, int* ng);
size_t cvd_conv2_packed_bytes(int cin_gemm, int cout_gemm, int k);
int cvd_conv2_pack_batch(const void* descs_dev, int n, void* stream);

/* Convolution (stride 1, "same"), replaces nn.Conv2d forward (hourglass.py:27,39,42) and, with flip-packed weights and a
 * CVD_XF_BNBWD-prepared operand, its input gradient.  z / zc8 / zc8_off: operand planes from cvd_prep_operand (the conv
 * reads ceil16(cin) channels from chunk zc8_off); cin / cout in GEMM terms; bias, dst, flags, bn as cvd_conv_fwd(_bn)
 * (bn may be NULL).  bf16x3 split precision (fp32-class) only. */
int cvd_conv2_fwd(const void* z, int zc8, int zc8_off, const void* packed_w, const float* bias,
                  const cvd_dst_t* dst, int N, int H, int W, int cin, int cout, int k,
                  int flags, const cvd_bn_t* bn, void* stream);

/* Weight gradient on the pre-split operand planes (csrc/wgrad2.cu): dW (fp32 OIHW [cout][cin][k][k], RED-accumulated,
 * caller zeroes it) += sum over pixels of G (x) X, with xz the planes of the conv's INPUT as the forward saw it (cin
 * channels from chunk x_off; the planes cvd_conv2_fwd read) and gz the planes of the gradient wrt the conv's raw output
 * (cout channels from chunk g_off; the planes the dgrad cvd_conv2_fwd reads).  TMA-fed; kx taps fused into GEMM N, ky taps
 * stacked into GEMM M.  Returns 0 = launched, 1 = error, 2 = shape not covered (cin not 16/32/64/128 for k > 1, ...):
 * the caller then uses cvd_conv_wgrad. */
int cvd_conv2_wgrad(const void* xz, int xc8, int x_off, const void* gz, int gc8, int g_off, float* dw_oihw,
                    int N, int H, int W, int cin, int cout, int k, void* stream);

/* Weight gradient of one channel chunk of a grouped convolution (see cvd_conv_pack_weights): gsrc / xsrc are views
 * of the chunk's c output / input channels, dw points at the chunk's rows of the (Cout, group_size, k, k) gradient. */
int cvd_conv_wgrad_grouped(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_ogkk,
                           int N, int H, int W, int c, int group_size, int k, int precision, void* stream);

/* All nchunks chunks of a grouped convolution's weight gradient in ONE launch (blockIdx.z = chunk): gsrc / xsrc are
 * gap-free views of chunk 0, chunk j reads channels [j*c, (j+1)*c) behind them; dw is the whole (Cout, group_size, k, k)
 * gradient.  Same result as nchunks cvd_conv_wgrad_grouped calls (ResNeXt conv2 of midas_v2: up to 32 chunks). */
int cvd_conv_wgrad_grouped_chunks(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_ogkk,
                                  int N, int H, int W, int c, int nchunks, int group_size, int k, int precision,
                                  void* stream);

/* Weight gradient of the same convolution: dW (fp32 OIHW, [cout][cin][k][k]) += sum over all
 * pixels of G (x) X, the wgrad half of autograd's conv backward (depth_fine_tuning.py:282).
 * gsrc: view of the gradient wrt the conv's raw output (CVD_XF_BNBWD of the following BatchNorm,
 * or CVD_XF_AFFINE for a plain gradient tensor); xsrc: the conv's input as the forward saw it.
 * Partial sums are RED-accumulated: the caller zeroes dW first. */
int cvd_conv_wgrad(const cvd_src_t* gsrc, const cvd_src_t* xsrc, float* dw_oihw,
                   int N, int H, int W, int cin, int cout, int k, int precision, void* stream);

/* ---- BatchNorm bookkeeping / elementwise glue (csrc/elementwise.cu) -------
 * nn.BatchNorm2d in TRAIN mode (hourglass.py:28,40,43,165): batch statistics of the raw conv
 * output x (physical channels [c_off, c_off+C) of a c_total-wide NHWC buffer) over npix pixels ->
 *   a[c_off+c] = gamma*rstd, b[c_off+c] = beta - mean*a   (consumed by CVD_XF_AFFINE on load)
 *   rstd / mean saved for backward; running_mean/var updated (momentum, unbiased var).
 * gamma/beta/running_* are the layer's own [C] arrays (NULL: affine=False / no running stats).
 * scratch: cvd_bn_scratch_bytes(C) bytes, zeroed once by the caller, self-cleaning afterwards. */
size_t cvd_bn_scratch_bytes(int C);
int cvd_bn_stats(const float* x, int c_total, int c_off, int C, long long npix, void* scratch,
                 const float* gamma, const float* beta, float eps, float momentum,
                 float* running_mean, float* running_var,
                 float* a, float* b, float* rstd, float* mean, void* stream);

/* BatchNorm(+ReLU) backward reductions for the same channel range: with y = a x + b,
 * g = dy*[y>0 or !relu]:  bw[c_off+c] = (c0, c1, c2, 0) such that d x = c0 g - c1 - c2 y
 * (what CVD_XF_BNBWD applies on load), plus dgamma/dbeta (affine BN) and the gradient of the
 * conv bias feeding this BN (dbias = sum d x), each [C] or NULL.  dy uses its own view;
 * dy_lc0 = logical channel of dy that corresponds to x's physical channel c_off. */
int cvd_bn_bwd_reduce(const float* x, int x_ctotal, int x_coff,
                      const float* dy, int dy_ctotal, int dy_coff, int dy_n0, int dy_gap, int dy_lc0,
                      const float* a, const float* b, const float* rstd, const float* mean,
                      const float* gamma, const float* beta, int relu,
                      long long npix, int C, void* scratch,
                      float* bw, float* dgamma, float* dbeta, float* dbias, void* stream);

/* nn.AvgPool2d(2) (hourglass.py:70,95,113,138) of relu?(a x + b) read through a view -> plain
 * (N,H/2,W/2,C); backward: dx (N,H,W through a channel view) (+)= 0.25 * dp. */
int cvd_pool_fwd(const float* x, int c_total, int c_off, int n0, int gap, const float* a, const float* b,
                 int relu, float* p, int N, int H, int W, int C, void* stream);
int cvd_pool_bwd(const float* dp, float* dx, int c_total, int c_off, int n0, int gap, int accumulate,
                 int N, int H, int W, int C, void* stream);

/* ChannelsN.forward (hourglass.py:81,106,131,156): z = relu(a1 x1 + b1) + up2x(relu(a2 x2 + b2)),
 * nn.UpsamplingBilinear2d(scale_factor=2) (align_corners=True); x2 is (N,H/2,W/2,.).  z plain.
 * Backward: dy2 (view, half res) = up2x^T(dz) (gather form, deterministic);
 *           dy1 (view, full res, may be NULL) (+)= dz. */
int cvd_merge_up_fwd(const float* x1, int ct1, int c01, int n01, int gap1, const float* a1, const float* b1,
                     const float* x2, int ct2, int c02, int n02, int gap2, const float* a2, const float* b2,
                     float* z, int N, int H, int W, int C, void* stream);
int cvd_merge_up_bwd(const float* dz, float* dy2, int ct2, int c02, int n02, int gap2,
                     float* dy1, int ct1, int c01, int n01, int gap1, int accumulate1,
                     int N, int H, int W, int C, void* stream);

/* (N,3,H,W) BGR image (DepthModel.forward input, depth_model.py:12-16) -> (N,H,W,4), 4th channel 0 */
int cvd_image_to_nhwc4(const float* img_nchw, float* out, int N, int H, int W, void* stream);

/* exp() backward of mannequin_challenge_model.py:66: out4[i] = (grad_depth[i]*depth[i], 0, 0, 0);
 * dbias (pred_layer.bias gradient, 1 float, accumulated) may be NULL. */
int cvd_dlogdepth(const float* grad_depth, const float* depth, float* out4, long long n, float* dbias, void* stream);

/* ------------------------------------------------------------------------------------------------
 * monodepth2 backbone: the passes around the conv engine.  All tensors NHWC fp32, plain
 * (channel offset 0) unless a (c_total, c_off) pair is given; C multiples of 4.
 * ------------------------------------------------------------------------------------------------ */

/* F.interpolate(images, size=feed, mode='bicubic', align_corners=False) (monodepth2_model.py:72-74) fused with
 * (x - 0.45) / 0.225 (resnet_encoder.py:89): (N,3,H,W) -> (N,oh,ow,4), 4th channel 0. */
int cvd_bicubic_image_fwd(const float* img_nchw, int N, int H, int W, float* out_nhwc4, int oh, int ow,
                          float mean, float inv_std, void* stream);

/* depth = 1 / bicubic(disp -> (H,W)) (monodepth2_model.py:79-82); disp (N,fh,fw), depth (N,H,W).
 * Backward: ddisp (N,fh,fw) = bicubic^T(-ddepth * depth^2); ddisp is zeroed by the call. */
int cvd_disp_to_depth_fwd(const float* disp, int N, int fh, int fw, float* depth, int H, int W, void* stream);
int cvd_disp_to_depth_bwd(const float* ddepth, const float* depth, int N, int fh, int fw, int H, int W,
                          float* ddisp, void* stream);

/* nn.Sigmoid of depth_decoder.py:63 on channel 0 of the dispconv output, which lives on the reflect-padded grid
 * (N,fh+2,fw+2,c_total); backward writes channel 0 of the interior of draw_padded (border / other channels untouched). */
int cvd_sigmoid_fwd(const float* raw_padded, int c_total, int N, int fh, int fw, float* disp, void* stream);
int cvd_sigmoid_bwd(const float* ddisp, const float* disp, int N, int fh, int fw, float* draw_padded, int c_total, void* stream);

/* Stride-2 convolutions of torchvision's resnet18 (conv1, layerN.0.conv1, layerN.0.downsample.0) run as the stride-1
 * conv at the input resolution followed by picking pixels (2y,2x):  dst (N,ceil(H/2),ceil(W/2),C) = src[:, ::2, ::2].
 * cvd_stuff2 is its transpose: dst (N,H,W,C) at (2y,2x) (+)= src; other pixels are not touched. */
int cvd_subsample2(const float* src, int N, int H, int W, int C, float* dst, void* stream);
int cvd_stuff2(const float* src, int N, int H, int W, int C, float* dst, int accumulate, void* stream);

/* BatchNorm(+ReLU) backward, materialised: dst[(stride*y, stride*x)] = c0 g - c1 - c2 (a x + b), g = dy [relu: * (a x + b > 0)],
 * with (c0,c1,c2) = bw[c] from cvd_bn_bwd_reduce; x, dy (N,h,w,C); dst (N,H,W,C), pixels off the stride grid untouched. */
int cvd_bnbwd_stuff(const float* x, const float* dy, const float* a, const float* b, const float* bw, int relu,
                    int N, int h, int w, int C, float* dst, int H, int W, int stride, void* stream);

/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) (resnet_encoder.py:93) of relu?(a x + b): (N,H,W,C) ->
 * (N,ceil(H/2),ceil(W/2),C) + arg-max tap (0..8, first maximum in scan order like torch) per output element.
 * Backward: dx (N,H,W,C) (+)= sum of dout over the windows whose arg-max is this pixel. */
int cvd_maxpool3s2_fwd(const float* x, const float* a, const float* b, int relu, int N, int H, int W, int C,
                       float* out, unsigned char* argmax, void* stream);
int cvd_maxpool3s2_bwd(const float* dout, const unsigned char* argmax, int N, int H, int W, int C, float* dx,
                       int accumulate, void* stream);

/* torchvision BasicBlock tail: out = relu(a y + b + r'), r' = ra r + rb (downsample branch) or r (ra = rb = NULL).
 * Backward: dout <- dout * [out > 0] in place; dres (+)= that (dres may be NULL). */
int cvd_bn_add_relu(const float* y, const float* a, const float* b, const float* res, const float* ra, const float* rb,
                    long long npix, int C, float* out, void* stream);
int cvd_relu_bwd_add(float* dout, const float* out, float* dres, int accumulate, long long n, void* stream);

/* Input of a Conv3x3 of monodepth2 (layers.py:121-136): dst[(N, hu+2, wu+2, d_ctotal) at channels d_coff..] =
 * ReflectionPad2d(1)( T(src) upsampled x2 nearest if `upsample` ), hu = hs << upsample; torch.cat = two calls with
 * different d_coff.  src is (N, hs + 2 s_pad, ws + 2 s_pad, s_ctotal) read at its interior (s_pad = 1 for a conv output
 * that itself lives on a padded grid).  T: identity, ELU (layers.py:113), or relu(a x + b) (encoder feature 0).
 * Backward (transpose, gather form, deterministic): dsrc (own geometry) (+)= T'(src) * sum of dpad over the pre-images;
 * for CVD_GATHER_AFFINE_RELU T' = 1 (the ReLU / BatchNorm backward happens where dsrc is consumed). */
#define CVD_GATHER_IDENTITY 0
#define CVD_GATHER_ELU 1
#define CVD_GATHER_AFFINE_RELU 2
int cvd_gather_pad_fwd(const float* src, int s_ctotal, int s_coff, int s_pad, const float* a, const float* b,
                       float* dst, int d_ctotal, int d_coff, int N, int hs, int ws, int C, int upsample, int mode,
                       void* stream);
int cvd_gather_pad_bwd(const float* dpad, int p_ctotal, int p_coff, const float* src, int s_ctotal, int s_coff, int s_pad,
                       float* dsrc, int ds_ctotal, int ds_coff, int ds_pad, int N, int hs, int ws, int C, int upsample,
                       int mode, int accumulate, void* stream);

/* out[c] += sum over pixels of x[p, c_off + c] (bias gradient of a conv without BatchNorm); C a power of two <= 256. */
int cvd_channel_sum(const float* x, int c_total, int c_off, int C, long long npix, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MiDaS-v2 network: the passes around the conv engine.  NHWC fp32, C multiples of 4.
 * ------------------------------------------------------------------------------------------------ */

/* (images - mean) / std per channel (midas_v2_model.py:58-59): (N,3,H,W) -> (N,H,W,4), 4th channel 0.
 * mean3_host / std3_host: 3 host floats each (passed by value to the kernel). */
int cvd_image_normalize_nhwc4(const float* img_nchw, int N, int H, int W, const float* mean3_host,
                              const float* std3_host, float* out_nhwc4, void* stream);

/* out = relu(x) + other (other may be NULL): the skip term of ResidualConvUnit (blocks.py:111-117, whose in-place
 * ReLU makes the skip relu(x)) and the FeatureFusionBlock sum (blocks.py:146-147); the conv then accumulates into out. */
int cvd_relu_add(const float* x, const float* other, float* out, long long n, void* stream);

/* out (N,2h,2w,C) = bilinear_x2(x (N,h,w,C)) [+ r or relu(r)], align_corners as given (blocks.py:151-153 True,
 * midas_net.py:40 False).  Backward (gather form): dx (N,h,w,C) (+)= transpose applied to dout. */
int cvd_up2_bilinear_fwd(const float* x, const float* r, int relu_r, int N, int h, int w, int C, int align_corners,
                         float* out, void* stream);
int cvd_up2_bilinear_bwd(const float* dout, int N, int h, int w, int C, int align_corners, float* dx, int accumulate, void* stream);

/* depth = 1 / relu(raw) (midas_net.py:43 ReLU + midas_v2_model.py:67 reciprocal) on channel 0 of 4-channel pixels;
 * backward: draw4 = (raw > 0 ? -ddepth * depth^2 : 0, 0, 0, 0). */
int cvd_recip_relu_fwd(const float* raw4, float* depth, long long n, void* stream);
int cvd_recip_relu_bwd(const float* ddepth, const float* depth, const float* raw4, float* draw4, long long n, void* stream);


/* ------------------------------------------------------------------------------------------------
 * the flow + photometric consistency masks of
 * utils/consistency.py:53-67 (flow.py:199-228) for B frame pairs, both directions, in one launch.
 * flows (B, 2 directions, 2, H, W): [b,0] = flow frame0 -> frame1, [b,1] = flow frame1 -> frame0 (pixels);
 * colors (B, 2 frames, 3, H, W); masks (B, 2 directions, H, W) float {0, 1}.
 * ------------------------------------------------------------------------------------------------ */
int cvd_flow_consistency_masks(const float* flows, const float* colors, float* masks, int B, int H, int W,
                               float flow_thresh, float color_thresh, void* stream);

/* ------------------------------------------------------------------------------------------------
 * the FlowNet2 custom ops, forward only, NCHW fp32 (run on a B200 against the oracle).
 * cvd_correlation_fwd  = correlation_package (correlation_cuda_kernel.cu:51-128; FlowNetC.py:28-31 uses pad = md = 20,
 *                        K = 1, s1 = 1, s2 = 2): out (B, D*D, Ho, Wo), D = 2 (md / s2) + 1, sizes from cvd_correlation_out_size
 * cvd_resample2d_fwd   = resample2d_package (resample2d_kernel.cu:17-73, kernel_size 1, bilinear): out = in1 sampled at (x + flow_x, y + flow_y)
 * cvd_channelnorm_fwd  = channelnorm_package (channelnorm_kernel.cu:16-60, norm_deg 2): out (B,1,H,W) = sqrt(sum_c in^2)
 * ------------------------------------------------------------------------------------------------ */
int cvd_correlation_out_size(int H, int W, int pad, int K, int md, int s1, int s2, int* channels, int* Ho, int* Wo);
int cvd_correlation_fwd(const float* in1, const float* in2, float* out, int B, int C, int H, int W,
                        int pad, int K, int md, int s1, int s2, void* stream);
int cvd_resample2d_fwd(const float* in1, const float* flow, float* out, int B, int C, int H, int W, void* stream);
int cvd_channelnorm_fwd(const float* in, float* out, int B, int C, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVD_H_ */
