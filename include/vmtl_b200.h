/*
 * vmtl_b200.h -- C ABI of libvmtl_b200.so: hand-written sm_100a CUDA kernels for the
 * per-step multi-task hot path of kirilllzaitsev/vision_mtl.
 *
 * The reference is pure Python and has no FFI; the "interface each entry point replaces"
 * is therefore the ATen op sequence behind a Python call site.  Citations are
 * /root/reference paths (file:line).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - feature maps are NHWC ("channels_last"): a [B,C,H,W] map is a row-major
 *     [npix = B*H*W, C] matrix, 16-byte aligned, C % 4 == 0;
 *   - the caller owns every buffer, including workspaces (size them with the
 *     *_workspace_bytes helpers); workspaces need no initialisation unless noted;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never
 *     synchronise, allocate or free;
 *   - return value: 0 (VMTL_OK) or a negative VMTL_E* code; vmtl_strerror() names it.
 *   - all floating point is fp32; reductions are accumulated in fp64 and are
 *     deterministic (fixed-order two-stage reductions, integer atomics only).
 */
#ifndef VMTL_B200_H_
#define VMTL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMTL_OK 0
#define VMTL_EINVAL (-1)       /* bad argument value                       */
#define VMTL_EALIGN (-2)       /* pointer / channel count not 16B friendly */
#define VMTL_ECUDA (-3)        /* a CUDA runtime call or launch failed      */
#define VMTL_EUNSUPPORTED (-4) /* shape outside what the kernels cover      */
#define VMTL_EWORKSPACE (-5)   /* workspace too small                       */

#define VMTL_MAX_TASKS 4

/* cross-stitch modes */
#define VMTL_XS_REFERENCE_DIAG 0 /* y[a] = alpha[a,a(,c)] * x[a]  -- what the reference einsum computes */
#define VMTL_XS_FULL_MIX 1       /* y[a] = sum_b alpha[a,b(,c)] * x[b] */

/* gate contraction precision */
#define VMTL_GATE_FP32_FFMA 0 /* CUDA-core fp32 FMA contraction                        */
#define VMTL_GATE_TC_3XTF32 1 /* tcgen05 kind::tf32, hi/lo split (3xTF32, fp32-grade) */
#define VMTL_GATE_TC_TF32 2   /* tcgen05 kind::tf32, single pass                      */

/* logits layouts for the *_logits loss kernels */
#define VMTL_LAYOUT_NCHW 0
#define VMTL_LAYOUT_NHWC 1

int vmtl_version(void);
const char* vmtl_strerror(int code);
/* number of SMs of the current device (grid sizing); <0 on error */
int vmtl_sm_count(void);

/* ------------------------------------------------------------------------------------
 * Cross-stitch unit.  Replaces torch.stack + torch.einsum at
 * vision_mtl/models/cross_stitch_model.py:32-37 (forward) and :143-156 (call site), and
 * their autograd backward.
 *   x_host / y_host : host arrays of T device pointers, each an NHWC [npix, C] map
 *   alpha           : [T,T] (channel_wise == 0) or [T,T,C] (channel_wise != 0)
 * ---------------------------------------------------------------------------------- */
int vmtl_xstitch_fwd(const float* const* x_host, float* const* y_host, const float* alpha, int T,
                     int64_t npix, int C, int channel_wise, int mode, void* stream);

size_t vmtl_xstitch_bwd_workspace_bytes(int T, int64_t npix, int C, int channel_wise);

/* dx[b] = sum_a alpha[a,b]*dy[a];  dalpha[a,b(,c)] = sum_pixels dy[a]*x[b].
 * In VMTL_XS_REFERENCE_DIAG mode only the diagonal is used / gets a gradient and the
 * off-diagonal entries of dalpha are written as exact zeros (matches the reference).
 * dx_host may be NULL (skip the input gradient). */
int vmtl_xstitch_bwd(const float* const* dy_host, const float* const* x_host,
                     float* const* dx_host, const float* alpha, float* dalpha, int T,
                     int64_t npix, int C, int channel_wise, int mode, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Cross-stitch unit fused with the tensor assembly in front of it at CSNet's decoder sites
 * (vision_mtl/models/cross_stitch_model.py:121-156, utils/model_utils.py:46-58): per task, the stitched map is
 *   up2 == 0 : cat([skip, zero_pad(x)], channel)   skip [B,Ho,Wo,Cs] (Cs may be 0), x [B,Hi,Wi,Cx] centred in
 *              the [Ho,Wo] frame (Ho >= Hi, Wo >= Wi), C = Cs + Cx
 *   up2 != 0 : nearest x2 up-sampling of x          (Cs == 0, Ho == 2 Hi, Wo == 2 Wi)
 * The kernels gather from (skip, x) and write y [B,Ho,Wo,C] per task; the concatenated / up-sampled tensor
 * never exists.  alpha is [T,T] or [T,T,C].  Backward: dskip / dx (entries may be NULL), dalpha. */
int vmtl_xstitch_cat_fwd(const float* const* skip_host, const float* const* x_host, float* const* y_host,
                         const float* alpha, int T, int B, int Ho, int Wo, int Cs, int Hi, int Wi, int Cx,
                         int up2, int channel_wise, int mode, void* stream);

size_t vmtl_xstitch_cat_bwd_workspace_bytes(int T, int B, int Ho, int Wo, int Cs, int Hi, int Wi, int Cx,
                                            int up2, int channel_wise);

int vmtl_xstitch_cat_bwd(const float* const* dy_host, const float* const* skip_host,
                         const float* const* x_host, float* const* dskip_host, float* const* dx_host,
                         const float* alpha, float* dalpha, int T, int B, int Ho, int Wo, int Cs, int Hi,
                         int Wi, int Cx, int up2, int channel_wise, int mode, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * MTAN attention gate:  y = s * sigmoid(BN(h @ W^T + bias)).
 * Replaces conv2 -> bn2 -> sigmoid -> mul at vision_mtl/models/mtan_model.py:71-75
 * (encoder) and :158-162 (decoder).
 *   h [M,K]  hidden activations (the reference uses K = 128)
 *   s [M,N]  shared features    (N % 4 == 0, N <= 1024)
 *   W [N,K], bias/gamma/beta/running_mean/running_var [N]
 * The tcgen05 kernels cover K == 128 with N == 32 or N % 64 == 0; any other shape (and
 * VMTL_GATE_FP32_FFMA) runs the CUDA-core contraction of the same library.
 * training != 0: batch statistics (biased var for normalisation, unbiased for the
 *   running update, momentum as nn.BatchNorm2d); z = h@W^T+bias is written to save_z and
 *   (mean, invstd) to save_mean/save_invstd for the backward.  running_* may be NULL.
 * training == 0: running statistics, save_* may be NULL (save_z == NULL: single fused pass on the
 *   tensor-core path, z is never stored).
 * h_coef (forward and backward): NULL, or [A1 | B1] ([2][K]) -- then `h` holds the PRE-activation c of the
 *   hidden layer (conv1's output, mtan_model.py:65-69 / :152-156) and the kernels rebuild
 *   h = max(A1 c + B1, 0), the folded bn1 + ReLU (coef of vmtl_bnrelu_fwd with y == NULL), while they convert
 *   their operand: the [M,K] hidden tensor is never written to or read from HBM.  Tensor-core shapes only
 *   (vmtl_gate_tc_supported); the backward's dh is then the gradient w.r.t. h (feed it, with c, to
 *   vmtl_bnrelu_bwd).
 * ---------------------------------------------------------------------------------- */
/* 1 when the tcgen05 kernels cover (K, N) */
int vmtl_gate_tc_supported(int K, int N);

/* backward != 0: size for vmtl_gate_bwd, else for vmtl_gate_fwd */
size_t vmtl_gate_workspace_bytes(int64_t M, int K, int N, int precision, int backward);

int vmtl_gate_fwd(const float* h, const float* h_coef, const float* s, const float* W, const float* bias,
                  const float* gamma, const float* beta, float* running_mean,
                  float* running_var, float momentum, float eps, int training, int precision,
                  int64_t M, int K, int N, float* y, float* save_z, float* save_mean,
                  float* save_invstd, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the training-mode gate (formulas: SURVEY Appendix B).
 * Inputs: dy, h, s, saved z/mean/invstd.  Outputs: dh [M,K], ds [M,N], dW [N,K],
 * dbias/dgamma/dbeta [N].  dh or ds may be NULL to skip them.  training == 0 uses
 * running statistics passed in save_mean/save_invstd (dz = gamma*invstd*du). */
int vmtl_gate_bwd(const float* dy, const float* h, const float* h_coef, const float* s, const float* z,
                  const float* W, const float* gamma, const float* beta, const float* save_mean,
                  const float* save_invstd, int training, int precision, int64_t M, int K,
                  int N, float* dh, float* ds, float* dW, float* dbias, float* dgamma,
                  float* dbeta, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * BatchNorm2d (+ ReLU, + 2x2 max-pool) on NHWC maps, batch or running statistics.
 * Replaces the batch_norm / relu / max_pool2d ATen kernels behind every conv -> BN -> ReLU chain of the
 * MTAN network: DoubleConv (vision_mtl/utils/model_utils.py:61-80), the attention modules' conv1/bn1/relu
 * (mtan_model.py:65-69, :152-156), conv3/bn3/relu(/maxpool) (mtan_model.py:77-81, :141-142) and
 * conv_out/bn_out/relu (mtan_model.py:165-167).
 *   x [M,C] (C % 4 == 0, C <= 1024), gamma/beta/running_* [C]
 *   training != 0: batch statistics (running_* updated with `momentum`, may be NULL); else running statistics
 *   relu != 0: y = max(bn(x), 0)
 *   y == NULL: statistics only -- save_mean / save_invstd / coef are produced, nothing is applied (the gate
 *              kernels consume x and fold max(A x + B, 0) into their operand conversion: vmtl_gate_*_pre)
 *   coef [2][C]: A = gamma*invstd, B = beta - mean*A, the fp32 coefficients every consumer folds
 * pool variants: x [B,H,W,C] -> y [B,H/2,W/2,C], nn.MaxPool2d(2) after the ReLU.
 * Backward: dx [M,C] (NULL to skip), dgamma, dbeta [C]; dy has the shape of y.
 *   conv_bias [C] (or NULL), training only: the bias of the convolution that produced x, which the caller did NOT
 *     add (x = conv(.) without bias).  Batch statistics cancel a per-channel shift exactly, so y is unchanged and
 *     only the running mean moves: running_mean <- (1-m) running_mean + m (mean(x) + conv_bias).  This removes the
 *     reference's bias-add pass over x and, in the backward, its reduction of dx over all pixels:
 *   dconv_bias [C] (or NULL): gradient of that bias = sum over pixels of dx = A (sum g - M c1), which is zero up to
 *     round-off under batch statistics (as in the reference) and A sum g under running statistics.
 * ---------------------------------------------------------------------------------- */
size_t vmtl_bnrelu_workspace_bytes(int64_t M, int C);

int vmtl_bnrelu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int training, int relu, int64_t M, int C,
                    float* y, float* save_mean, float* save_invstd, float* coef, const float* conv_bias,
                    void* workspace, size_t workspace_bytes, void* stream);

int vmtl_bnrelu_bwd(const float* dy, const float* x, const float* coef, const float* save_mean,
                    const float* save_invstd, int training, int relu, int64_t M, int C, float* dx,
                    float* dgamma, float* dbeta, float* dconv_bias, void* workspace, size_t workspace_bytes,
                    void* stream);

int vmtl_bnrelu_pool_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, int training, int relu, int B, int H,
                         int W, int C, float* y, float* save_mean, float* save_invstd, float* coef,
                         const float* conv_bias, void* workspace, size_t workspace_bytes, void* stream);

int vmtl_bnrelu_pool_bwd(const float* dy, const float* x, const float* coef, const float* save_mean,
                         const float* save_invstd, int training, int relu, int B, int H, int W, int C,
                         float* dx, float* dgamma, float* dbeta, float* dconv_bias, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Segmentation head fused with cross-entropy, argmax and the confusion matrix.
 * Replaces nn.Conv2d(32,C,1) (mtan_model.py:367-376,401-404), F.softmax + argmax
 * (lit_module.py:137-138), nn.CrossEntropyLoss (lit_module.py:31,123) and the
 * torchmetrics confusion statistics (lit_module.py:109-111).
 *   feat [P,Cin] NHWC (Cin % 4 == 0, Cin <= 64), W [C,Cin], b [C], target int64 [P]
 *   out (double[2]) : {sum of per-pixel losses over valid pixels, n_valid}
 *   loss (float[1]) : out[0]/out[1]   (mean over valid pixels; NaN when n_valid == 0)
 *   pred            : optional uint8 [P] argmax (lowest index wins ties)
 *   conf            : optional int64 [C,C], rows = target; ACCUMULATED (+=)
 * Pixels with target == ignore_index (or outside [0,C)) contribute to nothing.
 * ---------------------------------------------------------------------------------- */
size_t vmtl_loss_workspace_bytes(int64_t P);

int vmtl_head_ce_fwd(const float* feat, const float* W, const float* b, const int64_t* target,
                     int64_t P, int Cin, int C, int64_t ignore_index, double* out, float* loss,
                     uint8_t* pred, int64_t* conf, void* workspace, size_t workspace_bytes,
                     void* stream);

/* dfeat [P,Cin], dW [C,Cin], db [C] for loss = mean CE; logits are recomputed from feat.
 * gscale (float[1], device): upstream dL/dloss.  n_valid comes from out[1] of the forward. */
int vmtl_head_ce_bwd(const float* feat, const float* W, const float* b, const int64_t* target,
                     int64_t P, int Cin, int C, int64_t ignore_index, const double* fwd_out,
                     const float* gscale, float* dfeat, float* dW, float* db, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Same loss/argmax/confusion on precomputed logits (basic / csnet 3x3 heads,
 * basic_model.py:30-41, model_utils.py:125-130).  layout: VMTL_LAYOUT_NCHW
 * ([B,C,HW], P = B*HW) or VMTL_LAYOUT_NHWC ([P,C]). */
int vmtl_ce_logits_fwd(const float* logits, const int64_t* target, int64_t P, int64_t HW, int C,
                       int layout, int64_t ignore_index, double* out, float* loss, uint8_t* pred,
                       int64_t* conf, void* workspace, size_t workspace_bytes, void* stream);

int vmtl_ce_logits_bwd(const float* logits, const int64_t* target, int64_t P, int64_t HW, int C,
                       int layout, int64_t ignore_index, const double* fwd_out,
                       const float* gscale, float* dlogits, void* stream);

/* ------------------------------------------------------------------------------------
 * Depth head fused with sigmoid, the SILog loss moments and the depth error sums.
 * Replaces nn.Conv2d(32,1,1) (mtan_model.py:367-376), sigmoid+permute
 * (lit_module.py:139), SILogLoss.forward (losses.py:14-36) and MeanAbsoluteError
 * (lit_module.py:68,112).
 *   feat [P,Cin] (Cin == 1: feat already holds the depth logits, w/b ignored -> basic/csnet)
 *   out (double[8]) : {n_mask, sum g, sum g^2, sum|p-t| (all px), sum|p-t|/t (masked),
 *                      mean g, D = var_unbiased(g) + 0.15*mean(g)^2, P}
 *   scalars (float[3]) : {silog = 10*sqrt(D), mae = sum|p-t|/P, abs_rel}
 *   pred : optional float [P] = sigmoid(depth logit)
 * ---------------------------------------------------------------------------------- */
int vmtl_head_silog_fwd(const float* feat, const float* w, const float* b, const float* target,
                        int64_t P, int Cin, float min_depth, double* out, float* scalars,
                        float* pred, void* workspace, size_t workspace_bytes, void* stream);

/* dfeat [P,Cin], dw [Cin], db [1] for loss = silog; gscale: upstream dL/dsilog (device). */
int vmtl_head_silog_bwd(const float* feat, const float* w, const float* b, const float* target,
                        int64_t P, int Cin, float min_depth, const double* fwd_out,
                        const float* gscale, float* dfeat, float* dw, float* db,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Global-batch statistics for data-parallel training.
 * The reference is single-process: BatchNorm (mtan_model.py:72, model_utils.py:66-72) and SILog
 * (losses.py:32-36) reduce over the whole batch.  With the batch sharded over GPUs, every op above that has a
 * batch reduction in its middle is also exported as its two halves, so the caller can all-reduce (SUM, NCCL,
 * same stream) the fp64 moments between them; an N-GPU step then equals the single-GPU step on the
 * concatenated batch.  M_global = rows of all replicas.  Parameter gradients (dgamma, dbeta, dW, dbias) stay
 * this replica's contributions: the data-parallel wrapper averages them like every other gradient.
 *
 *   moments (double [2][C]):
 *     forward   (sum x, sum x^2)            backward  (sum g, sum g*xhat)   [= this replica's dbeta, dgamma]
 *
 * BatchNorm (+ReLU, +2x2 max-pool; pool != 0 needs y / the pooled dy as in vmtl_bnrelu_pool_*): x is [B,H,W,C].
 * ---------------------------------------------------------------------------------- */
int vmtl_bn_moments(const float* x, int64_t M, int C, double* moments, void* workspace, size_t workspace_bytes,
                    void* stream);

/* finalize from the all-reduced moments (running statistics updated with the global mean / unbiased variance)
 * + apply pass; y == NULL: coefficients only (folded hidden layer of the gate) */
int vmtl_bnrelu_fwd_global(const float* x, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float momentum, float eps, int relu, int pool, int B, int H, int W,
                           int C, const double* moments, int64_t M_global, float* y, float* save_mean,
                           float* save_invstd, float* coef, const float* conv_bias, void* workspace,
                           size_t workspace_bytes, void* stream);

int vmtl_bnrelu_bwd_moments(const float* dy, const float* x, const float* coef, const float* save_mean,
                            const float* save_invstd, int relu, int pool, int B, int H, int W, int C,
                            double* moments, void* workspace, size_t workspace_bytes, void* stream);

int vmtl_bnrelu_bwd_global(const float* dy, const float* x, const float* coef, const float* save_mean,
                           const float* save_invstd, int relu, int pool, int B, int H, int W, int C,
                           const double* moments, int64_t M_global, float* dx, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Gate forward, first half: contraction -> save_z, moments (sum z, sum z^2) [2][N]. */
int vmtl_gate_fwd_moments(const float* h, const float* h_coef, const float* W, const float* bias, int precision,
                          int64_t M, int K, int N, float* save_z, double* moments, void* workspace,
                          size_t workspace_bytes, void* stream);

/* second half: statistics from the all-reduced moments, y = s * sigmoid(BN(z)) */
int vmtl_gate_fwd_global(const float* s, const float* z, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float momentum, float eps, int precision, int64_t M, int K, int N,
                         const double* moments, int64_t M_global, float* y, float* save_mean, float* save_invstd,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Gate backward, first half: ds, this replica's sums -> moments (sum du, sum du*zhat) [2][N]; the workspace keeps
 * the dW pieces and must reach vmtl_gate_bwd_global untouched. */
int vmtl_gate_bwd_moments(const float* dy, const float* h, const float* h_coef, const float* s, const float* z,
                          const float* W, const float* gamma, const float* beta, const float* save_mean,
                          const float* save_invstd, int precision, int64_t M, int K, int N, float* ds,
                          double* moments, void* workspace, size_t workspace_bytes, void* stream);

int vmtl_gate_bwd_global(const float* dy, const float* h, const float* h_coef, const float* s, const float* z,
                         const float* W, const float* gamma, const float* beta, const float* save_mean,
                         const float* save_invstd, int precision, int64_t M, int K, int N, const double* moments,
                         int64_t M_global, float* dh, float* dW, float* dbias, float* dgamma, float* dbeta,
                         void* workspace, size_t workspace_bytes, void* stream);

/* SILog: `out` of vmtl_head_silog_fwd with out[0..4] and out[7] all-reduced -> out[5] (mean g), out[6] (D) and
 * the three scalars of the GLOBAL batch; feed that `out` to vmtl_head_silog_bwd with gscale multiplied by the
 * number of replicas (their gradients are averaged afterwards). */
int vmtl_silog_finalize(double* out, float* scalars, void* stream);

/* ------------------------------------------------------------------------------------
 * Validation reductions (torchmetrics call sites lit_module.py:106-118).
 * ---------------------------------------------------------------------------------- */
/* conf[target, pred] += 1 for every pixel with 0 <= target,pred < C and target != ignore_index.
 * pred_is_u8 != 0: pred is uint8 [P]; else int64 [P]. */
int vmtl_confusion_accum(const void* pred, int pred_is_u8, const int64_t* target, int64_t P, int C,
                         int64_t ignore_index, int64_t* conf, void* stream);

/* out (double[4]) = {P, sum|p-t|, n(t > min_depth), sum |p-t|/t over t > min_depth} */
int vmtl_depth_err_sums(const float* pred, const float* target, int64_t P, float min_depth,
                        double* out, void* workspace, size_t workspace_bytes, void* stream);

/* metrics (float[3]) = {accuracy (micro), jaccard (macro, absent = 0), F1 (weighted)}
 * from an int64 [C,C] confusion matrix (lit_module.py:48-67 configuration). */
int vmtl_seg_metrics(const int64_t* conf, int C, float* metrics, void* stream);

/* ------------------------------------------------------------------------------------
 * Bilinear x2 up-sampling, align_corners = True, NHWC -- nn.Upsample(scale_factor=2, mode="bilinear",
 * align_corners=True) of the decoder attention modules (vision_mtl/models/mtan_model.py:125, applied at :143-144
 * in front of torch.cat((conv1_shared, prev), dim=1), :145).  ATen's arithmetic (UpSampleBilinear2d.cu).
 *   x [B,Hi,Wi,C] dense (C % 4 == 0); the up-sampled side is addressed with a row stride (floats between consecutive
 *   pixels, >= C, % 4 == 0): y / dy may be the channel slice of a wider (concatenated) NHWC tensor, so the
 *   concatenation never copies the up-sampled half and its backward never repacks the slice.
 *   forward:  y[(b,oy,ox)*ldy + c], oy < 2 Hi, ox < 2 Wi.   backward: dx [B,Hi,Wi,C] dense, a deterministic gather.
 * ---------------------------------------------------------------------------------- */
int vmtl_up2_bilinear_fwd(const float* x, float* y, int B, int Hi, int Wi, int C, int64_t ldy, void* stream);
int vmtl_up2_bilinear_bwd(const float* dy, int64_t lddy, float* dx, int B, int Hi, int Wi, int C, void* stream);

/* Shape coverage of the fused heads: kind 0 = vmtl_head_ce_* (Cin, C), 1 = vmtl_head_silog_* (Cin),
 * 2 = vmtl_ce_logits_* (C).  Returns 1 when the sm_100a kernels cover the shape (else the calls return
 * VMTL_EUNSUPPORTED and the host side runs the 1x1 projection through cuDNN and the loss on its logits). */
int vmtl_head_supported(int kind, int Cin, int C);

/* ------------------------------------------------------------------------------------
 * Multi-tensor Adam: torch.optim.Adam(params, lr) of training_lit.py:51 / lit_module.py:194,
 * stepped at training_lit.py:87 -- every tensor of a parameter group in one launch (28 bytes per parameter).
 * Arithmetic of torch's _single_tensor_adam (no amsgrad, not maximize; weight_decay is the L2 form g += wd * p).
 *   params / grads / numel / moment_offset: HOST arrays with n_tensors entries -- device pointers of the fp32
 *     parameter and gradient tensors, their element counts, and each tensor's element offset (a multiple of 4)
 *     inside the flat moment buffers;
 *   exp_avg / exp_avg_sq: DEVICE, flat fp32 moment buffers (16-byte aligned), updated in place;
 *   lr_dev: DEVICE float[1] or NULL (then `lr`); hyper-parameters are doubles like torch's Python scalars (1 - beta
 *     is formed in double before it is rounded to fp32); step_dev: DEVICE float[1], steps taken so far, incremented by the
 *     call; ticket: DEVICE uint32[1], must be zero before the first call and is left zero.
 * Tables travel in the kernel parameter space (vmtl_adam_max_tensors_per_launch tensors per launch), so the call is
 * capture-safe without any host buffer outliving it.
 * ---------------------------------------------------------------------------------- */
int vmtl_adam_max_tensors_per_launch(void);
int vmtl_adam_step(const void* const* params, const void* const* grads, const int64_t* numel,
                   const int64_t* moment_offset, int n_tensors, float* exp_avg, float* exp_avg_sq,
                   const float* lr_dev, double lr, float* step_dev, unsigned int* ticket, double beta1, double beta2,
                   double eps, double weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMTL_B200_H_ */
