/* libvfd_b200.so -- C-ABI of the B200-native vfd_gan training hot path.
 *
 * The reference (umaionigiri/vfd_gan) has no native layer: its hot path bottoms out in torch
 * library calls (nn.Conv3d / nn.BatchNorm3d / nn.AvgPool3d / nn.Upsample / nn.Dropout ...).
 * Each entry point below replaces one of those call sites; the citation is reference file:line.
 * The host side (the vfd_gan_b200 Python package) binds these with ctypes and registers them as torch.library
 * ops; see INTEGRATION.md.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer (plain void / float / double), sizes are plain integers;
 *    no torch types cross this boundary. `stream` is a cudaStream_t passed as void*.
 *  - Activations are channels-last bf16: element (n,d,h,w,c) lives at
 *    ((((n*D+d)*H+h)*W+w) * ld + c); `ld` >= channel count lets a tensor be a channel slice of a
 *    wider (concat) buffer. Channel counts are padded to a multiple of 8 and padded channels hold 0.
 *  - Every function only enqueues work on `stream`; none synchronises the device.
 *  - Return value 0 = success; otherwise an error code, and vfd_last_error() returns a
 *    thread-local message. There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef VFD_B200_H
#define VFD_B200_H

#ifdef __cplusplus
#define VFD_API extern "C" __attribute__((visibility("default")))
#else
#define VFD_API __attribute__((visibility("default")))
#endif

VFD_API const char* vfd_last_error(void);
VFD_API int vfd_abi_version(void);

/* ---- conv3d, stride 1, "same" zero padding, kernel extents in {1,3} ------------------------------
 * Replaces nn.Conv3d forward at models/spatiotempconv.py:49-50,59-60,63-64, conv_last at
 * models/mygannet.py:52,97 and nn.Conv2d at models/convlstm.py:36-40,46 (D = 1).
 * Implicit GEMM on tcgen05: out[v][n] = bias[n] + sum_{tap,c} x[v+tap][c] * w_packed[n][tap][c].
 *   x        bf16 channels-last, `cin` visible channels (multiple of 8), row pitch x_ld
 *   w_packed bf16 [w_rows][kd*kh*kw][cin_k]  (vfd_pack_weight), w_rows % 16 == 0, cin_k % kc == 0
 *   bias     fp32 [w_rows] or NULL
 *   out      bf16 (out_fp32 = 0) or fp32 (out_fp32 = 1) channels-last, row pitch out_ld;
 *            columns [0, out_cols) are written
 *   stats    optional double [2][stats_ld] (zeroed): the epilogue adds the per-channel sum and sum of
 *            squares of the stored bf16 values, i.e. the vfd_bn_stats result for the BatchNorm
 *            that follows, without re-reading the tensor (bf16 output only)
 *   kc       channel block per MMA K-slab: 16, 32 or 64 (selects the 32/64/128-byte swizzle)
 * The input gradient (dgrad) is the same call on dy with mode-1 packed weights. */
VFD_API int vfd_conv3d_fwd(const void* x, long long x_ld, int cin, const void* w_packed, int w_rows,
                           int cin_k, const float* bias, void* out, long long out_ld, int out_cols,
                           int out_fp32, double* stats, int stats_ld, int N, int D, int H, int W, int kd,
                           int kh, int kw, int kc, void* stream);

/* Weight gradient of the same conv (autograd of nn.Conv3d, reached from err_g.backward() /
 * err_d.backward() at models/mygannet.py:311,344):
 *   acc[tap][ci][co] += sum_v dy[v][co] * x[v+tap][ci]     (fp32, red.add across voxel splits)
 * layout 0: acc is [kd*kh*kw][ci_pad][co_pad]; layout 1: acc is [kd*kh*kw][co_pad][ci_pad] (the kernel that
 * puts the input channels on the GEMM M dimension; only where vfd_conv3d_wgrad_layout returns 1). acc must
 * be zeroed by the caller; cout / cin are the valid channel counts. */
VFD_API int vfd_conv3d_wgrad(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                             int cin, float* acc, int co_pad, int ci_pad, int layout, int N, int D, int H, int W,
                             int kd, int kh, int kw, void* stream);
/* Accumulator layout (0 or 1) the tcgen05 weight-gradient path wants for this geometry. */
VFD_API int vfd_conv3d_wgrad_layout(int cout, int cin, int kd, int kh, int kw, int H, int W);

/* Thin weight gradient (1 <= cin, cout <= 32): acc[ci][co] += sum_v x[v][ci] * dy[v][co], acc fp32
 * [ci_pad][co_pad], zeroed by the caller. A streaming warp-level mma.sync kernel: these layers are HBM
 * streams, not GEMMs. fold = 0: a plain 1x1x1 gradient (kd = kh = kw = 1). fold = 1 / 2: the x / dy side is
 * tap-folded on the fly (as vfd_tap_gather would: the tensor holds cs channels per voxel and the folded
 * channel count is kd*kh*kw*cs), so the gradient of a thin kd x kh x kw conv needs no im2col tensor.
 * Supported folds: x (1,3,3) cs 3, x (3,1,1) cs 2, dy (3,3,3) cs 1. */
VFD_API int vfd_conv3d_wgrad_thin(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                                  int cin, float* acc, int co_pad, int ci_pad, int fold, int cs, int N, int D,
                                  int H, int W, int kd, int kh, int kw, void* stream);

/* Forward of a 3x3x3 conv from 32 input channels to ONE output channel with fp32 output (NetG's conv_last,
 * models/mygannet.py:52,97): same operands and result as vfd_conv3d_fwd (row 0 of the packed weights, packed K = 32;
 * out column 0 = the logit, the remaining out_cols - 1 columns 0), but the 27 taps are the GEMM's N dimension and the
 * neighbourhood sum runs over fp32 partial products in shared memory (csrc/conv_narrow.cu). */
VFD_API int vfd_conv3d_fwd_narrow(const void* x, long long x_ld, int cin, const void* w_packed, int cin_k,
                                  const float* bias, float* out, long long out_ld, int out_cols, int N, int D, int H,
                                  int W, void* stream);

/* One ConvLSTM step (models/convlstm.py:46-58) in one launch: the gate conv (nn.Conv2d over cat[x, h], kd = 1) on
 * tcgen05 with sigmoid / tanh and c' = f*c + i*g, h' = o*tanh(c') in its epilogue; the gates never reach HBM.
 * comb: bf16 channels-last [N][H][W][comb_ld] (cin = in + hid channels); w_packed_perm: forward-packed gate weights
 * [4*hid][taps][cin_k] with the rows re-ordered [hid/64][gate i,f,o,g][64] (so one 256-column accumulator tile holds the
 * four gates of 64 hidden channels), bias_perm in the same order (or NULL); c_cur / c_next fp32 [N][H][W][hid]; h_out
 * bf16 [N][H][W][h_ld] (may be a channel slice of the next step's concat buffer); act (optional) fp32 [N][H][W][4*hid]
 * gate-major post-activation values for vfd_convlstm_cell_bwd. hid must be a multiple of 64. */
VFD_API int vfd_convlstm_step_fwd(const void* comb, long long comb_ld, int cin, const void* w_packed_perm, int cin_k,
                                  const float* bias_perm, const float* c_cur, int hid, float* c_next, void* h_out,
                                  long long h_ld, float* act, int N, int H, int W, int kh, int kw, int kc,
                                  void* stream);

/* dgrad of the same layer: g bf16 channels-last [N][D][H][W][g_ld] whose column 0 is the gradient of the logit,
 * w_dgrad_packed = the dgrad operand vfd_conv3d_fwd takes for this layer ([32][27 taps, mirrored][w_ck], column 0 used),
 * dx bf16 [N][D][H][W][dx_ld] (32 channels). Same result as vfd_conv3d_fwd on (g, w_dgrad_packed) up to summation order. */
VFD_API int vfd_conv3d_dgrad_narrow(const void* g, long long g_ld, const void* w_dgrad_packed, int w_rows, int w_ck,
                                    void* dx, long long dx_ld, int N, int D, int H, int W, void* stream);

/* weight gradient of the same layer: acc[c][tap] += sum_u g[u - off(tap)] * x[u][c] (fp32 [32][acc_ld], the layout
 * vfd_conv3d_wgrad_thin fills for the tap-folded gradient), g and x each read once. Uses fp32 atomics across CTAs; the
 * deterministic mode keeps vfd_tap_gather + vfd_conv3d_wgrad_thin_det for this layer. */
VFD_API int vfd_conv3d_wgrad_narrow(const void* g, long long g_ld, const void* x, long long x_ld, float* acc, int acc_ld,
                                    int N, int D, int H, int W, void* stream);

/* weight gradient of the first layers (1x3x3 over three input channels, cout <= 32; NetG / SDisc dconv1 spatial
 * convs): acc[tap * 3 + c][co] += sum_v dy[v][co] * x[v + off(tap)][c] (fp32 [32][acc_ld], the layout of the tap-folded
 * path), x and dy read once, no folded scratch tensor. fp32 atomics across CTAs; the deterministic mode keeps
 * vfd_tap_gather + vfd_conv3d_wgrad_thin_det. */
VFD_API int vfd_conv3d_wgrad_first(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, float* acc,
                                   int acc_ld, int N, int D, int H, int W, void* stream);

/* Deterministic variants (opt-in, VFD_DETERMINISTIC=1 / ops.set_deterministic): the voxel-range splits (thin
 * kernels: the blocks) keep their own partial accumulators in `workspace` instead of meeting in fp32 atomics, and an
 * ordered second pass adds them to acc. Same arguments and accumulator layout as the plain entry points; workspace =
 * 16-byte aligned device scratch of at least the matching *_workspace(...) bytes. */
VFD_API long long vfd_conv3d_wgrad_det_workspace(int cout, int cin, int co_pad, int ci_pad, int layout, int N, int D,
                                                 int H, int W, int kd, int kh, int kw);
VFD_API int vfd_conv3d_wgrad_det(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld, int cin,
                                 float* acc, int co_pad, int ci_pad, int layout, int N, int D, int H, int W, int kd,
                                 int kh, int kw, void* workspace, long long ws_bytes, void* stream);
VFD_API long long vfd_conv3d_wgrad_thin_det_workspace(int cout, int cin, int fold, int cs, int N, int D, int H, int W,
                                                      int kd, int kh, int kw);
VFD_API int vfd_conv3d_wgrad_thin_det(const void* dy, long long dy_ld, int cout, const void* x, long long x_ld,
                                      int cin, float* acc, int co_pad, int ci_pad, int fold, int cs, int N, int D,
                                      int H, int W, int kd, int kh, int kw, void* workspace, long long ws_bytes,
                                      void* stream);

/* ---- layout / weight packing -------------------------------------------------------------------
 * fp32 NCDHW [N][Csrc][S] -> bf16 channels-last [N][S][ld] (Cp channels, zero beyond C).
 * replicate = 1 repeats the source channels (gray2rgb, lib/utils.py:91-92). */
VFD_API int vfd_pack_ncdhw(const float* src, void* dst, int N, int Csrc, long long S, int C,
                           long long ld, int Cp, int replicate, void* stream);
/* channels-last bf16 (src_fp32 = 0) or fp32 (1) [N][S][ld] -> fp32 NCDHW [N][C][S] */
VFD_API int vfd_unpack_ncdhw(const void* src, int src_fp32, float* dst, int N, int C, long long S,
                             long long ld, void* stream);
/* fp32 nn.Conv3d weight [Cout][Cin][taps] -> bf16 GEMM operand [rows][taps][ck];
 * mode 0: forward (row = cout, col = cin); mode 1: dgrad (row = cin, col = cout, taps mirrored). */
VFD_API int vfd_pack_weight(const float* w, void* w_packed, int Cout, int Cin, int taps, int rows,
                            int ck, int mode, void* stream);
/* The same packing for many weights in one launch. jobs: device array of njobs records
 *   { const float* w; void* dst; int cout, cin, taps, rows, ck, mode; long long begin; }   (48 bytes)
 * where begin is the running sum of rows*taps*ck over the preceding jobs and total the overall sum. */
VFD_API int vfd_pack_weights_batched(const void* jobs, int njobs, long long total, void* stream);
/* wgrad accumulator [taps][ci_pad][co_pad] -> fp32 weight gradient [Cout][Cin][taps] (accumulate != 0: added to gw) */
VFD_API int vfd_unpack_wgrad(const float* acc, float* gw, int Cout, int Cin, int taps, int co_pad,
                             int ci_pad, int accumulate, void* stream);

/* ---- BatchNorm3d + (Leaky)ReLU (+ AvgPool3d, + Dropout) ----------------------------------------
 * Replaces nn.BatchNorm3d + nn.ReLU (models/spatiotempconv.py:51-52,63), nn.BatchNorm3d +
 * nn.LeakyReLU(0.2) / nn.LeakyReLU() (models/mygannet.py:19-20,25-26,109-110,114-115), the
 * nn.AvgPool3d that always follows them (models/mygannet.py:41,132,174) and nn.Dropout(p=0.25)
 * (models/mygannet.py:49,76,81,86,91).
 * vfd_bn_stats accumulates per-channel sum / sum of squares into `sums` (double [2*C], must be
 * zero on entry; vfd_bn_finalize clears it again). */
VFD_API int vfd_bn_stats(const void* x, long long ld, int C, long long V, double* sums, void* stream);
/* sums -> mean / invstd / scale (= gamma*invstd) / shift (= beta - mean*scale), fp32 [C] each, and
 * the running-stat update (momentum, unbiased variance). train = 0 uses the running statistics.
 * pre_bias (fp32 [Cvalid] or NULL) is the producing conv's bias when it was deliberately not added
 * to the stored tensor: it cancels in the normalisation and only enters running_mean / shift. */
VFD_API int vfd_bn_finalize(double* sums, int C, int Cvalid, long long V, const float* pre_bias,
                            const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum,
                            float eps, int train, float* mean, float* invstd, float* scale,
                            float* shift, void* stream);
/* out = dropout(act(y*scale + shift)), act(z) = z > 0 ? z : slope*z. out_full (full resolution)
 * and out_pool (average over pd x ph x pw windows, floor semantics) are optional (NULL to skip).
 * drop_p = 0 disables dropout; the mask is Philox4x32-10 keyed by (seed, voxel, channel group).
 * seed_dev (optional device pointer) is added to seed inside the kernel, so a captured CUDA graph draws a
 * fresh mask at every replay from a counter the graph itself advances. 0 <= slope <= 1. */
VFD_API int vfd_bn_act_fwd(const void* y, long long y_ld, int N, int D, int H, int W, int C,
                           const float* scale, const float* shift, float slope, void* out_full,
                           long long full_ld, void* out_pool, long long pool_ld, int pd, int ph, int pw,
                           float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                           void* stream);
/* Backward of the above through BatchNorm: given the gradients of out_full / out_pool (either may
 * be NULL) computes dy (bf16), dgamma and dbeta (fp32 [Cvalid]). sums: zeroed double [2*C] scratch,
 * c1 / c2: fp32 [C] scratch. train: bit 0 = training-mode BatchNorm (batch statistics); bit 1 = g_pool is
 * [N][D/pd][C] and broadcast over the H and W axes (the gradient of the global spatial mean of TDisc,
 * models/mygannet.py:175,189-191), only for windows (1,1,1) and (2,1,1); bit 2 = add to dgamma / dbeta instead of
 * overwriting them (a BatchNorm applied twice per step, NetD on the real and on the generated clip).
 * ticket (optional): a zeroed device counter; with it the reduce pass's last block derives dgamma / dbeta / c1 / c2
 * itself (and leaves sums and the counter zero again) instead of a separate finalize launch. One counter must not be
 * shared by calls that may run concurrently on different streams. */
VFD_API int vfd_bn_act_bwd(const void* y, long long y_ld, int N, int D, int H, int W, int C, int Cvalid,
                           const float* mean, const float* invstd, const float* scale,
                           const float* shift, float slope, const void* g_full, long long gf_ld,
                           const void* g_pool, long long gp_ld, int pd, int ph, int pw, float drop_p,
                           unsigned long long seed, const unsigned long long* seed_dev, int train,
                           double* sums, float* c1, float* c2, float* dgamma, float* dbeta, void* dy,
                           long long dy_ld, unsigned int* ticket, void* stream);
/* Tap folding for the weight gradient of thin convs (autograd of nn.Conv3d, models/mygannet.py:311,344):
 * dst[v][t*cs + c] = src[v + sign*offset(t)][c] for the kd*kh*kw taps t and c < cs (zero outside the volume,
 * remaining columns zero; dst_cols = taps*cs rounded up to 8, <= 32). With taps folded into channels,
 * vfd_conv3d_wgrad of a 1x1x1 kernel on (dy, dst) (sign +1, src = x) or on (dst, x) (sign -1, src = dy)
 * yields every tap's gradient from dense MMAs. */
VFD_API int vfd_tap_gather(const void* src, long long src_ld, int cs, void* dst, long long dst_ld, int dst_cols,
                           int N, int D, int H, int W, int kd, int kh, int kw, int sign, void* stream);
/* out[c] += sum_v x[v][c]  (conv bias gradient when no BatchNorm follows); double accumulators: the blocks' fp32
 * partial sums add up exactly, so the result does not depend on the order of the atomics */
VFD_API int vfd_channel_sum(const void* x, long long ld, int C, long long V, double* out, void* stream);

/* ---- nn.Upsample(scale_factor=2, 'trilinear', align_corners=True) + torch.cat -------------------
 * (models/mygannet.py:50,77-94). Forward writes straight into a channel slice of the concat
 * buffer (out_ld = total channels); backward gathers the gradient of the low-resolution input. */
VFD_API int vfd_upsample2x_fwd(const void* x, long long x_ld, int N, int D, int H, int W, int C,
                               void* out, long long out_ld, void* stream);
/* workspace (optional): N*2D*2H*W*C*2 + N*2D*H*W*C*2 bytes of device scratch enable the separable
 * three-pass adjoint (W, H, D axes; bf16 temporaries); without it a direct gather kernel runs. */
VFD_API int vfd_upsample2x_bwd(const void* gout, long long go_ld, int N, int D, int H, int W, int C,
                               void* gx, long long gx_ld, void* workspace, long long ws_bytes, void* stream);

/* ---- heads and losses ---------------------------------------------------------------------------
 * nn.Sigmoid after conv_last (models/mygannet.py:53,99): logits fp32 [V][ld] column 0 -> predict. */
VFD_API int vfd_sigmoid_head_fwd(const float* logits, long long ld, long long V, float* predict,
                                 void* stream);
VFD_API int vfd_sigmoid_head_bwd(const float* gpred, const float* predict, long long V, void* dlogit,
                                 void* stream);
/* weighted_bce (lib/utils.py:65-71): *loss_sum += sum(t*log p + pos_weight*(1-t)*log(1-p)) (the
 * caller negates and divides by V); gpred (optional) = grad_scale * d(-sum)/dp. */
VFD_API int vfd_weighted_bce(const float* predict, const float* target, long long V, float pos_weight,
                             float grad_scale, double* loss_sum, float* gpred, void* stream);
/* l2_loss numerator (lib/utils.py:59-63): *out += sum((a-b)^2) over channels-last bf16 tensors */
VFD_API int vfd_sqdiff(const void* a, long long a_ld, const void* b, long long b_ld, int C, long long V,
                       double* out, void* stream);

/* ---- ConvLSTM cell update (models/convlstm.py:49-58) ---------------------------------------------
 * gates fp32 channels-last [V][g_ld] in split order i,f,o,g; c/h fp32 [V][hid]; act (optional)
 * receives the activated gates [V][4*hid] for the backward pass. */
VFD_API int vfd_convlstm_cell_fwd(const float* gates, long long g_ld, const float* c_cur, int hid,
                                  long long V, float* h_next, float* c_next, float* act, void* stream);
VFD_API int vfd_convlstm_cell_bwd(const float* act, const float* c_cur, const float* c_next,
                                  const float* dh, const float* dc_in, int hid, long long V,
                                  void* dgates, long long dg_ld, float* dc_cur, void* stream);

/* ---- scoring: anomaly score, latent / contextual losses, evaluation (vfd_gan_b200/csrc/scoring.cu) ----
 * per_clip[n] += sum over clip n of (a - b)^2; a, b channels-last bf16 latents [N * rows_per_clip][ld].
 * Replaces torch.mean(torch.pow(latent_i - latent_o, 2), dim=1) (models/ganomaly.py:372) generalised to the
 * 3-D latent (mean over every non-batch dim); the caller zeroes per_clip. The sum over all
 * clips is the numerator of the latent L2 loss l_enc (models/ganomaly.py:439,477). */
VFD_API int vfd_latent_score(const void* a, long long a_ld, const void* b, long long b_ld, int C,
                             long long rows_per_clip, int N, double* per_clip, void* stream);
/* gradient of scale * (*gscale) * sum((a-b)^2): ga = 2*scale*(*gscale)*(a-b), gb = -ga (bf16 channels-last;
 * either may be NULL; gscale is an optional device scalar, the upstream autograd gradient) */
VFD_API int vfd_sqdiff_bwd(const void* a, long long a_ld, const void* b, long long b_ld, int C, long long V,
                           const float* gscale, float scale, void* ga, long long ga_ld, void* gb, long long gb_ld,
                           void* stream);
/* nn.L1Loss numerator, the contextual loss l_con (models/ganomaly.py:438,476): *sum += sum|a-b| over fp32
 * tensors; ga (optional) = grad_scale * sign(a-b) */
VFD_API int vfd_l1_loss(const float* a, const float* b, long long V, float grad_scale, double* sum, float* ga,
                        void* stream);
/* nn.BCELoss numerator (models/mygannet.py:267,326-332; lib/train_stcnn.py:90,107): *sum += -(t*max(log p,-100)
 * + (1-t)*max(log(1-p),-100)); gp (optional) = grad_scale * (p-t)/max(p*(1-p),1e-12) */
VFD_API int vfd_bce_loss(const float* p, const float* t, long long V, float grad_scale, double* sum, float* gp,
                         void* stream);
/* scores[i] = per_clip[i] * inv_count (fp32); minmax (optional, device float[2] pre-set to +inf/-inf by the
 * caller) is updated with the running min / max of the sweep */
VFD_API int vfd_score_finalize(const double* per_clip, int n, double inv_count, float* scores, float* minmax,
                               void* stream);
/* (s - min) / (max - min) over the whole sweep (models/ganomaly.py:396); minmax is device float[2] */
VFD_API int vfd_score_scale(const float* scores, long long n, const float* minmax, float* out, void* stream);
/* threshold (lib/utils.py:149-152) + morphology_proc (lib/utils.py:139-147): t_out (optional) = predict > thr;
 * m_out = 5x5 opening of t in the (D, H) plane for every (clip, w) -- what cv2.morphologyEx does with the
 * (D, H, W) array the reference hands it (rows = D, cols = H, channels = W; W <= 512). fp32 [N][D][H][W]. */
VFD_API int vfd_threshold_open(const float* predict, int N, int D, int H, int W, float thr, float* t_out,
                               float* m_out, void* stream);
/* counts[0..3] += TP, FP, FN, TN with label = labels > 0.5 and prediction = scores >= thr: the inputs of the
 * roc / pr / f1_score metrics of lib/evaluate.py:14-91 when the scores are binary masks (MyGAN.test) */
VFD_API int vfd_confusion_counts(const float* labels, const float* scores, long long n, float thr,
                                 unsigned long long* counts, void* stream);
/* exact tie-aware ROC area of n <= 16384 (score, label) pairs = sklearn auc(roc_curve(...))
 * (lib/evaluate.py:37-38) and the precision-recall area auc(recall, precision) of precision_recall_curve
 * (lib/evaluate.py:67-68); out = device double[4]: ROC area, #positives, #negatives, PR area */
VFD_API int vfd_roc_auc(const float* scores, const float* labels, int n, double* out, void* stream);
/* the same four numbers for any n < 2^31 (the voxel-level evaluation of test.py:175-202 and
 * models/mygannet.py:444-470: every voxel of the test set is a (score, label) pair): stable multi-block radix sort of
 * (score, label) keys, negative prefix counts, one pass over the runs of equal scores. Exact integer pair counts for the
 * ROC area, blocks summed in index order for the PR area (deterministic). workspace: 256-byte aligned device scratch
 * of at least vfd_roc_auc_large_workspace(n) bytes (about 16 bytes per pair). */
VFD_API long long vfd_roc_auc_large_workspace(long long n);
VFD_API int vfd_roc_auc_large(const float* scores, const float* labels, long long n, double* out, void* workspace,
                              long long ws_bytes, void* stream);

/* ---- video_to_flow (lib/utils.py:94-129; called at models/mygannet.py:281-282,404-405) on the device -----
 * video fp32 [B][3][D][H][W] in [-1, 1] -> out fp32 [B][3][D][H][W] in [-1, 1]: per-frame-index normalize over the
 * batch, RGB2GRAY, cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0) for every frame pair,
 * cartToPolar / HSV (S = 255) / HSV2RGB on float32 / np.uint8 wrap / ClipToTensor / *2-1, last frame repeated.
 * raw_flow (optional) receives the Farneback fields [B][D-1][H][W][2]. workspace: 256-byte aligned device scratch of
 * at least vfd_video_to_flow_workspace(B, D, H, W) bytes. */
VFD_API long long vfd_video_to_flow_workspace(int B, int D, int H, int W);
VFD_API int vfd_video_to_flow(const float* video, int B, int D, int H, int W, float* out, float* raw_flow,
                              void* workspace, long long ws_bytes, void* stream);

/* ---- clip pipeline from decoded uint8 frames (lib/data.py:14-161) --------------------------------------------
 * vfd_resize_frames_u8: Resize((isize, isize)) of the test transform (test.py:150-153, lib/data.py:143-146; the
 * reference's videotransforms/functional.py:54-58 selects PIL.Image.BILINEAR): Pillow's antialiased two-pass
 * resample on 8-bit channels, bit-exact. src uint8 [n][Hin][Win][C] -> dst uint8 [n][Hout][Wout][C], C in {1, 3}.
 * workspace: 256-byte aligned device scratch of at least vfd_resize_frames_u8_workspace(...) bytes.
 * vfd_frames_to_clip: ClipToTensor (videotransforms/volume_transforms.py:17-58) and the `*2-1` of lib/data.py:78:
 * frames uint8 [B][T][H][W][C] -> out float32 [B][Cout][T][H][W] = x / 255 (pm1 = 1: then 2x - 1); C == Cout, or
 * C == 1 broadcast over Cout channels (an 'L' mask through ClipToTensor(channel_nb=3)). */
VFD_API long long vfd_resize_frames_u8_workspace(long long n, int Hin, int Win, int C, int Hout, int Wout);
VFD_API int vfd_resize_frames_u8(const void* src, long long n, int Hin, int Win, int C, void* dst, int Hout,
                                 int Wout, void* workspace, long long ws_bytes, void* stream);
VFD_API int vfd_frames_to_clip(const void* frames, long long B, int T, int H, int W, int C, int Cout, int pm1,
                               float* out, void* stream);

#endif /* VFD_B200_H */
