/* libegm_b200 -- C ABI of the B200-native EGM-UNet hot path (sm_100a only).
 *
 * The reference (feiyeha/EGM-Unet) is pure PyTorch and has no FFI of its own: every GPU
 * kernel it runs is an ATen / cuDNN / cuBLAS / cuFFT library kernel chosen by
 * torch.nn.  Each entry point below names the reference call site(s)
 * whose library kernels it replaces.  All functions:
 *   - take plain device pointers + sizes (no torch types), NHWC activations of `dtype`
 *     (EGM_F32 = fp32 check mode, EGM_BF16 = production), fp32 parameters/statistics;
 *   - launch asynchronously on `stream` (a cudaStream_t), never allocate or free device
 *     memory, never synchronise -- they are CUDA-graph capturable;
 *   - return EGM_OK or a negative EGM_E_* code; egm_last_error() (thread-local) has the text.
 * Channel-strided views: where a tensor has `*_cstride/*_coff` arguments, element (m, c)
 * lives at ptr[m * cstride + coff + c] (used to read/write slices of concatenated tensors).
 */
#ifndef EGM_B200_H
#define EGM_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define EGM_ABI_VERSION 1
enum { EGM_F32 = 0, EGM_BF16 = 1 };
enum { EGM_OK = 0, EGM_E_BADARG = -1, EGM_E_SHAPE = -2, EGM_E_ALIGN = -3, EGM_E_ARCH = -4, EGM_E_LAUNCH = -5 };

/* ---- library ---- */
int egm_abi_version(void);
const char* egm_last_error(void);
int egm_device_check(void);                       /* fails unless the current device is CC 10.x */
/* Kernel-boundary overlap (A/B switch): launch every kernel with programmatic stream serialization (each kernel lets its successor's
 * CTAs be scheduled early and waits for its predecessor's completion before its first global access).  Results are identical either
 * way; measured slower than plain stream order inside the captured step graph, so it is off by default.  Returns the previous
 * setting.  (The reference's eager ATen launches have no counterpart.) */
int egm_set_launch_overlap(int enabled);

/* ---- layout / glue (replaces ATen copy_/cat/split/add kernels around src/EGM-UNet.py:1527-1541) ---- */
int egm_nchw_to_nhwc(const float* x, void* y, int dtype, int N, int C, int H, int W, void* stream);
int egm_nhwc_to_nchw(const void* x, float* y, int dtype, int N, int C, int H, int W, void* stream);
int egm_copy_slice(const void* src, void* dst, int dtype, long long M, int C, long long s_cstride, long long s_coff,
                   long long d_cstride, long long d_coff, int accumulate, void* stream);
int egm_axpby(void* dst, const void* src, int dtype, long long n, float alpha, float beta, void* stream);
int egm_memset_zero(void* p, long long bytes, void* stream);
int egm_cast_from_f32(const float* src, void* dst, int dtype, long long n, void* stream);
int egm_cast_to_f32(const void* src, float* dst, int dtype, long long n, int accumulate, void* stream);

/* ---- convolution, CUDA-core implicit GEMM (nn.Conv2d stride 1 "same": src/EGM-UNet.py:49,52,962,1211-1215,1258-1292; cuDNN/cuBLAS today) ---- */
int egm_pack_conv_weight(const float* w, float* wf, float* wd, int Cout, int Cin_g, int kh, int kw, int groups, void* stream);
int egm_unpack_conv_wgrad(const float* dw_packed, float* dw, int Cout, int Cin_g, int kh, int kw, float beta, void* stream);
int egm_conv_weight_lift(float* w, float* wp, int Cout, int Cin_g, int groups, int taps, int CoutP, int CinP, int mode, void* stream);
int egm_kernel_embed(float* big, float* small_, long long CoCi, int kb, int ks, int mode, int accumulate, void* stream);
/* EdgeAwareFeatureEnhancer (src/EGM-UNet.py:872-886): x - AvgPool2d(3,1,1)(x) followed by the 1x1 conv is ONE 3x3 conv with
 * w3[co][ci][t] = w1[co][ci] * (delta[t] - 1/9).  mode 0 builds w3 from w1 (round_bf16: off-centre taps rounded to bf16, centre = -8 * that,
 * so the bf16 operand keeps the exact zero response to constant inputs); mode 1 reduces a w3-shaped gradient to dw1. */
int egm_highpass_compose(float* w1, float* w3, long long CoCi, int mode, int round_bf16, void* stream);
/* Batched form of the per-conv weight preparation / gradient extraction above (csrc/wbatch.cu): `jobs` is a DEVICE array of
 * n_jobs 160-byte EgmWJob records (layout documented in csrc/wbatch.cu and mirrored by engine.WeightPlan); one launch packs
 * every tcgen05 conv's bf16 forward/dgrad operands (+ padded / summed bias) from the fp32 master parameters of
 * nn.Conv2d.weight (src/EGM-UNet.py DoubleConv :44-55, BasicConv :958-975, FusionConv :1202-1236), one launch scatters every
 * packed fp32 weight gradient back into the parameters' .grad layout.  total_elems = sum over jobs of the element space. */
int egm_wjob_bytes(void);
int egm_weight_prep_batch(const void* jobs, int n_jobs, long long total_elems, void* stream);
int egm_wgrad_unpack_batch(const void* jobs, int n_jobs, long long total_elems, void* stream);
int egm_conv2d_direct(const void* x, long long x_cstride, long long x_coff, const float* w_packed, const float* bias, void* y,
                      long long y_cstride, long long y_coff, int accumulate, int dtype, int N, int H, int W, int Cin, int Cout,
                      int kh, int kw, int dil, int groups, void* stream);
int egm_conv2d_wgrad_direct(const void* x, long long x_cstride, long long x_coff, const void* dy, long long dy_cstride, long long dy_coff,
                            float* dw_packed, int dtype, int N, int H, int W, int Cin, int Cout, int kh, int kw, int dil, int groups,
                            void* stream);

/* ---- convolution, tcgen05 tensor-core implicit GEMM, bf16 NHWC, TMA-staged (DoubleConv: src/EGM-UNet.py:44-55, src/unet.py:7-18) ---- */
long long egm_conv2d_tc_workspace_bytes(int N, int H, int W, int Cin, int Cout, int kh, int dil);
int egm_conv2d_tc_supported(int Cin, int Cout, int kh, int kw, int dil, int groups);
int egm_pack_conv_weight_tc(const float* w, void* wf_bf16, void* wd_bf16, int Cout, int Cin, int kh, int kw, void* stream);
int egm_conv2d_tc(const void* x, const void* w_packed_bf16, const float* bias, void* y, int N, int H, int W, int Cin, int Cout,
                  int kh, int kw, int dil, void* stream);
int egm_conv2d_wgrad_tc(const void* x, const void* dy, float* dw_packed, int N, int H, int W, int Cin, int Cout, int kh, int kw, int dil,
                        void* stream);
/* the tcgen05 wgrad kernels leave dL/dW packed as fp32 [taps][Cout][Cin] (16-byte vector reductions from TMEM); this turns it into the
 * nn.Conv2d layout [Cout][Cin][kh][kw]: dw = beta*dw + unpacked */
int egm_unpack_conv_wgrad_tc(const float* dw_packed, float* dw, int Cout, int Cin, int kh, int kw, float beta, void* stream);
/* General forms: operands are channel-strided views (element (pixel p, channel c < valid) at ptr[p*cstride + coff + c], as in
 * egm_copy_slice).  Cin / Cout are the padded (multiple-of-16) channel counts of the packed weight; channels >= *_valid read
 * as zero through TMA out-of-bounds fill and are never written, so thin (C < 16), odd and channel-sliced tensors -- GRFB branch
 * convs on slices of cat(x, dir, edge, ctx), RGA split, 1/2/3-channel heads -- need no staging copies.  TMA-read views need
 * cstride % 8 == 0 and coff % 8 == 0; the output view may be arbitrary (unaligned rows fall back to 2-byte stores).
 * accumulate != 0: y += conv. */
int egm_conv2d_tc_view(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* w_packed_bf16, const float* bias,
                       void* y, long long y_cstride, long long y_coff, int cout_valid, int accumulate, int N, int H, int W, int Cin,
                       int Cout, int kh, int kw, int dil, void* stream);
/* Extended forward form -- BatchNorm / ReLU in the conv epilogue (BASELINE.json north_star subsystem 1; src/EGM-UNet.py:44-55 conv -> BN -> ReLU,
 * :958-975 BasicConv).
 *   relu != 0   y = relu(conv + bias): inference, with the BatchNorm folded into the packed weights (w * gamma*rstd via egm_scale_rows)
 *               and the bias (beta - mean*gamma*rstd) -- no BN kernel runs at all.
 *   stats       training: besides y (= the pre-BN tensor z) the kernel adds, per output channel, sum(z) and sum(z^2) of its fp32
 *               accumulators over all N*H*W pixels into stats[0..cout_valid) / stats[cout_valid..2*cout_valid) (fp64, zeroed by the
 *               caller); egm_bn_finalize turns them into scale / shift / running statistics.  Available where
 *               egm_conv2d_tc_stats_supported (padded Cout of 16 / 32 / 64: the layers with the large pre-BN maps). */
int egm_conv2d_tc_stats_supported(int Cin, int Cout, int kh, int kw, int dil);
int egm_conv2d_tc_stats_profitable(int Cin, int Cout, int kh, int kw, int dil);   /* supported AND faster than a separate statistics pass (measured) */
int egm_conv2d_tc_ex(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* w_packed_bf16, const float* bias,
                     void* y, long long y_cstride, long long y_coff, int cout_valid, int accumulate, int N, int H, int W, int Cin,
                     int Cout, int kh, int kw, int dil, int relu, double* stats, void* stream);
/* out[r][:] = w[r][:] * scale[r] (+ nothing else): folds an inference BatchNorm's per-output-channel scale into a conv weight [rows][cols] */
int egm_scale_rows(const float* w, const float* scale, float* out, int rows, long long cols, void* stream);
int egm_conv2d_wgrad_tc_view(const void* x, long long x_cstride, long long x_coff, int cin_valid, const void* dy, long long dy_cstride,
                             long long dy_coff, int cout_valid, float* dw_packed, int N, int H, int W, int Cin, int Cout, int kh, int kw,
                             int dil, void* stream);

/* ---- BatchNorm2d (+ReLU / sigmoid gate / residual) (nn.BatchNorm2d: src/EGM-UNet.py:50,53,879,966; ATen native_batch_norm today) ---- */
int egm_bn_stats(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, void* stream);
/* bn_stats + bn_finalize (training) in one launch: the last block to finish turns the sums into scale/shift/mean/rstd and
 * updates the running statistics.  `sums` must hold 2*C + 1 doubles (sums, sums of squares, block ticket counter). */
int egm_bn_stats_finalize(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* sums, const float* gamma,
                          const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                          float eps, float* scale, float* shift, float* mean, float* rstd, void* stream);
/* bn_act_bwd_reduce + bn_bwd_finalize (training) in one launch; `sums` holds 2*C + 1 doubles. */
int egm_bn_act_bwd_reduce_finalize(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                                   const float* mean, const float* rstd, int act, int mode, const void* aux, float alpha, int dtype,
                                   long long M, int C, double* sums, const float* gamma, float* coef, float* dgamma, float* dbeta,
                                   void* stream);
int egm_bn_finalize(const double* sums, long long M, const float* gamma, const float* beta, float* running_mean, float* running_var,
                    long long* num_batches_tracked, float momentum, float eps, int training, int C,
                    float* scale, float* shift, float* mean, float* rstd, void* stream);
int egm_bn_act_fwd(const void* z, long long z_cstride, long long z_coff, const float* scale, const float* shift, int act, int mode,
                   const void* aux, float alpha, void* y, long long y_cstride, long long y_coff, int dtype, long long M, int C, void* stream);
/* egm_bn_finalize (training) + egm_bn_act_fwd in one launch, for batch sums taken in a conv epilogue (egm_conv2d_tc_ex): every thread
 * derives the scale / shift of its channels from `sums` ([2][C] doubles over Mstat pixels); scale / shift / mean / rstd are still
 * written for the backward pass and the running statistics updated (nn.BatchNorm2d training semantics, src/EGM-UNet.py:50,53,966). */
int egm_bn_finalize_act_fwd(const double* sums, long long Mstat, const float* gamma, const float* beta, float* running_mean, float* running_var,
                            long long* num_batches_tracked, float momentum, float eps, float* scale, float* shift, float* mean, float* rstd,
                            const void* z, long long z_cstride, long long z_coff, int act, int mode, const void* aux, float alpha, void* y,
                            long long y_cstride, long long y_coff, int dtype, long long M, int C, void* stream);
int egm_bn_act_bwd_reduce(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                          const float* mean, const float* rstd, int act, int mode, const void* aux, float alpha, int dtype, long long M, int C,
                          double* sums, void* stream);
int egm_bn_bwd_finalize(const double* sums, long long M, const float* gamma, const float* rstd, int C, float* coef, float* dgamma,
                        float* dbeta, int training, void* stream);
int egm_bn_act_bwd_apply(const void* dy, long long dy_cstride, long long dy_coff, const void* z, const float* scale, const float* shift,
                         const float* mean, const float* rstd, const float* coef, int act, int mode, const void* aux, float alpha,
                         void* dz, void* daux, int daux_accumulate, int dtype, long long M, int C, void* stream);
/* out[c] = sum over pixels of x[., c] (conv bias gradient).  scratch: 2*C doubles. */
int egm_channel_sum(const void* x, int dtype, long long M, int C, long long cstride, long long coff, double* scratch, float* out, void* stream);

/* ---- down / up sampling (nn.MaxPool2d :908, nn.Upsample+F.pad+cat :931-947) ---- */
int egm_maxpool2x2_fwd(const void* x, void* y, int dtype, int N, int H, int W, int C, void* stream);
int egm_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int accumulate, int dtype, int N, int H, int W, int C, void* stream);
/* channel-strided forms: x (and dx) may be the first Cs channels of an Up level's concat buffer -- the skip connection is produced
 * straight into cat([skip, up]) (src/EGM-UNet.py:938-947 materialises the cat; here it is virtual) and pooled from there */
int egm_maxpool2x2_fwd_view(const void* x, long long x_cstride, long long x_coff, void* y, int dtype, int N, int H, int W, int C, void* stream);
int egm_maxpool2x2_bwd_view(const void* x, long long x_cstride, long long x_coff, const void* dy, void* dx, long long dx_cstride, long long dx_coff,
                            int accumulate, int dtype, int N, int H, int W, int C, void* stream);
int egm_upsample_concat_fwd(const void* skip, const void* low, void* out, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream);
int egm_upsample_concat_bwd_low(const void* dcat, void* dlow, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream);

int egm_deconv_weight_pack(float* w, float* wt, int Cin, int Cout, int mode, void* stream);      /* nn.ConvTranspose2d(k2,s2): src/unet.py:35-37 */
int egm_pixel_shuffle_concat_fwd(const void* skip, const void* z, void* out, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream);
int egm_pixel_shuffle_concat_bwd(const void* dcat, void* dz, int dtype, int N, int Hl, int Wl, int H, int W, int Cs, int Cu, void* stream);

/* ---- OutConv (src/EGM-UNet.py:952-956, src/unet.py:54-58) fused with the NHWC <-> NCHW change at the model boundary: fp32 logits out,
 * fp32 dlogits in (neither is rounded to the activation dtype); backward yields dL/dy, dL/dW [K][C] and dL/dbias [K] in one pass ---- */
int egm_outconv_supported(int C, int K);
int egm_outconv_fwd(const void* y, const float* w, const float* bias, float* logits, int dtype, int N, long long HW, int C, int K, void* stream);
int egm_outconv_bwd(const void* y, const float* w, const float* dlogits, void* dy, float* dw, float* dbias, int dtype, int N, long long HW, int C,
                    int K, void* stream);

/* ---- MCALayer (src/EGM-UNet.py:686-791, MCAGate :836-869) ---- */
long long egm_mca_vec_len(int N, int H, int W, int C);
long long egm_mca_vec_off_w(int N, int H);
long long egm_mca_vec_off_c(int N, int H, int W);
int egm_mca_stats(const void* x, int dtype, int N, int H, int W, int C, double* sums, void* stream);
int egm_mca_prod_sums(const void* a, const void* b, int dtype, int N, int H, int W, int C, double* sums, void* stream);
int egm_mca_gates(const double* sums, int N, int H, int W, int C, const float* w_h, const float* k_h, int ks_h, const float* w_w,
                  const float* k_w, int ks_w, const float* w_c, const float* k_c, int ks_c, float* gates, float* avg, float* stdv, void* stream);
int egm_mca_apply(const void* x, const float* gates, void* y, unsigned char* argidx, void* u_scratch, void* d2_scratch, int dtype, int N, int H,
                  int W, int C, void* stream);
int egm_mca_bwd_du(const void* x, const float* gates, const void* dy, const unsigned char* argidx, void* E_scratch, void* du, int dtype, int N,
                   int H, int W, int C, void* stream);
int egm_mca_gates_bwd(const double* dG, int N, int H, int W, int C, const float* gates, const float* avg, const float* stdv, const float* w_h,
                      const float* k_h, int ks_h, const float* w_w, const float* k_w, int ks_w, const float* w_c, const float* k_c, int ks_c,
                      float* coef_a, float* coef_b, float* dw_h, float* dk_h, float* dw_w, float* dk_w, float* dw_c, float* dk_c, void* stream);
int egm_mca_bwd_dx(const void* du, const void* x, const float* gates, const float* coef_a, const float* coef_b, void* dx, int dtype, int N,
                   int H, int W, int C, void* stream);
/* Fused forms (csrc/mca_fused.cu; C % 64 == 0): the whole blend of MCALayer.forward (:755-790) -- gating, 3x3 range, 3x3 local variance,
 * channel shuffle -- in ONE pass over x (row-walking CTAs, shared-memory neighbour exchange, fp32 intermediates), and its backward
 * (du from x, dy and the arg map) in one pass.  egm_mca_fwd replaces egm_mca_apply, egm_mca_bwd replaces egm_mca_bwd_du. */
int egm_mca_fused_supported(int C);
int egm_mca_fwd(const void* x, const float* gates, void* y, unsigned char* argidx, int dtype, int N, int H, int W, int C, void* stream);
int egm_mca_bwd(const void* x, const float* gates, const void* dy, const unsigned char* argidx, void* du, int dtype, int N, int H, int W, int C,
                void* stream);

/* ---- edge enhancer / FusionConv attention / GRFB tail / RGA gating (src/EGM-UNet.py:872-886, 1171-1236, 1289-1323, 458-547) ---- */
int egm_highpass3(const void* in, void* out, int accumulate, int dtype, int N, int H, int W, int C, void* stream);
int egm_pixel_dot(const void* a, const void* b, const float* cvec, float* out, int dtype, int N, long long HW, int C, void* stream);
int egm_sample_chan_dot(const void* a, const void* b, const float* pvec, float* out, int dtype, int N, long long HW, int C, void* stream);
int egm_mul_pixel_gate(const void* a, const void* g, void* y, int dtype, long long M, int C, int Gc, int mode, void* stream);
int egm_pixel_gate_bwd(const float* dot, const void* g, void* dg, int dtype, long long M, int Gc, int mode, void* stream);
int egm_unary(const void* a, const void* b, const float* scalar_dev, void* y, int dtype, long long n, int op, void* stream);
int egm_dot_all(const void* a, const void* b, float* out, int dtype, long long n, void* stream);
int egm_chan_meanmax(const void* s, float* mm, unsigned char* amax, int dtype, long long M, int C, void* stream);
int egm_sa_conv_fwd(const float* mm, const float* w, float* sa, int N, int H, int W, void* stream);
int egm_sa_conv_bwd(const float* dsa, const float* sa, const float* mm, const float* w, float* dmm, float* dw, int N, int H, int W, void* stream);   /* dmm or dw may be NULL: only the other half is computed */
int egm_gap_gmp(const void* f, float* avg, float* mx, int* arg, void* scratch, int dtype, int N, long long HW, int C, void* stream);
int egm_ca_mlp_fwd(const float* avg, const float* mx, const float* w0, const float* w2, float* ca, float* hid, int N, int C, int Cr, void* stream);
int egm_ca_mlp_bwd(const float* dca, const float* ca, const float* avg, const float* mx, const float* hid, const float* w0, const float* w2,
                   float* dw0, float* dw2, float* davg, float* dmx, int N, int C, int Cr, void* stream);
int egm_fuse_mix_fwd(const void* f, const void* s, const float* sa, const float* ca, void* t, int dtype, int N, long long HW, int C, void* stream);
int egm_fuse_mix_bwd_s(const void* dt, const float* sa, const float* ca, const float* dmm, const unsigned char* amax, void* ds, int dtype, int N,
                       long long HW, int C, void* stream);
int egm_fuse_df_finish(void* df, const void* dt, const float* davg, const float* dmx, const int* arg, int accumulate, int dtype, int N, long long HW, int C, void* stream);

/* ---- loss, optimiser, metrics (train_utils/train_and_eval.py:7-19, dice_coefficient_loss.py, train.py:113-118, distributed_utils.py:76-167) ---- */
long long egm_loss_workspace_bytes(int N, int C, int H, int W);
int egm_loss_fwd_bwd(const float* logits, const long long* target, const float* class_weight, int N, int C, int H, int W, int ignore_index,
                     int with_dice, float grad_scale, float* loss_out, float* dlogits, void* workspace, long long workspace_bytes, void* stream);
int egm_sgd_step(float* params, const float* grads, float* momentum_buf, long long n, const float* hyper_dev, void* stream);
int egm_eval_metrics(const float* logits, const long long* target, int N, int C, int H, int W, int ignore_index, long long* confmat,
                     double* dice_acc, void* stream);

/* ---- CLIPSeg-ensemble fusion downstream of the UNet logits (csrc/ensemble.cu) -----------------------------
 * Replaces, per validation image and alpha: F.interpolate(clip_logits, bilinear, align_corners=False) (eval_CLIPseg.py:885-888),
 * fused = clip + alpha*unet, torch.argmax, cv2.resize(INTER_NEAREST) and ConfusionMatrix.update/compute inside
 * search_best_alpha (eval_CLIPseg.py:656-748), and the final mask of eval_CLIPseg.py:901-912 / predict_CLIPseg.py:519-526.
 * clip_logits [C,hc,wc] and unet_logits [C,H,W] are fp32 NCHW planes of ONE image, label is uint8 [Hl,Wl] (values >= C are
 * ignored), confusion is [n_alpha][C][C] uint64 and is ACCUMULATED (zero it once, call per image); best = {alpha, mIoU}. */
int egm_ensemble_confusion(const float* clip_logits, int hc, int wc, const float* unet_logits, int H, int W, const unsigned char* label, int Hl,
                           int Wl, int num_classes, const double* alphas, int n_alpha, unsigned long long* confusion, void* stream);
int egm_ensemble_best_alpha(const unsigned long long* confusion, const double* alphas, int n_alpha, int num_classes, float* miou, double* best,
                            void* stream);
int egm_ensemble_predict(const float* clip_logits, int hc, int wc, const float* unet_logits, int H, int W, int num_classes, float alpha,
                         unsigned char* mask, int Ho, int Wo, void* stream);

/* ---- input pipeline on the device (csrc/pipeline.cu) ------------------------------------------------------
 * One decoded image -> one entry of the model-input batch: replaces transforms.py:30-110 (RandomResize [Pillow bilinear /
 * nearest], flips, RandomCrop + pad_if_smaller, ToTensor, Normalize) and my_dataset.py:106-133 (mask / 255, collate_fn padding).
 * img [H,W,3] / mask [H,W] are uint8 DEVICE buffers; the resample tables are DEVICE int32 arrays computed by the host exactly as
 * Pillow's precompute_coeffs / ImagingScaleAffine do (egm_unet_b200.data): hmin/hcnt [rw], hk [rw][hks] (null when rw == W),
 * vmin/vcnt [rh], vk [rh][vks] (null when rh == H), nnx [rw] / nny [rh] (null = identity).  The valid_h x valid_w region (the
 * crop, or the resized image in eval mode) starts at (top, left) of the zero-padded resized image; the rest of the out_h x out_w
 * plane gets the collate padding (0.0 / 255).  mean_std_host: 6 HOST floats.  out_img fp32 [3][out_h][out_w], out_tgt int64. */
int egm_input_transform(const unsigned char* img, const unsigned char* mask, int H, int W, int rh, int rw, const int* hmin, const int* hcnt,
                        const int* hk, int hks, const int* vmin, const int* vcnt, const int* vk, int vks, const int* nnx, const int* nny,
                        int hflip, int vflip, int top, int left, int valid_h, int valid_w, int out_h, int out_w, const float* mean_std_host,
                        float* out_img, long long* out_tgt, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGM_B200_H */
