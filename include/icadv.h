/* icadv.h -- C ABI of libicadv_b200.so: the sm_100a CUDA implementation of the per-image
 * adversarial-perturbation hot path of tongxyh/ImageCompression_Adversarial.
 *
 * Boundary contract:
 *   - plain C, raw DEVICE pointers + sizes + a cudaStream_t (passed as void*); no torch types;
 *   - the caller owns every buffer; the library allocates nothing persistent except plan objects
 *     (tensor maps + launch geometry) that the caller creates and destroys;
 *   - every entry point returns 0 or a negative ICADV_E* code; icadv_last_error() gives the text;
 *   - all work is stream-ordered and CUDA-graph capturable (no host synchronisation inside);
 *   - reductions are fixed-order (deterministic); no floating-point atomics;
 *   - activations are channels-last (NHWC) fp32; GEMM-shaped work runs as tcgen05 kind::tf32 with
 *     fp32 accumulation in TMEM (the reference's GPU path is cuDNN with TF32 allowed);
 *   - there is no CPU path: without an sm_100 device every compute entry point fails.
 *
 * Each entry point cites the reference interface (file:line under /root/reference) it replaces.
 */
#ifndef ICADV_H_
#define ICADV_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICADV_OK 0
#define ICADV_EINVAL (-1)   /* bad argument / unsupported shape */
#define ICADV_ECUDA (-2)    /* CUDA runtime / driver error */
#define ICADV_EARCH (-3)    /* device is not sm_100 */
#define ICADV_ENOMEM (-4)

typedef void* icadv_stream_t; /* cudaStream_t */

const char* icadv_last_error(void);
int icadv_version(void);
/* 0 if the current device is sm_100 (B200), ICADV_EARCH otherwise */
int icadv_check_device(void);

/* ------------------------------------------------------------------------------------------
 * Convolution-shaped contractions: nn.Conv2d / nn.ConvTranspose2d of the codec stacks
 *   net.g_a / net.g_s / net.h_a / net.h_s   (anchors/utils.py:112-130, attack_rd.py:344,349)
 * forward AND input-gradient (autograd of attack_rd.py:547), with GDN / IGDN
 * (utils/ops.py:58-97; compressai.layers.GDN) fused into the epilogue in both directions.
 *
 * Two geometric forms cover fwd and dgrad of both layer types (stride s, kernel k, pad k/2):
 *   ICADV_FORM_SCONV  out[n,oh,ow,:] = sum_taps in[n, s*oh+kh-p, s*ow+kw-p, :] * W[tap]
 *                     (Conv2d forward; ConvTranspose2d input-gradient)
 *   ICADV_FORM_TCONV  out[n, s*i+kh-p, s*j+kw-p, :] += in[n,i,j,:] * W[tap]      (output_padding s-1)
 *                     (ConvTranspose2d forward; Conv2d input-gradient).  Output pixels no tap reaches (k < s: the
 *                     input gradient of a strided 1x1 conv) are NOT written: the caller pre-fills them (bias / zero).
 * W is pre-packed [k*k taps][n_ch][k_ch] fp32 (icadv_pack_weight).
 * ------------------------------------------------------------------------------------------ */
#define ICADV_FORM_SCONV 0
#define ICADV_FORM_TCONV 1

#define ICADV_EPI_LINEAR 0   /* out = acc + bias, optional activation */
#define ICADV_EPI_GDN_FWD 1  /* x = acc+bias; sc = rsqrt(beta + G x^2); out = x*sc; out_scale = sc */
#define ICADV_EPI_IGDN_FWD 2 /* ... sc = sqrt(...) */
#define ICADV_EPI_GDN_BWD 3  /* g = acc; t = g*y*sc^2; s = G^T t; out = g*sc - (y/sc)*s */
#define ICADV_EPI_IGDN_BWD 4 /* t = g*y/sc^2;        out = g*sc + (y/sc)*s */

#define ICADV_ACT_NONE 0
#define ICADV_ACT_RELU 1
#define ICADV_ACT_LEAKY 2 /* slope 0.01 */
#define ICADV_ACT_ABS 3   /* used for h_a(|y|), anchors/balle.py:38 */

typedef struct {
  int form;            /* ICADV_FORM_* */
  int ksize, stride;   /* 5/2 (codec stacks), 3/1, 3/2, 1/1, 5/1 */
  int n_img;           /* images in the buffers */
  int in_h, in_w;      /* spatial size of `in` */
  int k_ch, n_ch;      /* input / output channels of this contraction */
  const float* in;     /* [n_img, in_h, in_w, k_ch] */
  const float* wpack;  /* [k*k][n_ch][k_ch] */
  const float* bias;   /* [n_ch] or NULL */
  float* out;          /* [n_img, out_h, out_w, n_ch] */
  int epi;             /* ICADV_EPI_* */
  int act;             /* ICADV_ACT_* (EPI_LINEAR only) */
  const float* gmat;   /* [n_ch][n_ch]: effective gamma (FWD) or its transpose (BWD) */
  const float* beta;   /* [n_ch] effective beta (FWD) */
  float* out_scale;    /* FWD: sc, same shape as out */
  const float* y_prev; /* BWD: saved out of the matching FWD */
  const float* sc_prev;/* BWD: saved out_scale of the matching FWD */
  int acc_from_in;     /* 1: no contraction, acc := in (1x1, k_ch == n_ch); GDN/IGDN as a stand-alone op */
  int round_out_tf32;  /* 1: round `out` to TF32 (nearest) when stored -- set when its consumer is a tensor-path
                          contraction, so operands are rounded once where produced instead of truncated by the MMA */
  int in_pad4;         /* 1: `in` is the padded RGB0 layout [n_img][in_h+4][in_w+8][4] (pixel (h,w) at (h+2,w+2),
                          zero border, 4th channel 0) written by icadv_pad_rgb4; first-layer form: SCONV 5x5/2,
                          k_ch = 3, wpack from icadv_pack_weight_rgb */
  const int* active;   /* optional [n_img] image indirection (device) */
  const int* n_active; /* optional device scalar: number of valid entries of `active` */
} icadv_conv_desc;

/* output spatial size implied by a descriptor */
int icadv_conv_out_hw(const icadv_conv_desc* d, int* out_h, int* out_w);

/* tcgen05/TMEM/TMA implicit-GEMM path.  Requires k_ch % 32 == 0, n_ch % 32 == 0 (n_ch > 256 is tiled over N for the
 * linear epilogue; GDN epilogues need 2 * n_ch <= 512 TMEM columns), or one of the two RGB end-layer forms (in_pad4;
 * transposed conv with n_ch <= 4).  A plan caches the TMA tensor maps for fixed buffers; launching is stream-ordered and
 * graph-capturable.  Two kernels sit behind a plan: a per-tile kernel (two CTAs per SM) and a persistent kernel (one CTA
 * per SM, TMEM double buffer, all parity classes of a transposed conv in one launch); the plan picks per shape
 * (environment: ICADV_TC_PERSIST = 0 never / 1 default / 2 wherever eligible -- results agree to fp32 round-off). */
typedef struct icadv_conv_plan icadv_conv_plan;
int icadv_conv_plan_create(const icadv_conv_desc* d, icadv_conv_plan** plan);
int icadv_conv_plan_launch(const icadv_conv_plan* plan, icadv_stream_t stream);
int icadv_conv_plan_destroy(icadv_conv_plan* plan);
int icadv_conv_plan_num_launches(const icadv_conv_plan* plan); /* kernels per icadv_conv_plan_launch */
/* developer profiling: per-CTA phase timestamps (clock64; 16 slots per CTA of the plan's first launch) */
int icadv_conv_plan_set_debug(icadv_conv_plan* plan, long long* dbg);
int icadv_conv_tc(const icadv_conv_desc* d, icadv_stream_t stream); /* create + launch + destroy */
int icadv_conv_tc_supported(const icadv_conv_desc* d);             /* 1 / 0 */

/* CUDA-core path for the shapes the tensor path does not take (3-channel end layers, odd widths).
 * Same descriptor; only ICADV_EPI_LINEAR. */
int icadv_conv_simt(const icadv_conv_desc* d, icadv_stream_t stream);

/* weight gradient of a contraction (train.py:357-359 `out_criterion["loss"].backward()` -- Conv2d / ConvTranspose2d
 * .weight.grad; the attack never reads parameter grads):
 * dW[tap][n][k] = sum_px gout[px,n] * in[tap-shifted px,k]; `d` describes the forward contraction (d->in = its input),
 * gout has the shape of its output; dwpack [taps][n_ch][k_ch] is overwritten; deterministic (fixed-order split reduction).
 * Two kernels: tcgen05 kind::tf32 with the pixel axis as the reduction axis (both operands MN-major straight from the
 * channels-last tensors; k_ch % 32 == 0 and n_ch % 32 == 0; operands are read as TF32, round them where produced) and an
 * fp32 CUDA-core kernel for every other shape and for the parity mode. */
#define ICADV_WGRAD_AUTO 0
#define ICADV_WGRAD_SIMT 1
#define ICADV_WGRAD_TC 2
int icadv_conv_wgrad(const icadv_conv_desc* d, const float* gout, float* dwpack, float* dbias,
                     icadv_stream_t stream); /* = _ex with ICADV_WGRAD_AUTO */
int icadv_conv_wgrad_ex(const icadv_conv_desc* d, const float* gout, float* dwpack, float* dbias, int path,
                        icadv_stream_t stream);
int icadv_conv_wgrad_tc_supported(const icadv_conv_desc* d); /* 1 / 0 */

/* torch layouts -> packed [taps][n_ch][k_ch].  kind: 0 Conv2d.weight [Co,Ci,k,k] for its forward,
 * 1 Conv2d.weight for its input-gradient, 2 ConvTranspose2d.weight [Ci,Co,k,k] for its forward,
 * 3 ConvTranspose2d.weight for its input-gradient. */
int icadv_pack_weight(const float* w, float* wpack, int kind, int c_out, int c_in, int ksize,
                      int round_tf32, icadv_stream_t stream);
/* first-layer form: w[n][3][5][5] (Conv2d.weight [Co,3,5,5] for its forward, or ConvTranspose2d.weight [Ci,3,5,5]
 * for its input-gradient -- same indexing) -> [5 kh][n_ch][32] with k = kw*4 + c, zero elsewhere */
int icadv_pack_weight_rgb(const float* w, float* wpack, int n_ch, int round_tf32, icadv_stream_t stream);
/* dense channels-last RGB [n_img][h][w][3] -> padded RGB0 layout (borders must have been zeroed once) */
int icadv_pad_rgb4(const float* src, float* dst, int n_img, int h, int w, int round_tf32, const int* active,
                   const int* n_active, icadv_stream_t stream);
/* inverse of the above for gradients: packed dW -> torch layout (accumulate = add into dst) */
int icadv_unpack_weight(const float* dwpack, float* dw, int kind, int c_out, int c_in, int ksize,
                        int accumulate, icadv_stream_t stream);

/* layout conversion at the operator surface: NCHW <-> NHWC fp32 */
int icadv_nchw_to_nhwc(const float* src, float* dst, int n, int c, int h, int w, icadv_stream_t stream);
int icadv_nhwc_to_nchw(const float* src, float* dst, int n, int c, int h, int w, icadv_stream_t stream);
/* The [0, 1] clamp of the reconstruction fused into the image layout copies around the MS-SSIM terms (attack_rd.py:353-362;
 * bounds and their gradient rules: utils/ops.py:28-56).  Forward: y = Up_bound(Low_bound(x, 0), 1) written as planes.
 * Backward: the gradient arriving as planes goes back through both bounds, gated by the unclamped interleaved x, and is
 * written interleaved.  Images only (c <= 4). */
int icadv_clamp01_nhwc_to_nchw(const float* x_nhwc, float* y_nchw, int n, int c, int h, int w, icadv_stream_t stream);
int icadv_clamp01_backward_nchw_to_nhwc(const float* g_nchw, const float* x_nhwc, float* gx_nhwc, int n, int c, int h,
                                        int w, icadv_stream_t stream);

/* nn.PixelShuffle(r) on channels-last tensors (compressai subpel_conv3x3, cheng2020_anchor g_s / h_s):
 * dst[n, h*r+i, w*r+j, c] = src[n, h, w, c*r*r + i*r + j]; inverse != 0 applies the inverse permutation (its gradient).
 * (h, w) is the LOW-resolution size, c_out the channel count after shuffling. */
int icadv_pixel_shuffle(const float* src, float* dst, int n, int h, int w, int c_out, int r, int inverse,
                        icadv_stream_t stream);
/* torch.cat / chunk along channels (anchors/model.py:104-105): dst[px, dst_off + c] = src[px, src_off + c], c < count */
int icadv_copy_channels(const float* src, float* dst, int64_t n_px, int c_src, int c_dst, int src_off, int dst_off,
                        int count, icadv_stream_t stream);

/* GDN parameter reparametrisation (compressai NonNegativeParametrizer; utils/ops.py:83-90):
 * eff = max(raw, bound)^2 - pedestal.  transpose != 0 writes eff^T (rows x rows). */
int icadv_gdn_reparam(const float* raw, float* eff, int rows, int cols, float bound, float pedestal,
                      int transpose, int round_tf32, icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Perturbation step (attack_rd.py:507,517,546-554; utils/ops.py:28-56; torch.optim.Adam +
 * MultiStepLR).  Elementwise and layout-agnostic; per_img = elements per image (C*H*W, % 4 == 0).
 * Per-image semantics: every image behaves like one N=1 call of the reference's attack_().
 * ------------------------------------------------------------------------------------------ */
#define ICADV_RED_BLOCKS 128 /* partial sums per image; reduction workspaces hold n_img*128 floats */

/* Device-resident per-image state (all arrays of n_img unless noted): no host sync in the loop. */
typedef struct {
  float* sum_d2;    /* sum (im_s - im_in)^2 */
  float* loss_i;    /* sum_d2 / per_img               (attack_rd.py:333) */
  int* branch;      /* 0 = A (over budget), 1 = B (network pass)   (attack_rd.py:334) */
  int* active;      /* compacted indices of the branch-B images */
  int* n_active;    /* [1] */
  int* step;        /* iteration counter i (0-based before the call, incremented by perturb_forward) */
  float* lr;        /* lr used at this iteration (MultiStepLR([1,2,3]) stepped every steps//3) */
  float* step_size; /* lr / (1 - beta1^t) */
  float* bc2_sqrt;  /* sqrt(1 - beta2^t) */
  unsigned int* counter;          /* [1], zero before the first call: arrival counter of perturb_forward's blocks (the last
                                   * one finalises the state) */
  unsigned long long cond_handle; /* 0, or the handle of a CUDA-graph IF node (icadv_graph_if_create) whose condition
                                   * perturb_forward sets to n_active > 0 */
} icadv_perturb_state;

/* noise -> clamp(+-eps) -> im_in = clamp(im_s + noise_clipped, 0, 1); loss_i; branch (force_branch
 * -1 = by budget, 0/1 = forced); compaction; LR schedule and Adam coefficients for this iteration. */
int icadv_perturb_forward(const float* im_s, const float* noise, float* im_in, float* ws,
                          const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                          float noise_budget, int force_branch, double lr0, double lr_gamma,
                          int sched_period, double beta1, double beta2, icadv_stream_t stream);

/* Backward of the two clamp pairs (custom Low/Up_bound rule) + Adam on the perturbation, fused.
 * Branch A forms d loss_i / d im_in in-kernel (x gradA_scale = 1/per_img); branch B reads g_in =
 * dLoss/d im_in from the network backward (x gradB_scale).  g_a_ext (nullable): externally computed branch-A
 * gradient (att_metric ms-ssim: d(1 - ms_ssim(im_s, im_in))/d im_in, attack_rd.py:336), x gradA_scale. */
int icadv_perturb_update_adam(const float* im_s, float* noise, const float* g_in, const float* g_a_ext, float* m,
                              float* v,
                              const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                              double beta1, double beta2, double adam_eps, float gradA_scale,
                              float gradB_scale, icadv_stream_t stream);

/* ROI / targeted variant of the three calls above (flags coder.py:198-203; formulas attack_cv.py:149-163,
 * attack_data.py:204-221 -- the reference's own call path for them is dead code, so the semantics are the oracle's
 * restatement, oracle/attack.py attack_our_roi):
 *   loss_i = mean(w_in d_in^2),  w_in = mask_tar + lamb_bkg_in * mask_bkg;  branch A iff loss_i >= budget (ge_test);
 *   loss_o = mean(w_out (ref - o)^2), ref = output_t inside the ROI / output_s outside (mixed by the caller),
 *            w_out = lamb_tar * mask_tar + lamb_bkg_out * mask_bkg; minimised: pass grad_scale = -1/per_img.
 * w_in / w_out: [per_img] floats in the image layout, shared by all images of the batch; NULL = all ones.
 * ge_test is a flag word: bit 0 = switch on ">=" (ROI) instead of ">" (attack_rd.py:334); bit 1 = test the budget on the
 * BATCH mean of loss_i, one shared branch for all images -- torch.mean over the batch at attack_rd.py:333, the semantics
 * train.py:342 gets on a training batch (the per-image test, bit 1 clear, is N independent CLI runs). */
int icadv_perturb_forward_roi(const float* im_s, const float* noise, float* im_in, float* ws,
                              const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                              float noise_budget, int force_branch, double lr0, double lr_gamma, int sched_period,
                              double beta1, double beta2, const float* w_in, int ge_test, icadv_stream_t stream);
int icadv_perturb_update_adam_roi(const float* im_s, float* noise, const float* g_in, const float* g_a_ext, float* m,
                                  float* v, const icadv_perturb_state* st, int n_img, int64_t per_img, float eps,
                                  double beta1, double beta2, double adam_eps, float gradA_scale, float gradB_scale,
                                  const float* w_in, icadv_stream_t stream);
int icadv_output_loss_roi(const float* x, const float* ref, float* g_x, float* ws, float* sum_d2, int n_img,
                          int64_t per_img, int do_clamp, float grad_scale, const int* active, const int* n_active,
                          const float* w_out, icadv_stream_t stream);

/* I-FGSM / PGD step (attack_ifgsm.py:409-418): x += alpha*sign(g); project to [x0-eps, x0+eps]. */
int icadv_ifgsm_update(const float* im_s, float* im_adv, const float* g, int64_t n, float alpha,
                       float eps, icadv_stream_t stream);

/* MI-FGSM step (attack_ifgsm.py:348-362), per image: g_mom = mu g_mom + g/||g||_1; x = clamp(x + alpha sign(g_mom), 0, 1);
 * then the eps projection.  ws: n_img*128 floats; l1: n_img floats (written). */
int icadv_mifgsm_update(const float* im_s, float* im_adv, const float* g, float* g_mom, float* ws, float* l1,
                        int n_img, int64_t per_img, float alpha, float eps, float mu, icadv_stream_t stream);

/* C&W step (attack_cw.py:111-140: loss = loss_i + c (1 - MSE_o); c = 0 once MSE_o > 1.1 x the image's target level):
 * g_out = 2 (im_in - im_s) / per_img + c_n * g_net, with c_n = (sum_d2[n] / per_img > 1.1 level[n]) ? 0 : c[n].
 * g_net = gradient of (1 - MSE_o) wrt im_in from the stack backward; c, level, sum_d2: [n_img] device floats. */
int icadv_cw_combine(const float* g_net, const float* im_in, const float* im_s, float* g_out, const float* c,
                     const float* level, const float* sum_d2, int n_img, int64_t per_img, icadv_stream_t stream);

/* Output clamp + distortion (attack_rd.py:353-364) and its gradient seed:
 * o = clamp(x,0,1) (if do_clamp); sum_d2[n] = sum (ref - o)^2;
 * g_x = grad_scale * 2 (ref - o) passed through the Low/Up_bound backward rule (g_x may be NULL). */
int icadv_output_loss(const float* x, const float* ref, float* g_x, float* ws, float* sum_d2, int n_img,
                      int64_t per_img, int do_clamp, float grad_scale, const int* active,
                      const int* n_active, icadv_stream_t stream);

/* Low_bound / Up_bound as stand-alone ops (utils/ops.py:28-56) */
int icadv_bound_forward(const float* x, float* y, int64_t n, float bound, int upper, icadv_stream_t stream);
int icadv_bound_backward(const float* x, const float* gy, float* gx, int64_t n, float bound, int upper,
                         icadv_stream_t stream);

/* per-image sum of squared differences, deterministic two-stage reduction */
int icadv_sum_sqdiff(const float* a, const float* b, float* ws, float* out, int n_img, int64_t per_img,
                     icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Entropy models (compressai EntropyBottleneck / GaussianConditional; call sites
 * anchors/model.py:87-106, attack_rd.py:419, self_ensemble.py:222, train.py:60-64).
 * Channels-last tensors; bits[n] = -sum log2(max(lik, bits_floor)) per image (fixed-order sum).
 * mode 0 = eval (round(x - mean) + mean), mode 1 = train (x + noise[], noise ~ U(-.5,.5) supplied
 * by the caller so two implementations can share the sample).
 * ------------------------------------------------------------------------------------------ */
/* softplus(matrix_i), bias_i, tanh(factor_i) of the 1-3-3-3-3-1 logits chain -> table [58][C] */
int icadv_eb_prepare(const float* const* matrices /*[5]*/, const float* const* biases /*[5]*/,
                     const float* const* factors /*[4]*/, float* table, int C, icadv_stream_t stream);
int icadv_eb_forward(const float* x, const float* noise, const float* table, const float* medians,
                     float* x_hat, float* lik, float* ws, float* bits, int n_img, int64_t per_img, int C,
                     int mode, float lik_bound, float bits_floor, icadv_stream_t stream);
/* lik = Phi((.5-|v|)/s) - Phi((-.5-|v|)/s), s = max(scale, scale_bound), v = y_hat - mean */
int icadv_gc_forward(const float* y, const float* scales, const float* means, const float* noise,
                     float* y_hat, float* lik, float* ws, float* bits, int n_img, int64_t per_img, int mode,
                     float scale_bound, float lik_bound, float bits_floor, icadv_stream_t stream);
/* elementwise helpers: op 0 abs (h_a(|y|), anchors/balle.py:38), 1 relu, 2 leaky 0.01, 3 round, 4 y = x + b,
 * 5 round-to-nearest TF32 (operand preparation for the tensor path), 6 clamp to [0,1] (attack_rd.py:411,
 * self_ensemble.py:182,207), 7 copy; op + 256: the result is rounded to TF32 on store (its readers are tensor-path
 * contractions) */
int icadv_unary(const float* x, const float* b, float* y, int64_t n, int op, icadv_stream_t stream);
/* gradient of op 0 abs / 1 relu / 2 leaky given the forward input (relu/leaky: or the output); op + 256: rounded to TF32
 * on store */
int icadv_act_backward(const float* x, const float* g, float* gx, int64_t n, int op, icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Codec update of adversarial training (train.py:335-366 under --adv; RateDistortionLoss train.py:37-96;
 * optimisers coder.py:50-86).  Only the update step uses these; the attack loop never does.
 * ------------------------------------------------------------------------------------------ */
/* Backward of the EntropyBottleneck likelihood (train mode, x_hat = x + noise): g_x = dL/dx; g_raw [58][C] = dL/d(raw
 * parameters) in the row order of the prepared table (rows 0-2 matrix0, 3-5 bias0, 6-8 factor0, then 15 rows per
 * middle layer: 9 matrix, 3 bias, 3 factor; 54-56 matrix4, 57 bias4) -- softplus / tanh chain rule included.
 * rows = n_img * pixels (channels-last rows of C).  ws: icadv_eb_backward_workspace_floats(rows, C) floats. */
int icadv_eb_backward_workspace_floats(int64_t rows, int C);
int icadv_eb_backward(const float* x_hat, const float* g_lik, const float* table, const float* const* matrices /*[5]*/,
                      const float* const* factors /*[4]*/, float* g_x, float* g_raw, float* ws, int64_t rows, int C,
                      float lik_bound, icadv_stream_t stream);
/* Backward of the GaussianConditional likelihood (train mode): g_y, g_scales, g_means (NULL iff means is NULL);
 * LowerBound backward rule on both the likelihood and the scale bound. */
int icadv_gc_backward(const float* y_hat, const float* scales, const float* means, const float* g_lik, float* g_y,
                      float* g_scales, float* g_means, int64_t n, float scale_bound, float lik_bound,
                      icadv_stream_t stream);
/* out[0] = sum log(max(lik, floor))  (train.py:62-64; ws: 128 floats) and its gradient scale/lik (0 below floor);
 * scale = scale_host * scale_dev[0] (scale_dev nullable) */
int icadv_log_sum(const float* lik, float* ws, float* out, int64_t n, float floor_, icadv_stream_t stream);
int icadv_log_sum_backward(const float* lik, float* g_lik, int64_t n, float floor_, const float* scale_dev,
                           float scale_host, icadv_stream_t stream);
/* out = scale * (a - b): gradient of the MSE distortion term (train.py:71) */
int icadv_scaled_diff(const float* a, const float* b, float* out, int64_t n, const float* scale_dev, float scale_host,
                      icadv_stream_t stream);
/* Gradients of the GDN / IGDN parameters (compressai GDN; utils/ops.py:58-97) w.r.t. the RAW beta [C] and gamma [C][C]
 * (non-negative reparametrisation + LowerBound rule included), from g = dL/dy, the saved output y and scale sc
 * (channels-last rows of C).  ws: icadv_gdn_param_grad_workspace_floats(C) floats. */
int icadv_gdn_param_grad_workspace_floats(int C);
int icadv_gdn_param_grad(const float* g, const float* y, const float* sc, const float* beta_raw, const float* gamma_raw,
                         float* g_beta, float* g_gamma, float* ws, int64_t n_px, int C, int inverse, float beta_bound,
                         float gamma_bound, icadv_stream_t stream);
/* The same gradients on the tensor path: icadv_gdn_param_operands writes T = -+1/2 g y sc^(+-2) and X2 = (y / sc)^2
 * (TF32-rounded, same layout as g); d gamma_eff [C][C] = sum_px T_i X2_j is then the 1x1 case of icadv_conv_wgrad_ex
 * (input X2, output gradient T) and d beta_eff [C] its bias gradient; icadv_gdn_param_grad_finalize applies the chain rule
 * through the reparametrisation (as the last step of icadv_gdn_param_grad). */
int icadv_gdn_param_operands(const float* g, const float* y, const float* sc, float* T, float* X2, int64_t n,
                             int inverse, icadv_stream_t stream);
int icadv_gdn_param_grad_finalize(const float* d_gamma_eff, const float* d_beta_eff, const float* beta_raw,
                                  const float* gamma_raw, float* g_beta, float* g_gamma, int C, float beta_bound,
                                  float gamma_bound, icadv_stream_t stream);
/* out[0] = sum g^2 over a flat gradient buffer (ws: 128 floats): the global norm of clip_grad_norm_ (train.py:360) */
int icadv_sumsq(const float* g, float* ws, float* out, int64_t n, icadv_stream_t stream);
/* clip_grad_norm_(max_norm) + torch.optim.Adam step over ONE flat buffer (after the gradient all-reduce): gradients
 * are g * grad_scale (1/world after a SUM all-reduce); the clip coefficient min(1, max_norm / (grad_scale *
 * sqrt(sumsq[0]) + 1e-6)) is formed on the device; sumsq NULL = no clipping. */
int icadv_adam_clip_step(float* params, const float* grads, float* m, float* v, int64_t n, const float* sumsq,
                         float max_norm, float grad_scale, double lr, double beta1, double beta2, double eps, int step,
                         icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MS-SSIM (pytorch_msssim.ms_ssim, call sites attack_rd.py:336,362, self_ensemble.py:225,228,
 * train.py:44,88; and utils/torch_msssim.py:18-76).  NCHW fp32 planes (planes = N*C).  One level
 * per call: per-plane sums of the ssim and cs maps over the output region; the host composes the
 * 5-level product.  same_pad = 0: valid separable window (variant 1); 1: zero "same" padding
 * (variant 2; 2-D window = outer product of win_taps).  win_taps is a HOST array of `win` floats.
 * ------------------------------------------------------------------------------------------ */
int icadv_ssim_workspace_floats(int planes, int h, int w, int win, int same_pad);
int icadv_ssim_level(const float* X, const float* Y, float* ws, float* ssim_sum, float* cs_sum, int planes,
                     int h, int w, const float* win_taps_host, int win, int same_pad, float c1, float c2,
                     icadv_stream_t stream);
/* Backward of one level w.r.t. X (same_pad as above: 0 = variant 1, 1 = variant 2).  coef_cs / coef_ss [planes]:
 * upstream weights on the per-plane SUMS of the cs / ssim maps; dXnext (nullable): gradient w.r.t. the
 * 2x2-average-pooled next level, folded in.  Statistics are recomputed in shared memory (nothing saved by the forward).
 * The SSIM map is symmetric in (X, Y): the gradient w.r.t. Y is this call with the two images swapped.  Used by the
 * ms-ssim attack metric and the ms-ssim RD loss (attack_rd.py:336,362 and train.py:44,88 under autograd). */
int icadv_ssim_level_backward(const float* X, const float* Y, const float* coef_cs, const float* coef_ss,
                              const float* dXnext, float* dX, int planes, int h, int w, int next_h, int next_w,
                              int pad_h, int pad_w, const float* win_taps_host, int win, int same_pad, float c1,
                              float c2, icadv_stream_t stream);
/* Value AND gradient of one level in one pass (11-tap window; csrc/icadv_msssim.cu, row-marching kernel): the per-plane
 * sums as icadv_ssim_level, plus U = d(sum of the level's map)/dX at unit upstream weight -- the cs map, or the ssim map
 * when last_level != 0.  The upstream weights depend on every level's value (ms_ssim = prod_l relu(v_l)^w_l,
 * pytorch_msssim.ms_ssim at attack_rd.py:336,362), so the host applies them afterwards, coarse to fine, with
 * icadv_ssim_combine: U_l <- coef[plane] * U_l + 0.25 * D_{l+1}[pool parent] (Dnext nullable at the coarsest level;
 * pad_h / pad_w = the padding of the avg_pool2d that produced level l+1). */
int icadv_ssim_vg_workspace_floats(int planes, int h, int w, int same_pad);
int icadv_ssim_level_value_grad(const float* X, const float* Y, float* U, float* ws, float* ssim_sum, float* cs_sum,
                                int planes, int h, int w, const float* win_taps_host, int win, int same_pad, float c1,
                                float c2, int last_level, icadv_stream_t stream);
/* The scalar step between the level launches and the combine launches, in one launch: sums the per-block partials the
 * level launches left in ws_all (call icadv_ssim_level_value_grad with ssim_sum = cs_sum = NULL; level l's partials start
 * at ws_all + ws_offset_host[l], nblocks_host[l] = icadv_ssim_vg_workspace_floats / (2 planes) blocks per plane),
 * value[b] = mean_c prod_l relu(v_l)^w_l  (v_l = mean cs map, mean ssim map at the last level; npx_host[l] = map size) and
 * coef [levels][planes] = d(sum_b upstream[b] value[b]) / d(sum of level l's map) -- what icadv_ssim_combine consumes.
 * pytorch_msssim.ms_ssim (attack_rd.py:336,362; train.py:44,88) under autograd. */
int icadv_msssim_coefficients(const float* ws_all, const int* ws_offset_host, const int* nblocks_host,
                              const float* npx_host, const float* weights_host, int levels, const float* upstream,
                              float* value, float* coef, int batch, int channels, icadv_stream_t stream);
int icadv_ssim_combine(float* U, const float* coef, const float* Dnext, int planes, int h, int w, int next_h, int next_w,
                       int pad_h, int pad_w, icadv_stream_t stream);
/* F.avg_pool2d(x, 2, stride 2, padding (pad_h, pad_w)) between levels */
int icadv_avgpool2(const float* x, float* y, int planes, int h, int w, int pad_h, int pad_w,
                   icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3xTF32 parity mode (csrc/icadv_split.cu).  The contractions of the codec stacks
 * (anchors/utils.py:112-130; compressai GDN, utils/ops.py:58-97) stay on the tcgen05 kernels, but every fp32
 * operand is split x = hi + lo (both TF32) and the product is hi*Whi + lo*Whi + hi*Wlo, accumulated in fp32 in
 * TMEM.  The split is a layout: activations [px][C] -> [hi | lo | hi | 0] along the channels (layout 0), packed
 * weights [rows][C] -> [hi | hi | lo | 0] (layout 1), written as G slices of Ks channels (Ks * G >= 3C); the contraction
 * is then launched once per slice with k_ch = Ks (see icadv_sum_slices).  op 1 squares the input first (operand of the GDN normalisation).  GDN / IGDN run unfused in this
 * mode: split3(op 1) -> 1x1 contraction with gamma, bias beta -> icadv_gdn_apply; backward: icadv_gdn_bwd_operand_split3
 * -> 1x1 contraction with gamma^T -> icadv_gdn_bwd_combine.  Purpose: per-step losses of attack_rd.py:332-379,506-560
 * within 1e-3 of the fp32 reference.
 * ------------------------------------------------------------------------------------------ */
int icadv_split3(const float* x, float* out, int64_t n_px, int C, int Ks, int G, int op, int layout,
                 icadv_stream_t stream);
int icadv_gdn_bwd_operand_split3(const float* g, const float* y, const float* sc, float* out, int64_t n_px, int C, int Ks,
                                 int G, int inverse, icadv_stream_t stream);
/* K-slicing (the tensor core truncates its fp32 accumulator after every instruction, so long accumulation chains come
 * out low by ~2^-24 per step): the split form is cut into G slices of Ks channels -- activations [G][px][Ks], weights
 * [G][rows][Ks] -- each slice is contracted by its own launch (k_ch = Ks) into its own partial output, and the partials
 * [G][n] are added in round-to-nearest fp32: */
int icadv_sum_slices(const float* parts, float* out, int64_t n, int G, icadv_stream_t stream);
/* sc = norm^(-1/2) (inverse: ^(+1/2)); y = x * sc */
int icadv_gdn_apply(const float* x, const float* nrm, float* y, float* sc, int64_t n, int inverse,
                    icadv_stream_t stream);
/* out = g sc - (y/sc) w   (inverse: +) */
int icadv_gdn_bwd_combine(const float* g, const float* y, const float* sc, const float* w, float* out, int64_t n,
                          int inverse, icadv_stream_t stream);

/* U[lo, hi) noise of the training-mode quantiser (compressai quantize(x, "noise"): x + U(-.5,.5), call site
 * anchors/model.py:102; EntropyBottleneck / GaussianConditional in train mode): Philox4x32-10 keyed by `seed`,
 * counter `offset` + element block, so a (seed, offset) pair reproduces the sample on any launch geometry. */
int icadv_uniform_noise(float* out, int64_t n, uint64_t seed, uint64_t offset, float lo, float hi,
                        icadv_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Entropy coding: what compressai's CPU coder does behind
 * EntropyBottleneck / GaussianConditional .compress() / .decompress() and model.compress() / .decompress()
 * (compressai/cpp_exts/rans/rans_interface.cpp encode_with_indexes / decode_with_indexes on ryg_rans rans64.h;
 * compressai/cpp_exts/ops/ops.cpp pmf_to_quantized_cdf; reference call sites attack_TIC.py:106-110,
 * InvCompress/attack_inv.py:112-116, InvCompress/ours.py:100-175, InvCompress/train.py:452 net.update()).
 * Per-symbol arithmetic is compressai's bit for bit (64-bit state, 32-bit renormalisation, 16-bit probabilities, 4-bit
 * bypass digits for out-of-range symbols).  An image's symbol sequence is dealt round-robin to `lanes` independent
 * coders (one GPU thread each); the string of an image is [lane word counts: lanes x u32 | lane 0 words | lane 1 ...],
 * little-endian u32 words; with lanes == 1 the header is omitted and the string is compressai's.
 * symbols / indexes / means / out: channels-last [n_img][hw][C].  mode 0: flat (c, h, w) sequence order (compressai's
 * order outside the autoregressive models), lane l takes flat indexes l, l + lanes, ...; mode 1: position-major over
 * `order` (n_pos hw-indexes, NULL = raster 0 .. n_pos - 1), lane l takes channels l, l + lanes, ... of every position
 * (lanes == 1 + raster = compressai's _compress_ar order; a wavefront list makes the autoregressive decode parallel).
 * cdf: [n_cdf][cdf_stride] int32 rows as built by icadv_pmf_to_quantized_cdf, cdf_sizes = entries per row (pmf
 * length + 2), offsets = symbol value of slot 0.
 * ------------------------------------------------------------------------------------------ */
/* HOST function (parameter-sized): cdf[0 .. n] from pmf[0 .. n-1], total 2^precision, no zero-width slot */
int icadv_pmf_to_quantized_cdf(const float* pmf, int n, int precision, int* cdf);
/* words of scratch per (image, lane) that icadv_rans_encode needs */
int icadv_rans_lane_capacity(int n_img, int hw, int C, int mode, int n_pos, int lanes);
/* scratch: n_img * lanes * capacity words; lane_words: n_img * lanes ints; packed: n_img rows of packed_stride words
 * (>= lanes * capacity + lanes); total_words[n_img] = length of each string in words */
int icadv_rans_encode(const int* symbols, const int* indexes, int n_img, int hw, int C, const int* cdf,
                      const int* cdf_sizes, const int* offsets, int cdf_stride, int mode, const int* order, int n_pos,
                      int lanes, uint32_t* scratch, int* lane_words, uint32_t* packed, int packed_stride,
                      int* total_words, icadv_stream_t stream);
/* out = decoded symbol (+ means when given) */
int icadv_rans_decode(const uint32_t* packed, int packed_stride, const int* indexes, const float* means, float* out,
                      int n_img, int hw, int C, const int* cdf, const int* cdf_sizes, const int* offsets,
                      int cdf_stride, int mode, const int* order, int n_pos, int lanes, icadv_stream_t stream);
/* incremental decoding (mode 1) for the autoregressive models, whose indexes depend on symbols decoded earlier: the
 * per-lane coder state (state, cursor: n_img * lanes each) lives in device memory between steps; each step decodes the
 * channels of `n_positions` listed positions */
int icadv_rans_decode_init(const uint32_t* packed, int packed_stride, int n_img, int lanes, uint64_t* state,
                           int* cursor, icadv_stream_t stream);
int icadv_rans_decode_step(const uint32_t* packed, int packed_stride, uint64_t* state, int* cursor, const int* indexes,
                           const float* means, float* out, int n_img, int hw, int C, const int* cdf,
                           const int* cdf_sizes, const int* offsets, int cdf_stride, const int* positions,
                           int n_positions, int lanes, icadv_stream_t stream);
/* compressai GaussianConditional.build_indexes: out = #(table[0 .. levels-2] < max(scale, bound)) */
int icadv_build_indexes(const float* scales, const float* table, int levels, float bound, int* out, long long n,
                        icadv_stream_t stream);

/* compressai.layers.AttentionBlock gate (cheng2020_attn): y = a * sigmoid(b) + x; backward:
 * ga = g * sigmoid(b), gb = g * a * sigmoid(b) * (1 - sigmoid(b)) (the gradient to x is g itself) */
int icadv_attention_gate(const float* a, const float* b, const float* x, float* y, int64_t n, icadv_stream_t stream);
int icadv_attention_gate_backward(const float* a, const float* b, const float* g, float* ga, float* gb, int64_t n,
                                  icadv_stream_t stream);

/* Conditional section of a captured launch sequence (CUDA graph IF node; replaces the reference's per-step host branch
 * `if loss_i >= noise_thres` of attack_rd.py:333-334 for the launches only one branch needs).  Call while
 * `capture_stream` is capturing: launches made on `body_stream` between begin and end become the body of an IF node of
 * the capturing graph and run at replay time only if *flag > 0 (read on the device when the graph reaches that point);
 * work captured on `capture_stream` after begin depends on the IF node.  The body must not allocate. */
int icadv_graph_if_begin(const int* flag, icadv_stream_t capture_stream, icadv_stream_t body_stream);
int icadv_graph_if_end(icadv_stream_t body_stream);
/* Two-step form: the handle is created first and handed to the launch that computes the condition (perturb_forward, through
 * icadv_perturb_state.cond_handle), which saves the one-thread launch icadv_graph_if_begin makes to set it. */
int icadv_graph_if_create(icadv_stream_t capture_stream, unsigned long long* handle_out);
int icadv_graph_if_begin_handle(unsigned long long handle, icadv_stream_t capture_stream, icadv_stream_t body_stream);

/* Roofline denominator for the contraction kernels (bench.py): one launch of a bare tcgen05.mma kind::tf32 loop
 * (128 x n x 8 instructions on static shared-memory operands, one CTA per SM, `iters` K-blocks of four MMAs each).
 * The caller times the launch with CUDA events; *flops_out receives the FLOP it performs. */
int icadv_probe_tf32_peak(int iters, int n, double* flops_out, icadv_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ICADV_H_ */
