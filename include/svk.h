/*
 * svk.h — C-ABI of libsvk.so, the sm_100a kernel library behind the speaker-embedding hot path.
 *
 * Drop-in boundary: the reference (ZihanLiao/pytorch-kaldi-resnet) has no native code;
 * its hot path is `scripts/model.py` (NeuralSpeakerModel), `scripts/train_resnet.py` (train loop),
 * `scripts/decode.py` (extraction) and `scripts/cosine_score.py` / `compute_topk_mean_std.py` /
 * `adaptive_snorm.py` (scoring), all of which reach the GPU only through torch library calls.  Each entry
 * point below replaces one of those torch calls; the citation after "replaces:" is the reference call site.
 * The Python host side (pytorch-kaldi-resnet_b200/svk) binds these with ctypes.
 *
 * Conventions
 *  - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller (torch), contiguous.
 *  - activations are NHWC ("channels last"); `dtype` selects their storage: SVK_F32 (validation mode) or
 *    SVK_BF16 (product mode).  Accumulation is always fp32; cross-CTA statistics are fp64.
 *  - `stream` is a cudaStream_t passed as void*.
 *  - return 0 on success; negative SVK_E_* on bad arguments / unsupported shapes; positive = cudaError_t.
 *    svk_last_error_string() describes the last failure on the calling thread.  There is no CPU fallback.
 *  - kernels never allocate, free or retain pointers.
 */
#ifndef SVK_H_
#define SVK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVK_VERSION 100

enum { SVK_F32 = 0, SVK_BF16 = 1 };
enum { SVK_IMPL_SIMT = 0, SVK_IMPL_TCGEN05 = 1 };
enum {
  SVK_E_BADARG = -1,      /* null pointer / non-positive size */
  SVK_E_UNSUPPORTED = -2, /* shape, dtype or impl not supported by this kernel */
  SVK_E_ALIGN = -3,       /* pointer or channel count not aligned as required */
  SVK_E_DRIVER = -4       /* cuTensorMapEncodeTiled / driver entry point failure */
};

int svk_version(void);
const char* svk_last_error_string(void);
/* Number of kernels launched by this library in this process (for bench.py's gpu_launches). */
long long svk_launch_count(void);

/* ---------------------------------------------------------------- convolution ---------------------------- */
/* One conv layer of the ResNet trunk.  replaces: nn.Conv2d forward/backward, model.py:12-15 (conv3x3),
 * model.py:233-236 (1x1 stride-2 downsample).  Weights are pre-packed by svk_pack_conv_weight. */
typedef struct svk_conv_desc {
  int N, H, W, Cin;   /* input  NHWC */
  int Ho, Wo, Cout;   /* output NHWC; Ho = (H-1)/stride+1 (pad = R/2) */
  int R;              /* 1 or 3 (square filter, pad R/2) */
  int stride;         /* 1 or 2 */
  int dtype;          /* SVK_F32 | SVK_BF16 (activation + packed-weight storage) */
  int impl;           /* SVK_IMPL_SIMT (any dtype) | SVK_IMPL_TCGEN05 (bf16 only) */
} svk_conv_desc;

/* OIHW fp32 master weights -> packed [R*R][Cout][Cin] (w_fwd, K-major for fprop) and
 * [R*R][Cin][Cout] (w_dgrad), both in `dtype`.  Either output may be NULL. */
int svk_pack_conv_weight(const float* w_oihw, void* w_fwd, void* w_dgrad, int Cout, int Cin, int R,
                         int dtype, void* stream);

/* y = conv(x, w).  Epilogue (all optional, applied in this order on the fp32 accumulator):
 *   v = acc * scale[c] + shift[c]          (scale/shift both non-NULL: folded eval-mode BatchNorm)
 *   v += residual[idx]                      (residual non-NULL, same NHWC shape as y)
 *   v = max(v, 0)                           (relu != 0)
 *   y[n, :, w >= valid_wo[n], :] = 0        (valid_wo non-NULL: per-utterance valid output width — batched
 *                                            extraction keeps exact batch-1 semantics, decode.py:198)
 *   stats[c] += sum(y), stats[Cout+c] += sum(y^2) over the stored (rounded) values   (stats non-NULL)
 * replaces: model.py:51,55,59 (+ bn/relu/add at :52-62 in eval mode). */
int svk_conv2d_fwd(const svk_conv_desc* d, const void* x, const void* w_fwd, void* y, double* stats,
                   const float* scale, const float* shift, const void* residual, int relu,
                   const int* valid_wo, void* stream);
/* dx = conv_transpose(dy, w) [+ res] [+ res_m * (mask > 0)]  — data gradient; desc describes the FORWARD conv
 * (dy is N,Ho,Wo,Cout; dx is N,H,W,Cin).  replaces: cuDNN dgrad under loss.backward(), train_resnet.py:327. */
int svk_conv2d_dgrad(const svk_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, const void* res,
                     const void* res_m, const void* mask, void* stream);
/* Data gradient with the first half of the BatchNorm backward of the layer BELOW fused into its epilogue.  dx is the
 * gradient w.r.t. relu(bn(c) [+ shortcut]); the kernel zeroes it where bn->mask <= 0 (the ReLU mask; mask is the stored
 * activation, shaped like dx) and, when bn->c is non-NULL, accumulates the two sums BatchNorm's backward needs over the
 * values it stores:  sums[ch] += sum(dx),  sums[Cin + ch] += sum(dx * (c - mean[ch]) * rstd[ch])  — exactly what
 * svk_bn_bwd_reduce(dx, mask, c, mean, rstd, ...) would compute in a separate pass over dx, mask and c.
 * svk_bn_bwd_apply is then called with out = NULL (dx is already masked).  bn == NULL: plain svk_conv2d_dgrad.
 * Not available for the 1x1/stride-2 accumulate form, nor together with res_m (the epilogue streams at most three
 * tensors: pass an already-masked gradient as res).  replaces: the ReLU and batch_norm backward nodes autograd runs
 * after each conv's dgrad under loss.backward(), train_resnet.py:327 (model.py:52-53,56-62). */
typedef struct svk_bn_bwd_fuse {
  const void* mask;     /* NHWC, same shape/dtype as dx */
  const void* c;        /* raw conv output normalised by the BatchNorm (same shape), or NULL: mask only */
  const float* mean;    /* [Cin] batch mean (c non-NULL) */
  const float* rstd;    /* [Cin] 1/sqrt(var + eps) */
  double* sums;         /* [2][Cin], accumulated (zero it first) */
} svk_bn_bwd_fuse;
int svk_conv2d_dgrad_bn(const svk_conv_desc* d, const void* dy, const void* w_dgrad, void* dx, const void* res,
                        const void* res_m, const void* mask, const svk_bn_bwd_fuse* bn, void* stream);
/* Block-input gradient of a downsample block in one call: dx = dgrad(conv1, 3x3/s2) + dgrad(shortcut conv, 1x1/s2), with
 * the same optional BatchNorm-backward fusion for the layer below (every dx pixel gets its final value in exactly one
 * epilogue: the 1x1 launch writes the even/even pixels, the 3x3 launch of that parity class adds them back in).
 * replaces: the two ConvolutionBackward nodes + the add autograd runs for model.py:59-62 under loss.backward(). */
int svk_downsample_dgrad_bn(const svk_conv_desc* d_conv1, const void* dy1, const void* w1_dgrad,
                            const svk_conv_desc* d_convd, const void* dyd, const void* wd_dgrad, void* dx,
                            const svk_bn_bwd_fuse* bn, void* stream);
/* g = mask > 0 ? g : 0 in place (n elements; n a multiple of the 16-byte vector width). */
int svk_relu_mask_inplace(void* g, const void* mask, long long n, int dtype, void* stream);
/* All convs of a network in ONE launch: table[nconv][6] (int64, device) = {offset of the OIHW weight in `flat_params`
 * (floats), offset of its packed copies in w_fwd/w_dgrad (elements), Cout, Cin, R*R, cumulative element start};
 * total = sum Cout*Cin*R*R.  Same layouts as svk_pack_conv_weight. */
int svk_pack_conv_weights_batched(const float* flat_params, void* w_fwd, void* w_dgrad, const long long* table, int nconv,
                                  long long total, int dtype, void* stream);

/* dw_oihw[Cout][Cin][R][R] (fp32, overwritten) = dy^T * im2col(x).  Split-K over pixel tiles: every CTA writes its
 * partial sum into `workspace` with plain stores, then one reduction kernel sums the partials in a fixed order and
 * transposes to OIHW (deterministic, no atomics).  workspace_bytes >= svk_conv2d_wgrad_workspace_bytes(d).
 * replaces: cuDNN wgrad under loss.backward(), train_resnet.py:327. */
size_t svk_conv2d_wgrad_workspace_bytes(const svk_conv_desc* d);
int svk_conv2d_wgrad(const svk_conv_desc* d, const void* x, const void* dy, float* dw_oihw, void* workspace,
                     size_t workspace_bytes, void* stream);

/* Stem: 3x3 s1 p1 conv, Cin = 1, x is the (B,F,T) fp32 feature tensor itself.  replaces: model.py:247-249.
 * stats (nullable): [2*Cout] doubles, += per-channel sum / sum of squares of the stored values (caller zeroes). */
int svk_stem_conv_fwd(const float* x, const float* w /*[Cout][9]*/, void* y /*N,H,W,Cout*/, int N, int H, int W,
                      int Cout, int dtype, const float* scale, const float* shift, int relu,
                      const int* valid_w /*nullable, per-utterance width: y[n,:,w>=valid_w[n],:] = 0*/, double* stats,
                      void* stream);
int svk_stem_conv_wgrad(const float* x, const void* dy, float* dw /*[Cout][9], overwritten*/, int N, int H, int W,
                        int Cout, int dtype, void* stream);

/* ---------------------------------------------------------------- BatchNorm2d ---------------------------- */
/* stats[c] += sum_m x[m,c]; stats[C+c] += sum_m x[m,c]^2  (x is [M,C]).  Caller zeroes stats. */
int svk_channel_stats(const void* x, long long M, int C, int dtype, double* stats, void* stream);
/* Training-mode BN finalise. replaces: native_batch_norm statistics + running-stat update, model.py:52,56,250.
 * mean = s1/M, var_b = s2/M - mean^2; scale = gamma*rstd, shift = beta - mean*scale; running stats updated with
 * momentum and UNBIASED variance; save_mean/save_rstd kept for backward. */
int svk_bn_finalize(const double* stats, long long M, int C, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* scale, float* shift,
                    float* save_mean, float* save_rstd, void* stream);
/* Training-mode BN finalise + apply in one pass: out = act(bn(x) + R), R = 0 | res | bn_b(res) (downsample branch).
 * Every thread derives the coefficients of its channels from the fp64 sums; the first threads also publish
 * coef[0..3][cstride] = scale, shift, mean, rstd (kept for backward) and update the running statistics exactly like
 * svk_bn_finalize.  replaces: model.py:52-53, :56-62 (+ :233-236 downsample BN) in training mode. */
int svk_bn_train_act_fwd(const void* x, const double* stats, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, float* coef, const void* res, const double* stats_b, const float* gamma_b,
                         const float* beta_b, float* running_mean_b, float* running_var_b, float* coef_b, int cstride,
                         float momentum, float eps, int relu, void* out, long long M, int C, int dtype, void* stream);
/* Eval-mode BN coefficients from running stats: scale = gamma/sqrt(rv+eps), shift = beta - rm*scale. */
int svk_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                       float eps, int C, float* scale, float* shift, void* stream);
/* out = act(scale*x + shift + R) with R = 0 | res | rscale*res + rshift.  replaces: model.py:52-53,56-62. */
int svk_bn_act_fwd(const void* x, const float* scale, const float* shift, const void* res, const float* rscale,
                   const float* rshift, int relu, void* out, long long M, int C, int dtype, void* stream);
/* Backward reductions for one BN (optionally two sharing the same upstream gradient — main path + downsample):
 * g = dout * (out > 0 if out else 1);  sums[0:C] += sum g;  sums[C:2C] += sum g*xhat(c,mean,rstd);
 * sums[2C:3C] += sum g*xhat(c_b,mean_b,rstd_b) if c_b.  replaces: native_batch_norm_backward reductions. */
int svk_bn_bwd_reduce(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                      const void* c_b, const float* mean_b, const float* rstd_b, double* sums, long long M, int C,
                      int dtype, void* stream);
/* dc = gamma*rstd*(g - s1/M - xhat*s2/M) (and dc_b likewise); dgamma = s2, dbeta = s1 written to fp32 grads. */
int svk_bn_bwd_apply(const void* dout, const void* out, const void* c, const float* mean, const float* rstd,
                     const float* gamma, void* dc, const void* c_b, const float* mean_b, const float* rstd_b,
                     const float* gamma_b, void* dc_b, const double* sums, float* dgamma, float* dbeta,
                     float* dgamma_b, float* dbeta_b, long long M, int C, int dtype, void* stream);
/* out = a + b  |  out = a + b*(mask>0)  (elementwise, residual-gradient merge).  mask may be NULL. */
int svk_add_masked(const void* a, const void* b, const void* mask, void* out, long long n, int dtype, void* stream);
/* dx[n, 2i, 2j, :] += d[n, i, j, :]  (merge the 1x1/s2 downsample data gradient into the block-input gradient). */
int svk_add_strided2(void* dx, const void* d, int N, int H, int W, int Ho, int Wo, int C, int dtype, void* stream);

/* ---------------------------------------------------------------- pooling + FC --------------------------- */
/* StatsPooling. mode 0 = 'mean' -> out[n, c*H + h]; mode 1 = 'mean+std' reproducing the reference's swapped
 * var_mean unpack: out[n, c*2H + h] = unbiased var over time, out[n, c*2H + H + h] = sqrt(mean over time).
 * valid_w (nullable) = per-utterance width.  replaces: model.py:441-455 (+ Flatten :381). */
int svk_statspool_fwd(const void* x, float* out, int N, int H, int W, int C, int mode, const int* valid_w,
                      int dtype, void* stream);
/* relu_mask != 0: x is a ReLU output and dx is additionally multiplied by (x > 0) (model.py:62 backward). */
int svk_statspool_bwd(const void* x, const float* dout, void* dx, int N, int H, int W, int C, int mode, int relu_mask,
                      int dtype, void* stream);
/* C[M,N] = alpha * op(A)[M,K] * op(B)[K,N] (+ bias[N]) (+ beta*C); fp32; element (m,k) of op(A) at
 * A[m*a_sm + k*a_sk], element (k,n) of op(B) at B[k*b_sk + n*b_sn].
 * replaces: nn.Linear fc1 (model.py:384) and its backward, F.linear in AAMLayer (model.py:485). */
int svk_sgemm(const float* A, long long a_sm, long long a_sk, const float* B, long long b_sk, long long b_sn,
              float* C, long long ldc, int M, int N, int K, float alpha, float beta, const float* bias,
              void* stream);
/* Tensor-core GEMM (tcgen05 kind::tf32, fp32 accumulate): C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).  a_kmajor / b_kmajor:
 * 1 = the operand is stored with K contiguous (A[m*lda + k], B[n*ldb + k]); 0 = stored transposed (A[k*lda + m],
 * B[k*ldb + n]) — the views the backward GEMMs need, consumed without a transposing copy.  lda/ldb must be multiples of
 * 4 floats.  Small-tile-count problems are split along K into `workspace` (>= svk_gemm_tf32_workspace_bytes) and summed
 * in a fixed order.  Product-mode replacement of svk_sgemm for fc1 (model.py:384), the AAM cosine GEMM (model.py:485),
 * their gradients, and the s-norm cohort score matrix (compute_topk_mean_std.py:17). */
size_t svk_gemm_tf32_workspace_bytes(int M, int N, int K);
int svk_gemm_tf32(const float* A, long long lda, int a_kmajor, const float* B, long long ldb, int b_kmajor, float* C,
                  long long ldc, int M, int N, int K, const float* bias, void* workspace, size_t workspace_bytes,
                  void* stream);
/* out[n] = sum_m x[m*ld + n]  (bias gradient). */
int svk_colsum(const float* x, float* out, int M, int N, long long ld, void* stream);

/* ---------------------------------------------------------------- AAM-softmax head ----------------------- */
/* xhat = x / max(||x||, eps) row-wise; inv[r] = 1/max(||x||,eps).  replaces: F.normalize, model.py:485. */
int svk_l2norm_rows_fwd(const float* x, float* xhat, float* inv, int rows, int cols, float eps, void* stream);
/* dx = (dxhat - xhat * <xhat, dxhat>) * inv.  (exact for ||x|| > eps) */
int svk_l2norm_rows_bwd(const float* dxhat, const float* xhat, const float* inv, float* dx, int rows, int cols,
                        void* stream);
/* In place on cos[B,C]: target column -> phi (cos(theta+m) with the monotonic fallback), everything * s.
 * cos_t[b] keeps the raw target cosine for backward.  replaces: model.py:487-499. */
int svk_aam_margin_fwd(float* cos_logits, const long long* label, float* cos_t, int B, int C, float cos_m,
                       float sin_m, float th, float mm, float s, void* stream);
/* In place on dlogits (B rows of C values, row pitch ld >= C) -> dcos (chain rule through the margin and the scale). */
int svk_aam_margin_bwd(float* dlogits, const long long* label, const float* cos_t, int B, int C, int ld, float cos_m,
                       float sin_m, float th, float s, void* stream);
/* Cross entropy (mean over batch). loss_rows[b] = lse_b - logits[b,y_b]; lse[b] saved; rank[b] = number of
 * logits strictly greater than the target's (top-k correct iff rank < k).
 * replaces: nn.CrossEntropyLoss (train_resnet.py:201,317), accuracy.py:4-16. */
int svk_ce_fwd(const float* logits, const long long* label, float* loss_rows, float* lse, int* rank,
               float* loss_mean /*nullable; += loss_rows[b]/B, caller zeroes*/, int B, int C, void* stream);
/* dlogits = (softmax - onehot) * (*gout) * mult   (gout = upstream gradient, a DEVICE scalar; mult = 1/B). */
int svk_ce_bwd(const float* logits, const long long* label, const float* lse, const float* gout, float mult,
               float* dlogits, int B, int C, void* stream);

/* Fused AAM-softmax head + cross-entropy (csrc/aam_fused.cu), embedding width E = 256 only.
 * replaces: AAMLayer.forward (model.py:483-501) + nn.CrossEntropyLoss (train_resnet.py:201,317) + accuracy.py:4-16 and
 * their autograd backward, in four launches; d_logits, x_hat and W_hat never exist in memory.
 *   fwd: logits[B,C] (margin on the target column, x s), cos_t[B] (raw target cosine), lse[B], loss_rows[B] (nullable),
 *        rank[B] (nullable; classes scoring strictly above the target), loss_mean (nullable; zeroed, then += loss_rows/B).
 *   bwd: dh[B,E], dW[C,E] from (logits, lse, cos_t) and the upstream scalar *gout (nullable = 1): d loss_mean.
 *   exact != 0: 3xTF32 split products (fp32 validation mode, <= 1e-5 of the fp32 oracle); 0: single tf32 products.
 *   workspace: svk_aam_ce_workspace_bytes(B, E, C) bytes, 16-byte aligned; the same buffer serves fwd and bwd. */
size_t svk_aam_ce_workspace_bytes(int B, int E, int C);
int svk_aam_ce_fwd(const float* h, const float* W, const long long* label, float* logits, float* cos_t, float* lse,
                   float* loss_rows, int* rank, float* loss_mean, int B, int E, int C, float cos_m, float sin_m, float th,
                   float mm, float s, int exact, void* workspace, size_t workspace_bytes, void* stream);
int svk_aam_ce_bwd(const float* h, const float* W, const long long* label, const float* logits, const float* lse,
                   const float* cos_t, const float* gout, float* dh, float* dW, int B, int E, int C, float cos_m, float sin_m,
                   float th, float s, int exact, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- optimiser ------------------------------ */
/* torch.optim.SGD semantics on flat buffers: d = g*gscale + wd*p; buf = mom*buf + d; p -= lr*buf.
 * (buf zero-initialised makes the first step equal torch's buf = d.)  replaces: train_resnet.py:203,328. */
int svk_sgd_step(float* p, const float* g, float* buf, long long n, float lr, float momentum, float wd,
                 float gscale, void* stream);
int svk_cast(const void* src, void* dst, long long n, int src_dtype, int dst_dtype, void* stream);

/* ---------------------------------------------------------------- scoring -------------------------------- */
/* score[t] = cos(E[ie[t]] - mean, T[it[t]] - mean), eps 1e-8 on the norm product as F.cosine_similarity.
 * replaces: cosine_score.py:60-65. */
int svk_cosine_score_pairs(const float* E, const float* T, const float* mean, const int* ie, const int* it,
                           float* score, long long ntrials, int D, void* stream);
/* Per row of scores[rows, ncoh]: top-k values -> mean and UNBIASED std.  replaces: compute_topk_mean_std.py:17-19. */
int svk_topk_meanstd(const float* scores, int rows, int ncoh, int topk, float* mean, float* stdv, void* stream);
/* out[t] = 0.5*((s-me[ie])/max(se[ie],1e-8) + (s-mt[it])/max(st[it],1e-8)).  replaces: adaptive_snorm.py:33-34. */
int svk_snorm_apply(const float* score, const int* ie, const int* it, const float* mean_e, const float* std_e,
                    const float* mean_t, const float* std_t, float* out, long long ntrials, void* stream);

/* ---------------------------------------------------------------- scoring back end ------- */
/* Stable ascending sort of (float64 key, int32 value) pairs (LSD radix, 8 x 8 bits; ties keep their input order).
 * replaces: the stable `sorted(..., key=itemgetter(1))` over Python floats of compute_eer.py:38-40 and
 * local/compute_min_dcf.py:58-61. */
size_t svk_sort_pairs_f64_workspace_bytes(long long n);
int svk_sort_pairs_f64(const double* keys, const int* vals, double* keys_out, int* vals_out, long long n,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Labels (1 = target) in ascending score order -> out6 = { eer = max(fpr, fnr) at the first index minimising
 * |fnr - fpr|, that index, min over i of c_miss*fnr*p_target + c_fa*fpr*(1-p_target) (not normalised), its first index,
 * n_target, n_nontarget }, float64 with the reference's operation order.
 * replaces: compute_eer.py:35-70,99-100; local/compute_min_dcf.py:54-106. */
size_t svk_det_metrics_workspace_bytes(long long n);
int svk_det_metrics(const int* sorted_labels, long long n, long long n_target, double p_target, double c_miss, double c_fa,
                    double* out6, void* workspace, size_t workspace_bytes, void* stream);
/* out[s] = mean of the rows X[order[offsets[s] .. offsets[s+1])], accumulated in that order (float32).
 * replaces: compute_speaker_mean.py:16-27. */
int svk_segment_mean(const float* X, const int* order, const int* offsets, int n_seg, int D, float* out, void* stream);
/* out[d] = mean over the n rows of X[n, D] (float64 accumulation).  replaces: compute_mean.py:9-20. */
size_t svk_col_mean_workspace_bytes(long long n, int D);
int svk_col_mean(const float* X, long long n, int D, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- diagnostics -------------------------- */
/* With SVK_PROF=1 in the environment the tensor-core convolution kernels add per-role cycle counters to a 16-entry device
 * table ([0] CTAs; MMA warp: [1] loop cycles, [2] waiting for operands, [3] waiting for a free accumulator; producer: [4]
 * loop, [5] waiting for a free stage; first epilogue warp: [6] loop, [7] waiting for an accumulator; [8..12] globaltimer
 * sums / extrema of the MMA loops; [14] the peer producer of a CTA pair).  Copies the table to `out16` (host) and clears
 * it; SVK_E_UNSUPPORTED unless SVK_PROF=1.  No reference counterpart: tests/prof_conv.py, tests/prof_wgrad.py. */
int svk_debug_prof_read(unsigned long long* out16);

#ifdef __cplusplus
}
#endif
#endif /* SVK_H_ */
