/*
 * pdb200.h — C ABI of libpdb200.so: the sm_100a kernels behind PartDistillation's Mask2Former
 * training hot path.  Citations (file:line) are into the reference checkout
 * (facebookresearch/PartDistillation, part_distillation/...), naming the interface each entry
 * point replaces.
 *
 * Conventions (all entry points):
 *   - every tensor argument is a raw DEVICE pointer to a contiguous buffer of the documented
 *     shape; small shape tables (spatial shapes, level starts, per-image offsets) are HOST arrays;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on that stream; the
 *     library never synchronises, never allocates and never touches the default stream;
 *   - return value 0 = launched; negative = rejected (nothing launched), the message is available
 *     from pdb_last_error() (thread-local).  Kernel launch failures are returned, not printed
 *     (the reference only printf()s them: ops/src/cuda/ms_deform_im2col_cuda.cuh:953-957);
 *   - there is no CPU implementation (as in the reference: ops/src/ms_deform_attn.h:44).
 */
#ifndef PDB200_H_
#define PDB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PDB_API __attribute__((visibility("default")))
#else
#define PDB_API
#endif

#define PDB_OK 0
#define PDB_ERR_INVALID (-1)
#define PDB_ERR_LAUNCH (-2)
#define PDB_ERR_UNSUPPORTED (-3)

#define PDB_F32 0
#define PDB_F64 1

/* ABI version of this header; bumped whenever a signature changes. */
PDB_API int pdb_abi_version(void);
/* Message of the last failing call on this thread ("" if none). */
PDB_API const char* pdb_last_error(void);
/* Number of kernel launches issued through this library by this process (bench.py's gpu_launches). */
PDB_API int64_t pdb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * MSDeformAttn — replaces MSDA.ms_deform_attn_forward / ms_deform_attn_backward
 * (ops/src/vision.cpp:19-22, ops/src/ms_deform_attn.h:27-67, ops/src/cuda/ms_deform_attn_cuda.cu:26-159)
 * and the production arithmetic ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:55-75).
 *   value  (N, S, M, D)        loc  (N, Lq, M, L, P, 2) in [0,1] as (x, y)
 *   attn   (N, Lq, M, L, P)    out  (N, Lq, M*D)
 *   shapes_hw: HOST int64[L*2] (H_l, W_l); level_start: HOST int64[L]
 * dtype PDB_F32 (any D; D==32 takes the vectorised path) or PDB_F64 (generic path).
 * backward ZERO-FILLS grad_value itself (the reference callee allocates zeros: .cu:127-129), then
 * accumulates into it; grad_loc / grad_attn are fully overwritten.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_msda_forward(const void* value, const int64_t* shapes_hw, const int64_t* level_start,
                     const void* loc, const void* attn, void* out,
                     int N, int S, int M, int D, int Lq, int L, int P, int dtype, void* stream);
PDB_API int pdb_msda_backward(const void* value, const int64_t* shapes_hw, const int64_t* level_start,
                      const void* loc, const void* attn, const void* grad_out,
                      void* grad_value, void* grad_loc, void* grad_attn,
                      int N, int S, int M, int D, int Lq, int L, int P, int dtype, void* stream);

/* Encoder path (Lq == S: the queries are the pixels of the value pyramid; D == 32, P == 4, L <= 4) with the value stored
 * as fp16, head-major (N, M, S, 32): pdb_msda_pack_value_h repacks the fp32 (N, S, M, 32) value once, pdb_msda_forward_h
 * is pdb_msda_forward on that copy (fp32 locations, weights, products and accumulation; out fp32 (N, Lq, M*32)).
 * Opt-in reduced-precision staging: halves the shared-memory bytes per tap corner. */
PDB_API int pdb_msda_pack_value_h(const float* value, void* value_h, int N, int S, int M, int D, void* stream);
PDB_API int pdb_msda_forward_h(const void* value_h, const int64_t* shapes_hw, const int64_t* level_start,
                       const float* loc, const float* attn, float* out,
                       int N, int S, int M, int D, int Lq, int L, int P, void* stream);
/* Debug / A-B timing of the encoder forward: 0 = measured dispatch (TMA-staged value tiles on >= 4 levels, L1-resident tiled
 * kernel otherwise), 1 = L1-resident kernels only, 4 = TMA-staged tiles wherever eligible; bit 1 (value 2): experimental launch
 * shape of the fp16-staged kernel (2 CTAs / SM). */
PDB_API int pdb_debug_set_msda_path(int path);

/* ------------------------------------------------------------------------------------------------
 * Mask head einsum — replaces torch.einsum("bqc,bchw->bqhw")
 * (mask2former_transformer_decoder.py:449, part_distillation_transformer_decoder.py:244).
 *   embed (B, Q, C) f32;  feat (B, HW, C) f32 PIXEL-MAJOR (the NCHW mask_features tensor in
 *   channels_last memory format);  out (B, Q, HW) f32.  C % 4 == 0, 16-byte aligned bases.
 *   forward: tcgen05 kind::tf32 with a 3-term hi/lo split of both operands (fp32-accurate logits);
 *   embed_lo (B, Q, C) = pdb_split_lo(embed) or NULL (then the kernel splits embed itself, once per pixel tile).
 * backward:  grad_embed (B,Q,C) = grad_out x feat (overwritten);
 *            grad_feat  (B,HW,C) (+)= grad_out^T x embed  (accumulate != 0 adds into grad_feat).
 * Either grad pointer may be NULL to skip it.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_mask_einsum_forward(const float* embed, const float* embed_lo, const float* feat, float* out,
                            int B, int Q, int C, int64_t HW, void* stream);
PDB_API int pdb_mask_einsum_backward(const float* embed, const float* feat, const float* grad_out,
                             float* grad_embed, float* grad_feat, int accumulate,
                             int B, int Q, int C, int64_t HW, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense fp32 contraction on the tcgen05 tensor cores (3xTF32: fp32-accurate, fp32 TMEM accumulation) —
 * replaces the cuBLAS sgemm behind nn.Linear forward / backward of the encoder and decoder layers
 * (msdeformattn.py:120-135, ops/modules/ms_deform_attn.py:102-130,
 * mask2former_transformer_decoder.py:148-208) and behind autograd of the mask-head einsum (:449).
 * For each batch item b:   C_b[m][n] (+)= sum_k A_b(m,k) * B_b(n,k)  (+ bias[n]) (ReLU)
 *   a_mn = 0: A_b(m,k) = A[b*sa + m*lda + k]   (K-major)      a_mn = 1: A[b*sa + k*lda + m]   (MN-major)
 *   b_mn = 0: B_b(n,k) = B[b*sb + n*ldb + k]                  b_mn = 1: B[b*sb + k*ldb + n]
 *   c_trans = 0: C[b*sc + m*ldc + n]                          c_trans = 1: C[b*sc + n*ldc + m]
 *   accumulate != 0: C += (red.add; C must be initialised); ksplit > 1 (split-K) requires accumulate.
 * A, B 16-byte aligned; lda, ldb, sa, sb multiples of 4 floats.  bias may be NULL.
 * relu: epilogue activation after the bias: 0 none, 1 ReLU, 2 GELU (erf form, as nn.GELU()).
 * B_lo: NULL, or the low parts of B (same layout as B) from pdb_split_lo — worthwhile when B is a weight matrix
 * shared by many row tiles (every 128-row tile would otherwise re-split it).
 * nn.Linear:  y = x W^T + b      -> A = x (K-major), B = W (K-major), bias, relu optional
 *             dx = dy W          -> A = dy (K-major), B = W (MN-major)
 *             dW = dy^T x        -> A = dy (MN-major), B = x (MN-major), split-K over the rows, accumulate
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_gemm_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N,
                    int K, int batch,
                    int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn, int b_mn,
                    int c_trans, int relu, int accumulate, int ksplit, void* stream);
/* pdb_gemm_tf32x3 with the backward of a ReLU fused into the store: C[m][n] = gate[m][n] > 0 ? sum_k A(m,k) B(n,k) : 0, gate laid
 * out like C (row-major, pitch ldc, batch stride sc).  The input gradient of the Linear behind a ReLU, dh = (dy W) * (h > 0) with
 * gate = h, in one pass instead of a GEMM and a threshold_backward pass (FFN of the encoder and decoder layers,
 * msdeformattn.py:120-124, mask2former_transformer_decoder.py:167-171 via autograd). */
PDB_API int pdb_gemm_tf32x3_gated(const float* A, const float* B, const float* B_lo, float* C, const float* gate, int M, int N, int K,
                          int batch, int64_t lda, int64_t ldb, int64_t ldc, int64_t sa, int64_t sb, int64_t sc, int a_mn,
                          int b_mn, void* stream);
/* bf16 contraction on the tensor cores (tcgen05 kind::f16, fp32 accumulation) for the autocast path (BASELINE configs[2]; the
 * reference trains under AMP: Base-COCO-InstanceSegmentation.yaml:34-35) — replaces the cuBLAS bf16 GEMM behind nn.Linear of the
 * Swin backbone and the transformer decoder under torch.autocast:
 *   C[m][n] = sum_k A[m*lda + k] * B[n*ldb + k]  (+ bias[n]) (act: 0 none, 1 ReLU, 2 GELU)
 * A (M x K), B (N x K): bf16, K contiguous, 16-byte aligned, K / lda / ldb multiples of 8; bias fp32 or NULL;
 * C row-major with pitch ldc, fp32 (out_bf16 = 0) or bf16 (out_bf16 = 1).
 * accumulate != 0: C += product (red.add; C fp32, act = 0) — e.g. straight into a parameter's gradient.  ksplit > 1 (weight
 * gradients: a few output tiles over a very long K) cuts K into that many slices whose partial sums meet in C the same way; it
 * implies accumulate, so the caller zero-fills C when it wants the plain product.
 * layout = 3: both operands MN-major — A(m,k) = A[k*lda + m], B(n,k) = B[k*ldb + n] (the contraction index is the row index of
 * the stored matrices): the weight gradient dW = dy^T x with dy and x read in place; K arbitrary.  layout = 0: as above. */
PDB_API int pdb_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int64_t lda, int64_t ldb,
                  int64_t ldc, int act, int out_bf16, int ksplit, int accumulate, int layout, void* stream);
/* out[n] (+)= sum_r x[r*N + n]: bias gradient of a Linear layer over few rows (the decoder's B*Q = 200 rows; autograd's db);
 * accumulate != 0 adds into out (a preallocated parameter gradient). */
PDB_API int pdb_col_sum(const float* x, float* out, int rows, int N, int accumulate, void* stream);
/* Same for a bf16 matrix (output gradients under torch.autocast), fp32 sums; N even. */
PDB_API int pdb_col_sum_bf16(const void* x, float* out, int rows, int N, int accumulate, void* stream);
/* Short-A variant of pdb_gemm_tf32x3 for the transformer decoder's B*Q = 200-row products (nn.Linear forward and input gradient,
 * mask2former_transformer_decoder.py:148-208): C[m][n] = sum_k A[m*lda + k] * B(n,k) (+ bias[n]) (ReLU), same 3xTF32 arithmetic
 * on mma.sync 32 x 64 tiles with no TMEM / TMA set-up.  b_mn = 0: B(n,k) = B[n*ldb + k];  b_mn = 1: B(n,k) = B[k*ldb + n].
 * K, lda, ldb multiples of 4 (N too when b_mn); A, B 16-byte aligned.  ksplit > 1 splits K over blockIdx.z and red.adds into C,
 * which the caller must have zero-filled; it excludes relu. */
PDB_API int pdb_gemm_small_tf32x3(const float* A, const float* B, float* C, const float* bias, int M, int N, int K, int64_t lda,
                          int64_t ldb, int64_t ldc, int b_mn, int relu, int ksplit, void* stream);
/* Convolution-shaped variant: K = taps * Ck, and the k range of tap t reads the A rows shifted by tap_off[t]:
 *   C_b[m][n] = sum_t sum_c A_b[m + tap_off[t]][c] * B[n][t * Ck + c]  (+ bias[n]) (ReLU)
 * A_b = A + b*sa is (a_rows x Ck), K-major, rows beyond a_rows read as 0; B is (N x taps*Ck), K-major, shared by all batch
 * items; Ck % 32 == 0; tap_off: HOST int32[taps] >= 0.  With the zero-padded NHWC image as A, M = H * (W + 2) and
 * tap_off[ky*3+kx] = ky * (W + 2) + kx this is the 3x3 convolution of the pixel decoder's output layer
 * (msdeformattn.py:275-287, F.conv2d in detectron2's Conv2d) as ONE tensor-core GEMM over the padded-width pixel grid. */
PDB_API int pdb_gemm_taps_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N,
                         int Ck, int batch, int a_rows, int64_t lda, int64_t ldc, int64_t sa, int64_t sc, int taps,
                         const int32_t* tap_off, int relu, void* stream);
/* Same, with a store that drops the garbage columns of the padded-width grid: row m = y * wp + x of the product goes to row
 * y * w + x of C when x < w — C is the (H, w, N) map itself (sc = H * w * N), no crop pass afterwards.  N > 112, M % wp == 0. */
PDB_API int pdb_gemm_taps_cropped_tf32x3(const float* A, const float* B, const float* B_lo, float* C, const float* bias, int M, int N,
                                 int Ck, int batch, int a_rows, int64_t lda, int64_t ldc, int64_t sa, int64_t sc, int taps,
                                 const int32_t* tap_off, int relu, int wp, int w, void* stream);
/* lo[i] = x[i] - trunc_tf32(x[i]) (x with its low 13 mantissa bits cleared); n % 4 == 0, 16-byte aligned. */
PDB_API int pdb_split_lo(const float* x, float* lo, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attention-mask build — replaces F.interpolate(bilinear, align_corners=False) -> sigmoid() < 0.5
 * -> repeat over heads (mask2former_transformer_decoder.py:453-457) and the all-masked-row reset
 * of the next layer (:405).
 *   logits (B, Q, H, W) f32  ->  mask (B, Q, h*w) uint8, 1 = key NOT attended.  The mask is the same
 *   for every head, so it is stored once per (b, q) instead of B*heads times.
 *   row_any (B*Q) int32 workspace: pass zero-filled; after the call row_any[r] != 0 iff row r
 *   has at least one attended key.  Rows with no attended key are treated as "attend everywhere"
 *   by pdb_masked_xattn_* (that is the reference's reset at :405); pdb_attn_mask_reset_rows applies
 *   the same reset to the stored mask for callers that want the reference's tensor.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_attn_mask_build(const float* logits, uint8_t* mask, int32_t* row_any,
                        int B, int Q, int H, int W, int h, int w, void* stream);
PDB_API int pdb_attn_mask_reset_rows(uint8_t* mask, const int32_t* row_any, int rows, int64_t hw, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Masked cross-attention core — replaces the scaled-dot-product inside nn.MultiheadAttention as used by
 * CrossAttentionLayer (mask2former_transformer_decoder.py:84,102-114): softmax(q k^T + mask) v per head.
 *   q (B, Q, heads*d) f32, already multiplied by 1/sqrt(d);  k, v (B, Lk, heads*d) f32;
 *   mask (B, Q, Lk) uint8 (1 = masked) or NULL;  row_any (B*Q) int32 or NULL (see above);
 *   out (B, Q, heads*d);  lse (B, heads, Q) f32 log-sum-exp saved for backward.
 *   workspace: pdb_masked_xattn_workspace_bytes(...) bytes.
 * backward overwrites grad_q, grad_k, grad_v.  d must be 32.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int64_t pdb_masked_xattn_workspace_bytes(int B, int heads, int Q, int Lk, int d);
PDB_API int pdb_masked_xattn_forward(const float* q, const float* k, const float* v, const uint8_t* mask,
                             const int32_t* row_any, float* out, float* lse, void* workspace,
                             int B, int heads, int Q, int Lk, int d, void* stream);
PDB_API int pdb_masked_xattn_backward(const float* q, const float* k, const float* v, const uint8_t* mask,
                              const int32_t* row_any, const float* out, const float* lse,
                              const float* grad_out, float* grad_q, float* grad_k, float* grad_v,
                              int B, int heads, int Q, int Lk, int d, void* stream);
/* Arithmetic of the tensor-core attention kernels for the calls that follow (process-wide; a captured CUDA graph keeps what was
 * set at capture): 3 = 3xTF32 products, fp32-accurate (default); 1 = one TF32 product per MMA — for torch.autocast(bfloat16)
 * regions, where the reference's SDPA rounds q, k, p and v to bf16 (8-bit mantissas; TF32 keeps 10).  Returns the previous value. */
PDB_API int pdb_set_xattn_passes(int passes);

/* ------------------------------------------------------------------------------------------------
 * Point sampling — replaces detectron2 point_sample == F.grid_sample(input, 2*coords-1,
 * bilinear, zeros, align_corners=False) at its call sites criterion.py:178-196, matcher.py:130-140.
 *   src: R_src maps of (H, W), f32 (src_dtype 0), uint8 0/1 (src_dtype 1) or bit-packed (src_dtype 2: int32 words
 *   (H, ceil(W / 32)) per map, bit x % 32 of word x / 32, the layout of pdb_pack_bits — ground-truth masks sampled straight
 *   from the words they arrived in);
 *   map_index (R) int32 or NULL (identity): which map row r samples from;
 *   coords (Rc, P, 2) f32 (x, y) in [0,1]; coord_index (R) int32 or NULL (identity; a constant 0
 *   table shares one point set across rows as matcher.py:128 does);
 *   out (R, P) f32.
 * backward (f32 maps only): grad_src (R_src, H, W) must be zero-filled by the caller; accumulated.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_point_sample_forward(const void* src, int src_dtype, const int32_t* map_index,
                             const float* coords, const int32_t* coord_index, float* out,
                             int R, int P, int H, int W, void* stream);
PDB_API int pdb_point_sample_backward(const float* grad_out, const int32_t* map_index, const float* coords,
                              const int32_t* coord_index, float* grad_src,
                              int R, int P, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Matcher cost — replaces batch_sigmoid_ce_loss / batch_dice_loss and the weighted sum
 * (matcher.py:19-66,108-158) for every image of the batch in one launch.
 *   pred_pts (B*Q, P) f32 sampled logits; tgt_pts (Ktot, P) f32 sampled gt (0..1);
 *   cls_prob (B*Q, Kc) f32 class probabilities (softmax / sigmoid already applied);
 *   tgt_label (Ktot) int32; tgt_offset: HOST int32[B+1] prefix offsets of each image's targets;
 *   cost (sum_b Q*K_b) f32, image b's (Q, K_b) row-major block starting at Q*tgt_offset[b].
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_matcher_cost(const float* pred_pts, const float* tgt_pts, const float* cls_prob,
                     const int32_t* tgt_label, const int32_t* tgt_offset, float* cost,
                     int B, int Q, int Kc, int P, float w_class, float w_mask, float w_dice, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched rectangular linear sum assignment — replaces C.cpu() + scipy linear_sum_assignment +
 * the ascending-cost re-ordering (matcher.py:159-163).  One warp per image; float64 arithmetic
 * and tie-breaking follow SciPy's shortest-augmenting-path solver.
 *   cost as produced by pdb_matcher_cost; tgt_offset HOST int32[B+1];
 *   pred_idx / tgt_idx (Ktot) int64: image b's min(Q, K_b) matches at [tgt_offset[b], ...),
 *   ordered by ascending matched cost; when K_b > Q the tail of the block is filled with -1.
 * Limits: Q <= 1024, K_b <= 1024.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_lsap_batched(const float* cost, const int32_t* tgt_offset, int64_t* pred_idx, int64_t* tgt_idx,
                     int B, int Q, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Point-sampled mask loss — replaces point_sample x2 + sigmoid_ce_loss + dice_loss
 * (criterion.py:25-69,188-206) fused over all matched pairs.
 *   pred (Rp, H, W) f32 mask logits; pred_index (Nm) int64 (-1 entries are skipped);
 *   gt (Rg, Hg, Wg) uint8 (gt_bits 0) or bit-packed int32 words (Rg, Hg, ceil(Wg / 32)) (gt_bits 1); gt_index (Nm) int64;
 *   coords (Nm, P, 2) f32;
 *   sums (Nm, 4) f32 out: [sum BCE, sum s*t, sum s, sum t] per pair (loss assembly is host-side:
 *   loss_mask = sum_i BCE_i/P / num_masks, loss_dice = sum_i (1-(2 st+1)/(s+t+1)) / num_masks).
 * backward: g_bce, g_dice (Nm) f32 = d loss / d (BCE_i/P), d loss / d dice_i; grad_pred (Rp, H, W)
 * must be zero-filled by the caller; accumulated.
 * splits > 1 (few pairs: a dozen per decoder output at B = 2): the P points of a pair are cut into `splits` chunks handled by
 * different CTAs; forward writes their sums to `partial` (Nm * splits * 4 f32) and a second launch adds them in split order
 * (deterministic); backward only needs the count.  splits <= 1: one CTA of 1024 threads per pair, partial may be NULL.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_point_loss_forward(const float* pred, const int64_t* pred_index, const void* gt,
                           const int64_t* gt_index, const float* coords, float* sums, float* partial, int splits,
                           int Nm, int P, int H, int W, int Hg, int Wg, int gt_bits, void* stream);
PDB_API int pdb_point_loss_backward(const float* pred, const int64_t* pred_index, const void* gt,
                            const int64_t* gt_index, const float* coords, const float* sums,
                            const float* g_bce, const float* g_dice, float* grad_pred, int splits,
                            int Nm, int P, int H, int W, int Hg, int Wg, int gt_bits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PartDistillation classifier rows — replaces the float64 Linear(256, P*O+1) followed by
 * apply_gradient_mask (part_distillation_transformer_decoder.py:107,215-230,237-238): only the
 * P columns of each image's object class plus the last (no-object) column are ever used.
 *   x (B, Q, C) f32 (promoted to f64 inside); weight (Ncls, C) f64; bias (Ncls) f64;
 *   obj (B) int32 object class per image; out (B, Q, Pn+1) f64.
 * backward: grad_x (B,Q,C) f32 overwritten; grad_weight (Ncls, C) / grad_bias (Ncls) f64 must be
 * zero-filled by the caller (dense zero rows keep AdamW parity); rows are accumulated atomically.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_class_rows_forward(const float* x, const double* weight, const double* bias, const int32_t* obj,
                           double* out, int B, int Q, int C, int Pn, int64_t Ncls, void* stream);
PDB_API int pdb_class_rows_backward(const float* x, const double* weight, const int32_t* obj, const double* grad_out,
                            float* grad_x, double* grad_weight, double* grad_bias,
                            int B, int Q, int C, int Pn, int64_t Ncls, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step on the flat gradient buffer — replaces FullModelGradientClippingOptimizer
 * (clip_grad_norm_ over all parameters, CLIP_VALUE 0.01) + torch.optim.AdamW with per-parameter lr / weight
 * decay groups (base_trainer.py:65-147), which launch one kernel per parameter group.
 *   pdb_grad_sumsq: out (device float64 scalar) = sum_i (grad_scale * grad[i])^2; n % 4 == 0.
 *   pdb_adamw_flat: param / grad / exp_avg / exp_avg_sq are flat fp32 buffers of n elements (n % 4 == 0, 16-byte
 *     aligned) in which every parameter occupies a 4-element-aligned segment; seg_start (DEVICE int64[num_segs],
 *     ascending, seg_start[0] == 0), seg_lr / seg_wd (DEVICE float[num_segs]).  Gradients are multiplied by
 *     grad_scale (1 / world size after the all-reduce) and by min(1, clip_norm / (sqrt(*sumsq) + 1e-6)) when
 *     clip_norm > 0 and sumsq != NULL; `step` points to the 1-based step count in DEVICE memory (bias
 *     corrections are computed on the device, so a captured CUDA graph of the step stays valid).
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_grad_sumsq(const float* grad, int64_t n, float grad_scale, double* out, void* stream);
PDB_API int pdb_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   const int64_t* seg_start, const float* seg_lr, const float* seg_wd, int num_segs, float beta1,
                   float beta2, float eps, const int64_t* step, float grad_scale, float clip_norm, const double* sumsq,
                   void* stream);

/* Whole (shifted-)window attention of a Swin block on the un-partitioned token grid: F.pad, torch.roll,
 * window_partition, the shift mask, window_reverse, the inverse roll and the crop of SwinTransformerBlock.forward
 * (modeling/backbone/swin.py:239-300) become index arithmetic of the kernel.
 *   qkv (B, H, W, 3, heads, 32) f32 (qkv Linear of norm1(x), token order); qkv_bias (3*heads*32) f32 or NULL: the
 *   qkv of padded tokens (the reference pads the normalised map with zeros); bias (heads, ws*ws, ws*ws);
 *   out (B, H, W, heads*32).  ws*ws <= 256, 0 <= shift < ws, head dim 32. */
PDB_API int pdb_swin_window_attention_forward(const float* qkv, const float* qkv_bias, const float* bias, float* out, int B,
                                      int H, int W, int heads, int d, int ws, int shift, float scale, void* stream);
/* Same contract on the tensor cores (warp-level mma.sync m16n8k8 TF32; csrc/window_attn_mma.cu): passes = 3 is fp32-accurate
 * (hi / lo operand split), passes = 1 a single TF32 pass for the bf16-autocast path.  Window sizes 12, 8, 4; d = 32.
 * out_bf16 != 0: out is bf16 (the input of the bf16 projection GEMM under autocast), else f32. */
PDB_API int pdb_swin_window_attention_forward_tc(const float* qkv, const float* qkv_bias, const float* bias, void* out, int B,
                                         int H, int W, int heads, int d, int ws, int shift, float scale, int passes,
                                         int out_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Pixel grouping affinity — replaces, inside PixelGroupingModel.generate_part_segments
 * (pixel_grouping_model.py:139-144,197-211), the bilinear up-sampling of the backbone features to image size,
 * the host copy of the masked pixels, measure_distance() and topk(1), and the label-map scatter.
 *   feat (C, h, w) f32 backbone-resolution features; centroids (Kc, C) f32 (k-means centres, Kc <= 16);
 *   mask (H, W) uint8 object mask at image size; labels (H, W) int32 out: 0 outside the mask, else
 *   1 + argmax_k score_k with score = <f, c_k> (metric 0, "dot") or 2<f, c_k> - |c_k|^2 (metric 1, "l2"),
 *   f = the feature bilinearly interpolated (align_corners=False) at that pixel.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_group_affinity(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                       int C, int Kc, int h, int w, int H, int W, int metric, void* stream);
/* Stage 1 of the two-stage grouping (what functional.group_affinity runs): bilinear up-sampling is linear, so the contraction
 * with the centroids is done once at feature resolution, scores (Kc, h, w): <f, c_k> (metric 0) or 2<f, c_k> - |c_k|^2
 * (metric 1); pdb_group_affinity[_resized] is then called on the Kc score maps with the identity as centroids (metric 0):
 * C*h*w reads instead of 4*C reads per output pixel.  Labels can differ from the one-stage evaluation only at numerical
 * near-ties of the two best scores (different rounding order). */
PDB_API int pdb_group_scores(const float* feat, const float* centroids, float* scores, int C, int Kc, int h, int w, int metric,
                     void* stream);
/* B images of one geometry in TWO launches (both stages batched over the grid): feat (B, C, h, w), centroids (B, Kc, C),
 * mask / labels (B, H, W); scores: workspace of B*Kc*h*w floats; identity: the Kc x Kc identity matrix (device). */
PDB_API int pdb_group_affinity_batched(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                               float* scores, const float* identity, int B, int C, int Kc, int h, int w, int H, int W,
                               int metric, void* stream);
/* Same, for an evaluation size different from the padded batch size (pixel_grouping_model.py:139-160,
 * proposal_generation_model.py:139-155): feat (C, h, w) -> bilinear -> (Hp, Wp) -> crop (Hi, Wi) -> bilinear -> (Ho, Wo)
 * (detectron2 sem_seg_postprocess), both passes composed per output pixel; mask and labels are (Ho, Wo). */
PDB_API int pdb_group_affinity_resized(const float* feat, const float* centroids, const uint8_t* mask, int32_t* labels,
                               int C, int Kc, int h, int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo, int metric,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Swin window attention, forward only (frozen backbone) — replaces q @ k^T * scale + relative-position bias
 * (+ shift mask) -> softmax -> @ v -> head transpose inside WindowAttention.forward (modeling/backbone/swin.py:78-176).
 *   qkv (Bw, N, 3, heads, 32) f32; bias (heads, N, N) f32; mask (nW, N, N) f32 additive or NULL (row bw uses window
 *   bw % nW); out (Bw, N, heads*32) f32.  N <= 256, head dim 32.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_window_attention_forward(const float* qkv, const float* bias, const float* mask, float* out, int Bw, int N,
                                 int heads, int d, int nW, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm forward, optionally fused with the residual add in front of it — replaces nn.LayerNorm of the encoder
 * layers (msdeformattn.py:129-133), decoder layers (mask2former_transformer_decoder.py:44-54,102-114,167-171) and Swin
 * blocks (swin.py:239-300):   z = x (+ residual);  y = (z - mean) * rstd * weight + bias.
 *   x, residual (or NULL), y, sum_out (or NULL; receives z): (rows, C) f32; weight, bias (C); mean, rstd (rows) f32
 *   are saved for the backward pass (same meaning as ATen's native_layer_norm).  C % 4 == 0, C <= 2048.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_layer_norm_forward(const float* x, const float* residual, const float* weight, const float* bias, float* y,
                           float* sum_out, float* mean, float* rstd, int64_t rows, int C, float eps, void* stream);
/* Same with stochastic depth on the residual branch (swin_transformer.py:131-134, x = shortcut + drop_path(branch)):
 * y = LayerNorm(x + res_scale[row / rows_per_sample] * residual); res_scale: one f32 factor per sample (keep / (1 - p)), NULL = 1.
 * y_bf16 != 0: y is (rows, C) bf16 (round to nearest even) — the input of a bf16 GEMM under torch.autocast, without the separate
 * conversion pass; sum_out, mean, rstd stay f32. */
PDB_API int pdb_layer_norm_forward_scaled(const float* x, const float* residual, const float* res_scale, int64_t rows_per_sample,
                                  const float* weight, const float* bias, void* y, float* sum_out, float* mean, float* rstd,
                                  int64_t rows, int C, float eps, int y_bf16, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FPN top-down step of the pixel decoder (msdeformattn.py:352-356): out = lateral + F.interpolate(x, size=(H, W),
 * mode="bilinear", align_corners=False) on channels-last maps, and the gradient with respect to x as a gather (deterministic).
 *   x (B, h, w, C) f32 pixel-major, batch stride x_batch_stride elements; lateral (B, H, W, C) or NULL; out / grad_out
 *   (B, H, W, C); grad_x (B, h, w, C) overwritten.  C % 4 == 0, 16-byte aligned.  Source index and weights as ATen's
 *   upsample_bilinear2d (scale = in / out, negative source clamped to 0, upper neighbour clamped to the border).
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_upsample_add_forward(const float* x, const float* lateral, float* out, int B, int h, int w, int H, int W, int C,
                             int64_t x_batch_stride, void* stream);
PDB_API int pdb_upsample_backward(const float* grad_out, float* grad_x, int B, int h, int w, int H, int W, int C, void* stream);
/* Zero-padded copy of a channels-last map, (B, H, W, C) -> (B, top + H + bottom, left + W + right, C), borders written in the same
 * pass: the A operand of the tap-shifted 3x3 convolution GEMM (top 1, bottom 2, left 1, right 1).  C % 4 == 0. */
PDB_API int pdb_pad_nhwc(const float* x, float* out, int B, int H, int W, int C, int top, int bottom, int left, int right,
                 void* stream);

/* ------------------------------------------------------------------------------------------------
 * GroupNorm (+ ReLU) over channels-last maps — replaces nn.GroupNorm(32, C) and the F.relu behind it on the pixel decoder's
 * input_proj / lateral / output convolutions (msdeformattn.py:249-287 via detectron2 Conv2d(norm=get_norm("GN", C),
 * activation=F.relu)).  ATen's kernel wants NCHW; the convolutions here produce pixel-major maps.
 *   x, y, dy, dx: (B, HW, C) f32 = the channels-last memory of the logical (B, C, H, W) tensor; G groups of C / G
 *   consecutive channels (C / G a multiple of 4, C / 4 a divisor of 256); weight, bias (C) f32; mean, rstd (B, G) f32.
 * forward:  stats (B, G, 2) f64 workspace, ZERO-FILLED by the caller (sum, sum of squares); relu != 0 applies max(., 0).
 * backward: chan_sums (B, C, 2) f64, ZERO-FILLED by the caller; on return [b][c][0] = sum_hw dy' * xhat and
 *   [b][c][1] = sum_hw dy' (dy' = dy, or dy where the forward output was positive): grad_weight = sum_b [..][0],
 *   grad_bias = sum_b [..][1] (host side); coef (B, G, 2) f32 workspace; dx may be NULL (parameter gradients only).
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_group_norm_forward(const float* x, const float* weight, const float* bias, float* y, double* stats, float* mean,
                           float* rstd, int B, int64_t HW, int C, int G, float eps, int relu, void* stream);
PDB_API int pdb_group_norm_backward(const float* dy, const float* x, const float* weight, const float* bias, const float* mean,
                            const float* rstd, double* chan_sums, float* coef, float* dx, int B, int64_t HW, int C, int G,
                            int relu, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inference post-processing (ProposalModel eval branch) — replaces, per image,
 * F.interpolate(pred_masks -> padded size) (proposal_model.py:225-230), detectron2 sem_seg_postprocess (crop +
 * second bilinear resize, :243-245), masking_with_object_mask (:369-376), _unique_assignment's `> 0`, top-1 object
 * map and score * sigmoid argmax (:258-302), and get_iou_all_cocoapi's host RLE round trip (utils/utils.py:35-42).
 * Masks live bit-packed: `bits` rows are (Ho, Ww) uint32 words, Ww = ceil(Wo / 32), bit i of a word = pixel
 * 32 * word + i of that image row.
 *
 * pdb_postprocess_masks:  logits (Q, h, w) f32 (one image of pred_masks); sel (K) int32 query indices (the top-k);
 *   scores (K) f32 (NULL unless label or score_bits is requested); gate (Ho, Wo) uint8 object mask or NULL;
 *   bits (K + 1, Ho, Ww) uint32 or NULL: row k = [resized(logits[sel[k]]) * gate > 0], row K = OR of the K rows
 *   (the reference's `topk(1, dim=0)[0] > 0` object map); label (Ho, Wo) int32 or NULL = argmax_k scores[k] *
 *   sigmoid(resized * gate) (first maximum).  (h, w) -> bilinear -> (Hp, Wp) -> crop (Hi, Wi) -> bilinear ->
 *   (Ho, Wo), both passes align_corners=False, composed per output pixel.  score_bits (K, Ho, Ww) uint32 or NULL:
 *   row k = [scores[k] * sigmoid(resized * gate) > score_thr] (PartDistillationModel's `predmask > 0.5` area filter and
 *   its `predmask > 0` masks, part_distillation_model.py:379-394).
 * pdb_resize_masks_u8:  masks (G, Hp, Wp) uint8 0/1 -> out (G, Ho, Wo) uint8 = [bilinear(float(crop (Hi, Wi))) != 0]
 *   (sem_seg_postprocess(target["masks"].float(), ...).bool(), :244-245).
 * pdb_pack_bits / pdb_unpack_bits:  (R, Ho, Wo) uint8 (non-zero = set) <-> (R, Ho, Ww) words; unpack gathers rows[r]
 *   (int32, NULL = identity).
 * pdb_bits_popcount:  counts[r] += popcount(bits[r, :words]); counts (rows) int64, ZERO-FILLED by the caller.
 * pdb_bits_intersect: inter[i, j] += popcount(a[i] & b[j]); inter (Ka, Kb) int64, ZERO-FILLED by the caller.
 *   IoU (pycocotools rleIou, iscrowd = 0) = inter / (|a| + |b| - inter), exactly 0 where inter == 0.
 * ---------------------------------------------------------------------------------------------- */
PDB_API int pdb_postprocess_masks(const float* logits, const int32_t* sel, const float* scores, const uint8_t* gate,
                          uint32_t* bits, int32_t* label, uint32_t* score_bits, float score_thr, int Q, int K, int h,
                          int w, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo, void* stream);
PDB_API int pdb_resize_masks_u8(const uint8_t* masks, uint8_t* out, int G, int Hp, int Wp, int Hi, int Wi, int Ho, int Wo,
                        void* stream);
PDB_API int pdb_pack_bits(const uint8_t* in, uint32_t* bits, int R, int Ho, int Wo, void* stream);
PDB_API int pdb_unpack_bits(const uint32_t* bits, const int32_t* rows, uint8_t* out, int R, int Ho, int Wo, void* stream);
PDB_API int pdb_bits_popcount(const uint32_t* bits, int64_t* counts, int rows, int64_t words, void* stream);
PDB_API int pdb_bits_intersect(const uint32_t* a, const uint32_t* b, int64_t* inter, int Ka, int Kb, int64_t words,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PDB200_H_ */
