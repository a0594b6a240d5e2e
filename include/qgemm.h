/*
 * qgemm.h -- C ABI of the B200-native quantized-linear hot path (libqgemm.so).
 *
 * Drop-in boundary for the reference's operator layer (header-only C++ templates over
 * Tensor<T>, the .cuh files under /root/reference/src/ops).  Each entry point names the reference interface it
 * replaces (file:line relative to /root/reference).  The reference-shaped C++ templates
 * (op_quantized_mm, op_mm, op_absmax, ... over a Tensor<T> view) that forward to these
 * functions live in quantized-gemm-for-transformer-inference_b200/cpp/; INTEGRATION.md shows
 * the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrix pointers are DEVICE pointers unless the name
 *     ends in _host; matrices are row-major with a leading dimension in ELEMENTS (the
 *     reference's stride_h with stride_w == 1, src/utils/tensor.cuh:14,62-63).
 *   - parameter names are lower case on purpose: the reference defines a macro `N`
 *     (src/ops/op_elemwise.cuh:10) that would otherwise rewrite these prototypes.
 *   - X is [M,K] (activations), W is [K,N] (in_dim x out_dim, src/modules/linear.cuh:18),
 *     bias is [N] (the reference's [1,N]), Cx is [M], Cw is [N], codes are int8.
 *   - return value: 0 = QG_OK, QG_E* (negative) for argument errors, positive = cudaError_t.
 *     qg_last_error() gives a per-thread message.  (The reference asserts; the C++ shim turns a
 *     non-zero status back into an assert.)
 *   - every launch goes to `stream` (a cudaStream_t; NULL = the legacy default stream the
 *     reference uses).  No call synchronises the device; no call allocates unless stated.
 *   - there is no CPU fallback: without an sm_100 device every compute call fails.
 *   - re-entrancy: calls on different devices are independent, and so are calls on different streams of
 *     one device as long as each passes its OWN workspace (qg_workspace_bytes covers every temporary of
 *     qg_quantized_mm / qg_linear_forward*, the int32 slice matrices of a split-K product included) and its
 *     own prepared weights.  The column-maximum scratch of the weight quantizers is kept per (device,
 *     stream).  What IS shared per device, as in the reference (everything on the legacy default stream):
 *     the grow-only buffers behind workspace == NULL, qg_quantized_mm_host and qg_attention_forward, and the
 *     split-K slices of the entry points that take no workspace argument (qg_gemm_s8*); growing one of those
 *     buffers allocates and synchronises the device once.
 */
#ifndef QGEMM_H_
#define QGEMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QG_API __attribute__((visibility("default")))

typedef void *qg_stream_t; /* cudaStream_t */

enum {
  QG_OK = 0,
  QG_EINVAL = -22,  /* bad shape / pointer / alignment */
  QG_ENOTSUP = -95, /* valid request this build does not implement */
  QG_ENOMEM = -12,  /* workspace too small */
  QG_ENODEV = -19   /* no sm_100 device */
};

/* I/O element types of X / W / O (all arithmetic is fp32 / int8 / int32). */
enum { QG_F32 = 0, QG_F16 = 1, QG_BF16 = 2, QG_S32 = 3 };

/* QG_MODE_REF_EXACT reproduces op_absmax's signed first element
 * (src/ops/op_reduction.cuh:80,105); QG_MODE_TRUE_ABSMAX folds |x[0]| like every other entry. */
enum { QG_MODE_REF_EXACT = 0, QG_MODE_TRUE_ABSMAX = 1 };

/* gemm kernel selection (QG_GEMM_AUTO in production; the others pin one variant for tests) */
enum {
  QG_GEMM_AUTO = 0,
  QG_GEMM_SIMT = 1,      /* generic CUDA-core kernel: any shape / leading dimension        */
  QG_GEMM_TC_1SM = 2,    /* tcgen05 cta_group::1, 128x256 tile per CTA                      */
  QG_GEMM_TC_2SM = 3     /* tcgen05 cta_group::2, 256x256 tile per CTA pair                 */
};

QG_API int qg_version(void);
QG_API const char *qg_last_error(void);
/* fills whichever pointers are non-NULL for the current device */
QG_API int qg_device_info(int *sm_count, int *cc_major, int *cc_minor);
/* process-wide kernel choice for the GEMM entry points (tests / bench ablations) */
QG_API int qg_set_gemm_variant(int variant);
/* SMs the tensor-core GEMM may occupy (0 = all, the default).  The GEMM is persistent and takes every SM it is given; a caller that
 * overlaps its own small kernels with it (the Megatron pair's reduce + gather of one row block under the next block's
 * products) leaves a few SMs free for them.  Process-wide. */
QG_API int qg_set_gemm_sm_limit(int sms);
/* number of kernels this library has launched since load / last reset (bench "gpu_launches") */
QG_API int64_t qg_launch_count(int reset);

/* ---- a1 / a2: op_absmax(in, out)  src/ops/op_reduction.cuh:195-204 ------------------------ */
/* out [M,1]: per-row reduce (op_reduction_kernel_colwise, :71-92) */
QG_API int qg_absmax_rows(const void *X, int dtype, int m, int k, int64_t ldx, int mode, float *Cx,
                          qg_stream_t stream);
/* out [1,N]: per-column reduce (op_reduction_kernel_rowwise, :96-117) */
QG_API int qg_absmax_cols(const void *W, int dtype, int k, int n, int64_t ldw, int mode, float *Cw,
                          qg_stream_t stream);

/* ---- a3: op_inv_divide(a, b, out) = b / a   src/ops/op_elemwise.cuh:657-667 --------------- */
QG_API int qg_inv_divide_f32(const float *a, int64_t n, float b, float *out, qg_stream_t stream);

/* ---- a4: op_multiply<T,int8_t>(a, scale, out)  src/ops/op_elemwise.cuh:629-640 ------------- */
/* scale [M,1] broadcast along columns (:416-419) */
QG_API int qg_quantize_rows(const void *X, int dtype, int m, int k, int64_t ldx, const float *sx,
                            int8_t *Xq, int64_t ldq, qg_stream_t stream);
/* scale [1,N] broadcast along rows (:412-415) */
QG_API int qg_quantize_cols(const void *W, int dtype, int k, int n, int64_t ldw, const float *sw,
                            int8_t *Wq, int64_t ldq, qg_stream_t stream);

/* ---- a1+a3+a4 fused: one pass over X writes int8 codes and the fp32 absmax ---------------- */
/* replaces op_absmax + op_inv_divide + op_multiply at src/ops/op_mm.cuh:76-77,82-83,86-87 */
QG_API int qg_absmax_quant_rows(const void *X, int dtype, int m, int k, int64_t ldx, float range,
                                int mode, int8_t *Xq, int64_t ldq, float *Cx, qg_stream_t stream);
/* replaces src/ops/op_mm.cuh:78-79,84-85,88-89.  scratch: unused since the library keeps its own
 * epoch-tagged column-max scratch (parameter kept for ABI stability; pass NULL). */
QG_API int qg_absmax_quant_cols(const void *W, int dtype, int k, int n, int64_t ldw, float range,
                                int mode, int8_t *Wq, int64_t ldq, float *Cw, float *scratch,
                                qg_stream_t stream);
/* Both quantizers of one op_quantized_mm call (a1-a4 on X and on W) -- same codes and scales as the two calls above.  Large
 * problems whose W still fits in L2 run column pass 1, then column pass 2 (L2-bound) SIDE BY SIDE with the row quantizer
 * (HBM-bound) in one launch, which is what qg_quantized_mm does internally. */
QG_API int qg_absmax_quant_rows_cols(const void *X, int64_t ldx, const void *W, int64_t ldw, int dtype, int m, int n, int k,
                                     float range, int mode, int8_t *Xq, int64_t ldxq, float *Cx, int8_t *Wq, int64_t ldwq,
                                     float *Cw, qg_stream_t stream);

/* ---- a5: op_mm<int8_t,int>(A, B, C)  src/ops/op_mm.cuh:49-65 ------------------------------ */
/* exact int32 accumulation (equals the reference's fp32-FMA accumulator while every partial
 * sum stays below 2^24).  A [M,K] lda, B [K,N] ldb, C [M,N] int32 ldc. */
QG_API int qg_gemm_s8s8s32(const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int m, int n,
                           int k, int32_t *C, int64_t ldc, qg_stream_t stream);

/* ---- a6+a7+a8 (+a10): op_mm(Cx,Cw,outer); op_dequantize; op_multiply(O,1/range^2) ---------- */
/* src/ops/op_mm.cuh:96-99, src/ops/op_elemwise.cuh:614-625,644-654; bias add = op_add at
 * src/modules/linear.cuh:54 (bias may be NULL).  Unfused form, for callers holding accumulators. */
QG_API int qg_dequantize_s32(const int32_t *acc, int64_t ldacc, const float *Cx, const float *Cw,
                             const float *bias, int m, int n, float range, void *O, int out_dtype,
                             int64_t ldo, qg_stream_t stream);

/* ---- a5..a8 (+a10) fused: int8 GEMM whose epilogue dequantizes, adds bias and casts --------- */
QG_API int qg_gemm_s8_dequant(const int8_t *Xq, int64_t ldxq, const int8_t *Wq, int64_t ldwq,
                              const float *Cx, const float *Cw, const float *bias, int m, int n, int k,
                              float range, void *O, int out_dtype, int64_t ldo, qg_stream_t stream);

/* ---- a9: op_quantized_mm(X, W, O, range)  src/ops/op_mm.cuh:67-101 ------------------------- */
/* bytes of device scratch qg_quantized_mm / qg_linear_forward need for this shape */
QG_API size_t qg_workspace_bytes(int m, int n, int k);
/* X,W of in_dtype, O of out_dtype; bias optional (LinearLayer::forward, linear.cuh:49-56).
 * The reference allocates 8 temporaries per call (op_mm.cuh:76-96); here the caller passes one
 * scratch block (workspace == NULL: the library keeps a grow-only per-device arena). */
QG_API int qg_quantized_mm(const void *X, int64_t ldx, const void *W, int64_t ldw, int in_dtype, void *O,
                           int64_t ldo, int out_dtype, int m, int n, int k, float range, int mode,
                           const float *bias, void *workspace, size_t workspace_bytes,
                           qg_stream_t stream);

/* ---- LinearLayer<T>::forward(x, y) with weights quantized once ----------------------------- */
/* Column-quantize W [K,N] once into the layout the tensor-core GEMM runs fastest on: Wt [N, ldwt]
 * (codes of column j contiguous along K, ldwt >= K and a multiple of 16) plus Cw [N].  Same
 * arithmetic as qg_absmax_quant_cols; only the placement of the codes differs. */
QG_API int qg_prepare_weights(const void *W, int dtype, int k, int n, int64_t ldw, float range, int mode,
                              int8_t *Wt, int64_t ldwt, float *Cw, qg_stream_t stream);
/* a5..a8 (+a10) on prepared weights; out_dtype QG_S32 writes raw accumulators (Cx, Cw unused) */
QG_API int qg_gemm_s8t_dequant(const int8_t *Xq, int64_t ldxq, const int8_t *Wt, int64_t ldwt,
                               const float *Cx, const float *Cw, const float *bias, int m, int n, int k,
                               float range, void *O, int out_dtype, int64_t ldo, qg_stream_t stream);
/* src/modules/linear.cuh:49-56: y = x @ w + b.  Wt/Cw come from qg_prepare_weights. */
QG_API int qg_linear_forward(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt,
                             const float *Cw, const float *bias, void *Y, int64_t ldy, int out_dtype,
                             int m, int n, int k, float range, int mode, void *workspace,
                             size_t workspace_bytes, qg_stream_t stream);

/* ---- column-parallel LinearLayer::forward with the output gather fused into the GEMM epilogue -- */
/* Rank p owns output columns [lo, hi): Wt/Cw/bias are its shard (n = hi - lo).  Its [m, n] block is
 * written to y_local and to y_peers[0..n_peers) -- the other GPUs' [m, N_total] output matrices,
 * mapped into this process (CUDA IPC / symmetric memory), every pointer already offset by `lo`
 * columns, common leading dimension ldy.  Peer stores are issued by the epilogue (TMA over NVLink)
 * and overlap the main loop; the caller places a cross-GPU barrier before consuming the result. */
QG_API int qg_linear_forward_multi(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt,
                                   const float *Cw, const float *bias, void *y_local, void *const *y_peers,
                                   int n_peers, int64_t ldy, int out_dtype, int m, int n, int k, float range,
                                   int mode, void *workspace, size_t workspace_bytes, qg_stream_t stream);

/* General form of the fused GEMM (a5..a8, a10): B = Wq [K,N] (b_kmajor 0) or Wt [N,K] (1); the block is
 * also written to peers[0..n_peers) (see qg_linear_forward_multi for the pointer convention). */
QG_API int qg_gemm_s8_dequant_ex(const int8_t *Xq, int64_t ldxq, const int8_t *B, int64_t ldb, int b_kmajor,
                                 const float *Cx, const float *Cw, const float *bias, int m, int n, int k,
                                 float range, void *O, void *const *peers, int n_peers, int out_dtype,
                                 int64_t ldo, qg_stream_t stream);
/* The same exchange done by the NVSwitch (NVLS): O_mc is the multicast (multimem) address of the block O in a symmetric
 * allocation bound on every GPU of the group -- the epilogue issues ONE multimem.st per 16 bytes and the switch writes it into
 * every GPU's matrix, the caller's own included, so the sender's egress is 1x the block instead of (P-1)x.  Floating-point
 * output, n a multiple of 16 bytes' worth of elements. */
QG_API int qg_gemm_s8_dequant_mc(const int8_t *Xq, int64_t ldxq, const int8_t *B, int64_t ldb, int b_kmajor, const float *Cx,
                                 const float *Cw, const float *bias, int m, int n, int k, float range, void *O, void *O_mc,
                                 int out_dtype, int64_t ldo, qg_stream_t stream);

/* ---- host-buffer form of a9 (pageable or pinned host pointers; H2D + compute + D2H) --------- */
/* what a caller holding host tensors (Tensor<T>{h,w,false}, toDevice/toHost at
 * src/utils/tensor.cuh:77-119) would otherwise spell by hand.  Blocks until O_host is written. */
QG_API int qg_quantized_mm_host(const float *X_host, const float *W_host, float *O_host, int m, int n,
                                int k, float range, int mode, const float *bias_host);

/* ---- a11: op_outlier_extractor(a, b, out)  src/ops/op_elemwise.cuh:698-708 ----------------- */
/* elementwise mask out = (|a| <= thr) ? 0 : 1, same dtype as the reference (float mask) */
QG_API int qg_outlier_mask_f32(const float *A, int m, int k, int64_t lda, float thr, float *mask,
                               int64_t ldm, qg_stream_t stream);

/* ---- LLM.int8()-style outlier decomposition (new; the reference stops at the mask above) ------ */
/* Column reduction of the mask: ascending K-indices with at least one |x| > thr (strict, NaN counts),
 * written to idx[0..min(count,max_idx)) on the device; *count (device) receives the total. */
QG_API int qg_outlier_cols(const void *X, int dtype, int m, int k, int64_t ldx, float thr, int *idx,
                           int max_idx, int *count, qg_stream_t stream);
QG_API size_t qg_outlier_workspace_bytes(int m, int n, int k);
/* LinearLayer::forward with the feature columns idx[0..n_idx) (device; any order -- the set is what counts: duplicates and
 * out-of-range entries are dropped, both operands of the side product are filed by the column's rank; n_idx <= 64) taken
 * out of the int8 path: X's outlier columns are zeroed before the row quantizer, and
 * fp16(X[:,idx]) @ fp16(W[idx,:]) (bf16 when X is bf16) is accumulated in fp32 inside the GEMM
 * epilogue: y = fl(fl(dequant + side) + bias), the side sum an fma chain in ascending idx order (the
 * order is part of the result, which is why it runs on the CUDA cores).  Up to 16 columns the epilogue
 * double-buffers its fp32 copy of W[idx, tile]; from 17 to 64 it gives up one stage of the operand ring
 * for it.  More than 64: QG_ENOTSUP.  W is the original [K,N] weight, Wt/Cw its prepared int8 form
 * (qg_prepare_weights: column scales over ALL rows, so they do not depend on idx). */
QG_API int qg_linear_forward_outlier(const void *X, int64_t ldx, int in_dtype, const void *W, int64_t ldw,
                                     int w_dtype, const int8_t *Wt, int64_t ldwt, const float *Cw,
                                     const float *bias, const int *idx, int n_idx, void *Y, int64_t ldy,
                                     int out_dtype, int m, int n, int k, float range, int mode,
                                     void *workspace, size_t workspace_bytes, qg_stream_t stream);

/* ---- op_mm<float,float>(A, B, C)  src/ops/op_mm.cuh:49-65 ---------------------------------- */
/* fp32 product with the reference's k-ascending fma chain per output (bit-identical results);
 * general strides (sa_h, sa_w ...) because attention.cuh:58-60 passes K.transpose(). */
QG_API int qg_mm_f32(const float *A, int64_t sa_h, int64_t sa_w, const float *B, int64_t sb_h, int64_t sb_w,
                     int m, int n, int k, float *C, int64_t ldc, qg_stream_t stream);

/* ---- LinearLayer::forward followed by op_relu  src/transformer.cu:65-67 (ll1.forward; op_relu) ---------- */
/* qg_linear_forward with an activation applied after the bias add, inside the GEMM epilogue.
 * QG_ACT_RELU is ReluFunc (src/ops/op_elemwise.cuh:181-195): x < 0 ? 0 : x (NaN and -0 pass through). */
enum { QG_ACT_NONE = 0, QG_ACT_RELU = 1 };
QG_API int qg_linear_forward_act(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt,
                                 const float *Cw, const float *bias, int act, void *Y, int64_t ldy, int out_dtype,
                                 int m, int n, int k, float range, int mode, void *workspace, size_t workspace_bytes,
                                 qg_stream_t stream);

/* ---- op_add(A, R, T); op_layernorm(T, B)  src/transformer.cu:57-58,74-75, src/ops/op_layernorm.cuh:6-44 -- */
/* The stack's "ADD & NORM" with the reference kernel's arithmetic: t = fl(a + r); mean and variance by
 * ascending-order fp32 sums; b = (t - mean) / var  (the reference divides by the variance, no epsilon).
 * R == NULL: plain op_layernorm.  B may alias A or R.  Every row is computed (the reference sizes its grid
 * from the width, op_layernorm.cuh:41, and leaves rows beyond 256*ceil(w/256) untouched). */
QG_API int qg_add_layernorm_f32(const float *A, int64_t lda, const float *R, int64_t ldr, int m, int n, float *B,
                                int64_t ldb, qg_stream_t stream);

/* ---- op_multiply(A, scale, T); op_softmax(T, B)  src/modules/attention.cuh:62-68,
 *      src/ops/op_elemwise.cuh:644-655, src/ops/op_softmax.cuh:6-41 ---------------------------- */
/* Row softmax of fl(A * scale) with the reference kernel's arithmetic: max by strict '>' from column 0,
 * e = expf(t - max), sum over ascending columns, e / sum (IEEE division).  scale = 1 gives plain
 * op_softmax.  B may alias A.  (The reference sizes its grid from the width, op_softmax.cuh:38, and
 * so skips rows beyond 256*ceil(w/256); every row is computed here.) */
QG_API int qg_softmax_rows_f32(const float *A, int64_t lda, int m, int n, float scale, float *B, int64_t ldb,
                               qg_stream_t stream);

/* ---- AttentionLayer<float>::forward(X, output)  src/modules/attention.cuh:47-70, and the 3-argument
 *      forward(Xq, Xkv, output) that src/transformer.cu:37,39,104,106,132 calls ------------------- */
/* Scaled dot-product attention whose three projections (attention.cuh:54-56) go through the quantized
 * linear path (qg_quantized_mm semantics, REF_EXACT arithmetic); Q K^T, the 1/sqrt(d_k) scaling, the
 * softmax and P V keep the reference's fp32 arithmetic bit for bit.
 *   Xq  [batch*sq,  d_model]  queries' input;  Xkv [batch*skv, d_model] keys'/values' input
 *                             (Xq == Xkv, sq == skv: self-attention, one fused projection)
 *   Wqkv [d_model, heads*(2*d_k + d_v)]: columns [ W_q of head 0..H-1 | W_k of head 0..H-1 | W_v ... ];
 *        heads = 1 is the reference's AttentionLayer (W_q | W_k | W_v side by side)
 *   out [batch*sq, heads*d_v]: head h in columns h*d_v .. (the concat of transformer.cu:43-50)
 * batch sequences attend independently (the reference has no batch dimension: batch = 1). */
QG_API int qg_attention_forward(const float *Xq, int64_t ldxq, const float *Xkv, int64_t ldxkv, int batch, int sq,
                                int skv, int d_model, const float *Wqkv, int64_t ldw, int heads, int d_k, int d_v,
                                float range, int mode, float *out, int64_t ldo, qg_stream_t stream);
/* The same layer with the projection weights quantized ONCE (a layer object that keeps its weights, which the reference's
 * AttentionLayer does -- attention.cuh:20-26 -- although it re-draws them in transformer.cu's loops): Wt [heads*(2*d_k+d_v),
 * ldwt] and Cw come from qg_prepare_weights(Wqkv).  Column scales depend only on their own column, so the result is
 * bit-identical to qg_attention_forward on the same fp32 weights; the per-call column quantizer pass disappears. */
QG_API int qg_attention_forward_prepared(const float *Xq, int64_t ldxq, const float *Xkv, int64_t ldxkv, int batch, int sq,
                                         int skv, int d_model, const int8_t *Wt, int64_t ldwt, const float *Cw, int heads,
                                         int d_k, int d_v, float range, int mode, float *out, int64_t ldo,
                                         qg_stream_t stream);

/* ---- quantization carried across layers (the FFN of src/transformer.cu:63-71) -----
 * The reference's stack runs ll1.forward -> op_relu -> ll2.forward; on the quantized path every linear starts with
 * a row quantizer pass over its input (reduce every row, then quantize it).  Here the producing GEMM's epilogue
 * already knows the row maxima of what it writes, and the ADD & NORM kernel holds whole rows on chip:
 *   - qg_linear_forward_q: LinearLayer::forward (+ activation) on activations that are ALREADY int8 codes + Cx;
 *     with y_rowmax != NULL the epilogue also leaves max_{j>=1} |y[i,j]| (of the values as stored) in y_rowmax[i];
 *   - qg_quantize_rows_given_max: the next layer's quantizer without its reduction: the scale comes from
 *     (Y[i,0], y_rowmax[i]) with the reference's signed-first-element rule, then the codes -- same bits as
 *     qg_absmax_quant_rows on Y;
 *   - qg_ffn_forward: the whole ll1 -> relu -> ll2 chain in one call (X given as floats, or as codes + Cx);
 *   - qg_add_layernorm_quant_f32: ADD & NORM (qg_add_layernorm_f32) that also emits the codes + Cx of its result,
 *     for the linear layer that consumes it.
 * Results are bit-identical to the unfused sequence (LinearLayer::forward calls one after the other). */
QG_API int qg_linear_forward_q(const int8_t *Xq, int64_t ldxq, const float *Cx, const int8_t *Wt, int64_t ldwt,
                               const float *Cw, const float *bias, int act, void *Y, int64_t ldy, int out_dtype, int m,
                               int n, int k, float range, float *y_rowmax, void *workspace, size_t workspace_bytes,
                               qg_stream_t stream);
QG_API int qg_quantize_rows_given_max(const void *Y, int dtype, int m, int k, int64_t ldy, float range, int mode,
                                      const float *rowmax, int8_t *Xq, int64_t ldq, float *Cx, qg_stream_t stream);
QG_API size_t qg_ffn_workspace_bytes(int m, int d_in, int d_ff, int d_out);
/* H [m, d_ff] = relu(x @ W1 + b1) (h_dtype), Y [m, d_out] = H @ W2 + b2.  Either X (floats, in_dtype) or
 * Xq_in + Cx_in (codes from qg_add_layernorm_quant_f32 / qg_absmax_quant_rows; X may then be NULL). */
QG_API int qg_ffn_forward(const void *X, int64_t ldx, int in_dtype, const int8_t *Xq_in, int64_t ldxq_in,
                          const float *Cx_in, const int8_t *W1t, int64_t ldw1t, const float *Cw1, const float *b1,
                          const int8_t *W2t, int64_t ldw2t, const float *Cw2, const float *b2, void *H, int64_t ldh,
                          int h_dtype, void *Y, int64_t ldy, int out_dtype, int m, int d_in, int d_ff, int d_out,
                          float range, int mode, void *workspace, size_t workspace_bytes, qg_stream_t stream);
QG_API int qg_add_layernorm_quant_f32(const float *A, int64_t lda, const float *R, int64_t ldr, int m, int n, float *B,
                                      int64_t ldb, float range, int mode, int8_t *Xq, int64_t ldq, float *Cx,
                                      qg_stream_t stream);

/* ---- Megatron pairing: column-parallel fc1 -> row-parallel fc2 ----------------------
 * The FFN of src/transformer.cu:63-71 (ll1.forward -> op_relu -> ll2.forward) split over P GPUs so that the only exchange is a
 * reduce-scatter of fc2's partial products instead of an all-gather of fc1's (much larger) output:
 *   rank p holds W1[:, F_p] and W2[F_p, :] for its slice F_p of the d_ff hidden features; it computes
 *   H_p = relu(x W1[:, F_p] + b1[F_p]) (bit-identical to the same columns of the single-GPU layer), quantizes H_p row-wise
 *   with scales taken over ITS slice, and forms the partial product part_p = dequant(Hq_p . Wq2_p) with the scales of the slice
 *   (Cx2_p per row, Cw2_p per column of W2[F_p, :]).  y = ((part_0 + part_1) + ... + part_{P-1}) + b2, every addition rounded
 *   to fp32 in ascending rank order.  This differs from the single-GPU result in the last bits (per-slice scales) -- it is its
 *   own, explicitly specified mode, matched bit for bit by the CPU restatement (oracle.megatron_ffn).
 * The exchange is carried by the kernels themselves over peer-mapped memory: the GEMM epilogue of fc2 stores column block b
 * of part_p straight into slot p of the GPU that owns block b (qg_ffn_forward_rowpar / qg_gemm_s8_dequant_scatter: part_dst[b]
 * is that slot, an [m, block_cols] matrix with leading dimension ld_part; block_cols a multiple of 32 columns for fp32 and
 * 64 for 16-bit partials), and qg_reduce_partials adds the P slots in order, adds the bias and writes the owner's block to
 * `out` and to the same block of the peers' matrices (the all-gather that completes the all-reduce; n_peers = 0: keep it
 * sharded; out_mc != NULL: the block's NVSwitch multicast address, one multimem.st instead of n_peers + 1 stores; max_ctas > 0 bounds the
 * grid for a caller that runs the reduction on a side stream under a GEMM, see qg_set_gemm_sm_limit).  The caller brackets the two with barriers (all slots written before the reduce; all reduces done before the
 * next forward's stores). */
QG_API int qg_gemm_s8_dequant_scatter(const int8_t *Xq, int64_t ldxq, const int8_t *Wt, int64_t ldwt, const float *Cx,
                                      const float *Cw, int m, int n, int k, float range, void *const *part_dst, int n_dst,
                                      int block_cols, int64_t ld_part, int part_dtype, qg_stream_t stream);
QG_API int qg_ffn_forward_rowpar(const void *X, int64_t ldx, int in_dtype, const int8_t *Xq_in, int64_t ldxq_in,
                                 const float *Cx_in, const int8_t *W1t, int64_t ldw1t, const float *Cw1, const float *b1,
                                 const int8_t *W2t, int64_t ldw2t, const float *Cw2, void *H, int64_t ldh, int h_dtype,
                                 void *const *part_dst, int n_dst, int block_cols, int64_t ld_part, int part_dtype, int m,
                                 int d_in, int d_ff_local, int d_out, float range, int mode, void *workspace,
                                 size_t workspace_bytes, qg_stream_t stream);
/* [rows, width_bytes] block copy between device buffers (peer-mapped ones included) on the copy engines, no SM involved */
QG_API int qg_copy_2d_async(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes, int64_t width_bytes,
                            int rows, qg_stream_t stream);
QG_API int qg_reduce_partials(const void *slots, int64_t slot_stride, int n_slots, int part_dtype, int64_t ld_part,
                              const float *bias, void *out, void *const *peers, int n_peers, void *out_mc, int64_t ldo,
                              int out_dtype, int m, int n, int max_ctas, qg_stream_t stream);

/* ---- the elementwise tail of the pipeline, op by op ------------------------------------------------
 * The fused epilogue makes these unnecessary on the fast path; they let the reference's step-by-step
 * sequence (src/timing_quantize.cu:38-58,67-70) be re-pointed one call at a time with identical bits.
 * B is [b_rows, b_cols]: the same shape as A, [1, n] (repeated down the rows) or [m, 1] (across the
 * columns) -- the broadcast rule of op_elemwise_binary_w_bcast_kernel, src/ops/op_elemwise.cuh:404-424.
 * O may alias A. */
/* op_add(a, b, out)  src/ops/op_elemwise.cuh:501-512 (AddFunc :57-65) */
QG_API int qg_add_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O,
                      int64_t ldo, int m, int n, qg_stream_t stream);
/* op_subtract(a, b, out)  src/ops/op_elemwise.cuh:531-542 (timing_quantize.cu:68: C - qC) */
QG_API int qg_subtract_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O,
                           int64_t ldo, int m, int n, qg_stream_t stream);
/* op_multiply(a, b, out) with T == OutT == float  src/ops/op_elemwise.cuh:629-640 (MultiplyFunc :81-91) */
QG_API int qg_multiply_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O,
                           int64_t ldo, int m, int n, qg_stream_t stream);
/* op_multiply(a, T b, out)  src/ops/op_elemwise.cuh:644-654 (MultiplyConstFunc :118-129; op_mm.cuh:99) */
QG_API int qg_multiply_const_f32(const float *A, int64_t lda, float c, float *O, int64_t ldo, int m, int n,
                                 qg_stream_t stream);
/* op_relu(t, out)  src/ops/op_elemwise.cuh:454-465 (ReluFunc :181-195: x < 0 ? 0 : x) */
QG_API int qg_relu_f32(const float *A, int64_t lda, float *O, int64_t ldo, int m, int n, qg_stream_t stream);
/* op_dequantize(a, b, out)  src/ops/op_elemwise.cuh:614-625 (DequantizeFunc :93-103): out = (float)acc * outer,
 * over a MATERIALISED outer-product matrix, exactly as op_mm.cuh:96-98 spells it */
QG_API int qg_dequantize_outer_f32(const int32_t *acc, int64_t ldacc, const float *outer, int64_t ldouter, float *O,
                                   int64_t ldo, int m, int n, qg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QGEMM_H_ */
