// ref_driver.cu -- C API over the REFERENCE's own CUDA kernels.  TEST INFRASTRUCTURE ONLY.
//
// Built by oracle/Makefile with `nvcc -I /root/reference/src` into oracle/_ref/libref_qmm.so.
// The reference headers are included where they lie; nothing of them is copied into this repo.
// It lets tests/ and bench.py run the unmodified reference operators (op_absmax, op_inv_divide,
// op_multiply<float,int8_t>, op_mm<int8_t,int>, op_mm<float,float>, op_dequantize,
// op_multiply(const), op_quantized_mm; src/ops/*.cuh) on a B200, on the same inputs as our
// kernels, to pin the CPU oracle and to time "the reference's own GPU kernel".
#include <sys/time.h>

#include <cstdint>
#include <cstring>

#include "utils/tensor.cuh"
#include "ops/op_elemwise.cuh"
#undef N  // src/ops/op_elemwise.cuh:10 defines N as 256
#include "ops/op_reduction.cuh"
#include "ops/op_mm.cuh"
#include "ops/op_softmax.cuh"
#include "ops/op_layernorm.cuh"

unsigned long long randgen_seed = 0;  // the ops `extern` this (src/ops/op_elemwise.cuh:12)

namespace {

template <typename T>
Tensor<T> view(T *dptr, int h, int w) {  // non-owning [h,w] row-major view of device memory
  Tensor<T> t;
  t.h = h; t.w = w; t.stride_h = w; t.stride_w = 1; t.offset = 0;
  t.rawp = dptr; t.on_device = true;
  return t;
}

template <typename T>
void d2h(void *dst, const Tensor<T> &t) {
  if (dst) cudaMemcpy(dst, t.rawp, sizeof(T) * (size_t)t.h * t.w, cudaMemcpyDeviceToHost);
}

int status() {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  return (int)e;
}

// The op sequence of op_quantized_mm / timing_quantize.cu:38-58 with every intermediate kept.
struct Pipeline {
  Tensor<float> Cx, Cw, sx, sw, Outer;
  Tensor<int8_t> Xq, Wq;
  Tensor<int> acc;
  Pipeline(int m, int n, int k)
      : Cx{m, 1, true}, Cw{1, n, true}, sx{m, 1, true}, sw{1, n, true}, Outer{m, n, true},
        Xq{m, k, true}, Wq{k, n, true}, acc{m, n, true} {}
  void run(const Tensor<float> &X, const Tensor<float> &W, Tensor<float> &O, float range) {
    op_absmax(X, Cx);
    op_absmax(W, Cw);
    op_inv_divide(Cx, range, sx);
    op_inv_divide(Cw, range, sw);
    op_multiply(X, sx, Xq);
    op_multiply(W, sw, Wq);
    op_mm(Xq, Wq, acc);
    op_mm(Cx, Cw, Outer);
    op_dequantize(acc, Outer, O);
    op_multiply(O, 1 / (range * range), O);
  }
};

}  // namespace

extern "C" {

int ref_version() { return 1; }

// Host in / host out; returns every intermediate that is non-NULL.
int ref_quantized_mm_parts(const float *X, const float *W, int m, int n, int k, float range,
                           float *O, float *Cx, float *Cw, int8_t *Xq, int8_t *Wq, int32_t *acc) {
  Tensor<float> dX{m, k, true}, dW{k, n, true}, dO{m, n, true};
  cudaMemcpy(dX.rawp, X, sizeof(float) * (size_t)m * k, cudaMemcpyHostToDevice);
  cudaMemcpy(dW.rawp, W, sizeof(float) * (size_t)k * n, cudaMemcpyHostToDevice);
  Pipeline p(m, n, k);
  p.run(dX, dW, dO, range);
  int rc = status();
  d2h(O, dO); d2h(Cx, p.Cx); d2h(Cw, p.Cw); d2h(Xq, p.Xq); d2h(Wq, p.Wq); d2h(acc, p.acc);
  return rc;
}

// The reference's own entry point, src/ops/op_mm.cuh:67-101, host in / host out.
int ref_op_quantized_mm(const float *X, const float *W, int m, int n, int k, float range, float *O) {
  Tensor<float> dX{m, k, true}, dW{k, n, true}, dO{m, n, true};
  cudaMemcpy(dX.rawp, X, sizeof(float) * (size_t)m * k, cudaMemcpyHostToDevice);
  cudaMemcpy(dW.rawp, W, sizeof(float) * (size_t)k * n, cudaMemcpyHostToDevice);
  op_quantized_mm(dX, dW, dO, range);
  int rc = status();
  d2h(O, dO);
  return rc;
}

// op_mm<float,float>, host in / host out (the unquantized comparison product).
int ref_mm_f32(const float *A, const float *B, int m, int n, int k, float *C) {
  Tensor<float> dA{m, k, true}, dB{k, n, true}, dC{m, n, true};
  cudaMemcpy(dA.rawp, A, sizeof(float) * (size_t)m * k, cudaMemcpyHostToDevice);
  cudaMemcpy(dB.rawp, B, sizeof(float) * (size_t)k * n, cudaMemcpyHostToDevice);
  op_mm(dA, dB, dC);
  int rc = status();
  d2h(C, dC);
  return rc;
}

// op_mm<int8_t,int>, host in / host out.
int ref_mm_s8(const int8_t *A, const int8_t *B, int m, int n, int k, int32_t *C) {
  Tensor<int8_t> dA{m, k, true}, dB{k, n, true};
  Tensor<int> dC{m, n, true};
  cudaMemcpy(dA.rawp, A, (size_t)m * k, cudaMemcpyHostToDevice);
  cudaMemcpy(dB.rawp, B, (size_t)k * n, cudaMemcpyHostToDevice);
  op_mm(dA, dB, dC);
  int rc = status();
  d2h(C, dC);
  return rc;
}

// Inputs the way the timing driver draws them (src/timing_quantize.cu:17-20): cuRAND XORWOW,
// U(-1,1], X first then W from one generator seeded with `seed`.  Host out.
int ref_uniform_inputs(unsigned long long seed, int m, int n, int k, float *X, float *W) {
  curandGenerator_t gen;
  if (curandCreateGenerator(&gen, CURAND_RNG_PSEUDO_DEFAULT) != CURAND_STATUS_SUCCESS) return -1;
  curandSetPseudoRandomGeneratorSeed(gen, seed);
  Tensor<float> dX{m, k, true}, dW{k, n, true};
  curandGenerateUniform(gen, dX.rawp, (size_t)m * k);
  op_add<float>(dX, -1.0f / 2.0f, dX);      // op_uniform_init: op_add(t, min/(max-min), t)
  op_multiply(dX, 2.0f, dX);                //                  op_multiply(t, max-min, t)
  curandGenerateUniform(gen, dW.rawp, (size_t)k * n);
  op_add<float>(dW, -1.0f / 2.0f, dW);
  op_multiply(dW, 2.0f, dW);
  int rc = status();
  d2h(X, dX); d2h(W, dW);
  curandDestroyGenerator(gen);
  return rc;
}

// Device-resident timing of the reference kernels on caller-owned device buffers.
//   ms_events : CUDA-event time per iteration of the 10-kernel pipeline with every temporary
//               pre-allocated (kernels only).
//   ms_wall   : reference-style per-call time of op_quantized_mm itself -- gettimeofday around the
//               call + cudaDeviceSynchronize, temporaries cudaMalloc'ed/cudaFree'd inside
//               (src/timing_quantize.cu:38-65).
int ref_time_quantized_mm_dev(float *dXp, float *dWp, float *dOp, int m, int n, int k, float range,
                              int warmup, int iters, double *ms_events, double *ms_wall) {
  Tensor<float> X = view(dXp, m, k), W = view(dWp, k, n), O = view(dOp, m, n);
  {
    Pipeline p(m, n, k);
    for (int i = 0; i < warmup; i++) p.run(X, W, O, range);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    for (int i = 0; i < iters; i++) p.run(X, W, O, range);
    cudaEventRecord(e1, 0);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms_events) *ms_events = (double)ms / iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (ms_wall) {
    for (int i = 0; i < warmup; i++) op_quantized_mm(X, W, O, range);
    cudaDeviceSynchronize();
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    for (int i = 0; i < iters; i++) {
      op_quantized_mm(X, W, O, range);
      cudaDeviceSynchronize();
    }
    gettimeofday(&t1, NULL);
    *ms_wall = ((t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_usec - t0.tv_usec)) / 1e3 / iters;
  }
  return status();
}

// Same for the unquantized fp32 op_mm (src/timing_quantize.cu:27-35).
int ref_time_mm_f32_dev(float *dXp, float *dWp, float *dOp, int m, int n, int k, int warmup,
                        int iters, double *ms_events) {
  Tensor<float> X = view(dXp, m, k), W = view(dWp, k, n), O = view(dOp, m, n);
  for (int i = 0; i < warmup; i++) op_mm(X, W, O);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < iters; i++) op_mm(X, W, O);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_events) *ms_events = (double)ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return status();
}

// Host-buffer end-to-end call of the reference op (H2D, op_quantized_mm, D2H), wall clock.
int ref_time_quantized_mm_host(const float *X, const float *W, float *O, int m, int n, int k,
                               float range, int warmup, int iters, double *ms_wall) {
  for (int i = 0; i < warmup; i++) ref_op_quantized_mm(X, W, m, n, k, range, O);
  struct timeval t0, t1;
  gettimeofday(&t0, NULL);
  int rc = 0;
  for (int i = 0; i < iters; i++) rc |= ref_op_quantized_mm(X, W, m, n, k, range, O);
  gettimeofday(&t1, NULL);
  *ms_wall = ((t1.tv_sec - t0.tv_sec) * 1e6 + (t1.tv_usec - t0.tv_usec)) / 1e3 / iters;
  return rc;
}

// op_softmax (src/ops/op_softmax.cuh:31-41).  Its grid covers rows 0 .. 256*ceil(w/256)-1 only
// (op_softmax.cuh:38 sizes it from the width): callers keep h inside that.
int ref_softmax(const float *A, int h, int w, float *B) {
  Tensor<float> a{h, w, true}, b{h, w, true};
  cudaMemcpy(a.rawp, A, sizeof(float) * (size_t)h * w, cudaMemcpyHostToDevice);
  op_softmax(a, b);
  int rc = status();
  d2h(B, b);
  return rc;
}

// AttentionLayer<float>::forward (src/modules/attention.cuh:47-70) statement by statement, with the
// three projections (:54-56) calling the reference's own op_quantized_mm instead of op_mm -- the
// re-pointing of SURVEY.md F2 -- and queries / keys+values taken from two inputs as
// src/transformer.cu:37 expects (pass the same X twice for the 2-argument form).  Every
// intermediate can be read back.
int ref_attention_quantized(const float *Xq, const float *Xkv, const float *Wq, const float *Wk, const float *Wv,
                            int sq, int skv, int d_model, int d_k, int d_v, float range, float *Q_out, float *K_out,
                            float *V_out, float *S_out, float *P_out, float *out) {
  Tensor<float> xq{sq, d_model, true}, xkv{skv, d_model, true}, wq{d_model, d_k, true}, wk{d_model, d_k, true},
      wv{d_model, d_v, true};
  cudaMemcpy(xq.rawp, Xq, sizeof(float) * (size_t)sq * d_model, cudaMemcpyHostToDevice);
  cudaMemcpy(xkv.rawp, Xkv, sizeof(float) * (size_t)skv * d_model, cudaMemcpyHostToDevice);
  cudaMemcpy(wq.rawp, Wq, sizeof(float) * (size_t)d_model * d_k, cudaMemcpyHostToDevice);
  cudaMemcpy(wk.rawp, Wk, sizeof(float) * (size_t)d_model * d_k, cudaMemcpyHostToDevice);
  cudaMemcpy(wv.rawp, Wv, sizeof(float) * (size_t)d_model * d_v, cudaMemcpyHostToDevice);
  Tensor<float> Q(sq, d_k, true), K(skv, d_k, true), V(skv, d_v, true);
  op_quantized_mm(xq, wq, Q, range);
  op_quantized_mm(xkv, wk, K, range);
  op_quantized_mm(xkv, wv, V, range);
  Tensor<float> K_transpose = K.transpose();
  Tensor<float> QK_T(Q.h, K_transpose.w, true);
  op_mm(Q, K_transpose, QK_T);
  Tensor<float> scaled_QK_T(QK_T.h, QK_T.w, true);
  float scale_factor = 1.0 / std::sqrt(d_k);
  op_multiply(QK_T, scale_factor, scaled_QK_T);
  Tensor<float> softmax_QK_T(QK_T.h, QK_T.w, true);
  op_softmax(scaled_QK_T, softmax_QK_T);
  Tensor<float> output(sq, d_v, true);
  op_mm(softmax_QK_T, V, output);
  int rc = status();
  d2h(Q_out, Q); d2h(K_out, K); d2h(V_out, V); d2h(S_out, QK_T); d2h(P_out, softmax_QK_T); d2h(out, output);
  return rc;
}

// op_add(a, r, a); op_layernorm(a, a) -- the in-place ADD & NORM of src/transformer.cu:57-58.  The
// layernorm grid covers rows 0 .. 256*ceil(w/256)-1 only (op_layernorm.cuh:41); r == NULL skips the add.
int ref_add_layernorm(const float *A, const float *R, int h, int w, float *B) {
  Tensor<float> a{h, w, true}, r{h, w, true};
  cudaMemcpy(a.rawp, A, sizeof(float) * (size_t)h * w, cudaMemcpyHostToDevice);
  if (R) {
    cudaMemcpy(r.rawp, R, sizeof(float) * (size_t)h * w, cudaMemcpyHostToDevice);
    op_add(a, r, a);
  }
  op_layernorm(a, a);
  int rc = status();
  d2h(B, a);
  return rc;
}

// One iteration of the Encoder loop, src/transformer.cu:24-76, statement by statement on the reference's
// own kernels, with: weights supplied by the caller instead of re-drawn (op_uniform_init) every call;
// every op_mm on weights routed to op_quantized_mm; the per-head result copied into
// its column block on the device instead of through the host (:43-50, same values); ffnOut sized
// [h, d_ff] (the reference's [h, d_model] buffer only works for d_ff == d_model).
// Wq/Wk/Wv: heads consecutive [d_model, d] matrices each.
int ref_encoder_block(const float *X, int h, int d_model, int heads, int d_ff, const float *Wq, const float *Wk,
                      const float *Wv, const float *W_O, const float *W1, const float *b1, const float *W2,
                      const float *b2, float range, float *out) {
  const int d = d_model / heads;
  auto up = [](const float *src, int r, int c) {
    Tensor<float> t{r, c, true};
    cudaMemcpy(t.rawp, src, sizeof(float) * (size_t)r * c, cudaMemcpyHostToDevice);
    return t;
  };
  Tensor<float> x = up(X, h, d_model), wo = up(W_O, d_model, d_model), w1 = up(W1, d_model, d_ff), bb1 = up(b1, 1, d_ff),
                w2 = up(W2, d_ff, d_model), bb2 = up(b2, 1, d_model);
  Tensor<float> multiHeadOut{h, d_model, true}, output{h, d_model, true};
  for (int j = 0; j < heads; j++) {
    Tensor<float> wq = up(Wq + (size_t)j * d_model * d, d_model, d), wk = up(Wk + (size_t)j * d_model * d, d_model, d),
                  wv = up(Wv + (size_t)j * d_model * d, d_model, d);
    Tensor<float> Q(h, d, true), K(h, d, true), V(h, d, true);
    op_quantized_mm(x, wq, Q, range);
    op_quantized_mm(x, wk, K, range);
    op_quantized_mm(x, wv, V, range);
    Tensor<float> K_transpose = K.transpose();
    Tensor<float> QK_T(h, h, true), scaled(h, h, true), P(h, h, true), attnHeadOut(h, d, true);
    op_mm(Q, K_transpose, QK_T);
    float scale_factor = 1.0 / std::sqrt(d);
    op_multiply(QK_T, scale_factor, scaled);
    op_softmax(scaled, P);
    op_mm(P, V, attnHeadOut);
    cudaMemcpy2D(multiHeadOut.rawp + j * d, sizeof(float) * d_model, attnHeadOut.rawp, sizeof(float) * d, sizeof(float) * d, h,
                 cudaMemcpyDeviceToDevice);
  }
  op_quantized_mm(multiHeadOut, wo, output, range);
  op_add(output, multiHeadOut, output);
  op_layernorm(output, output);
  Tensor<float> ffnOut{h, d_ff, true};
  op_quantized_mm(output, w1, ffnOut, range);  // LinearLayer::forward: op_mm + op_add(bias), linear.cuh:53-54
  op_add(ffnOut, bb1, ffnOut);
  op_relu(ffnOut, ffnOut);
  op_quantized_mm(ffnOut, w2, output, range);
  op_add(output, bb2, output);
  op_add(output, multiHeadOut, output);
  op_layernorm(output, output);
  int rc = status();
  d2h(out, output);
  return rc;
}

// One iteration of the Decoder loop, src/transformer.cu:91-166, statement by statement on the reference's own
// kernels, under the same conventions as ref_encoder_block (weights supplied, every product on weights through
// op_quantized_mm, heads concatenated on the device, ffnOut sized [h, d_ff]).  The 3-argument
// attn.forward(Xq, Xkv, out) the reference calls (:104,106,132) does not exist in attention.cuh;
// it is spelt here as the op sequence of AttentionLayer::forward (attention.cuh:51-70) with the queries
// projected from the first argument and the keys / values from the second.
// X [h, d_model] decoder input, E [h_enc, d_model] encoder output.  Wq1/Wk1/Wv1, Wq2/Wk2/Wv2: heads consecutive
// [d_model, d] matrices (self- and cross-attention); W_O1, W_O2 [d_model, d_model].
int ref_decoder_block(const float *X, const float *E, int h, int h_enc, int d_model, int heads, int d_ff,
                      const float *Wq1, const float *Wk1, const float *Wv1, const float *W_O1, const float *Wq2,
                      const float *Wk2, const float *Wv2, const float *W_O2, const float *W1, const float *b1,
                      const float *W2, const float *b2, float range, float *out) {
  const int d = d_model / heads;
  auto up = [](const float *src, int r, int c) {
    Tensor<float> t{r, c, true};
    cudaMemcpy(t.rawp, src, sizeof(float) * (size_t)r * c, cudaMemcpyHostToDevice);
    return t;
  };
  Tensor<float> x = up(X, h, d_model), enc_output = up(E, h_enc, d_model), wo1 = up(W_O1, d_model, d_model),
                wo2 = up(W_O2, d_model, d_model), w1 = up(W1, d_model, d_ff), bb1 = up(b1, 1, d_ff),
                w2 = up(W2, d_ff, d_model), bb2 = up(b2, 1, d_model);
  Tensor<float> multiHeadOut{h, d_model, true}, output{h, d_model, true};
  // attention.cuh:51-70 with separate query and key/value inputs; result into column block j of multiHeadOut
  auto head = [&](const Tensor<float> &xq, const Tensor<float> &xkv, const float *Wq, const float *Wk, const float *Wv, int j) {
    Tensor<float> wq = up(Wq + (size_t)j * d_model * d, d_model, d), wk = up(Wk + (size_t)j * d_model * d, d_model, d),
                  wv = up(Wv + (size_t)j * d_model * d, d_model, d);
    Tensor<float> Q(xq.h, d, true), K(xkv.h, d, true), V(xkv.h, d, true);
    op_quantized_mm(xq, wq, Q, range);
    op_quantized_mm(xkv, wk, K, range);
    op_quantized_mm(xkv, wv, V, range);
    Tensor<float> K_transpose = K.transpose();
    Tensor<float> QK_T(Q.h, K_transpose.w, true), scaled(Q.h, K_transpose.w, true), P(Q.h, K_transpose.w, true),
        attnHeadOut(Q.h, d, true);
    op_mm(Q, K_transpose, QK_T);
    float scale_factor = 1.0 / std::sqrt(d);
    op_multiply(QK_T, scale_factor, scaled);
    op_softmax(scaled, P);
    op_mm(P, V, attnHeadOut);
    cudaMemcpy2D(multiHeadOut.rawp + j * d, sizeof(float) * d_model, attnHeadOut.rawp, sizeof(float) * d, sizeof(float) * d,
                 Q.h, cudaMemcpyDeviceToDevice);
  };
  for (int j = 0; j < heads; j++) head(x, x, Wq1, Wk1, Wv1, j);           // :97-116 (the reference applies no causal mask)
  op_quantized_mm(multiHeadOut, wo1, output, range);                         // :117-119
  op_add(output, multiHeadOut, output);                                      // :123
  op_layernorm(output, output);                                              // :124
  for (int j = 0; j < heads; j++) head(output, enc_output, Wq2, Wk2, Wv2, j);  // :127-140
  op_quantized_mm(multiHeadOut, wo2, output, range);                         // :143-144
  op_add(output, multiHeadOut, output);                                      // :148
  op_layernorm(output, output);                                              // :149
  Tensor<float> ffnOut{h, d_ff, true};
  op_quantized_mm(output, w1, ffnOut, range);                                // :154-156 (LinearLayer::forward)
  op_add(ffnOut, bb1, ffnOut);
  op_relu(ffnOut, ffnOut);                                                   // :157
  op_quantized_mm(ffnOut, w2, output, range);                                // :159-161
  op_add(output, bb2, output);
  op_add(output, multiHeadOut, output);                                      // :165
  op_layernorm(output, output);                                              // :166
  int rc = status();
  d2h(out, output);
  return rc;
}

}  // extern "C"
