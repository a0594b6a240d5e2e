// qg_dropin.cuh -- reference-shaped C++ operator layer over the C ABI (include/qgemm.h).
//
// The reference's hot path is a set of header-only function templates over its Tensor<T> view
// (/root/reference/src/ops/op_mm.cuh, op_reduction.cuh, op_elemwise.cuh).  This header offers the
// same functions -- same names, argument order and error behaviour (assert) -- in namespace
// qg_dropin, written against ANY tensor type that has the reference's public fields
//     h, w, stride_h, stride_w, offset, rawp, on_device        (src/utils/tensor.cuh:47-54)
// so it compiles against the reference's own Tensor<T> unchanged (INTEGRATION.md shows the
// three-line edit of op_quantized_mm that re-points the reference at it) as well as against the
// minimal qg_dropin::Tensor<T> below, which tests/cpp uses on machines without the reference.
//
// Host code only needs a C++14 compiler and libqgemm.so; no CUDA headers are required except for
// Tensor's own allocation helpers.
#pragma once

#include <assert.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <math.h>
#include <string.h>

#include <memory>
#include <vector>
#include <type_traits>

#include "../../include/qgemm.h"

namespace qg_dropin {

// ------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------
template <typename T> struct io_dtype;
template <> struct io_dtype<float> { static constexpr int value = QG_F32; };

inline void check(int rc, const char *what) {
  if (rc != QG_OK) {
    fprintf(stderr, "%s failed (%d): %s\n", what, rc, qg_last_error());
    assert(0 && "qgemm call failed");  // the reference's error convention: print, then assert(0)
  }
}

template <class TensorT>
auto *base_ptr(const TensorT &t) { return t.rawp + t.offset; }

// the fast C-ABI entry points take row-major matrices with unit inner stride
template <class TensorT>
bool unit_inner(const TensorT &t) { return t.stride_w == 1 || t.w == 1; }

// ------------------------------------------------------------------------------------------
// ops on the quantized path
// ------------------------------------------------------------------------------------------

// op_absmax(in, out): src/ops/op_reduction.cuh:195-204; direction from the output shape (:143)
template <class TensorT>
void op_absmax(const TensorT &in, TensorT &out, int mode = QG_MODE_REF_EXACT) {
  assert((out.h == 1 && in.w == out.w) || (out.w == 1 && in.h == out.h));
  assert(in.on_device && out.on_device);
  assert(unit_inner(in));
  if (in.h > out.h) {
    assert(out.stride_w == 1 || out.w == 1);
    check(qg_absmax_cols(base_ptr(in), QG_F32, in.h, in.w, in.stride_h, mode, base_ptr(out), nullptr), "qg_absmax_cols");
  } else {
    assert(out.stride_h == 1 || out.h == 1);
    check(qg_absmax_rows(base_ptr(in), QG_F32, in.h, in.w, in.stride_h, mode, base_ptr(out), nullptr), "qg_absmax_rows");
  }
}

// op_inv_divide(a, b, out) = b / a: src/ops/op_elemwise.cuh:657-667 (contiguous vectors)
template <class TensorT, typename T>
void op_inv_divide(const TensorT &a, T b, TensorT &out) {
  assert(out.h == a.h && out.w == a.w);
  assert(a.on_device && out.on_device);
  check(qg_inv_divide_f32(base_ptr(a), (int64_t)a.h * a.w, (float)b, base_ptr(out), nullptr), "qg_inv_divide_f32");
}

// rows of a [h,w] view as a leading dimension the C ABI accepts (a size-1 dimension may carry any stride)
template <class TensorT>
int64_t ld_of(const TensorT &t) { return t.h > 1 ? t.stride_h : (t.stride_h >= t.w ? t.stride_h : t.w); }

// op_multiply<T,OutT>(a, b, out): src/ops/op_elemwise.cuh:629-640.  OutT == int8_t is the quantizing cast with a
// broadcast scale vector (a4 of the hot path); OutT == T is the plain elementwise product (:598-610).
template <class TensorA, class TensorQ>
void op_multiply(const TensorA &a, const TensorA &b, TensorQ &out) {
  assert(out.h == a.h && out.w == a.w);
  assert((a.h == b.h && a.w == b.w) || (a.h == b.h && b.w == 1) || (a.w == b.w && b.h == 1));
  assert(a.on_device && b.on_device && out.on_device);
  assert(unit_inner(a) && unit_inner(out));
  if (sizeof(*out.rawp) == 1) {
    assert((a.h == b.h && b.w == 1) || (a.w == b.w && b.h == 1));
    if (b.w == 1 && a.h == b.h && a.w != b.w)
      check(qg_quantize_rows(base_ptr(a), QG_F32, a.h, a.w, ld_of(a), (const float *)base_ptr(b), (int8_t *)base_ptr(out),
                             ld_of(out), nullptr), "qg_quantize_rows");
    else
      check(qg_quantize_cols(base_ptr(a), QG_F32, a.h, a.w, ld_of(a), (const float *)base_ptr(b), (int8_t *)base_ptr(out),
                             ld_of(out), nullptr), "qg_quantize_cols");
  } else {
    check(qg_multiply_f32((const float *)base_ptr(a), ld_of(a), (const float *)base_ptr(b), ld_of(b), b.h, b.w,
                          (float *)base_ptr(out), ld_of(out), a.h, a.w, nullptr), "qg_multiply_f32");
  }
}

// op_multiply(a, T b, out): src/ops/op_elemwise.cuh:644-654 (the 1/range^2 step of op_mm.cuh:99)
template <class TensorT, typename T, typename = typename std::enable_if<std::is_arithmetic<T>::value>::type>
void op_multiply(const TensorT &a, T b, TensorT &out) {
  assert(out.h == a.h && out.w == a.w);
  assert(a.on_device && out.on_device && unit_inner(a) && unit_inner(out));
  check(qg_multiply_const_f32(base_ptr(a), ld_of(a), (float)b, base_ptr(out), ld_of(out), a.h, a.w, nullptr),
        "qg_multiply_const_f32");
}

// op_dequantize(a, b, out): src/ops/op_elemwise.cuh:614-625 -- out = (float)a * b over a materialised b
template <class TensorI, class TensorF>
void op_dequantize(const TensorI &a, const TensorF &b, TensorF &out) {
  assert(out.h == a.h && out.w == a.w);
  assert(a.h == b.h && a.w == b.w);  // the pipeline's call site (op_mm.cuh:98) passes the full outer product
  assert(a.on_device && b.on_device && out.on_device && unit_inner(a) && unit_inner(b) && unit_inner(out));
  check(qg_dequantize_outer_f32((const int32_t *)base_ptr(a), ld_of(a), base_ptr(b), ld_of(b), base_ptr(out), ld_of(out), a.h,
                                a.w, nullptr), "qg_dequantize_outer_f32");
}

// op_add(a, b, out) / op_subtract(a, b, out): src/ops/op_elemwise.cuh:501-512, 531-542 (broadcast rule :410-421)
template <class TensorT>
void op_add(const TensorT &a, const TensorT &b, TensorT &out) {
  assert(out.h == a.h && out.w == a.w);
  assert((a.h == b.h && a.w == b.w) || (a.h == b.h && b.w == 1) || (a.w == b.w && b.h == 1));
  assert(a.on_device && b.on_device && out.on_device && unit_inner(a) && unit_inner(b) && unit_inner(out));
  check(qg_add_f32(base_ptr(a), ld_of(a), base_ptr(b), ld_of(b), b.h, b.w, base_ptr(out), ld_of(out), a.h, a.w, nullptr),
        "qg_add_f32");
}
template <class TensorT>
void op_subtract(const TensorT &a, const TensorT &b, TensorT &out) {
  assert(out.h == a.h && out.w == a.w);
  assert((a.h == b.h && a.w == b.w) || (a.h == b.h && b.w == 1) || (a.w == b.w && b.h == 1));
  assert(a.on_device && b.on_device && out.on_device && unit_inner(a) && unit_inner(b) && unit_inner(out));
  check(qg_subtract_f32(base_ptr(a), ld_of(a), base_ptr(b), ld_of(b), b.h, b.w, base_ptr(out), ld_of(out), a.h, a.w, nullptr),
        "qg_subtract_f32");
}

// op_relu(t, out): src/ops/op_elemwise.cuh:454-465
template <class TensorT>
void op_relu(const TensorT &t, TensorT &out) {
  assert(out.h == t.h && out.w == t.w);
  assert(t.on_device && out.on_device && unit_inner(t) && unit_inner(out));
  check(qg_relu_f32(base_ptr(t), ld_of(t), base_ptr(out), ld_of(out), t.h, t.w, nullptr), "qg_relu_f32");
}

// op_layernorm(A, B): src/ops/op_layernorm.cuh:34-44 -- (x - mean) / var with ascending-order sums; every row
template <class TensorT>
void op_layernorm(const TensorT &A, TensorT &B) {
  assert(A.h == B.h && A.w == B.w);
  assert(A.on_device && B.on_device && unit_inner(A) && unit_inner(B));
  check(qg_add_layernorm_f32(base_ptr(A), ld_of(A), nullptr, 0, A.h, A.w, base_ptr(B), ld_of(B), nullptr), "qg_add_layernorm_f32");
}

// op_mm<T,OutT>(A, B, C): src/ops/op_mm.cuh:49-65
template <class TensorA, class TensorC>
void op_mm(const TensorA &A, const TensorA &B, TensorC &C) {
  assert(A.h == C.h && B.w == C.w && A.w == B.h);
  assert(A.on_device && B.on_device && C.on_device);
  using AT = typename std::remove_cv<typename std::remove_reference<decltype(*A.rawp)>::type>::type;
  if (std::is_same<AT, int8_t>::value) {
    assert(unit_inner(A) && unit_inner(B) && unit_inner(C));
    check(qg_gemm_s8s8s32((const int8_t *)base_ptr(A), A.stride_h, (const int8_t *)base_ptr(B), B.stride_h, A.h, B.w, A.w,
                          (int32_t *)base_ptr(C), C.stride_h, nullptr), "qg_gemm_s8s8s32");
  } else {
    assert(unit_inner(C));
    check(qg_mm_f32((const float *)base_ptr(A), A.stride_h, A.stride_w, (const float *)base_ptr(B), B.stride_h, B.stride_w,
                    A.h, B.w, A.w, (float *)base_ptr(C), C.stride_h, nullptr), "qg_mm_f32");
  }
}

// op_quantized_mm(X, W, O, range): src/ops/op_mm.cuh:67-101 -- the whole pipeline in one call
template <class TensorT, typename T>
void op_quantized_mm(const TensorT &X, const TensorT &W, TensorT &O, T range, int mode = QG_MODE_REF_EXACT) {
  assert(X.h == O.h && W.w == O.w && X.w == W.h);
  assert(X.on_device && W.on_device && O.on_device);
  assert(unit_inner(X) && unit_inner(W) && unit_inner(O));
  check(qg_quantized_mm(base_ptr(X), X.stride_h, base_ptr(W), W.stride_h, QG_F32, base_ptr(O), O.stride_h, QG_F32, X.h, W.w,
                        X.w, (float)range, mode, nullptr, nullptr, 0, nullptr), "qg_quantized_mm");
}

// LinearLayer<T>::forward(x, y) = x @ w + b with the product quantized: src/modules/linear.cuh:49-56
template <class TensorT>
void linear_forward(const TensorT &x, const TensorT &w, const TensorT &b, TensorT &y, float range = 127.0f,
                    int mode = QG_MODE_REF_EXACT) {
  assert(x.w == w.h && y.h == x.h && y.w == w.w && b.h == 1 && b.w == w.w);
  assert(x.on_device && w.on_device && b.on_device && y.on_device);
  check(qg_quantized_mm(base_ptr(x), x.stride_h, base_ptr(w), w.stride_h, QG_F32, base_ptr(y), y.stride_h, QG_F32, x.h, w.w,
                        x.w, range, mode, base_ptr(b), nullptr, 0, nullptr), "qg_quantized_mm(+bias)");
}

// op_softmax(A, B): src/ops/op_softmax.cuh:31-41 (contiguous columns, as every reference call site has)
template <class TensorT>
void op_softmax(const TensorT &A, TensorT &B) {
  assert(A.h == B.h && A.w == B.w);
  assert(A.on_device && B.on_device);
  assert(A.stride_w == 1 && B.stride_w == 1);
  check(qg_softmax_rows_f32(base_ptr(A), A.stride_h, A.h, A.w, 1.0f, base_ptr(B), B.stride_h, nullptr), "qg_softmax_rows_f32");
}

// AttentionLayer<T>::forward(X, output), src/modules/attention.cuh:47-70, and the (Xq, Xkv, output) form
// src/transformer.cu:37,132 calls: the three projections on the quantized path, the rest in the
// reference's fp32 arithmetic.  W_q / W_k / W_v are separate tensors in the reference; they are packed
// side by side into one scratch matrix (3 strided device copies) so that the projections run as one
// product.  A layer that keeps its weights packed calls qg_attention_forward directly.
template <class TensorT>
void attention_forward(const TensorT &Xq, const TensorT &Xkv, const TensorT &W_q, const TensorT &W_k, const TensorT &W_v,
                       TensorT &output, float range = 127.0f, int mode = QG_MODE_REF_EXACT) {
  const int d_model = Xq.w, d_k = W_q.w, d_v = W_v.w, ntot = 2 * d_k + d_v;
  assert(Xkv.w == d_model && W_q.h == d_model && W_k.h == d_model && W_v.h == d_model && W_k.w == d_k);
  assert(output.h == Xq.h && output.w == d_v);
  assert(Xq.on_device && Xkv.on_device && W_q.on_device && W_k.on_device && W_v.on_device && output.on_device);
  assert(Xq.stride_w == 1 && Xkv.stride_w == 1 && W_q.stride_w == 1 && W_k.stride_w == 1 && W_v.stride_w == 1);
  float *packed = nullptr;
  cudaError_t e = cudaMalloc(&packed, sizeof(float) * (size_t)d_model * ntot);
  assert(e == cudaSuccess);
  const TensorT *ws[3] = {&W_q, &W_k, &W_v};
  int col = 0;
  for (int i = 0; i < 3; i++) {
    e = cudaMemcpy2DAsync(packed + col, sizeof(float) * ntot, base_ptr(*ws[i]), sizeof(float) * ws[i]->stride_h,
                          sizeof(float) * ws[i]->w, d_model, cudaMemcpyDeviceToDevice, nullptr);
    assert(e == cudaSuccess);
    col += ws[i]->w;
  }
  (void)e;
  const bool self = base_ptr(Xq) == base_ptr(Xkv) && Xq.h == Xkv.h;
  check(qg_attention_forward(base_ptr(Xq), Xq.stride_h, self ? base_ptr(Xq) : base_ptr(Xkv), Xkv.stride_h, 1, Xq.h, Xkv.h,
                             d_model, packed, ntot, 1, d_k, d_v, range, mode, base_ptr(output), output.stride_h, nullptr),
        "qg_attention_forward");
  cudaFree(packed);  // synchronises with the legacy stream the work above ran on
}
template <class TensorT>
void attention_forward(const TensorT &X, const TensorT &W_q, const TensorT &W_k, const TensorT &W_v, TensorT &output) {
  attention_forward(X, X, W_q, W_k, W_v, output);
}

// op_outlier_extractor(a, b, out): src/ops/op_elemwise.cuh:698-708
template <class TensorT, typename T>
void op_outlier_extractor(const TensorT &a, T b, TensorT &out) {
  assert(out.h == a.h && out.w == a.w);
  assert(a.on_device && out.on_device);
  check(qg_outlier_mask_f32(base_ptr(a), a.h, a.w, a.stride_h, (float)b, base_ptr(out), out.stride_h, nullptr),
        "qg_outlier_mask_f32");
}

// ------------------------------------------------------------------------------------------
// A minimal tensor with the reference's field names, for builds without the reference tree.
// ------------------------------------------------------------------------------------------
template <typename T>
class Tensor {
 public:
  int32_t h = 0, w = 0, stride_h = 0, stride_w = 0, offset = 0;
  T *rawp = nullptr;
  std::shared_ptr<T> ref;
  bool on_device = false;

  Tensor() = default;
  Tensor(int32_t h_, int32_t w_, bool on_device_ = false)
      : h(h_), w(w_), stride_h(w_), stride_w(1), offset(0), on_device(on_device_) {
    const size_t bytes = sizeof(T) * (size_t)h * w;
    if (on_device) {
      void *p = nullptr;
      cudaError_t e = cudaMalloc(&p, bytes);
      assert(e == cudaSuccess);
      (void)e;
      rawp = static_cast<T *>(p);
      ref = std::shared_ptr<T>(rawp, [](T *q) { cudaFree(q); });
    } else {
      rawp = static_cast<T *>(malloc(bytes));
      ref = std::shared_ptr<T>(rawp, [](T *q) { free(q); });
    }
  }
  T &at(int r, int c) { return rawp[offset + r * stride_h + c * stride_w]; }
  const T &at(int r, int c) const { return rawp[offset + r * stride_h + c * stride_w]; }

  Tensor<T> toDevice() const {
    assert(!on_device && stride_w == 1 && stride_h == w && offset == 0);
    Tensor<T> t(h, w, true);
    cudaMemcpy(t.rawp, rawp, sizeof(T) * (size_t)h * w, cudaMemcpyHostToDevice);
    return t;
  }
  Tensor<T> toHost() const {
    assert(on_device && stride_w == 1 && stride_h == w && offset == 0);
    Tensor<T> t(h, w, false);
    cudaMemcpy(t.rawp, rawp, sizeof(T) * (size_t)h * w, cudaMemcpyDeviceToHost);
    return t;
  }
  Tensor<T> transpose() const {
    Tensor<T> t = *this;
    t.h = w; t.w = h; t.stride_h = stride_w; t.stride_w = stride_h;
    return t;
  }
  T mean() const {  // sequential fp32 sum / (h*w), the figure timing_quantize prints
    assert(!on_device);
    T s = 0;
    for (int i = 0; i < h; i++)
      for (int j = 0; j < w; j++) s += at(i, j);
    return s / (h * w);
  }
};

// non-owning view over any tensor type with the reference's public fields
template <class AnyTensor>
Tensor<float> view_of(const AnyTensor &t) {
  Tensor<float> v;
  v.h = t.h; v.w = t.w; v.stride_h = t.stride_h; v.stride_w = t.stride_w; v.offset = t.offset;
  v.rawp = const_cast<float *>(reinterpret_cast<const float *>(t.rawp));
  v.on_device = t.on_device;
  return v;
}

// ------------------------------------------------------------------------------------------
// The reference's module classes on the quantized path (inference side): same names, members and
// call signatures as src/modules/param.cuh, linear.cuh:7-72 and attention.cuh:10-70, templated on the
// tensor type so that they work with the reference's Tensor<T> as well as the one above.
//   LinearLayer:    forward(x, y) = x @ w + b with the int8 codes of w prepared on first use
//                   (qg_prepare_weights -> qg_linear_forward), optional ReLU (transformer.cu:65-67)
//   AttentionLayer: forward(X, out) and the forward(Xq, Xkv, out) of transformer.cu:37,132
// Weights are plain public members; call invalidate() after changing them.  backward(), SGD and the
// loss stay with the reference (training is out of scope).
// ------------------------------------------------------------------------------------------
template <typename T, template <typename> class TensorT = Tensor>
class Parameter {
 public:
  Parameter() = default;
  Parameter(int t_h, int t_w, bool gpu) : t(t_h, t_w, gpu), dt(t_h, t_w, gpu) {}
  TensorT<T> t, dt;
};

// U(lo, hi) from a 64-bit LCG (the reference draws with cuRAND, op_elemwise.cuh:12-47; any values do)
template <class TensorT>
void uniform_init(TensorT &t, float lo, float hi, uint64_t seed) {
  std::vector<float> h((size_t)t.h * t.w);
  uint64_t s = seed * 6364136223846793005ULL + 1442695040888963407ULL;
  for (float &v : h) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    v = lo + (hi - lo) * (float)((s >> 40) & 0xffffff) / 16777216.0f;
  }
  assert(t.stride_w == 1 && t.stride_h == t.w);
  if (t.on_device) cudaMemcpy(base_ptr(t), h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice);
  else memcpy(base_ptr(t), h.data(), sizeof(float) * h.size());
}

template <typename T, template <typename> class TensorT = Tensor>
class LinearLayer {
 public:
  int in_dim = 0, out_dim = 0;
  Parameter<T, TensorT> w, b;

  LinearLayer() = default;
  LinearLayer(int in_dim_, int out_dim_, bool gpu) : in_dim(in_dim_), out_dim(out_dim_), w(in_dim_, out_dim_, gpu), b(1, out_dim_, gpu) {
    assert(gpu);
  }
  std::vector<Parameter<T, TensorT> *> parameters() { return {&w, &b}; }
  void init_uniform(uint64_t seed = 1) {  // linear.cuh:33-39
    const float mx = 1.0f / std::sqrt((float)in_dim);
    uniform_init(w.t, -mx, mx, seed);
    uniform_init(b.t, -mx, mx, seed + 1);
    invalidate();
  }
  void invalidate() { prepared_ = false; }
  void forward(const TensorT<T> &x, TensorT<T> &y, int act = QG_ACT_NONE) {  // linear.cuh:49-56
    assert(x.w == in_dim && y.h == x.h && y.w == out_dim && x.on_device && y.on_device);
    if (!prepared_) {
      wt_ = TensorT<int8_t>(out_dim, (in_dim + 15) / 16 * 16, true);
      cw_ = TensorT<float>(1, out_dim, true);
      check(qg_prepare_weights(base_ptr(w.t), QG_F32, in_dim, out_dim, w.t.stride_h, 127.0f, QG_MODE_REF_EXACT, wt_.rawp,
                               wt_.stride_h, cw_.rawp, nullptr), "qg_prepare_weights");
      prepared_ = true;
    }
    check(qg_linear_forward_act(base_ptr(x), x.stride_h, QG_F32, wt_.rawp, wt_.stride_h, cw_.rawp, base_ptr(b.t), act,
                                base_ptr(y), y.stride_h, QG_F32, x.h, out_dim, in_dim, 127.0f, QG_MODE_REF_EXACT, nullptr, 0,
                                nullptr), "qg_linear_forward_act");
  }

 private:
  TensorT<int8_t> wt_;
  TensorT<float> cw_;
  bool prepared_ = false;
};

template <typename T, template <typename> class TensorT = Tensor>
class AttentionLayer {
 public:
  int d_model, d_k, d_v;
  Parameter<T, TensorT> W_q, W_k, W_v;

  AttentionLayer(int d_model_, int d_k_, int d_v_, bool gpu)
      : d_model(d_model_), d_k(d_k_), d_v(d_v_), W_q(d_model_, d_k_, gpu), W_k(d_model_, d_k_, gpu), W_v(d_model_, d_v_, gpu) {
    assert(gpu);
  }
  std::vector<Parameter<T, TensorT> *> parameters() { return {&W_q, &W_k, &W_v}; }
  void init_uniform(uint64_t seed = 1) {  // attention.cuh:40-45
    const float mx = 1.0f / std::sqrt((float)d_k);
    uniform_init(W_q.t, -mx, mx, seed);
    uniform_init(W_k.t, -mx, mx, seed + 1);
    uniform_init(W_v.t, -mx, mx, seed + 2);
  }
  void forward(const TensorT<T> &X, TensorT<T> &output) { attention_forward(X, X, W_q.t, W_k.t, W_v.t, output); }
  void forward(const TensorT<T> &Xq, const TensorT<T> &Xkv, TensorT<T> &output) {
    attention_forward(Xq, Xkv, W_q.t, W_k.t, W_v.t, output);
  }
};

// ------------------------------------------------------------------------------------------
// Encoder / Decoder of src/transformer.cu:14-168 on the quantized path.
//
// The reference functions re-draw every weight (op_uniform_init) inside every call, run one single-head
// AttentionLayer per head and concatenate the heads through the host.  Here a block owns its weights (drawn once,
// int8 codes prepared on first use); all heads' projections are one quantized product (qg_attention_forward with
// heads > 1, W_q | W_k | W_v of all heads packed side by side); W_O and the FFN go through qg_linear_forward_act
// (ReLU inside the GEMM epilogue); ADD & NORM is qg_add_layernorm_f32.  Every step keeps the reference's
// arithmetic, including its quirks: the residual added is the attention output (:57,74,123,148,165), W_O is drawn
// from U(-1, 1) (:53), the "layernorm" divides by the variance.  `batch` independent sequences per call (the
// reference has none: batch = 1).  The free functions Encoder(...) / Decoder(...) keep the reference's
// signatures and its "fresh random weights on every call" behaviour.
// ------------------------------------------------------------------------------------------
template <template <typename> class TensorT = Tensor>
class PreparedLinear {  // y = act(x @ w [+ b]); bias optional (W_O is a bare product, transformer.cu:52-54)
 public:
  int in_dim = 0, out_dim = 0;
  TensorT<float> w, b;
  bool has_bias = true;
  PreparedLinear() = default;
  PreparedLinear(int in_dim_, int out_dim_, bool bias) : in_dim(in_dim_), out_dim(out_dim_), w(in_dim_, out_dim_, true),
                                                         b(1, out_dim_, true), has_bias(bias) {}
  void init_uniform(uint64_t seed, float lo = 0.0f, float hi = 0.0f) {
    const float mx = 1.0f / std::sqrt((float)in_dim);  // linear.cuh:33-39
    uniform_init(w, lo == hi ? -mx : lo, lo == hi ? mx : hi, seed);
    uniform_init(b, -mx, mx, seed + 1);
    prepared_ = false;
  }
  void invalidate() { prepared_ = false; }
  void forward(const TensorT<float> &x, TensorT<float> &y, int act = QG_ACT_NONE) {
    assert(x.w == in_dim && y.h == x.h && y.w == out_dim && x.on_device && y.on_device);
    if (!prepared_) {
      wt_ = TensorT<int8_t>(out_dim, (in_dim + 15) / 16 * 16, true);
      cw_ = TensorT<float>(1, out_dim, true);
      check(qg_prepare_weights(base_ptr(w), QG_F32, in_dim, out_dim, w.stride_h, 127.0f, QG_MODE_REF_EXACT, wt_.rawp,
                               wt_.stride_h, cw_.rawp, nullptr), "qg_prepare_weights");
      prepared_ = true;
    }
    check(qg_linear_forward_act(base_ptr(x), ld_of(x), QG_F32, wt_.rawp, wt_.stride_h, cw_.rawp, has_bias ? base_ptr(b) : nullptr,
                                act, base_ptr(y), ld_of(y), QG_F32, x.h, out_dim, in_dim, 127.0f, QG_MODE_REF_EXACT, nullptr, 0,
                                nullptr), "qg_linear_forward_act");
  }

 private:
  TensorT<int8_t> wt_;
  TensorT<float> cw_;
  bool prepared_ = false;
};

// all heads of one attention sub-layer: W_qkv [d_model, heads * (2 d_k + d_v)] = [ W_q of every head | W_k ... | W_v ... ]
template <template <typename> class TensorT = Tensor>
class MultiHeadAttention {
 public:
  int d_model = 0, heads = 0, d_k = 0, d_v = 0;
  TensorT<float> W_qkv;
  MultiHeadAttention() = default;
  MultiHeadAttention(int d_model_, int heads_) : d_model(d_model_), heads(heads_), d_k(d_model_ / heads_), d_v(d_model_ / heads_),
                                                 W_qkv(d_model_, heads_ * 3 * (d_model_ / heads_), true) {}
  void init_uniform(uint64_t seed) {  // attention.cuh:40-45, every head
    const float mx = 1.0f / std::sqrt((float)d_k);
    uniform_init(W_qkv, -mx, mx, seed);
    prepared_ = false;
  }
  void invalidate() { prepared_ = false; }  // after writing W_qkv directly
  // queries from xq, keys / values from xkv (the same tensor for self-attention); out [rows, heads * d_v].
  // The projection weights are column-quantized on first use (same bits as quantizing them on every call).
  void forward(const TensorT<float> &xq, const TensorT<float> &xkv, TensorT<float> &out, int batch = 1) {
    assert(xq.w == d_model && xkv.w == d_model && out.h == xq.h && out.w == heads * d_v);
    assert(xq.h % batch == 0 && xkv.h % batch == 0);
    const int ntot = heads * (2 * d_k + d_v);
    if (!prepared_) {
      wt_ = TensorT<int8_t>(ntot, (d_model + 15) / 16 * 16, true);
      cw_ = TensorT<float>(1, ntot, true);
      check(qg_prepare_weights(base_ptr(W_qkv), QG_F32, d_model, ntot, W_qkv.stride_h, 127.0f, QG_MODE_REF_EXACT, wt_.rawp,
                               wt_.stride_h, cw_.rawp, nullptr), "qg_prepare_weights");
      prepared_ = true;
    }
    check(qg_attention_forward_prepared(base_ptr(xq), ld_of(xq), base_ptr(xkv), ld_of(xkv), batch, xq.h / batch, xkv.h / batch,
                                        d_model, wt_.rawp, wt_.stride_h, cw_.rawp, heads, d_k, d_v, 127.0f, QG_MODE_REF_EXACT,
                                        base_ptr(out), ld_of(out), nullptr), "qg_attention_forward_prepared");
  }

 private:
  TensorT<int8_t> wt_;
  TensorT<float> cw_;
  bool prepared_ = false;
};

template <class TensorF>
void add_layernorm(const TensorF &a, const TensorF &r, TensorF &out) {  // op_add(a, r, out); op_layernorm(out, out)
  check(qg_add_layernorm_f32(base_ptr(a), ld_of(a), base_ptr(r), ld_of(r), a.h, a.w, base_ptr(out), ld_of(out), nullptr),
        "qg_add_layernorm_f32");
}

template <template <typename> class TensorT = Tensor>
class EncoderBlock {  // one iteration of the loop at transformer.cu:24-76
 public:
  int d_model, heads, d_ff;
  MultiHeadAttention<TensorT> attn;
  PreparedLinear<TensorT> W_O, ll1, ll2;
  EncoderBlock(int d_model_, int heads_, int d_ff_)
      : d_model(d_model_), heads(heads_), d_ff(d_ff_), attn(d_model_, heads_), W_O(d_model_, d_model_, false),
        ll1(d_model_, d_ff_, true), ll2(d_ff_, d_model_, true) {}
  void init_uniform(uint64_t seed) {
    attn.init_uniform(seed);
    W_O.init_uniform(seed + 10, -1.0f, 1.0f);  // transformer.cu:53
    ll1.init_uniform(seed + 20);
    ll2.init_uniform(seed + 30);
  }
  void forward(const TensorT<float> &x, TensorT<float> &out, int batch = 1) {
    TensorT<float> mh(x.h, d_model, true), ffn(x.h, d_ff, true);
    attn.forward(x, x, mh, batch);       // :27-50
    W_O.forward(mh, out);                // :52-54
    add_layernorm(out, mh, out);         // :57-58
    ll1.forward(out, ffn, QG_ACT_RELU);  // :63-67
    ll2.forward(ffn, out);               // :69-71
    add_layernorm(out, mh, out);         // :74-75
    cudaStreamSynchronize(nullptr);      // the scratch tensors above are freed on return
  }
};

template <template <typename> class TensorT = Tensor>
class DecoderBlock {  // one iteration of the loop at transformer.cu:91-166
 public:
  int d_model, heads, d_ff;
  MultiHeadAttention<TensorT> self_attn, cross_attn;
  PreparedLinear<TensorT> W_O1, W_O2, ll1, ll2;
  DecoderBlock(int d_model_, int heads_, int d_ff_)
      : d_model(d_model_), heads(heads_), d_ff(d_ff_), self_attn(d_model_, heads_), cross_attn(d_model_, heads_),
        W_O1(d_model_, d_model_, false), W_O2(d_model_, d_model_, false), ll1(d_model_, d_ff_, true), ll2(d_ff_, d_model_, true) {}
  void init_uniform(uint64_t seed) {
    self_attn.init_uniform(seed);
    cross_attn.init_uniform(seed + 5);
    W_O1.init_uniform(seed + 10, -1.0f, 1.0f);  // transformer.cu:117-118
    W_O2.init_uniform(seed + 15, -1.0f, 1.0f);  // transformer.cu:143
    ll1.init_uniform(seed + 20);
    ll2.init_uniform(seed + 30);
  }
  void forward(const TensorT<float> &x, const TensorT<float> &enc_output, TensorT<float> &out, int batch = 1) {
    TensorT<float> mh(x.h, d_model, true), ffn(x.h, d_ff, true);
    self_attn.forward(x, x, mh, batch);              // :97-116
    W_O1.forward(mh, out);                           // :117-119
    add_layernorm(out, mh, out);                     // :123-124
    cross_attn.forward(out, enc_output, mh, batch);  // :127-141 (queries: decoder stream; keys / values: encoder output)
    W_O2.forward(mh, out);                           // :143-144
    add_layernorm(out, mh, out);                     // :148-149
    ll1.forward(out, ffn, QG_ACT_RELU);              // :152-157
    ll2.forward(ffn, out);                           // :159-161
    add_layernorm(out, mh, out);                     // :165-166
    cudaStreamSynchronize(nullptr);
  }
};

// void Encoder(const Tensor<float> &X, Tensor<float> &output, int n_heads, int n_blocks, int d_ff)  transformer.cu:14
template <class TensorF>
void Encoder(const TensorF &X, TensorF &output, int n_heads, int n_blocks, int d_ff, uint64_t seed = 0) {
  assert(X.h == output.h && X.w == output.w && X.w % n_heads == 0);
  for (int i = 0; i < n_blocks; i++) {
    EncoderBlock<Tensor> blk(X.w, n_heads, d_ff);
    blk.init_uniform(seed + 100 * (uint64_t)i);  // fresh weights per block and call, as the reference draws them
    Tensor<float> in = view_of(i == 0 ? X : output), out = view_of(output);
    blk.forward(in, out);  // block 0 reads X, later blocks the running output (:35-39); in place is safe (see transformer.py)
  }
}
// void Decoder(const Tensor<float> &X, Tensor<float> &enc_output, Tensor<float> &output, ...)  transformer.cu:79-80
template <class TensorF>
void Decoder(const TensorF &X, TensorF &enc_output, TensorF &output, int n_heads, int n_blocks, int d_ff, uint64_t seed = 0) {
  assert(X.h == output.h && X.w == output.w && enc_output.w == X.w && X.w % n_heads == 0);
  for (int i = 0; i < n_blocks; i++) {
    DecoderBlock<Tensor> blk(X.w, n_heads, d_ff);
    blk.init_uniform(seed + 100 * (uint64_t)i);
    Tensor<float> in = view_of(i == 0 ? X : output), enc = view_of(enc_output), out = view_of(output);
    blk.forward(in, enc, out);
  }
}

}  // namespace qg_dropin
