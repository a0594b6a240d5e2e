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

// op_multiply<T,int8_t>(a, scale, out): src/ops/op_elemwise.cuh:629-640
template <class TensorA, class TensorQ>
void op_multiply(const TensorA &a, const TensorA &b, TensorQ &out) {
  static_assert(sizeof(*out.rawp) == 1, "quantizing overload: out must be int8");
  assert(out.h == a.h && out.w == a.w);
  assert((a.h == b.h && b.w == 1) || (a.w == b.w && b.h == 1));
  assert(a.on_device && b.on_device && out.on_device);
  if (b.w == 1 && a.h == b.h && a.w != b.w)
    check(qg_quantize_rows(base_ptr(a), QG_F32, a.h, a.w, a.stride_h, base_ptr(b), base_ptr(out), out.stride_h, nullptr),
          "qg_quantize_rows");
  else
    check(qg_quantize_cols(base_ptr(a), QG_F32, a.h, a.w, a.stride_h, base_ptr(b), base_ptr(out), out.stride_h, nullptr),
          "qg_quantize_cols");
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

}  // namespace qg_dropin
