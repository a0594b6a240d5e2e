/*
 * qfast.c -- TIMED CPU BASELINE of the quantized-linear hot path.  TEST / BENCH INFRASTRUCTURE ONLY.
 *
 * Same pipeline and the same results, bit for bit, as the oracle's qo_quantized_mm_f32
 * (oracle/qoracle.c, a restatement of /root/reference/src/ops/op_mm.cuh:67-101), written the way a
 * CPU implementation that wants to be fast would write it: every stage threaded with OpenMP, the
 * quantizers as vectorisable loops, and the int8 GEMM (op_mm<int8_t,int>, src/ops/op_mm.cuh:9-46) as
 * a register-blocked AVX-512 VNNI kernel (vpdpbusd) with run-time dispatch -- the prebuilt library
 * travels to the GPU box, whose host CPU may differ from the build container's -- and the oracle's
 * blocked scalar product as the fallback.  bench.py times THIS file for `cpu_baseline` and for
 * `--impl reference` (the reference itself has no CPU implementation of the path: every op asserts
 * device residency, src/ops/op_elemwise.cuh:459-463); tests/test_oracle_cpu.py checks that it
 * agrees with qoracle.c on every stage, edge rows included.  Nothing in the product links it.
 *
 * fp32 arithmetic: IEEE, no contraction (-ffp-contract=off), no fast-math -- as in qoracle.c.
 */
#include <immintrin.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QF_API __attribute__((visibility("default")))
#define QF_MODE_TRUE_ABSMAX 1

QF_API void qf_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
QF_API int qf_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* 2: AVX-512 VNNI kernel, 0: portable fallback */
QF_API int qf_gemm_kernel(void) {
  __builtin_cpu_init();
  return (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
          __builtin_cpu_supports("avx512vnni")) ? 2 : 0;
}

/* AbsMaxFunc (src/ops/op_reduction.cuh:7-25) folded over a strided vector, exactly as the oracle's
 * absmax_step does it: candidate = x > 0 ? x : -x, replace on strict '>'.  NaN never wins. */
static inline void fold(float x, float *acc) {
  const float c = x > 0 ? x : -x;
  if (c > *acc) *acc = c;
}

/* Rows: op_reduction_kernel_colwise (src/ops/op_reduction.cuh:71-92) + InvDivideConstFunc (:132-143 of
 * op_elemwise.cuh) + MultiplyWithTypecastFunc<float,int8_t> (:106-114), one pass per row.  The
 * maximum is taken over 16 independent lanes (any order gives the same VALUE under strict '>'); only
 * the sign of a zero result depends on the order, so rows whose maximum is zero are redone serially. */
static inline int32_t cvt_rzi(float f) { /* PTX cvt.rzi.s32.f32 */
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}

static float absmax_strided(const float *x, int n, int64_t stride, int mode) {
  float first = x[0];
  if (mode == QF_MODE_TRUE_ABSMAX) first = fabsf(first);
  float lane[16];
  for (int l = 0; l < 16; l++) lane[l] = -INFINITY;
  int j = 1;
  if (stride == 1) {
    for (; j + 16 <= n; j += 16)
      for (int l = 0; l < 16; l++) {
        const float v = x[j + l], c = v > 0 ? v : -v;
        lane[l] = c > lane[l] ? c : lane[l];
      }
  }
  for (; j < n; j++) fold(x[(int64_t)j * stride], &lane[0]);
  float m = lane[0];
  for (int l = 1; l < 16; l++) m = lane[l] > m ? lane[l] : m;
  float acc = first;
  if (m > acc) acc = m;
  if (acc == 0.0f) { /* +-0: the reference's visiting order decides the sign */
    acc = first;
    for (j = 1; j < n; j++) fold(x[(int64_t)j * stride], &acc);
  }
  return acc;
}

QF_API void qf_absmax_quant_rows_f32(const float *X, int M, int K, int64_t ld, float range, int mode, int8_t *Xq,
                                     int64_t ldq, float *Cx) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) {
    const float *row = X + (int64_t)i * ld;
    const float c = absmax_strided(row, K, 1, mode);
    Cx[i] = c;
    const float s = range / c;
    int8_t *q = Xq + (int64_t)i * ldq;
    for (int k = 0; k < K; k++) q[k] = (int8_t)(uint8_t)((uint32_t)cvt_rzi(row[k] * s) & 0xffu);
  }
}

/* Columns: op_reduction_kernel_rowwise (src/ops/op_reduction.cuh:96-117), then the same scale and cast.
 * Threads own blocks of 64 columns and walk the rows once for the maxima and once for the codes. */
QF_API void qf_absmax_quant_cols_f32(const float *W, int K, int N, int64_t ld, float range, int mode, int8_t *Wq,
                                     int64_t ldq, float *Cw) {
  enum { CB = 64 };
#pragma omp parallel for schedule(static)
  for (int j0 = 0; j0 < N; j0 += CB) {
    const int jb = N - j0 < CB ? N - j0 : CB;
    float acc[CB], s[CB];
    for (int j = 0; j < jb; j++) acc[j] = -INFINITY;
    for (int k = 1; k < K; k++) {
      const float *w = W + (int64_t)k * ld + j0;
      for (int j = 0; j < jb; j++) {
        const float v = w[j], c = v > 0 ? v : -v;
        acc[j] = c > acc[j] ? c : acc[j];
      }
    }
    for (int j = 0; j < jb; j++) {
      float first = W[j0 + j];
      if (mode == QF_MODE_TRUE_ABSMAX) first = fabsf(first);
      float c = first;
      if (acc[j] > c) c = acc[j];
      if (c == 0.0f) c = absmax_strided(W + j0 + j, K, ld, mode);
      Cw[j0 + j] = c;
      s[j] = range / c;
    }
    for (int k = 0; k < K; k++) {
      const float *w = W + (int64_t)k * ld + j0;
      int8_t *q = Wq + (int64_t)k * ldq + j0;
      for (int j = 0; j < jb; j++) q[j] = (int8_t)(uint8_t)((uint32_t)cvt_rzi(w[j] * s[j]) & 0xffu);
    }
  }
}

/* ---- int8 GEMM ------------------------------------------------------------------------------
 * acc[i,j] = sum_k A[i,k] * B[k,j], exact int32.  vpdpbusd multiplies UNSIGNED by signed bytes, so A
 * is biased by +128 and 128 * colsum(B) is taken off again: sum (a+128) b = sum a b + 128 sum b.
 * B is repacked once per call as Bp[k/4][j][4] (four consecutive k of one column in a dword). */
__attribute__((target("avx512f,avx512bw,avx512vnni")))
static void gemm_vnni_block(const uint8_t *Au, int64_t lda, const int8_t *Bp, int Np, int K4, int rows, int j0,
                            const int32_t *corr, int32_t *C, int64_t ldc, int N) {
  /* rows <= 6 rows of A against 64 columns starting at j0 */
  __m512i acc[6][4];
  for (int r = 0; r < 6; r++)
    for (int c = 0; c < 4; c++) acc[r][c] = _mm512_setzero_si512();
  for (int k4 = 0; k4 < K4; k4++) {
    const int8_t *bp = Bp + ((int64_t)k4 * Np + j0) * 4;
    const __m512i b0 = _mm512_loadu_si512(bp), b1 = _mm512_loadu_si512(bp + 64), b2 = _mm512_loadu_si512(bp + 128),
                  b3 = _mm512_loadu_si512(bp + 192);
#pragma GCC unroll 6
    for (int r = 0; r < 6; r++) {
      if (r < rows) {
        const __m512i a = _mm512_set1_epi32(*(const int32_t *)(Au + (int64_t)r * lda + 4 * k4));
        acc[r][0] = _mm512_dpbusd_epi32(acc[r][0], a, b0);
        acc[r][1] = _mm512_dpbusd_epi32(acc[r][1], a, b1);
        acc[r][2] = _mm512_dpbusd_epi32(acc[r][2], a, b2);
        acc[r][3] = _mm512_dpbusd_epi32(acc[r][3], a, b3);
      }
    }
  }
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < 4; c++) {
      const int j = j0 + 16 * c;
      if (j >= N) break;
      const __m512i v = _mm512_sub_epi32(acc[r][c], _mm512_loadu_si512(corr + j));
      if (j + 16 <= N) {
        _mm512_storeu_si512(C + (int64_t)r * ldc + j, v);
      } else {
        int32_t tmp[16];
        _mm512_storeu_si512(tmp, v);
        for (int t = 0; t < N - j; t++) C[(int64_t)r * ldc + j + t] = tmp[t];
      }
    }
}

static void gemm_portable(const int8_t *A, const int8_t *B, int M, int N, int K, int64_t lda, int64_t ldb, int32_t *C,
                          int64_t ldc) {
  enum { JB = 1024, IB = 4 };
#pragma omp parallel for schedule(dynamic, 1)
  for (int i0 = 0; i0 < M; i0 += IB) {
    int32_t acc[IB][JB];
    const int ib = (M - i0 < IB) ? (M - i0) : IB;
    for (int j0 = 0; j0 < N; j0 += JB) {
      const int jb = (N - j0 < JB) ? (N - j0) : JB;
      for (int r = 0; r < ib; r++) memset(acc[r], 0, sizeof(int32_t) * (size_t)jb);
      for (int k = 0; k < K; k++) {
        const int8_t *brow = B + (int64_t)k * ldb + j0;
        for (int r = 0; r < ib; r++) {
          const int32_t a = A[(int64_t)(i0 + r) * lda + k];
          int32_t *ar = acc[r];
          for (int j = 0; j < jb; j++) ar[j] += a * (int32_t)brow[j];
        }
      }
      for (int r = 0; r < ib; r++) memcpy(C + (int64_t)(i0 + r) * ldc + j0, acc[r], sizeof(int32_t) * (size_t)jb);
    }
  }
}

QF_API int qf_gemm_s8s8s32(const int8_t *A, const int8_t *B, int M, int N, int K, int64_t lda, int64_t ldb, int32_t *C,
                           int64_t ldc) {
  if (qf_gemm_kernel() != 2) {
    gemm_portable(A, B, M, N, K, lda, ldb, C, ldc);
    return 0;
  }
  const int K4 = (K + 3) / 4, Kp = 4 * K4, Np = (N + 63) / 64 * 64;
  uint8_t *Au = (uint8_t *)aligned_alloc(64, ((size_t)M * Kp + 63) / 64 * 64);
  int8_t *Bp = (int8_t *)aligned_alloc(64, (size_t)K4 * Np * 4);
  int32_t *corr = (int32_t *)aligned_alloc(64, sizeof(int32_t) * (size_t)Np);
  if (!Au || !Bp || !corr) { free(Au); free(Bp); free(corr); return -1; }
#pragma omp parallel
  {
#pragma omp for schedule(static) nowait
    for (int i = 0; i < M; i++) {
      const int8_t *a = A + (int64_t)i * lda;
      uint8_t *u = Au + (int64_t)i * Kp;
      for (int k = 0; k < K; k++) u[k] = (uint8_t)((int)a[k] + 128);
      for (int k = K; k < Kp; k++) u[k] = 128; /* bias of a zero: cancels against the zero rows of Bp */
    }
#pragma omp for schedule(static)
    for (int k4 = 0; k4 < K4; k4++) {
      int8_t *bp = Bp + (int64_t)k4 * Np * 4;
      for (int j = 0; j < Np; j++)
        for (int t = 0; t < 4; t++) {
          const int k = 4 * k4 + t;
          bp[4 * j + t] = (k < K && j < N) ? B[(int64_t)k * ldb + j] : 0;
        }
    }
#pragma omp for schedule(static)
    for (int j = 0; j < Np; j++) {
      int32_t s = 0;
      if (j < N)
        for (int k = 0; k < K; k++) s += B[(int64_t)k * ldb + j];
      corr[j] = 128 * s;
    }
    /* one task = 64 columns x a band of rows; the B panel (K x 64 bytes x 4) stays in the core's L2 */
    const int bands = (M + 95) / 96;
#pragma omp for schedule(dynamic, 1) collapse(2)
    for (int jp = 0; jp < Np / 64; jp++)
      for (int b = 0; b < bands; b++) {
        const int i1 = (b + 1) * 96 < M ? (b + 1) * 96 : M;
        for (int i = b * 96; i < i1; i += 6)
          gemm_vnni_block(Au + (int64_t)i * Kp, Kp, Bp, Np, K4, i1 - i < 6 ? i1 - i : 6, jp * 64, corr,
                          C + (int64_t)i * ldc, ldc, N);
      }
  }
  free(Au); free(Bp); free(corr);
  return 0;
}

/* a6-a8 (+a10): see qo_dequant_f32 in qoracle.c for the citations; three (four) separately rounded operations */
QF_API void qf_dequant_f32(const int32_t *acc, int64_t lda, const float *Cx, const float *Cw, const float *bias, int M,
                           int N, float range, float *O, int64_t ldo) {
  const float c = 1 / (range * range);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) {
    const float cx = Cx[i];
    const int32_t *a = acc + (int64_t)i * lda;
    float *o = O + (int64_t)i * ldo;
    for (int j = 0; j < N; j++) {
      float outer = cx * Cw[j];
      outer = outer + 0.0f;
      float v = (float)a[j] * outer;
      v = v * c;
      if (bias) v = v + bias[j];
      o[j] = v;
    }
  }
}

/* op_quantized_mm<float> (src/ops/op_mm.cuh:67-101) end to end; scratch is allocated per call like the
 * reference does (op_mm.cuh:76-96).  Optional outputs as in qo_quantized_mm_f32. */
QF_API int qf_quantized_mm_f32(const float *X, const float *W, float *O, int M, int N, int K, int64_t ldx, int64_t ldw,
                               int64_t ldo, float range, int mode, const float *bias, float *Cx_out, float *Cw_out,
                               int8_t *Xq_out, int8_t *Wq_out, int32_t *acc_out) {
  float *Cx = (float *)malloc(sizeof(float) * (size_t)M);
  float *Cw = (float *)malloc(sizeof(float) * (size_t)N);
  int8_t *Xq = (int8_t *)malloc((size_t)M * K);
  int8_t *Wq = (int8_t *)malloc((size_t)K * N);
  int32_t *acc = (int32_t *)malloc(sizeof(int32_t) * (size_t)M * N);
  int rc = -1;
  if (Cx && Cw && Xq && Wq && acc) {
    qf_absmax_quant_rows_f32(X, M, K, ldx, range, mode, Xq, K, Cx);
    qf_absmax_quant_cols_f32(W, K, N, ldw, range, mode, Wq, N, Cw);
    rc = qf_gemm_s8s8s32(Xq, Wq, M, N, K, K, N, acc, N);
    if (rc == 0) {
      qf_dequant_f32(acc, N, Cx, Cw, bias, M, N, range, O, ldo);
      if (Cx_out) memcpy(Cx_out, Cx, sizeof(float) * (size_t)M);
      if (Cw_out) memcpy(Cw_out, Cw, sizeof(float) * (size_t)N);
      if (Xq_out) memcpy(Xq_out, Xq, (size_t)M * K);
      if (Wq_out) memcpy(Wq_out, Wq, (size_t)K * N);
      if (acc_out) memcpy(acc_out, acc, sizeof(int32_t) * (size_t)M * N);
    }
  }
  free(Cx); free(Cw); free(Xq); free(Wq); free(acc);
  return rc;
}
