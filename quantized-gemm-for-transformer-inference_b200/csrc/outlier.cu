// outlier.cu -- LLM.int8()-style mixed decomposition: the pieces around the GEMM.
//
// The reference only has the elementwise mask primitive (op_outlier_extractor,
// src/ops/op_elemwise.cuh:292-306,698-708), never called and never reduced: detection of outlier
// feature columns, the split and the 16-bit side product are new here (specification in DESIGN.md,
// CPU restatement in oracle.quantized_mm_outlier; parity unpinned).
//
//   outlier columns  O = { k : some |X[i,k]| > thr }   (strict; NaN counts, like the reference functor)
//   int8 path        X with the columns in O zeroed, row scales from the remaining entries
//   side product     fp16(X[:,O]) @ fp16(W[O,:]) accumulated in fp32 inside the GEMM epilogue
//
// Kernels: column detection (HBM-bound, one read of X), index compaction, bit-mask construction
// from a given index list, the masking row quantizer (also gathers X[:,O]), and the W[O,:] gather.
#include "quant_common.cuh"

namespace qg {

namespace {

// ---- detection: mask bit k set <=> column k holds an outlier --------------------------------
// 32 x 8 threads; a thread owns one 16-byte vector of columns and walks rows with stride 8.
template <typename T>
__global__ void __launch_bounds__(kThreads)
outlier_detect_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float thr, int rows_per_cta,
                      uint32_t *__restrict__ mask) {
  constexpr int EPV = Unpack<T>::EPV;
  __shared__ uint32_t s_bits[8][32];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * EPV;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  griddep_wait();
  griddep_trigger_early();
  uint32_t bits = 0;
  if (col < K) {
    const T *base = X + col;
    auto scan = [&](const uint4 &v) {
      float f[EPV];
      Unpack<T>::run(v, f);
#pragma unroll
      for (int e = 0; e < EPV; e++) {
        // AbsCompareLTEConstFunc: inlier iff (a>=0 & a<=thr) | (a<=0 & -a<=thr), i.e. |a| <= thr with NaN an outlier
        bits |= (fabsf(f[e]) <= thr ? 0u : 1u) << e;
      }
    };
    int r = r0 + ty;
    for (; r + 56 < r1; r += 64) {  // eight independent 16-byte loads in flight per thread
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = ldg16(base + (int64_t)(r + 8 * u) * ldx);
#pragma unroll
      for (int u = 0; u < 8; u++) scan(v[u]);
    }
    for (; r < r1; r += 8) scan(ldg16(base + (int64_t)r * ldx));
  }
  s_bits[ty][tx] = bits;
  __syncthreads();
  if (ty == 0 && col < K) {
#pragma unroll
    for (int y = 1; y < 8; y++) bits |= s_bits[y][tx];
    if (bits) atomicOr(mask + (col >> 5), bits << (col & 31));
  }
}

template <typename T>
__global__ void outlier_detect_generic_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float thr,
                                              uint32_t *__restrict__ mask) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  if (k >= K) return;
  bool out = false;
  for (int i = 0; i < M; i++) {
    const float a = to_f32(X[(int64_t)i * ldx + k]);
    out |= !(((a >= 0) & (a <= thr)) | ((a <= 0) & (-a <= thr)));
  }
  if (out) atomicOr(mask + (k >> 5), 1u << (k & 31));
}

// ---- block-wide exclusive prefix of per-word popcounts (single block of 1024 threads) -------
// wbase[w] = number of set bits in mask words < w; returns the total in every thread.
__device__ int mask_prefix(const uint32_t *mask, int words, int *wbase) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int w0 = 0; w0 < words; w0 += 1024) {
    const int w = w0 + threadIdx.x;
    const int cnt = w < words ? __popc(mask[w]) : 0;
    int incl = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, off);
      if ((threadIdx.x & 31) >= off) incl += t;
    }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
      int v = s_warp[threadIdx.x];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, off);
        if (threadIdx.x >= off) v += t;
      }
      s_warp[threadIdx.x] = v;
    }
    __syncthreads();
    const int carry = s_carry;
    const int before = carry + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0) + incl - cnt;
    if (w < words) wbase[w] = before;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = before + cnt;
    __syncthreads();
  }
  return s_carry;
}

// mask -> ascending index list (first max_idx entries), total count, per-word prefix
__global__ void __launch_bounds__(1024)
outlier_index_kernel(const uint32_t *__restrict__ mask, int K, int *__restrict__ idx, int max_idx,
                     int *__restrict__ count, int *__restrict__ wbase) {
  griddep_wait();
  griddep_trigger_early();
  const int words = (K + 31) / 32;
  const int total = mask_prefix(mask, words, wbase);
  for (int w = threadIdx.x; w < words; w += 1024) {
    uint32_t m = mask[w];
    int pos = wbase[w];
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      if (pos < max_idx && idx != nullptr) idx[pos] = w * 32 + b;
      pos++;
    }
  }
  if (threadIdx.x == 0 && count != nullptr) *count = total;
}

// given ascending indices -> bit mask + per-word prefix
__global__ void __launch_bounds__(1024)
outlier_mask_from_idx_kernel(const int *__restrict__ idx, int n_idx, int K, uint32_t *__restrict__ mask,
                             int *__restrict__ wbase) {
  griddep_wait();
  griddep_trigger_early();
  const int words = (K + 31) / 32;
  for (int w = threadIdx.x; w < words; w += 1024) mask[w] = 0;
  __syncthreads();
  for (int o = threadIdx.x; o < n_idx; o += 1024) {
    const int k = idx[o];
    if (k >= 0 && k < K) atomicOr(mask + (k >> 5), 1u << (k & 31));
  }
  __syncthreads();
  const int total = mask_prefix(mask, words, wbase);
  if (threadIdx.x == 0) wbase[words] = total;  // number of distinct outlier columns (one int past the per-word prefixes)
}

// The masking quantizer IS the row quantizer of quantize.cu (quant_rows_body: G threads per row, NV vectors per thread kept in
// registers between the reduction and the codes, the next row block loaded ahead) with one hook: every vector passes
// through strip_outliers before it is reduced, so the outlier entries read as +0 for the scale and the codes and are
// written to the side operand.  (Round 1 had its own one-CTA-per-row kernel without the look-ahead: 51 us against 37 us
// for the plain quantizer at 16384 x 4096 fp16.)
template <typename T, int G, int NV>
__global__ void __launch_bounds__(kThreads, (NV <= 4 ? 4 : 2))  // the strip code would otherwise take 96 registers: 2 CTAs per SM
quant_rows_outlier_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                          int8_t *__restrict__ Xq, int64_t ldq, float *__restrict__ Cx, RowOutlier ro) {
  griddep_wait();
  griddep_trigger_early();
  quant_rows_body<T, G, NV, false, true>(X, M, K, ldx, range, mode, nullptr, Xq, ldq, Cx, RowMaxIo(), (int)blockIdx.x,
                                         (int)gridDim.x, ro);
}

// generic shapes: one warp per row, scalar accesses
template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_rows_outlier_generic_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                                  int8_t *__restrict__ Xq, int64_t ldq, float *__restrict__ Cx, RowOutlier ro) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  griddep_wait();
  griddep_trigger_early();
  if (row >= M) return;
  const T *xr = X + (int64_t)row * ldx;
  if (lane == 0) zero_side_tail(row, ro);
  auto is_out = [&](int j) { return ((__ldg(ro.mask + (j >> 5)) >> (j & 31)) & 1u) != 0; };
  auto pos_of = [&](int j) { return __ldg(ro.wbase + (j >> 5)) + __popc(__ldg(ro.mask + (j >> 5)) & ((1u << (j & 31)) - 1u)); };
  float m = -INFINITY;
  for (int j = lane; j < K; j += 32) {
    const float x = to_f32(xr[j]);
    if (is_out(j)) {
      if (ro.side_bf16) reinterpret_cast<__nv_bfloat16 *>(ro.Xo)[(int64_t)row * ro.ldxo + pos_of(j)] = __float2bfloat16_rn(x);
      else reinterpret_cast<__half *>(ro.Xo)[(int64_t)row * ro.ldxo + pos_of(j)] = __float2half_rn(x);
    } else if (j > 0) {
      m = fmaxf(m, fabsf(x));
    }
    if (is_out(j) && j > 0) m = fmaxf(m, 0.0f);
  }
  m = warp_max(m);
  const float x0 = is_out(0) ? 0.0f : to_f32(xr[0]);
  float c;
  if (fold_first(x0, m, mode, c)) {
    for (int j = 1; j < K; j++) {
      const float xj = is_out(j) ? 0.0f : to_f32(xr[j]);
      if (xj == xj) { c = -xj; break; }
    }
  }
  if (lane == 0 && Cx != nullptr) Cx[row] = c;
  const float scale = __fdiv_rn(range, c);
  for (int j = lane; j < K; j += 32)
    Xq[(int64_t)row * ldq + j] = (int8_t)quant_code_u8(is_out(j) ? 0.0f : to_f32(xr[j]), scale);
}

// Wo[o][j] = 16-bit(W[k_o, j]) where k_o is the o-th set bit of the mask (the same rank the row quantizer files X[:, k_o]
// under), zero rows for o >= count.  Taking the row from the MASK rather than from idx[o] keeps the two operands of the
// side product paired whatever order (or duplicates, or out-of-range entries) the caller's index list has.
template <typename T>
__global__ void gather_wo_kernel(const T *__restrict__ W, int64_t ldw, const uint32_t *__restrict__ mask,
                                 const int *__restrict__ wbase, int words, int N, void *__restrict__ Wo, int64_t ldwo,
                                 int side_bf16) {
  __shared__ int s_k;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int o = blockIdx.y;
  griddep_wait();
  griddep_trigger_early();
  if (threadIdx.x == 0) {
    int k = -1;
    if (o < wbase[words]) {  // wbase[words] = number of set bits
      int lo = 0, hi = words - 1;  // last word whose prefix count is <= o
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (wbase[mid] <= o) lo = mid; else hi = mid - 1;
      }
      uint32_t m = mask[lo];
      for (int r = o - wbase[lo]; r > 0; r--) m &= m - 1;
      k = lo * 32 + __ffs(m) - 1;
    }
    s_k = k;
  }
  __syncthreads();
  if (j >= N) return;
  const int k = s_k;
  const float v = k >= 0 ? to_f32(W[(int64_t)k * ldw + j]) : 0.0f;
  if (side_bf16) reinterpret_cast<__nv_bfloat16 *>(Wo)[(int64_t)o * ldwo + j] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half *>(Wo)[(int64_t)o * ldwo + j] = __float2half_rn(v);
}

inline bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

template <typename T>
int detect_t(const T *X, int M, int K, int64_t ldx, float thr, uint32_t *mask, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  cudaError_t e = cudaMemsetAsync(mask, 0, sizeof(uint32_t) * (size_t)ceil_div(K, 32), st);
  if (e != cudaSuccess) return (int)e;
  if (K % EPV == 0 && aligned_to(X, 16) && (ldx * sizeof(T)) % 16 == 0) {
    const int col_tiles = (int)ceil_div(K, 32 * EPV);
    int64_t want = ceil_div((int64_t)148 * 6, col_tiles);  // one wave of resident CTAs, as the column quantizer's pass 1
    const int rpc = (int)round_up(ceil_div(M, want > 0 ? want : 1) < 32 ? 32 : ceil_div(M, want > 0 ? want : 1), 32);
    return (int)launch_kernel(outlier_detect_kernel<T>, dim3(col_tiles, (unsigned)ceil_div(M, rpc)), dim3(kThreads), st, X, M, K,
                              ldx, thr, rpc, mask);
  }
  return (int)launch_kernel(outlier_detect_generic_kernel<T>, dim3((unsigned)ceil_div(K, 256)), dim3(256), st, X, M, K, ldx,
                            thr, mask);
}

template <typename T>
int rows_outlier_t(const T *X, int M, int K, int64_t ldx, float range, int mode, int8_t *Xq, int64_t ldq, float *Cx,
                   const RowOutlier &ro, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  const bool vec_ok = (K % EPV == 0) && aligned_to(X, 16) && ((ldx * sizeof(T)) % 16 == 0) && aligned_to(Xq, EPV) &&
                      ldq % EPV == 0;
  if (!vec_ok)
    return (int)launch_kernel(quant_rows_outlier_generic_kernel<T>, dim3((unsigned)ceil_div(M, kThreads / 32)), dim3(kThreads),
                              st, X, M, K, ldx, range, mode, Xq, ldq, Cx, ro);
  const int nvec = K / EPV;
#define QG_ROWS_OUTL(G, NV)                                                                                                   \
  do {                                                                                                                        \
    constexpr int RPB = kThreads / (G);                                                                                       \
    static int resident = 0;                                                                                                  \
    if (resident == 0) {                                                                                                      \
      int per_sm = 0, dev = 0, sms = 0;                                                                                       \
      cudaGetDevice(&dev);                                                                                                    \
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                                                      \
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quant_rows_outlier_kernel<T, G, NV>, kThreads, 0);               \
      resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);                                                           \
    }                                                                                                                         \
    const int64_t nrb = ceil_div(M, RPB);                                                                                     \
    const unsigned grid = (unsigned)(nrb < resident ? nrb : resident);                                                        \
    return (int)launch_kernel(quant_rows_outlier_kernel<T, G, NV>, dim3(grid), dim3(kThreads), st, X, M, K, ldx, range, mode, \
                              Xq, ldq, Cx, ro);                                                                               \
  } while (0)
  // the same (G, NV) table as the plain quantizer (quantize.cu: rows_dispatch)
  if (nvec <= 32) QG_ROWS_OUTL(32, 1);
  else if (nvec <= 64) QG_ROWS_OUTL(32, 2);
  else if (nvec <= 128) QG_ROWS_OUTL(32, 4);
  else if (nvec <= 256) QG_ROWS_OUTL(64, 4);
  else if (nvec <= 512) QG_ROWS_OUTL(128, 4);
  else if (nvec <= 1024) QG_ROWS_OUTL(256, 4);
  else if (nvec <= 2048) QG_ROWS_OUTL(256, 8);
  else if (nvec <= 4096) QG_ROWS_OUTL(256, 16);
  else QG_ROWS_OUTL(256, 0);
#undef QG_ROWS_OUTL
}

}  // namespace

// ---- entry points used by capi.cu ----
int outlier_detect(const void *X, int dtype, int M, int K, int64_t ldx, float thr, uint32_t *mask, cudaStream_t st) {
  switch (dtype) {
    case QG_F32: return detect_t((const float *)X, M, K, ldx, thr, mask, st);
    case QG_F16: return detect_t((const __half *)X, M, K, ldx, thr, mask, st);
    case QG_BF16: return detect_t((const __nv_bfloat16 *)X, M, K, ldx, thr, mask, st);
  }
  return QG_EINVAL;
}

int outlier_index(const uint32_t *mask, int K, int *idx, int max_idx, int *count, int *wbase, cudaStream_t st) {
  return (int)launch_kernel(outlier_index_kernel, dim3(1), dim3(1024), st, mask, K, idx, max_idx, count, wbase);
}

int outlier_mask_from_idx(const int *idx, int n_idx, int K, uint32_t *mask, int *wbase, cudaStream_t st) {
  return (int)launch_kernel(outlier_mask_from_idx_kernel, dim3(1), dim3(1024), st, idx, n_idx, K, mask, wbase);
}

int quant_rows_outlier(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, int8_t *Xq, int64_t ldq,
                       float *Cx, const uint32_t *mask, const int *wbase, void *Xo, int64_t ldxo, int side_bf16,
                       cudaStream_t st) {
  RowOutlier ro = {mask, wbase, Xo, ldxo, side_bf16, wbase + (K + 31) / 32};
  switch (dtype) {
    case QG_F32: return rows_outlier_t((const float *)X, M, K, ldx, range, mode, Xq, ldq, Cx, ro, st);
    case QG_F16: return rows_outlier_t((const __half *)X, M, K, ldx, range, mode, Xq, ldq, Cx, ro, st);
    case QG_BF16: return rows_outlier_t((const __nv_bfloat16 *)X, M, K, ldx, range, mode, Xq, ldq, Cx, ro, st);
  }
  return QG_EINVAL;
}

int gather_wo(const void *W, int dtype, int64_t ldw, const uint32_t *mask, const int *wbase, int K, int no_pad, int N, void *Wo,
              int64_t ldwo, int side_bf16, cudaStream_t st) {
  if (no_pad <= 0) return 0;
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)no_pad);
  const int words = (K + 31) / 32;
  switch (dtype) {
    case QG_F32: return (int)launch_kernel(gather_wo_kernel<float>, grid, dim3(256), st, (const float *)W, ldw, mask, wbase, words, N, Wo, ldwo, side_bf16);
    case QG_F16: return (int)launch_kernel(gather_wo_kernel<__half>, grid, dim3(256), st, (const __half *)W, ldw, mask, wbase, words, N, Wo, ldwo, side_bf16);
    case QG_BF16: return (int)launch_kernel(gather_wo_kernel<__nv_bfloat16>, grid, dim3(256), st, (const __nv_bfloat16 *)W, ldw, mask, wbase, words, N, Wo, ldwo, side_bf16);
  }
  return QG_EINVAL;
}

}  // namespace qg
