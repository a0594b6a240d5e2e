// quantize.cu -- vector-wise absmax quantizers.
//
// Replaces, per matrix, the reference's three launches
//   op_absmax        (src/ops/op_reduction.cuh:195-204; one thread per vector, serial loop)
//   op_inv_divide    (src/ops/op_elemwise.cuh:657-667)
//   op_multiply<T,int8_t> (src/ops/op_elemwise.cuh:629-640; one element per thread, byte stores)
// with HBM-bound kernels that read the matrix with 16-byte loads, reduce with warp shuffles and
// write packed int8 codes plus the fp32 absmax.  The arithmetic (and its quirks: signed first
// element, IEEE 127/x, truncate-and-wrap cast) is reproduced exactly; see oracle/qoracle.c.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace qg {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg16(const void *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}

// unpack one 16-byte vector into fp32 lanes
template <typename T> struct Unpack;
template <> struct Unpack<float> {
  static constexpr int EPV = 4;
  __device__ static __forceinline__ void run(const uint4 &r, float (&f)[4]) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
};
template <> struct Unpack<__half> {
  static constexpr int EPV = 8;
  __device__ static __forceinline__ void run(const uint4 &r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 p = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
      f[2 * i] = p.x; f[2 * i + 1] = p.y;
    }
  }
};
template <> struct Unpack<__nv_bfloat16> {
  static constexpr int EPV = 8;
  __device__ static __forceinline__ void run(const uint4 &r, float (&f)[8]) {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {  // bf16 -> fp32 is a 16-bit shift
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// Fold the signed first element x0 with m = max_{j>=1, x_j not NaN} |x_j| (m = -inf when there is
// no such j) the way AbsMaxFunc does (src/ops/op_reduction.cuh:7-25,80-83): strict '>' updates.
// Returns true when the +-0 tie-break path is needed (all later entries are zeros, x0 < 0): the
// reference then ends with -x_j of the first non-NaN j >= 1, whose zero sign the caller must fetch.
__device__ __forceinline__ bool fold_first(float x0, float m, int mode, float &c) {
  if (mode == QG_MODE_TRUE_ABSMAX) {
    const float a0 = fabsf(x0);
    c = (x0 != x0) ? x0 : ((m > a0) ? m : a0);
    return false;
  }
  c = (x0 != x0) ? x0 : ((m > x0) ? m : x0);
  return (m == 0.0f) && (x0 < 0.0f);
}

__device__ __forceinline__ float warp_max(float m) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  return m;
}

// ------------------------------------------------------------------------------------------
// Rows (activations): G threads cooperate on one row, NV 16-byte vectors per thread kept in
// registers between the reduction and the quantizing pass (NV == 0: row too long, re-read it;
// the second read is served by L2).  One read of X, one write of Xq.
//   sx_in != NULL : scales given (plain op_multiply<T,int8_t>), no reduction
//   Xq   == NULL  : reduction only (plain op_absmax)
// ------------------------------------------------------------------------------------------
template <typename T, int G, int NV>
__global__ void __launch_bounds__(kThreads)
quant_rows_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                  const float *__restrict__ sx_in, int8_t *__restrict__ Xq, int64_t ldq,
                  float *__restrict__ Cx) {
  constexpr int EPV = Unpack<T>::EPV;
  constexpr int RPB = kThreads / G;  // rows per block iteration
  constexpr int WPR = G / 32;        // warps per row
  constexpr int NVC = NV > 0 ? NV : 1;
  constexpr bool kPrefetch = NV > 0 && NV <= 8;  // next row block's vectors are loaded ahead
  __shared__ float s_m[2][kThreads / 32];
  __shared__ float s_x0[2][RPB];

  const int rib = threadIdx.x / G;
  const int g = threadIdx.x % G;
  const int nvec = K / EPV;
  const int nrb = (M + RPB - 1) / RPB;

  auto load = [&](uint4 (&dst)[NVC], int rb) {
    const int row = rb * RPB + rib;
    const T *xr = X + (int64_t)(row < M ? row : 0) * ldx;
#pragma unroll
    for (int v = 0; v < NVC; v++) {
      const int idx = v * G + g;
      dst[v] = (row < M && idx < nvec) ? ldg16(xr + (int64_t)idx * EPV) : make_uint4(0, 0, 0, 0);
    }
  };

  uint4 raw[NVC], nxt[NVC];
  int rb = blockIdx.x;
  griddep_wait();
  if (NV > 0 && rb < nrb) load(raw, rb);
  for (int it = 0; rb < nrb; rb += gridDim.x, it++) {
    if (kPrefetch && rb + (int)gridDim.x < nrb) load(nxt, rb + gridDim.x);
    else griddep_launch_dependents();  // last row block of this CTA: let the next kernel ramp up
    const int row = rb * RPB + rib;
    const bool active = row < M;
    const T *xr = X + (int64_t)(active ? row : 0) * ldx;
    float scale;
    if (sx_in == nullptr) {
      float m = -INFINITY, x0 = 0.0f;
      if (NV > 0) {
#pragma unroll
        for (int v = 0; v < NVC; v++) {
          const int idx = v * G + g;
          if (idx < nvec) {
            float f[EPV];
            Unpack<T>::run(raw[v], f);
            if (idx == 0) x0 = f[0];
#pragma unroll
            for (int e = 0; e < EPV; e++)
              if (e > 0 || idx > 0) m = fmaxf(m, fabsf(f[e]));
          }
        }
      } else {
        for (int idx = g; idx < nvec && active; idx += G) {
          float f[EPV];
          Unpack<T>::run(ldg16(xr + (int64_t)idx * EPV), f);
          if (idx == 0) x0 = f[0];
#pragma unroll
          for (int e = 0; e < EPV; e++)
            if (e > 0 || idx > 0) m = fmaxf(m, fabsf(f[e]));
        }
      }
      m = warp_max(m);
      if (WPR > 1) {  // double-buffered by iteration parity: one barrier per iteration is enough
        float *sm = s_m[it & 1], *sx0 = s_x0[it & 1];
        if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
        if (g == 0) sx0[rib] = x0;
        __syncthreads();
        m = sm[rib * WPR];
#pragma unroll
        for (int w = 1; w < WPR; w++) m = fmaxf(m, sm[rib * WPR + w]);
        x0 = sx0[rib];
      } else {
        x0 = __shfl_sync(0xffffffffu, x0, 0);
      }
      float c;
      if (fold_first(x0, m, mode, c) && active) {
        for (int j = 1; j < K; j++) {  // rare: sign of the first later zero decides (+-0 tie-break)
          const float xj = to_f32(xr[j]);
          if (xj == xj) { c = -xj; break; }
        }
      }
      if (active && g == 0 && Cx != nullptr) Cx[row] = c;
      scale = __fdiv_rn(range, c);  // InvDivideConstFunc: b / x, IEEE division
    } else {
      scale = active ? sx_in[row] : 0.0f;
    }
    if (Xq != nullptr && active) {
      int8_t *qr = Xq + (int64_t)row * ldq;
      auto emit = [&](const uint4 &r, int idx) {
        float f[EPV];
        Unpack<T>::run(r, f);
        uint32_t w[EPV / 4];
#pragma unroll
        for (int q = 0; q < EPV / 4; q++)
          w[q] = quant_code_u8(f[4 * q], scale) | (quant_code_u8(f[4 * q + 1], scale) << 8) |
                 (quant_code_u8(f[4 * q + 2], scale) << 16) | (quant_code_u8(f[4 * q + 3], scale) << 24);
        if (EPV == 4) *reinterpret_cast<uint32_t *>(qr + (int64_t)idx * 4) = w[0];
        else *reinterpret_cast<uint2 *>(qr + (int64_t)idx * 8) = make_uint2(w[0], w[EPV / 4 - 1]);
      };
      if (NV > 0) {
#pragma unroll
        for (int v = 0; v < NVC; v++) {
          const int idx = v * G + g;
          if (idx < nvec) emit(raw[v], idx);
        }
      } else {
        for (int idx = g; idx < nvec; idx += G) emit(ldg16(xr + (int64_t)idx * EPV), idx);
      }
    }
    if (kPrefetch) {
#pragma unroll
      for (int v = 0; v < NVC; v++) raw[v] = nxt[v];
    } else if (NV > 0 && rb + (int)gridDim.x < nrb) {
      load(raw, rb + gridDim.x);
    }
  }
}

// Any K, any alignment: one warp per row, scalar accesses.  Same arithmetic.
template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_rows_generic_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                          const float *__restrict__ sx_in, int8_t *__restrict__ Xq, int64_t ldq,
                          float *__restrict__ Cx) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  griddep_wait();
  if (row >= M) return;
  const T *xr = X + (int64_t)row * ldx;
  float scale;
  if (sx_in == nullptr) {
    float m = -INFINITY;
    for (int j = lane; j < K; j += 32)
      if (j > 0) m = fmaxf(m, fabsf(to_f32(xr[j])));
    m = warp_max(m);
    const float x0 = to_f32(xr[0]);
    float c;
    if (fold_first(x0, m, mode, c)) {
      for (int j = 1; j < K; j++) {
        const float xj = to_f32(xr[j]);
        if (xj == xj) { c = -xj; break; }
      }
    }
    if (lane == 0 && Cx != nullptr) Cx[row] = c;
    scale = __fdiv_rn(range, c);
  } else {
    scale = sx_in[row];
  }
  if (Xq == nullptr) return;
  for (int j = lane; j < K; j += 32)
    Xq[(int64_t)row * ldq + j] = (int8_t)quant_code_u8(to_f32(xr[j]), scale);
}

// ------------------------------------------------------------------------------------------
// Columns (weights, [K,N] row-major): pass 1 reduces |W[k,j]|, k >= 1, per column into
// part[j] (fp32 bit pattern, combined across row-chunks with a signed-int atomicMax: every
// candidate is >= +0 and the initial value is -inf); pass 2 folds in the signed row 0, forms the
// scale and writes codes.  The second read of W is served largely by the 126 MB L2.
// Thread layout: 32 x 8; a thread owns one 16-byte vector of columns and walks rows with stride 8.
// ------------------------------------------------------------------------------------------
__global__ void fill_f32_kernel(float *p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (i < n) p[i] = v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
absmax_cols_partial_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, int rows_per_cta,
                           float *__restrict__ part) {
  constexpr int EPV = Unpack<T>::EPV;
  __shared__ float s_m[8][32 * EPV + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * EPV;
  const int k0 = 1 + blockIdx.y * rows_per_cta;
  const int k1 = min(K, k0 + rows_per_cta);
  griddep_wait();
  float m[EPV];
#pragma unroll
  for (int e = 0; e < EPV; e++) m[e] = -INFINITY;
  if (col < N) {
    const T *base = W + col;
    int k = k0 + ty;
    for (; k + 24 < k1; k += 32) {  // 4 independent 16-byte loads in flight per thread
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; u++) r[u] = ldg16(base + (int64_t)(k + 8 * u) * ldw);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        float f[EPV];
        Unpack<T>::run(r[u], f);
#pragma unroll
        for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
      }
    }
    for (; k < k1; k += 8) {
      float f[EPV];
      Unpack<T>::run(ldg16(base + (int64_t)k * ldw), f);
#pragma unroll
      for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
    }
  }
#pragma unroll
  for (int e = 0; e < EPV; e++) s_m[ty][tx * EPV + e] = m[e];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * EPV; c += kThreads) {
    float r = s_m[0][c];
#pragma unroll
    for (int y = 1; y < 8; y++) r = fmaxf(r, s_m[y][c]);
    const int gc = blockIdx.x * 32 * EPV + c;
    if (gc < N && r >= 0.0f) atomicMax(reinterpret_cast<int *>(part) + gc, __float_as_int(r));
  }
}

// finalize only (plain op_absmax on a [K,N] matrix): Cw[j] from row 0 and part[j]
template <typename T>
__global__ void absmax_cols_finalize_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, int mode,
                                            const float *__restrict__ part, float *__restrict__ Cw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (j >= N) return;
  float c;
  if (fold_first(to_f32(W[j]), part[j], mode, c)) {
    for (int k = 1; k < K; k++) {
      const float x = to_f32(W[(int64_t)k * ldw + j]);
      if (x == x) { c = -x; break; }
    }
  }
  Cw[j] = c;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_cols_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode,
                  int rows_per_cta, const float *__restrict__ part, const float *__restrict__ sw_in,
                  int8_t *__restrict__ Wq, int64_t ldq, float *__restrict__ Cw) {
  constexpr int EPV = Unpack<T>::EPV;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * EPV;
  griddep_wait();
  if (col >= N) return;
  const T *base = W + col;
  float s[EPV];
  if (sw_in == nullptr) {
    float x0[EPV];
    Unpack<T>::run(ldg16(base), x0);
#pragma unroll
    for (int e = 0; e < EPV; e++) {
      float c;
      if (fold_first(x0[e], part[col + e], mode, c)) {
        for (int k = 1; k < K; k++) {
          const float x = to_f32(base[(int64_t)k * ldw + e]);
          if (x == x) { c = -x; break; }
        }
      }
      if (blockIdx.y == 0 && ty == 0 && Cw != nullptr) Cw[col + e] = c;
      s[e] = __fdiv_rn(range, c);
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPV; e++) s[e] = sw_in[col + e];
  }
  const int k0 = blockIdx.y * rows_per_cta;
  const int k1 = min(K, k0 + rows_per_cta);
  auto emit = [&](const uint4 &r, int k) {
    float f[EPV];
    Unpack<T>::run(r, f);
    uint32_t w[EPV / 4];
#pragma unroll
    for (int q = 0; q < EPV / 4; q++)
      w[q] = quant_code_u8(f[4 * q], s[4 * q]) | (quant_code_u8(f[4 * q + 1], s[4 * q + 1]) << 8) |
             (quant_code_u8(f[4 * q + 2], s[4 * q + 2]) << 16) | (quant_code_u8(f[4 * q + 3], s[4 * q + 3]) << 24);
    int8_t *dst = Wq + (int64_t)k * ldq + col;
    if (EPV == 4) *reinterpret_cast<uint32_t *>(dst) = w[0];
    else *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[EPV / 4 - 1]);
  };
  int k = k0 + ty;
  for (; k + 24 < k1; k += 32) {
    uint4 r[4];
#pragma unroll
    for (int u = 0; u < 4; u++) r[u] = ldg16(base + (int64_t)(k + 8 * u) * ldw);
#pragma unroll
    for (int u = 0; u < 4; u++) emit(r[u], k + 8 * u);
  }
  for (; k < k1; k += 8) emit(ldg16(base + (int64_t)k * ldw), k);
}

// ------------------------------------------------------------------------------------------
// Columns, ONE kernel and (for strips that stay in L2) one HBM read of W: a thread-block cluster
// owns a strip of 512 bytes of columns (128 fp32 / 256 half) over all K rows, rank r handling rows
// [r*rpc, (r+1)*rpc).  Phase 1 streams the strip and reduces |w| per column; the per-CTA maxima
// are combined across the cluster through distributed shared memory; phase 2 re-reads the strip
// (a few MB, read microseconds earlier by the same cluster: L2 hits) and writes the codes.
// kTranspose: codes go out as Wt[n][k] (K contiguous), the K-major operand layout the tensor-core
// GEMM runs fastest on -- staged through a swizzled smem tile so the stores are 128-byte rows.
// (A first design kept the strip in registers: one read, but the 128-byte strip rows it could
// afford used HBM badly -- 29 us vs 26 us for two passes at 4096^2 -- and it was dropped.)
// ------------------------------------------------------------------------------------------
template <typename T, bool kTranspose>
__global__ void __launch_bounds__(kThreads)
quant_cols_cluster_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode, int rpc,
                          int8_t *__restrict__ Wq, int64_t ldq, float *__restrict__ Cw) {
  namespace cg = cooperative_groups;
  constexpr int EPV = Unpack<T>::EPV;
  constexpr int SC = 32 * EPV;  // columns per strip
  __shared__ float s_red[kThreads / 32][SC];
  __shared__ float s_cmax[SC];   // this CTA's maxima over its rows (rows >= 1 only)
  __shared__ float s_x0[SC];     // row 0 of the strip (meaningful in cluster rank 0)
  __shared__ float s_final[SC];  // Cw of the strip
  __shared__ uint32_t s_tile[kTranspose ? SC * 32 : 1];  // [column][32 words = 128 k-bytes], word-swizzled
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned cs = cluster.num_blocks(), rank = cluster.block_rank();
  const int strip = blockIdx.x / cs;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = strip * SC + lane * EPV;
  const int r0 = (int)rank * rpc;
  const int r1 = min(K, r0 + rpc);
  const bool col_ok = col < N;
  const T *base = W + col;

  griddep_wait();
  // ---------------- phase 1: column maxima of this CTA's rows ----------------
  float m[EPV];
#pragma unroll
  for (int e = 0; e < EPV; e++) m[e] = -INFINITY;
  if (col_ok) {
    int r = r0 + warp;
    if (r0 == 0 && warp == 0 && r < r1) {  // row 0 is folded in signed, not by magnitude
      float f[EPV];
      Unpack<T>::run(ldg16(base), f);
#pragma unroll
      for (int e = 0; e < EPV; e++) s_x0[lane * EPV + e] = f[e];
      r += 8;
    }
    for (; r + 56 < r1; r += 64) {  // 8 independent 16-byte loads in flight per thread
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = ldg16(base + (int64_t)(r + 8 * u) * ldw);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        float f[EPV];
        Unpack<T>::run(v[u], f);
#pragma unroll
        for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
      }
    }
    for (; r < r1; r += 8) {
      float f[EPV];
      Unpack<T>::run(ldg16(base + (int64_t)r * ldw), f);
#pragma unroll
      for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
    }
  }
#pragma unroll
  for (int e = 0; e < EPV; e++) s_red[warp][lane * EPV + e] = m[e];
  __syncthreads();
  for (int c = threadIdx.x; c < SC; c += kThreads) {
    float r = s_red[0][c];
#pragma unroll
    for (int w = 1; w < kThreads / 32; w++) r = fmaxf(r, s_red[w][c]);
    s_cmax[c] = r;
  }
  cluster.sync();
  // ---------------- cluster-wide combine through DSMEM ----------------
  for (int c = threadIdx.x; c < SC; c += kThreads) {
    float mm = -INFINITY;
    for (unsigned rk = 0; rk < cs; rk++) mm = fmaxf(mm, cluster.map_shared_rank(s_cmax, rk)[c]);
    const int gc = strip * SC + c;
    float cw = 0.0f;
    if (gc < N) {
      const float x0 = cluster.map_shared_rank(s_x0, 0)[c];
      if (fold_first(x0, mm, mode, cw)) {
        for (int k = 1; k < K; k++) {  // rare +-0 tie-break: sign of the first later non-NaN zero
          const float x = to_f32(W[(int64_t)k * ldw + gc]);
          if (x == x) { cw = -x; break; }
        }
      }
      if (rank == 0 && Cw != nullptr) Cw[gc] = cw;
    }
    s_final[c] = cw;
  }
  cluster.sync();  // also keeps every CTA's smem alive until all remote reads are done
  griddep_launch_dependents();
  if (Wq == nullptr) return;
  // ---------------- phase 2: codes (strip re-read is served by L2) ----------------
  float s[EPV];
#pragma unroll
  for (int e = 0; e < EPV; e++) s[e] = __fdiv_rn(range, s_final[lane * EPV + e]);
  if (!kTranspose) {
    if (!col_ok) return;
    auto emit = [&](const uint4 &v, int r) {
      float f[EPV];
      Unpack<T>::run(v, f);
      uint32_t w[EPV / 4];
#pragma unroll
      for (int q = 0; q < EPV / 4; q++)
        w[q] = quant_code_u8(f[4 * q], s[4 * q]) | (quant_code_u8(f[4 * q + 1], s[4 * q + 1]) << 8) |
               (quant_code_u8(f[4 * q + 2], s[4 * q + 2]) << 16) | (quant_code_u8(f[4 * q + 3], s[4 * q + 3]) << 24);
      int8_t *dst = Wq + (int64_t)r * ldq + col;
      if (EPV == 4) *reinterpret_cast<uint32_t *>(dst) = w[0];
      else *reinterpret_cast<uint2 *>(dst) = make_uint2(w[0], w[EPV / 4 - 1]);
    };
    int r = r0 + warp;
    for (; r + 56 < r1; r += 64) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = ldg16(base + (int64_t)(r + 8 * u) * ldw);
#pragma unroll
      for (int u = 0; u < 8; u++) emit(v[u], r + 8 * u);
    }
    for (; r < r1; r += 8) emit(ldg16(base + (int64_t)r * ldw), r);
  } else {
    // chunks of 128 rows; warp w owns row quads w, w+8, w+16, w+24 of the chunk
    for (int rb = r0; rb < r1; rb += 128) {
      uint4 v[4][4];
#pragma unroll
      for (int qi = 0; qi < 4; qi++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int r = rb + 4 * (warp + 8 * qi) + i;
          v[qi][i] = (col_ok && r < r1) ? ldg16(base + (int64_t)r * ldw) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
      for (int qi = 0; qi < 4; qi++) {
        const int rq = warp + 8 * qi;
        float f[4][EPV];
#pragma unroll
        for (int i = 0; i < 4; i++) Unpack<T>::run(v[qi][i], f[i]);
#pragma unroll
        for (int e = 0; e < EPV; e++) {
          uint32_t w = 0;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int r = rb + 4 * rq + i;
            const uint32_t code = (r < r1) ? quant_code_u8(f[i][e], s[e]) : 0u;  // rows past K: zero padding
            w |= code << (8 * i);
          }
          // word index rotated by the writer's lane: conflict-free here and in the read-back below
          s_tile[(lane * EPV + e) * 32 + ((rq + lane) & 31)] = w;
        }
      }
      __syncthreads();
      for (int c = warp; c < SC; c += kThreads / 32) {  // one warp writes one 128-byte row of Wt
        const int gc = strip * SC + c;
        const int r = rb + 4 * lane;
        if (gc < N && r < r1)
          *reinterpret_cast<uint32_t *>(Wq + (int64_t)gc * ldq + r) = s_tile[c * 32 + ((lane + c / EPV) & 31)];
      }
      __syncthreads();
    }
  }
}

// Any N / alignment: one thread per column (coalesced across threads), scalar accesses.
template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_cols_generic_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode,
                          const float *__restrict__ sw_in, int8_t *__restrict__ Wq, int64_t ldq,
                          float *__restrict__ Cw, bool transpose) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (j >= N) return;
  float scale;
  if (sw_in == nullptr) {
    float m = -INFINITY;
    for (int k = 1; k < K; k++) m = fmaxf(m, fabsf(to_f32(W[(int64_t)k * ldw + j])));
    float c;
    if (fold_first(to_f32(W[j]), m, mode, c)) {
      for (int k = 1; k < K; k++) {
        const float x = to_f32(W[(int64_t)k * ldw + j]);
        if (x == x) { c = -x; break; }
      }
    }
    if (Cw != nullptr) Cw[j] = c;
    scale = __fdiv_rn(range, c);
  } else {
    scale = sw_in[j];
  }
  if (Wq == nullptr) return;
  for (int k = 0; k < K; k++) {
    const int8_t q = (int8_t)quant_code_u8(to_f32(W[(int64_t)k * ldw + j]), scale);
    if (transpose) Wq[(int64_t)j * ldq + k] = q;
    else Wq[(int64_t)k * ldq + j] = q;
  }
}

__global__ void inv_divide_kernel(const float *__restrict__ a, int64_t n, float b, float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (i < n) out[i] = __fdiv_rn(b, a[i]);
}

// AbsCompareLTEConstFunc, src/ops/op_elemwise.cuh:292-306
__global__ void outlier_mask_kernel(const float *__restrict__ A, int M, int K, int64_t lda, float thr,
                                    float *__restrict__ mask, int64_t ldm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (k >= K) return;
  for (int i = blockIdx.y; i < M; i += gridDim.y) {
    const float a = A[(int64_t)i * lda + k];
    const bool inl = ((a >= 0) & (a <= thr)) | ((a <= 0) & (-a <= thr));
    mask[(int64_t)i * ldm + k] = inl ? 0.0f : 1.0f;
  }
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- row launcher ----
template <typename T, int G, int NV>
void launch_rows(const T *X, int M, int K, int64_t ldx, float range, int mode, const float *sx, int8_t *Xq,
                 int64_t ldq, float *Cx, cudaStream_t st) {
  constexpr int RPB = kThreads / G;
  static int resident = 0;  // CTAs that fit on the device at once (per instantiation)
  if (resident == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quant_rows_kernel<T, G, NV>, kThreads, 0);
    resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
  }
  const int64_t nrb = ceil_div(M, RPB);
  const unsigned grid = (unsigned)(nrb < resident ? nrb : resident);
  launch_kernel(quant_rows_kernel<T, G, NV>, dim3(grid), dim3(kThreads), st, X, M, K, ldx, range, mode, sx, Xq, ldq, Cx);
}

template <typename T>
int rows_dispatch(const T *X, int M, int K, int64_t ldx, float range, int mode, const float *sx, int8_t *Xq,
                  int64_t ldq, float *Cx, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  const bool vec_ok = (K % EPV == 0) && aligned(X, 16) && ((ldx * sizeof(T)) % 16 == 0) &&
                      (Xq == nullptr || (aligned(Xq, EPV) && ldq % EPV == 0));
  if (!vec_ok) {
    return (int)launch_kernel(quant_rows_generic_kernel<T>, dim3((unsigned)ceil_div(M, kThreads / 32)), dim3(kThreads), st,
                              X, M, K, ldx, range, mode, sx, Xq, ldq, Cx);
  }
  const int nvec = K / EPV;
#define QG_ROWS(G, NV) launch_rows<T, G, NV>(X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, st)
  if (nvec <= 32) QG_ROWS(32, 1);
  else if (nvec <= 64) QG_ROWS(32, 2);
  else if (nvec <= 128) QG_ROWS(32, 4);
  else if (nvec <= 256) QG_ROWS(64, 4);
  else if (nvec <= 512) QG_ROWS(128, 4);
  else if (nvec <= 1024) QG_ROWS(256, 4);
  else if (nvec <= 2048) QG_ROWS(256, 8);
  else if (nvec <= 4096) QG_ROWS(256, 16);
  else QG_ROWS(256, 0);
#undef QG_ROWS
  return (int)cudaGetLastError();
}

// ---- column launcher ----
inline int cols_rows_per_cta(int K, int col_tiles) {
  // enough CTAs for ~4 waves of 148 SMs x 4 resident blocks, at least 32 rows per CTA
  int64_t want = ceil_div((int64_t)148 * 16, col_tiles);
  int64_t rows = ceil_div(K, want > 0 ? want : 1);
  rows = round_up(rows < 32 ? 32 : rows, 32);
  return (int)rows;
}

template <typename T>
int cols_dispatch(const T *W, int K, int N, int64_t ldw, float range, int mode, const float *sw, int8_t *Wq,
                  int64_t ldq, float *Cw, float *scratch, bool transpose, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  if (transpose && Wq != nullptr && !(aligned(Wq, 4) && ldq % 4 == 0))
    return (int)launch_kernel(quant_cols_generic_kernel<T>, dim3((unsigned)ceil_div(N, kThreads)), dim3(kThreads), st, W, K,
                              N, ldw, range, mode, sw, Wq, ldq, Cw, transpose);
  const bool vec_ok = (N % EPV == 0) && aligned(W, 16) && ((ldw * sizeof(T)) % 16 == 0) &&
                      (Wq == nullptr || (aligned(Wq, EPV) && ldq % EPV == 0)) &&
                      (sw != nullptr || scratch != nullptr || Wq != nullptr) && !(transpose && sw != nullptr);
  if (!vec_ok) {
    return (int)launch_kernel(quant_cols_generic_kernel<T>, dim3((unsigned)ceil_div(N, kThreads)), dim3(kThreads), st, W, K,
                              N, ldw, range, mode, sw, Wq, ldq, Cw, transpose);
  }
  // fused absmax + quantize: one cluster kernel (QG_COLS_TWO_PASS=1 keeps the three-launch path)
  if (sw == nullptr && Wq != nullptr && (transpose || getenv("QG_COLS_TWO_PASS") == nullptr)) {
    int cs = 1;
    while (cs < 8 && (int64_t)K > (int64_t)cs * 512) cs *= 2;      // >= ~512 rows per CTA, at most 8 CTAs
    const int rpc_c = (int)round_up(ceil_div(K, cs), 128);
    const int strips = (int)ceil_div(N, 32 * EPV);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(strips * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    count_launch();
    if (transpose)
      return (int)cudaLaunchKernelEx(&cfg, quant_cols_cluster_kernel<T, true>, W, K, N, ldw, range, mode, rpc_c, Wq, ldq, Cw);
    return (int)cudaLaunchKernelEx(&cfg, quant_cols_cluster_kernel<T, false>, W, K, N, ldw, range, mode, rpc_c, Wq, ldq, Cw);
  }
  const int col_tiles = (int)ceil_div(N, 32 * EPV);
  const int rpc = cols_rows_per_cta(K, col_tiles);
  if (sw == nullptr) {
    launch_kernel(fill_f32_kernel, dim3((unsigned)ceil_div(N, 256)), dim3(256), st, scratch, N, -INFINITY);
    if (K > 1) {
      dim3 grid(col_tiles, (unsigned)ceil_div(K - 1, rpc));
      launch_kernel(absmax_cols_partial_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, rpc, scratch);
    }
    if (Wq == nullptr) {
      return (int)launch_kernel(absmax_cols_finalize_kernel<T>, dim3((unsigned)ceil_div(N, 256)), dim3(256), st, W, K, N,
                                ldw, mode, scratch, Cw);
    }
  }
  dim3 grid(col_tiles, (unsigned)ceil_div(K, rpc));
  return (int)launch_kernel(quant_cols_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, range, mode, rpc, scratch, sw, Wq,
                            ldq, Cw);
}

}  // namespace

// ---- entry points used by capi.cu ----
int quant_rows(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, const float *sx,
               int8_t *Xq, int64_t ldq, float *Cx, cudaStream_t st) {
  switch (dtype) {
    case QG_F32: return rows_dispatch((const float *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, st);
    case QG_F16: return rows_dispatch((const __half *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, st);
    case QG_BF16: return rows_dispatch((const __nv_bfloat16 *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, st);
  }
  return QG_EINVAL;
}

// transpose: codes are written as Wt[n][k] with leading dimension ldq (K-major operand layout)
int quant_cols(const void *W, int dtype, int K, int N, int64_t ldw, float range, int mode, const float *sw,
               int8_t *Wq, int64_t ldq, float *Cw, float *scratch, bool transpose, cudaStream_t st) {
  switch (dtype) {
    case QG_F32: return cols_dispatch((const float *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, scratch, transpose, st);
    case QG_F16: return cols_dispatch((const __half *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, scratch, transpose, st);
    case QG_BF16:
      return cols_dispatch((const __nv_bfloat16 *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, scratch, transpose, st);
  }
  return QG_EINVAL;
}

int inv_divide(const float *a, int64_t n, float b, float *out, cudaStream_t st) {
  return (int)launch_kernel(inv_divide_kernel, dim3((unsigned)ceil_div(n, 256)), dim3(256), st, a, n, b, out);
}

int outlier_mask(const float *A, int M, int K, int64_t lda, float thr, float *mask, int64_t ldm, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(K, 256), (unsigned)(M < 32768 ? M : 32768));
  return (int)launch_kernel(outlier_mask_kernel, grid, dim3(256), st, A, M, K, lda, thr, mask, ldm);
}

}  // namespace qg
