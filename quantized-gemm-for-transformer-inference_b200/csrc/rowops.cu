// rowops.cu -- row-wise fp32 helpers of the attention / encoder-decoder widening (DESIGN.md sections 10-11):
// softmax(scale * A) and ADD & NORM with the reference kernels' arithmetic (src/ops/op_softmax.cuh,
// src/ops/op_layernorm.cuh).  Both are defined by ascending-order running sums per row, which is what
// the kernel structures below are built around.  Not on the quantized hot path.
#include "quant_common.cuh"

namespace qg {

// op_multiply(A, scale, T) + op_softmax(T, B) (attention.cuh:62-68; op_softmax.cuh:6-29) in one pass
// over the same arithmetic: t_j = fl(a_j * scale); max by strict '>' starting from column 0;
// e_j = expf(t_j - max); the sum runs over ascending j; b_j = e_j / sum.  One thread per row, as in
// the reference, because the ascending-order fp32 sum is part of the result; a CTA's rows are
// staged through shared memory in 32-column tiles so that global accesses stay coalesced.
// 128 threads stage ROWS x TW tiles (every load of a tile in flight at once: the tile loop is a chain of
// global-latency round trips, so wide tiles matter); the first ROWS threads own one row each.
// ROWS = 32 spreads a short matrix over four times as many SMs (the per-row work is serial either way).
constexpr int kSmThreads = 128;
template <int ROWS, int TW>
__global__ void __launch_bounds__(kSmThreads)
softmax_rows_kernel(const float *A, int64_t lda, int M, int N, float scale, float *B, int64_t ldb) {  // A may alias B
  __shared__ float tile[ROWS][TW + 1];
  const int r0 = blockIdx.x * ROWS;
  const int rows = min(ROWS, M - r0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();
  griddep_trigger_early();
  // tile <- fl(A[r0.., c0..c0+TW-1] * scale): warp w loads rows w, w+4, ..., 32 consecutive floats per load
  auto load_tile = [&](const float *src, int64_t ld, int c0, bool mul) {
    for (int r = warp; r < rows; r += kSmThreads / 32) {
#pragma unroll
      for (int cc = lane; cc < TW; cc += 32) {
        const int c = c0 + cc;
        float v = 0.0f;
        if (c < N) v = src[(int64_t)(r0 + r) * ld + c];
        tile[r][cc] = mul ? __fmul_rn(v, scale) : v;
      }
    }
  };
  auto store_tile = [&](int c0) {
    for (int r = warp; r < rows; r += kSmThreads / 32) {
#pragma unroll
      for (int cc = lane; cc < TW; cc += 32) {
        const int c = c0 + cc;
        if (c < N) B[(int64_t)(r0 + r) * ldb + c] = tile[r][cc];
      }
    }
  };
  const int t = threadIdx.x;
  float mx = 0.0f, sum = 0.0f;
  for (int c0 = 0; c0 < N; c0 += TW) {  // pass 1: row max
    __syncthreads();
    load_tile(A, lda, c0, true);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) {
        const float v = tile[t][j];
        if (c0 + j == 0) mx = v;
        else if (v > mx) mx = v;
      }
    }
  }
  for (int c0 = 0; c0 < N; c0 += TW) {  // pass 2: e_j, running sum; e_j parked in B
    __syncthreads();
    load_tile(A, lda, c0, true);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) {
        const float e = expf(__fsub_rn(tile[t][j], mx));
        tile[t][j] = e;
        sum = __fadd_rn(sum, e);
      }
    }
    __syncthreads();
    store_tile(c0);
  }
  for (int c0 = 0; c0 < N; c0 += TW) {  // pass 3: divide
    __syncthreads();
    load_tile(B, ldb, c0, false);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) tile[t][j] = __fdiv_rn(tile[t][j], sum);
    }
    __syncthreads();
    store_tile(c0);
  }
}

// op_add(A, R, T) + op_layernorm(T, B): the "ADD & NORM" of src/transformer.cu:57-58,74-75, arithmetic of
// AddFunc (op_elemwise.cuh:57-65) and layernorm_kernel (src/ops/op_layernorm.cuh:6-32):
//   t_j = fl(a_j + r_j);  mean = (sum of t_j, ascending j, from 0) / w;  var = (sum of pow(t_j - mean, 2)) / w;
//   b_j = (t_j - mean) / var            -- divides by the variance, not its square root, and has no epsilon
// One thread per row (ascending-order sums), rows staged through shared memory like the softmax.
template <int ROWS, int TW>
__global__ void __launch_bounds__(kSmThreads)
add_layernorm_rows_kernel(const float *A, int64_t lda, const float *R, int64_t ldr, int M, int N, float *B, int64_t ldb) {
  __shared__ float tile[ROWS][TW + 1];
  const int r0 = blockIdx.x * ROWS;
  const int rows = min(ROWS, M - r0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  griddep_wait();
  griddep_trigger_early();
  auto load_tile = [&](int c0) {
    for (int r = warp; r < rows; r += kSmThreads / 32) {
#pragma unroll
      for (int cc = lane; cc < TW; cc += 32) {
        const int c = c0 + cc;
        float v = 0.0f;
        if (c < N) {
          v = A[(int64_t)(r0 + r) * lda + c];
          if (R != nullptr) v = __fadd_rn(v, R[(int64_t)(r0 + r) * ldr + c]);
        }
        tile[r][cc] = v;
      }
    }
  };
  const int t = threadIdx.x;
  const int w = N;
  float mean = 0.0;
  float var = 0.0;
  for (int c0 = 0; c0 < N; c0 += TW) {
    __syncthreads();
    load_tile(c0);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) mean += tile[t][j];
    }
  }
  mean = mean / w;
  for (int c0 = 0; c0 < N; c0 += TW) {
    __syncthreads();
    load_tile(c0);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) {
        // the reference writes `var += pow(x - mean, 2)`: float base, int exponent -> the double overload,
        // so every step adds an exact double square and rounds the running sum back to float
        const double dd = (double)(tile[t][j] - mean);
        var = (float)((double)var + dd * dd);
      }
    }
  }
  var = var / w;
  // A (and R) are read for the last time in this pass, tile by tile, before the same tile of B is
  // written: B may alias A or R (the reference normalises in place)
  for (int c0 = 0; c0 < N; c0 += TW) {
    __syncthreads();
    load_tile(c0);
    __syncthreads();
    if (t < rows) {
      const int n = min(TW, N - c0);
      for (int j = 0; j < n; j++) tile[t][j] = (tile[t][j] - mean) / var;
    }
    __syncthreads();
    for (int r = warp; r < rows; r += kSmThreads / 32) {
#pragma unroll
      for (int cc = lane; cc < TW; cc += 32) {
        const int c = c0 + cc;
        if (c < N) B[(int64_t)(r0 + r) * ldb + c] = tile[r][cc];
      }
    }
  }
}


// ---- warp-per-row forms (rows up to kWarpRowMaxN columns) ----------------------------------------------
// The order-dependent part of both ops is one fp32 (softmax) or fp32/fp64 (layernorm) running sum per
// row; everything else -- loads, expf, squares, divisions, stores -- is independent per element.  A warp
// owns a row: all lanes do the independent work on a shared-memory copy of the row, lane 0 walks the
// sums in ascending column order.  4096 x 512 ADD & NORM: 73 us with one thread per row -> see DESIGN.md.
constexpr int kRowWarps = 4;
constexpr int kWarpRowMaxN = 4096;

__device__ __forceinline__ float warp_max_f32(float m) {  // fmaxf drops NaNs
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  return m;
}

__global__ void __launch_bounds__(kRowWarps * 32)
softmax_warp_rows_kernel(const float *A, int64_t lda, int M, int N, float scale, float *B, int64_t ldb) {  // A may alias B
  extern __shared__ float sm_rows[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float *row = sm_rows + (size_t)warp * N;
  griddep_wait();
  griddep_trigger_early();
  for (int r = blockIdx.x * kRowWarps + warp; r < M; r += gridDim.x * kRowWarps) {
    const float *a = A + (int64_t)r * lda;
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) {
      const float t = __fmul_rn(a[j], scale);
      row[j] = t;
      m = fmaxf(m, t);  // NaNs are skipped, as `t > max` skips them
    }
    m = warp_max_f32(m);
    __syncwarp();
    const float t0 = row[0];
    const float mx = (t0 != t0) ? t0 : m;  // max starts AT column 0: a NaN there is never replaced
    for (int j = lane; j < N; j += 32) row[j] = expf(__fsub_rn(row[j], mx));
    __syncwarp();
    float sum = 0.0f;
    if (lane == 0) {
#pragma unroll 8
      for (int j = 0; j < N; j++) sum = __fadd_rn(sum, row[j]);
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    float *b = B + (int64_t)r * ldb;
    for (int j = lane; j < N; j += 32) b[j] = __fdiv_rn(row[j], sum);
    __syncwarp();
  }
}

// ADD & NORM.  The reference's `var += pow(x - mean, 2)` is, per element, v <- RN32(v + d*d) with d*d
// exact: a double add stored back to float.  Done literally that is F2F.F64 -> DADD -> F2F.F32 in a
// dependent chain, ~176 cycles per element on this part (47 us for 4096 x 512).  Here the running
// sum stays in a double register and the store-to-float is done on its bit pattern (round to nearest
// even at bit 29: add 0x0fffffff + lsb, clear the low 29 bits), which leaves DADD + four integer
// instructions on the chain.  Valid while every partial sum is a normal float: rows with a non-zero
// square below 2^-100, a square above 1e30, or inf / NaN take the literal chain.  (An fp32-only
// chain -- error-free square + round-to-odd 3-sum, Boldo & Melquiond 2008 -- was bit-exact too but
// slower: 68 us.)
// Serial sums go one row per THREAD (warp 0 of the CTA owns up to 32 rows) while all 128 threads do
// the independent work -- loads, a + r, exact double squares, the final divisions, stores -- on a
// shared-memory copy of the CTA's rows.  Rows per CTA shrink with N so that rows + squares fit.
constexpr int kLnThreads = 256;

// QUANT: the normalised row is also absmax-quantized for the linear layer that consumes it (SURVEY.md section 8f rank 3:
// "fuse the preceding add & norm with the row quantizer"): Cx[i] = max(b[i,0], max_{j>=1} |b[i,j]|) and
// Xq[i,j] = low8(trunc(b[i,j] * (range / Cx[i]))), the arithmetic of quant_rows_kernel, from the shared-memory copy.
struct RowQuantOut {
  int8_t *Xq;
  int64_t ldq;
  float *Cx;
  float range;
  int mode;
};

template <bool QUANT>
__global__ void __launch_bounds__(kLnThreads)
add_layernorm_cta_rows_kernel(const float *A, int64_t lda, const float *R, int64_t ldr, int M, int N, int rows_per_cta,
                              float *B, int64_t ldb, RowQuantOut qo) {  // B may alias A or R
  extern __shared__ double sm_rows_d[];
  const int ldsq = N + 1, ldrow = N + 1;  // odd strides: 32 threads walking 32 rows hit 32 banks
  double *sq = sm_rows_d;                                                   // [rows_per_cta][N+1] exact squares
  __shared__ int s_slow[32];
  float *row = reinterpret_cast<float *>(sm_rows_d + (size_t)rows_per_cta * ldsq);  // [rows_per_cta][N+1] t = a + r
  __shared__ float s_mean[32], s_var[32];
  const int t = threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  const int w = N;
  for (int r0 = blockIdx.x * rows_per_cta; r0 < M; r0 += gridDim.x * rows_per_cta) {
    const int rows = min(rows_per_cta, M - r0);
    if (t < 32) s_slow[t] = 0;
    // Element-parallel phases (this load, the squares, the final division) walk the CTA's rows four columns at a time when the
    // operands allow 16-byte accesses: one index division per four elements, four times the bytes in flight.  (ncu, 4096 x 512:
    // the scalar form of this load alone was 31 % of the kernel's samples -- 4-byte loads, 16 KB in flight per SM.)
    const bool vec4 = (N & 3) == 0 && (lda & 3) == 0 && (ldb & 3) == 0 && (R == nullptr || (ldr & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(R)) & 15) == 0;
    const int n4 = N >> 2;
    if (vec4) {
      for (int e0 = t; e0 < rows * n4; e0 += kLnThreads * 4) {
        float4 va[4], vr[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int e = e0 + u * kLnThreads;
          const int r = e / n4, j = (e - r * n4) << 2;
          va[u] = vr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (e < rows * n4) {
            va[u] = *reinterpret_cast<const float4 *>(A + (int64_t)(r0 + r) * lda + j);
            if (R != nullptr) vr[u] = *reinterpret_cast<const float4 *>(R + (int64_t)(r0 + r) * ldr + j);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int e = e0 + u * kLnThreads;
          const int r = e / n4, j = (e - r * n4) << 2;
          if (e < rows * n4) {
            float *dst = row + r * ldrow + j;
            dst[0] = R != nullptr ? __fadd_rn(va[u].x, vr[u].x) : va[u].x;
            dst[1] = R != nullptr ? __fadd_rn(va[u].y, vr[u].y) : va[u].y;
            dst[2] = R != nullptr ? __fadd_rn(va[u].z, vr[u].z) : va[u].z;
            dst[3] = R != nullptr ? __fadd_rn(va[u].w, vr[u].w) : va[u].w;
          }
        }
      }
    }
    // eight independent loads per thread in flight (a plain loop here is a chain of DRAM round trips:
    // that, not the serial sums, was most of the first version's 73 us)
    for (int e0 = t; !vec4 && e0 < rows * N; e0 += kLnThreads * 4) {
      float va[4], vr[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * kLnThreads;
        const int r = e / N, j = e - r * N;
        va[u] = vr[u] = 0.0f;
        if (e < rows * N) {
          va[u] = A[(int64_t)(r0 + r) * lda + j];
          if (R != nullptr) vr[u] = R[(int64_t)(r0 + r) * ldr + j];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int e = e0 + u * kLnThreads;
        const int r = e / N, j = e - r * N;
        if (e < rows * N) row[r * ldrow + j] = R != nullptr ? __fadd_rn(va[u], vr[u]) : va[u];
      }
    }
    __syncthreads();
    if (t < rows) {
      float mean = 0.0f;
      const float *x = row + t * ldrow;
#pragma unroll 8
      for (int j = 0; j < N; j++) mean += x[j];
      s_mean[t] = mean / w;
    }
    __syncthreads();
    for (int e = t; e < rows * n4 && vec4; e += kLnThreads) {
      const int r = e / n4, j = (e - r * n4) << 2;
      const float mean = s_mean[r];
      bool slow = false;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double dd = (double)(row[r * ldrow + j + c] - mean);
        const double q = dd * dd;  // exact: 24-bit x 24-bit
        sq[r * ldsq + j + c] = q;
        slow |= (q != 0.0 && q < 7.888609052210118e-31) || !(q < 1e30);
      }
      if (slow) s_slow[r] = 1;
    }
    for (int e = t; e < rows * N && !vec4; e += kLnThreads) {
      const int r = e / N, j = e - r * N;
      const double dd = (double)(row[r * ldrow + j] - s_mean[r]);
      const double q = dd * dd;  // exact: 24-bit x 24-bit
      sq[r * ldsq + j] = q;
      // the bit-pattern rounding needs every partial sum to be zero or a normal, finite float
      if ((q != 0.0 && q < 7.888609052210118e-31) || !(q < 1e30)) s_slow[r] = 1;  // 2^-100
    }
    __syncthreads();
    if (t < rows) {
      float var = 0.0f;
      const double *q = sq + t * ldsq;
      if (!s_slow[t] && N <= 4096) {
        double v = 0.0;  // always exactly a float value
#pragma unroll 8
        for (int j = 0; j < N; j++) {
          long long b = __double_as_longlong(v + q[j]);
          b += 0x0fffffffLL + ((b >> 29) & 1);
          v = __longlong_as_double(b & ~0x1fffffffLL);
        }
        var = (float)v;
      } else {  // `var += pow(x - mean, 2)` literally: double add, float store
        const float *x = row + t * ldrow;
        const float mean = s_mean[t];
        for (int j = 0; j < N; j++) {
          const double dd = (double)(x[j] - mean);
          var = (float)((double)var + dd * dd);
        }
      }
      s_var[t] = var / w;
    }
    __syncthreads();
    for (int e = t; e < rows * n4 && vec4; e += kLnThreads) {
      const int r = e / n4, j = (e - r * n4) << 2;
      const float mean = s_mean[r], var = s_var[r];
      float *src = row + r * ldrow + j;
      float4 v;
      v.x = (src[0] - mean) / var; v.y = (src[1] - mean) / var;
      v.z = (src[2] - mean) / var; v.w = (src[3] - mean) / var;
      *reinterpret_cast<float4 *>(B + (int64_t)(r0 + r) * ldb + j) = v;
      if (QUANT) { src[0] = v.x; src[1] = v.y; src[2] = v.z; src[3] = v.w; }
    }
    for (int e = t; e < rows * N && !vec4; e += kLnThreads) {
      const int r = e / N, j = e - r * N;
      const float v = (row[r * ldrow + j] - s_mean[r]) / s_var[r];
      B[(int64_t)(r0 + r) * ldb + j] = v;
      if (QUANT) row[r * ldrow + j] = v;
    }
    __syncthreads();
    if (QUANT) {  // a warp per row: reduce, fold the signed first element, emit packed codes
      const int lane = t & 31, warp = t >> 5;
      for (int r = warp; r < rows; r += kLnThreads / 32) {
        const float *x = row + r * ldrow;
        float m = -INFINITY;
        for (int j = 1 + lane; j < N; j += 32) m = fmaxf(m, fabsf(x[j]));
        m = warp_max(m);
        float c;
        if (fold_first(x[0], m, qo.mode, c)) {
          for (int j = 1; j < N; j++)
            if (x[j] == x[j]) { c = -x[j]; break; }
        }
        if (lane == 0) qo.Cx[r0 + r] = c;
        const float scale = __fdiv_rn(qo.range, c);
        int8_t *q = qo.Xq + (int64_t)(r0 + r) * qo.ldq;
        if (((reinterpret_cast<uintptr_t>(q) | (uintptr_t)qo.ldq) & 3) == 0) {
          for (int j = 4 * lane; j < N; j += 128) {
            uint32_t wv = 0;
#pragma unroll
            for (int e = 0; e < 4; e++)
              if (j + e < N) wv |= quant_code_u8(x[j + e], scale) << (8 * e);
            if (j + 4 <= N) *reinterpret_cast<uint32_t *>(q + j) = wv;
            else for (int e = 0; j + e < N; e++) q[j + e] = (int8_t)((wv >> (8 * e)) & 0xffu);
          }
        } else {
          for (int j = lane; j < N; j += 32) q[j] = (int8_t)quant_code_u8(x[j], scale);
        }
      }
      __syncthreads();
    }
  }
}

// dynamic shared memory above 48 KB needs the opt-in once per kernel
template <typename... KArgs, typename... Args>
static void warp_rows_launch(void (*kern)(KArgs...), size_t smem, int M, cudaStream_t st, Args &&...args) {
  static bool opted[kMaxDevices] = {};  // per kernel (template instantiation of this launcher) and device
  if (smem > (48 << 10)) smem_optin(kern, 200 << 10, opted);
  const int64_t ctas = ceil_div(M, kRowWarps);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(ctas < 148 * 16 ? ctas : 148 * 16));
  cfg.blockDim = dim3(kRowWarps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// quant != nullptr: also emit int8 codes + Cx of the result (returns QG_ENOTSUP when this width has no fused form, so that
// the caller can run the row quantizer separately)
int add_layernorm_rows(const float *A, int64_t lda, const float *R, int64_t ldr, int M, int N, float *B, int64_t ldb,
                       cudaStream_t st, int8_t *Xq, int64_t ldq, float *Cx, float range, int mode) {
  const bool quant = Xq != nullptr;
  RowQuantOut qo = {Xq, ldq, Cx, range, mode};
  if (quant && N > kWarpRowMaxN) return QG_ENOTSUP;
  if (N <= kWarpRowMaxN) {
    // rows + squares of a CTA's rows in at most 192 KB of shared memory, at most 32 rows (one warp of serial sums)
    int rpc = (int)((192 << 10) / ((size_t)(N + 1) * 12));
    rpc = rpc > 32 ? 32 : rpc;
    while (rpc > 1 && ceil_div(M, rpc) < 148) rpc >>= 1;  // short matrices: spread over the SMs first
    const size_t smem = (size_t)rpc * (N + 1) * 12;
    static bool opted[kMaxDevices] = {}, opted_q[kMaxDevices] = {};
    if (quant) smem_optin(add_layernorm_cta_rows_kernel<true>, 200 << 10, opted_q);
    else smem_optin(add_layernorm_cta_rows_kernel<false>, 200 << 10, opted);
    const int64_t ctas = ceil_div(M, rpc);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(ctas < 148 * 8 ? ctas : 148 * 8));
    cfg.blockDim = dim3(kLnThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    count_launch();
    if (quant) cudaLaunchKernelEx(&cfg, add_layernorm_cta_rows_kernel<true>, A, lda, R, ldr, M, N, rpc, B, ldb, qo);
    else cudaLaunchKernelEx(&cfg, add_layernorm_cta_rows_kernel<false>, A, lda, R, ldr, M, N, rpc, B, ldb, qo);
    return (int)cudaGetLastError();
  }
  if (M >= 148 * 128)
    launch_kernel(add_layernorm_rows_kernel<128, 64>, dim3((unsigned)ceil_div(M, 128)), dim3(kSmThreads), st, A, lda, R, ldr, M, N, B, ldb);
  else
    launch_kernel(add_layernorm_rows_kernel<32, 128>, dim3((unsigned)ceil_div(M, 32)), dim3(kSmThreads), st, A, lda, R, ldr, M, N, B, ldb);
  return (int)cudaGetLastError();
}

int softmax_rows(const float *A, int64_t lda, int M, int N, float scale, float *B, int64_t ldb, cudaStream_t st) {
  if (N <= kWarpRowMaxN) {
    warp_rows_launch(softmax_warp_rows_kernel, (size_t)kRowWarps * N * 4, M, st, A, lda, M, N, scale, B, ldb);
    return (int)cudaGetLastError();
  }
  if (M >= 148 * 128)
    launch_kernel(softmax_rows_kernel<128, 64>, dim3((unsigned)ceil_div(M, 128)), dim3(kSmThreads), st, A, lda, M, N, scale, B, ldb);
  else
    launch_kernel(softmax_rows_kernel<32, 128>, dim3((unsigned)ceil_div(M, 32)), dim3(kSmThreads), st, A, lda, M, N, scale, B, ldb);
  return (int)cudaGetLastError();
}

}  // namespace qg
