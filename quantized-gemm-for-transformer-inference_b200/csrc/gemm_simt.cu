// gemm_simt.cu -- CUDA-core kernels that back the C ABI for shapes / layouts the tcgen05 path
// does not take (leading dimensions that are not multiples of 16 bytes, tiny matrices), the
// unfused dequantize, and the fp32 product used for the unquantized comparison.
//
// These are correctness paths, not the product's hot loop: the hot loop is gemm_i8_tc.cu.
#include "common.cuh"

namespace qg {

namespace {

constexpr int TM = 64, TN = 64, TK = 32;  // block tile; 256 threads, 4x4 outputs per thread

template <typename OutT>
__device__ __forceinline__ void store_out(void *O, int64_t ldo, int r, int c, float v) {
  reinterpret_cast<OutT *>(O)[(int64_t)r * ldo + c] = from_f32<OutT>(v);
}

// op_mm<int8_t,int> (src/ops/op_mm.cuh:9-46) with exact int32 accumulation; optional fused
// dequantize epilogue (kDequant) identical to the tcgen05 kernel's.
template <bool kDequant, typename OutT>
__global__ void __launch_bounds__(256)
gemm_s8_simt_kernel(const int8_t *__restrict__ A, int64_t lda, const int8_t *__restrict__ B, int64_t sb_k, int64_t sb_n,
                    int M, int N, int K, void *__restrict__ O, int64_t ldo, const float *__restrict__ Cx,
                    const float *__restrict__ Cw, const float *__restrict__ bias, float c, SideArgs side, int relu) {
  __shared__ int8_t sA[TM][TK + 4];
  __shared__ int8_t sB[TK][TN + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  griddep_wait();
  griddep_trigger_early();
  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0;

  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = e / TK, kk = e % TK;
      sA[r][kk] = (m0 + r < M && k0 + kk < K) ? A[(int64_t)(m0 + r) * lda + k0 + kk] : (int8_t)0;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      const int kk = e / TN, cidx = e % TN;
      sB[kk][cidx] = (k0 + kk < K && n0 + cidx < N) ? B[(int64_t)(k0 + kk) * sb_k + (int64_t)(n0 + cidx) * sb_n] : (int8_t)0;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < TK; kk++) {
      int a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sA[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      if (kDequant) {
        float v = dequant_ref(acc[i][j], Cx[r], Cw[col], c);
        if (side.no_pad > 0) {  // outlier side product, o ascending fp32 fma chain
          float sd = 0.0f;
          for (int o = 0; o < side.no_pad; o++) {
            float xo, wo;
            if (side.side_bf16) {
              xo = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(side.Xo)[(int64_t)r * side.ldxo + o]);
              wo = __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(side.Wo)[(int64_t)o * side.ldwo + col]);
            } else {
              xo = __half2float(reinterpret_cast<const __half *>(side.Xo)[(int64_t)r * side.ldxo + o]);
              wo = __half2float(reinterpret_cast<const __half *>(side.Wo)[(int64_t)o * side.ldwo + col]);
            }
            sd = __fmaf_rn(xo, wo, sd);
          }
          v = __fadd_rn(v, sd);
        }
        if (bias != nullptr) v = __fadd_rn(v, bias[col]);
        if (relu) v = v < 0.0f ? 0.0f : v;
        store_out<OutT>(O, ldo, r, col, v);
      } else {
        reinterpret_cast<int32_t *>(O)[(int64_t)r * ldo + col] = acc[i][j];
      }
    }
  }
}

// op_mm<float,float> (src/ops/op_mm.cuh:9-46): every output is the k-ascending chain
// res = fma(a, b, res) starting from +0, so any tiling gives bit-identical results; a partial last
// 32-wide tile adds fma(0,0,res) steps, which only turn -0 into +0.
__global__ void __launch_bounds__(256)
mm_f32_kernel(const float *__restrict__ A, int64_t sa_h, int64_t sa_w, const float *__restrict__ B, int64_t sb_h,
              int64_t sb_w, int M, int N, int K, float *__restrict__ C, int64_t ldc, MmBatch bt) {
  __shared__ float sA[TK][TM + 1];
  __shared__ float sB[TK][TN + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  {  // blockIdx.z = outer * n_inner + inner: one independent product per (outer, inner) pair
    const int zo = blockIdx.z / bt.n_inner, zi = blockIdx.z % bt.n_inner;
    A += zo * bt.a_outer + zi * bt.a_inner;
    B += zo * bt.b_outer + zi * bt.b_inner;
    C += zo * bt.c_outer + zi * bt.c_inner;
  }
  griddep_wait();
  griddep_trigger_early();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = e / TK, kk = e % TK;
      sA[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(int64_t)(m0 + r) * sa_h + (int64_t)(k0 + kk) * sa_w] : 0.0f;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      const int kk = e / TN, cidx = e % TN;
      sB[kk][cidx] = (k0 + kk < K && n0 + cidx < N) ? B[(int64_t)(k0 + kk) * sb_h + (int64_t)(n0 + cidx) * sb_w] : 0.0f;
    }
    __syncthreads();
    const int kmax = min(TK, K - k0);
    for (int kk = 0; kk < kmax; kk++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool pad = (K % 32) != 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int col = n0 + tx * 4 + j;
      if (col >= N) continue;
      C[(int64_t)r * ldc + col] = pad ? __fadd_rn(acc[i][j], 0.0f) : acc[i][j];
    }
  }
}

// a6+a7+a8(+a10) on stored accumulators: op_mm.cuh:96-99 / op_elemwise.cuh:93-103,118-129
template <typename OutT>
__global__ void __launch_bounds__(256)
dequantize_kernel(const int32_t *__restrict__ acc, int64_t ldacc, const float *__restrict__ Cx,
                  const float *__restrict__ Cw, const float *__restrict__ bias, int M, int N, float c,
                  OutT *__restrict__ O, int64_t ldo) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  if (col >= N) return;
  const float cw = Cw[col];
  const float b = bias ? bias[col] : 0.0f;
  for (int r = blockIdx.y; r < M; r += gridDim.y) {
    float v = dequant_ref(acc[(int64_t)r * ldacc + col], Cx[r], cw, c);
    if (bias != nullptr) v = __fadd_rn(v, b);
    O[(int64_t)r * ldo + col] = from_f32<OutT>(v);
  }
}

}  // namespace

// b_kmajor == 0: B is [K,N] with leading dimension ldb; 1: B is [N,K]
int gemm_s8_simt(const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int b_kmajor, int M, int N, int K, void *O,
                 int64_t ldo, int out_dtype, const float *Cx, const float *Cw, const float *bias, float c,
                 const SideArgs *side_in, cudaStream_t st, int act) {
  const int64_t sb_k = b_kmajor ? 1 : ldb, sb_n = b_kmajor ? ldb : 1;
  SideArgs side = {};
  if (side_in != nullptr) side = *side_in;
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM));
  switch (out_dtype) {
    case QG_S32:
      launch_kernel(gemm_s8_simt_kernel<false, float>, grid, dim3(256), st, A, lda, B, sb_k, sb_n, M, N, K, O, ldo, Cx, Cw, bias, c, side, act == QG_ACT_RELU ? 1 : 0);
      break;
    case QG_F32:
      launch_kernel(gemm_s8_simt_kernel<true, float>, grid, dim3(256), st, A, lda, B, sb_k, sb_n, M, N, K, O, ldo, Cx, Cw, bias, c, side, act == QG_ACT_RELU ? 1 : 0);
      break;
    case QG_F16:
      launch_kernel(gemm_s8_simt_kernel<true, __half>, grid, dim3(256), st, A, lda, B, sb_k, sb_n, M, N, K, O, ldo, Cx, Cw, bias, c, side, act == QG_ACT_RELU ? 1 : 0);
      break;
    case QG_BF16:
      launch_kernel(gemm_s8_simt_kernel<true, __nv_bfloat16>, grid, dim3(256), st, A, lda, B, sb_k, sb_n, M, N, K, O, ldo, Cx, Cw, bias, c, side, act == QG_ACT_RELU ? 1 : 0);
      break;
    default:
      return QG_EINVAL;
  }
  return (int)cudaGetLastError();
}

int mm_f32(const float *A, int64_t sa_h, int64_t sa_w, const float *B, int64_t sb_h, int64_t sb_w, int M, int N, int K,
           float *C, int64_t ldc, cudaStream_t st, const MmBatch *batch) {
  MmBatch bt = {};
  bt.n_outer = bt.n_inner = 1;
  if (batch != nullptr) bt = *batch;
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM), (unsigned)(bt.n_outer * bt.n_inner));
  launch_kernel(mm_f32_kernel, grid, dim3(256), st, A, sa_h, sa_w, B, sb_h, sb_w, M, N, K, C, ldc, bt);
  return (int)cudaGetLastError();
}

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

int add_layernorm_rows(const float *A, int64_t lda, const float *R, int64_t ldr, int M, int N, float *B, int64_t ldb,
                       cudaStream_t st) {
  if (M >= 148 * 128)
    launch_kernel(add_layernorm_rows_kernel<128, 64>, dim3((unsigned)ceil_div(M, 128)), dim3(kSmThreads), st, A, lda, R, ldr, M, N, B, ldb);
  else
    launch_kernel(add_layernorm_rows_kernel<32, 128>, dim3((unsigned)ceil_div(M, 32)), dim3(kSmThreads), st, A, lda, R, ldr, M, N, B, ldb);
  return (int)cudaGetLastError();
}

int softmax_rows(const float *A, int64_t lda, int M, int N, float scale, float *B, int64_t ldb, cudaStream_t st) {
  if (M >= 148 * 128)
    launch_kernel(softmax_rows_kernel<128, 64>, dim3((unsigned)ceil_div(M, 128)), dim3(kSmThreads), st, A, lda, M, N, scale, B, ldb);
  else
    launch_kernel(softmax_rows_kernel<32, 128>, dim3((unsigned)ceil_div(M, 32)), dim3(kSmThreads), st, A, lda, M, N, scale, B, ldb);
  return (int)cudaGetLastError();
}

int dequantize_s32(const int32_t *acc, int64_t ldacc, const float *Cx, const float *Cw, const float *bias, int M, int N,
                   float c, void *O, int out_dtype, int64_t ldo, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)(M < 8192 ? M : 8192));
  switch (out_dtype) {
    case QG_F32:
      launch_kernel(dequantize_kernel<float>, grid, dim3(256), st, acc, ldacc, Cx, Cw, bias, M, N, c, (float *)O, ldo);
      break;
    case QG_F16:
      launch_kernel(dequantize_kernel<__half>, grid, dim3(256), st, acc, ldacc, Cx, Cw, bias, M, N, c, (__half *)O, ldo);
      break;
    case QG_BF16:
      launch_kernel(dequantize_kernel<__nv_bfloat16>, grid, dim3(256), st, acc, ldacc, Cx, Cw, bias, M, N, c, (__nv_bfloat16 *)O, ldo);
      break;
    default:
      return QG_EINVAL;
  }
  return (int)cudaGetLastError();
}

}  // namespace qg
