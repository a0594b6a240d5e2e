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
    // consecutive threads walk whichever index is contiguous in memory (a transposed view, e.g. the
    // K^T of attention.cuh:58-60, has its unit stride along k)
    for (int e = threadIdx.x; e < TM * TK; e += 256) {
      const int r = sa_w == 1 ? e / TK : e % TM, kk = sa_w == 1 ? e % TK : e / TM;
      sA[kk][r] = (m0 + r < M && k0 + kk < K) ? A[(int64_t)(m0 + r) * sa_h + (int64_t)(k0 + kk) * sa_w] : 0.0f;
    }
    for (int e = threadIdx.x; e < TK * TN; e += 256) {
      const int kk = sb_w == 1 ? e / TN : e % TK, cidx = sb_w == 1 ? e % TN : e / TK;
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


// The same product on 128 x 128 block tiles, 8 x 8 outputs per thread, with the next k-tile's global
// loads issued into registers before the current tile is multiplied.  The 64 x 64 kernel above
// exposes one DRAM round trip per k-tile and spends half its issue slots on shared-memory loads
// (ncu: long-scoreboard stalls at the tile stores, short-scoreboard at the FFMAs; 46 us for the 256
// [128 x 64] x [64 x 128] score products of an attention call).  Same k-ascending fma chain per output.
constexpr int BM2 = 128, BN2 = 128, BK2 = 16;
__global__ void __launch_bounds__(256)
mm_f32_128_kernel(const float *__restrict__ A, int64_t sa_h, int64_t sa_w, const float *__restrict__ B, int64_t sb_h,
                  int64_t sb_w, int M, int N, int K, float *__restrict__ C, int64_t ldc, MmBatch bt) {
  __shared__ float sA[BK2][BM2 + 1];
  __shared__ float sB[BK2][BN2 + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * BM2, n0 = blockIdx.x * BN2;
  {
    const int zo = blockIdx.z / bt.n_inner, zi = blockIdx.z % bt.n_inner;
    A += zo * bt.a_outer + zi * bt.a_inner;
    B += zo * bt.b_outer + zi * bt.b_inner;
    C += zo * bt.c_outer + zi * bt.c_inner;
  }
  griddep_wait();
  griddep_trigger_early();
  constexpr int PER = BM2 * BK2 / 256;  // 8 elements of each operand tile per thread
  float ra[PER], rb[PER];
  // consecutive threads walk whichever index is contiguous in memory
  auto fetch = [&](int k0) {
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int e = threadIdx.x + u * 256;
      const int r = sa_w == 1 ? e / BK2 : e % BM2, kk = sa_w == 1 ? e % BK2 : e / BM2;
      ra[u] = (m0 + r < M && k0 + kk < K) ? A[(int64_t)(m0 + r) * sa_h + (int64_t)(k0 + kk) * sa_w] : 0.0f;
      const int kb = sb_w == 1 ? e / BN2 : e % BK2, c = sb_w == 1 ? e % BN2 : e / BK2;
      rb[u] = (k0 + kb < K && n0 + c < N) ? B[(int64_t)(k0 + kb) * sb_h + (int64_t)(n0 + c) * sb_w] : 0.0f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int u = 0; u < PER; u++) {
      const int e = threadIdx.x + u * 256;
      const int r = sa_w == 1 ? e / BK2 : e % BM2, kk = sa_w == 1 ? e % BK2 : e / BM2;
      sA[kk][r] = ra[u];
      const int kb = sb_w == 1 ? e / BN2 : e % BK2, c = sb_w == 1 ? e % BN2 : e / BK2;
      sB[kb][c] = rb[u];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += BK2) {
    stash();
    __syncthreads();
    if (k0 + BK2 < K) fetch(k0 + BK2);
    const int kmax = min(BK2, K - k0);
    for (int kk = 0; kk < kmax; kk++) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = sA[kk][ty + 16 * i];  // rows ty, ty+16, ...: conflict-free across ty
#pragma unroll
      for (int j = 0; j < 8; j++) b[j] = sB[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool pad = (K % 32) != 0;  // the reference's 32-wide k-tile, zero padded: fma(0,0,res) turns -0 into +0
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int r = m0 + ty + 16 * i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int col = n0 + tx + 16 * j;
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

// Second half of a split-K product: adds the `slices` int32 partial matrices (exact, any order) and runs
// the same epilogue as the fused GEMM -- dequantize, bias, ReLU, cast -- or stores the int32 sum (kRaw).
template <bool kRaw, typename OutT>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const int32_t *__restrict__ parts, int64_t slice_stride, int slices, int64_t ldp,
                     const float *__restrict__ Cx, const float *__restrict__ Cw, const float *__restrict__ bias, int M, int N,
                     float c, int relu, void *__restrict__ O, int64_t ldo) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  if (col >= N) return;
  const float cw = kRaw ? 0.0f : Cw[col];
  const float b = (!kRaw && bias) ? bias[col] : 0.0f;
  for (int r = blockIdx.y; r < M; r += gridDim.y) {
    int acc = 0;
    for (int sl = 0; sl < slices; sl++) acc += parts[sl * slice_stride + (int64_t)r * ldp + col];
    if (kRaw) {
      reinterpret_cast<int32_t *>(O)[(int64_t)r * ldo + col] = acc;
    } else {
      float v = dequant_ref(acc, Cx[r], cw, c);
      if (bias != nullptr) v = __fadd_rn(v, b);
      if (relu) v = v < 0.0f ? 0.0f : v;
      store_out<OutT>(O, ldo, r, col, v);
    }
  }
}

}  // namespace

int splitk_reduce(const int32_t *parts, int64_t slice_stride, int slices, int64_t ldp, const float *Cx, const float *Cw,
                  const float *bias, int M, int N, float c, int act, void *O, int out_dtype, int64_t ldo, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(N, 256), (unsigned)(M < 4096 ? M : 4096));
  const int relu = act == QG_ACT_RELU ? 1 : 0;
#define QG_SPLITK(RAW, T) launch_kernel(splitk_reduce_kernel<RAW, T>, grid, dim3(256), st, parts, slice_stride, slices, ldp, Cx, Cw, bias, M, N, c, relu, O, ldo)
  switch (out_dtype) {
    case QG_S32: QG_SPLITK(true, float); break;
    case QG_F32: QG_SPLITK(false, float); break;
    case QG_F16: QG_SPLITK(false, __half); break;
    case QG_BF16: QG_SPLITK(false, __nv_bfloat16); break;
    default: return QG_EINVAL;
  }
#undef QG_SPLITK
  return (int)cudaGetLastError();
}

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
  if (M >= 96 && N >= 96) {
    dim3 grid((unsigned)ceil_div(N, BN2), (unsigned)ceil_div(M, BM2), (unsigned)(bt.n_outer * bt.n_inner));
    launch_kernel(mm_f32_128_kernel, grid, dim3(256), st, A, sa_h, sa_w, B, sb_h, sb_w, M, N, K, C, ldc, bt);
    return (int)cudaGetLastError();
  }
  dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(M, TM), (unsigned)(bt.n_outer * bt.n_inner));
  launch_kernel(mm_f32_kernel, grid, dim3(256), st, A, sa_h, sa_w, B, sb_h, sb_w, M, N, K, C, ldc, bt);
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
