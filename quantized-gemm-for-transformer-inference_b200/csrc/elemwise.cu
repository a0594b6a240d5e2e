// elemwise.cu -- the elementwise tail of the reference pipeline, op by op.
//
// op_quantized_mm's last three steps and the module layer's helpers are separate elementwise launches
// in the reference (src/ops/op_elemwise.cuh): op_dequantize (:614-625, DequantizeFunc :93-103),
// op_multiply(a, T b, out) (:644-654, MultiplyConstFunc :118-129), op_add (:501-512, AddFunc :57-65),
// op_subtract (:531-542), op_relu (:454-465, ReluFunc :181-195), all through
// op_elemwise_unary_kernel / op_elemwise_binary_w_bcast_kernel (:404-424, one element per thread, 32x32
// blocks).  The fused GEMM epilogue makes them unnecessary on the fast path; they exist so that the
// reference's own step-by-step sequence (src/timing_quantize.cu:38-58,67-70) can be re-pointed one call at
// a time, with identical results.  HBM-bound: 16-byte accesses when the operands allow, one rounding
// per operation exactly as the reference's functors do it.
#include "common.cuh"

namespace qg {

namespace {

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DEQ = 3, OP_SCALE = 4, OP_RELU = 5 };

template <int OP>
__device__ __forceinline__ float apply(float a, float b) {
  if (OP == OP_ADD) return __fadd_rn(a, b);
  if (OP == OP_SUB) return __fsub_rn(a, b);
  if (OP == OP_MUL || OP == OP_DEQ || OP == OP_SCALE) return __fmul_rn(a, b);
  return a < 0.0f ? 0.0f : a;  // ReluFunc: NaN and -0 pass through
}

// B broadcast: bmode 0 = same shape, 1 = [1,N] (row vector repeated down the rows), 2 = [M,1]
// (op_elemwise.cuh:410-421), 3 = none (unary / constant `c`).  A is int32 for OP_DEQ (cvt.rn.f32.s32 first).
template <int OP, bool VEC>
__global__ void __launch_bounds__(256)
elemwise_kernel(const void *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb, int bmode, float c,
                float *__restrict__ O, int64_t ldo, int M, int N) {
  constexpr int W = VEC ? 4 : 1;
  const int j = (blockIdx.x * 256 + threadIdx.x) * W;
  griddep_wait();
  if (j >= N) return;
  float bv[W];
  if (bmode == 1) {
#pragma unroll
    for (int e = 0; e < W; e++) bv[e] = B[j + e];
  }
  for (int i = blockIdx.y; i < M; i += gridDim.y) {
    float a[W], r[W];
    if (VEC) {
      const uint4 v = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(A) + (int64_t)i * lda + j);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; e++) a[e] = OP == OP_DEQ ? __int2float_rn((int)w[e]) : __uint_as_float(w[e]);
    } else {
      const uint32_t w = reinterpret_cast<const uint32_t *>(A)[(int64_t)i * lda + j];
      a[0] = OP == OP_DEQ ? __int2float_rn((int)w) : __uint_as_float(w);
    }
    if (bmode == 0) {
      if (VEC) {
        const float4 b4 = *reinterpret_cast<const float4 *>(B + (int64_t)i * ldb + j);
        bv[0] = b4.x; bv[1 % W] = b4.y; bv[2 % W] = b4.z; bv[3 % W] = b4.w;
      } else {
        bv[0] = B[(int64_t)i * ldb + j];
      }
    } else if (bmode == 2) {
      const float b = B[(int64_t)i * ldb];
#pragma unroll
      for (int e = 0; e < W; e++) bv[e] = b;
    } else if (bmode == 3) {
#pragma unroll
      for (int e = 0; e < W; e++) bv[e] = c;
    }
#pragma unroll
    for (int e = 0; e < W; e++) r[e] = apply<OP>(a[e], bv[e]);
    if (VEC) *reinterpret_cast<float4 *>(O + (int64_t)i * ldo + j) = make_float4(r[0], r[1 % W], r[2 % W], r[3 % W]);
    else O[(int64_t)i * ldo + j] = r[0];
  }
}

inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Reduce half of the row-parallel linear's reduce-scatter: the partial products of the P
// ranks sit in P slots of this GPU's memory (written there by the peers' GEMM epilogues); out = ((s_0 + s_1) + ...) + bias
// in ascending slot order, every addition rounded to fp32 -- a fixed order, so every run and every rank count gives the
// bits the CPU restatement gives.  The result goes to `out` and, when peers are given, to the same block of their
// matrices (the all-gather that completes an all-reduce, as plain 16-byte stores over NVLink).
struct ReducePeers {
  void *dst[kMaxExtraOut];
  int n;
  void *mc;  // != NULL (vector path only): the block's NVSwitch multicast address -- one multimem.st reaches every GPU, `out` included
};
template <typename PT, typename OT, bool VEC>
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const PT *__restrict__ slots, int64_t slot_stride, int n_slots, int64_t ld_part,
                       const float *__restrict__ bias, OT *__restrict__ out, ReducePeers peers, int64_t ldo, int M, int N) {
  constexpr int W = VEC ? 4 : 1;
  const int j = (blockIdx.x * 256 + threadIdx.x) * W;
  griddep_wait();
  if (j >= N) return;
  float bv[W];
#pragma unroll
  for (int e = 0; e < W; e++) bv[e] = bias != nullptr ? bias[j + e] : 0.0f;
  for (int i = blockIdx.y; i < M; i += gridDim.y) {
    float acc[W];
    for (int s = 0; s < n_slots; s++) {
      const PT *src = slots + (int64_t)s * slot_stride + (int64_t)i * ld_part + j;
      float v[W];
      if (VEC) {
        if (sizeof(PT) == 4) {
          const float4 f = *reinterpret_cast<const float4 *>(src);
          v[0] = f.x; v[1 % W] = f.y; v[2 % W] = f.z; v[3 % W] = f.w;
        } else {
          const uint2 u = *reinterpret_cast<const uint2 *>(src);
          const PT *h = reinterpret_cast<const PT *>(&u);
#pragma unroll
          for (int e = 0; e < W; e++) v[e] = to_f32<PT>(h[e]);
        }
      } else {
        v[0] = to_f32<PT>(src[0]);
      }
#pragma unroll
      for (int e = 0; e < W; e++) acc[e] = s == 0 ? v[e] : __fadd_rn(acc[e], v[e]);
    }
    if (bias != nullptr) {
#pragma unroll
      for (int e = 0; e < W; e++) acc[e] = __fadd_rn(acc[e], bv[e]);
    }
    OT r[W];
#pragma unroll
    for (int e = 0; e < W; e++) r[e] = from_f32<OT>(acc[e]);
    if (VEC && peers.mc != nullptr) {
      OT *dst = reinterpret_cast<OT *>(peers.mc) + (int64_t)i * ldo + j;
      const uint32_t *w = reinterpret_cast<const uint32_t *>(r);
      if (sizeof(OT) == 4) multimem_st128(dst, w[0], w[1 % W], w[2 % W], w[3 % W]);
      else asm volatile("multimem.st.weak.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(__uint_as_float(w[0])), "f"(__uint_as_float(w[1 % W])) : "memory");
      continue;
    }
    for (int d = -1; d < peers.n; d++) {
      OT *dst = (d < 0 ? out : reinterpret_cast<OT *>(peers.dst[d])) + (int64_t)i * ldo + j;
      if (VEC) {
        if (sizeof(OT) == 4) *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(r);
        else *reinterpret_cast<uint2 *>(dst) = *reinterpret_cast<const uint2 *>(r);
      } else {
        dst[0] = r[0];
      }
    }
  }
}

template <typename PT, typename OT>
int launch_reduce(const void *slots, int64_t slot_stride, int n_slots, int64_t ld_part, const float *bias, void *out,
                  const ReducePeers &peers, int64_t ldo, int M, int N, cudaStream_t st, int max_ctas) {
  bool vec = N % 4 == 0 && ld_part % 4 == 0 && slot_stride % 4 == 0 && ldo % 4 == 0 &&
             (reinterpret_cast<uintptr_t>(slots) % (4 * sizeof(PT))) == 0 && (reinterpret_cast<uintptr_t>(out) % (4 * sizeof(OT))) == 0;
  for (int d = 0; d < peers.n; d++) vec = vec && (reinterpret_cast<uintptr_t>(peers.dst[d]) % (4 * sizeof(OT))) == 0;
  ReducePeers pr = peers;
  if (!vec || (reinterpret_cast<uintptr_t>(pr.mc) % (4 * sizeof(OT))) != 0) pr.mc = nullptr;  // multicast needs the vector path
  const int w = vec ? 4 : 1;
  const unsigned gx = (unsigned)ceil_div(N, 256 * w);
  unsigned gy = (unsigned)((max_ctas > 0 ? max_ctas : 148 * 16) / gx);  // max_ctas: a caller running this under a GEMM keeps it small
  gy = gy < 1 ? 1 : gy;
  gy = gy > (unsigned)M ? (unsigned)M : gy;
  dim3 grid(gx, gy);
  if (vec)
    return (int)launch_kernel(reduce_partials_kernel<PT, OT, true>, grid, dim3(256), st, (const PT *)slots, slot_stride, n_slots,
                              ld_part, bias, (OT *)out, pr, ldo, M, N);
  return (int)launch_kernel(reduce_partials_kernel<PT, OT, false>, grid, dim3(256), st, (const PT *)slots, slot_stride, n_slots,
                            ld_part, bias, (OT *)out, pr, ldo, M, N);
}
template <typename PT>
int reduce_out(int out_dtype, const void *slots, int64_t slot_stride, int n_slots, int64_t ld_part, const float *bias, void *out,
               const ReducePeers &peers, int64_t ldo, int M, int N, cudaStream_t st, int max_ctas) {
  switch (out_dtype) {
    case QG_F32: return launch_reduce<PT, float>(slots, slot_stride, n_slots, ld_part, bias, out, peers, ldo, M, N, st, max_ctas);
    case QG_F16: return launch_reduce<PT, __half>(slots, slot_stride, n_slots, ld_part, bias, out, peers, ldo, M, N, st, max_ctas);
    case QG_BF16: return launch_reduce<PT, __nv_bfloat16>(slots, slot_stride, n_slots, ld_part, bias, out, peers, ldo, M, N, st, max_ctas);
  }
  return QG_EINVAL;
}

template <int OP>
int launch_op(const void *A, int64_t lda, const float *B, int64_t ldb, int bmode, float c, float *O, int64_t ldo, int M, int N,
              cudaStream_t st) {
  const bool vec = N % 4 == 0 && al16(A) && al16(O) && lda % 4 == 0 && ldo % 4 == 0 &&
                   (bmode != 0 || (al16(B) && ldb % 4 == 0));
  const int w = vec ? 4 : 1;
  const unsigned gx = (unsigned)ceil_div(N, 256 * w);
  // enough row groups to fill the machine several times over, every thread walking rows with the same columns
  unsigned gy = (unsigned)(148 * 16 / gx);
  gy = gy < 1 ? 1 : gy;
  gy = gy > (unsigned)M ? (unsigned)M : gy;
  dim3 grid(gx, gy);
  if (vec) return (int)launch_kernel(elemwise_kernel<OP, true>, grid, dim3(256), st, A, lda, B, ldb, bmode, c, O, ldo, M, N);
  return (int)launch_kernel(elemwise_kernel<OP, false>, grid, dim3(256), st, A, lda, B, ldb, bmode, c, O, ldo, M, N);
}

}  // namespace

// op: 0 add, 1 subtract, 2 multiply (all with the reference's broadcast rule), 3 dequantize (A int32),
// 4 multiply by the constant c, 5 relu
int elemwise(int op, const void *A, int64_t lda, const float *B, int64_t ldb, int bmode, float c, float *O, int64_t ldo, int M,
             int N, cudaStream_t st) {
  switch (op) {
    case OP_ADD: return launch_op<OP_ADD>(A, lda, B, ldb, bmode, c, O, ldo, M, N, st);
    case OP_SUB: return launch_op<OP_SUB>(A, lda, B, ldb, bmode, c, O, ldo, M, N, st);
    case OP_MUL: return launch_op<OP_MUL>(A, lda, B, ldb, bmode, c, O, ldo, M, N, st);
    case OP_DEQ: return launch_op<OP_DEQ>(A, lda, B, ldb, bmode, c, O, ldo, M, N, st);
    case OP_SCALE: return launch_op<OP_SCALE>(A, lda, nullptr, 0, 3, c, O, ldo, M, N, st);
    case OP_RELU: return launch_op<OP_RELU>(A, lda, nullptr, 0, 3, 0.0f, O, ldo, M, N, st);
  }
  return QG_EINVAL;
}

// slots: [n_slots][M][ld_part] of part_dtype (slot_stride elements apart); out / peers: [M, ldo] of out_dtype
int reduce_partials(const void *slots, int64_t slot_stride, int n_slots, int part_dtype, int64_t ld_part, const float *bias,
                    void *out, void *const *peers, int n_peers, int64_t ldo, int out_dtype, int M, int N, cudaStream_t st,
                    void *out_mc, int max_ctas) {
  ReducePeers pr = {};
  pr.mc = out_mc;
  pr.n = n_peers;
  for (int d = 0; d < n_peers; d++) pr.dst[d] = peers[d];
  switch (part_dtype) {
    case QG_F32: return reduce_out<float>(out_dtype, slots, slot_stride, n_slots, ld_part, bias, out, pr, ldo, M, N, st, max_ctas);
    case QG_F16: return reduce_out<__half>(out_dtype, slots, slot_stride, n_slots, ld_part, bias, out, pr, ldo, M, N, st, max_ctas);
    case QG_BF16: return reduce_out<__nv_bfloat16>(out_dtype, slots, slot_stride, n_slots, ld_part, bias, out, pr, ldo, M, N, st, max_ctas);
  }
  return QG_EINVAL;
}

}  // namespace qg
