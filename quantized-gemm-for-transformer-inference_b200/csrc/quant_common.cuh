// quant_common.cuh -- helpers shared by the quantizer kernels (quantize.cu, outlier.cu).
#pragma once

#include "common.cuh"

namespace qg {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg16(const void *p) {
  return __ldg(reinterpret_cast<const uint4 *>(p));
}

// 16-byte read-only load carrying an L2 eviction policy (createpolicy.fractional.*)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg16_hint(const void *p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  return v;
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

}  // namespace
}  // namespace qg
