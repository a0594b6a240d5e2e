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

// ---- masking row quantizer --------------------------------------------------------------------
struct RowOutlier {
  const uint32_t *mask;  // K bits
  const int *wbase;      // prefix popcount per mask word
  void *Xo;              // [M, ldxo] 16-bit side operand, columns = rank of k inside O
  int64_t ldxo;
  int side_bf16;         // 0: fp16, 1: bf16
  const int *count;      // number of outlier columns (device): the quantizer zeroes columns [count, ldxo) of every row it writes
};

// the padding columns of a row of the side operand (the epilogue multiplies them with zero rows of Wo: they must not be NaN)
__device__ __forceinline__ void zero_side_tail(int row, const RowOutlier &ro) {
  uint16_t *x = reinterpret_cast<uint16_t *>(ro.Xo) + (int64_t)row * ro.ldxo;
  for (int p = __ldg(ro.count); p < (int)ro.ldxo; p++) x[p] = 0;
}

// Removes the outlier elements of one 16-byte vector (vector index vi of `row`) from the int8 path
// (sets them to +0) and, when `write`, stores them into the side operand.
// bits: the vector's EPV mask bits; pos: rank of its first outlier column among all outlier columns
template <typename T>
__device__ __forceinline__ void strip_bits(uint4 &r, uint32_t bits, int pos, int row, const RowOutlier &ro, bool write) {
  constexpr int EPV = Unpack<T>::EPV;
  if (bits == 0) return;
  float f[EPV];
  Unpack<T>::run(r, f);
  uint32_t *w = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
  for (int e = 0; e < EPV; e++) {
    if ((bits >> e) & 1u) {
      if (write) {
        if (ro.side_bf16) reinterpret_cast<__nv_bfloat16 *>(ro.Xo)[(int64_t)row * ro.ldxo + pos] = __float2bfloat16_rn(f[e]);
        else reinterpret_cast<__half *>(ro.Xo)[(int64_t)row * ro.ldxo + pos] = __float2half_rn(f[e]);
      }
      pos++;
      if (sizeof(T) == 4) w[e] = 0u;
      else w[e >> 1] &= (e & 1) ? 0x0000ffffu : 0xffff0000u;
    }
  }
}
// mask bits / rank of vector vi (they depend on the column only, not on the row)
template <typename T>
__device__ __forceinline__ void outlier_bits(int vi, const RowOutlier &ro, uint32_t &bits, int &pos) {
  constexpr int EPV = Unpack<T>::EPV;
  const int c0 = vi * EPV, b0 = c0 & 31;
  const uint32_t word = __ldg(ro.mask + (c0 >> 5));
  bits = (word >> b0) & ((1u << EPV) - 1u);
  pos = bits ? __ldg(ro.wbase + (c0 >> 5)) + __popc(word & ((1u << b0) - 1u)) : 0;
}
template <typename T>
__device__ __forceinline__ void strip_outliers(uint4 &r, int vi, int row, const RowOutlier &ro, bool write) {
  uint32_t bits;
  int pos;
  outlier_bits<T>(vi, ro, bits, pos);
  strip_bits<T>(r, bits, pos, row, ro, write);
}


// first_rb / rb_stride: the row blocks this CTA walks (blockIdx.x / gridDim.x when the kernel is on its own; a sub-range
// of the grid when it shares a launch with the column quantizer's second pass).  STREAM: X is loaded with an L2
// evict-first hint (it is read once, and must not push out the weight matrix the other half of the launch re-reads).
template <typename T, int G, int NV, bool STREAM = false, bool OUTL = false>
__device__ __forceinline__ void quant_rows_body(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                                                const float *__restrict__ sx_in, int8_t *__restrict__ Xq, int64_t ldq,
                                                float *__restrict__ Cx, RowMaxIo io, int first_rb, int rb_stride,
                                                RowOutlier ro = RowOutlier()) {
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

  const uint64_t pol_stream = STREAM ? l2_policy_evict_first() : 0;
  auto load = [&](uint4 (&dst)[NVC], int rb) {
    const int row = rb * RPB + rib;
    const T *xr = X + (int64_t)(row < M ? row : 0) * ldx;
#pragma unroll
    for (int v = 0; v < NVC; v++) {
      const int idx = v * G + g;
      if (row < M && idx < nvec) dst[v] = STREAM ? ldg16_hint(xr + (int64_t)idx * EPV, pol_stream) : ldg16(xr + (int64_t)idx * EPV);
      else dst[v] = make_uint4(0, 0, 0, 0);
    }
  };

  uint4 raw[NVC], nxt[NVC];
  // a thread's vectors sit in the same columns for every row: their mask bits and ranks are fetched once
  uint32_t obits[OUTL ? NVC : 1];
  int opos[OUTL ? NVC : 1];
  if (OUTL && NV > 0) {
#pragma unroll
    for (int v = 0; v < NVC; v++) {
      obits[v] = 0;
      opos[v] = 0;
      if (v * G + g < nvec) outlier_bits<T>(v * G + g, ro, obits[v], opos[v]);
    }
  }
  int rb = first_rb;
  if (NV > 0 && rb < nrb) load(raw, rb);
  for (int it = 0; rb < nrb; rb += rb_stride, it++) {
    if (kPrefetch && rb + rb_stride < nrb) load(nxt, rb + rb_stride);
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
            if (OUTL && active) strip_bits<T>(raw[v], obits[v], opos[v], row, ro, true);  // outlier entries leave the int8 path here
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
          uint4 rv = ldg16(xr + (int64_t)idx * EPV);
          if (OUTL) strip_outliers<T>(rv, idx, row, ro, true);
          Unpack<T>::run(rv, f);
          if (idx == 0) x0 = f[0];
#pragma unroll
          for (int e = 0; e < EPV; e++)
            if (e > 0 || idx > 0) m = fmaxf(m, fabsf(f[e]));
        }
      }
      if (OUTL && active && g == 0) zero_side_tail(row, ro);
      m = warp_max(m);
      if (io.m_in != nullptr) m = active ? io.m_in[row] : -INFINITY;  // the producer's epilogue already reduced columns 1..K-1
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
      if (active && g == 0) {
        if (io.m_out != nullptr) io.m_out[row] = m;
        if (io.init_out != nullptr) io.init_out[row] = -INFINITY;
      }
      float c;
      if (fold_first(x0, m, mode, c) && active) {
        for (int j = 1; j < K; j++) {  // rare: sign of the first later zero decides (+-0 tie-break); outlier columns read as +0
          const bool outl = OUTL && ((__ldg(ro.mask + (j >> 5)) >> (j & 31)) & 1u);
          const float xj = outl ? 0.0f : to_f32(xr[j]);
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
        for (int idx = g; idx < nvec; idx += G) {
          uint4 rv = ldg16(xr + (int64_t)idx * EPV);
          if (OUTL) strip_outliers<T>(rv, idx, row, ro, false);
          emit(rv, idx);
        }
      }
    }
    if (kPrefetch) {
#pragma unroll
      for (int v = 0; v < NVC; v++) raw[v] = nxt[v];
    } else if (NV > 0 && rb + rb_stride < nrb) {
      load(raw, rb + rb_stride);
    }
  }
}


}  // namespace
}  // namespace qg
