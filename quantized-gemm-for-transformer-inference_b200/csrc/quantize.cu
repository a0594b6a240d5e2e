// quantize.cu -- vector-wise absmax quantizers.
//
// Replaces, per matrix, the reference's three launches
//   op_absmax        (src/ops/op_reduction.cuh:195-204; one thread per vector, serial loop)
//   op_inv_divide    (src/ops/op_elemwise.cuh:657-667)
//   op_multiply<T,int8_t> (src/ops/op_elemwise.cuh:629-640; one element per thread, byte stores)
// with HBM-bound kernels that read the matrix with 16-byte loads, reduce with warp shuffles and
// write packed int8 codes plus the fp32 absmax.  The arithmetic (and its quirks: signed first
// element, IEEE 127/x, truncate-and-wrap cast) is reproduced exactly; see oracle/qoracle.c.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <type_traits>
#include <utility>

#include "quant_common.cuh"

namespace qg {

namespace {

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
                  float *__restrict__ Cx, RowMaxIo io) {
  griddep_wait();
  griddep_trigger_early();
  quant_rows_body<T, G, NV>(X, M, K, ldx, range, mode, sx_in, Xq, ldq, Cx, io, (int)blockIdx.x, (int)gridDim.x);
}

// Any K, any alignment: one warp per row, scalar accesses.  Same arithmetic.
template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_rows_generic_kernel(const T *__restrict__ X, int M, int K, int64_t ldx, float range, int mode,
                          const float *__restrict__ sx_in, int8_t *__restrict__ Xq, int64_t ldq,
                          float *__restrict__ Cx, RowMaxIo io) {
  const int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  griddep_wait();
  griddep_trigger_early();
  if (row >= M) return;
  const T *xr = X + (int64_t)row * ldx;
  float scale;
  if (sx_in == nullptr) {
    float m = -INFINITY;
    for (int j = lane; j < K; j += 32)
      if (j > 0) m = fmaxf(m, fabsf(to_f32(xr[j])));
    m = warp_max(m);
    if (io.m_in != nullptr) m = io.m_in[row];
    if (lane == 0) {
      if (io.m_out != nullptr) io.m_out[row] = m;
      if (io.init_out != nullptr) io.init_out[row] = -INFINITY;
    }
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
// part[j] = (epoch << 32) | fp32 bits of the running maximum; every call uses a fresh epoch, so stale
// entries lose every atomicMax and nothing has to be initialised between calls (twopass_scratch clears
// the buffer when the epoch wraps)
__device__ __forceinline__ float part_value(unsigned long long v, uint32_t epoch) {
  return (uint32_t)(v >> 32) == epoch ? __uint_as_float((uint32_t)v) : -INFINITY;
}

template <typename T>
__device__ __forceinline__ void absmax_cols_partial_body(const T *__restrict__ W, int K, int N, int64_t ldw, int rows_per_cta,
                                                         unsigned long long *__restrict__ part, uint32_t epoch, int col_tile,
                                                         int row_chunk, float (*s_m)[32 * Unpack<T>::EPV + 1]) {
  constexpr int EPV = Unpack<T>::EPV;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (col_tile * 32 + tx) * EPV;
  const int k0 = 1 + row_chunk * rows_per_cta;
  const int k1 = min(K, k0 + rows_per_cta);
#ifndef QG_COLS_NO_L2_HINTS
  const uint64_t pol = l2_policy_evict_last();
#define QG_LD1(p) ldg16_hint(p, pol)
#else
#define QG_LD1(p) ldg16(p)
#endif
  float m[EPV];
#pragma unroll
  for (int e = 0; e < EPV; e++) m[e] = -INFINITY;
  if (col < N) {
    const T *base = W + col;
    int k = k0 + ty;
    for (; k + 24 < k1; k += 32) {  // 4 independent 16-byte loads in flight per thread
      uint4 r[4];
#pragma unroll
      for (int u = 0; u < 4; u++) r[u] = QG_LD1(base + (int64_t)(k + 8 * u) * ldw);
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
      Unpack<T>::run(QG_LD1(base + (int64_t)k * ldw), f);
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
    const int gc = col_tile * 32 * EPV + c;
    if (gc < N && r >= 0.0f) atomicMax(part + gc, ((unsigned long long)epoch << 32) | __float_as_uint(r));
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
absmax_cols_partial_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, int rows_per_cta,
                           unsigned long long *__restrict__ part, uint32_t epoch) {
  __shared__ float s_m[8][32 * Unpack<T>::EPV + 1];
  griddep_wait();
  griddep_trigger_early();
  absmax_cols_partial_body<T>(W, K, N, ldw, rows_per_cta, part, epoch, (int)blockIdx.x, (int)blockIdx.y, s_m);
}

// finalize only (plain op_absmax on a [K,N] matrix): Cw[j] from row 0 and part[j]
template <typename T>
__global__ void absmax_cols_finalize_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, int mode,
                                            const unsigned long long *__restrict__ part, uint32_t epoch,
                                            float *__restrict__ Cw) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  if (j >= N) return;
  float c;
  if (fold_first(to_f32(W[j]), part_value(part[j], epoch), mode, c)) {
    for (int k = 1; k < K; k++) {
      const float x = to_f32(W[(int64_t)k * ldw + j]);
      if (x == x) { c = -x; break; }
    }
  }
  Cw[j] = c;
}

template <typename T, int UL = 4>  // UL: independent 16-byte loads in flight per thread
__device__ __forceinline__ void quant_cols_body(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode,
                                                int rows_per_cta, const unsigned long long *__restrict__ part, uint32_t epoch,
                                                const float *__restrict__ sw_in, int8_t *__restrict__ Wq, int64_t ldq,
                                                float *__restrict__ Cw, int col_tile, int row_chunk) {
  constexpr int EPV = Unpack<T>::EPV;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (col_tile * 32 + tx) * EPV;
  if (col >= N) return;
  const T *base = W + col;
  float s[EPV];
  if (sw_in == nullptr) {
    float x0[EPV];
    Unpack<T>::run(ldg16(base), x0);
#pragma unroll
    for (int e = 0; e < EPV; e++) {
      float c;
      if (fold_first(x0[e], part_value(part[col + e], epoch), mode, c)) {
        for (int k = 1; k < K; k++) {
          const float x = to_f32(base[(int64_t)k * ldw + e]);
          if (x == x) { c = -x; break; }
        }
      }
      if (row_chunk == 0 && ty == 0 && Cw != nullptr) Cw[col + e] = c;
      s[e] = __fdiv_rn(range, c);
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPV; e++) s[e] = sw_in[col + e];
  }
  const int k0 = row_chunk * rows_per_cta;
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
#ifndef QG_COLS_NO_L2_HINTS
  const uint64_t pol2 = l2_policy_evict_first();
#define QG_LD2(p) ldg16_hint(p, pol2)
#else
#define QG_LD2(p) ldg16(p)
#endif
  // Bottom-up (the rows pass 1 read last first), with L1::no_allocate / L2 evict-first loads (pass 1: evict-last).  In the
  // pipeline this pass IS served from L2 (tools/l2_reuse_probe.py: 14.3 us right after pass 1 against 20.5 us cold at 64 MiB);
  // round 1's "no hits" was ncu's kernel replay, which saves and restores memory between passes and evicts everything.
  int k = k1 - 1 - ty;
  for (; k - 8 * (UL - 1) >= k0; k -= 8 * UL) {
    uint4 r[UL];
#pragma unroll
    for (int u = 0; u < UL; u++) r[u] = QG_LD2(base + (int64_t)(k - 8 * u) * ldw);
#pragma unroll
    for (int u = 0; u < UL; u++) emit(r[u], k - 8 * u);
  }
  for (; k >= k0; k -= 8) emit(QG_LD2(base + (int64_t)k * ldw), k);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_cols_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode,
                  int rows_per_cta, const unsigned long long *__restrict__ part, uint32_t epoch,
                  const float *__restrict__ sw_in, int8_t *__restrict__ Wq, int64_t ldq, float *__restrict__ Cw) {
  griddep_wait();
  griddep_trigger_early();
  // the GEMM that follows may set itself up (barriers, TMEM, descriptors) while this pass streams:
  // a consistent 0.4-1.2 us per call (93.4 vs 93.8 us at 4096^3, same box).  The same trigger in
  // pass 1 or in every kernel is a loss (common.cuh: griddep_trigger_early).
  griddep_launch_dependents();
  quant_cols_body<T>(W, K, N, ldw, range, mode, rows_per_cta, part, epoch, sw_in, Wq, ldq, Cw, (int)blockIdx.x, (int)blockIdx.y);
}

// Wavefront form of the two passes: the columns are cut into slabs, and launch i runs pass 1 over slab i TOGETHER with
// pass 2 over slab i-1 (blockIdx.z picks the role).  Every CTA streams freely -- the only dependency is on the previous
// launch -- and pass 2 finds its slab where the previous launch's pass 1 left it, in L2.
template <typename T>
__global__ void __launch_bounds__(kThreads)
cols_wave_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode, int rows_per_cta1,
                 int rows_per_cta2, unsigned long long *__restrict__ part, uint32_t epoch, int8_t *__restrict__ Wq, int64_t ldq,
                 float *__restrict__ Cw, int tile1_first, int tile1_count, int tile2_first, int tile2_count, int chunks1,
                 int chunks2, int last) {
  __shared__ float s_m[8][32 * Unpack<T>::EPV + 1];
  griddep_wait();
  griddep_trigger_early();
  if (last) griddep_launch_dependents();  // the GEMM's set-up may overlap the final slab's codes
  if (blockIdx.z == 0) {
    if ((int)blockIdx.x < tile1_count && (int)blockIdx.y < chunks1)
      absmax_cols_partial_body<T>(W, K, N, ldw, rows_per_cta1, part, epoch, tile1_first + (int)blockIdx.x, (int)blockIdx.y, s_m);
  } else {
    if ((int)blockIdx.x < tile2_count && (int)blockIdx.y < chunks2)
      quant_cols_body<T>(W, K, N, ldw, range, mode, rows_per_cta2, part, epoch, nullptr, Wq, ldq, Cw,
                         tile2_first + (int)blockIdx.x, (int)blockIdx.y);
  }
}

// Pass 2 with transposed output: codes go out as Wt[n][k] (K contiguous), the K-major operand layout
// the tensor-core GEMM runs fastest on.  A CTA walks its row chunk in patches of 32 rows (one row
// quad per warp, next patch prefetched into registers), packs four k-consecutive codes per column
// into a word, stages them in a bank-rotated smem tile and writes 32-byte sectors of Wt; the
// neighbouring row chunks, which run concurrently, complete the 128-byte lines in L2.
//
// Single-launch designs were built and measured at 4096^2 fp32 before settling on two launches
// (26 us row-major, ~29 us transposed): a thread-block-cluster kernel holding a 128-byte-wide strip
// in registers with a DSMEM max exchange (one HBM read, but narrow rows: 29 us); a cluster kernel
// re-reading a 512-byte strip through L2 (34-40 us, too few CTAs in flight); and a "decoupled"
// kernel whose phase-2 blocks wait on per-panel arrival counters (ncu: 67 MB of DRAM reads, i.e. the
// re-read did hit L2, yet 35-48 us in three variants -- fence, atomic and wait latency per patch).
template <typename T>
__global__ void __launch_bounds__(kThreads)
quant_cols_t_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode, int rows_per_cta,
                    const unsigned long long *__restrict__ part, uint32_t epoch, int8_t *__restrict__ Wt, int64_t ldt,
                    float *__restrict__ Cw) {
  constexpr int EPV = Unpack<T>::EPV;
  constexpr int SC = 32 * EPV;  // columns per CTA
  __shared__ uint32_t s_buf[(kThreads / 32) * SC];  // [8 k-words][SC columns]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + lane) * EPV;
  const bool col_ok = col < N;
  const T *base = W + col;
  griddep_wait();
  griddep_trigger_early();
  float s[EPV];
  if (col_ok) {
    float x0[EPV];
    Unpack<T>::run(ldg16(base), x0);
#pragma unroll
    for (int e = 0; e < EPV; e++) {
      float c;
      if (fold_first(x0[e], part_value(part[col + e], epoch), mode, c)) {
        for (int k = 1; k < K; k++) {
          const float x = to_f32(base[(int64_t)k * ldw + e]);
          if (x == x) { c = -x; break; }
        }
      }
      if (blockIdx.y == 0 && warp == 0 && Cw != nullptr) Cw[col + e] = c;
      s[e] = __fdiv_rn(range, c);
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPV; e++) s[e] = 0.0f;
  }
  const int k0 = blockIdx.y * rows_per_cta, k1 = min(K, k0 + rows_per_cta);
  auto load = [&](int rb, uint4 (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = rb + 4 * warp + i;
      v[i] = (col_ok && row < k1) ? ldg16(base + (int64_t)row * ldw) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 v[4], vn[4];
  // bottom-up over the chunk's 32-row patches (see quant_cols_kernel: the rows pass 1 read last are
  // the ones still in L2)
  const int rb_last = k0 + ((k1 - k0 - 1) / 32) * 32;
  load(rb_last, v);
  for (int rb = rb_last; rb >= k0; rb -= 32) {
    if (rb - 32 >= k0) load(rb - 32, vn);
    float f[4][EPV];
#pragma unroll
    for (int i = 0; i < 4; i++) Unpack<T>::run(v[i], f[i]);
#pragma unroll
    for (int e = 0; e < EPV; e++) {
      uint32_t w = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint32_t code = (rb + 4 * warp + i < k1) ? quant_code_u8(f[i][e], s[e]) : 0u;  // rows past K: zeros
        w |= code << (8 * i);
      }
      // a lane's EPV columns are rotated by lane / (32 / EPV): 32 lanes (same e) hit 32 different banks
      s_buf[warp * SC + lane * EPV + ((e + lane / (32 / EPV)) & (EPV - 1))] = w;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SC * 2; t += kThreads) {  // a thread moves 16 of a column's 32 bytes
      const int h = t / SC, c = t - h * SC;
      const int gc = blockIdx.x * SC + c;
      const int wl = c / EPV;
      const int pos = wl * EPV + (((c & (EPV - 1)) + wl / (32 / EPV)) & (EPV - 1));
      uint32_t w4[4];
#pragma unroll
      for (int j = 0; j < 4; j++) w4[j] = s_buf[(4 * h + j) * SC + pos];
      const int row = rb + 16 * h;
      if (gc < N && row < k1)
        *reinterpret_cast<uint4 *>(Wt + (int64_t)gc * ldt + row) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; i++) v[i] = vn[i];
  }
}

#ifndef QG_FUSED_MIN_CTAS
#define QG_FUSED_MIN_CTAS 5
#endif
#ifndef QG_FUSED_COLS_LOADS
#define QG_FUSED_COLS_LOADS 4
#endif
// The op's two quantizers share one launch where they complement each other (qg_quantized_mm): the column quantizer's
// second pass re-reads W out of L2 and is bound by L2 throughput, the row quantizer streams X from HBM and is bound by
// HBM -- run side by side (CTAs [0, cols_ctas) take pass-2 tiles, the rest walk the rows of X) they overlap instead of
// queueing: pass 1 -> [pass 2 || rows] -> GEMM.  No dependency between the two halves; X is loaded evict-first so that
// it does not push W out of L2 before pass 2 has read it.
template <typename T, int G, int NV>
__global__ void __launch_bounds__(kThreads, (NV <= 4 ? QG_FUSED_MIN_CTAS : 2))
quant_cols2_rows_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode, int rows_per_cta,
                        const unsigned long long *__restrict__ part, uint32_t epoch, int8_t *__restrict__ Wq, int64_t ldq,
                        float *__restrict__ Cw, int col_tiles, int cols_ctas, const T *__restrict__ X, int M, int64_t ldx,
                        int8_t *__restrict__ Xq, int64_t ldxq, float *__restrict__ Cx) {
  griddep_wait();
  griddep_trigger_early();
  if ((int)blockIdx.x < cols_ctas) {
    griddep_launch_dependents();
    quant_cols_body<T, QG_FUSED_COLS_LOADS>(W, K, N, ldw, range, mode, rows_per_cta, part, epoch, nullptr, Wq, ldq, Cw,
                                            (int)blockIdx.x % col_tiles, (int)blockIdx.x / col_tiles);
  } else {
    quant_rows_body<T, G, NV, true>(X, M, K, ldx, range, mode, nullptr, Xq, ldxq, Cx, RowMaxIo(), (int)blockIdx.x - cols_ctas,
                                    (int)gridDim.x - cols_ctas);
  }
}

// ------------------------------------------------------------------------------------------
// Columns in ONE launch, one HBM read of W: the panel pipeline.
//
// W is cut into column panels small enough to stay in L2 (<= ~16 MiB).  A column group (32*EPV columns) goes
// through two phases -- phase 1 reduces |W[k,j]| (k >= 1) per column, phase 2 folds the signed row 0, forms the
// scale and writes the codes.  The persistent grid is split by role: "readers" (most CTAs) walk the phase-1 items
// of all panels in panel order and never wait, so HBM streams W once at full rate; "writers" walk the phase-2
// items in the same order, about one panel behind, and find their rows in L2.  Work is handed out per WARP with a
// STATIC round-robin and no block-level barrier exists in the loop: an item is one column group x R rows, streamed
// with 8 independent 16-byte loads in flight per lane.  A writer item waits until all phase-1 items of ITS column
// group have arrived (one counter per group, each on its own cache line).
// What this design replaced, measured (4096^2 fp32, two launches = 23.5 us): whole CTAs synchronising per item
// (38-48 us) and per-warp items claimed from a ticket counter (50 us) -- both read W from HBM once, as intended
// (ncu: 67 MB), and both were bound by the atomics: a contended same-address atomicAdd retires every ~4.4 ns on
// this part, so thousands of ticket claims serialise into tens of microseconds.  Hence no tickets.
// Every CTA of the grid is resident (grid <= occupancy x SMs), so writers spinning on a counter cannot keep a
// reader from running.  The scratch restores itself (the last CTA to leave resets counters and maxima): a call
// leaves no state behind -- safe under CUDA-graph replay; one scratch per (device, stream).
// ------------------------------------------------------------------------------------------
// Where the pipeline is used.  Measured on B200 (fp32, us per call, two launches -> one launch): 2048^2 (16 MiB) 16.5 -> 13.4-14.3;
// 4096^2 (64 MiB) 23.5 -> 28.5-31; 8192^2 (256 MiB) 93 -> 102-108.  It does read W from HBM once (ncu: 67 MB at 4096^2 against
// 134 MB), but at 64 MiB and beyond two free-running streaming launches at the HBM roof beat one launch whose second half
// waits on the first; it wins where a launch boundary costs more than the second read (QG_COLS_PIPE=1 forces it on).
// 16-bit inputs lose at every size tried (4096^2 fp16: 19.8 -> 25-31 us), so the window is fp32 only.
constexpr int64_t kPipeMinBytes = 8ll << 20, kPipeMaxBytes = 24ll << 20;
constexpr int kExitFan = 16;     // exit counting is two-level so that no address sees more than grid / 16 atomics
constexpr int kLineInts = 32;    // one counter per 128-byte line
struct ColsPipeGeom {
  int n_groups;          // column groups of 32*EPV columns
  int groups_per_panel;  // G
  int n_panels;
  int rows_per_item;     // R: rows of a phase-1 item
  int rows_per_item2;    // rows of a phase-2 item
  int c1, c2;            // row chunks per column group in phase 1 (rows 1..K-1) / phase 2 (rows 0..K-1)
  int items1, items2;    // phase-1 / phase-2 items over all panels
  int readers_of;        // role pattern: blockIdx % readers_of < readers_in  ->  reader
  int readers_in;
};
// device scratch, all zero / -inf between calls: [exit counters | arrive[n_groups] (one per line) | part[n_cols]]
struct ColsPipeScratch {
  int exit_top, pad[kLineInts - 1];
  int exit_leaf[kExitFan][kLineInts];
};

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
quant_cols_pipe_kernel(const T *__restrict__ W, int K, int N, int64_t ldw, float range, int mode, ColsPipeGeom g,
                       ColsPipeScratch *__restrict__ sc, int *__restrict__ arrive, int *__restrict__ part,
                       int8_t *__restrict__ Wq, int64_t ldq, float *__restrict__ Cw) {
  constexpr int EPV = Unpack<T>::EPV;
  constexpr int U = 8;             // loads in flight per lane
  constexpr int NW = kThreads / 32;
  constexpr uint32_t kNaN = sizeof(T) == 4 ? 0x7fc00000u : (std::is_same<T, __half>::value ? 0x7e007e00u : 0x7fc07fc0u);
  __shared__ float s_m[2][NW][32 * EPV + 1];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role_pos = (int)(blockIdx.x % g.readers_of), role_blk = (int)(blockIdx.x / g.readers_of);
  const bool reader = role_pos < g.readers_in;
  griddep_wait();
  griddep_trigger_early();
  const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  const int chunks = reader ? g.c1 : g.c2;
  const int n_items = reader ? g.items1 : g.items2;
  const int rows_item = reader ? g.rows_per_item : g.rows_per_item2;
  const int per_panel = g.groups_per_panel * chunks;  // items of a full panel (only the last panel may be narrower)
  // this CTA's position among the CTAs of its role, and their number (the grid is a multiple of the role pattern)
  const int role_ctas = (int)(gridDim.x / g.readers_of) * (reader ? g.readers_in : g.readers_of - g.readers_in);
  const int my_cta = reader ? role_blk * g.readers_in + role_pos
                            : role_blk * (g.readers_of - g.readers_in) + (role_pos - g.readers_in);
  auto decode = [&](int item, int &grp, int &chunk) {
    int panel = item / per_panel;
    if (panel >= g.n_panels) panel = g.n_panels - 1;
    const int local = item - panel * per_panel;
    const int gp = min(g.groups_per_panel, g.n_groups - panel * g.groups_per_panel);
    grp = panel * g.groups_per_panel + local % gp;
    chunk = local / gp;
  };

  if (reader) {
    // An item is one column group x R rows; warp w takes rows k0 + w, k0 + w + 8, ...  The loop is software-pipelined:
    // the loads of the NEXT item are issued before this item's reduction, atomics and fence, so the per-item
    // synchronisation chain runs under memory latency instead of adding to it.
    uint4 r[U];
    auto issue = [&](int item) {
      int grp, chunk;
      decode(item, grp, chunk);
      const int col = (grp * 32 + lane) * EPV;
      const int k0 = 1 + chunk * rows_item, k1 = min(K, k0 + rows_item);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int k = k0 + warp + NW * u;
        r[u] = (col < N && k < k1) ? ldg16_hint(W + col + (int64_t)k * ldw, pol_keep) : make_uint4(kNaN, kNaN, kNaN, kNaN);
      }
    };
    int item = my_cta;
    if (item < n_items) issue(item);
    int prev_grp = -1;  // the item whose arrival is still owed: it is published one iteration late, when the fence is free
    for (int it = 0; item < n_items; item += role_ctas, it++) {
      int grp, chunk;
      decode(item, grp, chunk);
      float m[EPV];
#pragma unroll
      for (int e = 0; e < EPV; e++) m[e] = -INFINITY;
#pragma unroll
      for (int u = 0; u < U; u++) {  // rows past the chunk hold NaNs, which fmaxf skips like the reduction skips real ones
        float f[EPV];
        Unpack<T>::run(r[u], f);
#pragma unroll
        for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
      }
      {  // a chunk taller than NW * U rows: the remaining rows, not pipelined
        const int col = (grp * 32 + lane) * EPV;
        const int k0 = 1 + chunk * rows_item, k1 = min(K, k0 + rows_item);
        for (int k = k0 + warp + NW * U; k < k1 && col < N; k += NW) {
          float f[EPV];
          Unpack<T>::run(ldg16_hint(W + col + (int64_t)k * ldw, pol_keep), f);
#pragma unroll
          for (int e = 0; e < EPV; e++) m[e] = fmaxf(m[e], fabsf(f[e]));
        }
      }
      // the previous item's atomics were issued a whole load round trip ago: this fence finds nothing to wait for
      if (prev_grp >= 0 && threadIdx.x < 32 * EPV) __threadfence();
      float(*sm)[32 * EPV + 1] = s_m[it & 1];
#pragma unroll
      for (int e = 0; e < EPV; e++) sm[warp][lane * EPV + e] = m[e];
      __syncthreads();  // the only barrier of an iteration (s_m alternates, so the next write needs none)
      if (threadIdx.x == 0 && prev_grp >= 0) atomicAdd(arrive + prev_grp * kLineInts, 1);
      for (int c = threadIdx.x; c < 32 * EPV; c += kThreads) {
        float v = sm[0][c];
#pragma unroll
        for (int y = 1; y < NW; y++) v = fmaxf(v, sm[y][c]);
        const int gc = grp * 32 * EPV + c;
        if (gc < N && v >= 0.0f) atomicMax(part + gc, __float_as_int(v));  // candidates >= +0, initial value -inf
      }
      if (item + role_ctas < n_items) issue(item + role_ctas);
      prev_grp = grp;
    }
    if (prev_grp >= 0) {
      if (threadIdx.x < 32 * EPV) __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(arrive + prev_grp * kLineInts, 1);
    }
  } else {
    for (int item = my_cta; item < n_items; item += role_ctas) {
      int grp, chunk;
      decode(item, grp, chunk);
      const int col = (grp * 32 + lane) * EPV;
      const T *base = W + col;
      if (g.c1 > 0) {  // all phase-1 items of this column group have arrived?
        if (lane == 0)
          while (ld_acquire_gpu(arrive + grp * kLineInts) < g.c1) __nanosleep(64);
        __syncwarp();
      }
      if (col < N) {
        float s[EPV];
        float x0[EPV];
        Unpack<T>::run(ldg16(base), x0);
#pragma unroll
        for (int e = 0; e < EPV; e++) {
          float c;
          if (fold_first(x0[e], __int_as_float(__ldcg(part + col + e)), mode, c)) {
            for (int k = 1; k < K; k++) {
              const float x = to_f32(base[(int64_t)k * ldw + e]);
              if (x == x) { c = -x; break; }
            }
          }
          if (chunk == 0 && warp == 0 && Cw != nullptr) Cw[col + e] = c;
          s[e] = __fdiv_rn(range, c);
        }
        const int k0 = chunk * rows_item, k1 = min(K, k0 + rows_item);
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
        // rows k0 + warp + 8 * (4 s + u): two register sets of four loads, one in flight while the other is converted
        constexpr int UH = U / 2;
        auto ld = [&](uint4 (&r)[UH], int st) {
#pragma unroll
          for (int u = 0; u < UH; u++) {
            const int k = k0 + warp + NW * (UH * st + u);
            if (k < k1) r[u] = ldg16_hint(base + (int64_t)k * ldw, pol_drop);
          }
        };
        auto put = [&](const uint4 (&r)[UH], int st) {
#pragma unroll
          for (int u = 0; u < UH; u++) {
            const int k = k0 + warp + NW * (UH * st + u);
            if (k < k1) emit(r[u], k);
          }
        };
        const int n_steps = (k1 - k0 + NW * UH - 1) / (NW * UH);
        uint4 ra[UH], rb[UH];
        ld(ra, 0);
        for (int st = 0; st < n_steps; st += 2) {
          ld(rb, st + 1);
          put(ra, st);
          ld(ra, st + 2);
          put(rb, st + 1);
        }
      }
    }
  }
  // the last CTA out puts the scratch back (two-level count: leaf = blockIdx % 16)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int leaf = (int)(blockIdx.x % kExitFan);
    const int leaf_size = (int)((gridDim.x - leaf + kExitFan - 1) / kExitFan);
    int last = 0;
    if (atomicAdd(&sc->exit_leaf[leaf][0], 1) == leaf_size - 1) {
      const int leaves = (int)min((unsigned)kExitFan, gridDim.x);
      last = atomicAdd(&sc->exit_top, 1) == leaves - 1;
    }
    s_last = last;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    const int ng = g.n_groups * 32 * EPV;
    const int neg_inf = __float_as_int(-INFINITY);
    for (int j = threadIdx.x; j < ng; j += kThreads) part[j] = neg_inf;
    for (int j = threadIdx.x; j < g.n_groups; j += kThreads) arrive[j * kLineInts] = 0;
    if (threadIdx.x < kExitFan) sc->exit_leaf[threadIdx.x][0] = 0;
    if (threadIdx.x == 0) sc->exit_top = 0;
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
  griddep_trigger_early();
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
  griddep_trigger_early();
  if (i < n) out[i] = __fdiv_rn(b, a[i]);
}

// AbsCompareLTEConstFunc, src/ops/op_elemwise.cuh:292-306
__global__ void outlier_mask_kernel(const float *__restrict__ A, int M, int K, int64_t lda, float thr,
                                    float *__restrict__ mask, int64_t ldm) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  griddep_trigger_early();
  if (k >= K) return;
  for (int i = blockIdx.y; i < M; i += gridDim.y) {
    const float a = A[(int64_t)i * lda + k];
    const bool inl = ((a >= 0) & (a <= thr)) | ((a <= 0) & (-a <= thr));
    mask[(int64_t)i * ldm + k] = inl ? 0.0f : 1.0f;
  }
}

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Epoch-tagged column-max scratch: zero-initialised once, never reset (each call uses a fresh epoch).
// One buffer per (device, stream): calls that quantize weights on different streams do not share it.
// Epoch 0 is never used (zero-initialised entries never match); when the 32-bit epoch wraps the buffer
// is cleared on the stream and the count restarts at 1, so stale tags can never win an atomicMax.
struct ColScratch {
  unsigned long long *part = nullptr;
  int n = 0;
  uint32_t epoch = 0;
};
int twopass_scratch(int N, cudaStream_t st, unsigned long long **part, uint32_t *epoch) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, ColScratch> states;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  if (dev < 0 || dev >= kMaxDevices) return (int)cudaErrorInvalidDevice;
  std::lock_guard<std::mutex> lk(mu);
  ColScratch &s = states[std::make_pair(dev, st)];
  if (s.n < N) {
    if (s.part) { cudaStreamSynchronize(st); cudaFree(s.part); s.part = nullptr; }
    s.n = (int)round_up(N < 16384 ? 16384 : N, 4096);
    if ((e = cudaMalloc(&s.part, sizeof(unsigned long long) * (size_t)s.n)) != cudaSuccess) { s = ColScratch(); return (int)e; }
    if ((e = cudaMemsetAsync(s.part, 0, sizeof(unsigned long long) * (size_t)s.n, st)) != cudaSuccess) return (int)e;
    s.epoch = 0;
  }
  if (++s.epoch == 0) {  // wrapped after 2^32 calls: forget every old tag
    if ((e = cudaMemsetAsync(s.part, 0, sizeof(unsigned long long) * (size_t)s.n, st)) != cudaSuccess) return (int)e;
    s.epoch = 1;
  }
  *part = s.part;
  *epoch = s.epoch;
  return 0;
}

// Scratch of the panel pipeline: [ColsPipeScratch | arrive[groups] one per line | part[cols]], zero / -inf at rest,
// one per (device, stream).
struct PipeState {
  ColsPipeScratch *sc = nullptr;
  int cols = 0;
};
__global__ void pipe_scratch_init_kernel(int *words, int n_zero, int n_total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_total) words[i] = i < n_zero ? 0 : __float_as_int(-INFINITY);
}
int pipe_scratch(int n_cols, cudaStream_t st, ColsPipeScratch **sc, int **arrive, int **part) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, PipeState> states;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return (int)e;
  std::lock_guard<std::mutex> lk(mu);
  PipeState &s = states[std::make_pair(dev, st)];
  if (s.cols < n_cols) {
    if (s.sc) { cudaStreamSynchronize(st); cudaFree(s.sc); s.sc = nullptr; }
    s.cols = (int)round_up(n_cols < 16384 ? 16384 : n_cols, 4096);
    const size_t head = sizeof(ColsPipeScratch) / sizeof(int), arr = (size_t)(s.cols / 128) * kLineInts;  // >= one group per 128 columns
    if ((e = cudaMalloc(&s.sc, sizeof(int) * (head + arr + (size_t)s.cols))) != cudaSuccess) { s = PipeState(); return (int)e; }
    const int n_total = (int)(head + arr + (size_t)s.cols);
    pipe_scratch_init_kernel<<<(unsigned)ceil_div(n_total, 256), 256, 0, st>>>(reinterpret_cast<int *>(s.sc), (int)(head + arr), n_total);
    if ((e = cudaGetLastError()) != cudaSuccess) return (int)e;
  }
  *sc = s.sc;
  *arrive = reinterpret_cast<int *>(s.sc) + sizeof(ColsPipeScratch) / sizeof(int);
  *part = *arrive + (size_t)(s.cols / 128) * kLineInts;
  return 0;
}

// Geometry of the panel pipeline.  Defaults: panels of <= 16 MiB, warp items of 32 rows, 4 CTAs per SM, 5 of 8 CTAs read.
// QG_COLS_PIPE = "0" switches the pipeline off (two launches); "G,R,readers_in,readers_of[,ctas_per_sm[,R2]]" overrides.
struct PipeTune { int on = 1, force = 0, G = 0, R = 0, rin = 0, rof = 0, per_sm = 0, R2 = 0; };
const PipeTune &pipe_tune() {
  static const PipeTune t = [] {
    PipeTune v;
    if (const char *e = getenv("QG_COLS_PIPE")) {
      int a = 0, b = 0, c = 0, d = 0, f = 0, h = 0;
      const int n = sscanf(e, "%d,%d,%d,%d,%d,%d", &a, &b, &c, &d, &f, &h);
      if (n == 1 && a == 0) v.on = 0;
      if (n == 1 && a == 1) v.force = 1;
      if (n >= 4) { v.G = a; v.R = b; v.rin = c; v.rof = d; v.force = 1; }
      if (n >= 5) v.per_sm = f;
      if (n >= 6) v.R2 = h;
    }
    return v;
  }();
  return t;
}

template <typename T>
int cols_pipe_launch(const T *W, int K, int N, int64_t ldw, float range, int mode, int8_t *Wq, int64_t ldq, float *Cw,
                     cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  const PipeTune &tune = pipe_tune();
  static int per_sm_occ = 0, sms = 0;
  if (per_sm_occ == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_occ, quant_cols_pipe_kernel<T>, kThreads, 0);
    if (per_sm_occ <= 0) per_sm_occ = 1;
    if (sms <= 0) sms = 148;
  }
  int per_sm = tune.per_sm > 0 ? tune.per_sm : 4;
  if (per_sm > per_sm_occ) per_sm = per_sm_occ;
  ColsPipeGeom g = {};
  g.readers_of = tune.rof > 0 ? tune.rof : 8;
  g.readers_in = tune.rin > 0 ? tune.rin : 5;
  if (g.readers_in >= g.readers_of) g.readers_in = g.readers_of - 1;
  if (g.readers_in < 1 || per_sm < 2) return -1;
  g.n_groups = (int)ceil_div(N, 32 * EPV);
  const int64_t group_row_bytes = 32 * 16;  // one column group of one row
  int G = tune.G > 0 ? tune.G : (int)((8ll << 20) / ((int64_t)K * group_row_bytes));
  if (G < 1) G = 1;
  if (G > g.n_groups) G = g.n_groups;
  g.groups_per_panel = G;
  g.n_panels = (int)ceil_div(g.n_groups, G);
  const int R = tune.R > 0 ? (int)round_up(tune.R, 8) : 64, R2 = tune.R2 > 0 ? (int)round_up(tune.R2, 8) : 256;
  g.rows_per_item = R;
  g.rows_per_item2 = R2;
  g.c1 = K > 1 ? (int)ceil_div(K - 1, R) : 0;
  g.c2 = (int)ceil_div(K, R2);
  const int64_t i1 = (int64_t)g.n_groups * g.c1, i2 = (int64_t)g.n_groups * g.c2;
  if (i1 + i2 > 0x3fffffff) return -1;  // caller falls back to the two-launch form
  g.items1 = (int)i1;
  g.items2 = (int)i2;
  static const int dbg = [] { const char *e = getenv("QG_COLS_PIPE_DBG"); return e ? atoi(e) : 0; }();
  if (dbg == 1) g.items2 = 0;               // timing experiment: readers alone (no codes are written)
  if (dbg == 2) { g.items1 = 0; g.c1 = 0; }  // timing experiment: writers alone, nothing to wait for (wrong scales)
  ColsPipeScratch *sc = nullptr;
  int *part = nullptr, *arrive = nullptr;
  int rc = pipe_scratch(g.n_groups * 32 * EPV, st, &sc, &arrive, &part);
  if (rc) return rc;
  // every CTA must be resident: a multiple of the role pattern, at most occupancy x SMs
  unsigned grid = (unsigned)(per_sm * sms);
  grid -= grid % (unsigned)g.readers_of;
  if (grid < (unsigned)g.readers_of) return -1;
  return (int)launch_kernel(quant_cols_pipe_kernel<T>, dim3(grid), dim3(kThreads), st, W, K, N, ldw, range, mode, g, sc, arrive,
                            part, Wq, ldq, Cw);
}

// ---- row launcher ----
template <typename T, int G, int NV>
void launch_rows(const T *X, int M, int K, int64_t ldx, float range, int mode, const float *sx, int8_t *Xq,
                 int64_t ldq, float *Cx, RowMaxIo io, cudaStream_t st) {
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
  launch_kernel(quant_rows_kernel<T, G, NV>, dim3(grid), dim3(kThreads), st, X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io);
}

template <typename T>
int rows_dispatch(const T *X, int M, int K, int64_t ldx, float range, int mode, const float *sx, int8_t *Xq,
                  int64_t ldq, float *Cx, RowMaxIo io, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  const bool vec_ok = (K % EPV == 0) && aligned(X, 16) && ((ldx * sizeof(T)) % 16 == 0) &&
                      (Xq == nullptr || (aligned(Xq, EPV) && ldq % EPV == 0));
  if (!vec_ok) {
    return (int)launch_kernel(quant_rows_generic_kernel<T>, dim3((unsigned)ceil_div(M, kThreads / 32)), dim3(kThreads), st,
                              X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io);
  }
  const int nvec = K / EPV;
#define QG_ROWS(G, NV) launch_rows<T, G, NV>(X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io, st)
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
inline int cols_rows_per_cta(int K, int col_tiles, int round) {
  // One wave: every CTA of the grid is resident at once (6 blocks of 256 threads per SM), so all of
  // them start together, stream the same number of rows and finish together.  Measured at 4096^2
  // (both passes, us; fp32 / fp16 input): 4 per SM 23.5 / 17.5, 6: 24.1 / 17.5, 8: 25.7 / 20.5,
  // ~14 (the earlier 2048 short blocks, 1.73 waves): 25.3 / 24.4; at 8192^2 fp32 4: 102.7, 6: 97.3,
  // 8: 95.8, 16: 93.9.  QG_COLS_CTAS_PER_SM overrides the 6 for experiments.
  static const int target = [] {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char *e = getenv("QG_COLS_CTAS_PER_SM");
    const int per_sm = e != nullptr && atoi(e) > 0 ? atoi(e) : 6;
    return (sms > 0 ? sms : 148) * per_sm;
  }();
  const int64_t chunks = target / col_tiles > 0 ? target / col_tiles : 1;
  int64_t rows = round_up(ceil_div(K, chunks), round);
  return (int)(rows < 32 ? 32 : rows);
}

template <typename T>
int cols_dispatch(const T *W, int K, int N, int64_t ldw, float range, int mode, const float *sw, int8_t *Wq,
                  int64_t ldq, float *Cw, bool transpose, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  if (transpose && Wq != nullptr && !(aligned(Wq, 16) && ldq % 16 == 0))
    return (int)launch_kernel(quant_cols_generic_kernel<T>, dim3((unsigned)ceil_div(N, kThreads)), dim3(kThreads), st, W, K,
                              N, ldw, range, mode, sw, Wq, ldq, Cw, transpose);
  const bool vec_ok = (N % EPV == 0) && aligned(W, 16) && ((ldw * sizeof(T)) % 16 == 0) &&
                      (Wq == nullptr || (aligned(Wq, EPV) && ldq % EPV == 0)) &&
                      !(transpose && sw != nullptr);
  if (!vec_ok) {
    return (int)launch_kernel(quant_cols_generic_kernel<T>, dim3((unsigned)ceil_div(N, kThreads)), dim3(kThreads), st, W, K,
                              N, ldw, range, mode, sw, Wq, ldq, Cw, transpose);
  }
  // one launch, one HBM read of W (panel pipeline) once the matrix is large enough for the second read to matter
  if (!transpose && sw == nullptr && Wq != nullptr && pipe_tune().on && (sizeof(T) == 4 || pipe_tune().force) &&
      (int64_t)K * N * (int64_t)sizeof(T) >= kPipeMinBytes &&
      ((int64_t)K * N * (int64_t)sizeof(T) <= kPipeMaxBytes || pipe_tune().force)) {
    const int rc = cols_pipe_launch(W, K, N, ldw, range, mode, Wq, ldq, Cw, st);
    if (rc >= 0) return rc;
  }
  const int col_tiles = (int)ceil_div(N, 32 * EPV);
  static const int wave_slabs = [] { const char *e = getenv("QG_COLS_WAVE"); return e ? atoi(e) : 0; }();
  if (wave_slabs >= 2 && !transpose && sw == nullptr && Wq != nullptr && K > 1 && col_tiles >= wave_slabs) {
    unsigned long long *part = nullptr;
    uint32_t epoch = 0;
    int rc = twopass_scratch(N, st, &part, &epoch);
    if (rc) return rc;
    const int per = (int)ceil_div(col_tiles, wave_slabs);
    // each role gets half of the resident CTAs of a launch
    const int rpc1 = cols_rows_per_cta(K, 2 * per, 8), rpc2 = rpc1;
    const int chunks1 = (int)ceil_div(K - 1, rpc1), chunks2 = (int)ceil_div(K, rpc2);
    for (int i = 0; i <= wave_slabs; i++) {
      const int t1_first = i * per, t1_count = i < wave_slabs ? std::min(per, col_tiles - t1_first) : 0;
      const int t2_first = (i - 1) * per, t2_count = i >= 1 ? std::min(per, col_tiles - t2_first) : 0;
      dim3 grid((unsigned)std::max(std::max(t1_count, t2_count), 1), (unsigned)std::max(chunks1, chunks2), 2);
      cudaError_t e = launch_kernel(cols_wave_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, range, mode, rpc1, rpc2, part,
                                    epoch, Wq, ldq, Cw, t1_first, t1_count > 0 ? t1_count : 0, t2_first, t2_count > 0 ? t2_count : 0,
                                    chunks1, chunks2, i == wave_slabs ? 1 : 0);
      if (e != cudaSuccess) return (int)e;
    }
    return 0;
  }
  const int rpc = cols_rows_per_cta(K, col_tiles, transpose ? 32 : 8);
  unsigned long long *part = nullptr;
  uint32_t epoch = 0;
  if (sw == nullptr) {
    int rc = twopass_scratch(N, st, &part, &epoch);
    if (rc) return rc;
    if (K > 1) {
      dim3 grid(col_tiles, (unsigned)ceil_div(K - 1, rpc));
      launch_kernel(absmax_cols_partial_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, rpc, part, epoch);
    }
    if (Wq == nullptr)
      return (int)launch_kernel(absmax_cols_finalize_kernel<T>, dim3((unsigned)ceil_div(N, 256)), dim3(256), st, W, K, N,
                                ldw, mode, part, epoch, Cw);
  }
  dim3 grid(col_tiles, (unsigned)ceil_div(K, rpc));
  if (transpose)
    return (int)launch_kernel(quant_cols_t_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, range, mode, rpc, part, epoch,
                              Wq, ldq, Cw);
  return (int)launch_kernel(quant_cols_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, range, mode, rpc, part, epoch, sw,
                            Wq, ldq, Cw);
}

// pass 1 over W, then [pass 2 over W || row quantizer over X] in one launch.  Returns -1 when the shapes are outside what
// the fused launch covers (the caller then runs the two quantizers one after the other).
template <typename T, int G, int NV>
int fused_launch(const T *W, int K, int N, int64_t ldw, float range, int mode, int8_t *Wq, int64_t ldq, float *Cw, const T *X,
                 int M, int64_t ldx, int8_t *Xq, int64_t ldxq, float *Cx, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  static int per_sm = 0, sms = 0;
  if (per_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, quant_cols2_rows_kernel<T, G, NV>, kThreads, 0);
    if (per_sm <= 0) per_sm = 1;
    if (sms <= 0) sms = 148;
  }
  static const int cols_share_pct = [] { const char *e = getenv("QG_FUSED_COLS_PCT"); return e && atoi(e) > 0 && atoi(e) < 100 ? atoi(e) : 45; }();
  const int resident = per_sm * sms;
  const int col_tiles = (int)ceil_div(N, 32 * EPV);
  // pass 2 on its share of the resident CTAs, one wave; the rest walk the rows of X
  int cols_target = resident * cols_share_pct / 100;
  if (cols_target < col_tiles) cols_target = col_tiles;
  const int chunks = cols_target / col_tiles > 0 ? cols_target / col_tiles : 1;
  int rpc2 = (int)round_up(ceil_div(K, chunks), 8);
  if (rpc2 < 32) rpc2 = 32;
  const int cols_ctas = col_tiles * (int)ceil_div(K, rpc2);
  constexpr int RPB = kThreads / G;
  int rows_ctas = resident - cols_ctas;
  const int nrb = (int)ceil_div(M, RPB);
  if (rows_ctas > nrb) rows_ctas = nrb;
  if (rows_ctas < sms) return -1;
  unsigned long long *part = nullptr;
  uint32_t epoch = 0;
  int rc = twopass_scratch(N, st, &part, &epoch);
  if (rc) return rc;
  {  // pass 1: all resident CTAs, as when it runs alone
    const int rpc1 = cols_rows_per_cta(K, col_tiles, 8);
    dim3 grid(col_tiles, (unsigned)ceil_div(K - 1, rpc1));
    cudaError_t e = launch_kernel(absmax_cols_partial_kernel<T>, grid, dim3(kThreads), st, W, K, N, ldw, rpc1, part, epoch);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)launch_kernel(quant_cols2_rows_kernel<T, G, NV>, dim3((unsigned)(cols_ctas + rows_ctas)), dim3(kThreads), st, W, K, N,
                            ldw, range, mode, rpc2, part, epoch, Wq, ldq, Cw, col_tiles, cols_ctas, X, M, ldx, Xq, ldxq, Cx);
}

template <typename T>
int fused_dispatch(const T *W, int K, int N, int64_t ldw, float range, int mode, int8_t *Wq, int64_t ldq, float *Cw, const T *X,
                   int M, int64_t ldx, int8_t *Xq, int64_t ldxq, float *Cx, cudaStream_t st) {
  constexpr int EPV = Unpack<T>::EPV;
  static const bool off = getenv("QG_NO_FUSED_QUANT") != nullptr;
  const bool w_ok = (N % EPV == 0) && aligned(W, 16) && ((ldw * sizeof(T)) % 16 == 0) && aligned(Wq, EPV) && ldq % EPV == 0 && K > 1;
  const bool x_ok = (K % EPV == 0) && aligned(X, 16) && ((ldx * sizeof(T)) % 16 == 0) && aligned(Xq, EPV) && ldxq % EPV == 0;
  // large enough for both halves to fill their share of the machine (small problems are launch-bound either way), and W
  // small enough to still be in L2 when pass 2 comes for it: beyond that pass 2 is HBM-bound like the row quantizer and
  // sharing a launch only halves what each gets (8192^2 fp32, 256 MiB: 541.6 us per op fused against 526.8 us apart)
  const int64_t w_bytes = (int64_t)K * N * (int64_t)sizeof(T);
  if (off || !w_ok || !x_ok || (int64_t)K * N < (4ll << 20) || (int64_t)M * K < (4ll << 20) || w_bytes > (80ll << 20)) return -1;
  const int nvec = K / EPV;
#define QG_FUSED(G, NV) return fused_launch<T, G, NV>(W, K, N, ldw, range, mode, Wq, ldq, Cw, X, M, ldx, Xq, ldxq, Cx, st)
  if (nvec > 256 && nvec <= 512) QG_FUSED(128, 4);
  if (nvec > 512 && nvec <= 1024) QG_FUSED(256, 4);
  if (nvec > 1024 && nvec <= 2048) QG_FUSED(256, 8);
#undef QG_FUSED
  return -1;
}

}  // namespace

// The op's two quantizers together (qg_quantized_mm, weights in the reference's [K,N] code layout): -1 = not covered,
// run quant_rows and quant_cols separately.
int quant_rows_cols_fused(const void *X, const void *W, int dtype, int M, int N, int K, int64_t ldx, int64_t ldw, float range, int mode,
                          int8_t *Xq, int64_t ldxq, float *Cx, int8_t *Wq, int64_t ldwq, float *Cw, cudaStream_t st) {
  switch (dtype) {
    case QG_F32:
      return fused_dispatch((const float *)W, K, N, ldw, range, mode, Wq, ldwq, Cw, (const float *)X, M, ldx, Xq, ldxq, Cx, st);
    case QG_F16:
      return fused_dispatch((const __half *)W, K, N, ldw, range, mode, Wq, ldwq, Cw, (const __half *)X, M, ldx, Xq, ldxq, Cx, st);
    case QG_BF16:
      return fused_dispatch((const __nv_bfloat16 *)W, K, N, ldw, range, mode, Wq, ldwq, Cw, (const __nv_bfloat16 *)X, M, ldx, Xq,
                            ldxq, Cx, st);
  }
  return -1;
}

// ---- entry points used by capi.cu ----
int quant_rows(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, const float *sx,
               int8_t *Xq, int64_t ldq, float *Cx, cudaStream_t st, RowMaxIo io) {
  switch (dtype) {
    case QG_F32: return rows_dispatch((const float *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io, st);
    case QG_F16: return rows_dispatch((const __half *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io, st);
    case QG_BF16: return rows_dispatch((const __nv_bfloat16 *)X, M, K, ldx, range, mode, sx, Xq, ldq, Cx, io, st);
  }
  return QG_EINVAL;
}

// transpose: codes are written as Wt[n][k] with leading dimension ldq (K-major operand layout)
int quant_cols(const void *W, int dtype, int K, int N, int64_t ldw, float range, int mode, const float *sw,
               int8_t *Wq, int64_t ldq, float *Cw, bool transpose, cudaStream_t st) {
  switch (dtype) {
    case QG_F32: return cols_dispatch((const float *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, transpose, st);
    case QG_F16: return cols_dispatch((const __half *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, transpose, st);
    case QG_BF16:
      return cols_dispatch((const __nv_bfloat16 *)W, K, N, ldw, range, mode, sw, Wq, ldq, Cw, transpose, st);
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
