// attention.cu -- the fp32 core of AttentionLayer::forward in one kernel.
//
// The reference's layer (src/modules/attention.cuh:58-69) runs, after the projections,
//   op_mm(Q, K^T, S); op_multiply(S, 1/sqrt(d_k)); op_softmax(S); op_mm(S, V, out)
// as separate launches over a [tokens, tokens] score matrix in global memory.  Re-pointed naively that is three
// kernels (scores, softmax, P.V) and four passes over the scores per (sequence, head): at BASELINE config 3
// (32 sequences x 128 tokens, 8 heads) they were 96 us of a 267 us encoder block.  Here one CTA owns one
// (sequence, head, 64-row block of queries): Q, K, V tiles and the score tile live in shared memory, nothing but
// the head's output goes back to HBM.
//
// The arithmetic is the reference's, operation by operation, so the result is bit-identical to the three-kernel path
// (and to the reference-composed fixtures):
//   S[i,j]  = k-ascending chain acc = fma(Q[i,k], K[j,k], acc) from +0      (op_mm.cuh:37-39; "+ 0.0f" when d_k % 32 != 0:
//             the reference's last 32-wide tile pads with fma(0,0,acc), which turns -0 into +0)
//   T[i,j]  = fl(S[i,j] * scale)                                             (op_multiply)
//   max_i   = T[i,0], then strict '>' updates over j (NaN at column 0 stays) (op_softmax.cuh)
//   E[i,j]  = expf(T[i,j] - max_i);  sum_i = E[i,0] + E[i,1] + ... ascending;  P = E / sum (IEEE division)
//   out[i,d]= j-ascending chain acc = fma(P[i,j], V[j,d], acc) from +0      ("+ 0.0f" when skv % 32 != 0)
#include <cstdlib>

#include "common.cuh"

namespace qg {

namespace {

constexpr int kAttRows = 64;   // query rows per CTA
constexpr int kAttKeys = 128;  // keys / values per (sequence, head) the fused kernel takes
constexpr int kAttDim = 64;    // d_k, d_v
constexpr int kAttThreads = 256;
constexpr int kLdQ = kAttRows + 4;   // Q^T tile [k][row]: rows padded so that 16-byte reads stay aligned
constexpr int kLdK = kAttKeys + 4;   // K^T tile [k][key]; the same region holds V [key][kAttDim] after the scores
constexpr int kLdSt = kAttRows + 4;  // score tile, TRANSPOSED [key][row]: a thread's four rows of one key are one 16-byte access
constexpr int kKVFloats = kAttDim * kLdK > kAttKeys * kAttDim ? kAttDim * kLdK : kAttKeys * kAttDim;
constexpr size_t kAttSmem =
    sizeof(float) * ((size_t)kAttDim * kLdQ + (size_t)kKVFloats + (size_t)kAttKeys * kLdSt + 5 * kAttRows);  // 85 KB: two CTAs per SM

// Where key `key` sits in a row of the staged K^T.  Thread tx of the score phase owns keys tx, tx + 16, ..., tx + 112 (so that the
// four rows x eight keys it produces go into the transposed score tile as eight conflict-free 16-byte stores); this placement
// puts them at positions tx*4 .. tx*4+3 and 64 + tx*4 .. : two 16-byte loads, sixteen lanes reading 256 contiguous bytes each.
__device__ __forceinline__ int key_pos(int key) {
  const int c = key >> 4, tx = key & 15;
  return ((c >> 2) << 6) + (tx << 2) + (c & 3);
}

__global__ void __launch_bounds__(kAttThreads, 2)
attention_core_kernel(const float *__restrict__ Q, int64_t ldq, const float *__restrict__ K, int64_t ldk,
                      const float *__restrict__ V, int64_t ldv, float *__restrict__ O, int64_t ldo, int sq, int skv, int d_k,
                      int d_v, float scale, MmBatch bt) {
  extern __shared__ float att_smem[];
  float *sQt = att_smem;                      // [d_k][kLdQ]   Q^T of this row block
  float *sKV = sQt + kAttDim * kLdQ;          // [d_k][kLdK]   K^T, then [skv][kAttDim] V
  float *sSt = sKV + kKVFloats;               // [keys][kLdSt] scores -> probabilities, transposed
  float *sStat = sSt + kAttKeys * kLdSt;      // [rows] row maximum, then row sum; [4][rows] partial maxima behind it
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int m0 = blockIdx.y * kAttRows;
  const int rows = min(kAttRows, sq - m0);
  {  // blockIdx.z = sequence * heads + head (offsets as in mm_f32's batch geometry: a = Q, b = K / V, c = out)
    const int zo = blockIdx.z / bt.n_inner, zi = blockIdx.z % bt.n_inner;
    Q += zo * bt.a_outer + zi * bt.a_inner + (int64_t)m0 * ldq;
    K += zo * bt.b_outer + zi * bt.b_inner;
    V += zo * bt.b_outer + zi * bt.c_inner;  // V's own base pointer: per head it advances by d_v columns
    O += zo * bt.c_outer + (int64_t)zi * d_v + (int64_t)m0 * ldo;
  }
  griddep_wait();
  griddep_trigger_early();

  // ---- V into registers now (it takes K^T's place in shared memory once the scores are done): 8 x 16 bytes per thread ----
  const int dv4 = d_v >> 2;
  float4 vreg[8];
#pragma unroll
  for (int u = 0; u < 8; u++) {
    const int e = t + u * kAttThreads;          // e < 128 * 16
    const int r = e >> 4, c4 = e & 15;
    vreg[u] = (r < skv && c4 < dv4) ? *reinterpret_cast<const float4 *>(V + (int64_t)r * ldv + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // ---- stage Q^T, K^T.  A warp fetches 8 rows x 64 bytes per request (whole 32-byte sectors; with one row per lane half of
  //      every sector was wasted and each request touched 32 lines).  All of a thread's loads (4 of Q, 8 of K, besides the 8 of
  //      V above) are issued before the first transposed store: a load-store-load-store loop was a chain of eight global-memory
  //      round trips per warp (ncu: 14 % of the kernel's samples sat on the two store instructions, waiting for their load) ----
  const int kq4 = d_k >> 2;
  {
    const int lane = t & 31, rl = lane & 7, kl = lane >> 3, wid = t >> 5;  // 8 rows x 4 vectors per warp request
    constexpr int kWarps = kAttThreads / 32;
    constexpr int kQReq = (kAttRows / 8) * (kAttDim / 16) / kWarps;   // 4 requests per warp
    constexpr int kKReq = (kAttKeys / 8) * (kAttDim / 16) / kWarps;   // 8
    float4 qreg[kQReq], kreg[kKReq];
#pragma unroll
    for (int u = 0; u < kQReq; u++) {
      const int w = wid + u * kWarps;
      const int r = (w % (kAttRows / 8)) * 8 + rl, k4 = (w / (kAttRows / 8)) * 4 + kl;
      qreg[u] = (r < rows && k4 < kq4) ? *reinterpret_cast<const float4 *>(Q + (int64_t)r * ldq + 4 * k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kKReq; u++) {
      const int w = wid + u * kWarps;
      const int r = (w % (kAttKeys / 8)) * 8 + rl, k4 = (w / (kAttKeys / 8)) * 4 + kl;
      kreg[u] = (r < skv && k4 < kq4) ? *reinterpret_cast<const float4 *>(K + (int64_t)r * ldk + 4 * k4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kQReq; u++) {
      const int w = wid + u * kWarps;
      const int r = (w % (kAttRows / 8)) * 8 + rl, k4 = (w / (kAttRows / 8)) * 4 + kl;
      if (k4 < kq4) {
        sQt[(4 * k4 + 0) * kLdQ + r] = qreg[u].x; sQt[(4 * k4 + 1) * kLdQ + r] = qreg[u].y;
        sQt[(4 * k4 + 2) * kLdQ + r] = qreg[u].z; sQt[(4 * k4 + 3) * kLdQ + r] = qreg[u].w;
      }
    }
#pragma unroll
    for (int u = 0; u < kKReq; u++) {
      const int w = wid + u * kWarps;
      const int r = (w % (kAttKeys / 8)) * 8 + rl, k4 = (w / (kAttKeys / 8)) * 4 + kl;
      if (k4 < kq4) {
        const int c = key_pos(r);
        sKV[(4 * k4 + 0) * kLdK + c] = kreg[u].x; sKV[(4 * k4 + 1) * kLdK + c] = kreg[u].y;
        sKV[(4 * k4 + 2) * kLdK + c] = kreg[u].z; sKV[(4 * k4 + 3) * kLdK + c] = kreg[u].w;
      }
    }
  }
  __syncthreads();

  // ---- scores: 4 rows x 8 keys per thread (keys tx + 16 j), k ascending ----
  {
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = 0.0f;
#pragma unroll 4
    for (int k = 0; k < d_k; k++) {
      const float4 a0 = *reinterpret_cast<const float4 *>(sQt + k * kLdQ + ty * 4);
      // positions tx*4 .. and 64 + tx*4 .. of the staged row hold this thread's keys (key_pos): sixteen lanes read 256 contiguous bytes
      const float4 b0 = *reinterpret_cast<const float4 *>(sKV + k * kLdK + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4 *>(sKV + k * kLdK + 64 + tx * 4);
      const float a[4] = {a0.x, a0.y, a0.z, a0.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    const bool pad = (d_k % 32) != 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      float sc[4];
#pragma unroll
      for (int i = 0; i < 4; i++) sc[i] = __fmul_rn(pad ? __fadd_rn(acc[i][j], 0.0f) : acc[i][j], scale);
      // key tx + 16 j, rows ty*4 .. ty*4+3: one 16-byte store; a warp's 32 stores cover all banks four times (no conflict)
      *reinterpret_cast<float4 *>(sSt + (tx + 16 * j) * kLdSt + ty * 4) = make_float4(sc[0], sc[1], sc[2], sc[3]);
    }
  }
  __syncthreads();  // every thread is done with K^T
#pragma unroll
  for (int u = 0; u < 8; u++) {  // V takes its place
    const int e = t + u * kAttThreads;
    *reinterpret_cast<float4 *>(sKV + (e >> 4) * kAttDim + 4 * (e & 15)) = vreg[u];
  }

  // ---- softmax.  Only the row maximum and the running sum walk a row in order (one thread per row; consecutive threads read
  //      consecutive words of a key's line); the exponentials and the divisions are independent per element, all threads ----
  // row maximum: mx = S[i,0], then strict '>' updates over j (a NaN never replaces, a NaN at column 0 stays, of equal values
  // the first stays).  Four threads per row take 32 keys each -- the three that do not own column 0 start from -inf, which no
  // value loses to and which never wins the strict comparison below -- and the partial results are merged in key order: the
  // same value as the one-thread walk, a quarter of its dependent chain.
  const int er = t & (kAttRows - 1), ej0 = t >> 6;
  {
    const int j0 = ej0 * (kAttKeys / 4), j1 = min(skv, j0 + kAttKeys / 4);
    float mx = ej0 == 0 ? sSt[er] : -INFINITY;
#pragma unroll 8
    for (int j = j0 + (ej0 == 0 ? 1 : 0); j < j1; j++) {
      const float v = sSt[j * kLdSt + er];
      if (v > mx) mx = v;
    }
    sStat[kAttRows + ej0 * kAttRows + er] = mx;
  }
  __syncthreads();
  if (t < rows) {
    float mx = sStat[kAttRows + t];
#pragma unroll
    for (int q = 1; q < 4; q++) {
      const float v = sStat[kAttRows + q * kAttRows + t];
      if (v > mx) mx = v;
    }
    sStat[t] = mx;
  }
  __syncthreads();
  // (thread = one query row, every fourth key: no index division, conflict-free)
  if (er < rows) {
    const float mx = sStat[er];
    for (int j = ej0; j < skv; j += kAttThreads / kAttRows) sSt[j * kLdSt + er] = expf(__fsub_rn(sSt[j * kLdSt + er], mx));
  }
  __syncthreads();
  if (t < rows) {
    float sum = 0.0f;
#pragma unroll 8
    for (int j = 0; j < skv; j++) sum = __fadd_rn(sum, sSt[j * kLdSt + t]);  // ascending order: part of the result
    sStat[t] = sum;
  }
  __syncthreads();
  if (er < rows) {
    const float sum = sStat[er];
    for (int j = ej0; j < skv; j += kAttThreads / kAttRows) sSt[j * kLdSt + er] = __fdiv_rn(sSt[j * kLdSt + er], sum);
  }
  __syncthreads();

  // ---- out = P V: 4 rows x 4 columns per thread, j ascending ----
  {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[i][c] = 0.0f;
#pragma unroll 8
    for (int j = 0; j < skv; j++) {
      const float4 v = *reinterpret_cast<const float4 *>(sKV + j * kAttDim + tx * 4);
      // this thread's four rows of key j in one 16-byte load (four 4-byte loads before the tile was transposed: the phase was
      // bound by shared-memory wavefronts, six per sixteen FMAs; now three)
      const float4 p4 = *reinterpret_cast<const float4 *>(sSt + j * kLdSt + ty * 4);
      const float pr[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const float p = pr[i];
        acc[i][0] = __fmaf_rn(p, v.x, acc[i][0]);
        acc[i][1] = __fmaf_rn(p, v.y, acc[i][1]);
        acc[i][2] = __fmaf_rn(p, v.z, acc[i][2]);
        acc[i][3] = __fmaf_rn(p, v.w, acc[i][3]);
      }
    }
    const bool pad = (skv % 32) != 0;
    if (tx * 4 < d_v) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int r = ty * 4 + i;
        if (r >= rows) continue;
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (pad) o = make_float4(__fadd_rn(o.x, 0.0f), __fadd_rn(o.y, 0.0f), __fadd_rn(o.z, 0.0f), __fadd_rn(o.w, 0.0f));
        *reinterpret_cast<float4 *>(O + (int64_t)r * ldo + tx * 4) = o;
      }
    }
  }
}

inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Fused scores -> softmax -> P.V for `batch` sequences x `heads` heads.  Q [batch*sq, ldq] with head h at columns
// h*d_k.., K / V [batch*skv, ldkv] with head h at columns h*d_k.. / v_col0 + h*d_v.. (relative to K), out [batch*sq, ldo] with head h
// at columns h*d_v...  Returns QG_ENOTSUP when the shape is outside what one CTA's shared memory takes (the caller then runs
// the three kernels): skv <= 128, d_k, d_v <= 64 and multiples of 4, 16-byte aligned rows.
int attention_core(const float *Q, int64_t ldq, const float *K, const float *V, int64_t ldkv, float *O, int64_t ldo, int batch,
                   int heads, int sq, int skv, int d_k, int d_v, float scale, cudaStream_t st) {
  static const bool off = getenv("QG_NO_FUSED_ATTENTION") != nullptr;
  if (off || skv > kAttKeys || d_k > kAttDim || d_v > kAttDim || (d_k & 3) || (d_v & 3) || (ldq & 3) || (ldkv & 3) || (ldo & 3) ||
      !al16(Q) || !al16(K) || !al16(V) || !al16(O) || (int64_t)batch * heads > 65535)
    return QG_ENOTSUP;
  static bool opted[kMaxDevices] = {};
  cudaError_t e = smem_optin(attention_core_kernel, (int)kAttSmem, opted);
  if (e != cudaSuccess) return (int)e;
  MmBatch bt;
  bt.n_outer = batch; bt.n_inner = heads;
  bt.a_outer = (int64_t)sq * ldq;   bt.a_inner = d_k;
  bt.b_outer = (int64_t)skv * ldkv; bt.b_inner = d_k;
  bt.c_outer = (int64_t)sq * ldo;   bt.c_inner = d_v;
  // V is addressed from its own base pointer: per head it advances by d_v columns (c_inner), per sequence like K (b_outer)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1, (unsigned)ceil_div(sq, kAttRows), (unsigned)(batch * heads));
  cfg.blockDim = dim3(kAttThreads);
  cfg.dynamicSmemBytes = kAttSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  return (int)cudaLaunchKernelEx(&cfg, attention_core_kernel, Q, ldq, K, ldkv, V, ldkv, O, ldo, sq, skv, d_k, d_v, scale, bt);
}

}  // namespace qg
