// capi.cu -- the extern "C" boundary declared in include/qgemm.h.
//
// Argument checking, workspace carving and kernel selection live here; the kernels are in
// quantize.cu (a1-a4), gemm_i8_tc.cu (a5-a8, a10 fused, tcgen05) and gemm_simt.cu (generic paths).
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace qg {

// ---- kernels / launchers defined in the other translation units ----
int quant_rows(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, const float *sx,
               int8_t *Xq, int64_t ldq, float *Cx, cudaStream_t st, RowMaxIo io = RowMaxIo());
int quant_cols(const void *W, int dtype, int K, int N, int64_t ldw, float range, int mode, const float *sw,
               int8_t *Wq, int64_t ldq, float *Cw, bool transpose, cudaStream_t st);
int quant_rows_cols_fused(const void *X, const void *W, int dtype, int M, int N, int K, int64_t ldx, int64_t ldw, float range, int mode,
                          int8_t *Xq, int64_t ldxq, float *Cx, int8_t *Wq, int64_t ldwq, float *Cw, cudaStream_t st);
int inv_divide(const float *a, int64_t n, float b, float *out, cudaStream_t st);
int outlier_mask(const float *A, int M, int K, int64_t lda, float thr, float *mask, int64_t ldm, cudaStream_t st);
int reduce_partials(const void *slots, int64_t slot_stride, int n_slots, int part_dtype, int64_t ld_part, const float *bias,
                    void *out, void *const *peers, int n_peers, int64_t ldo, int out_dtype, int M, int N, cudaStream_t st,
                    void *out_mc = nullptr, int max_ctas = 0);
int gemm_s8_simt(const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int b_kmajor, int M, int N, int K, void *O,
                 int64_t ldo, int out_dtype, const float *Cx, const float *Cw, const float *bias, float c,
                 const SideArgs *side, cudaStream_t st, int act = QG_ACT_NONE);
int outlier_detect(const void *X, int dtype, int M, int K, int64_t ldx, float thr, uint32_t *mask, cudaStream_t st);
int outlier_index(const uint32_t *mask, int K, int *idx, int max_idx, int *count, int *wbase, cudaStream_t st);
int outlier_mask_from_idx(const int *idx, int n_idx, int K, uint32_t *mask, int *wbase, cudaStream_t st);
int quant_rows_outlier(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, int8_t *Xq, int64_t ldq,
                       float *Cx, const uint32_t *mask, const int *wbase, void *Xo, int64_t ldxo, int side_bf16,
                       cudaStream_t st);
int gather_wo(const void *W, int dtype, int64_t ldw, const uint32_t *mask, const int *wbase, int K, int no_pad, int N, void *Wo,
              int64_t ldwo, int side_bf16, cudaStream_t st);
int mm_f32(const float *A, int64_t sa_h, int64_t sa_w, const float *B, int64_t sb_h, int64_t sb_w, int M, int N, int K,
           float *C, int64_t ldc, cudaStream_t st, const MmBatch *batch = nullptr);
int softmax_rows(const float *A, int64_t lda, int M, int N, float scale, float *B, int64_t ldb, cudaStream_t st);
int attention_core(const float *Q, int64_t ldq, const float *K, const float *V, int64_t ldkv, float *O, int64_t ldo, int batch,
                   int heads, int sq, int skv, int d_k, int d_v, float scale, cudaStream_t st);
int add_layernorm_rows(const float *A, int64_t lda, const float *R, int64_t ldr, int M, int N, float *B, int64_t ldb,
                       cudaStream_t st, int8_t *Xq = nullptr, int64_t ldq = 0, float *Cx = nullptr, float range = 127.0f,
                       int mode = QG_MODE_REF_EXACT);
int dequantize_s32(const int32_t *acc, int64_t ldacc, const float *Cx, const float *Cw, const float *bias, int M, int N,
                   float c, void *O, int out_dtype, int64_t ldo, cudaStream_t st);
bool gemm_i8_tc_supported(const void *A, int64_t lda, const void *B, int64_t ldb);
void gemm_i8_tc_set_stats(long long *dev_ptr);
int gemm_i8_tc(int cg, const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int b_kmajor, int M, int N, int K,
               void *O, int64_t ldo, int out_kind, const float *Cx, const float *Cw, const float *bias, float c,
               const SideArgs *side, const MultiOut *multi, int num_sms, cudaStream_t st, int act = QG_ACT_NONE,
               int split_k = 1, float *rowmax = nullptr);
int elemwise(int op, const void *A, int64_t lda, const float *B, int64_t ldb, int bmode, float c, float *O, int64_t ldo, int M,
             int N, cudaStream_t st);
int splitk_reduce(const int32_t *parts, int64_t slice_stride, int slices, int64_t ldp, const float *Cx, const float *Cw,
                  const float *bias, int M, int N, float c, int act, void *O, int out_dtype, int64_t ldo, cudaStream_t st);

// ---- error state ----
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return QG_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return (int)e;
}

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// programmatic dependent launch between the kernels of one call; QG_PDL=0 switches it off
bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("QG_PDL");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}

// ---- device state ----
struct DeviceState {
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  bool ok = false;
  void *arena = nullptr;  // grow-only scratch used when the caller passes workspace == NULL
  size_t arena_bytes = 0;
  // host-buffer entry point
  void *hx = nullptr, *hw = nullptr, *ho = nullptr, *hb = nullptr;
  size_t hx_bytes = 0, hw_bytes = 0, ho_bytes = 0, hb_bytes = 0;
  void *attn = nullptr;  // grow-only scratch of qg_attention_forward (projections, scores)
  size_t attn_bytes = 0;
  void *splitk = nullptr;  // grow-only scratch: int32 partial sums of split-K products
  size_t splitk_bytes = 0;
  // three streams: host->device copies, kernels, device->host copies (PCIe is full duplex)
  cudaStream_t hstream = nullptr, hstream_in = nullptr, hstream_out = nullptr;
  static constexpr int kHostChunksMax = 32;
  int host_chunks = 8;  // QG_HOST_CHUNKS
  cudaEvent_t ev_w = nullptr, ev_x[kHostChunksMax] = {}, ev_o[kHostChunksMax] = {};
};
static DeviceState g_dev[16];
static std::mutex g_mu;
static std::atomic<int> g_variant{QG_GEMM_AUTO};
std::atomic<int> g_sm_limit{0};  // qg_set_gemm_sm_limit

static int device_state(DeviceState **out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return QG_ENODEV;
  }
  if (dev < 0 || dev >= 16) { set_error("device ordinal %d out of range", dev); return QG_ENODEV; }
  DeviceState &d = g_dev[dev];
  if (!d.ok) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!d.ok) {
      QG_CUDA_OK(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
      QG_CUDA_OK(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
      QG_CUDA_OK(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
      if (d.cc_major != 10) {
        set_error("libqgemm is built for sm_100a only; device %d is sm_%d%d (no fallback path)", dev, d.cc_major,
                  d.cc_minor);
        return QG_ENODEV;
      }
      const char *env = getenv("QG_GEMM_VARIANT");
      if (env != nullptr) g_variant.store(atoi(env));
      d.ok = true;
    }
  }
  *out = &d;
  return QG_OK;
}

static int grow(void **p, size_t *have, size_t want) {
  if (*have >= want) return QG_OK;
  if (*p != nullptr) { QG_CUDA_OK(cudaDeviceSynchronize()); QG_CUDA_OK(cudaFree(*p)); *p = nullptr; *have = 0; }
  want = (size_t)round_up((int64_t)want, 1 << 20);
  QG_CUDA_OK(cudaMalloc(p, want));
  *have = want;
  return QG_OK;
}

// layout of the per-call weight codes: 0 (default) = the reference's [K,N] (row-major quantizer pass,
// MN-major GEMM operand), 1 = transposed (transposing quantizer pass, K-major operand).  Measured on
// B200 at 4096^3 / 8192^3: 98.1 / 520.6 us against 100.3 / 524.5 us.
static bool percall_kmajor() {
  static const bool on = [] {
    const char *e = getenv("QG_PERCALL_KMAJOR");
    return e != nullptr && atoi(e) != 0;
  }();
  return on;
}

static bool valid_io(int dt) { return dt == QG_F32 || dt == QG_F16 || dt == QG_BF16; }

// layout of the scratch block shared by qg_quantized_mm / qg_linear_forward
struct Workspace {
  int8_t *Xq, *Wq;
  float *Cx, *Cw;
  int64_t ldxq, ldwq;
  void *splitk;         // int32 partial sums of a split-K product (NULL when this shape never splits)
  size_t splitk_bytes;
  size_t bytes;
};
static size_t splitk_need(int M, int N, int K);
static Workspace carve(void *base, int M, int N, int K) {
  Workspace w;
  w.ldxq = round_up(K, 16);  // TMA: leading dimensions are multiples of 16 bytes
  w.ldwq = round_up(N, 16);  // per-call weight codes Wq [K, ldwq] (the reference's layout)
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (size_t)round_up((int64_t)(off + n), 256); return o; };
  // the weight-code region fits either layout: Wq [K, ceil16(N)] or Wt [N, ceil16(K)]
  const size_t wq_bytes = (size_t)K * w.ldwq > (size_t)N * w.ldxq ? (size_t)K * w.ldwq : (size_t)N * w.ldxq;
  const size_t oxq = take((size_t)M * w.ldxq), owq = take(wq_bytes);
  const size_t ocx = take(sizeof(float) * M), ocw = take(sizeof(float) * N);
  w.splitk_bytes = splitk_need(M, N, K);
  const size_t osk = take(w.splitk_bytes);
  char *b = reinterpret_cast<char *>(base);
  w.Xq = reinterpret_cast<int8_t *>(b + oxq);
  w.Wq = reinterpret_cast<int8_t *>(b + owq);
  w.Cx = reinterpret_cast<float *>(b + ocx);
  w.Cw = reinterpret_cast<float *>(b + ocw);
  w.splitk = w.splitk_bytes ? b + osk : nullptr;
  w.bytes = off;
  return w;
}

// Split-K factor for tile-starved shapes.  Tiles are (128*cg) x 256; with T tiles on P clusters a product
// takes ceil(T/P) rounds however few clusters the last one fills.  Cutting K into s slices gives
// ceil(T*s/P) shorter rounds plus a reduce pass over s int32 matrices.  Time model (us, measured on
// B200): a tile costs 11.5 per 4096 of K plus 1.5; the reduce pass 3 + bytes / 5 TB/s.  Taken only when
// the model promises more than 25 % (it pays for few tiles and long K -- M <= 256 against K >= 8192:
// 59.5 -> 31.5 us at 128 x 4096 x 16384 -- and loses elsewhere: 4096 x 1152 x 4096 went 38 -> 63 us
// under a rounds-only model).  QG_SPLIT_K=n forces n (1 = off).
static int choose_split_k(int sm_count, int cg, int M, int N, int K, int out_kind) {
  static const int forced = [] { const char *e = getenv("QG_SPLIT_K"); return e ? atoi(e) : 0; }();
  const int num_kb = (int)ceil_div(K, 128);
  if (forced >= 1) return (forced <= num_kb && forced <= kMaxExtraOut + 1 && (forced - 1) * (int)ceil_div(num_kb, forced) < num_kb) ? forced : 1;
  const int64_t T = ceil_div(M, 128 * cg) * ceil_div(N, 256);
  const int P = sm_count / cg;
  const double osz = (out_kind == QG_F16 || out_kind == QG_BF16) ? 2.0 : 4.0;
  auto cost = [&](int s) {
    const double gemm = (double)ceil_div(T * s, P) * (11.5 * K / s / 4096.0 + 1.5);
    const double reduce = s == 1 ? 0.0 : 3.0 + ((double)M * N * (8.0 * s + osz)) / 5e6;
    return gemm + reduce;
  };
  const double base = cost(1);
  double best = base;
  int pick = 1;
  for (int s = 2; s <= 4; s++) {
    if (num_kb / s < 8) break;
    const double c = cost(s);
    if (c < best) { best = c; pick = s; }
  }
  return best < 0.75 * base ? pick : 1;
}

// Bytes of int32 slice matrices the split-K form of an M x N x K product may ask for (0: it never splits).
// The split factor depends on the SM count (148 on every B200; the current device is asked when there is
// one) and, weakly, on the output width, so the larger of the fp32 / 16-bit answers is reserved.
static size_t splitk_need(int M, int N, int K) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v;
  } else {
    cudaGetLastError();
  }
  const int cg = M > 128 ? 2 : 1;
  const int sk = std::max(choose_split_k(sms, cg, M, N, K, QG_F32), choose_split_k(sms, cg, M, N, K, QG_F16));
  return sk > 1 ? sizeof(int32_t) * (size_t)sk * M * (size_t)round_up(N, 4) : 0;
}

// b_kmajor == 0: B is the reference's [K,N] (MN-major tensor-core operand); 1: B is Wt [N,K]
static int gemm_dispatch(DeviceState *d, const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int b_kmajor, int M,
                         int N, int K, void *O, int64_t ldo, int out_kind, const float *Cx, const float *Cw,
                         const float *bias, float c, cudaStream_t st, const SideArgs *side = nullptr,
                         const MultiOut *multi = nullptr, int act = QG_ACT_NONE, void *sk_buf = nullptr,
                         size_t sk_bytes = 0, bool sk_arena = true, float *rowmax = nullptr, bool *rowmax_done = nullptr) {
  // rowmax (optional): the tensor-core epilogue raises rowmax[i] to max_{j>=1} |y[i,j]| (RowMaxIo); *rowmax_done tells
  // the caller whether this path did it (the CUDA-core and split-K forms do not)
  if (rowmax_done) *rowmax_done = false;
  int variant = g_variant.load();
  const bool tc_ok = gemm_i8_tc_supported(A, lda, B, ldb);
  // the 2-SM tile (256x256 per CTA pair) halves the shared-memory traffic per MAC; one CTA row
  // of work is all a problem with M <= 128 has, so it takes the 1-SM kernel
  if (variant == QG_GEMM_AUTO) variant = !tc_ok ? QG_GEMM_SIMT : (M > 128 ? QG_GEMM_TC_2SM : QG_GEMM_TC_1SM);
  if (variant == QG_GEMM_SIMT || !tc_ok) {
    if (multi != nullptr && (multi->n > 0 || multi->mc != nullptr)) {
      set_error("extra destinations need the tensor-core path (16-byte aligned operands)");
      return QG_ENOTSUP;
    }
    return gemm_s8_simt(A, lda, B, ldb, b_kmajor, M, N, K, O, ldo, out_kind, Cx, Cw, bias, c, side, st, act);
  }
  const int cg = variant == QG_GEMM_TC_2SM ? 2 : 1;
  const int lim = g_sm_limit.load();
  const int sm_count = (lim > 0 && lim < d->sm_count) ? std::max(lim, 2) : d->sm_count;
  int sk = (side == nullptr && (multi == nullptr || (multi->n == 0 && multi->mc == nullptr))) ? choose_split_k(sm_count, cg, M, N, K, out_kind) : 1;
  // split-K: int32 partial sums of every k-slice, then one pass that adds them and runs the epilogue.
  // The slices live in the caller's workspace (qg_workspace_bytes reserves them); entry points without a
  // workspace argument use a grow-only per-device buffer (sk_arena; documented as shared in qgemm.h).  A
  // caller-provided block that is too small simply does not split.
  const int64_t ldp = round_up(N, 4);
  const size_t slice = (size_t)M * ldp;
  int32_t *parts = nullptr;
  if (sk > 1) {
    if (sk_buf != nullptr) {
      while (sk > 1 && sizeof(int32_t) * slice * sk > sk_bytes) sk--;
      if (sk > 1 && (sk - 1) * (int)ceil_div(ceil_div(K, 128), sk) >= (int)ceil_div(K, 128)) sk = 1;  // an empty slice
      parts = (int32_t *)sk_buf;
    } else if (sk_arena) {
      std::lock_guard<std::mutex> lk(g_mu);
      int rc = grow(&d->splitk, &d->splitk_bytes, sizeof(int32_t) * slice * sk);
      if (rc) return rc;
      parts = (int32_t *)d->splitk;
    } else {
      sk = 1;
    }
  }
  if (sk > 1) {
    MultiOut slices = {};
    slices.n = sk - 1;
    for (int i = 1; i < sk; i++) slices.dst[i - 1] = parts + (size_t)i * slice;
    int rc = gemm_i8_tc(cg, A, lda, B, ldb, b_kmajor, M, N, K, parts, ldp, QG_S32, nullptr, nullptr, nullptr, 0.0f, nullptr,
                        &slices, sm_count, st, QG_ACT_NONE, sk);
    if (rc) return rc;
    return cuda_status((cudaError_t)splitk_reduce(parts, (int64_t)slice, sk, ldp, Cx, Cw, bias, M, N, c, act, O, out_kind, ldo, st),
                       "split-K reduce");
  }
  if (rowmax_done) *rowmax_done = rowmax != nullptr && out_kind != QG_S32;
  return gemm_i8_tc(cg, A, lda, B, ldb, b_kmajor, M, N, K, O, ldo, out_kind, Cx, Cw, bias, c, side, multi, sm_count, st, act, 1,
                    rowmax);
}

}  // namespace qg

using namespace qg;

extern "C" {

int qg_version(void) { return 100; }
const char *qg_last_error(void) { return g_err; }

int qg_device_info(int *sm_count, int *cc_major, int *cc_minor) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  if (sm_count) *sm_count = d->sm_count;
  if (cc_major) *cc_major = d->cc_major;
  if (cc_minor) *cc_minor = d->cc_minor;
  return QG_OK;
}

/* SMs the tensor-core GEMM may occupy (0 = all).  A persistent GEMM takes every SM it is given; a caller that overlaps its own
 * communication kernels with the GEMM (megatron.MegatronFFN: the ordered reduce + gather of one row block under the next
 * block's products) leaves a few SMs free for them, as communication libraries reserve channels. */
int qg_set_gemm_sm_limit(int sms) {
  if (sms < 0) { set_error("bad SM limit %d", sms); return QG_EINVAL; }
  g_sm_limit.store(sms);
  return QG_OK;
}

int qg_set_gemm_variant(int variant) {
  if (variant < QG_GEMM_AUTO || variant > QG_GEMM_TC_2SM) { set_error("bad gemm variant %d", variant); return QG_EINVAL; }
  g_variant.store(variant);
  return QG_OK;
}

int64_t qg_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int qg_absmax_rows(const void *X, int dtype, int M, int K, int64_t ldx, int mode, float *Cx, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && Cx && M > 0 && K > 0 && ldx >= K && valid_io(dtype), "qg_absmax_rows: bad arguments");
  return quant_rows(X, dtype, M, K, ldx, 127.0f, mode, nullptr, nullptr, 0, Cx, (cudaStream_t)stream);
}

int qg_absmax_cols(const void *W, int dtype, int K, int N, int64_t ldw, int mode, float *Cw, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(W && Cw && K > 0 && N > 0 && ldw >= N && valid_io(dtype), "qg_absmax_cols: bad arguments");
  return quant_cols(W, dtype, K, N, ldw, 127.0f, mode, nullptr, nullptr, 0, Cw, false, (cudaStream_t)stream);
}

int qg_inv_divide_f32(const float *a, int64_t n, float b, float *out, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(a && out && n > 0, "qg_inv_divide_f32: bad arguments");
  return inv_divide(a, n, b, out, (cudaStream_t)stream);
}

int qg_quantize_rows(const void *X, int dtype, int M, int K, int64_t ldx, const float *sx, int8_t *Xq, int64_t ldq,
                     qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && sx && Xq && M > 0 && K > 0 && ldx >= K && ldq >= K && valid_io(dtype), "qg_quantize_rows: bad arguments");
  return quant_rows(X, dtype, M, K, ldx, 127.0f, 0, sx, Xq, ldq, nullptr, (cudaStream_t)stream);
}

int qg_quantize_cols(const void *W, int dtype, int K, int N, int64_t ldw, const float *sw, int8_t *Wq, int64_t ldq,
                     qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(W && sw && Wq && K > 0 && N > 0 && ldw >= N && ldq >= N && valid_io(dtype), "qg_quantize_cols: bad arguments");
  return quant_cols(W, dtype, K, N, ldw, 127.0f, 0, sw, Wq, ldq, nullptr, false, (cudaStream_t)stream);
}

int qg_absmax_quant_rows(const void *X, int dtype, int M, int K, int64_t ldx, float range, int mode, int8_t *Xq,
                         int64_t ldq, float *Cx, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && Xq && Cx && M > 0 && K > 0 && ldx >= K && ldq >= K && valid_io(dtype),
             "qg_absmax_quant_rows: bad arguments");
  return quant_rows(X, dtype, M, K, ldx, range, mode, nullptr, Xq, ldq, Cx, (cudaStream_t)stream);
}

int qg_absmax_quant_cols(const void *W, int dtype, int K, int N, int64_t ldw, float range, int mode, int8_t *Wq,
                         int64_t ldq, float *Cw, float *scratch, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(W && Wq && Cw && K > 0 && N > 0 && ldw >= N && ldq >= N && valid_io(dtype),
             "qg_absmax_quant_cols: bad arguments");
  (void)scratch;  // kept in the ABI; the library owns the column-maximum scratch (per device and stream)
  return quant_cols(W, dtype, K, N, ldw, range, mode, nullptr, Wq, ldq, Cw, false, (cudaStream_t)stream);
}

int qg_gemm_s8s8s32(const int8_t *A, int64_t lda, const int8_t *B, int64_t ldb, int M, int N, int K, int32_t *C,
                    int64_t ldc, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && lda >= K && ldb >= N && ldc >= N, "qg_gemm_s8s8s32: bad arguments");
  return gemm_dispatch(d, A, lda, B, ldb, 0, M, N, K, C, ldc, QG_S32, nullptr, nullptr, nullptr, 0.0f,
                       (cudaStream_t)stream);
}

int qg_dequantize_s32(const int32_t *acc, int64_t ldacc, const float *Cx, const float *Cw, const float *bias, int M,
                      int N, float range, void *O, int out_dtype, int64_t ldo, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(acc && Cx && Cw && O && M > 0 && N > 0 && ldacc >= N && ldo >= N && valid_io(out_dtype),
             "qg_dequantize_s32: bad arguments");
  return dequantize_s32(acc, ldacc, Cx, Cw, bias, M, N, 1 / (range * range), O, out_dtype, ldo, (cudaStream_t)stream);
}

int qg_gemm_s8_dequant(const int8_t *Xq, int64_t ldxq, const int8_t *Wq, int64_t ldwq, const float *Cx,
                       const float *Cw, const float *bias, int M, int N, int K, float range, void *O, int out_dtype,
                       int64_t ldo, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && Wq && Cx && Cw && O && M > 0 && N > 0 && K > 0 && ldxq >= K && ldwq >= N && ldo >= N &&
                 valid_io(out_dtype),
             "qg_gemm_s8_dequant: bad arguments");
  return gemm_dispatch(d, Xq, ldxq, Wq, ldwq, 0, M, N, K, O, ldo, out_dtype, Cx, Cw, bias, 1 / (range * range),
                       (cudaStream_t)stream);
}

size_t qg_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return carve(nullptr, M, N, K).bytes;
}

static int get_workspace(DeviceState *d, void *workspace, size_t workspace_bytes, int M, int N, int K, Workspace *w) {
  const size_t need = carve(nullptr, M, N, K).bytes;
  if (workspace == nullptr) {
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = grow(&d->arena, &d->arena_bytes, need);
    if (rc) return rc;
    workspace = d->arena;
  } else if (workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need);
    return QG_ENOMEM;
  } else if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) {
    set_error("workspace must be 256-byte aligned");
    return QG_EINVAL;
  }
  *w = carve(workspace, M, N, K);
  return QG_OK;
}

int qg_quantized_mm(const void *X, int64_t ldx, const void *W, int64_t ldw, int in_dtype, void *O, int64_t ldo,
                    int out_dtype, int M, int N, int K, float range, int mode, const float *bias, void *workspace,
                    size_t workspace_bytes, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && W && O && M > 0 && N > 0 && K > 0 && ldx >= K && ldw >= N && ldo >= N && valid_io(in_dtype) &&
                 valid_io(out_dtype),
             "qg_quantized_mm: bad arguments (M=%d N=%d K=%d)", M, N, K);
  Workspace w;
  rc = get_workspace(d, workspace, workspace_bytes, M, N, K, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // per-call weight quantization; QG_PERCALL_KMAJOR=0 keeps the reference's [K,N] code layout (MN-major operand)
  const bool kmajor = percall_kmajor();
  // large problems: column pass 1, then column pass 2 (L2-bound) side by side with the row quantizer (HBM-bound) in one launch
  rc = kmajor ? -1 : quant_rows_cols_fused(X, W, in_dtype, M, N, K, ldx, ldw, range, mode, w.Xq, w.ldxq, w.Cx, w.Wq, w.ldwq, w.Cw, st);
  if (rc > 0) return cuda_status((cudaError_t)rc, "fused quantizers");
  if (rc < 0) {
    rc = quant_rows(X, in_dtype, M, K, ldx, range, mode, nullptr, w.Xq, w.ldxq, w.Cx, st);
    if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
    rc = quant_cols(W, in_dtype, K, N, ldw, range, mode, nullptr, w.Wq, kmajor ? w.ldxq : w.ldwq, w.Cw, kmajor, st);
    if (rc) return cuda_status((cudaError_t)rc, "column quantizer");
  }
  return gemm_dispatch(d, w.Xq, w.ldxq, w.Wq, kmajor ? w.ldxq : w.ldwq, kmajor ? 1 : 0, M, N, K, O, ldo, out_dtype, w.Cx,
                       w.Cw, bias, 1 / (range * range), st, nullptr, nullptr, QG_ACT_NONE, w.splitk, w.splitk_bytes, false);
}

/* Both quantizers of the op in one call (a1-a4 for X and for W): the same codes and scales as qg_absmax_quant_rows +
 * qg_absmax_quant_cols; large problems run column pass 1, then column pass 2 side by side with the row quantizer. */
int qg_absmax_quant_rows_cols(const void *X, int64_t ldx, const void *W, int64_t ldw, int dtype, int M, int N, int K, float range,
                              int mode, int8_t *Xq, int64_t ldxq, float *Cx, int8_t *Wq, int64_t ldwq, float *Cw,
                              qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && W && Xq && Cx && Wq && Cw && M > 0 && N > 0 && K > 0 && ldx >= K && ldw >= N && ldxq >= K && ldwq >= N &&
                 valid_io(dtype),
             "qg_absmax_quant_rows_cols: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  rc = quant_rows_cols_fused(X, W, dtype, M, N, K, ldx, ldw, range, mode, Xq, ldxq, Cx, Wq, ldwq, Cw, st);
  if (rc > 0) return cuda_status((cudaError_t)rc, "fused quantizers");
  if (rc == 0) return QG_OK;
  rc = quant_rows(X, dtype, M, K, ldx, range, mode, nullptr, Xq, ldxq, Cx, st);
  if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
  rc = quant_cols(W, dtype, K, N, ldw, range, mode, nullptr, Wq, ldwq, Cw, false, st);
  return rc ? cuda_status((cudaError_t)rc, "column quantizer") : QG_OK;
}

int qg_prepare_weights(const void *W, int dtype, int K, int N, int64_t ldw, float range, int mode, int8_t *Wt,
                       int64_t ldwt, float *Cw, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(W && Wt && Cw && K > 0 && N > 0 && ldw >= N && ldwt >= K && valid_io(dtype), "qg_prepare_weights: bad arguments");
  return quant_cols(W, dtype, K, N, ldw, range, mode, nullptr, Wt, ldwt, Cw, true, (cudaStream_t)stream);
}

int qg_gemm_s8t_dequant(const int8_t *Xq, int64_t ldxq, const int8_t *Wt, int64_t ldwt, const float *Cx, const float *Cw,
                        const float *bias, int M, int N, int K, float range, void *O, int out_dtype, int64_t ldo,
                        qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && Wt && O && M > 0 && N > 0 && K > 0 && ldxq >= K && ldwt >= K && ldo >= N &&
                 (out_dtype == QG_S32 || (valid_io(out_dtype) && Cx && Cw)),
             "qg_gemm_s8t_dequant: bad arguments");
  return gemm_dispatch(d, Xq, ldxq, Wt, ldwt, 1, M, N, K, O, ldo, out_dtype, Cx, Cw, bias, 1 / (range * range),
                       (cudaStream_t)stream);
}

int qg_linear_forward(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt, const float *Cw,
                      const float *bias, void *Y, int64_t ldy, int out_dtype, int M, int N, int K, float range, int mode,
                      void *workspace, size_t workspace_bytes, qg_stream_t stream) {
  return qg_linear_forward_act(X, ldx, in_dtype, Wt, ldwt, Cw, bias, QG_ACT_NONE, Y, ldy, out_dtype, M, N, K, range, mode,
                               workspace, workspace_bytes, stream);
}

int qg_linear_forward_act(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt, const float *Cw,
                          const float *bias, int act, void *Y, int64_t ldy, int out_dtype, int M, int N, int K, float range,
                          int mode, void *workspace, size_t workspace_bytes, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && Wt && Cw && Y && M > 0 && N > 0 && K > 0 && ldx >= K && ldwt >= K && ldy >= N && valid_io(in_dtype) &&
                 valid_io(out_dtype),
             "qg_linear_forward: bad arguments (M=%d N=%d K=%d)", M, N, K);
  QG_REQUIRE(act == QG_ACT_NONE || act == QG_ACT_RELU, "qg_linear_forward: unknown activation %d", act);
  Workspace w;
  rc = get_workspace(d, workspace, workspace_bytes, M, N, K, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = quant_rows(X, in_dtype, M, K, ldx, range, mode, nullptr, w.Xq, w.ldxq, w.Cx, st);
  if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
  return gemm_dispatch(d, w.Xq, w.ldxq, Wt, ldwt, 1, M, N, K, Y, ldy, out_dtype, w.Cx, Cw, bias, 1 / (range * range), st,
                       nullptr, nullptr, act, w.splitk, w.splitk_bytes, false);
}

// ---- outlier decomposition ------------------------------------------------------------------
static const int kMaxOutlierCols = 64;  // what the fused epilogue takes (gemm_i8_tc.cu: kSideMax)

struct OutlierWs {
  uint32_t *mask;
  int *wbase;
  void *Xo, *Wo;
  int64_t ldxo, ldwo;
  size_t bytes;
};
static OutlierWs carve_outlier(void *base, int M, int N, int K) {
  OutlierWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (size_t)round_up((int64_t)(off + n), 256); return o; };
  const size_t words = (size_t)ceil_div(K, 32);
  w.ldxo = kMaxOutlierCols;
  w.ldwo = round_up(N, 8);
  const size_t om = take(4 * words), ob = take(4 * (words + 1)), ox = take(2 * (size_t)M * w.ldxo),
               ow = take(2 * (size_t)kMaxOutlierCols * w.ldwo);
  char *b = reinterpret_cast<char *>(base);
  w.mask = reinterpret_cast<uint32_t *>(b + om);
  w.wbase = reinterpret_cast<int *>(b + ob);
  w.Xo = b + ox;
  w.Wo = b + ow;
  w.bytes = off;
  return w;
}

int qg_outlier_cols(const void *X, int dtype, int M, int K, int64_t ldx, float thr, int *idx, int max_idx, int *count,
                    qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && count && M > 0 && K > 0 && ldx >= K && max_idx >= 0 && valid_io(dtype), "qg_outlier_cols: bad arguments");
  const size_t words = (size_t)ceil_div(K, 32);
  uint32_t *mask;
  int *wbase;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if ((rc = grow(&d->arena, &d->arena_bytes, 8 * words + 512))) return rc;
    mask = reinterpret_cast<uint32_t *>(d->arena);
    wbase = reinterpret_cast<int *>(reinterpret_cast<char *>(d->arena) + round_up((int64_t)(4 * words), 256));
  }
  cudaStream_t st = (cudaStream_t)stream;
  rc = outlier_detect(X, dtype, M, K, ldx, thr, mask, st);
  if (rc) return cuda_status((cudaError_t)rc, "outlier detection");
  rc = outlier_index(mask, K, idx, max_idx, count, wbase, st);
  return rc ? cuda_status((cudaError_t)rc, "outlier index") : QG_OK;
}

size_t qg_outlier_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return carve(nullptr, M, 1, K).bytes + carve_outlier(nullptr, M, N, K).bytes;
}

int qg_linear_forward_outlier(const void *X, int64_t ldx, int in_dtype, const void *W, int64_t ldw, int w_dtype,
                              const int8_t *Wt, int64_t ldwt, const float *Cw, const float *bias, const int *idx, int n_idx,
                              void *Y, int64_t ldy, int out_dtype, int M, int N, int K, float range, int mode,
                              void *workspace, size_t workspace_bytes, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && W && Wt && Cw && Y && M > 0 && N > 0 && K > 0 && ldx >= K && ldw >= N && ldwt >= K && ldy >= N &&
                 valid_io(in_dtype) && valid_io(w_dtype) && valid_io(out_dtype) && n_idx >= 0 && (n_idx == 0 || idx),
             "qg_linear_forward_outlier: bad arguments");
  if (n_idx > kMaxOutlierCols) {
    set_error("qg_linear_forward_outlier: %d outlier columns, the fused side product takes at most %d", n_idx,
              kMaxOutlierCols);
    return QG_ENOTSUP;
  }
  const size_t need = qg_outlier_workspace_bytes(M, N, K);
  if (workspace == nullptr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if ((rc = grow(&d->arena, &d->arena_bytes, need))) return rc;
    workspace = d->arena;
  } else {
    QG_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "qg_linear_forward_outlier: workspace too small or misaligned (%zu bytes needed)", need);
  }
  Workspace w = carve(workspace, M, 1, K);
  OutlierWs o = carve_outlier(reinterpret_cast<char *>(workspace) + w.bytes, M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  const int no_pad = (int)round_up(n_idx, 8);
  o.ldxo = no_pad <= 16 ? 16 : no_pad;  // rows of Xo as short as this call's outlier count allows (the block is sized for the maximum)
  const int side_bf16 = (in_dtype == QG_BF16) ? 1 : 0;
  rc = outlier_mask_from_idx(idx, n_idx, K, o.mask, o.wbase, st);
  if (rc) return cuda_status((cudaError_t)rc, "outlier mask");
  if (no_pad > 0) {  // (the side operand's padding columns are zeroed by the row quantizer itself: no memset node in the PDL chain)
    rc = gather_wo(W, w_dtype, ldw, o.mask, o.wbase, K, no_pad, N, o.Wo, o.ldwo, side_bf16, st);
    if (rc) return cuda_status((cudaError_t)rc, "outlier weight gather");
  }
  rc = quant_rows_outlier(X, in_dtype, M, K, ldx, range, mode, w.Xq, w.ldxq, w.Cx, o.mask, o.wbase, o.Xo, o.ldxo, side_bf16,
                          st);
  if (rc) return cuda_status((cudaError_t)rc, "masking row quantizer");
  SideArgs side = {o.Xo, o.ldxo, o.Wo, o.ldwo, no_pad, side_bf16};
  return gemm_dispatch(d, w.Xq, w.ldxq, Wt, ldwt, 1, M, N, K, Y, ldy, out_dtype, w.Cx, Cw, bias, 1 / (range * range), st,
                       no_pad > 0 ? &side : nullptr);
}

/* Column-parallel LinearLayer::forward: this rank's [m, n] block (n = its share
 * of the output columns, Wt/Cw/bias = its shard) is written into `y_local` and into the n_peers
 * matrices `y_peers[]` -- peer GPUs' [m, N_total] outputs mapped into this process, every pointer
 * already offset to this rank's first column, all with leading dimension ldy.  The GEMM epilogue
 * issues the peer stores itself (TMA over NVLink), so the gather overlaps the main loop. */
int qg_linear_forward_multi(const void *X, int64_t ldx, int in_dtype, const int8_t *Wt, int64_t ldwt, const float *Cw,
                            const float *bias, void *y_local, void *const *y_peers, int n_peers, int64_t ldy,
                            int out_dtype, int M, int N, int K, float range, int mode, void *workspace,
                            size_t workspace_bytes, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X && Wt && Cw && y_local && M > 0 && N > 0 && K > 0 && ldx >= K && ldwt >= K && ldy >= N && valid_io(in_dtype) &&
                 valid_io(out_dtype) && n_peers >= 0 && n_peers <= kMaxExtraOut && (n_peers == 0 || y_peers),
             "qg_linear_forward_multi: bad arguments (n_peers=%d)", n_peers);
  Workspace w;
  rc = get_workspace(d, workspace, workspace_bytes, M, 1, K, &w);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  rc = quant_rows(X, in_dtype, M, K, ldx, range, mode, nullptr, w.Xq, w.ldxq, w.Cx, st);
  if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
  MultiOut mo = {};
  mo.n = n_peers;
  for (int i = 0; i < n_peers; i++) mo.dst[i] = y_peers[i];
  return gemm_dispatch(d, w.Xq, w.ldxq, Wt, ldwt, 1, M, N, K, y_local, ldy, out_dtype, w.Cx, Cw, bias, 1 / (range * range),
                       st, nullptr, &mo);
}

/* General form of the fused GEMM: B in either layout, optional extra destinations for the epilogue. */
int qg_gemm_s8_dequant_ex(const int8_t *Xq, int64_t ldxq, const int8_t *B, int64_t ldb, int b_kmajor, const float *Cx,
                          const float *Cw, const float *bias, int M, int N, int K, float range, void *O,
                          void *const *peers, int n_peers, int out_dtype, int64_t ldo, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && B && Cx && Cw && O && M > 0 && N > 0 && K > 0 && ldxq >= K && ldb >= (b_kmajor ? K : N) && ldo >= N &&
                 valid_io(out_dtype) && n_peers >= 0 && n_peers <= kMaxExtraOut && (n_peers == 0 || peers),
             "qg_gemm_s8_dequant_ex: bad arguments");
  MultiOut mo = {};
  mo.n = n_peers;
  for (int i = 0; i < n_peers; i++) mo.dst[i] = peers[i];
  return gemm_dispatch(d, Xq, ldxq, B, ldb, b_kmajor ? 1 : 0, M, N, K, O, ldo, out_dtype, Cx, Cw, bias, 1 / (range * range),
                       (cudaStream_t)stream, nullptr, n_peers > 0 ? &mo : nullptr);
}

/* The same with the exchange done by the NVSwitch: O_mc is the multicast (multimem) address of the block O -- one store
 * reaches every GPU's matrix, the caller's own included (O itself is only used for alignment checks). */
int qg_gemm_s8_dequant_mc(const int8_t *Xq, int64_t ldxq, const int8_t *B, int64_t ldb, int b_kmajor, const float *Cx,
                          const float *Cw, const float *bias, int M, int N, int K, float range, void *O, void *O_mc,
                          int out_dtype, int64_t ldo, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && B && Cx && Cw && O && O_mc && M > 0 && N > 0 && K > 0 && ldxq >= K && ldb >= (b_kmajor ? K : N) && ldo >= N &&
                 valid_io(out_dtype),
             "qg_gemm_s8_dequant_mc: bad arguments");
  MultiOut mo = {};
  mo.mc = O_mc;
  return gemm_dispatch(d, Xq, ldxq, B, ldb, b_kmajor ? 1 : 0, M, N, K, O, ldo, out_dtype, Cx, Cw, bias, 1 / (range * range),
                       (cudaStream_t)stream, nullptr, &mo);
}

int qg_quantized_mm_host(const float *X_host, const float *W_host, float *O_host, int M, int N, int K, float range,
                         int mode, const float *bias_host) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(X_host && W_host && O_host && M > 0 && N > 0 && K > 0, "qg_quantized_mm_host: bad arguments");
  std::lock_guard<std::mutex> lk(g_mu);
  if (d->hstream == nullptr) {
    QG_CUDA_OK(cudaStreamCreateWithFlags(&d->hstream, cudaStreamNonBlocking));
    QG_CUDA_OK(cudaStreamCreateWithFlags(&d->hstream_in, cudaStreamNonBlocking));
    QG_CUDA_OK(cudaStreamCreateWithFlags(&d->hstream_out, cudaStreamNonBlocking));
    QG_CUDA_OK(cudaEventCreateWithFlags(&d->ev_w, cudaEventDisableTiming));
    const char *hc = getenv("QG_HOST_CHUNKS");
    if (hc != nullptr && atoi(hc) >= 1 && atoi(hc) <= DeviceState::kHostChunksMax) d->host_chunks = atoi(hc);
    for (int i = 0; i < DeviceState::kHostChunksMax; i++) {
      QG_CUDA_OK(cudaEventCreateWithFlags(&d->ev_x[i], cudaEventDisableTiming));
      QG_CUDA_OK(cudaEventCreateWithFlags(&d->ev_o[i], cudaEventDisableTiming));
    }
  }
  const size_t xb = sizeof(float) * (size_t)M * K, wb = sizeof(float) * (size_t)K * N, ob = sizeof(float) * (size_t)M * N;
  if ((rc = grow(&d->hx, &d->hx_bytes, xb))) return rc;
  if ((rc = grow(&d->hw, &d->hw_bytes, wb))) return rc;
  if ((rc = grow(&d->ho, &d->ho_bytes, ob))) return rc;
  if (bias_host && (rc = grow(&d->hb, &d->hb_bytes, sizeof(float) * (size_t)N))) return rc;
  if ((rc = grow(&d->arena, &d->arena_bytes, carve(nullptr, M, N, K).bytes))) return rc;
  // row chunks: multiples of 256 rows (one 2-SM tile), at most host_chunks of them.  A chunk's product may be
  // tile-starved enough to split K: its slice matrices are grown here, under the lock this function
  // already holds (gemm_dispatch must not take g_mu again), and handed over explicitly.
  int rows_per = (M + d->host_chunks - 1) / d->host_chunks;
  rows_per = ((rows_per + 255) / 256) * 256;
  const size_t sk_need = std::max(splitk_need(std::min(rows_per, M), N, K), splitk_need(M % rows_per ? M % rows_per : 1, N, K));
  if (sk_need && (rc = grow(&d->splitk, &d->splitk_bytes, sk_need))) return rc;
  // The reference's toDevice() / toHost() round trip (tensor.cuh:77-119) as a three-stream pipeline:
  //   in:      W, then X in row chunks                       (host -> device)
  //   compute: column quantizer once W is in; per chunk row quantizer + GEMM/dequantize
  //   out:     each finished row chunk of O                  (device -> host, concurrent with `in`)
  // Row chunk r of O depends on row chunk r of X and all of W, so after W has landed the two PCIe
  // directions run at the same time: about |W| + |X| + one chunk instead of |W| + |X| + |O|.
  cudaStream_t s_in = d->hstream_in, s_k = d->hstream, s_out = d->hstream_out;
  const float *bias_dev = bias_host ? (const float *)d->hb : nullptr;
  Workspace w = carve(d->arena, M, N, K);
  if (bias_host) QG_CUDA_OK(cudaMemcpyAsync(d->hb, bias_host, sizeof(float) * (size_t)N, cudaMemcpyHostToDevice, s_in));
  QG_CUDA_OK(cudaMemcpyAsync(d->hw, W_host, wb, cudaMemcpyHostToDevice, s_in));
  QG_CUDA_OK(cudaEventRecord(d->ev_w, s_in));
  QG_CUDA_OK(cudaStreamWaitEvent(s_k, d->ev_w, 0));
  rc = quant_cols(d->hw, QG_F32, K, N, N, range, mode, nullptr, w.Wq, w.ldwq, w.Cw, false, s_k);
  if (rc) return cuda_status((cudaError_t)rc, "column quantizer");
  const float *Xd = (const float *)d->hx;
  float *Od = (float *)d->ho;
  int ci = 0;
  for (int r0 = 0; r0 < M; r0 += rows_per, ci++) {
    const int rows = std::min(rows_per, M - r0);
    QG_CUDA_OK(cudaMemcpyAsync((void *)(Xd + (size_t)r0 * K), X_host + (size_t)r0 * K, sizeof(float) * (size_t)rows * K,
                               cudaMemcpyHostToDevice, s_in));
    QG_CUDA_OK(cudaEventRecord(d->ev_x[ci], s_in));
    QG_CUDA_OK(cudaStreamWaitEvent(s_k, d->ev_x[ci], 0));
    rc = quant_rows(Xd + (size_t)r0 * K, QG_F32, rows, K, K, range, mode, nullptr, w.Xq + (size_t)r0 * w.ldxq, w.ldxq,
                    w.Cx + r0, s_k);
    if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
    rc = gemm_dispatch(d, w.Xq + (size_t)r0 * w.ldxq, w.ldxq, w.Wq, w.ldwq, 0, rows, N, K, Od + (size_t)r0 * N, N, QG_F32,
                       w.Cx + r0, w.Cw, bias_dev, 1 / (range * range), s_k, nullptr, nullptr, QG_ACT_NONE,
                       sk_need ? d->splitk : nullptr, sk_need ? d->splitk_bytes : 0, false);
    if (rc) return rc;
    QG_CUDA_OK(cudaEventRecord(d->ev_o[ci], s_k));
    QG_CUDA_OK(cudaStreamWaitEvent(s_out, d->ev_o[ci], 0));
    QG_CUDA_OK(cudaMemcpyAsync(O_host + (size_t)r0 * N, Od + (size_t)r0 * N, sizeof(float) * (size_t)rows * N,
                               cudaMemcpyDeviceToHost, s_out));
  }
  QG_CUDA_OK(cudaStreamSynchronize(s_out));  // O_host is complete; `in` and `compute` finished before it
  return QG_OK;
}

int qg_outlier_mask_f32(const float *A, int M, int K, int64_t lda, float thr, float *mask, int64_t ldm,
                        qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && mask && M > 0 && K > 0 && lda >= K && ldm >= K, "qg_outlier_mask_f32: bad arguments");
  return outlier_mask(A, M, K, lda, thr, mask, ldm, (cudaStream_t)stream);
}

int qg_mm_f32(const float *A, int64_t sa_h, int64_t sa_w, const float *B, int64_t sb_h, int64_t sb_w, int M, int N,
              int K, float *C, int64_t ldc, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, "qg_mm_f32: bad arguments");
  return mm_f32(A, sa_h, sa_w, B, sb_h, sb_w, M, N, K, C, ldc, (cudaStream_t)stream);
}

int qg_add_layernorm_f32(const float *A, int64_t lda, const float *R, int64_t ldr, int m, int n, float *B, int64_t ldb,
                         qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && B && m > 0 && n > 0 && lda >= n && ldb >= n && (R == nullptr || ldr >= n), "qg_add_layernorm_f32: bad arguments");
  return cuda_status((cudaError_t)add_layernorm_rows(A, lda, R, ldr, m, n, B, ldb, (cudaStream_t)stream), "add + layernorm");
}

int qg_softmax_rows_f32(const float *A, int64_t lda, int m, int n, float scale, float *B, int64_t ldb,
                        qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && B && m > 0 && n > 0 && lda >= n && ldb >= n, "qg_softmax_rows_f32: bad arguments");
  return cuda_status((cudaError_t)softmax_rows(A, lda, m, n, scale, B, ldb, (cudaStream_t)stream), "softmax");
}

// shared by qg_attention_forward (fp32 weights, quantized on every call like op_quantized_mm) and
// qg_attention_forward_prepared (codes + column scales from qg_prepare_weights): exactly one of Wqkv / Wt is non-NULL
static int attention_forward_impl(const float *Xq, int64_t ldxq, const float *Xkv, int64_t ldxkv, int batch, int sq, int skv,
                                  int d_model, const float *Wqkv, int64_t ldw, const int8_t *Wt, int64_t ldwt, const float *Cw,
                                  int heads, int d_k, int d_v, float range, int mode, float *out, int64_t ldo,
                                  qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  const int nq = heads * d_k, nkv = heads * (d_k + d_v), ntot = nq + nkv;
  QG_REQUIRE(Xq && Xkv && out && batch > 0 && sq > 0 && skv > 0 && d_model > 0 && heads > 0 && d_k > 0 && d_v > 0 &&
                 ldxq >= d_model && ldxkv >= d_model && ldo >= heads * d_v &&
                 ((Wqkv && ldw >= ntot) || (Wt && Cw && ldwt >= d_model && ldwt % 16 == 0)),
             "qg_attention_forward: bad arguments");
  const bool self = (Xq == Xkv) && sq == skv && ldxq == ldxkv;
  const int64_t tq = (int64_t)batch * sq, tkv = (int64_t)batch * skv;
  QG_REQUIRE(tq < (1 << 30) && tkv < (1 << 30) && (int64_t)batch * heads <= 65535, "qg_attention_forward: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  // scratch: projections [tq, nq] + [tkv, nkv] (one [tq, ntot] matrix for self-attention), scores [batch*heads*sq, skv]
  const size_t proj_elems = self ? (size_t)tq * ntot : (size_t)tq * nq + (size_t)tkv * nkv;
  const size_t score_elems = (size_t)batch * heads * sq * skv;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if ((rc = grow(&d->attn, &d->attn_bytes, sizeof(float) * (proj_elems + score_elems) + 512))) return rc;
  }
  float *proj = (float *)d->attn;
  float *scores = proj + round_up((int64_t)proj_elems, 64);
  // 1. projections through the quantized linear path (attention.cuh:54-56 re-pointed).  Row scales
  //    depend only on X and column scales only on their own column, so one product against the
  //    concatenated [W_q | W_k | W_v] of all heads is bit-identical to the separate ones -- and a column block of the
  //    prepared codes (rows [c0, c1) of Wt, Cw + c0) to the prepared codes of that block alone.
  auto project = [&](const float *X, int64_t ldx, int64_t rows, int c0, int n, float *P) -> int {
    if (Wt) return qg_linear_forward(X, ldx, QG_F32, Wt + (int64_t)c0 * ldwt, ldwt, Cw + c0, nullptr, P, n, QG_F32, (int)rows, n,
                                     d_model, range, mode, nullptr, 0, stream);
    return qg_quantized_mm(X, ldx, Wqkv + c0, ldw, QG_F32, P, n, QG_F32, (int)rows, n, d_model, range, mode, nullptr, nullptr, 0,
                           stream);
  };
  const float *Q, *Kp, *Vp;
  int64_t ldq_, ldkv;
  if (self) {
    if ((rc = project(Xq, ldxq, tq, 0, ntot, proj))) return rc;
    Q = proj; Kp = proj + nq; Vp = proj + nq + heads * d_k;
    ldq_ = ldkv = ntot;
  } else {  // the 3-argument form transformer.cu:37,132 expects: queries from Xq, keys / values from Xkv
    float *pkv = proj + (size_t)tq * nq;
    if ((rc = project(Xq, ldxq, tq, 0, nq, proj))) return rc;
    if ((rc = project(Xkv, ldxkv, tkv, nq, nkv, pkv))) return rc;
    Q = proj; Kp = pkv; Vp = pkv + heads * d_k;
    ldq_ = nq; ldkv = nkv;
  }
  const float scale = (float)(1.0 / std::sqrt((double)d_k));
  // 2-4 in ONE kernel when a (sequence, head) fits a CTA's shared memory (attention.cu): same arithmetic, same bits
  rc = attention_core(Q, ldq_, Kp, Vp, ldkv, out, ldo, batch, heads, sq, skv, d_k, d_v, scale, st);
  if (rc != QG_ENOTSUP) return cuda_status((cudaError_t)rc, "attention core");
  // 2. scores = Q K^T per (sequence, head): fp32, k-ascending FMA chain like op_mm (attention.cuh:58-60)
  MmBatch bt;
  bt.n_outer = batch; bt.n_inner = heads;
  bt.a_outer = (int64_t)sq * ldq_;  bt.a_inner = d_k;
  bt.b_outer = (int64_t)skv * ldkv; bt.b_inner = d_k;
  bt.c_outer = (int64_t)heads * sq * skv; bt.c_inner = (int64_t)sq * skv;
  rc = mm_f32(Q, ldq_, 1, Kp, 1, ldkv, sq, skv, d_k, scores, skv, st, &bt);  // B = K^T: element (k, n) at K[n, k]
  if (rc) return cuda_status((cudaError_t)rc, "scores");
  // 3. softmax(scores / sqrt(d_k)) in place (attention.cuh:62-68)
  rc = softmax_rows(scores, skv, batch * heads * sq, skv, scale, scores, skv, st);
  if (rc) return cuda_status((cudaError_t)rc, "softmax");
  // 4. out[:, h*d_v : (h+1)*d_v] = P V per (sequence, head) (attention.cuh:69; the head concat of transformer.cu:43-50)
  bt.a_outer = (int64_t)heads * sq * skv; bt.a_inner = (int64_t)sq * skv;
  bt.b_outer = (int64_t)skv * ldkv;       bt.b_inner = d_v;
  bt.c_outer = (int64_t)sq * ldo;         bt.c_inner = d_v;
  rc = mm_f32(scores, skv, 1, Vp, ldkv, 1, sq, d_v, skv, out, ldo, st, &bt);
  return cuda_status((cudaError_t)rc, "P*V");
}

int qg_attention_forward(const float *Xq, int64_t ldxq, const float *Xkv, int64_t ldxkv, int batch, int sq, int skv,
                         int d_model, const float *Wqkv, int64_t ldw, int heads, int d_k, int d_v, float range, int mode,
                         float *out, int64_t ldo, qg_stream_t stream) {
  return attention_forward_impl(Xq, ldxq, Xkv, ldxkv, batch, sq, skv, d_model, Wqkv, ldw, nullptr, 0, nullptr, heads, d_k, d_v,
                                range, mode, out, ldo, stream);
}

int qg_attention_forward_prepared(const float *Xq, int64_t ldxq, const float *Xkv, int64_t ldxkv, int batch, int sq, int skv,
                                  int d_model, const int8_t *Wt, int64_t ldwt, const float *Cw, int heads, int d_k, int d_v,
                                  float range, int mode, float *out, int64_t ldo, qg_stream_t stream) {
  return attention_forward_impl(Xq, ldxq, Xkv, ldxkv, batch, sq, skv, d_model, nullptr, 0, Wt, ldwt, Cw, heads, d_k, d_v, range,
                                mode, out, ldo, stream);
}

/* ---- quantization carried across layers ---- */
namespace qg {
__global__ void fill_f32_kernel(float *p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  griddep_wait();
  if (i < n) p[i] = v;
}
}  // namespace qg

int qg_quantize_rows_given_max(const void *Y, int dtype, int M, int K, int64_t ldy, float range, int mode,
                               const float *rowmax, int8_t *Xq, int64_t ldq, float *Cx, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Y && rowmax && Xq && Cx && M > 0 && K > 0 && ldy >= K && ldq >= K && valid_io(dtype),
             "qg_quantize_rows_given_max: bad arguments");
  RowMaxIo io;
  io.m_in = rowmax;
  return quant_rows(Y, dtype, M, K, ldy, range, mode, nullptr, Xq, ldq, Cx, (cudaStream_t)stream, io);
}

int qg_linear_forward_q(const int8_t *Xq, int64_t ldxq, const float *Cx, const int8_t *Wt, int64_t ldwt, const float *Cw,
                        const float *bias, int act, void *Y, int64_t ldy, int out_dtype, int M, int N, int K, float range,
                        float *y_rowmax, void *workspace, size_t workspace_bytes, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && Cx && Wt && Cw && Y && M > 0 && N > 0 && K > 0 && ldxq >= K && ldwt >= K && ldy >= N && valid_io(out_dtype) &&
                 (act == QG_ACT_NONE || act == QG_ACT_RELU),
             "qg_linear_forward_q: bad arguments (M=%d N=%d K=%d)", M, N, K);
  cudaStream_t st = (cudaStream_t)stream;
  void *sk = nullptr;
  size_t sk_bytes = 0;
  if (workspace != nullptr) {  // only the split-K slices of a tile-starved shape live there
    QG_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "qg_linear_forward_q: workspace must be 256-byte aligned");
    sk = workspace;
    sk_bytes = workspace_bytes;
  }
  if (y_rowmax != nullptr)
    QG_CUDA_OK(launch_kernel(fill_f32_kernel, dim3((unsigned)ceil_div(M, 256)), dim3(256), st, y_rowmax, M, -INFINITY));
  bool done = false;
  rc = gemm_dispatch(d, Xq, ldxq, Wt, ldwt, 1, M, N, K, Y, ldy, out_dtype, Cx, Cw, bias, 1 / (range * range), st, nullptr, nullptr,
                     act, sk, sk_bytes, workspace == nullptr, y_rowmax, &done);
  if (rc) return rc;
  if (y_rowmax != nullptr && !done) {  // CUDA-core or split-K form: one reduction pass over Y instead
    RowMaxIo io;
    io.m_out = y_rowmax;
    rc = quant_rows(Y, out_dtype, M, N, ldy, range, QG_MODE_REF_EXACT, nullptr, nullptr, 0, nullptr, st, io);
    if (rc) return cuda_status((cudaError_t)rc, "row maxima");
  }
  return QG_OK;
}

struct FfnWs {
  int8_t *Xq1, *Xq2;
  float *Cx1, *Cx2, *rowmax;
  int64_t ldq1, ldq2;
  void *sk;
  size_t sk_bytes, bytes;
};
static FfnWs carve_ffn(void *base, int M, int d_in, int d_ff, int d_out) {
  FfnWs w;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = (size_t)round_up((int64_t)(off + n), 256); return o; };
  w.ldq1 = round_up(d_in, 16);
  w.ldq2 = round_up(d_ff, 16);
  const size_t o1 = take((size_t)M * w.ldq1), o2 = take((size_t)M * w.ldq2), c1 = take(4 * (size_t)M), c2 = take(4 * (size_t)M),
               rm = take(4 * (size_t)M);
  w.sk_bytes = std::max(splitk_need(M, d_ff, d_in), splitk_need(M, d_out, d_ff));
  const size_t osk = take(w.sk_bytes);
  char *b = reinterpret_cast<char *>(base);
  w.Xq1 = reinterpret_cast<int8_t *>(b + o1); w.Xq2 = reinterpret_cast<int8_t *>(b + o2);
  w.Cx1 = reinterpret_cast<float *>(b + c1); w.Cx2 = reinterpret_cast<float *>(b + c2); w.rowmax = reinterpret_cast<float *>(b + rm);
  w.sk = w.sk_bytes ? b + osk : nullptr;
  w.bytes = off;
  return w;
}
size_t qg_ffn_workspace_bytes(int M, int d_in, int d_ff, int d_out) {
  if (M <= 0 || d_in <= 0 || d_ff <= 0 || d_out <= 0) return 0;
  return carve_ffn(nullptr, M, d_in, d_ff, d_out).bytes;
}

}  // extern "C"

// scatter != NULL: the second GEMM is the row-parallel half of a Megatron pair -- its output (this rank's partial
// product) leaves through the scattering epilogue, Y being destination 0
static int ffn_chain(const void *X, int64_t ldx, int in_dtype, const int8_t *Xq_in, int64_t ldxq_in, const float *Cx_in,
                     const int8_t *W1t, int64_t ldw1t, const float *Cw1, const float *b1, const int8_t *W2t, int64_t ldw2t,
                     const float *Cw2, const float *b2, void *H, int64_t ldh, int h_dtype, void *Y, int64_t ldy, int out_dtype,
                     int M, int d_in, int d_ff, int d_out, float range, int mode, void *workspace, size_t workspace_bytes,
                     qg_stream_t stream, const MultiOut *scatter) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  const bool preq = Xq_in != nullptr;
  QG_REQUIRE((preq ? (Cx_in != nullptr && ldxq_in >= d_in) : (X != nullptr && ldx >= d_in && valid_io(in_dtype))) && W1t && Cw1 &&
                 W2t && Cw2 && H && Y && M > 0 && d_in > 0 && d_ff > 0 && d_out > 0 && ldw1t >= d_in && ldw2t >= d_ff &&
                 ldh >= d_ff && (scatter != nullptr || ldy >= d_out) && valid_io(h_dtype) && valid_io(out_dtype),
             "qg_ffn_forward: bad arguments (M=%d d_in=%d d_ff=%d d_out=%d)", M, d_in, d_ff, d_out);
  const size_t need = carve_ffn(nullptr, M, d_in, d_ff, d_out).bytes;
  if (workspace == nullptr) {
    std::lock_guard<std::mutex> lk(g_mu);
    if ((rc = grow(&d->arena, &d->arena_bytes, need))) return rc;
    workspace = d->arena;
  } else {
    QG_REQUIRE(workspace_bytes >= need && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
               "qg_ffn_forward: workspace too small or misaligned (%zu bytes needed)", need);
  }
  FfnWs w = carve_ffn(workspace, M, d_in, d_ff, d_out);
  cudaStream_t st = (cudaStream_t)stream;
  const float c = 1 / (range * range);
  const int8_t *xq1 = Xq_in;
  int64_t ldq1 = ldxq_in;
  const float *cx1 = Cx_in;
  if (!preq) {
    // ll1's row quantizer; it also arms the row-maximum buffer the first GEMM's epilogue fills
    RowMaxIo io;
    io.init_out = w.rowmax;
    rc = quant_rows(X, in_dtype, M, d_in, ldx, range, mode, nullptr, w.Xq1, w.ldq1, w.Cx1, st, io);
    if (rc) return cuda_status((cudaError_t)rc, "row quantizer");
    xq1 = w.Xq1; ldq1 = w.ldq1; cx1 = w.Cx1;
  } else {
    QG_CUDA_OK(launch_kernel(fill_f32_kernel, dim3((unsigned)ceil_div(M, 256)), dim3(256), st, w.rowmax, M, -INFINITY));
  }
  // ll1.forward + op_relu (transformer.cu:63-67): H = relu(x @ W1 + b1), row maxima of H from the epilogue
  bool done = false;
  rc = gemm_dispatch(d, xq1, ldq1, W1t, ldw1t, 1, M, d_ff, d_in, H, ldh, h_dtype, cx1, Cw1, b1, c, st, nullptr, nullptr, QG_ACT_RELU,
                     w.sk, w.sk_bytes, false, w.rowmax, &done);
  if (rc) return rc;
  // ll2's quantizer: scale from (H[i,0], rowmax[i]) -- no reduction pass over H -- then the codes
  RowMaxIo io2;
  if (done) io2.m_in = w.rowmax;
  rc = quant_rows(H, h_dtype, M, d_ff, ldh, range, mode, nullptr, w.Xq2, w.ldq2, w.Cx2, st, io2);
  if (rc) return cuda_status((cudaError_t)rc, "row quantizer (hidden)");
  // ll2.forward (transformer.cu:69-71)
  return gemm_dispatch(d, w.Xq2, w.ldq2, W2t, ldw2t, 1, M, d_out, d_ff, Y, ldy, out_dtype, w.Cx2, Cw2, b2, c, st, nullptr, scatter,
                       QG_ACT_NONE, w.sk, w.sk_bytes, false);
}

extern "C" {

int qg_ffn_forward(const void *X, int64_t ldx, int in_dtype, const int8_t *Xq_in, int64_t ldxq_in, const float *Cx_in,
                   const int8_t *W1t, int64_t ldw1t, const float *Cw1, const float *b1, const int8_t *W2t, int64_t ldw2t,
                   const float *Cw2, const float *b2, void *H, int64_t ldh, int h_dtype, void *Y, int64_t ldy, int out_dtype,
                   int M, int d_in, int d_ff, int d_out, float range, int mode, void *workspace, size_t workspace_bytes,
                   qg_stream_t stream) {
  return ffn_chain(X, ldx, in_dtype, Xq_in, ldxq_in, Cx_in, W1t, ldw1t, Cw1, b1, W2t, ldw2t, Cw2, b2, H, ldh, h_dtype, Y, ldy,
                   out_dtype, M, d_in, d_ff, d_out, range, mode, workspace, workspace_bytes, stream, nullptr);
}

/* ---- Megatron pairing: column-parallel fc1 -> row-parallel fc2 ---- */
static int scatter_args(const char *who, void *const *part_dst, int n_dst, int block_cols, int64_t ld_part, int n, MultiOut *mo) {
  QG_REQUIRE(part_dst && n_dst >= 1 && n_dst <= kMaxExtraOut + 1 && block_cols > 0 && ld_part >= block_cols &&
                 (int64_t)block_cols * n_dst >= n,
             "%s: bad scatter arguments (n_dst=%d block_cols=%d ld_part=%lld n=%d)", who, n_dst, block_cols, (long long)ld_part, n);
  for (int i = 0; i < n_dst; i++) QG_REQUIRE(part_dst[i] != nullptr, "%s: part_dst[%d] is NULL", who, i);
  *mo = MultiOut();
  mo->n = n_dst - 1;
  for (int i = 1; i < n_dst; i++) mo->dst[i - 1] = part_dst[i];
  mo->scatter_cols = block_cols;
  return QG_OK;
}

int qg_gemm_s8_dequant_scatter(const int8_t *Xq, int64_t ldxq, const int8_t *Wt, int64_t ldwt, const float *Cx, const float *Cw,
                               int M, int N, int K, float range, void *const *part_dst, int n_dst, int block_cols,
                               int64_t ld_part, int part_dtype, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(Xq && Wt && Cx && Cw && M > 0 && N > 0 && K > 0 && ldxq >= K && ldwt >= K && valid_io(part_dtype),
             "qg_gemm_s8_dequant_scatter: bad arguments");
  MultiOut mo;
  if ((rc = scatter_args("qg_gemm_s8_dequant_scatter", part_dst, n_dst, block_cols, ld_part, N, &mo))) return rc;
  return gemm_dispatch(d, Xq, ldxq, Wt, ldwt, 1, M, N, K, part_dst[0], ld_part, part_dtype, Cx, Cw, nullptr, 1 / (range * range),
                       (cudaStream_t)stream, nullptr, &mo);
}

int qg_ffn_forward_rowpar(const void *X, int64_t ldx, int in_dtype, const int8_t *Xq_in, int64_t ldxq_in, const float *Cx_in,
                          const int8_t *W1t, int64_t ldw1t, const float *Cw1, const float *b1, const int8_t *W2t, int64_t ldw2t,
                          const float *Cw2, void *H, int64_t ldh, int h_dtype, void *const *part_dst, int n_dst, int block_cols,
                          int64_t ld_part, int part_dtype, int M, int d_in, int d_ff_local, int d_out, float range, int mode,
                          void *workspace, size_t workspace_bytes, qg_stream_t stream) {
  MultiOut mo;
  int rc = scatter_args("qg_ffn_forward_rowpar", part_dst, n_dst, block_cols, ld_part, d_out, &mo);
  if (rc) return rc;
  // ld_part stands in for ldy; the chain's own check (ldy >= d_out) does not apply to a scattered output
  QG_REQUIRE(valid_io(part_dtype), "qg_ffn_forward_rowpar: bad partial dtype");
  return ffn_chain(X, ldx, in_dtype, Xq_in, ldxq_in, Cx_in, W1t, ldw1t, Cw1, b1, W2t, ldw2t, Cw2, nullptr, H, ldh, h_dtype,
                   part_dst[0], ld_part, part_dtype, M, d_in, d_ff_local, d_out, range, mode, workspace, workspace_bytes, stream, &mo);
}

/* [rows, width_bytes] block copy between device buffers (possibly on different GPUs of a peer-mapped allocation) by the COPY
 * ENGINES: the gather half of the Megatron exchange when it has to run under GEMMs that hold every SM. */
int qg_copy_2d_async(void *dst, int64_t dst_pitch_bytes, const void *src, int64_t src_pitch_bytes, int64_t width_bytes, int rows,
                     qg_stream_t stream) {
  QG_REQUIRE(dst && src && width_bytes > 0 && rows > 0 && dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes,
             "qg_copy_2d_async: bad arguments");
  QG_CUDA_OK(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes, (size_t)rows,
                               cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return QG_OK;
}

int qg_reduce_partials(const void *slots, int64_t slot_stride, int n_slots, int part_dtype, int64_t ld_part, const float *bias,
                       void *out, void *const *peers, int n_peers, void *out_mc, int64_t ldo, int out_dtype, int m, int n,
                       int max_ctas, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(slots && out && n_slots >= 1 && m > 0 && n > 0 && ld_part >= n && ldo >= n && valid_io(part_dtype) &&
                 valid_io(out_dtype) && n_peers >= 0 && n_peers <= kMaxExtraOut && (n_peers == 0 || peers) &&
                 (n_slots == 1 || slot_stride >= (int64_t)(m - 1) * ld_part + n),
             "qg_reduce_partials: bad arguments");
  return cuda_status((cudaError_t)reduce_partials(slots, slot_stride, n_slots, part_dtype, ld_part, bias, out, peers, n_peers, ldo,
                                                  out_dtype, m, n, (cudaStream_t)stream, out_mc, max_ctas),
                     "qg_reduce_partials");
}

int qg_add_layernorm_quant_f32(const float *A, int64_t lda, const float *R, int64_t ldr, int m, int n, float *B, int64_t ldb,
                               float range, int mode, int8_t *Xq, int64_t ldq, float *Cx, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && B && Xq && Cx && m > 0 && n > 0 && lda >= n && ldb >= n && ldq >= n && (R == nullptr || ldr >= n),
             "qg_add_layernorm_quant_f32: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  rc = add_layernorm_rows(A, lda, R, ldr, m, n, B, ldb, st, Xq, ldq, Cx, range, mode);
  if (rc == QG_ENOTSUP) {  // very wide rows: the two passes separately
    rc = add_layernorm_rows(A, lda, R, ldr, m, n, B, ldb, st);
    if (rc) return cuda_status((cudaError_t)rc, "add + layernorm");
    rc = quant_rows(B, QG_F32, m, n, ldb, range, mode, nullptr, Xq, ldq, Cx, st);
    return rc ? cuda_status((cudaError_t)rc, "row quantizer") : QG_OK;
  }
  return cuda_status((cudaError_t)rc, "add + layernorm + quantize");
}

/* ---- the elementwise tail of the pipeline, op by op (src/ops/op_elemwise.cuh) ---- */
static int bcast_mode(int b_rows, int b_cols, int m, int n) {
  // the reference's rule (op_elemwise.cuh:410-421): [1,n] repeats down the rows, [m,1] across the columns
  if (b_rows == 1 && b_cols == n && m != 1) return 1;
  if (b_cols == 1 && b_rows == m && n != 1) return 2;
  if (b_rows == m && b_cols == n) return 0;
  return -1;
}
static int binary_op(int op, const char *name, const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols,
                     float *O, int64_t ldo, int m, int n, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  const int bm = bcast_mode(b_rows, b_cols, m, n);
  QG_REQUIRE(A && B && O && m > 0 && n > 0 && lda >= n && ldo >= n && bm >= 0 && (bm != 0 || ldb >= n) && (bm != 2 || ldb >= 1),
             "%s: bad arguments (b is %d x %d against %d x %d)", name, b_rows, b_cols, m, n);
  return cuda_status((cudaError_t)elemwise(op, A, lda, B, ldb, bm, 0.0f, O, ldo, m, n, (cudaStream_t)stream), name);
}
int qg_add_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O, int64_t ldo, int m,
               int n, qg_stream_t stream) {
  return binary_op(0, "qg_add_f32", A, lda, B, ldb, b_rows, b_cols, O, ldo, m, n, stream);
}
int qg_subtract_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O, int64_t ldo,
                    int m, int n, qg_stream_t stream) {
  return binary_op(1, "qg_subtract_f32", A, lda, B, ldb, b_rows, b_cols, O, ldo, m, n, stream);
}
int qg_multiply_f32(const float *A, int64_t lda, const float *B, int64_t ldb, int b_rows, int b_cols, float *O, int64_t ldo,
                    int m, int n, qg_stream_t stream) {
  return binary_op(2, "qg_multiply_f32", A, lda, B, ldb, b_rows, b_cols, O, ldo, m, n, stream);
}
int qg_multiply_const_f32(const float *A, int64_t lda, float c, float *O, int64_t ldo, int m, int n, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && O && m > 0 && n > 0 && lda >= n && ldo >= n, "qg_multiply_const_f32: bad arguments");
  return cuda_status((cudaError_t)elemwise(4, A, lda, nullptr, 0, 3, c, O, ldo, m, n, (cudaStream_t)stream), "qg_multiply_const_f32");
}
int qg_relu_f32(const float *A, int64_t lda, float *O, int64_t ldo, int m, int n, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(A && O && m > 0 && n > 0 && lda >= n && ldo >= n, "qg_relu_f32: bad arguments");
  return cuda_status((cudaError_t)elemwise(5, A, lda, nullptr, 0, 3, 0.0f, O, ldo, m, n, (cudaStream_t)stream), "qg_relu_f32");
}
int qg_dequantize_outer_f32(const int32_t *acc, int64_t ldacc, const float *outer, int64_t ldouter, float *O, int64_t ldo, int m,
                            int n, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  QG_REQUIRE(acc && outer && O && m > 0 && n > 0 && ldacc >= n && ldouter >= n && ldo >= n, "qg_dequantize_outer_f32: bad arguments");
  return cuda_status((cudaError_t)elemwise(3, acc, ldacc, outer, ldouter, 0, 0.0f, O, ldo, m, n, (cudaStream_t)stream),
                     "qg_dequantize_outer_f32");
}

/* bring-up hook: device buffer (8 x int64 per CTA) that the tcgen05 GEMM fills with pipeline wait
 * counters; NULL switches the instrumentation off. */
QG_API int qg_debug_gemm_stats(long long *dev_ptr) {
  gemm_i8_tc_set_stats(dev_ptr);
  return QG_OK;
}

/* test hook (not part of the reference-facing surface): tcgen05 GEMM with B given as [N,K]
 * (K-major operand), used to validate the MN-major descriptor path against the classic one. */
QG_API int qg_test_gemm_s8_bt(int cg, const int8_t *A, int64_t lda, const int8_t *Bt, int64_t ldbt, int M, int N, int K,
                              int32_t *C, int64_t ldc, qg_stream_t stream) {
  DeviceState *d;
  int rc = device_state(&d);
  if (rc) return rc;
  return gemm_i8_tc(cg, A, lda, Bt, ldbt, 1, M, N, K, C, ldc, QG_S32, nullptr, nullptr, nullptr, 0.0f, nullptr, nullptr,
                    d->sm_count, (cudaStream_t)stream);
}

}  // extern "C"
