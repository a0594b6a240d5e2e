// common.cuh -- shared helpers for the sm_100a kernels: error plumbing, dtype access, and thin
// inline-PTX wrappers for mbarrier / TMA / tcgen05 (Blackwell 5th-gen tensor cores + TMEM).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/qgemm.h"

namespace qg {

// ------------------------------------------------------------------------------------------
// host-side error plumbing (C-ABI returns int; message kept per thread)
// ------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_status(cudaError_t e, const char *what);

#define QG_CUDA_OK(expr)                                   \
  do {                                                     \
    cudaError_t _e = (expr);                               \
    if (_e != cudaSuccess) return qg::cuda_status(_e, #expr); \
  } while (0)

#define QG_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      qg::set_error(__VA_ARGS__);        \
      return QG_EINVAL;                  \
    }                                    \
  } while (0)

void count_launch(int n = 1);
bool pdl_enabled();

// cudaFuncSetAttribute acts on the CURRENT device only, so the opt-in for more than 48 KB of dynamic shared
// memory is remembered per (kernel instantiation, device): `done` is a function-static array at the call site.
constexpr int kMaxDevices = 16;
template <typename Kern>
inline cudaError_t smem_optin(Kern kern, int bytes, bool (&done)[kMaxDevices]) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
  if (done[dev]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[dev] = true;
  return e;
}

// Launch through cudaLaunchKernelEx so that consecutive kernels of one call can overlap their
// launch latency and prologue (programmatic stream serialization); every kernel of this library
// executes griddep_wait() before touching global memory.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  count_launch();
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// optional outlier side product of the GEMM epilogue (see outlier.cu / gemm_i8_tc.cu)
struct SideArgs {
  const void *Xo;  // [M, ldxo] 16-bit, columns = rank of the outlier feature, zero padded to no_pad
  int64_t ldxo;
  const void *Wo;  // [no_pad, ldwo] 16-bit rows W[idx[o], :]
  int64_t ldwo;
  int no_pad;      // a multiple of 8, at most 64
  int side_bf16;   // 0: fp16 operands, 1: bf16
};

// Row-maximum hand-over between a producing GEMM epilogue and the next layer's row quantizer (SURVEY.md section 8f,
// rank 3).  m = max_{j >= 1, x_j not NaN} |x_j| (-inf when there is none), i.e. the reduction WITHOUT the signed first
// element, which the quantizer folds itself (op_reduction.cuh:80).
struct RowMaxIo {
  const float *m_in = nullptr;  // take m from here instead of reducing the row
  float *m_out = nullptr;       // store m (reduction-only use)
  float *init_out = nullptr;    // write -inf: arms a buffer the FOLLOWING GEMM's epilogue will atomicMax into
};

// batch geometry of the fp32 product: blockIdx.z = outer * n_inner + inner selects one independent
// product; element offsets per operand (attention: outer = sequence, inner = head)
struct MmBatch {
  int n_outer, n_inner;
  int64_t a_outer, a_inner, b_outer, b_inner, c_outer, c_inner;
};

// additional destinations of the GEMM epilogue: the same [M,N] block is also written (same leading
// dimension and dtype) to these buffers -- peers' output matrices mapped over NVLink in the
// column-parallel linear, so the all-gather rides inside the GEMM
constexpr int kMaxExtraOut = 7;
struct MultiOut {
  void *dst[kMaxExtraOut];
  int n;
  // != NULL: the NVSwitch multicast address of the same block (a multimem address covering EVERY GPU's matrix, the
  // caller's included).  The epilogue then issues one multimem.st per 16 bytes instead of one store per GPU: the
  // switch replicates, so the sender's NVLink egress is 1x the block instead of (P-1)x.  dst[] / n are ignored.
  void *mc;
  // > 0: scatter instead of replicate -- column block b (scatter_cols columns) goes to destination b only
  // (0 = the call's own output pointer, b >= 1 = dst[b-1]); see GemmParams::scatter_cols
  int scatter_cols;
};

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }
static inline size_t dtype_size(int dt) { return dt == QG_F32 ? 4 : 2; }

// ------------------------------------------------------------------------------------------
// device: element access for the three I/O dtypes (all math is done in fp32)
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// The reference's quantizing cast (src/ops/op_elemwise.cuh:106-114 -> mul.f32; cvt.rzi.s32.f32;
// st.u8): truncate toward zero, clamp to int32, NaN -> 0, keep the low byte.
__device__ __forceinline__ uint32_t quant_code_u8(float x, float s) {
  return (uint32_t)__float2int_rz(__fmul_rn(x, s)) & 0xffu;
}

// Reference dequantize chain (op_mm.cuh:96-99, op_elemwise.cuh:93-103,118-129): four separately
// rounded fp32 operations; `outer + 0.0f` reproduces the K=1 tiled product's -0 -> +0.
__device__ __forceinline__ float dequant_ref(int acc, float cx, float cw, float c) {
  float outer = __fadd_rn(__fmul_rn(cx, cw), 0.0f);
  return __fmul_rn(__fmul_rn(__int2float_rn(acc), outer), c);
}

// ------------------------------------------------------------------------------------------
// device: PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta_rank) : "memory");
}
// Same arrive without the release fence (MEMBAR.ALL + ERRBAR in SASS, ~1k cycles when bulk stores
// are in flight).  Only for hand-offs whose payload is TMEM: the reads being handed back were
// completed by tcgen05.wait::ld and ordered by tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta_rank) : "memory");
}
// explicit shared-window accesses (pointers rebuilt from an aligned uintptr_t are generic for nvcc)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done;
}
// Bounded wait: a barrier that never flips is a protocol bug; report and trap instead of hanging
// the GPU (about 2 s at 2 GHz).
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity, int tag) {
  printf("[qgemm] mbarrier timeout: block %d thread %d tag %d bar 0x%x parity %u\n", (int)blockIdx.x,
         (int)threadIdx.x, tag, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 4000000000LL) mbar_timeout(bar, parity, tag);
  }
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const void *desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(desc) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void *desc, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 2-SM form: data lands in this CTA's smem, completion is signalled on the barrier address given
// (a shared::cluster address, normally the pair leader's barrier).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void *desc, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(desc), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// 2-SM load multicast to the CTAs of `cta_mask`: the box lands at the same shared-memory offset in each of them and
// completes on the barrier at offset `bar` of each destination's pair (peer bit cleared by the caller)
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const void *desc, uint32_t bar, int c0, int c1, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(desc), "r"(bar), "h"(cta_mask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(desc), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// the same store carrying an L2 eviction policy (createpolicy.fractional.L2::*): output tiles are written once
// and not read again by the kernel that produces them
__device__ __forceinline__ void tma_store_2d_hint(const void *desc, uint32_t src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
               ::"l"(desc), "r"(src), "r"(c0), "r"(c1), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 1: evict_first, 2: evict_last, else evict_normal
  uint64_t pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// 16 bytes to every GPU bound to the multicast object behind `mc_addr` (NVLS; the switch replicates the write)
__device__ __forceinline__ void multimem_st128(void *mc_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};"
               ::"l"(mc_addr), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
template <int N> __device__ __forceinline__ void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- cluster ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- tcgen05 / TMEM ----
template <int CG> __device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int CG> __device__ __forceinline__ void tmem_relinquish() {
  if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG> __device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; int8 x int8 -> int32
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of all prior tcgen05 ops of this thread -> one arrive on an mbarrier
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 2-SM form: arrive on the barrier at this smem offset in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(cta_mask) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (thread t <- lane t)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) ----
// wait: blocks until every prerequisite grid has completed and its writes are visible (no-op when the
// kernel was launched without a programmatic dependency).  launch_dependents: lets the next kernel
// in the stream start its prologue while this grid is still running.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Early trigger, issued right after griddep_wait(): the next kernel of the stream is launched as soon
// as every CTA of this grid is running and then blocks in its own griddep_wait().  EXPERIMENT ONLY
// (build.py --tag early -DQG_EARLY_TRIGGER): measured on B200 it is a loss -- op_quantized_mm
// 4096^3 102.5 us against 95.7 us, 8192^3 531 against 513 us (profiles/r1_early_trigger_ab.json);
// parked dependents cost the running grid more than the hidden launch latency returns.  Two targeted
// triggers are kept because they measured as gains: the row quantizer's, on each CTA's last row
// block, and the column quantizer's second pass (quant_cols_kernel), whose dependent is the GEMM.
__device__ __forceinline__ void griddep_trigger_early() {
#ifdef QG_EARLY_TRIGGER
  griddep_launch_dependents();
#endif
}

// named barrier among a subset of warps
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- UMMA descriptors (bit layout: PTX ISA "tcgen05 shared memory / instruction descriptor") ----
// Shared-memory matrix descriptor, 128B swizzle.  Fields in 16-byte units:
//   [0,14) start address, [16,30) leading byte offset, [32,46) stride byte offset,
//   [46,48) version = 1 (Blackwell), [61,64) layout = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::i8: D = S32 (c_format 2 @ [4,6)), A/B = signed 8-bit (format 1
// @ [7,10) / [10,13)), a_major @15, b_major @16 (0 = K-major, 1 = MN-major), N>>3 @ [17,23),
// M>>4 @ [24,29); no saturation, dense.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n, int a_mn_major, int b_mn_major) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace qg
