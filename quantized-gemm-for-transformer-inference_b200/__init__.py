"""B200-native quantized-linear hot path: Python binding of the C ABI in include/qgemm.h.

This package is a thin ctypes layer over libqgemm.so (hand-written sm_100a CUDA: row / column
absmax quantizers and a tcgen05 kind::i8 GEMM with a fused dequantize epilogue).  The functions
carry the reference's operator names and argument order (/root/reference/src/ops/*.cuh) and
operate on torch CUDA tensors, which are used only as device-memory handles.

There is no CPU or PyTorch fallback: if libqgemm.so is missing, or no sm_100 GPU is visible,
calls raise.  The CPU oracle under oracle/ is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# QG_LIB selects an A/B build variant of the same library (build.py --tag); still in-tree, still CUDA only
LIB_PATH = os.path.join(_HERE, os.path.basename(os.environ.get("QG_LIB", "libqgemm.so")))

QG_F32, QG_F16, QG_BF16, QG_S32 = 0, 1, 2, 3
MODE_REF_EXACT, MODE_TRUE_ABSMAX = 0, 1
GEMM_AUTO, GEMM_SIMT, GEMM_TC_1SM, GEMM_TC_2SM = 0, 1, 2, 3

# every symbol include/qgemm.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "qg_version", "qg_last_error", "qg_device_info", "qg_set_gemm_variant", "qg_set_gemm_sm_limit", "qg_launch_count",
    "qg_absmax_rows", "qg_absmax_cols", "qg_inv_divide_f32", "qg_quantize_rows", "qg_quantize_cols",
    "qg_absmax_quant_rows", "qg_absmax_quant_cols", "qg_absmax_quant_rows_cols", "qg_gemm_s8s8s32", "qg_dequantize_s32",
    "qg_gemm_s8_dequant", "qg_workspace_bytes", "qg_quantized_mm", "qg_prepare_weights", "qg_gemm_s8t_dequant",
    "qg_linear_forward",
    "qg_quantized_mm_host", "qg_outlier_mask_f32", "qg_outlier_cols", "qg_outlier_workspace_bytes",
    "qg_linear_forward_outlier", "qg_linear_forward_multi", "qg_gemm_s8_dequant_ex", "qg_gemm_s8_dequant_mc", "qg_mm_f32",
    "qg_softmax_rows_f32", "qg_attention_forward", "qg_attention_forward_prepared", "qg_linear_forward_act", "qg_add_layernorm_f32",
    "qg_linear_forward_q", "qg_quantize_rows_given_max", "qg_ffn_workspace_bytes", "qg_ffn_forward",
    "qg_add_layernorm_quant_f32", "qg_gemm_s8_dequant_scatter", "qg_ffn_forward_rowpar", "qg_reduce_partials", "qg_copy_2d_async",
    "qg_add_f32", "qg_subtract_f32", "qg_multiply_f32", "qg_multiply_const_f32", "qg_relu_f32", "qg_dequantize_outer_f32",
]


class QGemmError(RuntimeError):
    pass


_lib = None


def lib() -> C.CDLL:
    """Load libqgemm.so (built in-tree by build.py / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QGemmError(
                f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        L.qg_last_error.restype = C.c_char_p
        L.qg_workspace_bytes.restype = C.c_size_t
        L.qg_outlier_workspace_bytes.restype = C.c_size_t
        L.qg_launch_count.restype = C.c_int64
        L.qg_ffn_workspace_bytes.restype = C.c_size_t
        _lib = L
    return _lib


def _check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().qg_last_error().decode(errors="replace")
        raise QGemmError(f"{what} failed with status {rc}: {msg}")


_DT = {torch.float32: QG_F32, torch.float16: QG_F16, torch.bfloat16: QG_BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise AssertionError(f"unsupported dtype {t.dtype}") from None


def _dev2d(t: torch.Tensor):
    """(pointer, leading dimension) of a 2-D CUDA tensor with unit inner stride."""
    assert t.is_cuda, "tensor must be on the device (the reference asserts on_device)"
    assert t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1), "row-major tensor with stride_w == 1 required"
    ld = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])  # size-1 dims carry arbitrary strides
    return C.c_void_p(t.data_ptr()), C.c_int64(ld)


def _vec(t: torch.Tensor, n: int):
    assert t.is_cuda and t.dtype == torch.float32 and t.numel() == n and t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_info():
    sm, major, minor = C.c_int(), C.c_int(), C.c_int()
    _check(lib().qg_device_info(C.byref(sm), C.byref(major), C.byref(minor)), "qg_device_info")
    return sm.value, major.value, minor.value


def set_gemm_variant(v: int) -> None:
    _check(lib().qg_set_gemm_variant(C.c_int(v)), "qg_set_gemm_variant")


def set_gemm_sm_limit(sms: int) -> None:
    """SMs the persistent tensor-core GEMM may occupy (0 = all); see qg_set_gemm_sm_limit."""
    _check(lib().qg_set_gemm_sm_limit(int(sms)), "qg_set_gemm_sm_limit")


def launch_count(reset: bool = False) -> int:
    return int(lib().qg_launch_count(C.c_int(1 if reset else 0)))


# ------------------------------------------------------------------------------------------
# Reference-named operators (same argument order as /root/reference/src/ops/*.cuh)
# ------------------------------------------------------------------------------------------

def op_absmax(inp: torch.Tensor, out: torch.Tensor, mode: int = MODE_REF_EXACT) -> None:
    """op_absmax(in, out), src/ops/op_reduction.cuh:195-204.  out [h,1] -> per-row reduce,
    out [1,w] -> per-column reduce (direction picked from the output shape, :143)."""
    h, w = inp.shape
    assert (out.shape[0] == 1 and out.shape[1] == w) or (out.shape[1] == 1 and out.shape[0] == h)
    assert out.dtype == torch.float32 and out.is_contiguous()
    p, ld = _dev2d(inp)
    if h > out.shape[0]:  # out [1,w]
        _check(lib().qg_absmax_cols(p, _dt(inp), h, w, ld, mode, _vec(out, w), _stream()), "qg_absmax_cols")
    else:
        _check(lib().qg_absmax_rows(p, _dt(inp), h, w, ld, mode, _vec(out, h), _stream()), "qg_absmax_rows")


def op_inv_divide(a: torch.Tensor, b: float, out: torch.Tensor) -> None:
    """op_inv_divide(a, b, out) = b / a, src/ops/op_elemwise.cuh:657-667."""
    assert out.shape == a.shape and a.is_contiguous() and out.is_contiguous()
    _check(lib().qg_inv_divide_f32(_vec(a, a.numel()), C.c_int64(a.numel()), C.c_float(b), _vec(out, out.numel()),
                                   _stream()), "qg_inv_divide_f32")


def op_multiply(a: torch.Tensor, b, out: torch.Tensor) -> None:
    """op_multiply<T,int8_t>(a, scale, out): quantizing multiply with broadcast,
    src/ops/op_elemwise.cuh:629-640 (b [h,1] or [1,w])."""
    assert out.dtype == torch.int8, "only the quantizing overload is on the hot path"
    h, w = a.shape
    assert out.shape == a.shape
    pa, lda = _dev2d(a)
    po, ldo = _dev2d(out)
    if b.shape[1] == 1 and b.shape[0] == h and not (b.shape[0] == 1 and w == 1):
        _check(lib().qg_quantize_rows(pa, _dt(a), h, w, lda, _vec(b, h), po, ldo, _stream()), "qg_quantize_rows")
    else:
        assert b.shape[0] == 1 and b.shape[1] == w
        _check(lib().qg_quantize_cols(pa, _dt(a), h, w, lda, _vec(b, w), po, ldo, _stream()), "qg_quantize_cols")


def op_mm(A: torch.Tensor, B: torch.Tensor, Cout: torch.Tensor) -> None:
    """op_mm<T,OutT>(A, B, C), src/ops/op_mm.cuh:49-65: int8 x int8 -> int32, or fp32."""
    assert A.shape[0] == Cout.shape[0] and B.shape[1] == Cout.shape[1] and A.shape[1] == B.shape[0]
    M, K = A.shape
    N = B.shape[1]
    if A.dtype == torch.int8:
        assert B.dtype == torch.int8 and Cout.dtype == torch.int32
        pa, lda = _dev2d(A)
        pb, ldb = _dev2d(B)
        pc, ldc = _dev2d(Cout)
        _check(lib().qg_gemm_s8s8s32(pa, lda, pb, ldb, M, N, K, pc, ldc, _stream()), "qg_gemm_s8s8s32")
    else:
        assert A.dtype == torch.float32 and B.dtype == torch.float32 and Cout.dtype == torch.float32
        assert A.is_cuda and B.is_cuda and Cout.is_cuda and Cout.stride(1) == 1
        _check(lib().qg_mm_f32(C.c_void_p(A.data_ptr()), C.c_int64(A.stride(0)), C.c_int64(A.stride(1)),
                               C.c_void_p(B.data_ptr()), C.c_int64(B.stride(0)), C.c_int64(B.stride(1)),
                               M, N, K, C.c_void_p(Cout.data_ptr()), C.c_int64(Cout.stride(0)), _stream()),
               "qg_mm_f32")


def op_dequantize(acc: torch.Tensor, Cx: torch.Tensor, Cw: torch.Tensor, out: torch.Tensor,
                  range_: float = 127.0, bias: torch.Tensor | None = None) -> None:
    """The reference's op_mm(Cx,Cw,outer); op_dequantize(acc, outer, O); op_multiply(O, 1/range^2, O)
    (src/ops/op_mm.cuh:96-99) as one pass over stored accumulators."""
    M, N = acc.shape
    assert acc.dtype == torch.int32 and out.shape == acc.shape
    pa, lda = _dev2d(acc)
    po, ldo = _dev2d(out)
    pb = None if bias is None else _vec(bias.reshape(-1), N)
    _check(lib().qg_dequantize_s32(pa, lda, _vec(Cx.reshape(-1), M), _vec(Cw.reshape(-1), N), pb, M, N,
                                   C.c_float(range_), po, _dt(out), ldo, _stream()), "qg_dequantize_s32")


def op_outlier_extractor(a: torch.Tensor, b: float, out: torch.Tensor) -> None:
    """op_outlier_extractor(a, b, out), src/ops/op_elemwise.cuh:698-708: out = |a| <= b ? 0 : 1."""
    assert out.shape == a.shape and a.dtype == torch.float32 and out.dtype == torch.float32
    pa, lda = _dev2d(a)
    po, ldo = _dev2d(out)
    _check(lib().qg_outlier_mask_f32(pa, a.shape[0], a.shape[1], lda, C.c_float(b), po, ldo, _stream()),
           "qg_outlier_mask_f32")


def absmax_quant_rows(X: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT,
                      Xq: torch.Tensor | None = None, Cx: torch.Tensor | None = None):
    """Fused a1+a3+a4 for activations: returns (Xq int8 [M,K], Cx f32 [M])."""
    M, K = X.shape
    if Xq is None:
        Xq = torch.empty((M, K), dtype=torch.int8, device=X.device)
    if Cx is None:
        Cx = torch.empty(M, dtype=torch.float32, device=X.device)
    px, ldx = _dev2d(X)
    pq, ldq = _dev2d(Xq)
    _check(lib().qg_absmax_quant_rows(px, _dt(X), M, K, ldx, C.c_float(range_), mode, pq, ldq, _vec(Cx, M),
                                      _stream()), "qg_absmax_quant_rows")
    return Xq, Cx


def absmax_quant_cols(W: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT,
                      Wq: torch.Tensor | None = None, Cw: torch.Tensor | None = None):
    """Fused a2+a3+a4 for weights: returns (Wq int8 [K,N], Cw f32 [N])."""
    K, N = W.shape
    if Wq is None:
        Wq = torch.empty((K, N), dtype=torch.int8, device=W.device)
    if Cw is None:
        Cw = torch.empty(N, dtype=torch.float32, device=W.device)
    pw, ldw = _dev2d(W)
    pq, ldq = _dev2d(Wq)
    _check(lib().qg_absmax_quant_cols(pw, _dt(W), K, N, ldw, C.c_float(range_), mode, pq, ldq, _vec(Cw, N),
                                      None, _stream()), "qg_absmax_quant_cols")
    return Wq, Cw


def absmax_quant_rows_cols(X: torch.Tensor, W: torch.Tensor, Xq: torch.Tensor, Cx: torch.Tensor, Wq: torch.Tensor,
                           Cw: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT) -> None:
    """Both quantizers of one op (qg_absmax_quant_rows_cols): same results as absmax_quant_rows + absmax_quant_cols; large
    problems overlap the column quantizer's second pass with the row quantizer in one launch."""
    M, K = X.shape
    N = W.shape[1]
    assert W.shape[0] == K and X.dtype == W.dtype
    px, ldx = _dev2d(X)
    pw, ldw = _dev2d(W)
    pxq, ldxq = _dev2d(Xq)
    pwq, ldwq = _dev2d(Wq)
    _check(lib().qg_absmax_quant_rows_cols(px, ldx, pw, ldw, _dt(X), M, N, K, C.c_float(range_), mode, pxq, ldxq, _vec(Cx, M),
                                           pwq, ldwq, _vec(Cw, N), _stream()), "qg_absmax_quant_rows_cols")


def gemm_s8_dequant(Xq, Wq, Cx, Cw, out, range_: float = 127.0, bias=None) -> None:
    """int8 GEMM with the dequantize / bias / cast epilogue fused (a5..a8, a10)."""
    M, K = Xq.shape
    N = Wq.shape[1]
    assert Wq.shape[0] == K and out.shape == (M, N)
    pa, lda = _dev2d(Xq)
    pb, ldb = _dev2d(Wq)
    po, ldo = _dev2d(out)
    pbias = None if bias is None else _vec(bias.reshape(-1), N)
    _check(lib().qg_gemm_s8_dequant(pa, lda, pb, ldb, _vec(Cx.reshape(-1), M), _vec(Cw.reshape(-1), N), pbias, M, N, K,
                                    C.c_float(range_), po, _dt(out), ldo, _stream()), "qg_gemm_s8_dequant")


def prepare_weights(W: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT,
                    Wt: torch.Tensor | None = None, Cw: torch.Tensor | None = None):
    """Column-quantize W [K,N] once into the GEMM's preferred operand layout: returns
    (Wt int8 [N, K] view with a 16-byte aligned leading dimension, Cw f32 [N])."""
    K, N = W.shape
    if Wt is None:
        ldk = (K + 15) // 16 * 16
        Wt = torch.zeros((N, ldk), dtype=torch.int8, device=W.device)[:, :K]
    if Cw is None:
        Cw = torch.empty(N, dtype=torch.float32, device=W.device)
    assert Wt.shape == (N, K) and Wt.dtype == torch.int8
    pw, ldw = _dev2d(W)
    pt, ldt = _dev2d(Wt)
    _check(lib().qg_prepare_weights(pw, _dt(W), K, N, ldw, C.c_float(range_), mode, pt, ldt, _vec(Cw, N), _stream()),
           "qg_prepare_weights")
    return Wt, Cw


def gemm_s8t_dequant(Xq, Wt, Cx, Cw, out, range_: float = 127.0, bias=None) -> None:
    """GEMM on prepared (transposed, K-major) weights; out int32 -> raw accumulators."""
    M, K = Xq.shape
    N = Wt.shape[0]
    assert Wt.shape[1] == K and out.shape == (M, N)
    pa, lda = _dev2d(Xq)
    pb, ldb = _dev2d(Wt)
    po, ldo = _dev2d(out)
    od = QG_S32 if out.dtype == torch.int32 else _dt(out)
    pcx = None if Cx is None else _vec(Cx.reshape(-1), M)
    pcw = None if Cw is None else _vec(Cw.reshape(-1), N)
    pbias = None if bias is None else _vec(bias.reshape(-1), N)
    _check(lib().qg_gemm_s8t_dequant(pa, lda, pb, ldb, pcx, pcw, pbias, M, N, K, C.c_float(range_), po, od, ldo,
                                     _stream()), "qg_gemm_s8t_dequant")


def gemm_s8_dequant_ex(Xq, B, b_kmajor: bool, Cx, Cw, out, peer_ptrs=(), range_: float = 127.0, bias=None) -> None:
    """Fused GEMM with B in either layout; `out` ([M,n] view, possibly a column block of a wider matrix)
    is also written at the device addresses `peer_ptrs` (ints: same block inside peers' matrices)."""
    M, K = Xq.shape
    N = B.shape[0] if b_kmajor else B.shape[1]
    assert out.shape == (M, N)
    pa, lda = _dev2d(Xq)
    pb, ldb = _dev2d(B)
    po, ldo = _dev2d(out)
    pbias = None if bias is None else _vec(bias.reshape(-1), N)
    n = len(peer_ptrs)
    arr = (C.c_void_p * max(n, 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
    _check(lib().qg_gemm_s8_dequant_ex(pa, lda, pb, ldb, 1 if b_kmajor else 0, _vec(Cx.reshape(-1), M),
                                       _vec(Cw.reshape(-1), N), pbias, M, N, K, C.c_float(range_), po, arr, n, _dt(out),
                                       ldo, _stream()), "qg_gemm_s8_dequant_ex")


def gemm_s8_dequant_mc(Xq, B, b_kmajor: bool, Cx, Cw, out, mc_ptr: int, range_: float = 127.0, bias=None) -> None:
    """The same GEMM with the exchange done by the NVSwitch: `mc_ptr` is the multicast (multimem) address of the block `out`
    inside a symmetric allocation; one store per 16 bytes reaches every GPU's matrix, this one's included."""
    M, K = Xq.shape
    N = B.shape[0] if b_kmajor else B.shape[1]
    assert out.shape == (M, N) and mc_ptr
    pa, lda = _dev2d(Xq)
    pb, ldb = _dev2d(B)
    po, ldo = _dev2d(out)
    pbias = None if bias is None else _vec(bias.reshape(-1), N)
    _check(lib().qg_gemm_s8_dequant_mc(pa, lda, pb, ldb, 1 if b_kmajor else 0, _vec(Cx.reshape(-1), M),
                                       _vec(Cw.reshape(-1), N), pbias, M, N, K, C.c_float(range_), po, C.c_void_p(int(mc_ptr)),
                                       _dt(out), ldo, _stream()), "qg_gemm_s8_dequant_mc")


def workspace_bytes(M: int, N: int, K: int) -> int:
    return int(lib().qg_workspace_bytes(M, N, K))


def op_quantized_mm(X: torch.Tensor, W: torch.Tensor, O: torch.Tensor, range_: float = 127.0,
                    mode: int = MODE_REF_EXACT, bias: torch.Tensor | None = None,
                    workspace: torch.Tensor | None = None) -> None:
    """op_quantized_mm(X, W, O, range), src/ops/op_mm.cuh:67-101 (+ optional LinearLayer bias)."""
    assert X.shape[0] == O.shape[0] and W.shape[1] == O.shape[1] and X.shape[1] == W.shape[0]
    assert X.dtype == W.dtype
    M, K = X.shape
    N = W.shape[1]
    px, ldx = _dev2d(X)
    pw, ldw = _dev2d(W)
    po, ldo = _dev2d(O)
    pbias = None if bias is None else _vec(bias.reshape(-1), N)
    ws_p, ws_n = (None, 0) if workspace is None else (C.c_void_p(workspace.data_ptr()), workspace.numel())
    _check(lib().qg_quantized_mm(px, ldx, pw, ldw, _dt(X), po, ldo, _dt(O), M, N, K, C.c_float(range_), mode, pbias,
                                 ws_p, C.c_size_t(ws_n), _stream()), "qg_quantized_mm")


def linear_forward(X: torch.Tensor, Wt: torch.Tensor, Cw: torch.Tensor, bias, Y: torch.Tensor, range_: float = 127.0,
                   mode: int = MODE_REF_EXACT, act: int = 0, workspace: torch.Tensor | None = None) -> None:
    """LinearLayer::forward on prepared weights (qg_linear_forward_act): y = act(x @ w + b); the caller's
    workspace (workspace_bytes(M, N, K)) holds every temporary of the call, split-K slices included."""
    M, K = X.shape
    N = Wt.shape[0]
    assert Wt.shape[1] == K and Y.shape == (M, N)
    px, ldx = _dev2d(X)
    pq, ldq = _dev2d(Wt)
    py, ldy = _dev2d(Y)
    pb = None if bias is None else _vec(bias.reshape(-1), N)
    ws_p, ws_n = (None, 0) if workspace is None else (C.c_void_p(workspace.data_ptr()), workspace.numel())
    _check(lib().qg_linear_forward_act(px, ldx, _dt(X), pq, ldq, _vec(Cw, N), pb, act, py, ldy, _dt(Y), M, N, K,
                                       C.c_float(range_), mode, ws_p, C.c_size_t(ws_n), _stream()), "qg_linear_forward_act")


def linear_forward_q(Xq: torch.Tensor, Cx: torch.Tensor, Wt: torch.Tensor, Cw: torch.Tensor, bias, Y: torch.Tensor,
                     range_: float = 127.0, act: int = 0, y_rowmax: torch.Tensor | None = None) -> None:
    """LinearLayer::forward on activations that are already int8 codes + Cx; optional row maxima of Y for the
    next layer's quantizer (qg_linear_forward_q)."""
    M, K = Xq.shape
    N = Wt.shape[0]
    assert Wt.shape[1] == K and Y.shape == (M, N)
    pa, lda = _dev2d(Xq)
    pq, ldq = _dev2d(Wt)
    py, ldy = _dev2d(Y)
    pb = None if bias is None else _vec(bias.reshape(-1), N)
    prm = None if y_rowmax is None else _vec(y_rowmax, M)
    _check(lib().qg_linear_forward_q(pa, lda, _vec(Cx.reshape(-1), M), pq, ldq, _vec(Cw, N), pb, act, py, ldy, _dt(Y), M, N, K,
                                     C.c_float(range_), prm, None, C.c_size_t(0), _stream()), "qg_linear_forward_q")


def quantize_rows_given_max(Y: torch.Tensor, rowmax: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT):
    """The row quantizer without its reduction: scale from (Y[i,0], rowmax[i]); returns (Xq, Cx)."""
    M, K = Y.shape
    Xq = torch.empty((M, K), dtype=torch.int8, device=Y.device)
    Cx = torch.empty(M, dtype=torch.float32, device=Y.device)
    py, ldy = _dev2d(Y)
    pq, ldq = _dev2d(Xq)
    _check(lib().qg_quantize_rows_given_max(py, _dt(Y), M, K, ldy, C.c_float(range_), mode, _vec(rowmax, M), pq, ldq, _vec(Cx, M),
                                            _stream()), "qg_quantize_rows_given_max")
    return Xq, Cx


def ffn_forward(X, W1t, Cw1, b1, W2t, Cw2, b2, H: torch.Tensor, Y: torch.Tensor, range_: float = 127.0,
                mode: int = MODE_REF_EXACT, Xq: torch.Tensor | None = None, Cx: torch.Tensor | None = None,
                workspace: torch.Tensor | None = None) -> None:
    """ll1.forward -> op_relu -> ll2.forward (transformer.cu:63-71) in one call (qg_ffn_forward): H = relu(x W1 + b1),
    Y = H W2 + b2; the hidden activation's row maxima come out of the first GEMM's epilogue.  Pass Xq / Cx instead of
    X when the input is already quantized (add_layernorm_quant)."""
    M = (Xq if Xq is not None else X).shape[0]
    d_in, d_ff, d_out = W1t.shape[1], W1t.shape[0], W2t.shape[0]
    assert W2t.shape[1] == d_ff and H.shape == (M, d_ff) and Y.shape == (M, d_out)
    px, ldx, dtx = (None, C.c_int64(0), QG_F32) if X is None else (*_dev2d(X), _dt(X))
    pxq, ldxq = (None, C.c_int64(0)) if Xq is None else _dev2d(Xq)
    pcx = None if Cx is None else _vec(Cx.reshape(-1), M)
    p1, ld1 = _dev2d(W1t)
    p2, ld2 = _dev2d(W2t)
    ph, ldh = _dev2d(H)
    py, ldy = _dev2d(Y)
    ws_p, ws_n = (None, 0) if workspace is None else (C.c_void_p(workspace.data_ptr()), workspace.numel())
    _check(lib().qg_ffn_forward(px, ldx, dtx, pxq, ldxq, pcx, p1, ld1, _vec(Cw1, d_ff), None if b1 is None else _vec(b1.reshape(-1), d_ff),
                                p2, ld2, _vec(Cw2, d_out), None if b2 is None else _vec(b2.reshape(-1), d_out), ph, ldh, _dt(H),
                                py, ldy, _dt(Y), M, d_in, d_ff, d_out, C.c_float(range_), mode, ws_p, C.c_size_t(ws_n), _stream()),
           "qg_ffn_forward")


def ffn_workspace_bytes(m: int, d_in: int, d_ff: int, d_out: int) -> int:
    return int(lib().qg_ffn_workspace_bytes(m, d_in, d_ff, d_out))


def _ptr_array(ptrs):
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
    return arr


def gemm_s8_dequant_scatter(Xq, Wt, Cx, Cw, part_ptrs, block_cols: int, ld_part: int, part_dtype: torch.dtype,
                            n: int, range_: float = 127.0) -> None:
    """Row-parallel linear's GEMM (qg_gemm_s8_dequant_scatter): column block b of dequant(Xq . Wt^T) is stored to
    part_ptrs[b] (device addresses of [M, block_cols] matrices, possibly on peer GPUs)."""
    M, K = Xq.shape
    pq, ldq = _dev2d(Xq)
    pw, ldw = _dev2d(Wt)
    _check(lib().qg_gemm_s8_dequant_scatter(pq, ldq, pw, ldw, _vec(Cx, M), _vec(Cw, n), M, n, K, C.c_float(range_),
                                            _ptr_array(part_ptrs), len(part_ptrs), block_cols, C.c_int64(ld_part),
                                            _DT[part_dtype], _stream()), "qg_gemm_s8_dequant_scatter")


def ffn_forward_rowpar(X, W1t, Cw1, b1, W2t, Cw2, H: torch.Tensor, part_ptrs, block_cols: int, ld_part: int,
                       part_dtype: torch.dtype, d_out: int, range_: float = 127.0, mode: int = MODE_REF_EXACT,
                       Xq: torch.Tensor | None = None, Cx: torch.Tensor | None = None,
                       workspace: torch.Tensor | None = None) -> None:
    """This rank's half of a Megatron FFN pair (qg_ffn_forward_rowpar): H = relu(x W1_p + b1_p), then the partial product
    H W2_p scattered by column block into part_ptrs (see include/qgemm.h)."""
    M = (Xq if Xq is not None else X).shape[0]
    d_in, d_ff = W1t.shape[1], W1t.shape[0]
    assert W2t.shape == (d_out, d_ff) and H.shape == (M, d_ff)
    px, ldx, dtx = (None, C.c_int64(0), QG_F32) if X is None else (*_dev2d(X), _dt(X))
    pxq, ldxq = (None, C.c_int64(0)) if Xq is None else _dev2d(Xq)
    pcx = None if Cx is None else _vec(Cx.reshape(-1), M)
    p1, ld1 = _dev2d(W1t)
    p2, ld2 = _dev2d(W2t)
    ph, ldh = _dev2d(H)
    ws_p, ws_n = (None, 0) if workspace is None else (C.c_void_p(workspace.data_ptr()), workspace.numel())
    _check(lib().qg_ffn_forward_rowpar(px, ldx, dtx, pxq, ldxq, pcx, p1, ld1, _vec(Cw1, d_ff),
                                       None if b1 is None else _vec(b1.reshape(-1), d_ff), p2, ld2, _vec(Cw2, d_out), ph, ldh,
                                       _dt(H), _ptr_array(part_ptrs), len(part_ptrs), block_cols, C.c_int64(ld_part),
                                       _DT[part_dtype], M, d_in, d_ff, d_out, C.c_float(range_), mode, ws_p, C.c_size_t(ws_n),
                                       _stream()), "qg_ffn_forward_rowpar")


def reduce_partials(slots: torch.Tensor, bias, out: torch.Tensor, peer_ptrs=(), n: int | None = None, mc_ptr: int = 0,
                    max_ctas: int = 0) -> None:
    """out[:, :n] = ((slots[0] + slots[1]) + ...) + bias in ascending slot order (qg_reduce_partials); the result is also
    stored to the same block of the peers' matrices (device addresses in peer_ptrs, leading dimension = out's)."""
    P, M, bc = slots.shape
    n = bc if n is None else n
    po, ldo = _dev2d(out)
    _check(lib().qg_reduce_partials(C.c_void_p(slots.data_ptr()), C.c_int64(slots.stride(0)), P, _dt(slots),
                                    C.c_int64(slots.stride(1)), None if bias is None else _vec(bias.reshape(-1), n), po,
                                    _ptr_array(peer_ptrs) if len(peer_ptrs) else None, len(peer_ptrs),
                                    C.c_void_p(int(mc_ptr)) if mc_ptr else None, ldo, _dt(out), M, n, int(max_ctas), _stream()),
           "qg_reduce_partials")


def copy_2d_async(dst_ptr: int, dst_pitch: int, src_ptr: int, src_pitch: int, width_bytes: int, rows: int) -> None:
    """Block copy on the copy engines (cudaMemcpy2DAsync), device addresses given as ints (peer-mapped ones included)."""
    _check(lib().qg_copy_2d_async(C.c_void_p(int(dst_ptr)), C.c_int64(dst_pitch), C.c_void_p(int(src_ptr)), C.c_int64(src_pitch),
                                  C.c_int64(width_bytes), int(rows), _stream()), "qg_copy_2d_async")


def add_layernorm_quant(A: torch.Tensor, R, B: torch.Tensor, range_: float = 127.0, mode: int = MODE_REF_EXACT):
    """ADD & NORM (transformer.cu:57-58) that also returns the int8 codes + Cx of its result (qg_add_layernorm_quant_f32)."""
    M, N = A.shape
    ldq = (N + 15) // 16 * 16
    # the padding columns up to the 16-byte leading dimension must read as zero codes; without padding nothing needs clearing
    Xq = (torch.empty if ldq == N else torch.zeros)((M, ldq), dtype=torch.int8, device=A.device)[:, :N]
    Cx = torch.empty(M, dtype=torch.float32, device=A.device)
    pa, lda = _dev2d(A)
    pb, ldb = _dev2d(B)
    pr, ldr = _dev2d(R) if R is not None else (None, C.c_int64(0))
    pq, ldq = _dev2d(Xq)
    _check(lib().qg_add_layernorm_quant_f32(pa, lda, pr, ldr, M, N, pb, ldb, C.c_float(range_), mode, pq, ldq, _vec(Cx, M),
                                            _stream()), "qg_add_layernorm_quant_f32")
    return Xq, Cx


def quantized_mm_host(X, W, range_: float = 127.0, mode: int = MODE_REF_EXACT, bias=None, out=None):
    """Host-buffer form: X, W (and out) are CPU float32 tensors (pinned for full PCIe rate)."""
    assert not X.is_cuda and not W.is_cuda and X.dtype == torch.float32 and W.dtype == torch.float32
    assert X.is_contiguous() and W.is_contiguous()
    M, K = X.shape
    N = W.shape[1]
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32)
    pb = None if bias is None else C.c_void_p(bias.data_ptr())
    _check(lib().qg_quantized_mm_host(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(out.data_ptr()),
                                      M, N, K, C.c_float(range_), mode, pb), "qg_quantized_mm_host")
    return out


MAX_OUTLIER_COLS = 64


def outlier_cols(X: torch.Tensor, thr: float, max_idx: int = 1024):
    """Feature columns of X [M,K] holding at least one |x| > thr: returns (idx int32 [count] on the
    device, ascending; count).  Reads the count back, i.e. synchronises the stream."""
    M, K = X.shape
    idx = torch.empty(max_idx, dtype=torch.int32, device=X.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=X.device)
    px, ldx = _dev2d(X)
    _check(lib().qg_outlier_cols(px, _dt(X), M, K, ldx, C.c_float(thr), C.c_void_p(idx.data_ptr()), max_idx,
                                 C.c_void_p(cnt.data_ptr()), _stream()), "qg_outlier_cols")
    n = int(cnt.item())
    return idx[: min(n, max_idx)], n


class LinearLayer:
    """LinearLayer<float> (src/modules/linear.cuh:7-72), inference only: y = x @ w + b with the
    product on the int8 tensor-core path.  w is [in_dim, out_dim], b is [1, out_dim]; the weights
    are quantized once (column-wise) and cached."""

    def __init__(self, in_dim: int, out_dim: int, device="cuda", dtype=torch.float32, range_: float = 127.0,
                 mode: int = MODE_REF_EXACT):
        self.in_dim, self.out_dim, self.range, self.mode = in_dim, out_dim, range_, mode
        self.w = torch.empty((in_dim, out_dim), dtype=dtype, device=device)
        self.b = torch.empty((1, out_dim), dtype=torch.float32, device=device)
        self._wq = None
        self._ws = None

    def init_uniform(self, generator=None):  # linear.cuh:33-39
        mx = 1.0 / (self.in_dim ** 0.5)
        self.w.uniform_(-mx, mx, generator=generator)
        self.b.uniform_(-mx, mx, generator=generator)
        self._wq = None

    def quantize_weights(self):
        self._wq = prepare_weights(self.w, self.range, self.mode)  # (Wt [N,K] int8, Cw)
        return self._wq

    def forward_outlier(self, x: torch.Tensor, y: torch.Tensor, idx: torch.Tensor) -> None:
        """forward with the feature columns `idx` (int32, device, ascending; e.g. from outlier_cols or a
        calibrated fixed set) routed through the 16-bit side product (LLM.int8()-style decomposition)."""
        assert x.shape[1] == self.in_dim and y.shape == (x.shape[0], self.out_dim)
        assert idx.dtype == torch.int32 and idx.is_cuda and idx.numel() <= MAX_OUTLIER_COLS
        if self._wq is None:
            self.quantize_weights()
        Wt, Cw = self._wq
        M, K, N = x.shape[0], self.in_dim, self.out_dim
        need = int(lib().qg_outlier_workspace_bytes(M, N, K))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        px, ldx = _dev2d(x)
        pw, ldw = _dev2d(self.w)
        pq, ldq = _dev2d(Wt)
        py, ldy = _dev2d(y)
        _check(lib().qg_linear_forward_outlier(px, ldx, _dt(x), pw, ldw, _dt(self.w), pq, ldq, _vec(Cw, N),
                                               _vec(self.b.reshape(-1), N), C.c_void_p(idx.data_ptr()), int(idx.numel()),
                                               py, ldy, _dt(y), M, N, K, C.c_float(self.range), self.mode,
                                               C.c_void_p(self._ws.data_ptr()), C.c_size_t(self._ws.numel()), _stream()),
               "qg_linear_forward_outlier")

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> None:  # linear.cuh:49-56
        assert x.shape[1] == self.in_dim and y.shape == (x.shape[0], self.out_dim)
        if self._wq is None:
            self.quantize_weights()
        Wt, Cw = self._wq
        M, K, N = x.shape[0], self.in_dim, self.out_dim
        need = workspace_bytes(M, N, K)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
        px, ldx = _dev2d(x)
        pq, ldq = _dev2d(Wt)
        py, ldy = _dev2d(y)
        _check(lib().qg_linear_forward(px, ldx, _dt(x), pq, ldq, _vec(Cw, N), _vec(self.b.reshape(-1), N), py, ldy,
                                       _dt(y), M, N, K, C.c_float(self.range), self.mode,
                                       C.c_void_p(self._ws.data_ptr()), C.c_size_t(self._ws.numel()), _stream()),
               "qg_linear_forward")


def op_softmax(A: torch.Tensor, B: torch.Tensor, scale: float = 1.0) -> None:
    """op_softmax (src/ops/op_softmax.cuh:31-41) of fl(A * scale), row-wise, fp32; B may be A."""
    assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.shape == B.shape
    pa, lda = _dev2d(A)
    pb, ldb = _dev2d(B)
    _check(lib().qg_softmax_rows_f32(pa, lda, A.shape[0], A.shape[1], C.c_float(scale), pb, ldb, _stream()),
           "qg_softmax_rows_f32")


def attention_forward(Xq: torch.Tensor, Xkv: torch.Tensor, Wqkv: torch.Tensor, out: torch.Tensor, heads: int, d_k: int,
                      d_v: int, batch: int = 1, range_: float = 127.0, mode: int = MODE_REF_EXACT) -> None:
    """qg_attention_forward: see include/qgemm.h.  Xq [batch*sq, d_model], Xkv [batch*skv, d_model],
    Wqkv [d_model, heads*(2*d_k+d_v)] = [all W_q | all W_k | all W_v], out [batch*sq, heads*d_v]."""
    assert all(t.dtype == torch.float32 and t.is_cuda for t in (Xq, Xkv, Wqkv, out))
    assert Xq.shape[0] % batch == 0 and Xkv.shape[0] % batch == 0
    sq, skv, d_model = Xq.shape[0] // batch, Xkv.shape[0] // batch, Xq.shape[1]
    assert Xkv.shape[1] == d_model and Wqkv.shape == (d_model, heads * (2 * d_k + d_v))
    assert out.shape == (batch * sq, heads * d_v)
    pq, ldq = _dev2d(Xq)
    pkv, ldkv = _dev2d(Xkv)
    pw, ldw = _dev2d(Wqkv)
    po, ldo = _dev2d(out)
    _check(lib().qg_attention_forward(pq, ldq, pkv, ldkv, batch, sq, skv, d_model, pw, ldw, heads, d_k, d_v,
                                      C.c_float(range_), mode, po, ldo, _stream()), "qg_attention_forward")


def attention_forward_prepared(Xq: torch.Tensor, Xkv: torch.Tensor, Wt: torch.Tensor, Cw: torch.Tensor, out: torch.Tensor,
                               heads: int, d_k: int, d_v: int, batch: int = 1, range_: float = 127.0,
                               mode: int = MODE_REF_EXACT) -> None:
    """qg_attention_forward_prepared: the same layer on (Wt, Cw) = prepare_weights(Wqkv) -- same bits, no per-call
    column quantizer."""
    assert all(t.dtype == torch.float32 and t.is_cuda for t in (Xq, Xkv, Cw, out)) and Wt.dtype == torch.int8
    assert Xq.shape[0] % batch == 0 and Xkv.shape[0] % batch == 0
    sq, skv, d_model = Xq.shape[0] // batch, Xkv.shape[0] // batch, Xq.shape[1]
    ntot = heads * (2 * d_k + d_v)
    assert Xkv.shape[1] == d_model and Wt.shape[0] == ntot and Wt.shape[1] >= d_model and Cw.numel() == ntot
    assert out.shape == (batch * sq, heads * d_v)
    pq, ldq = _dev2d(Xq)
    pkv, ldkv = _dev2d(Xkv)
    pw, ldw = _dev2d(Wt)
    po, ldo = _dev2d(out)
    _check(lib().qg_attention_forward_prepared(pq, ldq, pkv, ldkv, batch, sq, skv, d_model, pw, ldw, _vec(Cw.reshape(-1), ntot),
                                               heads, d_k, d_v, C.c_float(range_), mode, po, ldo, _stream()),
           "qg_attention_forward_prepared")


class AttentionLayer:
    """AttentionLayer<float> (src/modules/attention.cuh:10-70): single head, no bias, no mask.
    W_q, W_k [d_model, d_k] and W_v [d_model, d_v] are views of one [d_model, 2*d_k+d_v] parameter, so
    the three projections of forward() run as ONE quantized product (bit-identical to three).
    forward(X, out) is the reference's; forward(Xq, Xkv, out) is the cross-attention form
    src/transformer.cu:37,132 calls."""

    def __init__(self, d_model: int, d_k: int, d_v: int, device="cuda", range_: float = 127.0, mode: int = MODE_REF_EXACT):
        self.d_model, self.d_k, self.d_v, self.range, self.mode = d_model, d_k, d_v, range_, mode
        self.W_qkv = torch.empty((d_model, 2 * d_k + d_v), dtype=torch.float32, device=device)
        self.W_q = self.W_qkv[:, :d_k]
        self.W_k = self.W_qkv[:, d_k:2 * d_k]
        self.W_v = self.W_qkv[:, 2 * d_k:]

    def parameters(self):  # attention.cuh:32-38
        return [self.W_q, self.W_k, self.W_v]

    def init_uniform(self, generator=None):  # attention.cuh:40-45
        mx = 1.0 / (self.d_k ** 0.5)
        self.W_qkv.uniform_(-mx, mx, generator=generator)

    def forward(self, *args) -> None:
        if len(args) == 2:
            (X, out), Xkv = args, args[0]
        else:
            X, Xkv, out = args
        attention_forward(X, Xkv, self.W_qkv, out, 1, self.d_k, self.d_v, 1, self.range, self.mode)


class MultiHeadAttention:
    """The per-head AttentionLayer loop + host-side concat of src/transformer.cu:27-50 as one call:
    all heads' projections in one quantized product, scores / softmax / P V batched over
    (sequence, head), heads written side by side.  `batch` independent sequences per call."""

    def __init__(self, d_model: int, n_heads: int, device="cuda", range_: float = 127.0, mode: int = MODE_REF_EXACT):
        assert d_model % n_heads == 0
        self.d_model, self.n_heads, self.range, self.mode = d_model, n_heads, range_, mode
        self.d_k = self.d_v = d_model // n_heads  # transformer.cu:21-22
        self.W_qkv = torch.empty((d_model, n_heads * (2 * self.d_k + self.d_v)), dtype=torch.float32, device=device)
        self._wq = None  # (Wt, Cw) of quantize_weights(); dropped by init_uniform

    def head_weights(self, h: int):
        """(W_q, W_k, W_v) views of head h, each [d_model, d_k]."""
        H, dk, dv = self.n_heads, self.d_k, self.d_v
        return (self.W_qkv[:, h * dk:(h + 1) * dk], self.W_qkv[:, H * dk + h * dk:H * dk + (h + 1) * dk],
                self.W_qkv[:, 2 * H * dk + h * dv:2 * H * dk + (h + 1) * dv])

    def init_uniform(self, generator=None):
        mx = 1.0 / (self.d_k ** 0.5)
        self.W_qkv.uniform_(-mx, mx, generator=generator)
        self._wq = None

    def quantize_weights(self) -> None:
        """Column-quantize W_qkv once (call again after changing W_qkv); forward(prepared=True) then skips the per-call pass."""
        self._wq = prepare_weights(self.W_qkv, self.range, self.mode)

    def forward(self, Xq: torch.Tensor, Xkv: torch.Tensor, out: torch.Tensor, batch: int = 1, prepared: bool = False) -> None:
        """prepared=False: the weights are quantized on every call, as op_quantized_mm does (W_qkv may change between calls).
        prepared=True: the codes of quantize_weights() (made on first use) -- bit-identical output."""
        if not prepared:
            attention_forward(Xq, Xkv, self.W_qkv, out, self.n_heads, self.d_k, self.d_v, batch, self.range, self.mode)
            return
        if self._wq is None:
            self.quantize_weights()
        Wt, Cw = self._wq
        attention_forward_prepared(Xq, Xkv, Wt, Cw, out, self.n_heads, self.d_k, self.d_v, batch, self.range, self.mode)
