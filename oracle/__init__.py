"""ctypes front-end of the CPU oracle (oracle/qoracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this package; the product (quantized-gemm-for-transformer-inference_b200) never does.

Each wrapper takes/returns numpy arrays and maps 1:1 onto a C function whose comment cites the
reference file:line it restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqoracle.so")

MODE_REF_EXACT = 0
MODE_TRUE_ABSMAX = 1


def build(force: bool = False) -> str:
    """Compile oracle/qoracle.c (and oracle/_ref when /root/reference exists)."""
    src = os.path.join(_HERE, "qoracle.c")
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "libqoracle.so"], check=True, capture_output=True)
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.qo_quant_code.restype = C.c_int
        _lib.qo_quant_code.argtypes = [C.c_float, C.c_float]
        _lib.qo_max_partial_sum.restype = C.c_int64
        _lib.qo_signed_mean_f32.restype = C.c_float
        _lib.qo_num_threads.restype = C.c_int
        _lib.qo_quantized_mm_f32.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    a = np.asarray(a)
    if a.dtype in (np.float16,):
        a = a.astype(np.float32)
    assert a.dtype == np.float32, a.dtype
    return np.ascontiguousarray(a)


def num_threads() -> int:
    return int(lib().qo_num_threads())


def absmax_rows(X, mode=MODE_REF_EXACT):
    X = _f32(X)
    M, K = X.shape
    out = np.empty(M, np.float32)
    lib().qo_absmax_rows_f32(_p(X), C.c_int(M), C.c_int(K), C.c_int64(K), C.c_int(mode), _p(out))
    return out


def absmax_cols(W, mode=MODE_REF_EXACT):
    W = _f32(W)
    K, N = W.shape
    out = np.empty(N, np.float32)
    lib().qo_absmax_cols_f32(_p(W), C.c_int(K), C.c_int(N), C.c_int64(N), C.c_int(mode), _p(out))
    return out


def inv_divide(c, b=127.0):
    c = _f32(c)
    out = np.empty_like(c)
    lib().qo_inv_divide_f32(_p(c), C.c_int64(c.size), C.c_float(b), _p(out))
    return out


def quant_code(x: float, s: float) -> int:
    return int(lib().qo_quant_code(C.c_float(x), C.c_float(s)))


def quantize_rows(X, sx):
    X = _f32(X)
    sx = _f32(sx).reshape(-1)
    M, K = X.shape
    out = np.empty((M, K), np.int8)
    lib().qo_quantize_rows_f32(_p(X), C.c_int(M), C.c_int(K), C.c_int64(K), _p(sx), _p(out), C.c_int64(K))
    return out


def quantize_cols(W, sw):
    W = _f32(W)
    sw = _f32(sw).reshape(-1)
    K, N = W.shape
    out = np.empty((K, N), np.int8)
    lib().qo_quantize_cols_f32(_p(W), C.c_int(K), C.c_int(N), C.c_int64(N), _p(sw), _p(out), C.c_int64(N))
    return out


def absmax_quant_rows(X, range_=127.0, mode=MODE_REF_EXACT):
    """a1 + a3 + a4 of SURVEY section 8: returns (Xq int8 [M,K], Cx f32 [M])."""
    Cx = absmax_rows(X, mode)
    return quantize_rows(X, inv_divide(Cx, range_)), Cx


def absmax_quant_cols(W, range_=127.0, mode=MODE_REF_EXACT):
    """a2 + a3 + a4: returns (Wq int8 [K,N], Cw f32 [N])."""
    Cw = absmax_cols(W, mode)
    return quantize_cols(W, inv_divide(Cw, range_)), Cw


def gemm_s8s8s32(A, B):
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    M, K = A.shape
    K2, N = B.shape
    assert K == K2
    out = np.empty((M, N), np.int32)
    lib().qo_gemm_s8s8s32(_p(A), _p(B), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K), C.c_int64(N),
                          _p(out), C.c_int64(N))
    return out


def gemm_s8_reff32(A, B):
    """Literal emulation of the reference's fp32-FMA 'int' accumulator (small shapes)."""
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    M, K = A.shape
    _, N = B.shape
    out = np.empty((M, N), np.int32)
    lib().qo_gemm_s8_reff32(_p(A), _p(B), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K), C.c_int64(N),
                            _p(out), C.c_int64(N))
    return out


def max_partial_sum(A, B) -> int:
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    M, K = A.shape
    _, N = B.shape
    return int(lib().qo_max_partial_sum(_p(A), _p(B), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K),
                                        C.c_int64(N)))


def gemm_f32_ref(A, B):
    A = _f32(A)
    B = _f32(B)
    M, K = A.shape
    _, N = B.shape
    out = np.empty((M, N), np.float32)
    lib().qo_gemm_f32_ref(_p(A), _p(B), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K), C.c_int64(N),
                          _p(out), C.c_int64(N))
    return out


def dequant(acc, Cx, Cw, range_=127.0, bias=None):
    acc = np.ascontiguousarray(acc, np.int32)
    Cx = _f32(Cx).reshape(-1)
    Cw = _f32(Cw).reshape(-1)
    M, N = acc.shape
    b = None if bias is None else _f32(bias).reshape(-1)
    out = np.empty((M, N), np.float32)
    lib().qo_dequant_f32(_p(acc), C.c_int64(N), _p(Cx), _p(Cw), _p(b), C.c_int(M), C.c_int(N),
                         C.c_float(range_), _p(out), C.c_int64(N))
    return out


def quantized_mm(X, W, range_=127.0, mode=MODE_REF_EXACT, bias=None, return_parts=False):
    """op_quantized_mm (src/ops/op_mm.cuh:67-101), optionally followed by the LinearLayer bias add."""
    X = _f32(X)
    W = _f32(W)
    M, K = X.shape
    K2, N = W.shape
    assert K == K2
    O = np.empty((M, N), np.float32)
    b = None if bias is None else _f32(bias).reshape(-1)
    if return_parts:
        Cx = np.empty(M, np.float32)
        Cw = np.empty(N, np.float32)
        Xq = np.empty((M, K), np.int8)
        Wq = np.empty((K, N), np.int8)
        acc = np.empty((M, N), np.int32)
    else:
        Cx = Cw = Xq = Wq = acc = None
    rc = lib().qo_quantized_mm_f32(_p(X), _p(W), _p(O), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K),
                                   C.c_int64(N), C.c_int64(N), C.c_float(range_), C.c_int(mode), _p(b),
                                   _p(Cx), _p(Cw), _p(Xq), _p(Wq), _p(acc))
    if rc != 0:
        raise MemoryError("qo_quantized_mm_f32")
    if return_parts:
        return O, dict(Cx=Cx, Cw=Cw, Xq=Xq, Wq=Wq, acc=acc)
    return O


def outlier_mask(A, thr):
    A = _f32(A)
    M, K = A.shape
    out = np.empty((M, K), np.float32)
    lib().qo_outlier_mask_f32(_p(A), C.c_int(M), C.c_int(K), C.c_int64(K), C.c_float(thr), _p(out),
                              C.c_int64(K))
    return out


def signed_mean(A) -> float:
    A = _f32(A)
    M, N = A.shape
    return float(lib().qo_signed_mean_f32(_p(A), C.c_int(M), C.c_int(N), C.c_int64(N)))


# ---------------------------------------------------------------------------------------------
# Extensions the reference does not implement (parity UNPINNED; the spec is ours)
# ---------------------------------------------------------------------------------------------

def round_to(dtype: str, y):
    """fp32 -> output dtype conversion of the fused epilogue (round to nearest even)."""
    if dtype == "f32":
        return y
    if dtype == "f16":
        return y.astype(np.float16)
    if dtype == "bf16":
        import torch

        return torch.from_numpy(y).to(torch.bfloat16)
    raise ValueError(dtype)


def outlier_columns(X, thr):
    """K-indices that hold at least one |x| > thr (strict, NaN counts) -- column-reduction of
    the reference's elementwise mask primitive (src/ops/op_elemwise.cuh:292-306,698-708)."""
    return np.nonzero(outlier_mask(X, thr).max(axis=0) > 0)[0].astype(np.int32)


def side_gemm(Xo, Wo):
    """fp32-accumulated side product of already-rounded operands (k ascending fma from +0)."""
    Xo = _f32(Xo)
    Wo = _f32(Wo)
    M, no = Xo.shape
    N = Wo.shape[1]
    out = np.empty((M, N), np.float32)
    lib().qo_side_gemm_f32(_p(Xo), _p(Wo), C.c_int(M), C.c_int(N), C.c_int(no), C.c_int64(no), C.c_int64(N), _p(out),
                           C.c_int64(N))
    return out


def _round_side(a, side_dtype):
    if side_dtype == "bf16":
        import torch

        return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()
    return a.astype(np.float16).astype(np.float32)


def quantized_mm_outlier(X, W, thr, range_=127.0, mode=MODE_REF_EXACT, bias=None, side_dtype="f16", idx=None):
    """LLM.int8()-style mixed decomposition (parity unpinned; specification ours):
    int8 path on X with the outlier feature columns zeroed (row scales from the remaining entries),
    W column-quantized over ALL rows (so the weights can be prepared once), plus a 16-bit side
    product X[:,O] @ W[O,:] accumulated in fp32; out = fl(fl(dequant + side) + bias)."""
    X = _f32(X)
    W = _f32(W)
    if idx is None:
        idx = outlier_columns(X, thr)
    idx = np.asarray(idx, np.int32)
    Xr = X.copy()
    Xr[:, idx] = 0.0
    O, parts = quantized_mm(Xr, W, range_, mode, None, return_parts=True)
    if idx.size:
        side = side_gemm(_round_side(X[:, idx], side_dtype), _round_side(W[idx, :], side_dtype))
        O = (O + side).astype(np.float32)
    if bias is not None:
        O = (O + _f32(bias).reshape(1, -1)).astype(np.float32)
    parts["outlier_idx"] = idx
    return O, parts


# ---- attention: composition of the functions above -----------------------

def softmax_rows(A, scale=1.0):
    """op_multiply(A, scale) + op_softmax (attention.cuh:62-68, op_softmax.cuh:6-29) on the host."""
    A = _f32(A)
    M, N = A.shape
    out = np.empty((M, N), np.float32)
    lib().qo_softmax_rows_f32(_p(A), C.c_int64(N), C.c_int(M), C.c_int(N), C.c_float(scale), _p(out), C.c_int64(N))
    return out


def attention_forward(Xq, Xkv, Wq, Wk, Wv, range_=127.0, mode=MODE_REF_EXACT, return_parts=False):
    """AttentionLayer::forward (attention.cuh:47-70), one head, one sequence, with the three
    projections routed to op_quantized_mm (the re-pointing of SURVEY F2); Xq is Xkv for the
    reference's 2-argument form."""
    Q = quantized_mm(Xq, Wq, range_, mode)
    K = quantized_mm(Xkv, Wk, range_, mode)
    V = quantized_mm(Xkv, Wv, range_, mode)
    S = gemm_f32_ref(Q, np.ascontiguousarray(K.T))                 # op_mm(Q, K.transpose(), QK_T)
    d_k = Wq.shape[1]
    P = softmax_rows(S, np.float32(1.0 / np.sqrt(np.float64(d_k))))  # T scale_factor = 1.0 / std::sqrt(d_k)
    out = gemm_f32_ref(P, V)
    if return_parts:
        return {"Q": Q, "K": K, "V": V, "S": S, "P": P, "out": out}
    return out


def multi_head_attention(Xq, Xkv, Wqkv, heads, d_k, d_v, batch=1, range_=127.0, mode=MODE_REF_EXACT):
    """The head loop + concat of transformer.cu:27-50 over `batch` independent sequences.
    Wqkv columns: [W_q of every head | W_k of every head | W_v of every head]."""
    Xq, Xkv, Wqkv = _f32(Xq), _f32(Xkv), _f32(Wqkv)
    sq, skv = Xq.shape[0] // batch, Xkv.shape[0] // batch
    out = np.empty((batch * sq, heads * d_v), np.float32)
    for b in range(batch):
        xq, xkv = Xq[b * sq:(b + 1) * sq], Xkv[b * skv:(b + 1) * skv]
        for h in range(heads):
            wq = Wqkv[:, h * d_k:(h + 1) * d_k]
            wk = Wqkv[:, heads * d_k + h * d_k:heads * d_k + (h + 1) * d_k]
            wv = Wqkv[:, 2 * heads * d_k + h * d_v:2 * heads * d_k + (h + 1) * d_v]
            out[b * sq:(b + 1) * sq, h * d_v:(h + 1) * d_v] = attention_forward(xq, xkv, wq, wk, wv, range_, mode)
    return out


# ---- encoder block: src/transformer.cu:24-76 with persistent weights -----

def add_layernorm(A, R=None):
    """op_add(A, R) + op_layernorm (transformer.cu:57-58; op_layernorm.cuh:6-32) on the host."""
    A = _f32(A)
    M, N = A.shape
    out = np.empty((M, N), np.float32)
    if R is not None:
        R = _f32(R)
    lib().qo_add_layernorm_f32(_p(A), C.c_int64(N), _p(R) if R is not None else None, C.c_int64(N), C.c_int(M), C.c_int(N),
                               _p(out), C.c_int64(N))
    return out


def relu(A):
    """ReluFunc, src/ops/op_elemwise.cuh:181-195: x < 0 ? 0 : x (NaN and -0 pass through)."""
    A = _f32(A)
    return np.where(A < 0, np.float32(0), A).astype(np.float32)


def encoder_block(X, Wqkv, W_O, W1, b1, W2, b2, heads, batch=1, range_=127.0, mode=MODE_REF_EXACT):
    """One iteration of the Encoder loop (transformer.cu:24-76), every op_mm / LinearLayer on the quantized
    path, the residual taken from the attention output exactly as the reference does (:57, :74)."""
    d_model = X.shape[1]
    d = d_model // heads
    mh = multi_head_attention(X, X, Wqkv, heads, d, d, batch, range_, mode)        # :27-50
    out = quantized_mm(mh, W_O, range_, mode)                                        # :54  op_mm(multiHeadOut, W_O, output)
    out = add_layernorm(out, mh)                                                     # :57-58
    ffn = relu(quantized_mm(out, W1, range_, mode, bias=b1))                         # :63-67
    out = quantized_mm(ffn, W2, range_, mode, bias=b2)                               # :69-71
    return add_layernorm(out, mh)                                                    # :74-75


def megatron_ffn(X, W1, b1, W2, b2, bounds, range_=127.0, mode=MODE_REF_EXACT, h_dtype="f32", part_dtype="f32",
                 out_dtype="f32", return_parts=False):
    """Megatron pairing of the FFN of src/transformer.cu:63-71 (ll1 -> relu -> ll2) over len(bounds) ranks; the
    specification is ours (parity UNPINNED: the reference has no multi-GPU code), see include/qgemm.h.

    bounds[p] = (lo, hi): rank p's slice of the d_ff hidden features.  H_p = relu(quantized_mm(X, W1[:, lo:hi]) + b1[lo:hi])
    rounded to h_dtype (the same bits as those columns of the single-GPU layer); part_p = quantized_mm(H_p, W2[lo:hi, :])
    with the row scales of H_p and the column scales of W2[lo:hi, :] (per-slice scales), rounded to part_dtype;
    y = ((part_0 + part_1) + ...) + b2, fp32 additions in ascending rank order, rounded to out_dtype."""
    X = _f32(X)
    parts = []
    for lo, hi in bounds:
        h = relu(quantized_mm(X, W1[:, lo:hi], range_, mode, bias=None if b1 is None else np.asarray(b1).reshape(-1)[lo:hi]))
        h = np.asarray(_to_f32(round_to(h_dtype, h)))
        part = quantized_mm(h, W2[lo:hi, :], range_, mode)
        parts.append(np.asarray(_to_f32(round_to(part_dtype, part))))
    y = parts[0].copy()
    for p in parts[1:]:
        y = (y + p).astype(np.float32)
    if b2 is not None:
        y = (y + _f32(np.asarray(b2).reshape(1, -1))).astype(np.float32)
    y = round_to(out_dtype, y)
    return (y, parts) if return_parts else y


def _to_f32(a):
    """numpy float32 view of round_to()'s result (bf16 comes back as a torch tensor)."""
    if isinstance(a, np.ndarray):
        return a.astype(np.float32)
    return a.float().numpy()


# ---------------------------------------------------------------------------------------------
# Timed CPU baseline (oracle/qfast.c): the same pipeline, threaded and vectorised (AVX-512 VNNI
# int8 GEMM behind a run-time check).  Used by bench.py's cpu_baseline / --impl reference legs;
# tests/test_oracle_cpu.py checks it against the functions above bit for bit.
# ---------------------------------------------------------------------------------------------
_SO_FAST = os.path.join(_HERE, "libqfast.so")
_fast = None


def fast_lib() -> C.CDLL:
    global _fast
    if _fast is None:
        src = os.path.join(_HERE, "qfast.c")
        if not os.path.exists(_SO_FAST) or (os.path.exists(src) and os.path.getmtime(_SO_FAST) < os.path.getmtime(src)):
            subprocess.run(["make", "-C", _HERE, "libqfast.so"], check=True, capture_output=True)
        _fast = C.CDLL(_SO_FAST)
        _fast.qf_num_threads.restype = C.c_int
        _fast.qf_gemm_kernel.restype = C.c_int
        _fast.qf_gemm_s8s8s32.restype = C.c_int
        _fast.qf_quantized_mm_f32.restype = C.c_int
    return _fast


def fast_set_threads(n: int | None = None) -> int:
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm asks for the cores it may actually use."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    fast_lib().qf_set_threads(C.c_int(int(n)))
    return int(fast_lib().qf_num_threads())


def fast_kernel_name() -> str:
    return {2: "avx512_vnni vpdpbusd 6x64 register block", 0: "portable blocked int32 (compiler-vectorised)"}[
        int(fast_lib().qf_gemm_kernel())]


def fast_gemm_s8s8s32(A, B):
    A = np.ascontiguousarray(A, np.int8)
    B = np.ascontiguousarray(B, np.int8)
    M, K = A.shape
    _, N = B.shape
    out = np.empty((M, N), np.int32)
    rc = fast_lib().qf_gemm_s8s8s32(_p(A), _p(B), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K), C.c_int64(N), _p(out),
                                    C.c_int64(N))
    if rc != 0:
        raise MemoryError("qf_gemm_s8s8s32")
    return out


def fast_quantized_mm(X, W, range_=127.0, mode=MODE_REF_EXACT, bias=None, return_parts=False, out=None):
    X = _f32(X)
    W = _f32(W)
    M, K = X.shape
    _, N = W.shape
    O = np.empty((M, N), np.float32) if out is None else out
    b = None if bias is None else _f32(bias).reshape(-1)
    if return_parts:
        Cx, Cw = np.empty(M, np.float32), np.empty(N, np.float32)
        Xq, Wq, acc = np.empty((M, K), np.int8), np.empty((K, N), np.int8), np.empty((M, N), np.int32)
    else:
        Cx = Cw = Xq = Wq = acc = None
    rc = fast_lib().qf_quantized_mm_f32(_p(X), _p(W), _p(O), C.c_int(M), C.c_int(N), C.c_int(K), C.c_int64(K), C.c_int64(N),
                                        C.c_int64(N), C.c_float(range_), C.c_int(mode), _p(b), _p(Cx), _p(Cw), _p(Xq),
                                        _p(Wq), _p(acc))
    if rc != 0:
        raise MemoryError("qf_quantized_mm_f32")
    if return_parts:
        return O, dict(Cx=Cx, Cw=Cw, Xq=Xq, Wq=Wq, acc=acc)
    return O


def decoder_block(X, E, sa_Wqkv, sa_W_O, ca_Wqkv, ca_W_O, W1, b1, W2, b2, heads, batch=1, range_=127.0, mode=MODE_REF_EXACT):
    """One iteration of the Decoder loop (transformer.cu:91-166): self-attention, ADD & NORM, cross-attention
    (queries from the decoder stream, keys / values from the encoder output E), ADD & NORM, FFN, ADD & NORM --
    every product on weights through op_quantized_mm, residuals taken from the attention outputs (:123,148,165)."""
    d = X.shape[1] // heads
    mh = multi_head_attention(X, X, sa_Wqkv, heads, d, d, batch, range_, mode)      # :97-116
    out = add_layernorm(quantized_mm(mh, sa_W_O, range_, mode), mh)                  # :117-124
    mh = multi_head_attention(out, E, ca_Wqkv, heads, d, d, batch, range_, mode)    # :127-141
    out = add_layernorm(quantized_mm(mh, ca_W_O, range_, mode), mh)                  # :142-149
    ffn = relu(quantized_mm(out, W1, range_, mode, bias=b1))                         # :152-157
    out = quantized_mm(ffn, W2, range_, mode, bias=b2)                               # :159-161
    return add_layernorm(out, mh)                                                    # :165-166
