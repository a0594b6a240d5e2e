"""GPU parity tests: every call goes through the C ABI (libqgemm.so) and is compared with the CPU
oracle on the same seeded inputs.  Bars: int8 codes, absmax vectors and
int32 accumulators bit-exact; fp32 output bit-exact (same three roundings); fp16/bf16 output
equal to round-to-nearest of the fp32 result."""
import ctypes as C
import json
import os
import zlib

import numpy as np
import pytest
import torch

from conftest import make_edge_matrix

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
DEV = "cuda"
TORCH_DT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def seed_of(*args) -> int:
    return zlib.crc32(repr(args).encode())


def same_f32(a: np.ndarray, b: np.ndarray) -> bool:
    """Bit-level equality that treats every NaN alike but distinguishes +0 from -0."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    if a.shape != b.shape:
        return False
    nan = np.isnan(a)
    if not np.array_equal(nan, np.isnan(b)):
        return False
    return np.array_equal(a.view(np.int32)[~nan], b.view(np.int32)[~nan])


def to_dev(x: np.ndarray, dt: str = "f32") -> torch.Tensor:
    return torch.from_numpy(x).to(TORCH_DT[dt]).to(DEV)


def as_f32_np(t: torch.Tensor) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def padded(t: torch.Tensor, mult: int) -> torch.Tensor:
    """Copy of t whose leading dimension is a multiple of `mult` elements (stride(0) != shape[1])."""
    r, c = t.shape
    ld = (c + mult - 1) // mult * mult
    buf = torch.zeros((r, ld), dtype=t.dtype, device=t.device)
    buf[:, :c] = t
    return buf[:, :c]


# ------------------------------------------------------------------------------------------
# golden vector of the reference's own test
# ------------------------------------------------------------------------------------------
def test_golden_3x3_through_c_abi(qg):
    g = json.load(open(os.path.join(HERE, "golden", "test_quantize_3x3.json")))
    X = to_dev(np.array(g["X"], np.float32))
    W = to_dev(np.array(g["W"], np.float32))
    O = torch.empty((3, 2), device=DEV)
    qg.op_quantized_mm(X, W, O, g["range"])
    np.testing.assert_allclose(as_f32_np(O), np.array(g["out"], np.float32), rtol=0, atol=5e-7)
    Xq, Cx = qg.absmax_quant_rows(X)
    Wq, Cw = qg.absmax_quant_cols(W)
    assert Xq.cpu().numpy().tolist() == g["Xq"] and Wq.cpu().numpy().tolist() == g["Wq"]
    assert Cx.cpu().numpy().tolist() == g["Cx"] and Cw.cpu().numpy().tolist() == g["Cw"]
    acc = torch.empty((3, 2), dtype=torch.int32, device=DEV)
    qg.op_mm(Xq, Wq, acc)
    assert acc.cpu().numpy().tolist() == g["acc"]


# ------------------------------------------------------------------------------------------
# quantizers (a1-a4)
# ------------------------------------------------------------------------------------------
ROW_SHAPES = [(3, 3), (9, 1), (11, 5), (40, 128), (40, 512), (33, 1000), (70, 4096), (300, 1024),
              (20, 8192), (9, 16384), (6, 20000), (4, 36864), (5, 1002)]


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", ROW_SHAPES)
def test_row_quantizer_bit_exact(qg, oracle, shape, dt):
    rng = np.random.default_rng(seed_of(shape, dt))
    X = to_dev(make_edge_matrix(rng, *shape), dt)
    Xh = as_f32_np(X)  # what the kernel sees, widened exactly
    for mode in (qg.MODE_REF_EXACT, qg.MODE_TRUE_ABSMAX):
        Xq, Cx = qg.absmax_quant_rows(X, 127.0, mode)
        eq, ecx = oracle.absmax_quant_rows(Xh, 127.0, mode)
        assert same_f32(Cx.cpu().numpy(), ecx), f"Cx mismatch mode {mode}"
        assert np.array_equal(Xq.cpu().numpy(), eq), f"codes mismatch mode {mode}"


COL_SHAPES = [(3, 2), (1, 7), (5, 11), (128, 40), (512, 64), (1000, 36), (4096, 72), (1024, 300),
              (777, 1024), (4097, 260), (64, 4096)]


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", COL_SHAPES)
def test_col_quantizer_bit_exact(qg, oracle, shape, dt):
    rng = np.random.default_rng(seed_of(shape, dt, "c"))
    K, N = shape
    W = to_dev(np.ascontiguousarray(make_edge_matrix(rng, N, K).T), dt)
    Wh = as_f32_np(W)
    for mode in (qg.MODE_REF_EXACT, qg.MODE_TRUE_ABSMAX):
        Wq, Cw = qg.absmax_quant_cols(W, 127.0, mode)
        eq, ecw = oracle.absmax_quant_cols(Wh, 127.0, mode)
        assert same_f32(Cw.cpu().numpy(), ecw), f"Cw mismatch mode {mode}"
        assert np.array_equal(Wq.cpu().numpy(), eq), f"codes mismatch mode {mode}"


@pytest.mark.parametrize("shape,dt", [((2048, 2048, 2048), "f32"), ((1024, 4096, 4096), "f32"), ((1024, 512, 8192), "f32"),
                                      ((2048, 2048, 4096), "f16"), ((1024, 2048, 8192), "bf16"), ((600, 1000, 4096), "f32"),
                                      ((64, 64, 128), "f32")])
def test_both_quantizers_in_one_call_bit_exact(qg, oracle, shape, dt):
    """qg_absmax_quant_rows_cols: on large problems column pass 2 runs side by side with the row quantizer in one launch
    (the first five shapes take that path, the last two the separate launches) -- same codes and scales either way."""
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape, dt, "both"))
    X = to_dev(make_edge_matrix(rng, M, K), dt)
    W = to_dev(np.ascontiguousarray(make_edge_matrix(rng, N, K).T), dt)
    Xh, Wh = as_f32_np(X), as_f32_np(W)
    for mode in (qg.MODE_REF_EXACT, qg.MODE_TRUE_ABSMAX):
        Xq = torch.full((M, K), 77, dtype=torch.int8, device=DEV)
        Wq = torch.full((K, N), 77, dtype=torch.int8, device=DEV)
        Cx, Cw = torch.full((M,), -1.0, device=DEV), torch.full((N,), -1.0, device=DEV)
        qg.absmax_quant_rows_cols(X, W, Xq, Cx, Wq, Cw, 127.0, mode)
        ex, ecx = oracle.absmax_quant_rows(Xh, 127.0, mode)
        ew, ecw = oracle.absmax_quant_cols(Wh, 127.0, mode)
        assert same_f32(Cx.cpu().numpy(), ecx) and same_f32(Cw.cpu().numpy(), ecw), f"scales mismatch mode {mode}"
        assert np.array_equal(Xq.cpu().numpy(), ex), f"row codes mismatch mode {mode}"
        assert np.array_equal(Wq.cpu().numpy(), ew), f"column codes mismatch mode {mode}"


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", COL_SHAPES + [(4096, 4096), (300, 2050)])
def test_prepare_weights_is_transposed_column_quantizer(qg, oracle, shape, dt):
    """qg_prepare_weights: same codes and Cw as the column quantizer, stored as Wt [N,K]."""
    rng = np.random.default_rng(seed_of(shape, dt, "t"))
    K, N = shape
    W = to_dev(np.ascontiguousarray(make_edge_matrix(rng, N, K).T), dt)
    Wh = as_f32_np(W)
    for mode in (qg.MODE_REF_EXACT, qg.MODE_TRUE_ABSMAX):
        Wt, Cw = qg.prepare_weights(W, 127.0, mode)
        eq, ecw = oracle.absmax_quant_cols(Wh, 127.0, mode)
        assert same_f32(Cw.cpu().numpy(), ecw), f"Cw mismatch mode {mode}"
        assert np.array_equal(Wt.cpu().numpy(), eq.T), f"codes mismatch mode {mode}"


@pytest.mark.parametrize("variant", ["SIMT", "TC_1SM", "TC_2SM"])
@pytest.mark.parametrize("shape", [(128, 256, 128), (200, 300, 1000), (512, 768, 1024), (3, 2, 3), (129, 257, 129)])
def test_gemm_on_prepared_weights(qg, oracle, shape, variant):
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape, "prep"))
    A, B = codes(rng, M, K), codes(rng, K, N)
    Cx = rng.random(M, dtype=np.float32) + 0.05
    Cw = rng.random(N, dtype=np.float32) + 0.05
    bias = rng.standard_normal(N).astype(np.float32)
    dA = padded(torch.from_numpy(A).to(DEV), 16)
    dBt = padded(torch.from_numpy(np.ascontiguousarray(B.T)).to(DEV), 16)
    acc = padded(torch.empty((M, N), dtype=torch.int32, device=DEV), 4)
    out = torch.empty((M, N), device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.gemm_s8t_dequant(dA, dBt, None, None, acc)
        qg.gemm_s8t_dequant(dA, dBt, to_dev(Cx), to_dev(Cw), out, 127.0, to_dev(bias))
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    exp = oracle.gemm_s8s8s32(A, B)
    assert np.array_equal(acc.cpu().numpy(), exp)
    assert same_f32(out.cpu().numpy(), oracle.dequant(exp, Cx, Cw, 127.0, bias))


@pytest.mark.parametrize("variant", ["TC_1SM", "TC_2SM"])
@pytest.mark.parametrize("shape", [(4096, 4096, 512), (3000, 4000, 256), (4096, 4096, 4096)])
def test_gemm_prepared_weights_tail_split_tiles(qg, oracle, shape, variant):
    """Shapes whose last wave is cut into half-width tiles by the tile scheduler (more than one wave,
    remainder at most half the machine): sampled rows exact against the oracle + a checksum of all rows."""
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape, "tail"))
    A, B = codes(rng, M, K), codes(rng, K, N)
    dA = torch.from_numpy(A).to(DEV)
    dBt = torch.from_numpy(np.ascontiguousarray(B.T)).to(DEV)
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.gemm_s8t_dequant(dA, dBt, None, None, acc)
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    got = acc.cpu().numpy()
    rows = np.sort(rng.choice(M, 64, replace=False))
    assert np.array_equal(got[rows], oracle.gemm_s8s8s32(A[rows], B))
    bsum = B.astype(np.int64).sum(axis=1)
    assert np.array_equal(got.astype(np.int64).sum(axis=1), A.astype(np.int64) @ bsum)
    # column checksum as well (catches tiles written to the wrong column block)
    asum = A.astype(np.int64).sum(axis=0)
    assert np.array_equal(got.astype(np.int64).sum(axis=0), asum @ B.astype(np.int64))


def test_unfused_ops_match_reference_sequence(qg, oracle):
    """op_absmax -> op_inv_divide -> op_multiply<float,int8_t>, the reference's own call sequence
    (src/ops/op_mm.cuh:76-89), through the individual entry points."""
    rng = np.random.default_rng(5)
    X = make_edge_matrix(rng, 50, 640)
    W = np.ascontiguousarray(make_edge_matrix(rng, 48, 640).T)
    dX, dW = to_dev(X), to_dev(W)
    Cx = torch.empty((50, 1), device=DEV)
    Cw = torch.empty((1, 48), device=DEV)
    qg.op_absmax(dX, Cx)
    qg.op_absmax(dW, Cw)
    assert same_f32(Cx.cpu().numpy().ravel(), oracle.absmax_rows(X))
    assert same_f32(Cw.cpu().numpy().ravel(), oracle.absmax_cols(W))
    sx, sw = torch.empty_like(Cx), torch.empty_like(Cw)
    qg.op_inv_divide(Cx, 127.0, sx)
    qg.op_inv_divide(Cw, 127.0, sw)
    assert same_f32(sx.cpu().numpy().ravel(), oracle.inv_divide(oracle.absmax_rows(X)))
    Xq = torch.empty((50, 640), dtype=torch.int8, device=DEV)
    Wq = torch.empty((640, 48), dtype=torch.int8, device=DEV)
    qg.op_multiply(dX, sx, Xq)
    qg.op_multiply(dW, sw, Wq)
    assert np.array_equal(Xq.cpu().numpy(), oracle.absmax_quant_rows(X)[0])
    assert np.array_equal(Wq.cpu().numpy(), oracle.absmax_quant_cols(W)[0])


def test_outlier_extractor(qg, oracle):
    rng = np.random.default_rng(6)
    A = (rng.standard_normal((37, 300)) * 4).astype(np.float32)
    A[3, 7] = np.nan
    A[4, 8] = 6.0
    A[5, 9] = -6.0
    out = torch.empty((37, 300), device=DEV)
    qg.op_outlier_extractor(to_dev(A), 6.0, out)
    assert np.array_equal(out.cpu().numpy(), oracle.outlier_mask(A, 6.0))


# ------------------------------------------------------------------------------------------
# int8 GEMM (a5): every kernel variant, exact int32
# ------------------------------------------------------------------------------------------
GEMM_SHAPES = [(128, 256, 128), (1, 8, 16), (3, 2, 3), (100, 70, 33), (200, 300, 1000), (256, 512, 4096),
               (129, 257, 129), (512, 1024, 640), (1024, 1024, 1024), (384, 768, 96)]
VARIANTS = ["SIMT", "TC_1SM", "TC_2SM"]


def codes(rng, r, c):
    return rng.integers(-128, 128, (r, c), dtype=np.int8)


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape", GEMM_SHAPES)
def test_gemm_s8s8s32_bit_exact(qg, oracle, shape, variant):
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape))
    A, B = codes(rng, M, K), codes(rng, K, N)
    dA, dB = padded(torch.from_numpy(A).to(DEV), 16), padded(torch.from_numpy(B).to(DEV), 16)
    out = padded(torch.full((M, N), -7, dtype=torch.int32, device=DEV), 4)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.op_mm(dA, dB, out)
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    assert np.array_equal(out.cpu().numpy(), oracle.gemm_s8s8s32(A, B))


@pytest.mark.parametrize("shape", [(256, 2048, 4096), (128, 1024, 8192), (300, 1000, 6144), (64, 512, 16384),
                                   (256, 768, 12288), (4096, 512, 2048)])
def test_gemm_split_k_shapes(qg, oracle, shape):
    """Few tiles, long K: the dispatcher cuts K into 2-4 slices (int32 partial sums, then one reduce + epilogue
    pass).  Integer partial sums add exactly, so every output must still be bit-exact: raw accumulators, the
    dequantized fp32 / fp16 result with bias, and the prepared-weight linear with ReLU."""
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape))
    A, B = codes(rng, M, K), codes(rng, K, N)
    dA, dB = torch.from_numpy(A).to(DEV), torch.from_numpy(B).to(DEV)
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    qg.op_mm(dA, dB, acc)
    exp_acc = oracle.gemm_s8s8s32(A, B)
    assert np.array_equal(acc.cpu().numpy(), exp_acc)
    Cx = rng.random(M).astype(np.float32) + 0.5
    Cw = rng.random(N).astype(np.float32) + 0.5
    bias = rng.standard_normal(N).astype(np.float32)
    exp = oracle.dequant(exp_acc, Cx, Cw, 127.0, bias)
    Bt = torch.from_numpy(np.ascontiguousarray(B.T)).to(DEV)
    for dt in ("f32", "f16"):
        out = torch.empty((M, N), dtype=TORCH_DT[dt], device=DEV)
        qg.gemm_s8t_dequant(dA, Bt, to_dev(Cx), to_dev(Cw), out, 127.0, bias=to_dev(bias))
        if dt == "f32":
            assert same_f32(out.cpu().numpy(), exp)
        else:
            assert torch.equal(out.cpu(), torch.from_numpy(exp).to(torch.float16))


@pytest.mark.parametrize("split", [2, 3, 4])
def test_forced_split_k(split):
    """Every tensor-core product cut into `split` k-slices (QG_SPLIT_K, read once per process -> subprocess)."""
    import subprocess
    import sys as _sys

    env = dict(os.environ, QG_SPLIT_K=str(split))
    r = subprocess.run([_sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitk_check.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_two_pair_clusters():
    """The GEMM with two CTA pairs per cluster (512 x 256 cluster tiles, B quarter-loads TMA-multicast between the pairs;
    QG_GEMM_NP=2, read once per process -> subprocess): same bits as the default kernel and the oracle."""
    import subprocess
    import sys as _sys

    env = dict(os.environ, QG_GEMM_NP="2")
    r = subprocess.run([_sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "splitk_check.py")],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("variant", ["TC_1SM", "TC_2SM"])
def test_gemm_large_sampled_rows(qg, oracle, variant):
    """4096^3 (BASELINE target shape): full result against torch._int_mm is not the bar -- the
    oracle is; it checks 96 sampled rows exactly, and a checksum of every row against int64 math."""
    M = N = K = 4096
    rng = np.random.default_rng(11)
    A, B = codes(rng, M, K), codes(rng, K, N)
    dA, dB = torch.from_numpy(A).to(DEV), torch.from_numpy(B).to(DEV)
    out = torch.empty((M, N), dtype=torch.int32, device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.op_mm(dA, dB, out)
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    rows = np.sort(rng.choice(M, 96, replace=False))
    got = out.cpu().numpy()
    assert np.array_equal(got[rows], oracle.gemm_s8s8s32(A[rows], B))
    # linearity checksum over all rows: sum_j C[i,j] == A[i,:] . (B @ 1)
    bsum = B.astype(np.int64).sum(axis=1)
    assert np.array_equal(got.astype(np.int64).sum(axis=1), A.astype(np.int64) @ bsum)


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_kmajor_b_hook(qg, oracle, cg):
    """Same kernel with B supplied as [N,K] (classic K-major operand) -- cross-checks the MN-major
    shared-memory descriptors used for the reference's [K,N] weight layout."""
    M, N, K = 256, 512, 384
    rng = np.random.default_rng(12)
    A, B = codes(rng, M, K), codes(rng, K, N)
    dA = torch.from_numpy(A).to(DEV)
    dBt = torch.from_numpy(np.ascontiguousarray(B.T)).to(DEV)
    out = torch.empty((M, N), dtype=torch.int32, device=DEV)
    rc = qg.lib().qg_test_gemm_s8_bt(cg, C.c_void_p(dA.data_ptr()), C.c_int64(K), C.c_void_p(dBt.data_ptr()),
                                     C.c_int64(K), M, N, K, C.c_void_p(out.data_ptr()), C.c_int64(N),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, qg.lib().qg_last_error()
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), oracle.gemm_s8s8s32(A, B))


# ------------------------------------------------------------------------------------------
# fused dequantize epilogue (a6-a8, a10) and the unfused form
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape,with_bias", [((128, 256, 128), False), ((200, 300, 520), True),
                                             ((512, 768, 1024), True), ((64, 37, 200), True)])
def test_gemm_dequant_epilogue(qg, oracle, shape, with_bias, dt, variant):
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape, dt))
    A, B = codes(rng, M, K), codes(rng, K, N)
    Cx = (rng.random(M, dtype=np.float32) + 0.05)
    Cw = (rng.random(N, dtype=np.float32) + 0.05)
    Cx[0] = 0.0
    Cw[-1] = -0.0
    bias = rng.standard_normal(N).astype(np.float32) if with_bias else None
    dA, dB = padded(torch.from_numpy(A).to(DEV), 16), padded(torch.from_numpy(B).to(DEV), 16)
    out = torch.empty((M, N), dtype=TORCH_DT[dt], device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.gemm_s8_dequant(dA, dB, to_dev(Cx), to_dev(Cw), out, 127.0, None if bias is None else to_dev(bias))
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    expect = oracle.dequant(oracle.gemm_s8s8s32(A, B), Cx, Cw, 127.0, bias)
    if dt == "f32":
        assert same_f32(out.cpu().numpy(), expect)
    else:
        assert torch.equal(out.cpu(), torch.from_numpy(expect).to(TORCH_DT[dt]))
    # unfused entry point on stored accumulators gives the same bits
    acc = torch.from_numpy(oracle.gemm_s8s8s32(A, B)).to(DEV)
    out2 = torch.empty_like(out)
    qg.op_dequantize(acc, to_dev(Cx), to_dev(Cw), out2, 127.0, None if bias is None else to_dev(bias))
    assert torch.equal(out2.view(torch.int16 if dt != "f32" else torch.int32),
                       out.view(torch.int16 if dt != "f32" else torch.int32))


# ------------------------------------------------------------------------------------------
# the whole op (a9) and LinearLayer::forward
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("shape", [(3, 2, 3), (37, 51, 129), (256, 256, 256), (512, 768, 1024), (2048, 512, 512)])
def test_op_quantized_mm_bit_exact(qg, oracle, shape, variant):
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape))
    X = make_edge_matrix(rng, M, K)
    W = np.ascontiguousarray(make_edge_matrix(rng, N, K).T)
    O = torch.empty((M, N), device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        qg.op_quantized_mm(to_dev(X), to_dev(W), O, 127.0)
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    expect, parts = oracle.quantized_mm(X, W, 127.0, return_parts=True)
    assert oracle.max_partial_sum(parts["Xq"], parts["Wq"]) < 2**24  # reference accumulator == int32 here
    assert same_f32(O.cpu().numpy(), expect)


@pytest.mark.parametrize("size", [4096, 8192])
def test_op_quantized_mm_full_size_properties(qg, oracle, size):
    """BASELINE's headline shape (and 8192^3) at full size, through properties that do not need a full CPU
    product: exact scales, sampled rows bit-exact against the oracle, and the independence structure of the
    op -- rows of X are quantized independently (row permutation commutes with the op, bit for bit),
    columns of W independently (a column block of W gives the same block of O), repeat calls identical."""
    M = N = K = size
    g = torch.Generator(device=DEV).manual_seed(size)
    X = torch.rand((M, K), device=DEV, generator=g) * 2 - 1
    W = torch.rand((K, N), device=DEV, generator=g) * 2 - 1
    O = torch.empty((M, N), device=DEV)
    qg.op_quantized_mm(X, W, O, 127.0)
    Xh, Wh = X.cpu().numpy(), W.cpu().numpy()
    # scales of the whole problem
    Xq = torch.empty((M, K), dtype=torch.int8, device=DEV); Cx = torch.empty(M, device=DEV)
    Wq = torch.empty((K, N), dtype=torch.int8, device=DEV); Cw = torch.empty(N, device=DEV)
    qg.absmax_quant_rows(X, 127.0, qg.MODE_REF_EXACT, Xq, Cx)
    qg.absmax_quant_cols(W, 127.0, qg.MODE_REF_EXACT, Wq, Cw)
    assert same_f32(Cx.cpu().numpy(), oracle.absmax_rows(Xh)) and same_f32(Cw.cpu().numpy(), oracle.absmax_cols(Wh))
    # sampled rows, bit-exact (the oracle quantizes all of W, and only these rows of X)
    rows = np.sort(np.random.default_rng(size).choice(M, 24, replace=False))
    assert same_f32(O[torch.from_numpy(rows).to(DEV)].cpu().numpy(), oracle.quantized_mm(Xh[rows], Wh))
    # row permutation commutes with the op
    perm = torch.randperm(M, device=DEV, generator=g)
    O2 = torch.empty_like(O)
    qg.op_quantized_mm(X[perm].contiguous(), W, O2, 127.0)
    assert torch.equal(O2.view(torch.int32), O[perm].view(torch.int32))
    # a column block of W gives the same block of O (strided W view, narrower N)
    nb = N // 4 + 16
    O3 = torch.empty((M, nb), device=DEV)
    qg.op_quantized_mm(X, W[:, 256:256 + nb], O3, 127.0)
    assert torch.equal(O3.view(torch.int32), O[:, 256:256 + nb].contiguous().view(torch.int32))
    # repeat call: identical bits
    qg.op_quantized_mm(X, W, O2, 127.0)
    assert torch.equal(O2.view(torch.int32), O.view(torch.int32))


def test_op_quantized_mm_uniform_2048_error_stats(qg, oracle):
    """timing_quantize's default shape and distribution: report-level error figures
    (signed-mean ~1e-4, mean-abs ~0.17, max-abs ~1.1 for 2048^3)."""
    M = N = K = 2048
    g = torch.Generator(device="cpu").manual_seed(0)
    X = torch.rand((M, K), generator=g) * 2 - 1
    W = torch.rand((K, N), generator=g) * 2 - 1
    O = torch.empty((M, N), device=DEV)
    qg.op_quantized_mm(X.to(DEV), W.to(DEV), O, 127.0)
    Cref = (X.double() @ W.double())
    err = Cref - O.cpu().double()
    assert abs(err.mean().item()) < 2e-3
    assert 0.10 < err.abs().mean().item() < 0.25
    assert err.abs().max().item() < 2.0
    rows = [0, 1, 777, 2047]
    exp = oracle.quantized_mm(X.numpy(), W.numpy())[rows]
    assert same_f32(O.cpu().numpy()[rows], exp)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_linear_layer_forward(qg, oracle, dt):
    torch.manual_seed(3)
    lin = qg.LinearLayer(640, 384, device=DEV, dtype=TORCH_DT[dt])
    lin.init_uniform()
    x = (torch.rand((200, 640), device=DEV) * 2 - 1).to(TORCH_DT[dt])
    y = torch.empty((200, 384), dtype=TORCH_DT[dt], device=DEV)
    lin.forward(x, y)
    expect = oracle.quantized_mm(as_f32_np(x), as_f32_np(lin.w), 127.0, bias=lin.b.cpu().numpy())
    if dt == "f32":
        assert same_f32(y.cpu().numpy(), expect)
    else:
        assert torch.equal(y.cpu(), torch.from_numpy(expect).to(TORCH_DT[dt]))
    # second call reuses the cached int8 weights and must give the same bits
    y2 = torch.empty_like(y)
    lin.forward(x, y2)
    assert torch.equal(y.view(torch.int16 if dt != "f32" else torch.int32),
                       y2.view(torch.int16 if dt != "f32" else torch.int32))


def test_host_buffer_entry_point(qg, oracle):
    rng = np.random.default_rng(9)
    X = torch.from_numpy(make_edge_matrix(rng, 300, 520)).pin_memory()
    W = torch.from_numpy(np.ascontiguousarray(make_edge_matrix(rng, 260, 520).T)).pin_memory()
    out = qg.quantized_mm_host(X, W)
    assert same_f32(out.numpy(), oracle.quantized_mm(X.numpy(), W.numpy()))


@pytest.mark.parametrize("pinned", [True, False])
def test_host_buffer_entry_point_row_chunk_pipeline(qg, oracle, pinned):
    """M large enough for several row chunks of the copy/compute/copy pipeline, ragged last chunk, bias."""
    rng = np.random.default_rng(19)
    X = torch.from_numpy(make_edge_matrix(rng, 2200, 264))
    W = torch.from_numpy(np.ascontiguousarray(make_edge_matrix(rng, 136, 264).T))
    b = torch.from_numpy(rng.standard_normal(136).astype(np.float32))
    if pinned:
        X, W = X.pin_memory(), W.pin_memory()
    expect = oracle.quantized_mm(X.numpy(), W.numpy(), 127.0, bias=b.numpy())
    for _ in range(2):  # second call reuses the streams, events and device buffers
        out = qg.quantized_mm_host(X, W, bias=b)
        assert same_f32(out.numpy(), expect)


def test_fp32_product_matches_reference_fma_order(qg, oracle):
    rng = np.random.default_rng(10)
    A = rng.standard_normal((70, 200)).astype(np.float32)
    B = rng.standard_normal((200, 90)).astype(np.float32)
    out = torch.empty((70, 90), device=DEV)
    qg.op_mm(to_dev(A), to_dev(B), out)
    assert same_f32(out.cpu().numpy(), oracle.gemm_f32_ref(A, B))
    # transposed view as AttentionLayer passes K.transpose() (attention.cuh:58-60)
    Bt = to_dev(np.ascontiguousarray(B.T)).t()
    out2 = torch.empty((70, 90), device=DEV)
    qg.op_mm(to_dev(A), Bt, out2)
    assert torch.equal(out, out2)


def test_argument_errors_are_reported(qg):
    X = torch.zeros((4, 8), device=DEV)
    W = torch.zeros((9, 4), device=DEV)
    O = torch.zeros((4, 4), device=DEV)
    with pytest.raises(AssertionError):  # the reference asserts X.w == W.h (op_mm.cuh:71)
        qg.op_quantized_mm(X, W, O)
    rc = qg.lib().qg_gemm_s8s8s32(None, C.c_int64(8), None, C.c_int64(8), 4, 4, 8, None, C.c_int64(4), None)
    assert rc == -22 and b"bad arguments" in qg.lib().qg_last_error()


# ------------------------------------------------------------------------------------------
# BASELINE configs 4 / 5 at full size (OPT-6.7B linears at T = 16384, OPT-66B FFN shards at T = 4096, P = 8)
# ------------------------------------------------------------------------------------------
@pytest.mark.timeout(900)
@pytest.mark.parametrize("shape", [(16384, 16384, 4096), (16384, 4096, 16384), (4096, 4608, 9216), (4096, 1152, 36864)],
                         ids=["opt6.7b_fc1", "opt6.7b_fc2", "opt66b_fc1_shard", "opt66b_fc2_shard"])
def test_baseline_config_shapes_full_size(qg, oracle, shape):
    """fp16 in / fp16 out LinearLayer::forward with prepared weights at the shapes BASELINE.json times
    (M, N, K): N(0,1) activations with six outlier feature columns scaled by 20, N(0, 0.02^2) weights.
    Bars: column / row scales exact; sampled rows bit-exact against the oracle (int8 codes, int32 accumulators,
    fp16 output); every partial sum of the sampled rows below 2^24 (the reference's fp32 accumulator is then
    the exact integer); row and column checksums of ALL int32 accumulators against int64 arithmetic."""
    M, N, K = shape
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    X = torch.randn((M, K), device=DEV, generator=g)
    cols = torch.randperm(K, device=DEV, generator=g)[:6]
    X[:, cols] *= 20.0
    X = X.half()
    W = (torch.randn((K, N), device=DEV, generator=g) * 0.02).half()
    bias = torch.randn(N, device=DEV, generator=g)
    Wt, Cw = qg.prepare_weights(W, 127.0, qg.MODE_REF_EXACT)
    Y = torch.empty((M, N), dtype=torch.float16, device=DEV)
    lin = qg.LinearLayer(K, N, device=DEV, dtype=torch.float16)
    lin.w, lin.b = W, bias.reshape(1, N)
    lin.quantize_weights()
    lin.forward(X, Y)
    torch.cuda.synchronize()
    Xh, Wh = X.float().cpu().numpy(), W.float().cpu().numpy()
    # scales and codes of the whole problem
    Xq = torch.empty((M, K), dtype=torch.int8, device=DEV); Cx = torch.empty(M, device=DEV)
    qg.absmax_quant_rows(X, 127.0, qg.MODE_REF_EXACT, Xq, Cx)
    Wq_e, Cw_e = oracle.absmax_quant_cols(Wh)
    assert same_f32(Cw.cpu().numpy(), Cw_e)
    assert np.array_equal(Wt.cpu().numpy()[:, :K], Wq_e.T)
    assert same_f32(Cx.cpu().numpy(), oracle.absmax_rows(Xh))
    # sampled rows: codes, accumulators (exact int32 == the reference's fp32 chain), fp16 output
    rows = np.sort(np.random.default_rng(M ^ N).choice(M, 12, replace=False))
    Xq_e, Cx_e = oracle.absmax_quant_rows(Xh[rows])
    assert np.array_equal(Xq.cpu().numpy()[rows], Xq_e)
    assert oracle.max_partial_sum(Xq_e, Wq_e) < 2 ** 24
    acc_e = oracle.gemm_s8s8s32(Xq_e, Wq_e)
    y_e = oracle.dequant(acc_e, Cx_e, Cw_e, 127.0, bias.cpu().numpy())
    assert torch.equal(Y[torch.from_numpy(rows).to(DEV)].cpu(), torch.from_numpy(y_e).to(torch.float16))
    # all accumulators: row and column checksums in int64
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    qg.gemm_s8t_dequant(Xq, Wt, None, None, acc, 127.0)
    torch.cuda.synchronize()
    assert np.array_equal(acc.cpu().numpy()[rows], acc_e)
    A64, B64 = Xq.cpu().numpy().astype(np.int64), Wq_e.astype(np.int64)
    assert np.array_equal(acc.sum(dim=1, dtype=torch.int64).cpu().numpy(), A64 @ B64.sum(axis=1))
    assert np.array_equal(acc.sum(dim=0, dtype=torch.int64).cpu().numpy(), A64.sum(axis=0) @ B64)


# ------------------------------------------------------------------------------------------
# robustness: host path on tile-starved chunks, two streams, two devices in one process
# ------------------------------------------------------------------------------------------
@pytest.mark.timeout(300)
@pytest.mark.parametrize("shape", [(16, 4096, 4096), (64, 4096, 4096), (1, 4096, 4096), (256, 1024, 8192), (700, 512, 4352)])
def test_host_buffer_entry_point_split_k_shapes(qg, oracle, shape):
    """Host-buffer call whose per-chunk product is tile-starved enough for split-K (M <= 64 against K >= 4096 ...):
    this used to self-deadlock on the library mutex (the split-K scratch grew under a lock the host call already held)."""
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape))
    X = torch.from_numpy(rng.random((M, K), dtype=np.float32) * 2 - 1).pin_memory()
    W = torch.from_numpy(rng.random((K, N), dtype=np.float32) * 2 - 1).pin_memory()
    b = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    out = qg.quantized_mm_host(X, W, bias=b)
    assert same_f32(out.numpy(), oracle.quantized_mm(X.numpy(), W.numpy(), 127.0, bias=b.numpy()))


def test_split_k_slices_live_in_the_callers_workspace(qg, oracle):
    """Two streams running the same decode-shaped linear (split-K) with their OWN workspaces must not share
    scratch: results equal the oracle on both, many times over."""
    M, N, K = 128, 2048, 16384
    rng = np.random.default_rng(5)
    Wh = (rng.random((K, N), dtype=np.float32) * 2 - 1)
    Wt, Cw = qg.prepare_weights(to_dev(Wh), 127.0, qg.MODE_REF_EXACT)
    assert qg.workspace_bytes(M, N, K) > M * K + 4 * (M + N) + 2 * 4 * M * N  # the slices are part of the block
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    Xs = [rng.random((M, K), dtype=np.float32) * 2 - 1 for _ in range(2)]
    exp = [oracle.quantized_mm(x, Wh) for x in Xs]
    dX = [to_dev(x) for x in Xs]
    Ys = [torch.empty((M, N), device=DEV) for _ in range(2)]
    wss = [torch.empty(qg.workspace_bytes(M, N, K), dtype=torch.uint8, device=DEV) for _ in range(2)]
    torch.cuda.synchronize()
    for _ in range(20):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                qg.linear_forward(dX[i], Wt, Cw, None, Ys[i], workspace=wss[i])
    torch.cuda.synchronize()
    for i in range(2):
        assert same_f32(Ys[i].cpu().numpy(), exp[i])


def test_weight_quantization_on_two_streams(qg, oracle):
    """The column-maximum scratch is per (device, stream): concurrent per-call weight quantizations on two
    streams give the same bits as sequential ones."""
    rng = np.random.default_rng(6)
    Ws = [make_edge_matrix(rng, 1500, 2048).T.copy() for _ in range(2)]  # [2048, 1500]
    exp = [oracle.absmax_quant_cols(w) for w in Ws]
    dW = [to_dev(w) for w in Ws]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [(torch.empty(w.shape, dtype=torch.int8, device=DEV), torch.empty(w.shape[1], device=DEV)) for w in Ws]
    torch.cuda.synchronize()
    for _ in range(10):
        for i, st in enumerate(streams):
            with torch.cuda.stream(st):
                qg.absmax_quant_cols(dW[i], 127.0, qg.MODE_REF_EXACT, outs[i][0], outs[i][1])
    torch.cuda.synchronize()
    for i in range(2):
        assert np.array_equal(outs[i][0].cpu().numpy(), exp[i][0]) and same_f32(outs[i][1].cpu().numpy(), exp[i][1])


def test_second_device_in_one_process(qg, oracle):
    """cudaFuncSetAttribute (the > 48 KB shared-memory opt-in) is per device: the tcgen05 GEMM, the softmax and
    the ADD & NORM kernels must launch on device 1 after device 0 has used them, in the same process."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs in one process")
    rng = np.random.default_rng(8)
    X = rng.random((300, 640), dtype=np.float32) * 2 - 1
    W = rng.random((640, 512), dtype=np.float32) * 2 - 1
    exp = oracle.quantized_mm(X, W)
    S = rng.standard_normal((64, 600)).astype(np.float32)
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            dev = f"cuda:{d}"
            O = torch.empty((300, 512), device=dev)
            qg.op_quantized_mm(torch.from_numpy(X).to(dev), torch.from_numpy(W).to(dev), O, 127.0)
            P = torch.empty((64, 600), device=dev)
            qg.op_softmax(torch.from_numpy(S).to(dev), P)
            torch.cuda.synchronize()
            assert same_f32(O.cpu().numpy(), exp)
            np.testing.assert_allclose(P.cpu().numpy(), oracle.softmax_rows(S), rtol=2e-6)
