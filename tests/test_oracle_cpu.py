"""CPU tests: the oracle against the reference's only known-answer input and against the
arithmetic facts SURVEY.md section 8(c) / Appendix A pin (PTX-confirmed semantics)."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    with open(os.path.join(HERE, "golden", "test_quantize_3x3.json")) as f:
        return json.load(f)


def test_golden_3x3(oracle):
    g = golden()
    X = np.array(g["X"], np.float32)
    W = np.array(g["W"], np.float32)
    O, p = oracle.quantized_mm(X, W, g["range"], return_parts=True)
    assert np.array_equal(p["Cx"], np.array(g["Cx"], np.float32))
    assert np.array_equal(p["Cw"], np.array(g["Cw"], np.float32))
    assert np.array_equal(oracle.inv_divide(p["Cx"]), np.array(g["sx"], np.float32))
    assert np.array_equal(oracle.inv_divide(p["Cw"]), np.array(g["sw"], np.float32))
    assert np.array_equal(p["Xq"], np.array(g["Xq"], np.int8))
    assert np.array_equal(p["Wq"], np.array(g["Wq"], np.int8))
    assert np.array_equal(p["acc"], np.array(g["acc"], np.int32))
    np.testing.assert_allclose(O, np.array(g["out"], np.float32), rtol=0, atol=5e-7)
    C = oracle.gemm_f32_ref(X, W)
    assert np.array_equal(C, np.array(g["fp32_product"], np.float32))
    assert abs(oracle.signed_mean(C - O) - g["signed_mean_error"]) < 1e-8


def test_absmax_first_element_quirk(oracle):
    # op_reduction.cuh:80-83: accumulator starts from the SIGNED first element
    X = np.array([[-3.0, 1.0, 2.0], [3.0, -1.0, -2.0], [-0.5, 0.0, 0.0], [-0.5, -0.0, 0.0]], np.float32)
    c = oracle.absmax_rows(X)
    assert c[0] == 2.0 and c[1] == 3.0
    assert c[2] == 0.0 and np.signbit(c[2])       # -(+0.0): first later zero decides the sign
    assert c[3] == 0.0 and not np.signbit(c[3])   # -(-0.0)
    ct = oracle.absmax_rows(X, oracle.MODE_TRUE_ABSMAX)
    assert ct[0] == 3.0 and ct[2] == 0.5
    # columns behave the same way (op_reduction.cuh:105-108)
    assert np.array_equal(oracle.absmax_cols(X.T.copy()), c)


def test_quantizing_cast_truncates_and_wraps(oracle):
    # op_elemwise.cuh:106-114 -> mul.f32; cvt.rzi.s32.f32; st.u8
    assert oracle.quant_code(1.0, 126.99) == 126          # truncation, not rounding
    assert oracle.quant_code(-1.0, 126.99) == -126
    assert oracle.quant_code(-3.0, 63.5) == 66            # -190 wraps to 66 (low byte)
    assert oracle.quant_code(3.0, 63.5) == -66
    assert oracle.quant_code(float("nan"), 1.0) == 0      # cvt of NaN is 0
    assert oracle.quant_code(0.0, float("inf")) == 0      # 0*inf = NaN
    assert oracle.quant_code(1.0, float("inf")) == -1     # +inf clamps to INT_MAX, low byte 0xff
    assert oracle.quant_code(-1.0, float("inf")) == 0     # -inf clamps to INT_MIN, low byte 0x00


def test_wrapped_row_matches_sequential_definition(oracle):
    rng = np.random.default_rng(0)
    X = (rng.random((4, 64), dtype=np.float32) * 2 - 1)
    X[1, 0] = -5.0
    Xq, Cx = oracle.absmax_quant_rows(X)
    s = np.float32(127.0) / Cx[1]
    expect = np.trunc(X[1].astype(np.float32) * s).astype(np.int64)
    assert np.array_equal(Xq[1].astype(np.int64), ((expect + 128) % 256) - 128)
    assert abs(int(expect[0])) > 127 and int(Xq[1, 0]) != int(expect[0])  # the wrapped element


def test_f32_accumulator_equals_int32_below_2_24(oracle):
    rng = np.random.default_rng(1)
    A = rng.integers(-127, 128, (33, 700), dtype=np.int8)
    B = rng.integers(-127, 128, (700, 45), dtype=np.int8)
    assert oracle.max_partial_sum(A, B) < 2 ** 24
    exact = oracle.gemm_s8s8s32(A, B)
    assert np.array_equal(exact, A.astype(np.int64) @ B.astype(np.int64))
    assert np.array_equal(oracle.gemm_s8_reff32(A, B), exact)


def test_f32_accumulator_diverges_above_2_24(oracle):
    # documents the limit of "reference accumulator == int32": K * 127^2 must stay below 2^24
    A = np.full((1, 2048), 127, np.int8)
    B = np.full((2048, 1), 127, np.int8)
    assert oracle.max_partial_sum(A, B) >= 2 ** 24
    assert oracle.gemm_s8s8s32(A, B)[0, 0] == 2048 * 127 * 127
    assert oracle.gemm_s8_reff32(A, B)[0, 0] != 2048 * 127 * 127


def test_dequant_rounding_order(oracle):
    rng = np.random.default_rng(2)
    acc = rng.integers(-2_000_000, 2_000_000, (17, 19), dtype=np.int32)
    Cx = rng.random(17, dtype=np.float32) + 0.1
    Cw = rng.random(19, dtype=np.float32) + 0.1
    bias = rng.random(19, dtype=np.float32)
    c = np.float32(1.0) / (np.float32(127.0) * np.float32(127.0))
    outer = (Cx[:, None] * Cw[None, :]).astype(np.float32)
    expect = ((acc.astype(np.float32) * outer).astype(np.float32) * c).astype(np.float32)
    assert np.array_equal(oracle.dequant(acc, Cx, Cw), expect)
    assert np.array_equal(oracle.dequant(acc, Cx, Cw, bias=bias), (expect + bias[None, :]).astype(np.float32))
    # not the same as multiplying by the reciprocal scales (last paragraph)
    alt = acc.astype(np.float32) * (Cx[:, None] / np.float32(127)) * (Cw[None, :] / np.float32(127))
    assert not np.array_equal(alt.astype(np.float32), expect)


def test_outer_product_negative_zero(oracle):
    acc = np.array([[5]], np.int32)
    out = oracle.dequant(acc, np.array([-0.0], np.float32), np.array([1.0], np.float32))
    assert out[0, 0] == 0.0 and not np.signbit(out[0, 0])  # fma(-0,1,+0) = +0 in the K=1 tiled product


def test_full_pipeline_consistency(oracle):
    rng = np.random.default_rng(3)
    X = rng.random((37, 129), dtype=np.float32) * 2 - 1
    W = rng.random((129, 51), dtype=np.float32) * 2 - 1
    O, p = oracle.quantized_mm(X, W, return_parts=True)
    Xq, Cx = oracle.absmax_quant_rows(X)
    Wq, Cw = oracle.absmax_quant_cols(W)
    assert np.array_equal(Xq, p["Xq"]) and np.array_equal(Wq, p["Wq"])
    assert np.array_equal(oracle.dequant(oracle.gemm_s8s8s32(Xq, Wq), Cx, Cw), O)
    C = X.astype(np.float64) @ W.astype(np.float64)
    assert np.abs(O - C).mean() < 0.02 * np.sqrt((C ** 2).mean()) + 1e-3


def test_outlier_mask(oracle):
    A = np.array([[6.0, -6.0, 6.0001, -7.0, np.nan, 0.0]], np.float32)
    assert oracle.outlier_mask(A, 6.0).tolist() == [[0.0, 0.0, 1.0, 1.0, 1.0, 0.0]]
    assert oracle.outlier_columns(A, 6.0).tolist() == [2, 3, 4]


def _same_f32(a, b):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    nan = np.isnan(a)
    return a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and np.array_equal(
        a.view(np.int32)[~nan], b.view(np.int32)[~nan])


@pytest.mark.parametrize("name", ["3x3", "edge_24x40x56", "rand_64x48x96", "normal_33x65x130", "curand_seed0_32x16x64"])
def test_oracle_matches_reference_kernel_outputs(oracle, name):
    """tests/golden/ref_*.npz were produced by the REFERENCE's own CUDA kernels on a B200
    (tests/golden/make_ref_fixtures.py through oracle/_ref): every intermediate must match bit for bit
    (signed zeros included, NaNs position-wise)."""
    r = np.load(os.path.join(HERE, "golden", f"ref_{name}.npz"))
    O, p = oracle.quantized_mm(r["X"], r["W"], 127.0, return_parts=True)
    assert _same_f32(p["Cx"], r["Cx"]) and _same_f32(p["Cw"], r["Cw"])
    assert np.array_equal(p["Xq"], r["Xq"]) and np.array_equal(p["Wq"], r["Wq"])
    assert np.array_equal(p["acc"], r["acc"])
    assert np.array_equal(oracle.gemm_s8_reff32(p["Xq"], p["Wq"]), r["acc"])
    assert _same_f32(O, r["O"])          # step-by-step pipeline (timing_quantize.cu:38-58)
    assert _same_f32(O, r["O_op"])       # the reference's op_quantized_mm itself (op_mm.cuh:67-101)
    assert _same_f32(oracle.gemm_f32_ref(r["X"], r["W"]), r["C_fp32"])


# ---- attention widening: oracle against the reference kernels' outputs ----
import glob as _glob  # noqa: E402

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(_glob.glob(os.path.join(_GOLDEN, "ref_softmax_*.npz"))))
def test_oracle_softmax_vs_reference_kernel(oracle, path):
    g = np.load(path)
    # libm expf vs the device's expf (<= 2 ulp): tolerance, not bits
    np.testing.assert_allclose(oracle.softmax_rows(g["A"]), g["B"], rtol=2e-6, atol=1e-37)


@pytest.mark.parametrize("path", sorted(_glob.glob(os.path.join(_GOLDEN, "ref_attn_*.npz"))))
def test_oracle_attention_vs_reference_kernels(oracle, path):
    g = np.load(path)
    o = oracle.attention_forward(g["Xq"], g["Xkv"], g["Wq"], g["Wk"], g["Wv"], return_parts=True)
    for k in ("Q", "K", "V", "S"):  # quantized projections and the fp32 score product: bit-exact
        assert np.array_equal(o[k].view(np.uint32), g[k].view(np.uint32)), k
    np.testing.assert_allclose(o["P"], g["P"], rtol=2e-6, atol=1e-37)
    np.testing.assert_allclose(o["out"], g["out"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("path", sorted(_glob.glob(os.path.join(_GOLDEN, "ref_addnorm_*.npz"))))
def test_oracle_add_layernorm_vs_reference_kernel(oracle, path):
    g = np.load(path)
    got = oracle.add_layernorm(g["A"], g["R"] if "R" in g.files else None)
    assert np.array_equal(got.view(np.uint32), g["B"].view(np.uint32))


@pytest.mark.parametrize("path", sorted(_glob.glob(os.path.join(_GOLDEN, "ref_enc_*.npz"))))
def test_oracle_encoder_block_vs_reference_kernels(oracle, path):
    g = np.load(path)
    W = np.concatenate([np.concatenate(list(g[k]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1)
    o = oracle.encoder_block(g["X"], W, g["W_O"], g["W1"], g["b1"], g["W2"], g["b2"], int(g["heads"]))
    err = np.abs(o - g["out"])
    # libm vs device expf differ in the last bits and a 1-ulp change can move an int8 code downstream
    assert np.median(err) <= 1e-4 * max(1.0, np.abs(g["out"]).max())


# ---- the timed CPU baseline (oracle/qfast.c) must be the same function as the oracle ----------------------

def _edge_inputs(rng, M, K, N):
    X = (rng.random((M, K), dtype=np.float32) * 2 - 1)
    W = (rng.random((K, N), dtype=np.float32) * 2 - 1)
    X[0, :] = 0.0                                  # all-zero row: 127/0 = inf, 0*inf = NaN -> code 0
    if M > 1:
        X[1, 0], X[1, 1:] = -5.0, 0.0              # negative first element, zeros after: the maximum is a zero
    if M > 2:
        X[2, 0], X[2, 1:] = -3.0, -0.0             # ... whose SIGN depends on the visiting order
        X[2, min(5, K - 1)] = np.nan
    if M > 3:
        X[3, 0] = -2.5                             # largest magnitude negative and first: wrapped codes
    W[:, 0] = 0.0
    if N > 1:
        W[0, 1], W[1:, 1] = -4.0, 0.0
    if N > 2:
        W[min(3, K - 1), 2] = np.inf
    return X, W


@pytest.mark.parametrize("shape", [(3, 3, 2), (7, 70, 5), (70, 257, 130), (130, 96, 200), (260, 1000, 77)])
@pytest.mark.parametrize("mode", [0, 1])
def test_fast_cpu_baseline_equals_oracle(oracle, shape, mode):
    M, K, N = shape
    rng = np.random.default_rng(M * 1000 + N)
    X, W = _edge_inputs(rng, M, K, N)
    b = rng.random(N, dtype=np.float32)
    oracle.fast_set_threads(4)
    O1, p1 = oracle.quantized_mm(X, W, 127.0, mode, bias=b, return_parts=True)
    O2, p2 = oracle.fast_quantized_mm(X, W, 127.0, mode, bias=b, return_parts=True)
    for key in ("Cx", "Cw"):
        assert np.array_equal(p1[key].view(np.uint32), p2[key].view(np.uint32)), key
    for key in ("Xq", "Wq", "acc"):
        assert np.array_equal(p1[key], p2[key]), key
    assert np.array_equal(O1.view(np.uint32), O2.view(np.uint32))


def test_fast_cpu_gemm_extreme_codes(oracle):
    """+-127 / -128 codes and K not a multiple of 4: the +128 bias trick of the VNNI kernel stays exact."""
    rng = np.random.default_rng(9)
    A = rng.integers(-128, 128, (37, 1030), dtype=np.int8)
    B = rng.integers(-128, 128, (1030, 75), dtype=np.int8)
    A[0, :], B[:, 0] = -128, -128
    A[1, :], B[:, 1] = 127, -128
    assert np.array_equal(oracle.fast_gemm_s8s8s32(A, B), oracle.gemm_s8s8s32(A, B))
    assert np.array_equal(oracle.fast_gemm_s8s8s32(A, B), A.astype(np.int64) @ B.astype(np.int64))


@pytest.mark.parametrize("path", sorted(_glob.glob(os.path.join(_GOLDEN, "ref_dec_*.npz"))))
def test_oracle_decoder_block_vs_reference_kernels(oracle, path):
    """The decoder composite of the reference's kernels (oracle/ref_driver.cu: ref_decoder_block, run on a B200)
    against the CPU restatement.  libm's expf is not the device's, and a last-bit difference in a softmax can move
    an int8 code downstream, so the bound is statistical (as for the encoder); the bit-exact check of the CUDA path
    against this fixture is tests/test_gpu_transformer.py."""
    g = np.load(path)
    cat = lambda pre: np.ascontiguousarray(np.concatenate(
        [np.concatenate(list(g[f"{pre}_{k}"]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1))
    o = oracle.decoder_block(g["X"], g["E"], cat("sa"), g["sa_W_O"], cat("ca"), g["ca_W_O"], g["W1"], g["b1"], g["W2"], g["b2"],
                             int(g["heads"]))
    assert o.shape == g["out"].shape
    err = np.abs(o - g["out"])
    assert np.median(err) <= 1e-4 * max(1.0, np.abs(g["out"]).max())
