"""GPU parity of the encoder / decoder widening: ReLU in the GEMM epilogue,
ADD & NORM, and one Encoder-loop iteration against the REFERENCE's kernels composed statement by statement
(oracle/ref_driver.cu: ref_encoder_block, ref_add_layernorm) -- bit-exact, live and as committed fixtures."""
import ctypes as C
import glob
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from conftest import make_edge_matrix
from test_gpu_parity import same_f32, to_dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so")
sys.path.insert(0, GOLDEN)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_qmm.so not built (needs /root/reference at build time)")
    return C.CDLL(REF_SO)


@pytest.fixture(scope="module")
def tf(qg):
    return importlib.import_module(qg.__name__ + ".transformer")


def load_block(tf, g):
    """EncoderBlock with the fixture's weights ([heads, d_model, d] stacks -> the fused column layout)."""
    d_model, heads, d_ff = g["X"].shape[1], int(g["heads"]), int(g["d_ff"])
    blk = tf.EncoderBlock(d_model, heads, d_ff)
    W = np.concatenate([np.concatenate(list(g[k]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1)
    blk.attn.W_qkv.copy_(to_dev(np.ascontiguousarray(W)))
    blk.W_O.w.copy_(to_dev(g["W_O"]))
    blk.ll1.w.copy_(to_dev(g["W1"])); blk.ll1.b.copy_(to_dev(g["b1"]))
    blk.ll2.w.copy_(to_dev(g["W2"])); blk.ll2.b.copy_(to_dev(g["b2"]))
    return blk


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_addnorm_*.npz"))))
def test_add_layernorm_matches_reference_fixture_bit_for_bit(tf, oracle, path):
    g = np.load(path)
    A = to_dev(g["A"])
    R = to_dev(g["R"]) if "R" in g.files else None
    B = torch.empty_like(A)
    tf.add_layernorm(A, R, B)
    assert same_f32(B.cpu().numpy(), g["B"])
    assert same_f32(oracle.add_layernorm(g["A"], g["R"] if "R" in g.files else None), g["B"])
    tf.add_layernorm(A, R, A)  # in place, as transformer.cu:57-58 runs it
    assert same_f32(A.cpu().numpy(), g["B"])


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (130, 33), (300, 40), (1000, 512)])
def test_add_layernorm_vs_oracle_and_live_reference(tf, oracle, ref, shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    A = rng.standard_normal(shape).astype(np.float32)
    R = rng.standard_normal(shape).astype(np.float32)
    B = torch.empty(shape, device="cuda")
    tf.add_layernorm(to_dev(A), to_dev(R), B)
    got = B.cpu().numpy()
    assert same_f32(got, oracle.add_layernorm(A, R))
    Bref = np.zeros_like(A)
    assert ref.ref_add_layernorm(A.ctypes.data_as(C.c_void_p), R.ctypes.data_as(C.c_void_p), shape[0], shape[1],
                                 Bref.ctypes.data_as(C.c_void_p)) == 0
    covered = min(shape[0], 256 * -(-shape[1] // 256))  # rows the reference's grid reaches (op_layernorm.cuh:41)
    assert same_f32(got[:covered], Bref[:covered])


def test_add_layernorm_edge_rows_vs_oracle(tf, oracle):
    """Rows that leave the fp32 3-sum chain (tiny deviations, overflowing squares, inf, NaN), rows of very
    different magnitudes, and a wide row (rows-per-CTA shrinks) -- all bit-exact against the oracle's
    literal double-add / float-store restatement."""
    rng = np.random.default_rng(77)
    N = 300
    A = rng.standard_normal((40, N)).astype(np.float32)
    A[1] *= np.float32(1e-20); A[2] *= np.float32(1e18); A[3] *= np.float32(3e19)   # tiny / large / overflowing squares
    A[4] = np.float32(1.0); A[4, 7] = np.float32(1.0 + 2.0 ** -23)                   # deviations near 1e-7
    A[5] = np.float32(0.0)                                                           # var = 0 -> 0/0
    A[6, 5] = np.inf; A[7, 9] = np.nan; A[8] = np.float32(1e-30)
    A[9] = (rng.standard_normal(N) * 1e-12).astype(np.float32)
    for scale_row in range(10, 40):
        A[scale_row] *= np.float32(2.0 ** rng.integers(-30, 30))
    B = torch.empty((40, N), device="cuda")
    tf.add_layernorm(to_dev(A), None, B)
    assert same_f32(B.cpu().numpy(), oracle.add_layernorm(A))
    wide = rng.standard_normal((9, 3000)).astype(np.float32)
    Bw = torch.empty((9, 3000), device="cuda")
    tf.add_layernorm(to_dev(wide), to_dev(wide[::-1].copy()), Bw)
    assert same_f32(Bw.cpu().numpy(), oracle.add_layernorm(wide, wide[::-1].copy()))
    big = (rng.standard_normal((2000, 512)) * 3).astype(np.float32)
    Bb = torch.empty((2000, 512), device="cuda")
    tf.add_layernorm(to_dev(big), None, Bb)
    assert same_f32(Bb.cpu().numpy(), oracle.add_layernorm(big))


@pytest.mark.parametrize("shape", [(5, 7, 9), (200, 136, 264), (512, 768, 1024)])
def test_relu_epilogue(qg, tf, oracle, shape):
    M, N, K = shape
    rng = np.random.default_rng(M)
    X = make_edge_matrix(rng, M, K)
    lin = tf.PreparedLinear(K, N)
    lin.w.copy_(to_dev(np.ascontiguousarray(make_edge_matrix(rng, N, K).T)))
    lin.b.copy_(to_dev(rng.standard_normal((1, N)).astype(np.float32)))
    y = torch.empty((M, N), device="cuda")
    lin.forward(to_dev(X), y, tf.ACT_RELU)
    exp = oracle.relu(oracle.quantized_mm(X, lin.w.cpu().numpy(), 127.0, bias=lin.b.cpu().numpy()))
    assert same_f32(y.cpu().numpy(), exp)
    y0 = torch.empty((M, N), device="cuda")
    lin.forward(to_dev(X), y0)  # no activation: unchanged path
    assert same_f32(y0.cpu().numpy(), oracle.quantized_mm(X, lin.w.cpu().numpy(), 127.0, bias=lin.b.cpu().numpy()))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_enc_*.npz"))))
def test_encoder_block_matches_reference_fixture_bit_for_bit(tf, path):
    g = np.load(path)
    blk = load_block(tf, g)
    out = torch.empty(g["X"].shape, device="cuda")
    blk.forward(to_dev(g["X"]), out)
    assert same_f32(out.cpu().numpy(), g["out"])


@pytest.mark.parametrize("h,d_model,heads,d_ff", [(6, 8, 4, 8), (40, 32, 2, 64), (128, 128, 8, 256)])
def test_encoder_block_matches_live_reference_and_batches(tf, ref, oracle, h, d_model, heads, d_ff):
    from make_ref_fixtures import encoder_weights, run_ref_encoder_block

    rng = np.random.default_rng(h * 31 + d_model)
    w = encoder_weights(rng, d_model, heads, d_ff)
    X = (rng.random((h, d_model), dtype=np.float32) * 2 - 1)
    expect = run_ref_encoder_block(ref, X, heads, d_ff, **w)
    blk = load_block(tf, dict(X=X, heads=heads, d_ff=d_ff, **w))
    out = torch.empty((h, d_model), device="cuda")
    blk.forward(to_dev(X), out)
    assert same_f32(out.cpu().numpy(), expect)
    # two sequences per call == two calls (row scales are per row, attention per sequence)
    X2 = (rng.random((h, d_model), dtype=np.float32) * 2 - 1)
    both = torch.empty((2 * h, d_model), device="cuda")
    blk.forward(to_dev(np.concatenate([X, X2])), both, batch=2)
    assert same_f32(both[:h].cpu().numpy(), expect)
    assert same_f32(both[h:].cpu().numpy(), run_ref_encoder_block(ref, X2, heads, d_ff, **w))
    if h <= 40:  # CPU oracle: libm expf != device expf, and a 1-ulp change can move an int8 code -> loose bound
        W = np.concatenate([np.concatenate(list(w[k]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1)
        o = oracle.encoder_block(X, W, w["W_O"], w["W1"], w["b1"], w["W2"], w["b2"], heads)
        err = np.abs(o - expect)
        assert np.median(err) <= 1e-4 * max(1.0, np.abs(expect).max())


def load_decoder_block(tf, X, heads, d_ff, sa, ca, W1, b1, W2, b2):
    """DecoderBlock with the given weights ([heads, d_model, d] stacks -> the fused column layout)."""
    blk = tf.DecoderBlock(X.shape[1], int(heads), int(d_ff))
    for attn, wo, w in ((blk.self_attn, blk.W_O1, sa), (blk.cross_attn, blk.W_O2, ca)):
        W = np.concatenate([np.concatenate(list(w[k]), axis=1) for k in ("Wq", "Wk", "Wv")], axis=1)
        attn.W_qkv.copy_(to_dev(np.ascontiguousarray(W)))
        wo.w.copy_(to_dev(w["W_O"]))
    blk.ll1.w.copy_(to_dev(W1)); blk.ll1.b.copy_(to_dev(b1))
    blk.ll2.w.copy_(to_dev(W2)); blk.ll2.b.copy_(to_dev(b2))
    return blk


def _unflatten_decoder_fixture(g):
    sa = {k: g[f"sa_{k}"] for k in ("Wq", "Wk", "Wv", "W_O")}
    ca = {k: g[f"ca_{k}"] for k in ("Wq", "Wk", "Wv", "W_O")}
    return dict(sa=sa, ca=ca, W1=g["W1"], b1=g["b1"], W2=g["W2"], b2=g["b2"])


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_dec_*.npz"))))
def test_decoder_block_matches_reference_fixture_bit_for_bit(tf, path):
    """One Decoder-loop iteration (transformer.cu:91-166) composed from the REFERENCE's kernels on a B200
    (oracle/ref_driver.cu: ref_decoder_block), committed as a fixture."""
    g = np.load(path)
    blk = load_decoder_block(tf, g["X"], g["heads"], g["d_ff"], **_unflatten_decoder_fixture(g))
    out = torch.empty(g["X"].shape, device="cuda")
    blk.forward(to_dev(g["X"]), to_dev(g["E"]), out)
    assert same_f32(out.cpu().numpy(), g["out"])


@pytest.mark.parametrize("h,h_enc,d_model,heads,d_ff", [(6, 6, 8, 4, 8), (24, 40, 64, 4, 96), (128, 96, 128, 8, 256)])
def test_decoder_block_matches_live_reference_and_batches(tf, ref, h, h_enc, d_model, heads, d_ff):
    """The same composite run live beside the new path: bit-exact; a batched call equals per-sequence reference
    runs; an in-place call (later loop iterations, transformer.cu:104-106) equals the out-of-place one."""
    from make_ref_fixtures import decoder_weights, run_ref_decoder_block

    rng = np.random.default_rng(h * 17 + d_model)
    w = decoder_weights(rng, d_model, heads, d_ff)
    X = [(rng.random((h, d_model), dtype=np.float32) * 2 - 1) for _ in range(2)]
    E = [(rng.random((h_enc, d_model), dtype=np.float32) * 2 - 1) for _ in range(2)]
    expect = [run_ref_decoder_block(ref, X[b], E[b], heads, d_ff, **w) for b in range(2)]
    blk = load_decoder_block(tf, X[0], heads, d_ff, **w)
    out = torch.empty((h, d_model), device="cuda")
    blk.forward(to_dev(X[0]), to_dev(E[0]), out)
    assert same_f32(out.cpu().numpy(), expect[0])
    both = torch.empty((2 * h, d_model), device="cuda")
    blk.forward(to_dev(np.concatenate(X)), to_dev(np.concatenate(E)), both, batch=2)
    assert same_f32(both[:h].cpu().numpy(), expect[0]) and same_f32(both[h:].cpu().numpy(), expect[1])
    xin = to_dev(np.concatenate(X))
    blk.forward(xin, to_dev(np.concatenate(E)), xin, batch=2)
    assert torch.equal(xin.view(torch.int32), both.view(torch.int32))


# ---- quantization carried across layers ---------------------------------------------------
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(5, 9, 7), (200, 264, 136), (384, 512, 640), (1024, 2048, 512), (300, 1000, 1032)])
def test_epilogue_row_maxima_feed_the_next_quantizer(qg, oracle, shape, dt):
    """GEMM epilogue -> row maxima -> qg_quantize_rows_given_max: codes and Cx must equal the plain row quantizer
    run on the stored output (REF_EXACT: signed first element, NaN / zero rows included), which equals the oracle."""
    from test_gpu_parity import TORCH_DT, as_f32_np

    M, N, K = shape
    rng = np.random.default_rng(M + N)
    X = make_edge_matrix(rng, M, K)
    W = np.ascontiguousarray(make_edge_matrix(rng, N, K).T)
    b = rng.standard_normal(N).astype(np.float32)
    Wt, Cw = qg.prepare_weights(to_dev(W), 127.0, qg.MODE_REF_EXACT)
    Xq, Cx = qg.absmax_quant_rows(to_dev(X))
    for act in (0, 1):
        Y = torch.empty((M, N), dtype=TORCH_DT[dt], device="cuda")
        rm = torch.full((M,), 123.0, device="cuda")
        qg.linear_forward_q(Xq, Cx, Wt, Cw, to_dev(b), Y, act=act, y_rowmax=rm)
        Yq, Cy = qg.quantize_rows_given_max(Y, rm)
        Yq2, Cy2 = qg.absmax_quant_rows(Y)
        assert torch.equal(Yq, Yq2) and same_f32(Cy.cpu().numpy(), Cy2.cpu().numpy())
        exp_y = oracle.quantized_mm(X, W, 127.0, bias=b)
        if act:
            exp_y = oracle.relu(exp_y)
        exp_y = as_f32_np(torch.from_numpy(exp_y).to(TORCH_DT[dt]))
        eq, ec = oracle.absmax_quant_rows(exp_y)
        assert np.array_equal(Yq.cpu().numpy(), eq) and same_f32(Cy.cpu().numpy(), ec)


@pytest.mark.parametrize("shape", [(6, 8, 8, 8), (200, 136, 264, 72), (512, 512, 2048, 512), (130, 1000, 520, 1000)])
@pytest.mark.parametrize("hdt", ["f32", "f16"])
def test_ffn_chain_equals_two_linear_layers(qg, tf, oracle, shape, hdt):
    """qg_ffn_forward (ll1 -> relu -> ll2, transformer.cu:63-71) against the two LinearLayer::forward calls it replaces
    and against the oracle; input given as floats and as codes from the fused ADD & NORM."""
    from test_gpu_parity import TORCH_DT, as_f32_np

    M, d_in, d_ff, d_out = shape
    rng = np.random.default_rng(sum(shape))
    X = make_edge_matrix(rng, M, d_in)
    l1, l2 = tf.PreparedLinear(d_in, d_ff), tf.PreparedLinear(d_ff, d_out)
    g = torch.Generator(device="cuda").manual_seed(1)
    l1.init_uniform(generator=g); l2.init_uniform(generator=g)
    H = torch.empty((M, d_ff), dtype=TORCH_DT[hdt], device="cuda")
    Y = torch.empty((M, d_out), device="cuda")
    tf.ffn_chain(l1, l2, to_dev(X), H, Y)
    h_exp = oracle.relu(oracle.quantized_mm(X, l1.w.cpu().numpy(), 127.0, bias=l1.b.cpu().numpy()))
    h_exp = as_f32_np(torch.from_numpy(h_exp).to(TORCH_DT[hdt]))
    y_exp = oracle.quantized_mm(h_exp, l2.w.cpu().numpy(), 127.0, bias=l2.b.cpu().numpy())
    assert same_f32(as_f32_np(H), h_exp) and same_f32(Y.cpu().numpy(), y_exp)
    if hdt == "f32":  # the unfused sequence through the module calls
        H2, Y2 = torch.empty_like(H), torch.empty_like(Y)
        l1.forward(to_dev(X), H2, tf.ACT_RELU)
        l2.forward(H2, Y2)
        assert torch.equal(H.view(torch.int32), H2.view(torch.int32)) and torch.equal(Y.view(torch.int32), Y2.view(torch.int32))


@pytest.mark.parametrize("shape", [(1, 1), (6, 8), (130, 33), (300, 512), (1000, 2048), (40, 4096), (9, 5000)])
def test_add_layernorm_quant_equals_separate_passes(qg, tf, oracle, shape):
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    A = rng.standard_normal(shape).astype(np.float32)
    R = rng.standard_normal(shape).astype(np.float32)
    if shape[0] > 5:
        A[1] = 0.0; R[1] = 0.0            # zero variance: 0/0 rows
        A[2, 0] = -50.0                   # negative first element dominates after the norm
        A[3, 1] = np.nan
    B = torch.empty(shape, device="cuda")
    Xq, Cx = qg.add_layernorm_quant(to_dev(A), to_dev(R), B)
    b_exp = oracle.add_layernorm(A, R)
    assert same_f32(B.cpu().numpy(), b_exp)
    eq, ec = oracle.absmax_quant_rows(b_exp)
    assert np.array_equal(Xq.cpu().numpy(), eq) and same_f32(Cx.cpu().numpy(), ec)
