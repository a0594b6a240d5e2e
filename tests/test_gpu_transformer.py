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


def test_decoder_block_batched_equals_per_sequence(tf):
    """No reference composite for the decoder (transformer.cu does not compile): its pieces are the
    pinned ones; here the batched call must equal per-sequence calls bit for bit, in place included."""
    g = torch.Generator(device="cuda").manual_seed(5)
    d_model, heads, d_ff, s, s_enc = 64, 4, 96, 24, 40
    blk = tf.DecoderBlock(d_model, heads, d_ff)
    blk.init_uniform(g)
    x = torch.rand((2 * s, d_model), device="cuda", generator=g) * 2 - 1
    enc = torch.rand((2 * s_enc, d_model), device="cuda", generator=g) * 2 - 1
    both = torch.empty_like(x)
    blk.forward(x, enc, both, batch=2)
    for b in range(2):
        one = torch.empty((s, d_model), device="cuda")
        blk.forward(x[b * s:(b + 1) * s], enc[b * s_enc:(b + 1) * s_enc], one)
        assert torch.equal(one.view(torch.int32), both[b * s:(b + 1) * s].view(torch.int32))
    xin = x.clone()
    blk.forward(xin, enc, xin, batch=2)  # in place, as later loop iterations run (transformer.cu:104-106)
    assert torch.equal(xin.view(torch.int32), both.view(torch.int32))
    assert torch.isfinite(both).all()
