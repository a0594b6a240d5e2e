"""GPU parity of the attention widening: op_softmax and
AttentionLayer::forward with the three projections on the quantized linear path.

Anchors, strongest first:
  * the REFERENCE's kernels, live (oracle/_ref/libref_qmm.so: ref_softmax, ref_attention_quantized) and as
    committed fixtures (tests/golden/ref_attn_*.npz, ref_softmax_*.npz): BIT-EXACT -- the new kernels keep the
    reference's operation order and call the same device expf;
  * the CPU oracle (libm expf differs from the device's in the last bits): tolerance below.
"""
import ctypes as C
import glob
import os
import sys

import numpy as np
import pytest
import torch

from test_gpu_parity import same_f32, to_dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so")
sys.path.insert(0, GOLDEN)

# device expf vs libm expf: <= 2 ulp each on e_j, then a sum and a division
SOFTMAX_RTOL, SOFTMAX_ATOL = 2e-6, 1e-37
# attention output = P @ V with |V| = O(1): errors of P accumulate over skv terms
ATTN_RTOL, ATTN_ATOL = 1e-5, 2e-6


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_qmm.so not built (needs /root/reference at build time)")
    return C.CDLL(REF_SO)


def fused_w(Wq, Wk, Wv):
    return np.ascontiguousarray(np.concatenate([Wq, Wk, Wv], axis=1))


# ------------------------------------------------------------------------------------------ softmax
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_softmax_*.npz"))))
def test_softmax_matches_reference_fixture_bit_for_bit(qg, path):
    g = np.load(path)
    A = to_dev(g["A"])
    B = torch.empty_like(A)
    qg.op_softmax(A, B)
    assert same_f32(B.cpu().numpy(), g["B"])


@pytest.mark.parametrize("shape", [(1, 1), (1, 3), (5, 32), (129, 33), (300, 200), (1000, 64)])
@pytest.mark.parametrize("scale", [1.0, 0.125, 0.17677669])
def test_softmax_vs_oracle_and_in_place(qg, oracle, shape, scale):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    A = (rng.standard_normal(shape) * 5).astype(np.float32)
    dA = to_dev(A)
    dB = torch.empty_like(dA)
    qg.op_softmax(dA, dB, scale)
    got = dB.cpu().numpy()
    np.testing.assert_allclose(got, oracle.softmax_rows(A, scale), rtol=SOFTMAX_RTOL, atol=SOFTMAX_ATOL)
    np.testing.assert_allclose(got.sum(axis=1), 1.0, rtol=1e-5)
    qg.op_softmax(dA, dA, scale)  # in place
    assert same_f32(dA.cpu().numpy(), got)


def test_softmax_and_layernorm_wide_rows_take_the_tiled_kernels(qg, oracle):
    """More than 4096 columns: the warp-per-row / CTA-rows kernels hand over to the tiled thread-per-row ones."""
    import importlib

    tf = importlib.import_module(qg.__name__ + ".transformer")
    rng = np.random.default_rng(41)
    A = (rng.standard_normal((70, 5000)) * 3).astype(np.float32)
    B = torch.empty((70, 5000), device="cuda")
    qg.op_softmax(to_dev(A), B, 0.5)
    np.testing.assert_allclose(B.cpu().numpy(), oracle.softmax_rows(A, 0.5), rtol=SOFTMAX_RTOL, atol=SOFTMAX_ATOL)
    R = rng.standard_normal((70, 5000)).astype(np.float32)
    tf.add_layernorm(to_dev(A), to_dev(R), B)
    assert same_f32(B.cpu().numpy(), oracle.add_layernorm(A, R))


def test_softmax_strided_views_and_rows_the_reference_grid_skips(qg, ref):
    """300 rows x 40 columns: the reference's grid (ceil(40/256) = 1 block of 256 threads, op_softmax.cuh:38)
    leaves rows 256.. untouched; ours computes them.  Rows 0..255 agree bit for bit."""
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((300, 40)) * 3).astype(np.float32)
    big = torch.zeros((300, 64), device="cuda")
    big[:, 8:48] = to_dev(A)
    out = torch.full((300, 50), -7.0, device="cuda")
    qg.op_softmax(big[:, 8:48], out[:, 3:43])
    got = out.cpu().numpy()
    assert np.all(got[:, :3] == -7.0) and np.all(got[:, 43:] == -7.0)
    Bref = np.zeros_like(A)
    assert ref.ref_softmax(A.ctypes.data_as(C.c_void_p), 300, 40, Bref.ctypes.data_as(C.c_void_p)) == 0
    assert same_f32(got[:256, 3:43], Bref[:256])
    np.testing.assert_allclose(got[256:, 3:43].sum(axis=1), 1.0, rtol=1e-5)


# ---------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "ref_attn_*.npz"))))
def test_attention_matches_reference_fixture_bit_for_bit(qg, oracle, path):
    g = np.load(path)
    d_k, d_v = g["Wq"].shape[1], g["Wv"].shape[1]
    att = qg.AttentionLayer(g["Xq"].shape[1], d_k, d_v)
    att.W_qkv.copy_(to_dev(fused_w(g["Wq"], g["Wk"], g["Wv"])))
    out = torch.empty((g["Xq"].shape[0], d_v), device="cuda")
    xq = to_dev(g["Xq"])
    if "self" in os.path.basename(path):
        att.forward(xq, out)                      # AttentionLayer::forward(X, output), attention.cuh:47
    else:
        att.forward(xq, to_dev(g["Xkv"]), out)    # the 3-argument form of transformer.cu:37
    assert same_f32(out.cpu().numpy(), g["out"])
    # the projections on their own are the pinned quantized path
    Q = torch.empty((g["Xq"].shape[0], d_k), device="cuda")
    qg.op_quantized_mm(xq, to_dev(g["Wq"]), Q, 127.0)
    assert same_f32(Q.cpu().numpy(), g["Q"])
    # CPU oracle: projections and scores bit-exact, softmax / output within tolerance
    o = oracle.attention_forward(g["Xq"], g["Xkv"], g["Wq"], g["Wk"], g["Wv"], return_parts=True)
    assert same_f32(o["Q"], g["Q"]) and same_f32(o["K"], g["K"]) and same_f32(o["V"], g["V"]) and same_f32(o["S"], g["S"])
    np.testing.assert_allclose(o["P"], g["P"], rtol=SOFTMAX_RTOL, atol=SOFTMAX_ATOL)
    np.testing.assert_allclose(o["out"], g["out"], rtol=ATTN_RTOL, atol=ATTN_ATOL)


@pytest.mark.parametrize("sq,skv,d_model,d_k,d_v", [(2, 2, 3, 2, 4), (33, 33, 40, 8, 12), (64, 100, 128, 32, 16),
                                                    (128, 128, 512, 64, 64),
                                                    (250, 100, 64, 32, 32),    # four query row blocks of the one-kernel core (the reference's softmax grid stops at row 256)
                                                    (64, 200, 64, 32, 32),     # more keys than a CTA takes: the three-kernel path
                                                    (130, 128, 96, 64, 60)])   # ragged last row block, d_v below the tile width
def test_attention_matches_live_reference(qg, ref, sq, skv, d_model, d_k, d_v):
    from make_ref_fixtures import run_ref_attention

    rng = np.random.default_rng(sq * 7 + skv)
    u = lambda *s: (rng.random(s, dtype=np.float32) * 2 - 1)
    Xq = u(sq, d_model)
    Xkv = Xq if sq == skv else u(skv, d_model)
    s = np.float32(1 / d_k ** 0.5)
    Wq, Wk, Wv = u(d_model, d_k) * s, u(d_model, d_k) * s, u(d_model, d_v) * s
    r = run_ref_attention(ref, Xq, Xkv, Wq, Wk, Wv)
    out = torch.empty((sq, d_v), device="cuda")
    dq = to_dev(Xq)
    dkv = dq if Xkv is Xq else to_dev(Xkv)
    qg.attention_forward(dq, dkv, to_dev(fused_w(Wq, Wk, Wv)), out, 1, d_k, d_v)
    assert same_f32(out.cpu().numpy(), r["out"])


@pytest.mark.parametrize("batch,seq,d_model,heads", [(1, 6, 8, 4), (3, 20, 32, 4), (2, 128, 512, 8)])
def test_multi_head_fused_projection_equals_per_head_layers(qg, oracle, batch, seq, d_model, heads):
    """transformer.cu:27-50 loops heads with one AttentionLayer each and concatenates on the host; the fused
    call must give the same bits (row scales depend only on X, column scales only on their column)."""
    rng = np.random.default_rng(batch * 100 + seq)
    X = (rng.random((batch * seq, d_model), dtype=np.float32) * 2 - 1)
    mha = qg.MultiHeadAttention(d_model, heads)
    g = torch.Generator(device="cuda").manual_seed(3)
    mha.init_uniform(g)
    dX = to_dev(X)
    out = torch.empty((batch * seq, d_model), device="cuda")
    mha.forward(dX, dX, out, batch=batch)
    got = out.cpu().numpy()
    # per (sequence, head): the reference-shaped single-head layer
    d = d_model // heads
    for b in range(batch):
        xb = dX[b * seq:(b + 1) * seq]
        for h in range(heads):
            att = qg.AttentionLayer(d_model, d, d)
            wq, wk, wv = mha.head_weights(h)
            att.W_q.copy_(wq); att.W_k.copy_(wk); att.W_v.copy_(wv)
            o1 = torch.empty((seq, d), device="cuda")
            att.forward(xb, o1)
            assert same_f32(o1.cpu().numpy(), got[b * seq:(b + 1) * seq, h * d:(h + 1) * d])
    if batch * seq <= 64:
        exp = oracle.multi_head_attention(X, X, mha.W_qkv.cpu().numpy(), heads, d, d, batch)
        np.testing.assert_allclose(got, exp, rtol=ATTN_RTOL, atol=ATTN_ATOL)


def test_cross_attention_multi_head_vs_oracle(qg, oracle):
    rng = np.random.default_rng(11)
    batch, sq, skv, d_model, heads = 2, 10, 14, 24, 3
    Xq = (rng.random((batch * sq, d_model), dtype=np.float32) * 2 - 1)
    Xkv = (rng.random((batch * skv, d_model), dtype=np.float32) * 2 - 1)
    W = ((rng.random((d_model, heads * 24), dtype=np.float32) * 2 - 1) / 3).astype(np.float32)
    out = torch.empty((batch * sq, heads * 8), device="cuda")
    qg.attention_forward(to_dev(Xq), to_dev(Xkv), to_dev(W), out, heads, 8, 8, batch)
    exp = oracle.multi_head_attention(Xq, Xkv, W, heads, 8, 8, batch)
    np.testing.assert_allclose(out.cpu().numpy(), exp, rtol=ATTN_RTOL, atol=ATTN_ATOL)


@pytest.mark.parametrize("batch,sq,skv,d_model,heads", [(1, 6, 6, 8, 4), (3, 20, 20, 40, 4), (2, 10, 14, 24, 3), (2, 128, 128, 512, 8),
                                                         (2, 128, 96, 512, 8), (1, 70, 200, 72, 2)])
def test_prepared_projection_weights_give_the_same_bits(qg, batch, sq, skv, d_model, heads):
    """qg_attention_forward_prepared: W_q | W_k | W_v column-quantized once (qg_prepare_weights) instead of on every call.
    A column's scale and codes depend only on that column, so self- and cross-attention (whose K / V projection uses a row
    block of the prepared codes) must equal the per-call form bit for bit -- d_model not a multiple of 16 included."""
    rng = np.random.default_rng(batch * 1000 + sq * 10 + skv)
    Xq = to_dev(rng.random((batch * sq, d_model), dtype=np.float32) * 2 - 1)
    Xkv = Xq if sq == skv else to_dev(rng.random((batch * skv, d_model), dtype=np.float32) * 2 - 1)
    mha = qg.MultiHeadAttention(d_model, heads)
    mha.init_uniform(torch.Generator(device="cuda").manual_seed(5))
    a = torch.empty((batch * sq, d_model), device="cuda")
    b = torch.full_like(a, float("nan"))
    mha.forward(Xq, Xkv, a, batch=batch)
    mha.forward(Xq, Xkv, b, batch=batch, prepared=True)
    assert same_f32(a.cpu().numpy(), b.cpu().numpy())
    # new weights: the cache is dropped by init_uniform, and quantize_weights() refreshes it after a direct write
    mha.W_qkv.mul_(0.5)
    mha.quantize_weights()
    mha.forward(Xq, Xkv, a, batch=batch)
    mha.forward(Xq, Xkv, b, batch=batch, prepared=True)
    assert same_f32(a.cpu().numpy(), b.cpu().numpy())
