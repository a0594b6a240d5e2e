"""Generates tests/golden/ref_*.npz by running the REFERENCE's own kernels (oracle/_ref/libref_qmm.so,
compiled from /root/reference/src by oracle/Makefile) on a B200.

    gpurun -- python tests/golden/make_ref_fixtures.py gpurun_out/golden

then copy gpurun_out/golden/ref_*.npz into tests/golden/ and commit them.  Each file holds the
inputs and every intermediate of the reference pipeline (Cx, Cw, Xq, Wq, acc, O) plus the
reference's own op_quantized_mm output and its fp32 op_mm product.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import make_edge_matrix  # noqa: E402


def load_ref():
    return C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so"))


def p(a):
    return a.ctypes.data_as(C.c_void_p)


def run_ref(ref, X, W, range_=127.0):
    M, K = X.shape
    N = W.shape[1]
    O = np.empty((M, N), np.float32)
    O2 = np.empty((M, N), np.float32)
    Cfp = np.empty((M, N), np.float32)
    Cx = np.empty(M, np.float32)
    Cw = np.empty(N, np.float32)
    Xq = np.empty((M, K), np.int8)
    Wq = np.empty((K, N), np.int8)
    acc = np.empty((M, N), np.int32)
    rc = ref.ref_quantized_mm_parts(p(X), p(W), M, N, K, C.c_float(range_), p(O), p(Cx), p(Cw), p(Xq), p(Wq), p(acc))
    assert rc == 0, rc
    rc = ref.ref_op_quantized_mm(p(X), p(W), M, N, K, C.c_float(range_), p(O2))
    assert rc == 0, rc
    rc = ref.ref_mm_f32(p(X), p(W), M, N, K, p(Cfp))
    assert rc == 0, rc
    return dict(X=X, W=W, O=O, O_op=O2, C_fp32=Cfp, Cx=Cx, Cw=Cw, Xq=Xq, Wq=Wq, acc=acc)


def cases(ref):
    yield "3x3", np.array([[2, -1, -1], [0, 3, 2], [-1, -1, 0]], np.float32), np.array([[-1, 0], [0, -2], [-1, 2]], np.float32)
    rng = np.random.default_rng(0)
    yield "edge_24x40x56", make_edge_matrix(rng, 24, 56), np.ascontiguousarray(make_edge_matrix(rng, 40, 56).T)
    yield "rand_64x48x96", (rng.random((64, 96), dtype=np.float32) * 2 - 1), (rng.random((96, 48), dtype=np.float32) * 2 - 1)
    yield "normal_33x65x130", rng.standard_normal((33, 130)).astype(np.float32), (rng.standard_normal((130, 65)) * 0.02).astype(np.float32)
    # the timing driver's own input generator (cuRAND XORWOW, seed 0), small shape
    X = np.empty((32, 64), np.float32)
    W = np.empty((64, 16), np.float32)
    assert ref.ref_uniform_inputs(C.c_ulonglong(0), 32, 16, 64, p(X), p(W)) == 0
    yield "curand_seed0_32x16x64", X, W


def run_ref_attention(ref, Xq, Xkv, Wq, Wk, Wv, range_=127.0):
    """AttentionLayer::forward with the projections on op_quantized_mm, every intermediate kept."""
    sq, d_model = Xq.shape
    skv = Xkv.shape[0]
    d_k, d_v = Wq.shape[1], Wv.shape[1]
    Q, K, V = np.empty((sq, d_k), np.float32), np.empty((skv, d_k), np.float32), np.empty((skv, d_v), np.float32)
    S, P, out = np.empty((sq, skv), np.float32), np.empty((sq, skv), np.float32), np.empty((sq, d_v), np.float32)
    rc = ref.ref_attention_quantized(p(Xq), p(Xkv), p(Wq), p(Wk), p(Wv), sq, skv, d_model, d_k, d_v, C.c_float(range_),
                                     p(Q), p(K), p(V), p(S), p(P), p(out))
    assert rc == 0, rc
    return dict(Xq=Xq, Xkv=Xkv, Wq=Wq, Wk=Wk, Wv=Wv, Q=Q, K=K, V=V, S=S, P=P, out=out)


def attention_cases():
    rng = np.random.default_rng(7)
    u = lambda *s: (rng.random(s, dtype=np.float32) * 2 - 1)
    # the shape of src/test_attn.cu:100-134 (seq 2, d_model 3, d_k 2, d_v 4) with its hard-coded weights' ranges
    X = u(2, 3)
    yield "attn_self_2x3_dk2_dv4", X, X, u(3, 2), u(3, 2), u(3, 4)
    X = u(48, 64)
    s = np.float32(1 / 8)
    yield "attn_self_48x64_dk16_dv24", X, X, u(64, 16) * s, u(64, 16) * s, u(64, 24) * s
    yield "attn_cross_40_72x96_dk32_dv32", u(40, 96), u(72, 96), u(96, 32) * s, u(96, 32) * s, u(96, 32) * s
    X = rng.standard_normal((128, 512)).astype(np.float32)  # one head of BASELINE config 3
    s = np.float32(1 / 64 ** 0.5)
    yield "attn_self_128x512_dk64_dv64", X, X, u(512, 64) * s, u(512, 64) * s, u(512, 64) * s


def softmax_cases():
    rng = np.random.default_rng(8)
    yield "softmax_1x3", np.array([[1.0, 2.0, 3.0]], np.float32)  # src/test_softmax.cu:30-66
    A = (rng.standard_normal((200, 77)) * 4).astype(np.float32)
    A[3, 5] = np.float32(88.0); A[4, :] = np.float32(-1e30); A[5, 0] = np.float32(-0.0)
    yield "softmax_200x77", A


def run_ref_encoder_block(ref, X, heads, d_ff, Wq, Wk, Wv, W_O, W1, b1, W2, b2, range_=127.0):
    """Wq/Wk/Wv: [heads, d_model, d] stacks.  One Encoder-loop iteration on the reference's kernels."""
    h, d_model = X.shape
    out = np.empty((h, d_model), np.float32)
    rc = ref.ref_encoder_block(p(X), h, d_model, heads, d_ff, p(Wq), p(Wk), p(Wv), p(W_O), p(W1), p(b1), p(W2), p(b2),
                               C.c_float(range_), p(out))
    assert rc == 0, rc
    return out


def encoder_weights(rng, d_model, heads, d_ff):
    d = d_model // heads
    u = lambda lo, hi, *s: (rng.random(s, dtype=np.float32) * np.float32(hi - lo) + np.float32(lo))
    a = 1 / d ** 0.5
    Wq, Wk, Wv = u(-a, a, heads, d_model, d), u(-a, a, heads, d_model, d), u(-a, a, heads, d_model, d)
    W_O = u(-1, 1, d_model, d_model)
    a1, a2 = 1 / d_model ** 0.5, 1 / d_ff ** 0.5
    return dict(Wq=Wq, Wk=Wk, Wv=Wv, W_O=W_O, W1=u(-a1, a1, d_model, d_ff), b1=u(-a1, a1, 1, d_ff),
                W2=u(-a2, a2, d_ff, d_model), b2=u(-a2, a2, 1, d_model))


def encoder_cases():
    rng = np.random.default_rng(21)
    # the shape of transformer.cu:170-185 (6 x 8 input, 4 heads, d_ff 8), then a wider one
    for name, h, d_model, heads, d_ff in (("enc_6x8_h4_ff8", 6, 8, 4, 8), ("enc_48x64_h4_ff96", 48, 64, 4, 96)):
        X = (rng.random((h, d_model), dtype=np.float32) * 2 - 1)
        yield name, X, heads, d_ff, encoder_weights(rng, d_model, heads, d_ff)


def run_ref_decoder_block(ref, X, E, heads, d_ff, sa, ca, W1, b1, W2, b2, range_=127.0):
    """One Decoder-loop iteration (src/transformer.cu:91-166) on the reference's kernels.  sa / ca: dicts with
    Wq, Wk, Wv ([heads, d_model, d] stacks) and W_O of the self- and the cross-attention."""
    h, d_model = X.shape
    out = np.empty((h, d_model), np.float32)
    rc = ref.ref_decoder_block(p(X), p(E), h, E.shape[0], d_model, heads, d_ff, p(sa["Wq"]), p(sa["Wk"]), p(sa["Wv"]),
                               p(sa["W_O"]), p(ca["Wq"]), p(ca["Wk"]), p(ca["Wv"]), p(ca["W_O"]), p(W1), p(b1), p(W2), p(b2),
                               C.c_float(range_), p(out))
    assert rc == 0, rc
    return out


def decoder_weights(rng, d_model, heads, d_ff):
    sa = encoder_weights(rng, d_model, heads, d_ff)
    ca = encoder_weights(rng, d_model, heads, d_ff)
    return dict(sa={k: sa[k] for k in ("Wq", "Wk", "Wv", "W_O")}, ca={k: ca[k] for k in ("Wq", "Wk", "Wv", "W_O")},
                W1=ca["W1"], b1=ca["b1"], W2=ca["W2"], b2=ca["b2"])


def decoder_cases():
    rng = np.random.default_rng(23)
    # transformer.cu:170-185 shape (6 x 8, 4 heads, d_ff 8) with a 6-row encoder output; then a wider one whose
    # encoder sequence is longer than the decoder's (cross-attention scores are not square)
    for name, h, h_enc, d_model, heads, d_ff in (("dec_6x8_h4_ff8_enc6", 6, 6, 8, 4, 8), ("dec_40x64_h4_ff96_enc56", 40, 56, 64, 4, 96)):
        X = (rng.random((h, d_model), dtype=np.float32) * 2 - 1)
        E = (rng.random((h_enc, d_model), dtype=np.float32) * 2 - 1)
        yield name, X, E, heads, d_ff, decoder_weights(rng, d_model, heads, d_ff)


def save_decoder_case(path, X, E, heads, d_ff, w, out):
    flat = {f"sa_{k}": v for k, v in w["sa"].items()}
    flat.update({f"ca_{k}": v for k, v in w["ca"].items()})
    np.savez_compressed(path, X=X, E=E, heads=heads, d_ff=d_ff, out=out, W1=w["W1"], b1=w["b1"], W2=w["W2"], b2=w["b2"], **flat)


def addnorm_cases():
    rng = np.random.default_rng(22)
    A = rng.standard_normal((200, 77)).astype(np.float32)
    R = rng.standard_normal((200, 77)).astype(np.float32)
    yield "addnorm_200x77", A, R
    yield "addnorm_6x8_noadd", rng.standard_normal((6, 8)).astype(np.float32), None


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    ref = load_ref()
    for name, X, W in cases(ref):
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), **run_ref(ref, X, W))
        print("wrote", name)
    for name, Xq, Xkv, Wq, Wk, Wv in attention_cases():
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), **run_ref_attention(ref, Xq, Xkv, Wq, Wk, Wv))
        print("wrote", name)
    for name, X, heads, d_ff, w in encoder_cases():
        o = run_ref_encoder_block(ref, X, heads, d_ff, **w)
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), X=X, heads=heads, d_ff=d_ff, out=o, **w)
        print("wrote", name)
    for name, X, E, heads, d_ff, w in decoder_cases():
        o = run_ref_decoder_block(ref, X, E, heads, d_ff, **w)
        save_decoder_case(os.path.join(out, f"ref_{name}.npz"), X, E, heads, d_ff, w, o)
        print("wrote", name)
    for name, A, R in addnorm_cases():
        B = np.empty_like(A)
        assert ref.ref_add_layernorm(p(A), p(R) if R is not None else None, A.shape[0], A.shape[1], p(B)) == 0
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), A=A, B=B, **({"R": R} if R is not None else {}))
        print("wrote", name)
    for name, A in softmax_cases():
        B = np.empty_like(A)
        assert ref.ref_softmax(p(A), A.shape[0], A.shape[1], p(B)) == 0
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), A=A, B=B)
        print("wrote", name)


if __name__ == "__main__":
    main()
