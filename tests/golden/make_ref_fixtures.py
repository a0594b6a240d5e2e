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


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    ref = load_ref()
    for name, X, W in cases(ref):
        np.savez_compressed(os.path.join(out, f"ref_{name}.npz"), **run_ref(ref, X, W))
        print("wrote", name)


if __name__ == "__main__":
    main()
