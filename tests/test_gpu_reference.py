"""GPU tests against the REFERENCE's own kernels (oracle/_ref/libref_qmm.so, built from
/root/reference/src in the dev container and shipped as a binary): on identical inputs the oracle,
the reference and the new path must agree bit for bit on Cx/Cw, int8 codes, int32 accumulators and
the fp32 output.  This is what pins the oracle (the reference asserts nothing)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

from conftest import make_edge_matrix
from test_gpu_parity import same_f32, to_dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so")
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_qmm.so not built (needs /root/reference at build time)")
    return C.CDLL(REF_SO)


def inputs(kind, M, N, K, ref):
    rng = np.random.default_rng(M * 1000003 + N * 1009 + K)
    if kind == "edge":
        return make_edge_matrix(rng, M, K), np.ascontiguousarray(make_edge_matrix(rng, N, K).T)
    if kind == "curand":  # the timing driver's generator, src/timing_quantize.cu:17-20
        X = np.empty((M, K), np.float32)
        W = np.empty((K, N), np.float32)
        assert ref.ref_uniform_inputs(C.c_ulonglong(0), M, N, K, X.ctypes.data_as(C.c_void_p),
                                      W.ctypes.data_as(C.c_void_p)) == 0
        return X, W
    return (rng.random((M, K), dtype=np.float32) * 2 - 1), (rng.random((K, N), dtype=np.float32) * 2 - 1)


@pytest.mark.parametrize("kind,shape", [("edge", (3, 2, 3)), ("edge", (24, 40, 56)), ("uniform", (256, 256, 256)),
                                        ("edge", (130, 70, 300)), ("curand", (2048, 512, 512)),
                                        ("curand", (512, 512, 1024))])
def test_reference_oracle_and_new_path_agree(qg, oracle, ref, kind, shape):
    from make_ref_fixtures import run_ref

    M, N, K = shape
    X, W = inputs(kind, M, N, K, ref)
    r = run_ref(ref, X, W)
    # 1. oracle == reference (pins the oracle)
    O, p = oracle.quantized_mm(X, W, 127.0, return_parts=True)
    assert same_f32(p["Cx"], r["Cx"]) and same_f32(p["Cw"], r["Cw"])
    assert np.array_equal(p["Xq"], r["Xq"]) and np.array_equal(p["Wq"], r["Wq"])
    assert np.array_equal(p["acc"], r["acc"])
    assert same_f32(O, r["O"]) and same_f32(r["O"], r["O_op"])
    assert same_f32(oracle.gemm_f32_ref(X, W), r["C_fp32"])
    # 2. new path == reference
    dO = torch.empty((M, N), device="cuda")
    qg.op_quantized_mm(to_dev(X), to_dev(W), dO, 127.0)
    assert same_f32(dO.cpu().numpy(), r["O"])
    Xq, Cx = qg.absmax_quant_rows(to_dev(X))
    Wq, Cw = qg.absmax_quant_cols(to_dev(W))
    assert np.array_equal(Xq.cpu().numpy(), r["Xq"]) and np.array_equal(Wq.cpu().numpy(), r["Wq"])
    assert same_f32(Cx.cpu().numpy(), r["Cx"]) and same_f32(Cw.cpu().numpy(), r["Cw"])
    acc = torch.empty((M, N), dtype=torch.int32, device="cuda")
    qg.op_mm(Xq, Wq, acc)
    assert np.array_equal(acc.cpu().numpy(), r["acc"])
    Cf = torch.empty((M, N), device="cuda")
    qg.op_mm(to_dev(X), to_dev(W), Cf)
    assert same_f32(Cf.cpu().numpy(), r["C_fp32"])
