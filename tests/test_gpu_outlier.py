"""GPU tests of the LLM.int8()-style outlier decomposition (not in the reference: parity unpinned, the
oracle is our own restatement in oracle.quantized_mm_outlier).  Everything is still bit-exact against it."""
import numpy as np
import pytest
import torch

from conftest import make_edge_matrix
from test_gpu_parity import DEV, TORCH_DT, as_f32_np, same_f32, seed_of, to_dev

pytestmark = pytest.mark.gpu


def outlier_input(rng, M, K, cols, dt):
    X = rng.standard_normal((M, K)).astype(np.float32)
    X[:, cols] *= 20.0
    t = to_dev(X, dt)
    return t, as_f32_np(t)


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape", [(64, 256), (300, 1000), (1024, 4096), (33, 70), (16, 36864)])
def test_outlier_column_detection(qg, oracle, shape, dt):
    rng = np.random.default_rng(seed_of(shape, dt, "det"))
    M, K = shape
    cols = np.sort(rng.choice(K, min(6, K), replace=False))
    X, Xh = outlier_input(rng, M, K, cols, dt)
    thr = 6.0
    expect = oracle.outlier_columns(Xh, thr)
    idx, n = qg.outlier_cols(X, thr, max_idx=max(1, K))
    assert n == expect.size
    assert np.array_equal(idx.cpu().numpy(), expect)
    # NaN counts as an outlier; exact threshold value does not (strict comparison)
    Xh2 = Xh.copy()
    Xh2[:] = np.clip(Xh2, -1, 1)
    Xh2[1, 3] = np.nan
    Xh2[2, 5] = 6.0
    Xh2[0, K - 1] = -6.0001 if dt == "f32" else -6.5
    X2 = to_dev(Xh2, dt)
    idx2, n2 = qg.outlier_cols(X2, thr, max_idx=K)
    assert np.array_equal(idx2.cpu().numpy(), oracle.outlier_columns(as_f32_np(X2), thr))


@pytest.mark.parametrize("variant", ["SIMT", "TC_1SM", "TC_2SM"])
@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
@pytest.mark.parametrize("shape,n_out", [((128, 256, 512), 6), ((200, 300, 520), 3), ((512, 768, 1024), 13),
                                         ((64, 64, 4096), 16), ((256, 512, 256), 0), ((130, 264, 1000), 8),
                                         ((384, 512, 1024), 17), ((300, 520, 768), 40), ((256, 768, 2048), 64)])
def test_linear_forward_with_outlier_decomposition(qg, oracle, shape, n_out, dt, variant):
    M, N, K = shape
    rng = np.random.default_rng(seed_of(shape, dt, n_out))
    cols = np.sort(rng.choice(K, n_out, replace=False)).astype(np.int32)
    if n_out >= 3:
        cols[0] = 0  # the signed-first-element rule then applies to a zeroed entry
        cols = np.unique(cols)
    X, Xh = outlier_input(rng, M, K, cols, dt)
    lin = qg.LinearLayer(K, N, device=DEV, dtype=TORCH_DT[dt])
    lin.w.copy_(to_dev((rng.standard_normal((K, N)) * 0.05).astype(np.float32), dt))
    lin.b.copy_(to_dev(rng.standard_normal((1, N)).astype(np.float32)))
    y = torch.empty((M, N), dtype=TORCH_DT[dt], device=DEV)
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    try:
        lin.forward_outlier(X, y, torch.from_numpy(cols).to(DEV))
        torch.cuda.synchronize()
    finally:
        qg.set_gemm_variant(qg.GEMM_AUTO)
    expect, parts = oracle.quantized_mm_outlier(Xh, as_f32_np(lin.w), 6.0, bias=lin.b.cpu().numpy(),
                                                side_dtype="bf16" if dt == "bf16" else "f16", idx=cols)
    if dt == "f32":
        assert same_f32(y.cpu().numpy(), expect)
    else:
        assert torch.equal(y.cpu(), torch.from_numpy(expect).to(TORCH_DT[dt]))


def test_decomposition_reduces_error_on_outlier_features(qg):
    """The point of the split (LLM.int8, BASELINE config 4): with 6 outlier feature dims scaled by 20,
    routing them through the 16-bit side product cuts the error of the int8 linear several-fold."""
    torch.manual_seed(0)
    M, K, N = 2048, 4096, 4096
    X = torch.randn((M, K), device=DEV, dtype=torch.float16)
    cols = torch.tensor([11, 300, 1027, 2222, 3000, 4001], dtype=torch.int32, device=DEV)
    X[:, cols.long()] *= 20
    lin = qg.LinearLayer(K, N, device=DEV, dtype=torch.float16)
    lin.w.normal_(0, 0.02)
    lin.b.zero_()
    y0 = torch.empty((M, N), dtype=torch.float16, device=DEV)
    y1 = torch.empty_like(y0)
    lin.forward(X, y0)
    idx, n = qg.outlier_cols(X, 6.0)
    assert n == 6 and torch.equal(idx, cols)
    lin.forward_outlier(X, y1, idx)
    ref = X.float() @ lin.w.float()
    e0 = (y0.float() - ref).abs().mean().item()
    e1 = (y1.float() - ref).abs().mean().item()
    assert e1 < 0.5 * e0, (e0, e1)


def test_too_many_outlier_columns_is_an_error(qg):
    """More columns than the fused epilogue takes (64): a clean QG_ENOTSUP, nothing written out of bounds."""
    M, K, N = 128, 512, 256
    X = torch.randn((M, K), device=DEV)
    lin = qg.LinearLayer(K, N, device=DEV)
    y = torch.empty((M, N), device=DEV)
    idx = torch.arange(0, 65, dtype=torch.int32, device=DEV)
    with pytest.raises((qg.QGemmError, AssertionError)):
        lin.forward_outlier(X, y, idx)


def test_index_list_order_does_not_matter(qg, oracle):
    """The side product pairs X[:, k] with W[k, :] by the column's rank in the outlier MASK, so an unsorted index list, a
    duplicate or an out-of-range entry gives the result of the sorted, de-duplicated, in-range set."""
    M, N, K = 256, 384, 1024
    rng = np.random.default_rng(17)
    cols = np.array([900, 3, 512, 77, 640], dtype=np.int32)
    X, Xh = outlier_input(rng, M, K, cols, "f16")
    lin = qg.LinearLayer(K, N, device=DEV, dtype=torch.float16)
    lin.w.copy_(to_dev((rng.standard_normal((K, N)) * 0.05).astype(np.float32), "f16"))
    lin.b.copy_(to_dev(rng.standard_normal((1, N)).astype(np.float32)))
    y = torch.empty((M, N), dtype=torch.float16, device=DEV)
    messy = np.array([900, 3, 512, 77, 640, 3, 5000, -1], dtype=np.int32)  # unsorted, a duplicate, two out of range
    lin.forward_outlier(X, y, torch.from_numpy(messy).to(DEV))
    torch.cuda.synchronize()
    expect, _ = oracle.quantized_mm_outlier(Xh, as_f32_np(lin.w), 6.0, bias=lin.b.cpu().numpy(), side_dtype="f16", idx=np.sort(cols))
    assert torch.equal(y.cpu(), torch.from_numpy(expect).to(torch.float16))
