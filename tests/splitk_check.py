"""Run by tests/test_gpu_parity.py::test_forced_split_k / test_two_pair_clusters in a subprocess with QG_SPLIT_K=n or
QG_GEMM_NP=2 (the library reads the switches once): int32 accumulators, dequantized outputs (fp32 / fp16 / bf16, bias, ReLU) and the whole op must stay
bit-exact against the oracle when every tensor-core product is cut into n k-slices."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
tf = importlib.import_module(qg.__name__ + ".transformer")
oracle = importlib.import_module("oracle")
from conftest import make_edge_matrix  # noqa: E402
from test_gpu_parity import same_f32  # noqa: E402  (NaN-aware bit comparison)

DEV = "cuda"
rng = np.random.default_rng(int(os.environ.get("QG_SPLIT_K", "1")))
bad = 0
for (M, N, K) in [(128, 256, 1024), (200, 300, 1000), (130, 520, 2048), (384, 512, 640), (1024, 768, 4096), (700, 520, 1024)]:
    A = rng.integers(-127, 128, (M, K), dtype=np.int8)
    B = rng.integers(-127, 128, (K, N), dtype=np.int8)
    dA, dB = torch.from_numpy(A).to(DEV), torch.from_numpy(B).to(DEV)
    acc = torch.empty((M, N), dtype=torch.int32, device=DEV)
    qg.op_mm(dA, dB, acc)
    exp_acc = oracle.gemm_s8s8s32(A, B)
    bad += int(not np.array_equal(acc.cpu().numpy(), exp_acc))
    X = make_edge_matrix(rng, M, K)
    lin = tf.PreparedLinear(K, N)
    lin.w.copy_(torch.from_numpy(np.ascontiguousarray(make_edge_matrix(rng, N, K).T)).to(DEV))
    lin.b.copy_(torch.from_numpy(rng.standard_normal((1, N)).astype(np.float32)).to(DEV))
    exp = oracle.quantized_mm(X, lin.w.cpu().numpy(), 127.0, bias=lin.b.cpu().numpy())
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        y = torch.empty((M, N), dtype=dt, device=DEV)
        lin.forward(torch.from_numpy(X).to(DEV), y, tf.ACT_RELU)
        want = torch.from_numpy(oracle.relu(exp)).to(dt)
        bad += int(not same_f32(y.float().cpu().numpy(), want.float().numpy()))
    O = torch.empty((M, N), device=DEV)
    qg.op_quantized_mm(torch.from_numpy(X).to(DEV), lin.w, O, 127.0)
    e2 = oracle.quantized_mm(X, lin.w.cpu().numpy(), 127.0)
    bad += int(not same_f32(O.cpu().numpy(), e2))
print("split-k check:", "OK" if bad == 0 else f"{bad} mismatches")
sys.exit(1 if bad else 0)
