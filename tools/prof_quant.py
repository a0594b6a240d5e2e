"""Small driver for ncu captures of the quantizer kernels (n x n fp32)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X = torch.rand((n, n), device="cuda") * 2 - 1
W = torch.rand((n, n), device="cuda") * 2 - 1
Wq = torch.empty((n, n), dtype=torch.int8, device="cuda")
Wt = torch.empty((n, n), dtype=torch.int8, device="cuda")
Xq = torch.empty((n, n), dtype=torch.int8, device="cuda")
Cx = torch.empty(n, device="cuda"); Cw = torch.empty(n, device="cuda")
for _ in range(3):
    qg.absmax_quant_rows(X, 127.0, 0, Xq, Cx)
    qg.absmax_quant_cols(W, 127.0, 0, Wq, Cw)
    qg.prepare_weights(W, 127.0, 0, Wt, Cw)
torch.cuda.synchronize()
print("ok")
