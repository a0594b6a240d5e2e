"""Launch sequence of the outlier-decomposed linear at the OPT-6.7B out-projection shape (T = 16384, 4096 -> 4096, fp16), for
`ncu --metrics gpu__time_duration.sum`: the plain int8 linear, then 6 / 16 / 64 outlier columns."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
M, K, N = 16384, 4096, 4096
g = torch.Generator(device="cuda").manual_seed(5)
lin = qg.LinearLayer(K, N, device="cuda", dtype=torch.float16)
lin.w.normal_(0, 0.02, generator=g)
lin.b.zero_()
lin.quantize_weights()
y = torch.empty((M, N), dtype=torch.float16, device="cuda")
X = torch.randn((M, K), device="cuda", generator=g).to(torch.float16)
for rep in range(2):
    lin.forward(X, y)
    for n in (6, 16, 64):
        cols = torch.randperm(K, device="cuda", generator=g)[:n].sort().values.to(torch.int32)
        lin.forward_outlier(X, y, cols)
torch.cuda.synchronize()
print("ok")
for rep in range(3):
    qg.outlier_cols(X, 6.0)
torch.cuda.synchronize()
