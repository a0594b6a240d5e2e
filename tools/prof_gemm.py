"""Small driver for ncu captures of the tcgen05 GEMM (prepared K-major weights, fp32 output)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Bt = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Cx, Cw = torch.rand(n, device="cuda"), torch.rand(n, device="cuda")
O = torch.empty((n, n), dtype=torch.float32, device="cuda")
O16 = torch.empty((n, n), dtype=torch.float16, device="cuda")
for v in (qg.GEMM_TC_2SM, qg.GEMM_TC_1SM):
    qg.set_gemm_variant(v)
    for _ in range(3):
        qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O)
        qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O16)
torch.cuda.synchronize()
print("ok")
