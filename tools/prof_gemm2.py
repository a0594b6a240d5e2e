"""ncu driver, round 2: the 2-SM tcgen05 GEMM at n^3 with fp32, fp16 and raw int32 outputs (prepared K-major weights)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
A = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Bt = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Cx, Cw = torch.rand(n, device="cuda"), torch.rand(n, device="cuda")
outs = [torch.empty((n, n), dtype=dt, device="cuda") for dt in (torch.float32, torch.float16, torch.int32)]
qg.set_gemm_variant(qg.GEMM_TC_2SM)
for _ in range(3):
    for O in outs:
        qg.gemm_s8t_dequant(A, Bt, None if O.dtype == torch.int32 else Cx, None if O.dtype == torch.int32 else Cw, O)
torch.cuda.synchronize()
print("ok")
