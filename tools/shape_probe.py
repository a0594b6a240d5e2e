"""Prepared-weight linear (fp16 in/out) on a few shapes, 3 repeats each: same-box A/B of scheduler switches."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.bench_configs import timed
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
shapes = [("fc1_shard", 4096, 9216, 4608), ("fc2_shard", 4096, 36864, 1152), ("sq4096", 4096, 4096, 4096), ("n1152_k4096", 4096, 4096, 1152),
          ("opt_out", 16384, 4096, 4096), ("m256_8192", 256, 8192, 8192), ("m128_k16384", 128, 16384, 4096),
          ("ffn2_cfg3", 4096, 2048, 512)]
res = {}
for name, M, K, N in shapes:
    lin = qg.LinearLayer(K, N, device="cuda", dtype=torch.float16); lin.w.normal_(0, 0.02); lin.b.zero_(); lin.quantize_weights()
    X = [torch.randn((M, K), device="cuda", dtype=torch.float16) for _ in range(2)]
    y = torch.empty((M, N), device="cuda", dtype=torch.float16)
    res[name] = [round(timed(lambda i: lin.forward(X[i & 1], y)), 1) for _ in range(3)]
    del lin, X, y; torch.cuda.empty_cache()
print(json.dumps({"QG_SPLIT_K": os.environ.get("QG_SPLIT_K", "auto"), **res}))
