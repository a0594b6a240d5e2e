"""Does the second pass of the column quantizer find in L2 what the first pass streamed?  Timed without a profiler (ncu's
kernel replay saves / restores memory between passes and so destroys exactly the residency in question):
  hot   op_multiply<float,int8_t>(W_a) (pass 2) right after op_absmax(W_a) (pass 1)
  cold  the same right after op_absmax(W_b) (W_a was evicted by 3 other matrices first)
  torch the same right after a torch reduction over W_a (default cache policy instead of our hinted loads)
for K = 4096 and N = 1024 .. 4096 (16 .. 64 MiB of fp32)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
K = 4096
res = {}
for N in (1024, 2048, 3072, 4096, 6144):
    Ws = [torch.rand((K, N), device="cuda") * 2 - 1 for _ in range(5)]
    cw = torch.empty((1, N), device="cuda")
    s = torch.rand((1, N), device="cuda") + 100.0
    Wq = torch.empty((K, N), dtype=torch.int8, device="cuda")

    def timed(pre, iters=20):
        ts = []
        for i in range(iters):
            for j in range(1, 5):  # evict W_0
                Ws[j].mul_(1.0)
            pre()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            qg.op_multiply(Ws[0], s, Wq)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    res[N] = {"mib": K * N * 4 / 2**20,
              "pass2_hot_us": timed(lambda: qg.op_absmax(Ws[0], cw)),
              "pass2_cold_us": timed(lambda: qg.op_absmax(Ws[1], cw)),
              "pass2_after_torch_read_us": timed(lambda: Ws[0].sum()),
              "pass2_after_torch_write_us": timed(lambda: Ws[0].mul_(1.0))}
    print(N, json.dumps(res[N]), flush=True)
    del Ws
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/l2_reuse_probe.json", "w"), indent=1)
