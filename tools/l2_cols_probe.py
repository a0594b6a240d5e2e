"""Does the column quantizer's second pass find W in L2?  Times absmax_quant_cols on W[K, n] views of
a [4096, 4096] fp32 matrix (row stride 4096) for several n, with a 256 MiB L2-evicting write before
each call or not, and on column halves back to back."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
dev = "cuda"
K = N = 4096
Ws = [torch.rand((K, N), device=dev) * 2 - 1 for _ in range(4)]
Wq = torch.empty((K, N), dtype=torch.int8, device=dev)
Cw = torch.empty(N, device=dev)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fn, n=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        junk.fill_(i & 0xff)      # evict L2
        e0.record(); fn(i); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3

res = {}
for n in (512, 1024, 2048, 3072, 4096):
    us = timed(lambda i: qg.absmax_quant_cols(Ws[i & 3][:, :n], 127.0, qg.MODE_REF_EXACT, Wq[:, :n], Cw[:n]))
    res[f"cols_{n}"] = {"us": round(us, 2), "mb": K * n * 4 / 1e6, "us_per_64mb": round(us * 4096 / n, 2)}
for parts in (2, 4):
    w = N // parts
    def split(i):
        for p in range(parts):
            qg.absmax_quant_cols(Ws[i & 3][:, p * w:(p + 1) * w], 127.0, qg.MODE_REF_EXACT, Wq[:, p * w:(p + 1) * w], Cw[p * w:(p + 1) * w])
    res[f"split_{parts}"] = {"us": round(timed(split), 2)}
for parts in (2, 4):  # row panels are not a valid split for column scales; timing probe of the access pattern only
    h = K // parts
    def rsplit(i):
        for p in range(parts):
            qg.absmax_quant_cols(Ws[i & 3][p * h:(p + 1) * h], 127.0, qg.MODE_REF_EXACT, Wq[p * h:(p + 1) * h], Cw)
    res[f"rowpanel_{parts}_timing_only"] = {"us": round(timed(rsplit), 2)}
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "l2_cols_probe.json"), "w"), indent=1)
