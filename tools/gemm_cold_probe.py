"""GEMM + dequant at 4096^3 (prepared weights, fp32 out): back to back (operands and output in L2 from the previous
call) vs after a 512 MiB write that evicts L2 (operands come from DRAM, as inside op_quantized_mm)."""
import importlib, json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = 4096
A = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Bt = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda")
Cx, Cw = torch.rand(n, device="cuda"), torch.rand(n, device="cuda")
O = torch.empty((n, n), device="cuda")
junk = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def run(flush, iters=30, rewarm=None):
    ts = []
    for i in range(iters + 3):
        if flush: junk.fill_(i & 0xff)
        if rewarm == "operands":  # bring A and Bt back into L2, leave the output cold
            A.view(torch.int32).sum(); Bt.view(torch.int32).sum()
        if rewarm == "output":
            O.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O); e1.record()
        torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1) * 1e3)
    return round(statistics.median(ts), 2), round(min(ts), 2)
print(json.dumps({"hot_median_min_us": run(False), "cold_median_min_us": run(True), "cold_but_operands_rewarmed": run(True, rewarm="operands"),
                  "cold_but_output_rewarmed": run(True, rewarm="output"), "hot_again": run(False)}))
