"""How much of a re-read stream does the B200 L2 serve?  Back-to-back reductions over the same S-byte
buffer: effective GB/s per pass against S (HBM ~6.5 TB/s; anything above it is L2 hits)."""
import json, os, torch
res = {}
for mb in (8, 16, 24, 32, 48, 64, 80, 96, 128, 256):
    x = torch.ones(mb * (1 << 20) // 4, device="cuda")
    for _ in range(3): x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): x.sum()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    res[mb] = {"us_per_pass": round(us, 2), "gbs": round(mb * 1.048576 / us * 1e3, 0)}
    del x
print(json.dumps(res, indent=1))
json.dump(res, open("gpurun_out/l2_capacity_probe.json", "w"), indent=1)
