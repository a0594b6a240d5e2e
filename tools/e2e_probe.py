"""Host-buffer call (qg_quantized_mm_host) at 4096^3: time per call for the row-chunk count given in
QG_HOST_CHUNKS, next to the raw pinned H2D / D2H copy times that bound it."""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = 4096
Xh = (torch.rand((n, n)) * 2 - 1).pin_memory(); Wh = (torch.rand((n, n)) * 2 - 1).pin_memory(); Oh = torch.empty((n, n)).pin_memory()
for _ in range(3): qg.quantized_mm_host(Xh, Wh, out=Oh)
t0 = time.perf_counter()
for _ in range(10): qg.quantized_mm_host(Xh, Wh, out=Oh)
ms = (time.perf_counter() - t0) * 100
d = torch.empty((n, n), device="cuda")
def t(fn, k=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
h2d = t(lambda: d.copy_(Xh, non_blocking=True)); d2h = t(lambda: Oh.copy_(d, non_blocking=True))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty((n, n), device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(Xh, non_blocking=True)
    with torch.cuda.stream(s2): Oh.copy_(d2, non_blocking=True)
duplex = t(both)
print(json.dumps({"chunks": os.environ.get("QG_HOST_CHUNKS", "8"), "e2e_ms": round(ms, 3), "h2d_64MiB_ms": round(h2d, 3),
                  "d2h_64MiB_ms": round(d2h, 3), "h2d_and_d2h_concurrent_ms": round(duplex, 3)}))
