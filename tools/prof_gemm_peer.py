"""One process, two GPUs: the fused GEMM + exchange kernel of the column-parallel path with device 1 as its only peer.

Made for a single-process ncu capture (ncu must not wrap a multi-rank command): the launch on device 0 is the kernel
`bench.py --gpus 2` runs on every rank -- tcgen05 main loop, dequantize epilogue, each tile TMA-stored to the local
result and to the same block of the peer's result over NVLink.  `nvltx__bytes*` of that launch is the NVLink traffic the
exchange really causes, to set against the algorithmic M*N*4 bytes per peer.  Also checks the peer copy bit for bit.

  python tools/prof_gemm_peer.py            # 4096^3
"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
assert torch.cuda.device_count() >= 2 and torch.cuda.can_device_access_peer(0, 1)
torch.cuda.set_device(0)
# peer mapping both ways (torch enables it on the first cross-device copy)
probe = torch.zeros(1 << 20, device="cuda:0")
probe1 = probe.to("cuda:1"); probe.copy_(probe1)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)

g = torch.Generator(device="cuda:0").manual_seed(7)
A = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda:0", generator=g)
B = torch.randint(-127, 128, (n, n), dtype=torch.int8, device="cuda:0", generator=g)   # [K, N], the per-call layout
Cx = torch.rand(n, device="cuda:0", generator=g)
Cw = torch.rand(n, device="cuda:0", generator=g)
# the [M, 2N] results of "rank 0" (local) and "rank 1" (peer); rank 0 owns columns [0, N)
local = torch.zeros((n, 2 * n), dtype=torch.float32, device="cuda:0")
peer = torch.zeros((n, 2 * n), dtype=torch.float32, device="cuda:1")
plain = torch.empty((n, n), dtype=torch.float32, device="cuda:0")

reps = int(os.environ.get("QG_PEER_REPS", "5"))
for _ in range(reps):
    qg.gemm_s8_dequant_ex(A, B, False, Cx, Cw, local[:, :n], [peer.data_ptr()], 127.0)
torch.cuda.synchronize(0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    qg.gemm_s8_dequant_ex(A, B, False, Cx, Cw, local[:, :n], [peer.data_ptr()], 127.0)
e1.record()
torch.cuda.synchronize(0)
ms_peer = e0.elapsed_time(e1) / reps
e0.record()
for _ in range(reps):
    qg.gemm_s8_dequant(A, B, Cx, Cw, plain, 127.0)
e1.record()
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
ms_plain = e0.elapsed_time(e1) / reps
ok_local = torch.equal(local[:, :n].view(torch.int32), plain.view(torch.int32))
ok_peer = torch.equal(peer[:, :n].to("cuda:0").view(torch.int32), plain.view(torch.int32))
untouched = bool((peer[:, n:] == 0).all().item())
print(json.dumps({"shape": [n, n, n], "gemm_with_peer_store_ms": ms_peer, "gemm_local_only_ms": ms_plain,
                  "algorithmic_peer_bytes": n * n * 4, "egress_GBps": n * n * 4 / ms_peer / 1e6,
                  "local_block_equal": ok_local, "peer_block_equal": ok_peer, "peer_other_block_untouched": untouched}))
assert ok_local and ok_peer and untouched
