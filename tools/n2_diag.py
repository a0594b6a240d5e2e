"""Multi-GPU bench diagnosis (run under torchrun): where does a column-parallel step spend its time?

Times, per rank, with CUDA events over 20 iterations after 5 warm-ups:
  local      quantizers + GEMM, no exchange
  nccl_ag    all_gather_into_tensor of the [M,N] fp32 block alone
  local+ag   both
  symm_bar   two symmetric-memory barriers alone
  fused      barrier, GEMM with peer TMA stores, barrier
  copy_peer  plain tensor copy into the peer's symmetric buffer (NVLink store bandwidth)
"""
import importlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
M = N = K = 4096
X = torch.rand((M, K), device=dev) * 2 - 1
W = torch.rand((K, N), device=dev) * 2 - 1
O = torch.empty((M, N), device=dev)
Xq = torch.empty((M, K), dtype=torch.int8, device=dev)
Wq = torch.empty((K, N), dtype=torch.int8, device=dev)
Cx, Cw = torch.empty(M, device=dev), torch.empty(N, device=dev)
gathered = torch.empty((world, M, N), device=dev)
import torch.distributed._symmetric_memory as symm_mem
symm_out = symm_mem.empty((M, N * world), dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(symm_out, dist.group.WORLD)
peer_ptrs = [int(hdl.buffer_ptrs[r]) + rank * N * 4 for r in range(world) if r != rank]
peer_buf = hdl.get_buffer((rank + 1) % world, (M, N * world), torch.float32)


def local():
    qg.absmax_quant_rows(X, 127.0, qg.MODE_REF_EXACT, Xq, Cx)
    qg.absmax_quant_cols(W, 127.0, qg.MODE_REF_EXACT, Wq, Cw)
    qg.gemm_s8_dequant(Xq, Wq, Cx, Cw, O, 127.0)


def fused():
    qg.absmax_quant_rows(X, 127.0, qg.MODE_REF_EXACT, Xq, Cx)
    qg.absmax_quant_cols(W, 127.0, qg.MODE_REF_EXACT, Wq, Cw)
    hdl.barrier(channel=0)
    qg.gemm_s8_dequant_ex(Xq, Wq, False, Cx, Cw, symm_out[:, rank * N:(rank + 1) * N], peer_ptrs, 127.0)
    hdl.barrier(channel=1)


def bars():
    hdl.barrier(channel=0)
    hdl.barrier(channel=1)


cases = {
    "local": local,
    "nccl_ag": lambda: dist.all_gather_into_tensor(gathered, O),
    "local+ag": lambda: (local(), dist.all_gather_into_tensor(gathered, O)),
    "symm_bar": bars,
    "fused": fused,
    "copy_peer": lambda: peer_buf[:, rank * N:(rank + 1) * N].copy_(O),
}
res = {}
for name, fn in cases.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    host_ms = (time.perf_counter() - t0) * 1e3 / 20
    torch.cuda.synchronize()
    res[name] = {"dev_us": round(e0.elapsed_time(e1) / 20 * 1e3, 1), "host_enqueue_us": round(host_ms * 1e3, 1)}
    dist.barrier()
print(json.dumps({"rank": rank, **res}), flush=True)
dist.destroy_process_group()
