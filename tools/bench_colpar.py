"""BASELINE config 4 (OPT-66B-shaped FFN, column-parallel over the GPUs of one box), under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29571 tools/bench_colpar.py

For fc1 (9216 -> 36864) and fc2 (36864 -> 9216), T = 4096 tokens, fp16 activations and outputs, int8
weights prepared once: time per forward (max over ranks, CUDA events, 3 warm-up + 10 timed) of
  fused  FusedColumnParallelLinear: the GEMM epilogue TMA-stores every tile into all ranks' results
  nccl   ColumnParallelLinear: local GEMM, then all_gather_into_tensor + permute
  local  the rank's shard alone, no exchange (what the exchange costs)
Rank 0 prints one JSON line per layer and writes gpurun_out/colpar_P.json.
"""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
colpar = importlib.import_module(qg.__name__ + ".colpar")
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
T = 4096


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = []
for name, K, N in (("opt66b_fc1", 9216, 36864), ("opt66b_fc2", 36864, 9216)):
    g = torch.Generator(device=dev).manual_seed(7)  # the same full weight on every rank, sliced by the layer
    W = (torch.randn((K, N), device=dev, generator=g) * 0.02).to(torch.float16)
    x = torch.randn((T, K), device=dev, generator=g).to(torch.float16)
    fused = colpar.FusedColumnParallelLinear(W, None, rank, world)
    nccl = colpar.ColumnParallelLinear(W, None, rank, world)
    del W
    torch.cuda.empty_cache()
    y_f = fused.forward(x)
    y_n = nccl.forward(x)
    torch.cuda.synchronize()
    same = bool(torch.equal(y_f, y_n))  # value equality (the NCCL path adds a zero bias: -0 becomes +0)
    res = {"layer": name, "world": world, "T": T, "K": K, "N": N, "fused_eq_nccl_bits": same}
    res["fused_us"] = timed(lambda: fused.forward(x))
    res["nccl_us"] = timed(lambda: nccl.forward(x))
    res["local_us"] = timed(lambda: nccl.local_forward(x))
    ops = 2.0 * T * K * N
    res["fused_tops_total"] = ops / res["fused_us"] / 1e6
    res["nccl_tops_total"] = ops / res["nccl_us"] / 1e6
    res["gather_bytes_per_rank_in"] = T * (N - N // world) * 2
    out.append(res)
    if rank == 0:
        print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    del fused, nccl, x, y_f, y_n
    torch.cuda.empty_cache()
    dist.barrier()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/colpar_{world}.json", "w"), indent=1)
dist.destroy_process_group()
