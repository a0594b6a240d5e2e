"""BASELINE config 5 as a Megatron pair: one OPT-66B-shaped FFN (9216 -> 36864 -> 9216), T = 4096
tokens, fp16 activations, int8 weights prepared once, the d_ff hidden features sliced over the GPUs of one box.  Under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29573 tools/bench_megatron.py

Per forward of the WHOLE FFN (max over ranks, CUDA events, 3 warm-up + 10 timed):
  fused       MegatronFFN, exchange carried by the kernels (scattering GEMM epilogue over NVLink + ordered reduce, result on every rank)
  fused_rs    the same without the final all-gather (result stays sharded by columns: reduce-scatter only)
  collective  the same arithmetic with all_to_all_single + all_gather (torch.distributed / NCCL)
  colpar      the round-1 form: two column-parallel layers, each all-gathering its output in the GEMM epilogue
  local       this rank's compute alone (fc1 slice, quantizer, fc2 slice into local slots), no exchange
Rank 0 prints one JSON line and writes gpurun_out/megatron_P.json.  --small runs a reduced shape (bring-up)."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
colpar = importlib.import_module(qg.__name__ + ".colpar")
mg = importlib.import_module(qg.__name__ + ".megatron")
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.device("cuda", lr)
small = "--small" in sys.argv
only_fused = "--only-fused" in sys.argv  # skip the collective / column-parallel / local legs (already measured): fused exchange and its row-block variants only
T, D, F = (1024, 1024, 4096) if small else (4096, 9216, 36864)
part_dts = [torch.float32, torch.bfloat16]


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


g = torch.Generator(device=dev).manual_seed(7)  # the same full weights on every rank, sliced by the layers
W1 = (torch.randn((D, F), device=dev, generator=g) * 0.02).to(torch.float16)
W2 = (torch.randn((F, D), device=dev, generator=g) * 0.02).to(torch.float16)
b1 = torch.randn(F, device=dev, generator=g) * 0.1
b2 = torch.randn(D, device=dev, generator=g) * 0.1
x = torch.randn((T, D), device=dev, generator=g).to(torch.float16)
ops = 2.0 * T * D * F * 2
res = {"layer": "opt66b_ffn" if not small else "small_ffn", "world": world, "T": T, "d_model": D, "d_ff": F}
for pdt in part_dts:
    tag = {torch.float32: "f32", torch.bfloat16: "bf16"}[pdt]
    kw = dict(h_dtype=torch.float16, part_dtype=pdt, out_dtype=torch.float16)
    fused = mg.MegatronFFN(W1, b1, W2, b2, rank, world, exchange="fused", gather=True, **kw)
    coll = mg.MegatronFFN(W1, b1, W2, b2, rank, world, exchange="collective", gather=True, **kw)
    y_f = fused.forward(x).clone()
    y_c = coll.forward(x)
    torch.cuda.synchronize()
    res[f"fused_eq_collective_bits_{tag}"] = bool(torch.equal(y_f.view(torch.int16), y_c.view(torch.int16)))
    res[f"fused_us_{tag}"] = timed(lambda: fused.forward(x))
    res[f"fused_row_blocks_{tag}"] = len(fused._row_blocks(T))
    if not only_fused:
        res[f"collective_us_{tag}"] = timed(lambda: coll.forward(x))
        res[f"local_us_{tag}"] = timed(lambda: coll._partial_blocks(x))
    if world > 1:
        for c in (1, 2):
            rs = mg.MegatronFFN(W1, b1, W2, b2, rank, world, exchange="fused", gather=False, chunks=c, **kw)
            res[f"fused_rs_us_{tag}_chunks{c}"] = timed(lambda: rs.forward(x))
            del rs
        # row-block pipelining of the exchange tail, with 0 / 8 / 16 SMs kept free of the GEMMs for the side stream's kernels
        # row-block pipelining of the exchange tail; the gather by peer stores from the reduce kernel or by the copy engines
        for c, eng in ((1, "kernel"), (1, "copy"), (2, "kernel"), (2, "copy"), (4, "kernel"), (4, "copy")):
            fc = mg.MegatronFFN(W1, b1, W2, b2, rank, world, exchange="fused", gather=True, chunks=c, gather_engine=eng, **kw)
            res[f"fused_us_{tag}_chunks{c}_{eng}"] = timed(lambda: fc.forward(x))
            del fc
            torch.cuda.empty_cache()
    res[f"fused_tops_total_{tag}"] = ops / res[f"fused_us_{tag}"] / 1e6
    res[f"exchange_bytes_out_per_rank_{tag}"] = T * fused.bc * (world - 1) * torch.empty(0, dtype=pdt).element_size()
    del fused, coll, y_f, y_c
    torch.cuda.empty_cache()
    dist.barrier()
if world > 1 and not only_fused:
    l1 = colpar.FusedColumnParallelLinear(W1, b1, rank, world)
    l2 = colpar.FusedColumnParallelLinear(W2, b2, rank, world)

    def two_layers():
        h = l1.forward(x)
        h.relu_()
        return l2.forward(h)

    res["colpar_us"] = timed(two_layers)
    del l1, l2
res["clocks_note"] = "max over ranks of CUDA-event time per forward"
if rank == 0:
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in res.items()}), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/megatron_{world}{'_small' if small else ''}{'_fused' if only_fused else ''}.json", "w"), indent=1)
dist.destroy_process_group()
