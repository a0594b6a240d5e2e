"""CPU tests (gloo, world_size 2 and 3) of the column-parallel host logic: shard bounds, ragged
gather layout, and bit-identity of the sharded result with the single-rank result.  The per-rank
compute is the CPU oracle here (there is no GPU in this container); on the GPU box the same class
runs the C-ABI path (tests/test_gpu_multi.py, bench.py --gpus N)."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "quantized-gemm-for-transformer-inference_b200"


def colpar():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    return importlib.import_module(PKG + ".colpar")


def test_shard_bounds_cover_and_align():
    cp = colpar()
    for n, world, align in [(4096, 8, 16), (36864, 8, 16), (100, 3, 1), (100, 3, 16), (7, 8, 1), (48, 2, 16), (9216, 4, 16)]:
        spans = [cp.shard_bounds(n, world, r, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b
        for lo, hi in spans:
            assert lo % align == 0 or lo == n
        widths = [hi - lo for lo, hi in spans]
        units = -(-n // align)
        assert sum(widths) == n and max(widths) <= -(-units // world) * align


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, align, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        import oracle

        cp = importlib.import_module(PKG + ".colpar")
        M, N, K = shape
        rng = np.random.default_rng(42)  # same data on every rank: X replicated, W sharded from the full matrix
        X = rng.random((M, K), dtype=np.float32) * 2 - 1
        W = rng.random((K, N), dtype=np.float32) * 2 - 1
        b = rng.standard_normal(N).astype(np.float32)

        def compute(x, w, bias):
            return torch.from_numpy(oracle.quantized_mm(x.numpy(), w.numpy(), 127.0, bias=None if bias is None else bias.numpy()))

        layer = cp.ColumnParallelLinear(torch.from_numpy(W), torch.from_numpy(b), rank, world, compute=compute, align=align)
        y = layer.forward(torch.from_numpy(X))
        full = oracle.quantized_mm(X, W, 127.0, bias=b)
        ok = np.array_equal(y.numpy().view(np.int32), full.view(np.int32))
        ret[rank] = bool(ok) and tuple(y.shape) == (M, N)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,align", [(2, (33, 96, 70), 16), (2, (16, 50, 40), 1), (3, (20, 100, 64), 16)])
def test_sharded_result_is_bit_identical(world, shape, align):
    import oracle

    oracle.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, align, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_bench_reference_arm_contract(tmp_path):
    """`bench.py --impl reference` (the CPU arm): one JSON line with the contract's keys; under torchrun only
    rank 0 prints and the other ranks exit 0 without work."""
    import json
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    small = ["--impl", "reference", "--steps", "1", "--warmup", "0", "-m", "256", "-n", "128", "-k", "192", "-s", "3"]
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *small], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "TOPS" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["extrapolated"] is False and "seed 3" in line["data"]
    # both arms print the SAME config object for the same command line (the driver compares them)
    sys.path.insert(0, root)
    import bench

    assert line["config"] == bench.bench_config(256, 128, 192, 1)
    assert (line["config"]["M"], line["config"]["N"], line["config"]["K"]) == (256, 128, 192)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29613", os.path.join(root, "bench.py"), "--gpus", "2", *small],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line2 = json.loads(lines[0])
    assert line2["n_gpus"] == 2 and line2["config"] == bench.bench_config(256, 128, 192, 2)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must still use the cores it may use
    assert line2["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
