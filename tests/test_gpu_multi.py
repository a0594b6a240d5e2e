"""Multi-GPU test (needs >= 2 GPUs on the box; skipped otherwise): column-parallel quantized linear
over NCCL, one process per GPU -- the gathered result must equal the single-GPU op bit for bit."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "quantized-gemm-for-transformer-inference_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, dt, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sys.path.insert(0, ROOT)
        qg = importlib.import_module(PKG)
        cp = importlib.import_module(PKG + ".colpar")
        M, N, K = shape
        tdt = {"f32": torch.float32, "f16": torch.float16}[dt]
        g = torch.Generator().manual_seed(7)
        X = (torch.rand((M, K), generator=g) * 2 - 1).to(tdt).cuda()
        W = (torch.rand((K, N), generator=g) * 2 - 1).to(tdt).cuda()
        b = torch.randn(N, generator=g).cuda()
        layer = cp.ColumnParallelLinear(W, b, rank, world)
        y = layer.forward(X)
        full = torch.empty((M, N), dtype=tdt, device="cuda")
        qg.op_quantized_mm(X, W, full, 127.0, bias=b)
        torch.cuda.synchronize()
        ret[rank] = bool(torch.equal(y.view(torch.int32 if dt == "f32" else torch.int16),
                                     full.view(torch.int32 if dt == "f32" else torch.int16)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,dt", [((512, 1024, 768), "f32"), ((300, 2000, 520), "f16")])
def test_column_parallel_matches_single_gpu(shape, dt):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, dt, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def _worker_fused(rank, world, port, shape, dt, ret, multicast=None):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        sys.path.insert(0, ROOT)
        qg = importlib.import_module(PKG)
        cp = importlib.import_module(PKG + ".colpar")
        M, N, K = shape
        tdt = {"f32": torch.float32, "f16": torch.float16}[dt]
        g = torch.Generator().manual_seed(11)
        X = (torch.rand((M, K), generator=g) * 2 - 1).to(tdt).cuda()
        W = (torch.rand((K, N), generator=g) * 2 - 1).to(tdt).cuda()
        b = torch.randn(N, generator=g).cuda()
        layer = cp.FusedColumnParallelLinear(W, b, rank, world, multicast=multicast)
        ok = True
        for _ in range(3):  # repeated forwards reuse the symmetric buffer
            y = layer.forward(X)
            torch.cuda.synchronize()
            full = torch.empty((M, N), dtype=tdt, device="cuda")
            qg.op_quantized_mm(X, W, full, 127.0, bias=b)
            torch.cuda.synchronize()
            view = torch.int32 if dt == "f32" else torch.int16
            ok = ok and bool(torch.equal(y.view(view), full.view(view)))
        ret[rank] = ok
        ret["mc_%d" % rank] = bool(layer.mc_ptr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,dt", [((512, 1024, 768), "f32"), ((300, 2048, 520), "f16"), ((4096, 4096, 1024), "f16")])
def test_multicast_exchange_matches_single_gpu(shape, dt):
    """The gather done by the NVSwitch (multimem.st to the symmetric allocation's multicast address): same bits.  Skipped when
    the fabric offers no multicast mapping (the unicast form is then what FusedColumnParallelLinear runs)."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    ret = mp.Manager().dict()
    mp.spawn(_worker_fused, args=(world, _free_port(), shape, dt, ret, True), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
    if not all(ret.get("mc_%d" % r) for r in range(world)):
        pytest.skip("no multicast mapping on this box: the unicast path ran (and matched)")


@pytest.mark.parametrize("shape,dt", [((512, 1024, 768), "f32"), ((300, 2048, 520), "f16"), ((4096, 4096, 1024), "f16")])
def test_fused_epilogue_gather_matches_single_gpu(shape, dt):
    """Same check with the gather fused into the GEMM epilogue (peer TMA stores over NVLink)."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    ret = mp.Manager().dict()
    mp.spawn(_worker_fused, args=(world, _free_port(), shape, dt, ret, False), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
