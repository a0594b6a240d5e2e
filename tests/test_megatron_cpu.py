"""CPU tests (gloo, world_size 2 and 3) of the Megatron pairing's host logic: hidden-feature
slices, owner column blocks (ragged, padded), the ordered reduction and the final gather.  The per-rank compute is the
CPU oracle here; on the GPU box the same class runs the C-ABI path with the exchange carried by the kernels
(tests/test_gpu_multi.py)."""
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


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_block_cols_cover():
    sys.path.insert(0, ROOT)
    mg = importlib.import_module(PKG + ".megatron")
    for d_out, world in [(9216, 8), (4096, 8), (512, 2), (100, 3), (64, 8), (1, 2)]:
        bc = mg.block_cols_for(d_out, world)
        assert bc % 64 == 0 and bc * world >= d_out and (bc - 64) * world < d_out


def test_oracle_megatron_single_rank_equals_plain_ffn():
    """With one rank the pairing IS the reference's FFN chain: same bits as ll1 -> relu -> ll2 on the quantized path."""
    sys.path.insert(0, ROOT)
    import oracle

    oracle.build()
    rng = np.random.default_rng(5)
    X = rng.standard_normal((17, 40)).astype(np.float32)
    W1 = (rng.random((40, 72), dtype=np.float32) * 2 - 1) / 6
    W2 = (rng.random((72, 24), dtype=np.float32) * 2 - 1) / 8
    b1, b2 = rng.standard_normal(72).astype(np.float32), rng.standard_normal(24).astype(np.float32)
    y = oracle.megatron_ffn(X, W1, b1, W2, b2, [(0, 72)])
    ref = oracle.quantized_mm(oracle.relu(oracle.quantized_mm(X, W1, 127.0, bias=b1)), W2, 127.0, bias=b2)
    assert np.array_equal(y.view(np.int32), ref.view(np.int32))
    # two slices: close to, but not the same bits as, the single-GPU layer (per-slice scales)
    y2 = oracle.megatron_ffn(X, W1, b1, W2, b2, [(0, 32), (32, 72)])
    assert np.allclose(y2, ref, rtol=0, atol=0.05 * np.abs(ref).max())


def _worker(rank, world, port, shape, gather, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        import oracle

        mg = importlib.import_module(PKG + ".megatron")
        cp = importlib.import_module(PKG + ".colpar")
        M, d_in, d_ff, d_out = shape
        rng = np.random.default_rng(7)  # same data on every rank
        X = rng.standard_normal((M, d_in)).astype(np.float32)
        W1 = ((rng.random((d_in, d_ff), dtype=np.float32) * 2 - 1) / np.sqrt(d_in)).astype(np.float32)
        W2 = ((rng.random((d_ff, d_out), dtype=np.float32) * 2 - 1) / np.sqrt(d_ff)).astype(np.float32)
        b1, b2 = rng.standard_normal(d_ff).astype(np.float32), rng.standard_normal(d_out).astype(np.float32)

        def compute(x, w1, bb1, w2):
            h = oracle.relu(oracle.quantized_mm(x.numpy(), w1.numpy(), 127.0, bias=None if bb1 is None else bb1.numpy()))
            return torch.from_numpy(oracle.quantized_mm(h, w2.numpy(), 127.0))

        layer = mg.MegatronFFN(torch.from_numpy(W1), torch.from_numpy(b1), torch.from_numpy(W2), torch.from_numpy(b2), rank, world,
                               exchange="collective", gather=gather, compute=compute)
        y = layer.forward(torch.from_numpy(X)).numpy()
        bounds = [cp.shard_bounds(d_ff, world, r, 16) for r in range(world)]
        full = oracle.megatron_ffn(X, W1, b1, W2, b2, bounds)
        want = full if gather else full[:, layer.olo:layer.ohi]
        ret[rank] = bool(np.array_equal(y.view(np.int32), want.view(np.int32))) and y.shape == want.shape
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,gather", [(2, (19, 48, 96, 40), True), (3, (8, 32, 80, 200), True), (2, (5, 16, 64, 70), False)])
def test_megatron_host_logic_matches_oracle(world, shape, gather):
    import oracle

    oracle.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, gather, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
