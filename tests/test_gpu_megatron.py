"""Megatron pairing on the GPU, through the C ABI.

Single GPU: the scattering GEMM epilogue and the ordered reduction against the CPU oracle (the P ranks are played one after
the other on one device, each scattering into its own slot of a local buffer) -- bit-exact.  Two or more GPUs: MegatronFFN with
the exchange carried by the kernels over symmetric memory == the collective form == oracle.megatron_ffn, all bits."""
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TDT = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}


def _pkg():
    return importlib.import_module(PKG)


def _bits(t):
    return t.view(torch.int32 if t.dtype == torch.float32 else torch.int16)


def _data(M, d_in, d_ff, d_out, seed=3):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((M, d_in)).astype(np.float32)
    W1 = ((rng.random((d_in, d_ff), dtype=np.float32) * 2 - 1) / np.sqrt(d_in)).astype(np.float32)
    W2 = ((rng.random((d_ff, d_out), dtype=np.float32) * 2 - 1) / np.sqrt(d_ff)).astype(np.float32)
    b1 = rng.standard_normal(d_ff).astype(np.float32)
    b2 = rng.standard_normal(d_out).astype(np.float32)
    return X, W1, b1, W2, b2


@pytest.mark.parametrize("shape,P,dts", [
    ((256, 128, 512, 256), 2, ("f32", "f32", "f32")),
    ((300, 96, 640, 200), 3, ("f32", "f32", "f32")),       # ragged: d_out not a multiple of the block width
    ((384, 256, 1024, 1152), 4, ("f16", "f32", "f16")),    # OPT-66B-like block width (1152 / 4 = 288 -> 320-column blocks)
    ((130, 64, 256, 576), 2, ("f16", "bf16", "bf16")),     # 16-bit partial products
    ((512, 512, 2048, 512), 8, ("f32", "f32", "f32")),     # config 3's FFN over 8 slices
])
def test_rowpar_chain_single_device_matches_oracle(shape, P, dts):
    import oracle

    qg = _pkg()
    mg = importlib.import_module(PKG + ".megatron")
    cp = importlib.import_module(PKG + ".colpar")
    M, d_in, d_ff, d_out = shape
    h_dt, p_dt, o_dt = dts
    X, W1, b1, W2, b2 = _data(*shape)
    bounds = [cp.shard_bounds(d_ff, P, r, 16) for r in range(P)]
    bc = mg.block_cols_for(d_out, P)
    want, parts = oracle.megatron_ffn(X, W1, b1, W2, b2, bounds, h_dtype=h_dt, part_dtype=p_dt, out_dtype=o_dt, return_parts=True)
    Xd = torch.from_numpy(X).cuda()
    # owner-major slots: slots[b][p] = rank p's partial of block b
    slots = torch.full((P, P, M, bc), float("nan"), dtype=TDT[p_dt], device="cuda")
    es = slots.element_size()
    for p, (lo, hi) in enumerate(bounds):
        w1t, cw1 = qg.prepare_weights(torch.from_numpy(W1[:, lo:hi].copy()).cuda())
        w2t, cw2 = qg.prepare_weights(torch.from_numpy(W2[lo:hi, :].copy()).cuda())
        H = torch.empty((M, hi - lo), dtype=TDT[h_dt], device="cuda")
        ptrs = [slots[b, p].data_ptr() for b in range(P)]
        qg.ffn_forward_rowpar(Xd, w1t, cw1, torch.from_numpy(b1[lo:hi].copy()).cuda(), w2t, cw2, H, ptrs, bc, bc, TDT[p_dt], d_out)
        torch.cuda.synchronize()
        # the rank's partial product, reassembled from the owners' slots, equals the oracle's
        got = torch.cat([slots[b, p] for b in range(P)], dim=1)[:, :d_out]
        wantp = torch.from_numpy(parts[p]).to(TDT[p_dt]).cuda()
        assert torch.equal(_bits(got.contiguous()), _bits(wantp)), f"partial of rank {p}"
    out = torch.empty((M, d_out), dtype=TDT[o_dt], device="cuda")
    for b in range(P):
        lo, hi = min(b * bc, d_out), min((b + 1) * bc, d_out)
        if hi > lo:
            qg.reduce_partials(slots[b], torch.from_numpy(b2[lo:hi].copy()).cuda(), out[:, lo:hi], (), hi - lo)
    torch.cuda.synchronize()
    w = want if isinstance(want, torch.Tensor) else torch.from_numpy(np.asarray(want))
    assert torch.equal(_bits(out), _bits(w.to(TDT[o_dt]).cuda()))


def test_scatter_rejects_bad_blocks():
    qg = _pkg()
    Xq = torch.zeros((128, 64), dtype=torch.int8, device="cuda")
    Wt = torch.zeros((256, 64), dtype=torch.int8, device="cuda")
    Cx, Cw = torch.ones(128, device="cuda"), torch.ones(256, device="cuda")
    buf = torch.empty((2, 128, 144), device="cuda")
    with pytest.raises(qg.QGemmError):  # 144 is not a multiple of the 32 columns one fp32 store carries
        qg.gemm_s8_dequant_scatter(Xq, Wt, Cx, Cw, [buf[0].data_ptr(), buf[1].data_ptr()], 144, 144, torch.float32, 256)
    with pytest.raises(qg.QGemmError):  # two blocks of 96 columns do not cover 256
        qg.gemm_s8_dequant_scatter(Xq, Wt, Cx, Cw, [buf[0].data_ptr(), buf[1].data_ptr()], 96, 144, torch.float32, 256)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, dts, gather, ret, chunks=0, multicast=False, engine="auto"):
    import torch.distributed as dist

    if multicast:  # the result's gather through the NVSwitch multicast address (multimem.st in the reduce kernel)
        os.environ["QG_MULTICAST"] = "1"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import oracle

        mg = importlib.import_module(PKG + ".megatron")
        cp = importlib.import_module(PKG + ".colpar")
        M, d_in, d_ff, d_out = shape
        h_dt, p_dt, o_dt = dts
        X, W1, b1, W2, b2 = _data(*shape, seed=11)
        t = lambda a: torch.from_numpy(a).cuda()
        kw = dict(h_dtype=TDT[h_dt], part_dtype=TDT[p_dt], out_dtype=TDT[o_dt], gather=gather)
        fused = mg.MegatronFFN(t(W1), t(b1), t(W2), t(b2), rank, world, exchange="fused", chunks=chunks, gather_engine=engine, **kw)
        coll = mg.MegatronFFN(t(W1), t(b1), t(W2), t(b2), rank, world, exchange="collective", **kw)
        bounds = [cp.shard_bounds(d_ff, world, r, 16) for r in range(world)]
        want = oracle.megatron_ffn(X, W1, b1, W2, b2, bounds, h_dtype=h_dt, part_dtype=p_dt, out_dtype=o_dt)
        want = (want if isinstance(want, torch.Tensor) else torch.from_numpy(np.asarray(want))).to(TDT[o_dt]).cuda()
        if not gather:
            want = want[:, fused.olo:fused.ohi].contiguous()
        ok = True
        Xd = t(X)
        for it in range(3):  # repeated forwards reuse the symmetric slots
            yf = fused.forward(Xd)
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(_bits(yf.contiguous()), _bits(want)))
        yc = coll.forward(Xd)
        torch.cuda.synchronize()
        ok = ok and bool(torch.equal(_bits(yc.contiguous()), _bits(want)))
        ret[rank] = ok
        ret["mc_%d" % rank] = bool(getattr(fused, "out_mc", 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("shape,dts", [((512, 256, 1024, 512), ("f32", "f32", "f32")), ((768, 128, 768, 1152), ("f16", "bf16", "f16"))])
def test_megatron_ffn_multicast_gather_matches_oracle(shape, dts):
    """The all-gather half of the exchange as one multimem.st per 16 bytes (QG_MULTICAST=1): same bits.  Skipped when the
    fabric has no multicast mapping."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, dts, True, ret, 0, True), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
    if not all(ret.get("mc_%d" % r) for r in range(world)):
        pytest.skip("no multicast mapping on this box: the unicast gather ran (and matched)")


@pytest.mark.parametrize("chunks", [1, 2])
def test_megatron_ffn_copy_engine_gather_matches_oracle(chunks):
    """The gather of the reduced blocks by cudaMemcpy2DAsync (copy engines) instead of peer stores from the reduce kernel."""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), (1024, 256, 1024, 704), ("f16", "bf16", "f16"), True, ret, chunks, False, "copy"),
             nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


@pytest.mark.parametrize("shape,dts,gather,chunks", [
    ((512, 256, 1024, 512), ("f32", "f32", "f32"), True, 1),
    ((300, 128, 768, 200), ("f16", "f32", "f16"), True, 0),
    ((1024, 512, 2048, 1152), ("f16", "bf16", "f16"), False, 2),
    ((2304, 256, 1024, 512), ("f16", "f32", "f16"), True, 0),   # 4 row blocks, the last one ragged (2304 = 3 x 768 ... rounded to tiles)
    ((1100, 128, 512, 320), ("f32", "f32", "f32"), True, 3),
])
def test_megatron_ffn_fused_exchange_matches_oracle(shape, dts, gather, chunks):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import torch.multiprocessing as mp

    import oracle

    oracle.build()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), shape, dts, gather, ret, chunks), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)
