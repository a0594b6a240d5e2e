"""Performance experiments on the GPU box (each in its own subprocess so environment switches such
as QG_PDL / QG_COLS_TWO_PASS / QG_GEMM_VARIANT take effect).  Writes gpurun_out/perf.json.

    python tools/gpu_perf.py [--only name1,name2]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def pkg():
    return importlib.import_module("quantized-gemm-for-transformer-inference_b200")


def bench(fn, iters=30, warm=5):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def gemm_stats(variant, size, out="f32", kmajor=True):
    """Pipeline wait counters of the tcgen05 GEMM (cycles, averaged over CTAs)."""
    import torch

    qg = pkg()
    M = N = K = size
    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    B = torch.randint(-127, 128, (K, N), dtype=torch.int8, device="cuda")
    Cx, Cw = torch.rand(M, device="cuda"), torch.rand(N, device="cuda")
    dt = {"f32": torch.float32, "f16": torch.float16, "s32": torch.int32}[out]
    O = torch.empty((M, N), dtype=dt, device="cuda")
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))

    Bt = B.t().contiguous()

    def run():
        if kmajor:
            qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O)
        elif out == "s32":
            qg.op_mm(A, B, O)
        else:
            qg.gemm_s8_dequant(A, B, Cx, Cw, O)

    us = bench(run)
    stats = torch.zeros((148, 8), dtype=torch.int64, device="cuda")
    qg.lib().qg_debug_gemm_stats(C.c_void_p(stats.data_ptr()))
    run()
    torch.cuda.synchronize()
    us_stats = bench(run, iters=5, warm=1)
    qg.lib().qg_debug_gemm_stats(None)
    s = stats.cpu().double()
    used = s[:, 4] > 0 if variant == "TC_1SM" else s[:, 1] > 0
    names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_total"]
    res = {"us": us, "us_with_stats": us_stats, "tops": 2.0 * M * N * K / us / 1e6}
    # end-of-work wall clock per CTA (ns) and its start (end - loop cycles / 1.965 GHz): skew across the grid
    end_ns = s[:, 7]
    live = end_ns > 0
    if live.any():
        start_ns = end_ns - s[:, 6] / 1.965
        res["end_spread_us"] = float((end_ns[live].max() - end_ns[live].min()) / 1e3)
        res["start_spread_us"] = float((start_ns[live].max() - start_ns[live].min()) / 1e3)
        res["first_start_to_last_end_us"] = float((end_ns[live].max() - start_ns[live].min()) / 1e3)
        # per-cluster finish offsets relative to the first finisher (us), leader CTAs only, in cluster order
        raw = stats.cpu()[:, 7]
        lead_ns = raw[::2] if variant == "TC_2SM" else raw
        lead_ns = lead_ns[lead_ns > 0]
        res["end_offsets_us"] = [round(float(v - lead_ns.min()) / 1e3, 1) for v in lead_ns]
    lead = s[s[:, 4] > 0]
    for i, n in enumerate(names):
        col = s[:, i]
        nz = col[col > 0] if n.startswith("mma") else col[used]
        res[n] = float(nz.mean()) if nz.numel() else 0.0
    res["mma_total_max"] = float(lead[:, 4].max()) if lead.numel() else 0.0
    res["mma_total_min"] = float(lead[:, 4].min()) if lead.numel() else 0.0
    return res


def gemm_kmajor(cg, size):
    """B supplied as [N,K] (K-major operand) vs the reference's [K,N] (MN-major operand)."""
    import torch

    qg = pkg()
    M = N = K = size
    A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    B = torch.randint(-127, 128, (K, N), dtype=torch.int8, device="cuda")
    Bt = B.t().contiguous()
    O = torch.empty((M, N), dtype=torch.int32, device="cuda")
    qg.set_gemm_variant(qg.GEMM_TC_2SM if cg == 2 else qg.GEMM_TC_1SM)
    us_mn = bench(lambda: qg.op_mm(A, B, O))
    ref = O.clone()

    def km():
        rc = qg.lib().qg_test_gemm_s8_bt(cg, C.c_void_p(A.data_ptr()), C.c_int64(K), C.c_void_p(Bt.data_ptr()), C.c_int64(K),
                                         M, N, K, C.c_void_p(O.data_ptr()), C.c_int64(N), None)
        assert rc == 0

    us_k = bench(km)
    torch.cuda.synchronize()
    ops = 2.0 * M * N * K
    return {"mn_major_us": us_mn, "k_major_us": us_k, "mn_tops": ops / us_mn / 1e6, "k_tops": ops / us_k / 1e6,
            "same": bool(torch.equal(ref, O))}


def quantizers(size, dt="f32"):
    import torch

    qg = pkg()
    M = N = K = size
    tdt = {"f32": torch.float32, "f16": torch.float16}[dt]
    es = 4 if dt == "f32" else 2
    X = (torch.rand((M, K), device="cuda") * 2 - 1).to(tdt)
    W = (torch.rand((K, N), device="cuda") * 2 - 1).to(tdt)
    X2 = X.clone()
    W2 = W.clone()
    Xq = torch.empty((M, K), dtype=torch.int8, device="cuda")
    Wq = torch.empty((K, N), dtype=torch.int8, device="cuda")
    Cx, Cw = torch.empty(M, device="cuda"), torch.empty(N, device="cuda")
    flip = [0]

    def rows():
        flip[0] ^= 1
        qg.absmax_quant_rows(X if flip[0] else X2, 127.0, 0, Xq, Cx)

    def cols():
        flip[0] ^= 1
        qg.absmax_quant_cols(W if flip[0] else W2, 127.0, 0, Wq, Cw)

    Wt = torch.empty((N, K), dtype=torch.int8, device="cuda")

    def cols_t():
        flip[0] ^= 1
        qg.prepare_weights(W if flip[0] else W2, 127.0, 0, Wt, Cw)

    r, c, ct = bench(rows), bench(cols), bench(cols_t)
    return {"rows_us": r, "rows_gbs": (M * K * (es + 1) + 4 * M) / r / 1e3, "cols_us": c,
            "cols_gbs": (K * N * (es + 1) + 4 * N) / c / 1e3, "cols_t_us": ct, "cols_t_gbs": (K * N * (es + 1) + 4 * N) / ct / 1e3}


def cols_pipe(size, dt="f32", nbuf=3):
    """Column quantizer alone on rotating weight buffers (cold W) + a digest of its outputs for cross-configuration parity."""
    import hashlib

    import torch

    qg = pkg()
    K = N = size
    tdt = {"f32": torch.float32, "f16": torch.float16}[dt]
    es = 4 if dt == "f32" else 2
    torch.manual_seed(0)
    Ws = [(torch.rand((K, N), device="cuda") * 2 - 1).to(tdt) for _ in range(nbuf)]
    Wq = torch.empty((K, N), dtype=torch.int8, device="cuda")
    Cw = torch.empty(N, device="cuda")
    i = [0]

    def cols():
        i[0] = (i[0] + 1) % nbuf
        qg.absmax_quant_cols(Ws[i[0]], 127.0, 0, Wq, Cw)

    us = bench(cols, iters=60, warm=10)
    qg.absmax_quant_cols(Ws[0], 127.0, 0, Wq, Cw)
    torch.cuda.synchronize()
    h = hashlib.sha256(Wq.cpu().numpy().tobytes() + Cw.cpu().numpy().tobytes()).hexdigest()[:16]
    # same call captured in a CUDA graph and replayed on changing weights (the scratch must restore itself)
    g = torch.cuda.CUDAGraph()
    Wg = Ws[1].clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        qg.absmax_quant_cols(Wg, 127.0, 0, Wq, Cw)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            qg.absmax_quant_cols(Wg, 127.0, 0, Wq, Cw)
    Wg.copy_(Ws[0] * 0.5)
    g.replay()
    torch.cuda.synchronize()
    hg = hashlib.sha256(Wq.cpu().numpy().tobytes() + Cw.cpu().numpy().tobytes()).hexdigest()[:16]
    qg.absmax_quant_cols(Ws[0] * 0.5, 127.0, 0, Wq, Cw)
    torch.cuda.synchronize()
    hg_ref = hashlib.sha256(Wq.cpu().numpy().tobytes() + Cw.cpu().numpy().tobytes()).hexdigest()[:16]
    return {"cols_us": us, "cols_gbs": (K * N * (es + 1) + 4 * N) / us / 1e3, "digest": h, "graph_replay_ok": hg == hg_ref}


def full_op(size, out="f32"):
    import torch

    qg = pkg()
    M = N = K = size
    X = [torch.rand((M, K), device="cuda") * 2 - 1 for _ in range(2)]
    W = [torch.rand((K, N), device="cuda") * 2 - 1 for _ in range(2)]
    O = torch.empty((M, N), dtype=torch.float32 if out == "f32" else torch.float16, device="cuda")
    ws = torch.empty(qg.workspace_bytes(M, N, K), dtype=torch.uint8, device="cuda")
    flip = [0]

    def run():
        flip[0] ^= 1
        qg.op_quantized_mm(X[flip[0]], W[flip[0]], O, 127.0, workspace=ws)

    us = bench(run)
    # the same three calls captured in a CUDA graph
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        run()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            run()
            run()
    torch.cuda.synchronize()
    us_graph = bench(lambda: g.replay()) / 2
    lin = qg.LinearLayer(K, N)
    lin.init_uniform()
    lin.quantize_weights()
    us_lin = bench(lambda: lin.forward(X[0], O if out == "f32" else O))
    ops = 2.0 * M * N * K
    return {"full_us": us, "full_tops": ops / us / 1e6, "full_graph_us": us_graph, "full_graph_tops": ops / us_graph / 1e6,
            "linear_cached_w_us": us_lin, "linear_cached_w_tops": ops / us_lin / 1e6}


def lib_ref(size):
    import torch

    M = N = K = size
    a8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    b8 = torch.randint(-127, 128, (K, N), dtype=torch.int8, device="cuda")
    a16 = torch.randn((M, K), dtype=torch.float16, device="cuda")
    b16 = torch.randn((K, N), dtype=torch.float16, device="cuda")
    u8, u16 = bench(lambda: torch._int_mm(a8, b8)), bench(lambda: a16 @ b16)
    ops = 2.0 * M * N * K
    return {"cublaslt_int8_us": u8, "cublaslt_int8_tops": ops / u8 / 1e6, "cublas_fp16_us": u16, "cublas_fp16_tflops": ops / u16 / 1e6}


def gemm_inop(size, out="f32"):
    """The GEMM launch bracketed by CUDA events INSIDE the op sequence (rows quantizer, columns quantizer, GEMM),
    as bench.py's instrumented pass times it: start-to-end of one launch, not a back-to-back average."""
    import statistics

    import torch

    qg = pkg()
    M = N = K = size
    X = [torch.rand((M, K), device="cuda") * 2 - 1 for _ in range(2)]
    W = [torch.rand((K, N), device="cuda") * 2 - 1 for _ in range(2)]
    O = [torch.empty((M, N), dtype=torch.float32 if out == "f32" else torch.float16, device="cuda") for _ in range(2)]
    Xq = torch.empty((M, K), dtype=torch.int8, device="cuda")
    Wq = torch.empty((K, N), dtype=torch.int8, device="cuda")
    Cx, Cw = torch.empty(M, device="cuda"), torch.empty(N, device="cuda")
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(30)]
    for i in range(35):
        e = evs[max(0, i - 5)]
        e[0].record()
        qg.absmax_quant_rows(X[i & 1], 127.0, 0, Xq, Cx)
        e[1].record()
        qg.absmax_quant_cols(W[i & 1], 127.0, 0, Wq, Cw)
        e[2].record()
        qg.gemm_s8_dequant(Xq, Wq, Cx, Cw, O[i & 1], 127.0)
        e[3].record()
    torch.cuda.synchronize()
    med = lambda a, b: statistics.median(e[a].elapsed_time(e[b]) for e in evs) * 1e3
    return {"rows_us": med(0, 1), "cols_us": med(1, 2), "gemm_us": med(2, 3), "total_us": med(0, 3),
            "gemm_tops": 2.0 * M * N * K / med(2, 3) / 1e6}


EXPERIMENTS = {
    # name: (function, args, env)
    "lib_4096": (lib_ref, (4096,), {}),
    "lib_8192": (lib_ref, (8192,), {}),
    "stats_1sm_4096_f32": (gemm_stats, ("TC_1SM", 4096, "f32"), {}),
    "stats_2sm_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {}),
    "stats_2sm_4096_f32_nosplit": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_NO_TAIL_SPLIT": "1"}),
    "stats_1sm_4096_f32_nosplit": (gemm_stats, ("TC_1SM", 4096, "f32"), {"QG_NO_TAIL_SPLIT": "1"}),
    "stats_2sm_4096_s32": (gemm_stats, ("TC_2SM", 4096, "s32"), {}),
    "stats_2sm_4096_f16": (gemm_stats, ("TC_2SM", 4096, "f16"), {}),
    "stats_2sm_4096_f32_mnmajor": (gemm_stats, ("TC_2SM", 4096, "f32", False), {}),
    "stats_1sm_8192_f32": (gemm_stats, ("TC_1SM", 8192, "f32"), {}),
    "stats_2sm_8192_f32": (gemm_stats, ("TC_2SM", 8192, "f32"), {}),
    "stats_2sm_8192_f16": (gemm_stats, ("TC_2SM", 8192, "f16"), {}),
    "stats_2sm_2048_f32": (gemm_stats, ("TC_2SM", 2048, "f32"), {}),
    "stats_1sm_2048_f32": (gemm_stats, ("TC_1SM", 2048, "f32"), {}),
    "stats_2sm_1024_f32": (gemm_stats, ("TC_2SM", 1024, "f32"), {}),
    "stats_1sm_1024_f32": (gemm_stats, ("TC_1SM", 1024, "f32"), {}),
    "r2_stats_2sm_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {}),
    "r2_stats_2sm_4096_f32_noring": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_NO_LAST_RING": "1"}),
    "r2_stats_2sm_4096_f32_noepi": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1"}),
    "r2_stats_2sm_4096_f32_noepi_noload": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1", "QG_DBG_NOLOAD": "1"}),
    "r2_stats_2sm_4096_f32_mn": (gemm_stats, ("TC_2SM", 4096, "f32", False), {}),
    "r2_stats_2sm_4096_f32_mn_noring": (gemm_stats, ("TC_2SM", 4096, "f32", False), {"QG_NO_LAST_RING": "1"}),
    "r2_stats_2sm_8192_f32": (gemm_stats, ("TC_2SM", 8192, "f32"), {}),
    "r2_stats_2sm_8192_f32_noepi": (gemm_stats, ("TC_2SM", 8192, "f32"), {"QG_DBG_NOEPI": "1"}),
    "r2_stats_2sm_2048_f32": (gemm_stats, ("TC_2SM", 2048, "f32"), {}),
    "r2_clusters74_noepi_noload": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1", "QG_DBG_NOLOAD": "1", "QG_NO_TAIL_SPLIT": "1"}),
    "r2_clusters64_noepi_noload": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1", "QG_DBG_NOLOAD": "1", "QG_DBG_MAX_CLUSTERS": "64", "QG_NO_TAIL_SPLIT": "1"}),
    "r2_clusters32_noepi_noload": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1", "QG_DBG_NOLOAD": "1", "QG_DBG_MAX_CLUSTERS": "32", "QG_NO_TAIL_SPLIT": "1"}),
    "r2_clusters8_noepi_noload": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NOEPI": "1", "QG_DBG_NOLOAD": "1", "QG_DBG_MAX_CLUSTERS": "8", "QG_NO_TAIL_SPLIT": "1"}),
    "r2_clusters8_full": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_MAX_CLUSTERS": "8", "QG_NO_TAIL_SPLIT": "1"}),
    "r2_epi1_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_EPI_LEVEL": "1"}),
    "r2_epi1_4096_s32": (gemm_stats, ("TC_2SM", 4096, "s32"), {"QG_DBG_EPI_LEVEL": "1"}),
    "r2_epi2_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_EPI_LEVEL": "2"}),
    "r2_epi2_4096_s32": (gemm_stats, ("TC_2SM", 4096, "s32"), {"QG_DBG_EPI_LEVEL": "2"}),
    "r2_epi0_4096_s32": (gemm_stats, ("TC_2SM", 4096, "s32"), {}),
    "r2_epi0_4096_f16": (gemm_stats, ("TC_2SM", 4096, "f16"), {}),
    "r2_epi0_4096_f32_nostore_tma": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_DBG_NO_TMA_STORE": "1"}),
    "r2_hint0_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {}),
    "r2_hint1_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_STORE_HINT": "1"}),
    "r2_hint2_4096_f32": (gemm_stats, ("TC_2SM", 4096, "f32"), {"QG_STORE_HINT": "2"}),
    "r2_hint1_8192_f32": (gemm_stats, ("TC_2SM", 8192, "f32"), {"QG_STORE_HINT": "1"}),
    "r2_hint0_8192_f32": (gemm_stats, ("TC_2SM", 8192, "f32"), {}),
    "r2_inop_4096": (gemm_inop, (4096,), {}),
    "r2_inop_4096_noring": (gemm_inop, (4096,), {"QG_NO_LAST_RING": "1"}),
    "r2_inop_4096_f16": (gemm_inop, (4096, "f16"), {}),
    "r2_inop_8192": (gemm_inop, (8192,), {}),
    "r2_cpipe_off_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0"}),
    "r2_cpipe_def_4096": (cols_pipe, (4096,), {}),
    "r2_cpipe_1_64_5of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,5,8,4,128"}),
    "r2_cpipe_1_64_1of2_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,1,2,4,128"}),
    "r2_cpipe_1_64_3of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,3,8,4,128"}),
    "r2_cpipe_1_64_3of4_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,3,4,4,128"}),
    "r2_cpipe_1_64_5of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,5,8,4,256"}),
    "r2_cpipe_1_64_1of2_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,1,2,4,256"}),
    "r2_cpipe_1_64_3of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,3,8,4,256"}),
    "r2_cpipe_1_64_3of4_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "1,64,3,4,4,256"}),
    "r2_cpipe_2_64_5of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,5,8,4,128"}),
    "r2_cpipe_2_64_1of2_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,1,2,4,128"}),
    "r2_cpipe_2_64_3of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,3,8,4,128"}),
    "r2_cpipe_2_64_3of4_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,3,4,4,128"}),
    "r2_cpipe_2_64_5of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,5,8,4,256"}),
    "r2_cpipe_2_64_1of2_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,1,2,4,256"}),
    "r2_cpipe_2_64_3of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,3,8,4,256"}),
    "r2_cpipe_2_64_3of4_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "2,64,3,4,4,256"}),
    "r2_cpipe_4_64_5of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,5,8,4,128"}),
    "r2_cpipe_4_64_1of2_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,1,2,4,128"}),
    "r2_cpipe_4_64_3of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,3,8,4,128"}),
    "r2_cpipe_4_64_3of4_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,3,4,4,128"}),
    "r2_cpipe_4_64_5of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,5,8,4,256"}),
    "r2_cpipe_4_64_1of2_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,1,2,4,256"}),
    "r2_cpipe_4_64_3of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,3,8,4,256"}),
    "r2_cpipe_4_64_3of4_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,3,4,4,256"}),
    "r2_cpipe_8_64_5of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,5,8,4,128"}),
    "r2_cpipe_8_64_1of2_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,1,2,4,128"}),
    "r2_cpipe_8_64_3of8_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,3,8,4,128"}),
    "r2_cpipe_8_64_3of4_p4_128_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,3,4,4,128"}),
    "r2_cpipe_8_64_5of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,5,8,4,256"}),
    "r2_cpipe_8_64_1of2_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,1,2,4,256"}),
    "r2_cpipe_8_64_3of8_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,3,8,4,256"}),
    "r2_cpipe_8_64_3of4_p4_256_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "8,64,3,4,4,256"}),
    "r2_cdbg_readers_4_64_5of8": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,5,8,4,256", "QG_COLS_PIPE_DBG": "1"}),
    "r2_cdbg_writers_4_64_5of8": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,5,8,4,256", "QG_COLS_PIPE_DBG": "2"}),
    "r2_cdbg_readers_4_64_7of8": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,7,8,4,256", "QG_COLS_PIPE_DBG": "1"}),
    "r2_cdbg_writers_4_64_1of8": (cols_pipe, (4096,), {"QG_COLS_PIPE": "4,64,1,8,4,256", "QG_COLS_PIPE_DBG": "2"}),
    "r2_cwave2_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "2"}),
    "r2_cwave3_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "3"}),
    "r2_cwave4_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "4"}),
    "r2_cwave8_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "8"}),
    "r2_cwave0_4096": (cols_pipe, (4096,), {"QG_COLS_PIPE": "0"}),
    "r2_cwave2_8192": (cols_pipe, (8192,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "2"}),
    "r2_cwave8_8192": (cols_pipe, (8192,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "8"}),
    "r2_cwave16_8192": (cols_pipe, (8192,), {"QG_COLS_PIPE": "0", "QG_COLS_WAVE": "16"}),
    "r2_cwave0_8192": (cols_pipe, (8192,), {"QG_COLS_PIPE": "0"}),
    "r2_cpipe_off_4096_f16": (cols_pipe, (4096, "f16"), {"QG_COLS_PIPE": "0"}),
    "r2_cpipe_def_4096_f16": (cols_pipe, (4096, "f16"), {}),
    "r2_cpipe_off_8192": (cols_pipe, (8192,), {"QG_COLS_PIPE": "0"}),
    "r2_cpipe_def_8192": (cols_pipe, (8192,), {}),
    "r2_cpipe_off_2048": (cols_pipe, (2048,), {"QG_COLS_PIPE": "0"}),
    "r2_cpipe_def_2048": (cols_pipe, (2048,), {}),
    "quant_4096": (quantizers, (4096,), {}),
    "quant_4096_twopass": (quantizers, (4096,), {"QG_COLS_TWO_PASS": "1"}),
    "quant_4096_f16": (quantizers, (4096, "f16"), {}),
    "quant_8192": (quantizers, (8192,), {}),
    "quant_8192_twopass": (quantizers, (8192,), {"QG_COLS_TWO_PASS": "1"}),
    "full_4096_pdl": (full_op, (4096,), {}),
    "full_4096_nopdl": (full_op, (4096,), {"QG_PDL": "0"}),
    "full_4096_1sm": (full_op, (4096,), {"QG_GEMM_VARIANT": "2"}),
    "full_2048_pdl": (full_op, (2048,), {}),
    "full_2048_nopdl": (full_op, (2048,), {"QG_PDL": "0"}),
    "full_8192": (full_op, (8192,), {}),
    "full_1024": (full_op, (1024,), {}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    ap.add_argument("--only", default="")
    ap.add_argument("--out", default=os.path.join(OUT, "perf.json"))
    args = ap.parse_args()
    if args.one:
        fn, a, _ = EXPERIMENTS[args.one]
        print("PERF_RESULT " + json.dumps(fn(*a)))
        return
    os.makedirs(OUT, exist_ok=True)
    results = {}
    names = [n for n in EXPERIMENTS if not args.only or any(n.startswith(p) for p in args.only.split(","))]
    for name in names:
        env = dict(os.environ)
        env.update(EXPERIMENTS[name][2])
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], capture_output=True, text=True,
                               timeout=240, env=env)
            line = [l for l in r.stdout.splitlines() if l.startswith("PERF_RESULT ")]
            results[name] = json.loads(line[-1][12:]) if line else {"error": (r.stdout + r.stderr)[-1500:]}
        except subprocess.TimeoutExpired:
            results[name] = {"error": "timeout"}
        results[name]["secs"] = round(time.time() - t0, 1)
        print(name, json.dumps(results[name])[:500], flush=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
