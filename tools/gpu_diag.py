"""Bring-up diagnostics for the GPU box: runs each check in its own subprocess under a timeout so
a faulting kernel variant cannot take the others down, and writes gpurun_out/diag.json.

    python tools/gpu_diag.py            # all checks
    python tools/gpu_diag.py --one NAME # a single check, in-process (what the subprocesses run)
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


def _codes(rng, r, c):
    import numpy as np

    return rng.integers(-128, 128, (r, c), dtype=np.int8)


def gemm_case(variant, M, N, K, kmajor_hook=False, cg=1):
    import numpy as np
    import torch

    import oracle

    qg = pkg()
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A, B = _codes(rng, M, K), _codes(rng, K, N)
    dA = torch.from_numpy(A).cuda()
    out = torch.full((M, N), -7, dtype=torch.int32, device="cuda")
    if kmajor_hook:
        dBt = torch.from_numpy(np.ascontiguousarray(B.T)).cuda()
        rc = qg.lib().qg_test_gemm_s8_bt(cg, C.c_void_p(dA.data_ptr()), C.c_int64(K), C.c_void_p(dBt.data_ptr()),
                                         C.c_int64(K), M, N, K, C.c_void_p(out.data_ptr()), C.c_int64(N), None)
        assert rc == 0, qg.lib().qg_last_error()
    else:
        dB = torch.from_numpy(B).cuda()
        qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
        qg.op_mm(dA, dB, out)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    exp = oracle.gemm_s8s8s32(A, B)
    ok = bool(np.array_equal(got, exp))
    info = {"ok": ok}
    if not ok:
        eq = got == exp
        info["match_frac"] = float(eq.mean())
        info["rows_all_ok"] = int(eq.all(axis=1).sum())
        info["cols_all_ok"] = int(eq.all(axis=0).sum())
        info["untouched_frac"] = float((got == -7).mean())
        info["first_bad"] = [int(v) for v in np.argwhere(~eq)[0]]
        info["got_sample"] = got[:2, :8].tolist()
        info["exp_sample"] = exp[:2, :8].tolist()
        # column blocks of 8 that are fully right, to spot swizzle / chunk mix-ups
        cb = eq.all(axis=0).reshape(-1, 8).all(axis=1) if N % 8 == 0 else None
        if cb is not None:
            info["col8_ok"] = "".join("1" if v else "0" for v in cb.tolist())
    return info


def probe_mn(variant):
    """One-hot A: C[i,n] = B[i,n] for i < K.  With B encoding k (then n) the result shows which
    element of B the tensor core actually read for every (i, n)."""
    import numpy as np
    import torch

    qg = pkg()
    M, N, K = 128, 256, 128
    A = np.eye(M, K, dtype=np.int8)
    res = {}
    for name, B in (("k", np.repeat((np.arange(K) - 64).astype(np.int8)[:, None], N, 1)),
                    ("n", np.repeat(((np.arange(N) % 256) - 128).astype(np.int8)[None, :], K, 0))):
        out = torch.zeros((M, N), dtype=torch.int32, device="cuda")
        qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
        qg.op_mm(torch.from_numpy(A).cuda(), torch.from_numpy(np.ascontiguousarray(B)).cuda(), out)
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        res[name + "_ok"] = bool(np.array_equal(got, B.astype(np.int32)))
        res[name + "_rows0_3_cols0_40"] = (got[:4, :40] + (64 if name == "k" else 128)).tolist()
        res[name + "_col0_rows0_40"] = (got[:40, 0] + (64 if name == "k" else 128)).tolist()
        res[name + "_row1_cols120_140"] = (got[1, 120:140] + (64 if name == "k" else 128)).tolist()
    return res


def quant_case():
    import numpy as np
    import torch

    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import make_edge_matrix

    qg = pkg()
    rng = np.random.default_rng(0)
    res = {}
    for shape in [(70, 4096), (33, 1000), (9, 16384), (40, 512)]:
        X = make_edge_matrix(rng, *shape)
        Xq, Cx = qg.absmax_quant_rows(torch.from_numpy(X).cuda())
        eq, ecx = oracle.absmax_quant_rows(X)
        res[f"rows{shape}"] = bool(np.array_equal(Xq.cpu().numpy(), eq)) and bool(
            np.array_equal(np.nan_to_num(Cx.cpu().numpy(), nan=7.0), np.nan_to_num(ecx, nan=7.0)))
        W = np.ascontiguousarray(X.T)
        Wq, Cw = qg.absmax_quant_cols(torch.from_numpy(W).cuda())
        eq, ecw = oracle.absmax_quant_cols(W)
        res[f"cols{shape[::-1]}"] = bool(np.array_equal(Wq.cpu().numpy(), eq)) and bool(
            np.array_equal(np.nan_to_num(Cw.cpu().numpy(), nan=7.0), np.nan_to_num(ecw, nan=7.0)))
    res["ok"] = all(res.values())
    return res


def full_case(variant, M, N, K, dt="f32"):
    import numpy as np
    import torch

    import oracle

    qg = pkg()
    rng = np.random.default_rng(1)
    X = rng.random((M, K), dtype=np.float32) * 2 - 1
    W = rng.random((K, N), dtype=np.float32) * 2 - 1
    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[dt]
    dX, dW = torch.from_numpy(X).cuda().to(tdt), torch.from_numpy(W).cuda().to(tdt)
    bias = rng.standard_normal(N).astype(np.float32)
    O = torch.empty((M, N), dtype=tdt, device="cuda")
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    qg.op_quantized_mm(dX, dW, O, 127.0, bias=torch.from_numpy(bias).cuda())
    torch.cuda.synchronize()
    exp = oracle.quantized_mm(dX.float().cpu().numpy(), dW.float().cpu().numpy(), bias=bias)
    expt = torch.from_numpy(exp).to(tdt)
    ok = bool(torch.equal(O.cpu(), expt))
    info = {"ok": ok}
    if not ok:
        d = (O.cpu().float() - expt.float()).abs()
        info["max_abs_diff"] = float(d.max())
        info["mismatch_frac"] = float((d > 0).float().mean())
    return info


def timing(variant, M, N, K):
    import torch

    qg = pkg()
    dA = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    dB = torch.randint(-127, 128, (K, N), dtype=torch.int8, device="cuda")
    out = torch.empty((M, N), dtype=torch.int32, device="cuda")
    Cx = torch.rand(M, device="cuda")
    Cw = torch.rand(N, device="cuda")
    of = torch.empty((M, N), dtype=torch.float32, device="cuda")
    oh = torch.empty((M, N), dtype=torch.float16, device="cuda")
    res = {}

    def bench(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    ops = 2.0 * M * N * K
    if variant == "LIB":
        ms = bench(lambda: torch._int_mm(dA, dB))
        res["torch_int_mm_ms"] = ms
        res["torch_int_mm_tops"] = ops / ms / 1e9
        a16, b16 = torch.randn((M, K), dtype=torch.float16, device="cuda"), torch.randn((K, N), dtype=torch.float16, device="cuda")
        ms = bench(lambda: torch.matmul(a16, b16))
        res["torch_fp16_ms"] = ms
        res["torch_fp16_tflops"] = ops / ms / 1e9
        res["ok"] = True
        return res
    qg.set_gemm_variant(getattr(qg, "GEMM_" + variant))
    ms = bench(lambda: qg.op_mm(dA, dB, out))
    res["s32_ms"], res["s32_tops"] = ms, ops / ms / 1e9
    ms = bench(lambda: qg.gemm_s8_dequant(dA, dB, Cx, Cw, of))
    res["f32_ms"], res["f32_tops"] = ms, ops / ms / 1e9
    ms = bench(lambda: qg.gemm_s8_dequant(dA, dB, Cx, Cw, oh))
    res["f16_ms"], res["f16_tops"] = ms, ops / ms / 1e9
    X = torch.rand((M, K), device="cuda") * 2 - 1
    W = torch.rand((K, N), device="cuda") * 2 - 1
    ms = bench(lambda: qg.absmax_quant_rows(X))
    res["quant_rows_ms"], res["quant_rows_gbs"] = ms, (M * K * 5 + 4 * M) / ms / 1e6
    ms = bench(lambda: qg.absmax_quant_cols(W))
    res["quant_cols_ms"], res["quant_cols_gbs"] = ms, (N * K * 5 + 4 * N) / ms / 1e6
    ms = bench(lambda: qg.op_quantized_mm(X, W, of))
    res["full_ms"], res["full_tops"] = ms, ops / ms / 1e9
    res["ok"] = True
    return res


CHECKS = {
    "device": lambda: dict(zip(("sm", "major", "minor"), pkg().device_info()), ok=True),
    "quantizers": quant_case,
    "simt_small": lambda: gemm_case("SIMT", 100, 70, 33),
    "tc1_kmajor_128": lambda: gemm_case("TC_1SM", 128, 256, 128, kmajor_hook=True, cg=1),
    "tc1_kmajor_big": lambda: gemm_case("TC_1SM", 512, 1024, 640, kmajor_hook=True, cg=1),
    "tc1_mn_128": lambda: gemm_case("TC_1SM", 128, 256, 128),
    "tc1_mn_k32": lambda: gemm_case("TC_1SM", 128, 256, 32),
    "tc1_mn_n128": lambda: gemm_case("TC_1SM", 128, 128, 128),
    "tc1_mn_big": lambda: gemm_case("TC_1SM", 512, 1024, 640),
    "tc1_mn_odd": lambda: gemm_case("TC_1SM", 200, 304, 1008),
    "tc1_mn_probe": lambda: probe_mn("TC_1SM"),
    "tc2_kmajor_256": lambda: gemm_case("TC_2SM", 256, 256, 128, kmajor_hook=True, cg=2),
    "tc2_kmajor_big": lambda: gemm_case("TC_2SM", 512, 1024, 640, kmajor_hook=True, cg=2),
    "tc2_mn_256": lambda: gemm_case("TC_2SM", 256, 256, 128),
    "tc2_mn_big": lambda: gemm_case("TC_2SM", 512, 1024, 640),
    "tc2_mn_odd": lambda: gemm_case("TC_2SM", 200, 304, 1008),
    "tc2_mn_probe": lambda: probe_mn("TC_2SM"),
    "full_simt": lambda: full_case("SIMT", 200, 304, 520),
    "full_tc1_f32": lambda: full_case("TC_1SM", 512, 768, 1024),
    "full_tc1_f16": lambda: full_case("TC_1SM", 512, 768, 1024, "f16"),
    "full_tc1_bf16": lambda: full_case("TC_1SM", 200, 304, 520, "bf16"),
    "full_tc2_f32": lambda: full_case("TC_2SM", 512, 768, 1024),
    "full_tc2_f16": lambda: full_case("TC_2SM", 512, 768, 1024, "f16"),
    "time_lib_4096": lambda: timing("LIB", 4096, 4096, 4096),
    "time_tc1_4096": lambda: timing("TC_1SM", 4096, 4096, 4096),
    "time_tc2_4096": lambda: timing("TC_2SM", 4096, 4096, 4096),
    "time_lib_8192": lambda: timing("LIB", 8192, 8192, 8192),
    "time_tc1_8192": lambda: timing("TC_1SM", 8192, 8192, 8192),
    "time_tc2_8192": lambda: timing("TC_2SM", 8192, 8192, 8192),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one")
    ap.add_argument("--only", default="")
    ap.add_argument("--timeout", type=int, default=150)
    ap.add_argument("--out", default=os.path.join(OUT, "diag.json"))
    args = ap.parse_args()
    if args.one:
        print("DIAG_RESULT " + json.dumps(CHECKS[args.one]()))
        return
    os.makedirs(OUT, exist_ok=True)
    names = [n for n in CHECKS if not args.only or any(n.startswith(p) for p in args.only.split(","))]
    results = {}
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], capture_output=True,
                               text=True, timeout=args.timeout, env=dict(os.environ))
            line = [l for l in r.stdout.splitlines() if l.startswith("DIAG_RESULT ")]
            if line:
                results[name] = json.loads(line[-1][len("DIAG_RESULT "):])
            else:
                results[name] = {"ok": False, "rc": r.returncode, "stdout": r.stdout[-1500:], "stderr": r.stderr[-2500:]}
        except subprocess.TimeoutExpired as e:
            results[name] = {"ok": False, "timeout": True, "stdout": (e.stdout or b"")[-1500:].decode(errors="replace")
                             if isinstance(e.stdout, bytes) else str(e.stdout)[-1500:]}
        results[name]["secs"] = round(time.time() - t0, 1)
        print(name, json.dumps(results[name])[:600], flush=True)
        with open(args.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
