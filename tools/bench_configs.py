"""Measurements for the BASELINE.json configs other than the headline one (which bench.py owns):

  config 2  square sweep M=N=K in {1024, 2048, 4096, 8192}: quantized linear vs fp16 GEMM
  config 4  OPT-6.7B-shaped decoder linears, T = 8 x 2048 tokens, fp16 in/out, 6 injected outlier
            feature dims (x20), threshold 6.0: with / without the decomposition, error vs fp16 GEMM
  config 5  OPT-66B-shaped FFN shards as one GPU sees them at P = 8 (multi-GPU runs: bench.py --gpus N)

All timings: CUDA events, 3 warm-up + 10 timed calls, two rotating activation buffers.
Writes gpurun_out/configs.json.
"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
DEV = "cuda"


def timed(fn, iters=10, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def linear_case(name, M, K, N, dt=torch.float16, outliers=0, w_std=0.02, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    X = [torch.randn((M, K), device=DEV, generator=g).to(dt) for _ in range(2)]
    idx = None
    if outliers:
        cols = torch.randperm(K, device=DEV, generator=g)[:outliers].sort().values
        for x in X:
            x[:, cols] *= 20
        idx = cols.to(torch.int32)
    lin = qg.LinearLayer(K, N, device=DEV, dtype=dt)
    lin.w.normal_(0, w_std, generator=g)
    lin.b.zero_()
    lin.quantize_weights()
    y = torch.empty((M, N), dtype=dt, device=DEV)
    ops = 2.0 * M * N * K
    res = {"name": name, "M": M, "K": K, "N": N, "dtype": str(dt).split(".")[-1], "outlier_dims": outliers}
    us = timed(lambda i: lin.forward(X[i & 1], y))
    res["int8_linear_us"], res["int8_linear_tops"] = us, ops / us / 1e6
    ref = X[0].float() @ lin.w.float() if M * N <= 4096 * 16384 else None
    if ref is not None:
        lin.forward(X[0], y)
        res["int8_mean_abs_err"] = (y.float() - ref).abs().mean().item()
        res["ref_rms"] = ref.pow(2).mean().sqrt().item()
    if idx is not None:
        us = timed(lambda i: lin.forward_outlier(X[i & 1], y, idx))
        res["int8_outlier_linear_us"], res["int8_outlier_linear_tops"] = us, ops / us / 1e6
        found, n = qg.outlier_cols(X[0], 6.0)
        res["outliers_detected"] = n
        us = timed(lambda i: qg.outlier_cols(X[i & 1], 6.0))
        res["outlier_detect_us"] = us
        if ref is not None:
            lin.forward_outlier(X[0], y, idx)
            res["int8_outlier_mean_abs_err"] = (y.float() - ref).abs().mean().item()
    w16 = lin.w.to(torch.float16)
    x16 = [x.to(torch.float16) for x in X]
    us = timed(lambda i: torch.matmul(x16[i & 1], w16))
    res["fp16_gemm_us"], res["fp16_gemm_tflops"] = us, ops / us / 1e6
    if ref is not None:
        res["fp16_mean_abs_err"] = (torch.matmul(x16[0], w16).float() - ref).abs().mean().item()
    res["speedup_vs_fp16"] = res["fp16_gemm_us"] / res["int8_linear_us"]
    if not outliers and M == N == K:  # config 2: the three views SURVEY 8(d) asks for, plus the library int8 GEMM
        Wt, Cw = lin.quantize_weights()
        Xq = torch.empty((M, K), dtype=torch.int8, device=DEV)
        Cx = torch.empty(M, device=DEV)
        qg.absmax_quant_rows(X[0], 127.0, qg.MODE_REF_EXACT, Xq, Cx)
        us = timed(lambda i: qg.gemm_s8t_dequant(Xq, Wt, Cx, Cw, y))
        res["gemm_only_us"], res["gemm_only_tops"] = us, ops / us / 1e6
        us = timed(lambda i: qg.op_quantized_mm(X[i & 1], lin.w, y, 127.0))
        res["full_op_requantizing_w_us"], res["full_op_tops"] = us, ops / us / 1e6
        a8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=DEV)
        b8 = torch.randint(-127, 128, (K, N), dtype=torch.int8, device=DEV)
        us = timed(lambda i: torch._int_mm(a8, b8))
        res["cublaslt_int8_gemm_us"], res["cublaslt_int8_tops"] = us, ops / us / 1e6
        del a8, b8, Xq
    del lin, X, y, w16, x16
    torch.cuda.empty_cache()
    return res


def attention_case(batch=32, seq=128, d_model=512, heads=8):
    """config 3 building block: multi-head self-attention over `batch` sequences, all projections int8.
    fused = one qg_attention_forward call; looped = the reference's structure (transformer.cu:27-50):
    one single-head AttentionLayer call per (sequence, head)."""
    g = torch.Generator(device=DEV).manual_seed(3)
    X = [torch.randn((batch * seq, d_model), device=DEV, generator=g) for _ in range(2)]
    mha = qg.MultiHeadAttention(d_model, heads, device=DEV)
    mha.init_uniform(g)
    out = torch.empty((batch * seq, d_model), device=DEV)
    res = {"name": f"mha_b{batch}_s{seq}_d{d_model}_h{heads}", "tokens": batch * seq}
    us = timed(lambda i: mha.forward(X[i & 1], X[i & 1], out, batch=batch))
    res["fused_us"], res["fused_tok_per_s"] = us, batch * seq / us * 1e6
    d = d_model // heads
    layers = []
    for h in range(heads):
        att = qg.AttentionLayer(d_model, d, d, device=DEV)
        for dst, src in zip((att.W_q, att.W_k, att.W_v), mha.head_weights(h)):
            dst.copy_(src)
        layers.append(att)
    o1 = torch.empty((seq, d), device=DEV)

    def looped(i):
        x = X[i & 1]
        for b in range(batch):
            xb = x[b * seq:(b + 1) * seq]
            for att in layers:
                att.forward(xb, o1)

    us = timed(looped, iters=3, warm=1)
    res["per_head_per_sequence_calls_us"] = us
    res["speedup_fused_vs_looped"] = us / res["fused_us"]
    # fp32 library attention for scale (not the same arithmetic: cuBLAS/SDPA, no quantization)
    w = mha.W_qkv
    def torch_ref(i):
        qkv = X[i & 1] @ w
        H, dk = heads, d
        q = qkv[:, :H * dk].view(batch, seq, H, dk).transpose(1, 2)
        k = qkv[:, H * dk:2 * H * dk].view(batch, seq, H, dk).transpose(1, 2)
        v = qkv[:, 2 * H * dk:].view(batch, seq, H, dk).transpose(1, 2)
        return torch.nn.functional.scaled_dot_product_attention(q, k, v)
    res["torch_fp32_sdpa_us"] = timed(torch_ref)
    return res


def transformer_case(batch=32, seq=128, d_model=512, heads=8, d_ff=2048, n_blocks=6):
    """BASELINE config 3: encoder + decoder stack, 6 + 6 blocks, all projections int8, 32 x 128 tokens."""
    tf = importlib.import_module(qg.__name__ + ".transformer")
    g = torch.Generator(device=DEV).manual_seed(0)
    T = batch * seq
    enc, dec = tf.Encoder(d_model, heads, n_blocks, d_ff, DEV), tf.Decoder(d_model, heads, n_blocks, d_ff, DEV)
    enc.init_uniform(g); dec.init_uniform(g)
    X = [torch.randn((T, d_model), device=DEV, generator=g) for _ in range(2)]
    Y = [torch.randn((T, d_model), device=DEV, generator=g) for _ in range(2)]
    eo, do = torch.empty((T, d_model), device=DEV), torch.empty((T, d_model), device=DEV)

    def step(i):
        enc.forward(X[i & 1], eo, batch)
        dec.forward(Y[i & 1], eo, do, batch)

    us = timed(step, iters=5, warm=2)
    res = {"name": f"transformer_enc{n_blocks}_dec{n_blocks}_b{batch}_s{seq}_d{d_model}_h{heads}_ff{d_ff}", "tokens": T,
           "stack_us": us, "tok_per_s": T / us * 1e6, "finite": bool(torch.isfinite(do).all().item())}
    res["encoder_block_us"] = timed(lambda i: enc.blocks[0].forward(X[i & 1], eo, batch))
    res["decoder_block_us"] = timed(lambda i: dec.blocks[0].forward(Y[i & 1], eo, do, batch))
    return res


def timing_quantize_case(M=2048, N=512, K=512):
    """BASELINE config 0: the shape of src/timing_quantize.cu (README: 0.31954 ms fp32 vs 1.33682 ms quantized,
    hardware not stated): our op_quantized_mm and fp32 op_mm, the reference's own kernels on this GPU, and the
    CPU oracle on the host cores."""
    import ctypes as C
    import time

    import numpy as np

    X = torch.rand((M, K), device=DEV) * 2 - 1
    W = torch.rand((K, N), device=DEV) * 2 - 1
    O = torch.empty((M, N), device=DEV)
    res = {"name": f"timing_quantize_{M}x{N}x{K}"}
    res["ours_quantized_us"] = timed(lambda i: qg.op_quantized_mm(X, W, O, 127.0))
    res["ours_fp32_op_mm_us"] = timed(lambda i: qg.op_mm(X, W, O))
    res["torch_fp32_matmul_us"] = timed(lambda i: torch.matmul(X, W, out=O))
    so = os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so")
    if os.path.exists(so):
        ref = C.CDLL(so)
        ev, wall, fp = C.c_double(), C.c_double(), C.c_double()
        ref.ref_time_quantized_mm_dev(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(O.data_ptr()), M, N, K,
                                      C.c_float(127.0), 3, 10, C.byref(ev), C.byref(wall))
        ref.ref_time_mm_f32_dev(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(O.data_ptr()), M, N, K, 3, 10,
                                C.byref(fp))
        res["reference_quantized_kernels_us"] = ev.value * 1e3
        res["reference_quantized_call_us"] = wall.value * 1e3  # with its 9 cudaMalloc/cudaFree per call
        res["reference_fp32_op_mm_us"] = fp.value * 1e3
    sys.path.insert(0, ROOT)
    oracle = importlib.import_module("oracle")
    Xh, Wh = X.cpu().numpy(), W.cpu().numpy()
    oracle.quantized_mm(Xh[:64], Wh)
    t0 = time.perf_counter()
    oracle.quantized_mm(Xh, Wh)
    res["cpu_oracle_us"] = (time.perf_counter() - t0) * 1e6
    res["cpu_threads"] = oracle.num_threads()
    # error of the quantized result against the fp32 product (README: signed mean 4.58e-5, shape not stated)
    qg.op_quantized_mm(X, W, O, 127.0)
    err = (X.double() @ W.double()) - O.double()
    res["err_signed_mean"], res["err_mean_abs"], res["err_max_abs"] = err.mean().item(), err.abs().mean().item(), err.abs().max().item()
    return res


def ffn_chain_case(name, M, d_in, d_ff, d_out, dt=torch.float16):
    """ll1 -> relu -> ll2 (src/transformer.cu:63-71): two LinearLayer::forward calls + op_relu apart, against the one-call chain
    whose first epilogue applies the ReLU and hands the row maxima to the second layer's quantizer."""
    g = torch.Generator(device=DEV).manual_seed(3)
    X = [torch.randn((M, d_in), device=DEV, generator=g).to(dt) for _ in range(2)]
    l1 = qg.LinearLayer(d_in, d_ff, device=DEV, dtype=dt)
    l2 = qg.LinearLayer(d_ff, d_out, device=DEV, dtype=dt)
    for l in (l1, l2):
        l.w.normal_(0, 0.02, generator=g)
        l.b.normal_(0, 0.1, generator=g)
    W1t, Cw1 = l1.quantize_weights()
    W2t, Cw2 = l2.quantize_weights()
    H = torch.empty((M, d_ff), dtype=dt, device=DEV)
    Y = torch.empty((M, d_out), dtype=dt, device=DEV)
    Y2 = torch.empty_like(Y)

    def apart(i):
        l1.forward(X[i & 1], H)
        H.relu_()
        l2.forward(H, Y2)

    def chain(i):
        qg.ffn_forward(X[i & 1], W1t, Cw1, l1.b, W2t, Cw2, l2.b, H, Y)

    apart(0); chain(0)
    torch.cuda.synchronize()
    res = {"name": name, "M": M, "d_in": d_in, "d_ff": d_ff, "d_out": d_out, "dtype": str(dt).split(".")[-1],
           "same_bits": bool(torch.equal(Y.view(torch.int16 if dt != torch.float32 else torch.int32),
                                         Y2.view(torch.int16 if dt != torch.float32 else torch.int32)))}
    res["apart_us"] = timed(apart)
    res["chain_us"] = timed(chain)
    ops = 2.0 * M * d_ff * (d_in + d_out)
    res["chain_tops"] = ops / res["chain_us"] / 1e6
    w1, w2 = l1.w.to(torch.float16), l2.w.to(torch.float16)
    x16 = [x.to(torch.float16) for x in X]
    res["fp16_two_gemms_relu_us"] = timed(lambda i: torch.matmul(torch.relu_(torch.matmul(x16[i & 1], w1)), w2))
    del l1, l2, X, H, Y, Y2, w1, w2, x16
    torch.cuda.empty_cache()
    return res


def outlier_sweep_case(M=16384, K=4096, N=4096, dt=torch.float16):
    """config 4's out-projection with 0 .. 64 outlier feature dims: cost of the decomposition against the plain int8 linear."""
    g = torch.Generator(device=DEV).manual_seed(5)
    res = {"name": "opt6.7b_out_outlier_sweep", "M": M, "K": K, "N": N, "dtype": str(dt).split(".")[-1], "us": {}}
    lin = qg.LinearLayer(K, N, device=DEV, dtype=dt)
    lin.w.normal_(0, 0.02, generator=g)
    lin.b.zero_()
    lin.quantize_weights()
    y = torch.empty((M, N), dtype=dt, device=DEV)
    X = torch.randn((M, K), device=DEV, generator=g).to(dt)
    res["us"]["plain"] = timed(lambda i: lin.forward(X, y))
    for n in (6, 8, 16, 24, 32, 48, 64):
        cols = torch.randperm(K, device=DEV, generator=g)[:n].sort().values.to(torch.int32)
        res["us"][str(n)] = timed(lambda i: lin.forward_outlier(X, y, cols))
    del lin, X, y
    torch.cuda.empty_cache()
    return res


def main():
    out = []
    out.append(timing_quantize_case())  # config 0: both shapes found in the reference's timing driver
    out.append(timing_quantize_case(2048, 2048, 2048))
    for n in (1024, 2048, 4096, 8192):  # config 2
        out.append(linear_case(f"square_{n}_f16", n, n, n, torch.float16, w_std=1.0 / n ** 0.5))
        out.append(linear_case(f"square_{n}_f32", n, n, n, torch.float32, w_std=1.0 / n ** 0.5))
    T = 8 * 2048  # config 4
    for name, K, N in (("opt6.7b_qkv", 4096, 12288), ("opt6.7b_out", 4096, 4096), ("opt6.7b_fc1", 4096, 16384),
                       ("opt6.7b_fc2", 16384, 4096)):
        out.append(linear_case(name, T, K, N, torch.float16, outliers=6))
    for name, K, N in (("opt66b_fc1_shard_of_8", 9216, 36864 // 8), ("opt66b_fc2_shard_of_8", 36864, 9216 // 8)):  # config 5
        out.append(linear_case(name, 4096, K, N, torch.float16))
    out.append(outlier_sweep_case())
    out.append(ffn_chain_case("ffn_config3_4096x512x2048", 4096, 512, 2048, 512, torch.float32))
    out.append(ffn_chain_case("ffn_opt6.7b_T4096", 4096, 4096, 16384, 4096, torch.float16))
    out.append(ffn_chain_case("ffn_4096_cubed", 4096, 4096, 4096, 4096, torch.float16))
    out.append(attention_case())  # config 3 building block
    out.append(transformer_case())  # config 3
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w") as f:
        json.dump(out, f, indent=1)
    for r in out:
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}))


if __name__ == "__main__":
    main()
