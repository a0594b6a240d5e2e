"""Where do the stack's small GEMMs (config 3: 4096 tokens, d_model 512, d_ff 2048) spend their 12-19 us?
Back-to-back time of the prepared-weight GEMM + dequantize (fp32 out) per shape and variant, with the in-kernel wait counters
(qg_debug_gemm_stats: producer / MMA / epilogue role-loop durations and their barrier waits, medians over CTAs, in us at the
SM clock).  Debug switches are read once per process: run once per setting (QG_DBG_ALL_HALF, QG_DBG_NOEPI, ...)."""
import ctypes as C, importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

qg = importlib.import_module("quantized-gemm-for-transformer-inference_b200")
dev = "cuda"
shapes = [(4096, 512, 512), (4096, 1536, 512), (4096, 2048, 512), (4096, 512, 2048)]
tag = {k: os.environ[k] for k in os.environ if k.startswith("QG_")}
clk_mhz = 1965.0
for variant, vname in ((qg.GEMM_TC_2SM, "2sm"), (qg.GEMM_TC_1SM, "1sm")):
    qg.set_gemm_variant(variant)
    for (M, N, K) in shapes:
        A = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        Bt = torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev)
        Cx, Cw = torch.rand(M, device=dev), torch.rand(N, device=dev)
        O = torch.empty((M, N), dtype=torch.float32, device=dev)
        for _ in range(5):
            qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        st = torch.zeros((148, 8), dtype=torch.int64, device=dev)
        qg.lib().qg_debug_gemm_stats(C.c_void_p(st.data_ptr()))
        qg.gemm_s8t_dequant(A, Bt, Cx, Cw, O)
        torch.cuda.synchronize()
        qg.lib().qg_debug_gemm_stats(C.c_void_p(0))
        s = st.cpu()
        used = s[:, 6] > 0
        med = lambda c: float(s[used, c].double().median()) / clk_mhz
        mx = lambda c: float(s[used, c].double().max()) / clk_mhz
        lead = s[:, 4] > 0
        res = {"env": tag, "variant": vname, "shape": [M, N, K], "b2b_us": round(us, 2), "ctas": int(used.sum()),
               "producer_loop_us": round(med(1), 2), "producer_wait_empty_us": round(med(0), 2),
               "mma_loop_us": round(float(s[lead, 4].double().median()) / clk_mhz, 2),
               "mma_wait_full_us": round(float(s[lead, 2].double().median()) / clk_mhz, 2),
               "mma_wait_tmem_us": round(float(s[lead, 3].double().median()) / clk_mhz, 2),
               "epi_loop_us": round(med(6), 2), "epi_loop_max_us": round(mx(6), 2), "epi_wait_tfull_us": round(med(5), 2),
               "finish_spread_us": round(float((s[used, 7].max() - s[used, 7].min())) / 1e3, 2)}
        print(json.dumps(res), flush=True)
