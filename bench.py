#!/usr/bin/env python
"""bench.py -- headline measurement of the quantized-linear hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|reference-gpu]
                    [-m M] [-n N] [-k K] [-s SEED]      (the flags of src/timing_quantize.cu:83-101)

A step is one pass of the hot path over one batch of synthetic input: the reference's whole
`op_quantized_mm` (row absmax-quantize X, column absmax-quantize W, int8 GEMM, dequantize) at
M = N = K = 4096 with fp32 in / fp32 out, i.e. exactly what src/timing_quantize.cu:38-58 times.
Nothing is cached between steps (W is re-quantized every step, as the reference does).

  value     whole-job TOPS (2*M*N*K per GPU-step / device time), inputs resident in HBM
  e2e       same op through the host-buffer C-ABI call (H2D of X and W, D2H of O inside the timing)
  roofline  the dominant kernel (tcgen05 int8 GEMM + fused dequantize), CUDA-event timed per step
  cpu_baseline  the CPU port of the reference math (oracle/qfast.c: OpenMP + AVX-512 VNNI) on this box's
                host cores: WHOLE steps on the same inputs, whose output also checks the GPU result bit for bit
  sustained     the same step back to back for >= 2 s (clocks recorded), against the sustained tensor peak

N > 1: column-parallel linear -- rank p owns W[:, p*N:(p+1)*N] (N = 4096 columns per GPU, weak
scaling), quantizes the replicated X locally, and the fp32 outputs are gathered by the GEMM epilogue's
peer stores (NCCL all-gather when peer memory is unavailable).  After the timed loop every rank
recomputes the full [M, N*P] product on its own GPU and compares its gathered copy bit for bit
("parity_checked"); a mismatch fails the run.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "quantized-gemm-for-transformer-inference_b200"
METRIC = "quantized linear TOPS at MxNxK (absmax quantize -> int8 GEMM -> dequantize, op_quantized_mm)"
INT8_SPEC_TOPS = 4500.0


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """SM clock / power / throttle reasons DURING the timed region.  The timed region of this
    bench lasts milliseconds, so NVML is polled from a thread (about 1 kHz); `nvidia-smi -lms`
    (the B200_PROFILING.md recipe) is the fallback when pynvml is unavailable."""

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self.stop_flag = False
        self.thread = None
        self.smi = None

    def _poll(self):
        import pynvml as nv

        h = self.h
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((sm, pw, rs))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                     "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                     "clocks_event_reasons.sw_power_cap")
                fd, self.path = tempfile.mkstemp(suffix=".csv")
                os.close(fd)
                self.smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms",
                                             "50", "-i", str(self.idx)], stdout=open(self.path, "w"),
                                            stderr=subprocess.DEVNULL)
            except Exception:
                self.smi = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            import pynvml as nv

            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.samples:
                pmax = max(p for _, p, _ in self.samples)
                busy = [s for s, p, _ in self.samples if p >= 0.6 * pmax] or [s for s, _, _ in self.samples]
                bits = 0
                for _, _, r in self.samples:
                    bits |= r
                names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                         "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                         "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                         "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
                out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(self.max),
                       "reasons": sorted(k for k, v in names.items() if bits & v), "samples": len(self.samples),
                       "power_w_max": pmax, "source": "nvml polled in-process during the timed region"}
            return out
        if self.smi is not None:
            self.smi.terminate()
            try:
                self.smi.wait(timeout=5)
            except Exception:
                self.smi.kill()
            sm, mx, reasons = [], [], set()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except (ValueError, IndexError):
                    continue
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            os.unlink(self.path)
            if sm:
                out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                       "samples": len(sm), "source": "nvidia-smi -lms 50"}
        return out


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def bench_config(M, N, K, world):
    """The `config` object both arms print (identical for the same command line)."""
    return {
        "workload": f"op_quantized_mm {M}x{N}x{K} fp32 in / fp32 out per GPU, W re-quantized every step",
        "M": M, "N": N * world, "K": K, "mode": "REF_EXACT",
        "l2": "2 rotating buffer sets, 384 MiB touched per 2 steps at 4096^3 (> 126 MB L2); no explicit flush",
        "parallelism": f"column-parallel x{world}" if world > 1 else "single GPU",
        "exchange": ("output blocks gathered to every GPU (GEMM-epilogue peer stores over NVLink; NCCL all-gather "
                     "when peer memory is unavailable)") if world > 1 else "none",
    }


def synth_inputs(M, N, K, seed, block):
    """Deterministic host inputs: X is shared by every column block (it is replicated across GPUs),
    W depends on the block (= rank).  U(-1, 1), the distribution of src/timing_quantize.cu:17-20."""
    import numpy as np

    X = np.random.default_rng([seed, 0]).random((M, K), dtype=np.float32) * 2 - 1
    W = np.random.default_rng([seed, 1 + block]).random((K, N), dtype=np.float32) * 2 - 1
    return X, W


def cpu_step_seconds(X, Ws, O, repeats=1):
    """One whole step of the CPU port per column block: quantize X rows, quantize W columns, int8 GEMM,
    dequantize (oracle/qfast.c: qf_quantized_mm_f32, scratch allocated per call like the reference)."""
    import oracle

    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        for W in Ws:
            oracle.fast_quantized_mm(X, W, out=O)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_reference_cpu(args):
    """--impl reference: the reference has no CPU implementation of its own (every op asserts
    off-device, src/ops/op_elemwise.cuh:459-463), so the CPU arm is the port under oracle/ (qfast.c),
    run on every host core this process may use.  Each step is one WHOLE op_quantized_mm per GPU of the
    repo arm's workload (world column blocks), nothing extrapolated."""
    import numpy as np

    import oracle

    M, N, K = args.m, args.n, args.k
    world = max(1, args.gpus)
    cores = oracle.fast_set_threads()  # torchrun pins OMP_NUM_THREADS=1; ask for the cores we may use
    X, _ = synth_inputs(M, N, K, args.seed, 0)
    Ws = [synth_inputs(M, N, K, args.seed, b)[1] for b in range(world)]
    O = np.empty((M, N), np.float32)
    times = []
    t_wall0 = time.perf_counter()
    for i in range(args.warmup + args.steps):
        dt = cpu_step_seconds(X, Ws, O)
        if i >= args.warmup:
            times.append(dt)
    wall = time.perf_counter() - t_wall0
    sec = statistics.mean(times)
    tops = world * 2.0 * M * N * K / sec / 1e12
    sample = (f"every step: the whole op on the full {M}x{K} X and {world} x {K}x{N} W (row + column absmax quantize, "
              f"int8 GEMM [{oracle.fast_kernel_name()}], dequantize); {args.warmup}+{args.steps} steps ran in {wall:.1f} s")
    line = {
        "impl": "reference", "metric": METRIC, "value": tops, "unit": "TOPS", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8 x int8 -> int32, fp32 scales and output (CPU: vpdpbusd)", "data": f"synthetic U(-1,1), seed {args.seed}",
        "config": bench_config(M, N, K, world), "extrapolated": False,
        "cpu_baseline": {"value": tops, "unit": "TOPS", "cores": cores, "kind": "port", "sample": sample,
                         "cpu": cpu_model(), "nproc": os.cpu_count()},
        "e2e": {"value": tops, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu(args):
    """--impl reference-gpu: the reference's OWN CUDA kernels (oracle/_ref/libref_qmm.so, compiled
    from /root/reference/src for sm_100a) on the same device-resident inputs."""
    import ctypes as C

    import torch

    so = os.path.join(ROOT, "oracle", "_ref", "libref_qmm.so")
    if not os.path.exists(so):
        print(json.dumps({"impl": "reference-gpu", "unavailable": "oracle/_ref/libref_qmm.so not built"}))
        return
    ref = C.CDLL(so)
    M, N, K = args.m, args.n, args.k
    X = torch.rand((M, K), device="cuda") * 2 - 1
    W = torch.rand((K, N), device="cuda") * 2 - 1
    O = torch.empty((M, N), device="cuda")
    ms_ev, ms_wall, ms_fp32 = C.c_double(), C.c_double(), C.c_double()
    rc = ref.ref_time_quantized_mm_dev(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(O.data_ptr()),
                                       M, N, K, C.c_float(127.0), max(1, args.warmup), max(1, args.steps),
                                       C.byref(ms_ev), C.byref(ms_wall))
    assert rc == 0, rc
    rc = ref.ref_time_mm_f32_dev(C.c_void_p(X.data_ptr()), C.c_void_p(W.data_ptr()), C.c_void_p(O.data_ptr()), M, N, K,
                                 1, max(1, args.steps // 2), C.byref(ms_fp32))
    ops = 2.0 * M * N * K
    print(json.dumps({
        "impl": "reference-gpu", "metric": METRIC, "value": ops / ms_ev.value / 1e9, "unit": "TOPS", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_ev.value, "higher_is_better": True,
        "ms_per_call_reference_style": ms_wall.value, "fp32_op_mm_ms": ms_fp32.value,
        "config": {"workload": f"reference op_quantized_mm kernels recompiled for sm_100a, {M}x{N}x{K}"},
    }))


def pcie_rates(dev, nbytes=64 << 20, reps=4):
    """Pinned-host copy rates of this box, both directions at once (the e2e pipeline runs them concurrently):
    GB/s host->device and device->host.  CUDA events on the two copy streams."""
    import torch

    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = [0.0, 0.0]
    for _ in range(reps + 1):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            e[0].record()
            d_in.copy_(h_in, non_blocking=True)
            e[1].record()
        with torch.cuda.stream(s2):
            e[2].record()
            h_out.copy_(d_out, non_blocking=True)
            e[3].record()
        torch.cuda.synchronize()
        best[0] = max(best[0], nbytes / e[0].elapsed_time(e[1]) / 1e6)
        best[1] = max(best[1], nbytes / e[2].elapsed_time(e[3]) / 1e6)
    return best


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    # NCCL's INFO log (the driver counts ranks from it) goes to stderr so that stdout carries only the JSON line
    if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    qg = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    M, N, K = args.m, args.n, args.k
    peaks = measured_peaks()

    # two rotating buffer sets: 2 x (X 64 MiB + W 64 MiB + O 64 MiB) = 384 MiB >> 126 MB L2.
    # Set 0 is the seeded host data both arms use (X shared by all ranks, W per rank); set 1 is drawn on the device.
    nset = 2
    Xh0, Wh0 = synth_inputs(M, N, K, args.seed, rank)
    g = torch.Generator(device=dev).manual_seed(1234 + args.seed)
    Xs = [torch.from_numpy(Xh0).to(dev), torch.rand((M, K), device=dev, generator=g) * 2 - 1]
    g.manual_seed(4321 + args.seed + rank)
    Ws = [torch.from_numpy(Wh0).to(dev), torch.rand((K, N), device=dev, generator=g) * 2 - 1]
    Os = [torch.empty((M, N), device=dev) for _ in range(nset)]
    gathered = torch.empty((world, M, N), device=dev) if world > 1 else None
    # N > 1: the output gather is fused into the GEMM epilogue when symmetric (peer-mapped) memory is
    # available -- every rank's [M, world*N] result buffer is written directly by all ranks' epilogues.
    # Two result buffers alternate, so ONE cross-GPU barrier per step (after the product: "every block has
    # landed everywhere") also orders the reuse of a buffer two steps later.
    exchange, symm, hdls, peer_ptrs, mc_ptrs = "none", [], [], [], []
    if world > 1:
        exchange = "nccl all_gather_into_tensor"
        if os.environ.get("QG_BENCH_EXCHANGE", "fused") == "fused":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                for _ in range(nset):
                    buf = symm_mem.empty((M, N * world), dtype=torch.float32, device=dev)
                    h = symm_mem.rendezvous(buf, dist.group.WORLD)
                    symm.append(buf)
                    hdls.append(h)
                    peer_ptrs.append([int(h.buffer_ptrs[r]) + rank * N * 4 for r in range(world) if r != rank])
                    mc = int(getattr(h, "multicast_ptr", 0) or 0)
                    # opt-in (QG_MULTICAST=1): measured SLOWER than the unicast TMA stores for this gather -- every GPU must still
                    # receive (P-1) blocks, and ingress is what bounds it (N=8: 814.6 vs 775.5 us per step, profiles/r2_scaling_run19.json)
                    mc_ptrs.append(mc + rank * N * 4 if mc and os.environ.get("QG_MULTICAST") == "1" and N % 4 == 0 else 0)
                exchange = "fused: GEMM epilogue TMA-stores into every peer's result over NVLink (symmetric memory)"
                if all(mc_ptrs):  # NVSwitch multicast mapping: one multimem.st per 16 bytes, replicated by the switch
                    exchange = ("fused: GEMM epilogue stores every tile once to the symmetric buffer's NVSwitch multicast "
                                "address (multimem.st; egress 1x the block)")
                else:
                    mc_ptrs = []
            except Exception as ex:  # no peer access on this box: fall back to the NCCL collective
                exchange = f"nccl all_gather_into_tensor (symmetric memory unavailable: {str(ex)[:120]})"
                symm, hdls, peer_ptrs, mc_ptrs = [], [], [], []
    fused = bool(symm)
    Xq = torch.empty((M, K), dtype=torch.int8, device=dev)
    Wq = torch.empty((K, N), dtype=torch.int8, device=dev)
    Wt = torch.empty((N, K), dtype=torch.int8, device=dev)
    kmajor = os.environ.get("QG_PERCALL_KMAJOR", "0") != "0"  # same switch the library reads
    Cx = torch.empty(M, device=dev)
    Cw = torch.empty(N, device=dev)

    ws = torch.empty(qg.workspace_bytes(M, N, K), dtype=torch.uint8, device=dev)

    def step(i, ev=None):
        s = i % nset
        if ev is None and world == 1:
            # the public call: one C-ABI entry, PDL-chained launches
            qg.op_quantized_mm(Xs[s], Ws[s], Os[s], 127.0, workspace=ws)
            return
        # the same launches issued one by one, so that the dominant kernel can be bracketed by CUDA events
        # (instrumented pass) or given its peer destinations (N > 1)
        if kmajor:  # weight codes transposed on the fly (QG_PERCALL_KMAJOR=1)
            qg.absmax_quant_rows(Xs[s], 127.0, qg.MODE_REF_EXACT, Xq, Cx)
            qg.prepare_weights(Ws[s], 127.0, qg.MODE_REF_EXACT, Wt, Cw)
        elif ev is None:  # both quantizers as the op runs them (column pass 2 side by side with the row quantizer)
            qg.absmax_quant_rows_cols(Xs[s], Ws[s], Xq, Cx, Wq, Cw, 127.0, qg.MODE_REF_EXACT)
        else:
            qg.absmax_quant_rows(Xs[s], 127.0, qg.MODE_REF_EXACT, Xq, Cx)
            qg.absmax_quant_cols(Ws[s], 127.0, qg.MODE_REF_EXACT, Wq, Cw)
        if ev is not None:
            ev[0].record()
        if fused and mc_ptrs:
            qg.gemm_s8_dequant_mc(Xq, Wt if kmajor else Wq, kmajor, Cx, Cw, symm[s][:, rank * N:(rank + 1) * N],
                                  mc_ptrs[s], 127.0)
            hdls[s].barrier(channel=0)  # every rank's blocks of this step have landed everywhere
        elif fused:
            qg.gemm_s8_dequant_ex(Xq, Wt if kmajor else Wq, kmajor, Cx, Cw, symm[s][:, rank * N:(rank + 1) * N],
                                  peer_ptrs[s], 127.0)
            hdls[s].barrier(channel=0)  # every rank's blocks of this step have landed everywhere
        elif kmajor:
            qg.gemm_s8t_dequant(Xq, Wt, Cx, Cw, Os[s], 127.0)
        else:
            qg.gemm_s8_dequant(Xq, Wq, Cx, Cw, Os[s], 127.0)
        if ev is not None:
            ev[1].record()
        if world > 1 and not fused:
            dist.all_gather_into_tensor(gathered, Os[s])

    # the clock sampler starts BEFORE the warm-up steps: its start-up (a 250 ms pause) would otherwise sit between the
    # warm-up and the timed region and let the GPU fall idle right before the first timed step (one box measured 94.5 us per
    # step that way against 92.9 us sustained); samples are filtered by power, so the idle ones do not count
    sampler = ClockSampler(local_rank)
    sample_clocks = rank == 0 and os.environ.get("QG_BENCH_NO_CLOCKS") is None
    if sample_clocks:
        sampler.start()
        time.sleep(0.25)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    # every rank enters the timed region together: rank 0's sampler start-up above must not show up
    # as 250 ms of waiting inside the other ranks' first exchange
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    qg.launch_count(reset=True)
    t_host0 = time.perf_counter()
    t_start.record()
    for i in range(args.steps):
        step(i)
    t_end.record()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    launches = qg.launch_count()
    clocks = sampler.stop() if sample_clocks else None
    # instrumented pass, immediately after and on the same buffers: the same K steps with a CUDA event
    # pair around the dominant kernel (events between the launches would otherwise break the
    # programmatic-dependent-launch overlap of the timed pass above)
    for i in range(args.steps):
        step(i, evs[i])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    total_ms = t_start.elapsed_time(t_end)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    ops = 2.0 * M * N * K
    value = world * ops / ms_per_step / 1e9  # TOPS, whole job

    gemm_ms = statistics.median(e[0].elapsed_time(e[1]) for e in evs)

    # ---- N > 1 parity: this rank's gathered [M, N*P] result of set 0 against the same product computed here,
    # on one GPU, from the gathered weights (column blocks are independent, so the two must agree bit for bit)
    parity = None
    if world > 1:
        step(0)  # set 0 once more, so the buffers hold a known step
        torch.cuda.synchronize()
        dist.barrier()
        Wall = torch.empty((world, K, N), device=dev)
        dist.all_gather_into_tensor(Wall, Ws[0])
        Wfull = Wall.permute(1, 0, 2).reshape(K, world * N).contiguous()
        Ofull = torch.empty((M, world * N), device=dev)
        qg.op_quantized_mm(Xs[0], Wfull, Ofull, 127.0)
        got = symm[0] if fused else gathered.permute(1, 0, 2).reshape(M, world * N)
        ok = bool(torch.equal(got.view(torch.int32), Ofull.view(torch.int32)))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        parity = {"checked": True, "ok": bool(flag.item() == 1),
                  "how": "every rank: gathered [M, N*P] output == op_quantized_mm(X, all-gathered W) on one GPU, all bits"}
        del Wall, Wfull, Ofull
        torch.cuda.synchronize()

    # per-stage timings of the two HBM-bound quantizers, outside the timed region (same buffers)
    def stage_ms(fn, n=20):
        for _ in range(3):
            fn(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(n):
            fn(j)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    rows_ms = stage_ms(lambda j: qg.absmax_quant_rows(Xs[j % nset], 127.0, qg.MODE_REF_EXACT, Xq, Cx))
    cols_ms = stage_ms(lambda j: qg.prepare_weights(Ws[j % nset], 127.0, qg.MODE_REF_EXACT, Wt, Cw) if kmajor
                       else qg.absmax_quant_cols(Ws[j % nset], 127.0, qg.MODE_REF_EXACT, Wq, Cw))

    # both quantizers the way the op runs them: column pass 1, then column pass 2 side by side with the row quantizer
    both_ms = None if kmajor else stage_ms(lambda j: qg.absmax_quant_rows_cols(Xs[j % nset], Ws[j % nset], Xq, Cx, Wq, Cw, 127.0,
                                                                              qg.MODE_REF_EXACT))

    # ---- e2e at N > 1: every rank runs the host-buffer call on its own column block at the same time (pinned host X and this
    # rank's W in, this rank's O block out; no gather -- the caller's buffers are host memory); max over ranks ----
    e2e_multi = None
    if world > 1 and os.environ.get("QG_BENCH_NO_E2E") is None:
        e2e_ms, same, failed = 0.0, True, 0.0
        Xh = Wh = Oh = None
        try:
            Xh = torch.empty((M, K), dtype=torch.float32).pin_memory()
            Wh = torch.empty((K, N), dtype=torch.float32).pin_memory()
            Oh = torch.empty((M, N), dtype=torch.float32).pin_memory()
            Xh.copy_(Xs[0]); Wh.copy_(Ws[0])
            for _ in range(2):
                qg.quantized_mm_host(Xh, Wh, out=Oh)
            torch.cuda.synchronize()
        except Exception as ex:  # this leg must not take the device-resident measurement above down with it
            failed = 1.0
            print(f"[bench] rank {rank}: e2e leg unavailable: {str(ex)[:200]}", file=sys.stderr)
        dist.barrier()  # every rank reaches the collectives below, whatever happened locally
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 10))
        if not failed:
            try:
                t0 = time.perf_counter()
                for _ in range(n_e2e):
                    qg.quantized_mm_host(Xh, Wh, out=Oh)  # blocks until Oh is written
                e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
                same = torch.equal(Oh.to(dev).view(torch.int32), _device_result(qg, Xs[0], Ws[0], ws).view(torch.int32))
            except Exception as ex:
                failed = 1.0
                print(f"[bench] rank {rank}: e2e leg failed: {str(ex)[:200]}", file=sys.stderr)
        te = torch.tensor([e2e_ms, 0.0 if same else 1.0, failed], device=dev, dtype=torch.float64)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        if te[2].item() == 0.0:
            e2e_ms = float(te[0].item())
            e2e_multi = {"value": world * ops / e2e_ms / 1e9, "unit": "TOPS", "ms_per_step": e2e_ms,
                         "h2d_bytes_per_step": world * (M * K + K * N) * 4, "d2h_bytes_per_step": world * M * N * 4,
                         "api": "qg_quantized_mm_host on every rank at once (pinned host X and this rank's W -> this rank's host O "
                                "block); bytes are whole-job totals, time is the max over ranks",
                         "parity_vs_device_path": bool(te[1].item() == 0.0)}
        del Xh, Wh, Oh

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        if parity is not None and not parity["ok"]:
            sys.exit(3)
        return

    # ---- sustained leg (N = 1): the same step back to back for >= `--sustained-seconds`, clocks sampled ----
    sustained = None
    if world == 1 and args.sustained_seconds > 0:
        n_sus = max(args.steps, int(args.sustained_seconds / (ms_per_step * 1e-3)))
        s2 = ClockSampler(local_rank)
        s2.start()
        time.sleep(0.05)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_sus):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        c2 = s2.stop()
        sus_total_ms = e0.elapsed_time(e1)
        sus_ms = sus_total_ms / n_sus
        # the GEMM alone, back to back, for ~1/3 of that time: its fraction of the SUSTAINED tensor peak
        n_g = max(10, int(args.sustained_seconds / 3 / (gemm_ms * 1e-3)))
        s3 = ClockSampler(local_rank)
        s3.start()
        e0.record()
        for i in range(n_g):
            qg.gemm_s8_dequant(Xq, Wq, Cx, Cw, Os[i % nset], 127.0)
        e1.record()
        torch.cuda.synchronize()
        c3 = s3.stop()
        g_ms = e0.elapsed_time(e1) / n_g
        sustained = {"steps": n_sus, "seconds": sus_total_ms / 1e3, "ms_per_step": sus_ms,
                     "value": ops / sus_ms / 1e9, "unit": "TOPS", "clocks": c2,
                     "gemm": {"launches": n_g, "ms": g_ms, "tops": ops / g_ms / 1e9,
                              "peak": 2.0 * peaks["bf16_tflops_sustained"],
                              "frac": ops / g_ms / 1e9 / (2.0 * peaks["bf16_tflops_sustained"]),
                              "peak_note": "2 x bf16_tflops_sustained (" + peaks["source"] + ")", "clocks": c3}}

    # ---- e2e: host buffers through the C-ABI host call (H2D X, W; compute; D2H O) ----
    e2e = None
    if world == 1:
        Xh = torch.empty((M, K), dtype=torch.float32).pin_memory()
        Wh = torch.empty((K, N), dtype=torch.float32).pin_memory()
        Oh = torch.empty((M, N), dtype=torch.float32).pin_memory()
        Xh.copy_(Xs[0]); Wh.copy_(Ws[0])
        for _ in range(2):
            qg.quantized_mm_host(Xh, Wh, out=Oh)
        torch.cuda.synchronize()
        n_e2e = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            qg.quantized_mm_host(Xh, Wh, out=Oh)  # blocks until Oh is written
        e2e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        h2d_b, d2h_b = (M * K + K * N) * 4, M * N * 4
        h2d_gbs, d2h_gbs = pcie_rates(dev)
        floor_ms = max(h2d_b / h2d_gbs, d2h_b / d2h_gbs) / 1e6
        e2e = {"value": ops / e2e_ms / 1e9, "unit": "TOPS", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
               "api": "qg_quantized_mm_host (pinned host X, W -> host O)",
               "parity_vs_device_path": bool(torch.equal(Oh.to(dev).view(torch.int32), _device_result(qg, Xs[0], Ws[0], ws).view(torch.int32))),
               "roofline": {"bound": "pcie", "achieved": h2d_b / e2e_ms / 1e6, "peak": h2d_gbs, "unit": "GB/s",
                            "frac": floor_ms / e2e_ms,
                            "note": f"pinned copies measured on this box, both directions concurrently: H2D {h2d_gbs:.1f} GB/s, "
                                    f"D2H {d2h_gbs:.1f} GB/s; floor = max(h2d_bytes/H2D, d2h_bytes/D2H) = {floor_ms:.3f} ms; "
                                    "frac = floor / measured"}}
    else:
        e2e = e2e_multi or {"value": None, "unit": "TOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                            "note": "host-buffer call switched off (QG_BENCH_NO_E2E) or unavailable on a rank (see stderr)"}

    # ---- context: LinearLayer::forward with the weights prepared once (K-major int8), fp32 and fp16 I/O ----
    cached = {}
    if world == 1:
        for name, tdt in (("f32", torch.float32), ("f16", torch.float16)):
            lin = qg.LinearLayer(K, N, device=dev, dtype=tdt)
            lin.init_uniform()
            lin.quantize_weights()
            xin = [x.to(tdt) for x in Xs]
            yout = torch.empty((M, N), dtype=tdt, device=dev)
            ms = stage_ms(lambda j: lin.forward(xin[j % nset], yout))
            cached[name] = {"ms": ms, "tops": ops / ms / 1e9}
            del lin, xin, yout

    # ---- library context: cuBLASLt int8 and fp16 GEMMs of the same shape (not on our path) ----
    lib = {}
    try:
        a8 = torch.randint(-127, 128, (M, K), dtype=torch.int8, device=dev)
        b8 = torch.randint(-127, 128, (K, N), dtype=torch.int8, device=dev)
        a16 = torch.randn((M, K), dtype=torch.float16, device=dev)
        b16 = torch.randn((K, N), dtype=torch.float16, device=dev)
        for name, fn in (("cublaslt_int8_tops", lambda: torch._int_mm(a8, b8)), ("cublas_fp16_tflops", lambda: a16 @ b16)):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            lib[name] = ops / (e0.elapsed_time(e1) / 10) / 1e9
            lib[name.rsplit("_", 1)[0] + "_ms"] = e0.elapsed_time(e1) / 10
    except Exception as ex:  # context only
        lib["error"] = str(ex)[:200]

    # ---- CPU baseline: the CPU port on the host cores, whole steps on the inputs of buffer set 0; its output
    # doubles as the checker of the GPU result (N = 1 parity: every bit of the 4096 x 4096 output) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle

        cores = oracle.fast_set_threads()
        Ocpu = np.empty((M, N), np.float32)
        t_w0 = time.perf_counter()
        sec = cpu_step_seconds(Xh0, [Wh0], Ocpu, repeats=3 if M * N * K <= 4096 ** 3 else 1)
        cpu = {"value": ops / sec / 1e12, "unit": "TOPS", "cores": cores, "kind": "port", "cpu": cpu_model(), "nproc": os.cpu_count(),
               "sample": f"whole {M}x{N}x{K} steps (best of 3) on the inputs of buffer set 0: row + column absmax quantize, int8 GEMM "
                         f"[{oracle.fast_kernel_name()}], dequantize; {time.perf_counter() - t_w0:.1f} s of CPU work",
               "ms_per_step": sec * 1e3}
        got = _device_result(qg, Xs[0], Ws[0], ws).cpu().numpy()
        parity = {"checked": True, "ok": bool(np.array_equal(got.view(np.int32), Ocpu.view(np.int32))),
                  "how": "op_quantized_mm output of buffer set 0 == the CPU port's output on the same inputs, all bits"}

    # burst peak when the SM clock stayed at its maximum during the (short) timed region, else sustained
    at_max = bool(clocks and clocks.get("sm_mhz") and clocks["sm_mhz"] >= 0.95 * clocks["sm_max_mhz"])
    peak_key = "bf16_tflops" if at_max or not (clocks and clocks.get("sm_mhz")) else "bf16_tflops_sustained"
    int8_peak = 2.0 * peaks[peak_key]
    # dram__bytes_read.sum + dram__bytes_write.sum of the GEMM launch, from the committed ncu --set full capture
    # of this same command (never measured in this run: nothing here runs under a profiler)
    traffic, traffic_src = None, None
    for tname in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if world == 1 and (M, N, K) == (4096, 4096, 4096) and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
            break
    gemm_tops = ops / gemm_ms / 1e9
    cfg = bench_config(M, N, K, world)
    line = {
        "metric": METRIC, "value": value, "unit": "TOPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8 x int8 -> int32 (tcgen05 kind::i8), fp32 scales and output", "data": f"synthetic U(-1,1), seed {args.seed}",
        "config": cfg, "exchange_used": exchange,
        "parity_checked": bool(parity and parity["ok"]), "parity": parity,
        "roofline": {"bound": "tensor", "kernel": "gemm_i8_tc_kernel (tcgen05 kind::i8 + fused dequantize epilogue)",
                     "achieved": gemm_tops, "peak": int8_peak, "unit": "TOP/s", "frac": gemm_tops / int8_peak,
                     "peak_note": f"2 x {peak_key} from MEASURED_PEAKS.json ({peaks['source']}); int8 dense "
                                  f"rate is 2x bf16; spec 4500 TOP/s -> frac_spec {gemm_tops / INT8_SPEC_TOPS:.3f}",
                     "frac_spec": gemm_tops / INT8_SPEC_TOPS, "ms": gemm_ms, "traffic": traffic, "traffic_unit": "bytes per launch",
                     "traffic_source": traffic_src,
                     "timing": "CUDA events around this launch in every step of an instrumented pass of the same K steps, "
                               "run directly after the timed pass"},
        "stages": {
            "quant_rows": {"ms": rows_ms, "achieved": (M * K * 5 + 4 * M) / rows_ms / 1e6, "unit": "GB/s", "bound": "hbm",
                           "frac": (M * K * 5 + 4 * M) / rows_ms / 1e6 / peaks["hbm_gbs"]},
            "quant_cols": {"ms": cols_ms, "achieved": (K * N * 5 + 4 * N) / cols_ms / 1e6, "unit": "GB/s", "bound": "hbm",
                           "frac": (K * N * 5 + 4 * N) / cols_ms / 1e6 / peaks["hbm_gbs"],
                           "note": "two launches timed apart from the row quantizer: pass 1 (column maxima) streams W from HBM, pass 2 re-reads "
                                   "it from L2 (W <= 80 MiB) and is bound by L2 throughput, not HBM (DESIGN 4.2: hot 10.2 us vs 14.3 us cold at "
                                   "4096^2); scored against ONE read of W. In the timed step pass 2 shares a launch with the row quantizer "
                                   "(next entry)"},
            "quant_rows_and_cols": None if both_ms is None else {
                "ms": both_ms, "achieved": ((M * K + K * N) * 5 + 4 * (M + N)) / both_ms / 1e6, "unit": "GB/s", "bound": "hbm",
                "frac": ((M * K + K * N) * 5 + 4 * (M + N)) / both_ms / 1e6 / peaks["hbm_gbs"],
                "note": "what the timed step runs when W fits in L2: the two stages above are the same kernels launched apart"},
            "gemm_dequant": {"ms": gemm_ms, "tops": gemm_tops},
        },
        "sustained": sustained,
        "linear_prepared_weights": cached,
        "library_context": lib,
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "host_enqueue_ms": host_enqueue_ms,
        "clocks": clocks,
    }
    if world > 1:
        # exchange roofline: bytes every GPU must RECEIVE per step over NVLink against the measured peer bandwidth
        rx = (world - 1) * M * N * 4
        line["exchange_roofline"] = {"bound": "nvlink ingress", "bytes_received_per_gpu": rx,
                                     "achieved": rx / ms_per_step / 1e6, "peak": 770.0, "unit": "GB/s",
                                     "frac": rx / ms_per_step / 1e6 / 770.0,
                                     "peak_note": "measured peer copy per direction per GPU (B200_PROFILING.md); nominal 900"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        sys.stderr.write("bench.py: PARITY MISMATCH\n")
        sys.exit(3)


def _device_result(qg, X, W, ws):
    import torch

    O = torch.empty((X.shape[0], W.shape[1]), device=X.device)
    qg.op_quantized_mm(X, W, O, 127.0, workspace=ws)
    torch.cuda.synchronize()
    return O


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-gpu"])
    ap.add_argument("--size", type=int, default=4096, help="M = N = K (overridden by -m / -n / -k)")
    # the flags of the reference's timing driver, src/timing_quantize.cu:83-101
    ap.add_argument("-m", type=int, default=None, help="rows of X")
    ap.add_argument("-n", type=int, default=None, help="columns of W (per GPU)")
    ap.add_argument("-k", type=int, default=None, help="inner dimension")
    ap.add_argument("-s", "--seed", type=int, default=0, help="random seed (randgen_seed)")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.m = args.m or args.size
    args.n = args.n or args.size
    args.k = args.k or args.size
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference_cpu(args)
        return
    if args.impl == "reference-gpu":
        if rank == 0:
            run_reference_gpu(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
