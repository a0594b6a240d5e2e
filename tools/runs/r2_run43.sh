#!/bin/bash
# round 2, run 43 (1 GPU): the stack's small GEMMs -- role-loop durations and waits, 256- vs 128-column tiles, without the epilogue
mkdir -p gpurun_out
: > gpurun_out/r2_43_small_gemm.jsonl
timeout 120 python tools/small_gemm_diag.py >> gpurun_out/r2_43_small_gemm.jsonl 2> gpurun_out/r2_43.err; echo "rc=$?"
QG_DBG_ALL_HALF=1 timeout 120 python tools/small_gemm_diag.py >> gpurun_out/r2_43_small_gemm.jsonl 2>> gpurun_out/r2_43.err; echo "rc=$?"
QG_DBG_NOEPI=1 timeout 120 python tools/small_gemm_diag.py >> gpurun_out/r2_43_small_gemm.jsonl 2>> gpurun_out/r2_43.err; echo "rc=$?"
QG_PDL=0 timeout 120 python tools/small_gemm_diag.py >> gpurun_out/r2_43_small_gemm.jsonl 2>> gpurun_out/r2_43.err; echo "rc=$?"
cut -c1-420 gpurun_out/r2_43_small_gemm.jsonl; tail -3 gpurun_out/r2_43.err
