#!/bin/bash
for lib in libqgemm.so libqgemm_p2.so libqgemm_p12.so libqgemm.so libqgemm_p2.so libqgemm_p12.so; do echo "== $lib"; QG_LIB=$lib timeout 200 python tools/gpu_perf.py --only full_4096_pdl,full_8192,full_2048_pdl --out gpurun_out/perf_$lib.json 2>&1 | cut -c1-200; done
