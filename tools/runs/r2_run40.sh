#!/bin/bash
# round 2, run 40 (2 GPUs): NVLink bytes of the fused GEMM + exchange kernel counted by ncu (single process, device 1 as the peer),
# then bench.py --gpus 2 with the N > 1 e2e leg
mkdir -p gpurun_out
timeout 300 python tools/prof_gemm_peer.py > gpurun_out/r2_40_peer.json 2> gpurun_out/r2_40_peer.err; echo "peer rc=$?"; cat gpurun_out/r2_40_peer.json; tail -2 gpurun_out/r2_40_peer.err
QG_PEER_REPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --cache-control none -k regex:gemm_i8_tc -c 6 --csv --log-file gpurun_out/r2_40_peer_ncu.csv python tools/prof_gemm_peer.py > gpurun_out/r2_40_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2_40_ncu.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_40_bench_n2.json 2> gpurun_out/r2_40_bench_n2.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_40_bench_n2.json; tail -3 gpurun_out/r2_40_bench_n2.err | cut -c1-300
