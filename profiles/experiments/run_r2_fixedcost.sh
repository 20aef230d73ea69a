#!/bin/bash
# Fixed-cost profile at the 8-GPU shard shape (N = 1.25M rows per rank, d=1024, Q=64, k=10): launch list + full captures of the
# prep and finish kernels with source correlation (profiles/ncu_source_hot.py reads them).  Run through gpurun.
# The asynchronous API launches an early-exit finish after every real one: --launch-skip 6 lands on a real one (4th search).
mkdir -p gpurun_out
B="python bench.py --no-also --no-cpu-baseline --no-parity --rows 1250000"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r2_shard.csv $B --steps 5 --warmup 3 > gpurun_out/fx_ncu1.log 2>&1; echo ncu1 rc=$?
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:filter_finish --launch-skip 6 -c 1 -o gpurun_out/prof_r2_finish_shard $B --steps 2 --warmup 3 > gpurun_out/fx_ncu2.log 2>&1; echo ncu2 rc=$?
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:knn_scan_shadow --launch-skip 3 -c 1 -o gpurun_out/prof_r2_shadow_shard $B --steps 2 --warmup 3 > gpurun_out/fx_ncu3.log 2>&1; echo ncu3 rc=$?
NK_TC_DEBUG=64 $B --steps 2 --warmup 2 2>&1 >/dev/null | grep "shadow prof" | tail -3
$B --steps 50 --warmup 5 > gpurun_out/fx_bench_shard.log 2>&1; tail -1 gpurun_out/fx_bench_shard.log | cut -c1-300
