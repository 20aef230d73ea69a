#!/bin/bash
# One-GPU measurement sweep behind DESIGN.md §3 / README numbers.  Usage (on the GPU box): bash profiles/experiments/sweep.sh
cd "$(dirname "$0")/../.."
S="python profiles/experiments/bsum.py"
run() { timeout -s KILL 300 python bench.py "$@" 2>&1 | tail -1 | tee -a gpurun_out/sweep.jsonl | $S; }
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
run --steps 20 --warmup 3                                # headline, default path, with CPU baseline
run --impl reference --steps 2 --warmup 1
for p in filter tensor simt; do run --path $p --steps 5 --warmup 3 --no-cpu; done
for w in c1 c2 c3 c4 q1; do run --workload $w --steps 10 --warmup 3 --no-cpu; done
run --workload c3 --k 10 --steps 10 --warmup 3 --no-cpu   # Q=1024 k=10 (configs[4]'s per-GPU shape at N=10M)
run --workload q1 --path simt --steps 10 --warmup 3 --no-cpu
run --workload c2 --path filter --steps 10 --warmup 3 --no-cpu
