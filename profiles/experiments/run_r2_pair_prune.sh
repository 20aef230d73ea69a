#!/bin/bash
# CTA-pair kernel: prune select sized to the live key count vs always full width (NK_PRUNE_TRIGGER=-1), same box.
timeout -k 5 300 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "cta_pairs" 2>&1 | tail -2
B="python bench.py --no-also --no-cpu-baseline --no-parity"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  ms/step %.3f'%d['ms_per_step'])"; }
for v in 0 -1 0 -1; do
  echo "NK_PRUNE_TRIGGER=$v c3 (k=100)"; NK_PRUNE_TRIGGER=$v $B --workload c3 --steps 8 --warmup 3 | show
done
echo "sized c3 k=10"; $B --workload c3 --k 10 --steps 8 --warmup 3 | show
echo "full  c3 k=10"; NK_PRUNE_TRIGGER=-1 $B --workload c3 --k 10 --steps 8 --warmup 3 | show
