#!/bin/bash
# Programmatic dependent launch on / off, alternating twice: step time (device-resident, asynchronous API) and e2e
# (host-synchronous nk_search) at the 8-GPU shard shape and configs[1].
B="python bench.py --no-also --no-cpu-baseline --no-parity"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('  ms/step %.4f scan %.4f e2e %.4f ms'%(d['ms_per_step'],r['avg_launch_ms'],d['e2e']['ms_per_step']))"; }
for rep in 1 2; do for pdl in 1 0; do
  echo "NK_PDL=$pdl shard";  NK_PDL=$pdl $B --rows 1250000 --steps 100 --warmup 5 | show
  echo "NK_PDL=$pdl c2";     NK_PDL=$pdl $B --workload c2 --steps 100 --warmup 5 | show
done; done
