timeout -k 5 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash profiles/experiments/run_r2_timeline.sh 2>&1 | grep "finish prof"
B="python bench.py --no-also --no-cpu-baseline --no-parity"
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('  ms/step %.4f scan %.4f e2e %.4f ms'%(d['ms_per_step'],r['avg_launch_ms'],d['e2e']['ms_per_step']))"; }
echo shard; $B --rows 1250000 --steps 100 --warmup 5 | show
echo c2; $B --workload c2 --steps 100 --warmup 5 | show
echo c1; $B --workload c1 --steps 300 --warmup 10 | show
