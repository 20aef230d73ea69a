B="python bench.py --no-also --no-cpu-baseline --no-parity"
NK_TC_DEBUG=64 $B --rows 1250000 --steps 2 --warmup 2 2>&1 >/dev/null | grep "shadow prof" | tail -3
for v in 0 -2; do echo "NK_PRUNE_TRIGGER=$v"; NK_PRUNE_TRIGGER=$v $B --rows 1250000 --steps 50 --warmup 5 | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('  ms/step %.4f scan %.4f'%(d['ms_per_step'],r['avg_launch_ms']))"; done
