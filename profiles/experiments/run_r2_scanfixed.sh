#!/bin/bash
# Where do the ~50 us of fixed cost of one 16-bit scan launch go?  Timeline (NK_TC_DEBUG=64) + knob sweeps at the 8-GPU shard
# shape.  Run through gpurun; prints one line per variant: ms/step, scan ms, prep+finish+gaps.
mkdir -p gpurun_out
B="python bench.py --no-also --no-cpu-baseline --no-parity"
line() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=d['roofline']
print(f"  ms/step {d['ms_per_step']:.4f}  scan {r['avg_launch_ms']:.4f}  rest {d['ms_per_step']-r['avg_launch_ms']*r['scan_launches_per_step']:.4f}  frac {r['frac']:.3f}")
PY
}
run() { name=$1; shift; echo "== $name"; env "$@" $B --steps 50 --warmup 5 > gpurun_out/sf_$name.log 2>gpurun_out/sf_$name.err; line gpurun_out/sf_$name.log; }
echo "== timeline"; NK_TC_DEBUG=64 $B --rows 1250000 --steps 3 --warmup 2 2>&1 >/dev/null | grep "shadow prof" | tail -4

B1="$B"
B="$B1 --rows 1250000"
run base NK_X=0
if [ -z "$QUICK" ]; then
run trigger768 NK_PRUNE_TRIGGER=768
run trigger64 NK_PRUNE_TRIGGER=64
run trigger128 NK_PRUNE_TRIGGER=128
run sample512 NK_TAU_SAMPLE_S=512
run sample2048 NK_TAU_SAMPLE_S=2048
run sample512_t64 NK_TAU_SAMPLE_S=512 NK_PRUNE_TRIGGER=64
run nosample NK_TAU_SAMPLE=0
else
run trigger384 NK_PRUNE_TRIGGER=384
run trigger192 NK_PRUNE_TRIGGER=192
fi
B="$B1 --rows 625000"
run rows625k NK_X=0
B="$B1 --rows 2500000"
run rows2500k NK_X=0
