#!/bin/bash
# In-kernel timelines (NK_TC_DEBUG=64) of the 16-bit scan and the finish kernel at the 8-GPU shard shape and at N=10M.
B="python bench.py --no-also --no-cpu-baseline --no-parity"
NK_TC_DEBUG=64 $B --rows 1250000 --steps 2 --warmup 2 2>&1 >/dev/null | grep "prof" | tail -5
NK_TC_DEBUG=64 $B --steps 2 --warmup 2 2>&1 >/dev/null | grep "finish prof" | tail -2
