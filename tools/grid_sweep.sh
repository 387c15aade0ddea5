#!/bin/bash
# usage: tools/grid_sweep.sh <trajectories> <blocks-per-SM list...>
N=$1; shift
for g in "$@"; do
  MTM_FUSED_BLOCKS_PER_SM=$g python bench.py --steps 20 --warmup 3 --trajectories $N --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n', $N, 'blocks/SM', $g, 'ms', round(d['ms_per_step'],4), 'kernel', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],4))"
done
