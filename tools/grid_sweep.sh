#!/bin/bash
for g in 24 32 48 64 96; do
  MTM_FUSED_BLOCKS_PER_SM=$g python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('blocks/SM', $g, 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
done
