#!/bin/bash
# usage: tools/compare_libs.sh "<bench args>" lib1.so lib2.so ...   ("default" = in-tree library)
ARGS="$1"; shift
for v in "$@"; do
  if [ "$v" = "default" ]; then unset MTM_B200_LIB; else export MTM_B200_LIB="$PWD/$v"; fi
  python bench.py $ARGS 2>/tmp/err.log | tail -1 > /tmp/line.json
  python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open('/tmp/line.json')); r = d['roofline']
    print(f"{sys.argv[1]:34s} step {d['ms_per_step']:8.3f} ms  E {r['kernel_ms']:8.3f} ms  fin {r['finalize_ms']:7.3f} ms  fp64 frac {r['frac']:.3f}  "
          f"evals/s {d['value']:.4g}  parity {d['parity']['assignments_equal']} {d['parity']['max_rel_ll_err']:.1e}  clk {d['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e, open('/tmp/err.log').read()[-600:])
PY
done
