#!/bin/bash
# bench.py at N = the arguments given (e.g. "1 2"), --steps 20 --warmup 5, lines into gpurun_out/<tag>_bench_n<N>.json
# Usage (from the repo root): gpurun --gpus 2 -- 'bash tools/gpu_round_check_n.sh TAG 1 2'
tag=$1
shift
mkdir -p gpurun_out
for N in "$@"; do
    if [ "$N" = 1 ]; then
        python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
    else
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
            bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
    fi
    echo "N=$N rc=$?"
    python - "$tag" "$N" <<'PY'
import json, sys
tag, N = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/{tag}_bench_n{N}.json"))
s = d["statistics"]
mg = d["parity"].get("multi_gpu") or {}
print(N, round(d["ms_per_step"], 4), f"{d['value']:.4e}", round(d["roofline"]["frac"], 4), round(d["roofline"]["kernel_ms"], 4),
      s["mode_of_timed_iteration"], round(s["ms_per_step_full_statistics"], 4), s.get("labels_equal_delta_vs_full"),
      s["params_max_rel_err_delta_vs_full"], d.get("invalid"), d["state"]["assign_sha256"][:16],
      mg.get("assignments_equal"), mg.get("params_max_rel_err"), (d["parity"].get("class_sharded") or {}).get("ok"),
      round(d["e2e"]["ms_per_step"], 2), round(d["e2e"]["per_fit"]["ms_total"], 1),
      {k: (v.get("ms") or v.get("ms_per_step") or v.get("seconds_per_sweep_iteration") or v.get("seconds"))
       for k, v in d["configs"].items()})
PY
done
