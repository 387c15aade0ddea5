#!/bin/bash
# One GPU box, one GPU: the GPU test suite, the timed EM iteration on a 1.25 M shard (one rank's share at 8 GPUs)
# and the default bench line.  Usage (from the repo root): gpurun -- 'bash tools/gpu_round_check.sh TAG'
tag=${1:-check}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
python bench.py --trajectories 1250000 --steps 40 --warmup 5 --legs "" --no-e2e --no-cpu-baseline \
    > gpurun_out/${tag}_bench_shard.json 2> gpurun_out/${tag}_bench_shard.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
echo "N=1 rc=$?"
tail -3 gpurun_out/${tag}_bench_n1.err
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
for name in ("shard", "n1"):
    d = json.load(open(f"gpurun_out/{tag}_bench_{name}.json"))
    s = d["statistics"]
    print(name, round(d["ms_per_step"], 4), f"{d['value']:.4e}", round(d["roofline"]["frac"], 4),
          round(d["roofline"]["kernel_ms"], 4), s["mode_of_timed_iteration"], round(s["ms_per_step_full_statistics"], 4),
          s["params_max_rel_err_delta_vs_full"], d.get("invalid"), d["gpu_launches"], d["state"]["assign_sha256"][:16],
          (d["e2e"] or {}).get("per_fit", {}).get("ms_total"),
          {k: (v.get("ms") or v.get("ms_per_step") or v.get("seconds_per_sweep_iteration") or v.get("seconds"))
           for k, v in d["configs"].items()})
PY
