#!/bin/bash
# A/B bench of runtime / build variants: bash tools/gpu_ab.sh <tag> "<env assignments;...>" [bench args]
# each configuration: env <cfg> python bench.py --steps 30 --warmup 5 (no CPU baseline, 1 e2e step) -> one summary line
tag=$1; out=gpurun_out; mkdir -p $out
IFS=';' read -ra CFG <<< "$2"
k=0
for cfg in "${CFG[@]}"; do
  k=$((k+1))
  echo "== [$k] $cfg $3"
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 $3 > $out/${tag}_ab$k.json 2> $out/${tag}_ab$k.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/${tag}_ab$k.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("ms/step %.4f  nn %.4f  evals/q %.1f  fallback %.4f  stages %s  err %.5f" % (d["ms_per_step"], r["launch_ms"], r["nn_candidate_evaluations_per_query"],
          r["nn_queries_on_per_thread_fallback"], [round(v, 4) for v in r["stage_ms"].values()], d["result"]["translation_error_m"]))
    if r.get("grid_phases_us"): print("   grid phases us:", {k2: (round(v, 2) if not isinstance(v, dict) else {k3: round(v3, 2) for k3, v3 in v.items()}) for k2, v in r["grid_phases_us"].items()})
except Exception as e:
    print("unreadable:", e); print(open("$out/${tag}_ab$k.err").read()[-1500:])
PY
done
