#!/bin/bash
# NN tuning pass: bash tools/gpu_tune.sh "<env assignments;...>" [bench args]   e.g. "M3DREG_NN_RHO_DIV=8;M3DREG_NN_RHO_DIV=16" "--workload c3"
out=gpurun_out; mkdir -p $out
IFS=';' read -ra CFG <<< "$1"
for cfg in "${CFG[@]}"; do
  echo "== $cfg $2"
  env $cfg timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 1 $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('ms_per_step %.4f  nn_ms %.4f  evals/q %.1f  stages %s  e2e %.3g' % (d['ms_per_step'], r['launch_ms'], r['nn_candidate_evaluations_per_query'], [round(v, 4) for v in r['stage_ms'].values()], d['e2e']['value']))
"
done
