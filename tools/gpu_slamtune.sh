#!/bin/bash
# C4 sweep under different search heuristics: bash tools/gpu_slamtune.sh "<env;env;...>" [slam_bench args]
IFS=';' read -ra CFG <<< "$1"
for cfg in "${CFG[@]}"; do
  echo "== $cfg $2"
  env $cfg timeout 300 python tools/slam_bench.py --sweeps 4 $2 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ('ms_per_sweep', 'scans_per_s', 'points_per_s', 'pairs', 'nn_evaluations_per_query', 'nn_fallback_share')})
"
done
