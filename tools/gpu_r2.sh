#!/bin/bash
# Round-2 GPU pass: smoke first (a hang costs 2 minutes, not the call), parity tests, bench A/B of the grid paths, launch
# list, full ncu captures.  Usage (under gpurun, repo root): bash tools/gpu_r2.sh <tag> [ncu-kernel-regex]
tag=${1:-r2a}; rx=${2:-k_grid_build|k_normal_equations|k_nn_search_grid}
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv 2>&1
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; rc=$?; echo "smoke rc=$rc" >> $out/${tag}_smoke.log
tail -4 $out/${tag}_smoke.log
if [ $rc -ne 0 ]; then
  echo "smoke failed: trying the single-launch grid path"; M3DREG_GRID_MEGA=1 timeout 180 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
fi
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
tail -15 $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 3000 $out/${tag}_bench.json
M3DREG_GRID_MEGA=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_bench_mega.json 2> $out/${tag}_bench_mega.err
python - <<PY
import json
for f in ("${out}/${tag}_bench.json", "${out}/${tag}_bench_mega.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d["roofline"]
        print(f, "ms/step %.4f" % d["ms_per_step"], {k[:12]: round(v, 4) for k, v in r["stage_ms"].items()}, "launches/step", d["launches_per_step"], "e2e %.3g" % d["e2e"]["value"], d["result"])
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 12 -c 3 -f -o $out/${tag}_prof \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_prof.log 2>&1
tail -2 $out/${tag}_ncu_prof.log | cut -c1-200
