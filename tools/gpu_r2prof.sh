#!/bin/bash
# Round-2 profile pass: bench line, launch list (ncu time metric only) and full ncu captures of one steady-state iteration.
# Usage (under gpurun, repo root): bash tools/gpu_r2prof.sh <tag> [bench args]
tag=${1:-r2m}; shift
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --steps 30 --warmup 5 --slam none "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 2500 $out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --slam none "$@" > $out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 60 -c 12 -f -o $out/${tag}_prof \
    python bench.py --steps 6 --warmup 6 --no-cpu-baseline --e2e-steps 1 --slam none "$@" > $out/${tag}_ncu_prof.log 2>&1
tail -2 $out/${tag}_ncu_prof.log | cut -c1-200
