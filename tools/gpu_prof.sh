#!/bin/bash
# Full ncu capture of selected kernels inside a warmed bench run: bash tools/gpu_prof.sh <tag> <kernel-regex> [launch-skip] [count] [bench args...]
tag=$1; rx=$2; skip=${3:-40}; cnt=${4:-8}; shift 4
out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s $skip -c $cnt -f -o $out/${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 "$@" > $out/${tag}.log 2>&1
tail -2 $out/${tag}.log | cut -c1-300
