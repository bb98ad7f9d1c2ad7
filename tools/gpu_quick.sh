#!/bin/bash
# Quick GPU pass: parity tests, bench, launch list and one full ncu capture of one kernel.  bash tools/gpu_quick.sh <tag> [kernel-regex] [skip]
tag=${1:-q}; rx=${2:-k_nn_search_grid}; skip=${3:-20}
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $out/${tag}_prof \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_prof.log 2>&1
tail -4 $out/${tag}_pytest_gpu.log; cut -c1-2500 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
