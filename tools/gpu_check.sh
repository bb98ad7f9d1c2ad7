#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (ours + reference arm), launch list, one full ncu capture of the NN kernel.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn_search_grid -s 20 -c 1 -f -o $out/${tag}_nn \
    python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 > $out/${tag}_ncu_nn.log 2>&1
tail -5 $out/${tag}_pytest_gpu.log; cat $out/${tag}_smoke.log | tail -3; cat $out/${tag}_bench.json; cat $out/${tag}_bench_ref.json
