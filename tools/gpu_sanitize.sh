#!/bin/bash
# compute-sanitizer passes over tools/sanitize_small.py (bash tools/gpu_sanitize.sh <tag>) + the new GPU tests
tag=${1:-r2s}; out=gpurun_out; mkdir -p $out
{
echo "compute-sanitizer on tools/sanitize_small.py (fused ICP, batched sweep, stage-level NN, m3dreg_slam_sweep, NDT, pre-registration steps, yaw sweep, node replay), build $tag"
for tool in memcheck racecheck initcheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_small.py 2>&1 | grep -v "^=========$" | grep -E "^icp|^sweep|^nn|^slam|^ndt|^preproc|^yaw|^node|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race|Uninit|hazard|at .* in " | head -40
done
} > $out/${tag}_compute_sanitizer.txt 2>&1
cat $out/${tag}_compute_sanitizer.txt
timeout 600 python -m pytest tests/test_gpu_shim.py tests/test_gpu_preproc.py tests/test_gpu_node.py -x -q 2>&1 | tail -5
timeout 900 python bench.py --impl reference --workload c3 --steps 2 --warmup 1 > $out/${tag}_bench_ref_c3.json 2> $out/${tag}_bench_ref_c3.err; cut -c1-200 $out/${tag}_bench_ref_c3.json
timeout 600 python bench.py --workload c3 --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_bench_c3.json 2> $out/${tag}_bench_c3.err; cut -c1-200 $out/${tag}_bench_c3.json
timeout 600 python bench.py --workload c1 --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_bench_c1.json 2> $out/${tag}_bench_c1.err; cut -c1-200 $out/${tag}_bench_c1.json
timeout 600 python bench.py --impl reference --workload c1 --steps 5 --warmup 2 > $out/${tag}_bench_ref_c1.json 2> $out/${tag}_bench_ref_c1.err; cut -c1-200 $out/${tag}_bench_ref_c1.json
