#!/bin/bash
# Multi-GPU pass (run under: gpurun --gpus N -- bash tools/gpu_multi.sh <tag> N)
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_smi.csv
timeout 600 $TR bench.py --gpus $n --steps 30 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
timeout 900 $TR tools/slam_bench.py --scans 100 --sweeps 3 --check --dof 6 > $out/${tag}_slam_n$n.json 2> $out/${tag}_slam_n$n.err
timeout 900 python tools/slam_bench.py --scans 100 --sweeps 3 --dof 6 > $out/${tag}_slam_n1.json 2> $out/${tag}_slam_n1.err
timeout 900 $TR tools/slam_bench.py --scans 24 --kind sick --sweeps 2 --check --dof 6 > $out/${tag}_slam_sick_n$n.json 2> $out/${tag}_slam_sick_n$n.err
timeout 900 python tools/slam_bench.py --scans 24 --kind sick --sweeps 2 --dof 6 > $out/${tag}_slam_sick_n1.json 2> $out/${tag}_slam_sick_n1.err
for f in $out/${tag}_*.json; do echo "== $f"; cut -c1-900 $f; done
tail -n 3 $out/${tag}_*.err
