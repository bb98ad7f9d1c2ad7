#!/bin/bash
# C5-shaped strong-scaling set: registerAll over S rotating-SICK scans x 1 048 576, ICP and NDT sweeps alternating.
# bash tools/gpu_c5.sh <tag> <n_gpus> <n_scans>   (under gpurun --gpus <n_gpus> for n_gpus > 1)
tag=$1; n=$2; S=$3; out=gpurun_out; mkdir -p $out
ARGS="--gpus $n --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --slam c5 --slam-c5-scans $S --slam-sweeps 6"
if [ "$n" -gt 1 ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py $ARGS > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.err
else
  timeout 1200 python bench.py $ARGS > $out/${tag}_c5_n$n.json 2> $out/${tag}_c5_n$n.err
fi
python - <<PY
import json
d = json.loads(open("$out/${tag}_c5_n$n.json").read().strip().splitlines()[-1])
c5 = (d.get("slam") or {}).get("c5") or {}
print("N=%d" % d["n_gpus"], {k: c5.get(k) for k in ("n_scans", "pairs", "pairs_rank0", "ms_per_sweep", "scans_per_s", "points_per_s", "accumulate_ms_max_rank", "allreduce_wait_ms_max_rank", "sharded_vs_single_rank", "relative_pose_error_m", "solved_scans", "scan_generation_and_upload_s")})
PY
tail -n 3 $out/${tag}_c5_n$n.err
