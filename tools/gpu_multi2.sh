#!/bin/bash
# Round-2 multi-GPU pass (run under: gpurun --gpus N -- bash tools/gpu_multi2.sh <tag> N): the driver's bench launch at N ranks
# (C2 headline + C4 sweep record, strong scaling) and the 2-rank NCCL tests.
tag=$1; n=$2; out=gpurun_out; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name --format=csv > $out/${tag}_smi.csv
timeout 600 python -m pytest tests -m gpu -q -k "nccl or two_rank or multi" > $out/${tag}_pytest_n$n.log 2>&1; tail -3 $out/${tag}_pytest_n$n.log
timeout 900 $TR bench.py --gpus $n --steps 30 --warmup 5 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench_n$n.json").read().strip().splitlines()[-1])
print("N=%d value %.3g pt/s ms/step %.4f e2e %.3g allreduce_check %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["allreduce_check"]))
c4 = (d.get("slam") or {}).get("c4") or {}
print("   C4:", {k: c4.get(k) for k in ("ms_per_sweep", "scans_per_s", "pairs_rank0", "accumulate_ms_max_rank", "allreduce_wait_ms_max_rank", "sharded_vs_single_rank")})
PY
tail -n 3 $out/${tag}_bench_n$n.err
