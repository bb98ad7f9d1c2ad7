#!/bin/bash
# Final check pass of round 2: parity suite, smoke, the driver's default bench line (ours + reference arm), launch list.
tag=${1:-r4z}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log; tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log; tail -2 $out/${tag}_smoke.log
t0=$(date +%s); timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench wall $(( $(date +%s) - t0 )) s"
t0=$(date +%s); timeout 600 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "reference arm wall $(( $(date +%s) - t0 )) s"; cut -c1-400 $out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_ncu_bench.log 2>&1
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), r["kernel"], "frac %.4f iter frac %.4f" % (r["frac"], r["iteration"]["frac"]),
      "stages", [round(v, 4) for v in r["stage_ms"].values()], "clocks", d["clocks"], "launches", d["gpu_launches"])
print("cpu_baseline", d["cpu_baseline"])
print("slam", json.dumps(d.get("slam"))[:1200])
PY
