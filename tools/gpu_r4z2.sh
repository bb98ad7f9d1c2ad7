#!/bin/bash
# Final bench lines of round 2 (ours + reference arm), launch list, one full ncu capture of a steady-state iteration.
tag=${1:-r4z}; out=gpurun_out; mkdir -p $out
t0=$(date +%s); timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench wall $(( $(date +%s) - t0 )) s"
t0=$(date +%s); timeout 600 python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "reference arm wall $(( $(date +%s) - t0 )) s"; cut -c1-300 $out/${tag}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 60 -c 10 -f -o $out/${tag}_prof \
    python bench.py --steps 6 --warmup 6 --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_ncu_prof.log 2>&1
tail -2 $out/${tag}_ncu_prof.log | cut -c1-200
python - <<PY
import json
d = json.loads(open("$out/${tag}_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("ms/step %.4f value %.4g e2e %.4g" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), r["kernel"], "frac %.4f iter frac %.4f" % (r["frac"], r["iteration"]["frac"]),
      "stages", [round(v, 4) for v in r["stage_ms"].values()], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "launches", d["gpu_launches"])
print("schedule", d.get("schedule"))
PY
