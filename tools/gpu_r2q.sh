#!/bin/bash
# Round-2 check pass: parity suite, default bench line (ours + reference arm), NDT line, C5-shaped sweep record.
tag=${1:-r2q}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log; tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log; tail -2 $out/${tag}_smoke.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 1500 $out/${tag}_bench.json; echo
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; cut -c1-300 $out/${tag}_bench_ref.json
timeout 600 python bench.py --mode ndt --no-cpu-baseline --e2e-steps 1 --slam none > $out/${tag}_bench_ndt.json 2> $out/${tag}_bench_ndt.err
timeout 900 python bench.py --no-cpu-baseline --e2e-steps 1 --slam c5 --slam-c5-scans 16 --slam-sweeps 4 > $out/${tag}_bench_c5.json 2> $out/${tag}_bench_c5.err
python - <<PY
import json
for f in ("bench", "bench_ndt", "bench_c5"):
    try:
        d = json.loads(open("$out/${tag}_%s.json" % f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, "ms/step %.4f" % d["ms_per_step"], r["kernel"][:40], "frac %.3f" % r["frac"], "clocks", d["clocks"], "schedule", d.get("schedule"))
        if d.get("slam"): print("   slam", json.dumps(d["slam"])[:1500])
    except Exception as e:
        print(f, "unreadable", e)
PY
