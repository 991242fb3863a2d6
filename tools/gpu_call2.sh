#!/bin/bash
# round 2, call 2: batch-dependence diagnostic, full GPU suite with the new defaults, bench lines
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python tools/diag_batchdep.py > $O/c2_batchdep.out 2>&1; echo "batchdep rc=$?"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $O/c2_tests.out 2>&1; echo "tests rc=$?"
tail -15 $O/c2_tests.out
timeout 300 python bench.py --steps 20 --warmup 5 > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc=$?"
timeout 120 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras --e2e-api ae_step > $O/c2_bench_aestep.json 2> $O/c2_bench_aestep.err; echo "bench2 rc=$?"
cat $O/c2_batchdep.out
python - <<'PY'
import json
for f in ("gpurun_out/c2_bench.json", "gpurun_out/c2_bench_aestep.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches/step", d.get("launches_per_step"), "roof", round(d["roofline"]["frac"], 3), d["roofline"]["avg_launch_ms"], d.get("extras"), d.get("cpu_baseline", {}).get("value"), d["timing"])
    except Exception as e:
        print(f, "ERR", e)
PY
