#!/bin/bash
# two-stream encoder overlap: model / DP / graph tests first (abort on failure), then bench with and without
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_dp.py tests/test_gpu_properties.py > $O/ab_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/ab_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/ab_tests.out | head -20; exit 1; fi
for ov in 1 0; do
  AVC_OVERLAP=$ov timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/ab_bench_ov$ov.json 2> $O/ab_bench_ov$ov.err; echo "bench overlap=$ov rc=$?"
done
python - <<'PY'
import json
for f in ("gpurun_out/ab_bench_ov1.json", "gpurun_out/ab_bench_ov0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"], d["timing"]["window_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
