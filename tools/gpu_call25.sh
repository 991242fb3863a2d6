#!/bin/bash
# weight gradients forked onto their own stream: tests first (abort on failure), then bench with / without
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_dp.py tests/test_gpu_properties.py > $O/c25_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/c25_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/c25_tests.out | head -20; exit 1; fi
for ws in 1 0; do
  AVC_WGRAD_STREAM=$ws timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c25_bench_ws$ws.json 2> $O/c25_bench_ws$ws.err; echo "bench wgrad_stream=$ws rc=$?"
done
python - <<'PY'
import json
for f in ("gpurun_out/c25_bench_ws1.json", "gpurun_out/c25_bench_ws0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], d["last_losses"], d["timing"]["window_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
