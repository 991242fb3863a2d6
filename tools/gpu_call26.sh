#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
AVC_WGRAD_STREAM=1 timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_dp.py > $O/c26_tests.out 2>&1; rc=$?; echo "tests(wgrad stream) rc=$rc"; tail -3 $O/c26_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/c26_tests.out | head -20; exit 1; fi
AVC_WGRAD_STREAM=1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c26_bench_ws1.json 2> $O/c26_bench_ws1.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c26_bench_ws1.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d["last_losses"], d["timing"]["window_ms"])
    except Exception as e:
        print(f, "ERR", e)
PY
