#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 100 python tools/diag_store.py > $O/c10_store.out 2>&1; echo "store rc=$?"; cat $O/c10_store.out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_wgrad_acc.py > $O/c10_kern.out 2>&1; echo "kernel tests rc=$?"; tail -4 $O/c10_kern.out
timeout 200 python tools/diag_wgrad.py > $O/c10_wgrad.out 2>&1; echo "wgrad rc=$?"; cat $O/c10_wgrad.out | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c10_bench.json 2> $O/c10_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c10_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
