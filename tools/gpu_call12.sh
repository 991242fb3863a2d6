#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python tools/diag_wgrad.py > $O/c12_wgrad.out 2>&1; echo "wgrad diag rc=$?"; cat $O/c12_wgrad.out | cut -c1-300
timeout 400 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_wgrad_acc.py > $O/c12_kern.out 2>&1; echo "kernel tests rc=$?"; tail -6 $O/c12_kern.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c12_bench.json 2> $O/c12_bench.err; echo "bench rc=$?"
AVC_WGRAD_STAGE=cpasync timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c12_bench_cpasync.json 2> $O/c12_bench_cpasync.err; echo "bench cpasync rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c12_bench.json", "gpurun_out/c12_bench_cpasync.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/c12_bench.err
