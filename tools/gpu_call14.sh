#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python tools/diag_wgrad.py > $O/c14_wgrad.out 2>&1; echo "wgrad diag rc=$?"; cat $O/c14_wgrad.out | cut -c1-300
AVC_T2_VARIANT=4 timeout 100 python tools/diag_phases2.py 2>&1 | grep -A1 "conv5 T128 IN" | cut -c1-420
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests > $O/c14_tests.out 2>&1; echo "all tests rc=$?"; tail -8 $O/c14_tests.out
AVC_WGRAD_KERNEL=split timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_wgrad_acc.py > $O/c14_split.out 2>&1; echo "split tests rc=$?"; tail -3 $O/c14_split.out
timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu > $O/c14_bench.json 2> $O/c14_bench.err; echo "bench rc=$?"
AVC_WGRAD_KERNEL=split timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c14_bench_split.json 2> $O/c14_bench_split.err; echo "bench split rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c14_bench.json", "gpurun_out/c14_bench_split.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
        if "extras" in d: print("   extras:", {k: (round(v.get("value", 0)), v.get("e2e", {}).get("value")) if isinstance(v, dict) else v for k, v in d["extras"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/c14_bench.err
