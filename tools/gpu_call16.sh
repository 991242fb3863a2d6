#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python tools/diag_wgrad.py > $O/c16_wgrad.out 2>&1; echo "wgrad diag rc=$?"; grep -v "round-1  CTAs" $O/c16_wgrad.out | cut -c1-300
timeout 100 python tools/diag_phases2.py 2>&1 | grep -v "variant [13]" | cut -c1-420
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_wgrad_acc.py tests/test_gpu_fold_fused.py tests/test_gpu_normbwd_fused.py tests/test_gpu_model.py tests/test_gpu_kernels.py > $O/c16_tests.out 2>&1; echo "tests rc=$?"; tail -3 $O/c16_tests.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c16_bench.json 2> $O/c16_bench.err; echo "bench rc=$?"
AVC_T2_VARIANT=8 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c16_bench_nopf.json 2> $O/c16_bench_nopf.err; echo "bench no-prefetch rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c16_bench.json", "gpurun_out/c16_bench_nopf.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
