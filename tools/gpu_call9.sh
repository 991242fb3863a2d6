#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py tests/test_gpu_kernels.py tests/test_gpu_wgrad_acc.py > $O/c9_kern.out 2>&1; echo "kernel tests rc=$?"; tail -8 $O/c9_kern.out
timeout 200 python tools/diag_phases2.py > $O/c9_phases.out 2>&1; echo "phases rc=$?"; grep -v "variant [13]" $O/c9_phases.out | cut -c1-420
timeout 200 python tools/diag_wgrad.py > $O/c9_wgrad.out 2>&1; echo "wgrad rc=$?"; cat $O/c9_wgrad.out | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c9_bench.json 2> $O/c9_bench.err; echo "bench rc=$?"
AVC_T2_VARIANT=1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c9_bench_v1.json 2> $O/c9_bench_v1.err; echo "bench variant1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c9_bench.json", "gpurun_out/c9_bench_v1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/c9_bench.err
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests --deselect tests/test_gpu_tc_conv.py --deselect tests/test_gpu_fold_fused.py --deselect tests/test_gpu_kernels.py --deselect tests/test_gpu_wgrad_acc.py > $O/c9_tests.out 2>&1; echo "other tests rc=$?"; tail -12 $O/c9_tests.out
