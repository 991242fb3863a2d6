#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_normbwd_fused.py tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py tests/test_gpu_wgrad_acc.py > $O/c11_kern.out 2>&1; echo "kernel tests rc=$?"; tail -12 $O/c11_kern.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c11_bench.json 2> $O/c11_bench.err; echo "bench rc=$?"
AVC_NORM_BWD_FUSED=0 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c11_bench_nofuse.json 2> $O/c11_bench_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c11_bench.json", "gpurun_out/c11_bench_nofuse.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $O/c11_bench.err
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_properties.py tests/test_gpu_dp.py tests/test_gpu_tf32_accuracy.py > $O/c11_tests.out 2>&1; echo "model tests rc=$?"; tail -12 $O/c11_tests.out
