#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider -x tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py tests/test_gpu_kernels.py > $O/c8_tcconv.out 2>&1; echo "tc conv rc=$?"; tail -8 $O/c8_tcconv.out
timeout 200 python tools/diag_phases2.py > $O/c8_phases.out 2>&1; echo "phases rc=$?"; grep -v "variant [13]" $O/c8_phases.out | cut -c1-420
timeout 200 python tools/diag_wgrad.py > $O/c8_wgrad.out 2>&1; echo "wgrad rc=$?"; cat $O/c8_wgrad.out | cut -c1-300
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c8_bench.json 2> $O/c8_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c8_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
except Exception as e:
    print("ERR", e)
PY
tail -3 $O/c8_bench.err
