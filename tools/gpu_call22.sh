#!/bin/bash
# per-stage patch-warp ownership: correctness first (abort on failure), then ablation vs the single-owner structure, bench
set -u
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_model.py > $O/c22_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/c22_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/c22_tests.out | head -20; exit 1; fi
DIAG_PROBES=0,496,2048,2544 timeout 120 python tools/diag_ablate.py > $O/c22_ablate.out 2>&1; echo "ablate rc=$?"; cut -c1-500 $O/c22_ablate.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c22_bench.json 2> $O/c22_bench.err; echo "bench rc=$?"
AVC_T2_HS=1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c22_bench_hs1.json 2> $O/c22_bench_hs1.err; echo "bench hs1 rc=$?"
AVC_T2_RPAD=4 DIAG_PROBES=0,496 timeout 100 python tools/diag_ablate.py 2>&1 | grep "in_conv\|status" | cut -c1-300
python - <<'PY'
import json
for f in ("gpurun_out/c22_bench.json", "gpurun_out/c22_bench_hs1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
