#!/bin/bash
# round 2, call 3: first run of the persistent conv kernel (conv_tc2.cu)
set -u
O=gpurun_out; mkdir -p $O
T="python -m pytest -q -m gpu -p no:cacheprovider -x"
timeout 120 $T tests/test_gpu_tc.py > $O/c3_tc.out 2>&1; echo "tc probe rc=$?"; tail -3 $O/c3_tc.out
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py > $O/c3_tcconv.out 2>&1; echo "tc conv rc=$?"; tail -12 $O/c3_tcconv.out
timeout 120 python tools/diag_batchdep.py > $O/c3_batchdep.out 2>&1; echo "batchdep rc=$?"; cat $O/c3_batchdep.out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests --deselect tests/test_gpu_tc_conv.py > $O/c3_tests.out 2>&1; echo "tests rc=$?"; tail -12 $O/c3_tests.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c3_bench.json 2> $O/c3_bench.err; echo "bench rc=$?"
AVC_TC_CONV=v1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c3_bench_v1.json 2> $O/c3_bench_v1.err; echo "bench v1 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c3_bench.json", "gpurun_out/c3_bench_v1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches/step", d.get("launches_per_step"), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
tail -5 $O/c3_bench.err
