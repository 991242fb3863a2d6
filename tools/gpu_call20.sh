#!/bin/bash
# where does the per-stage fixed cost of the persistent conv kernel go?  spin waits, bulk x rows, phase counters of the empty pipeline
set -u
O=gpurun_out; mkdir -p $O
DIAG_PROBES=0,496,512,1008,1024,1536,256 timeout 300 python tools/diag_ablate.py > $O/c20_ablate.out 2>&1; echo "ablate rc=$?"; cut -c1-600 $O/c20_ablate.out
DIAG_VARIANTS=0,496,512,1008,1024,1536 timeout 200 python tools/diag_phases2.py 2>&1 | cut -c1-420 > $O/c20_phases.out; cat $O/c20_phases.out
AVC_T2_VARIANT=512 timeout 600 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_model.py > $O/c20_tests_spin.out 2>&1; echo "tests(spin) rc=$?"; tail -2 $O/c20_tests_spin.out
for v in 0 512; do
  AVC_T2_VARIANT=$v timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c20_bench_v$v.json 2> $O/c20_bench_v$v.err; echo "bench v=$v rc=$?"
done
python - <<'PY'
import json
for f in ("gpurun_out/c20_bench_v0.json", "gpurun_out/c20_bench_v512.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
