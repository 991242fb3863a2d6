#!/bin/bash
# round-robin patch warps: correctness, ablation vs the single-owner structure (bit 2048), hs = 1 revisited, row padding of the 1x1 layers, bench
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest -q -x -m gpu -p no:cacheprovider tests > $O/c21_tests.out 2>&1; echo "tests rc=$?"; tail -3 $O/c21_tests.out
for hs in 0 1; do
  AVC_T2_HS=$hs DIAG_PROBES=0,496,2048,2544,256 timeout 300 python tools/diag_ablate.py > $O/c21_ablate_hs$hs.out 2>&1; echo "ablate hs=$hs rc=$?"; cut -c1-500 $O/c21_ablate_hs$hs.out
  AVC_T2_HS=$hs timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c21_bench_hs$hs.json 2> $O/c21_bench_hs$hs.err; echo "bench hs=$hs rc=$?"
done
AVC_T2_RPAD=4 DIAG_PROBES=0,496 timeout 300 python tools/diag_ablate.py > $O/c21_ablate_rpad4.out 2>&1; echo "ablate rpad rc=$?"; grep in_conv $O/c21_ablate_rpad4.out | cut -c1-500
AVC_T2_RPAD=4 timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_model.py > $O/c21_tests_rpad.out 2>&1; echo "tests(rpad) rc=$?"; tail -2 $O/c21_tests_rpad.out
DIAG_VARIANTS=0,496 timeout 200 python tools/diag_phases2.py 2>&1 | cut -c1-420 > $O/c21_phases.out; cat $O/c21_phases.out
python - <<'PY'
import json
for f in ("gpurun_out/c21_bench_hs0.json", "gpurun_out/c21_bench_hs1.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
