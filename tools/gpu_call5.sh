#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/diag_phases2.py > $O/c5_phases.out 2>&1; echo "phases rc=$?"; cat $O/c5_phases.out | cut -c1-420
for v in 0 3; do
  AVC_T2_VARIANT=$v timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py > $O/c5_tcconv_v$v.out 2>&1; echo "tc conv variant $v rc=$?"; tail -4 $O/c5_tcconv_v$v.out
done
for v in 0 1 3; do
  AVC_T2_VARIANT=$v timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c5_bench_v$v.json 2> $O/c5_bench_v$v.err; echo "bench variant $v rc=$?"
done
python - <<'PY'
import json
for v in (0, 1, 3):
    f = f"gpurun_out/c5_bench_v{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
