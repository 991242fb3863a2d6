#!/bin/bash
# round 2, call 7: correctness of the reworked persistent kernel, phases, ncu stall profile (small outputs only!)
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py tests/test_gpu_tc.py > $O/c7_tcconv.out 2>&1; echo "tc conv rc=$?"; tail -6 $O/c7_tcconv.out
timeout 200 python tools/diag_phases2.py > $O/c7_phases.out 2>&1; echo "phases rc=$?"; cat $O/c7_phases.out | cut -c1-420
NCU="ncu --clock-control none"
for T in 128 16; do
  timeout 300 $NCU --set full --import-source on -k regex:conv_block_tc2 -s 3 -c 2 -o /tmp/c7_conv_T$T -f python tools/ncu_conv_tc.py $T > $O/c7_ncu_T$T.out 2>&1; echo "ncu T=$T rc=$?"
  ncu -i /tmp/c7_conv_T$T.ncu-rep --page raw --csv > $O/c7_conv_T${T}_raw.csv 2>/dev/null
  ncu -i /tmp/c7_conv_T$T.ncu-rep --page source --csv --print-source sass > $O/c7_conv_T${T}_sass.csv 2>/dev/null
  ls -la /tmp/c7_conv_T$T.ncu-rep $O/c7_conv_T${T}_sass.csv
done
timeout 600 python tools/diag_tf32.py > $O/c7_tf32.out 2>&1; echo "tf32 rc=$?"; tail -12 $O/c7_tf32.out
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c7_bench.json 2> $O/c7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/c7_bench.json").read().strip().splitlines()[-1])
    print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
except Exception as e:
    print("ERR", e)
PY
du -sh $O
