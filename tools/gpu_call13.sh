#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 200 compute-sanitizer --tool memcheck --print-limit 8 python tools/diag_wgrad_one.py > $O/c13_san.out 2>&1; echo "sanitizer rc=$?"; grep -v "^$" $O/c13_san.out | head -60 | cut -c1-260
AVC_WGRAD_TMA_SWZ=0 timeout 100 python tools/diag_wgrad_one.py > $O/c13_noswz.out 2>&1; echo "no-swizzle rc=$?"; tail -4 $O/c13_noswz.out | cut -c1-200
# conv kernel: merged store sweep
AVC_WGRAD_STAGE=cpasync timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_tc_conv.py tests/test_gpu_fold_fused.py tests/test_gpu_normbwd_fused.py > $O/c13_kern.out 2>&1; echo "conv tests rc=$?"; tail -5 $O/c13_kern.out
AVC_WGRAD_STAGE=cpasync timeout 200 python tools/diag_phases2.py > $O/c13_phases.out 2>&1; echo "phases rc=$?"; grep -v "variant [13]" $O/c13_phases.out | cut -c1-420
AVC_WGRAD_STAGE=cpasync timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c13_bench.json 2> $O/c13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/c13_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
