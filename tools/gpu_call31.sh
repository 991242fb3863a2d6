#!/bin/bash
# re-validation with the decoder's weight gradients on their own stream by default: full suite, smoke, A/B of the fused norm backward, default bench
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests > $O/c31_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/c31_tests.out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/c31_smoke.out 2>&1; echo "smoke rc=$?"; tail -1 $O/c31_smoke.out
AVC_NORM_BWD_FUSED=1 timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('norm_bwd_fused=1', round(d['value']), 'seg/s', round(d['ms_per_step'],3), 'ms launches', d['launches_per_step'], d['timing']['window_ms'])"
timeout 500 python bench.py > $O/c31_bench.json 2> $O/c31_bench.err; echo "bench rc=$?"; tail -2 $O/c31_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c31_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), d["last_losses"], {k: (round(v.get("value", 0)), round(v.get("e2e", {}).get("value", 0))) for k, v in d.get("extras", {}).items()}, d.get("cpu_baseline", {}).get("value"), d["clocks"])
PY
