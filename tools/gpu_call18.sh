#!/bin/bash
# re-entry call: ablation probes of the persistent conv kernel, phase counters, full GPU test suite, default bench line
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python tools/diag_ablate.py > $O/c18_ablate.out 2>&1; echo "ablate rc=$?"; cut -c1-900 $O/c18_ablate.out
timeout 150 python tools/diag_phases2.py 2>&1 | grep -v "variant [13]" | cut -c1-420 > $O/c18_phases.out; cat $O/c18_phases.out
timeout 200 python tools/diag_wgrad.py > $O/c18_wgrad.out 2>&1; echo "wgrad diag rc=$?"; grep "split " $O/c18_wgrad.out | cut -c1-300
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider tests > $O/c18_tests.out 2>&1; echo "tests rc=$?"; tail -4 $O/c18_tests.out
timeout 400 python bench.py > $O/c18_bench.json 2> $O/c18_bench.err; echo "bench rc=$?"; tail -2 $O/c18_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/c18_bench.json",):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"], {k: round(v.get("value", 0)) for k, v in d.get("extras", {}).items()}, d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
