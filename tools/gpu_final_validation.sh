#!/bin/bash
# final validation of the round's last tree: full GPU suite, smoke, default bench line
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest -q -x -m gpu -p no:cacheprovider tests > $O/final_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/final_tests.out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/final_smoke.out 2>&1; echo "smoke rc=$?"; tail -1 $O/final_smoke.out
timeout 500 python bench.py ${BENCH_FLAGS:-} > $O/final_bench.json 2> $O/final_bench.err; echo "bench rc=$?"; tail -2 $O/final_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), d["last_losses"], {k: (round(v.get("value", 0)), round(v.get("e2e", {}).get("value", 0))) for k, v in d.get("extras", {}).items()}, d.get("cpu_baseline", {}).get("value"), d["clocks"], d["timing"]["window_ms"])
PY
