#!/bin/bash
# half-slab pipeline stages of the persistent conv kernel: correctness, ablation timing old (AVC_T2_HS=2) vs new, bench
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest -q -x -m gpu -p no:cacheprovider tests > $O/c19_tests.out 2>&1; echo "tests rc=$?"; tail -4 $O/c19_tests.out
for hs in 0 2; do
  AVC_T2_HS=$hs timeout 300 python tools/diag_ablate.py > $O/c19_ablate_hs$hs.out 2>&1; echo "ablate hs=$hs rc=$?"; cut -c1-400 $O/c19_ablate_hs$hs.out
  AVC_T2_HS=$hs timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/c19_bench_hs$hs.json 2> $O/c19_bench_hs$hs.err; echo "bench hs=$hs rc=$?"
done
AVC_T2_HS=0 timeout 150 python tools/diag_phases2.py 2>&1 | grep -v "variant [13]" | cut -c1-420 > $O/c19_phases.out; cat $O/c19_phases.out
python - <<'PY'
import json
for f in ("gpurun_out/c19_bench_hs0.json", "gpurun_out/c19_bench_hs2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us", d["last_losses"])
    except Exception as e:
        print(f, "ERR", e)
PY
