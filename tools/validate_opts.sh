#!/bin/bash
# One-shot B200 validation of the opt-in paths (run through gpurun from the repo root):
#   uniform tcgen05 issue loops + unrolled wgrad reduction (AVC_TC_ISSUE / AVC_WGRAD_REDUCE)
#   programmatic dependent launch build (AVC_PDL=1)
# Every step is bounded by its own timeout and writes to gpurun_out/ as it goes.
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 20 --warmup 5 --skip-cpu"
run() { # name, env..., -- cmd
  local name=$1; shift
  local t0=$(date +%s%N)
  ( "$@" ) > $O/$name.out 2> $O/$name.err
  echo "$name rc=$? $(( ($(date +%s%N) - t0) / 1000000 )) ms" >> $O/val_summary.txt
}
: > $O/val_summary.txt
run val_bench_base      env timeout 90 $B
run val_tests_uniform   env AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2 timeout 150 python -m pytest tests -m gpu -x -q
run val_bench_uniform   env AVC_TC_ISSUE=uniform timeout 90 $B
run val_bench_uniform_v2 env AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2 timeout 90 $B
run val_tests_pdl       env AVC_PDL=1 AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2 timeout 150 python -m pytest tests -m gpu -x -q
run val_bench_pdl       env AVC_PDL=1 timeout 90 $B
run val_bench_all       env AVC_PDL=1 AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2 timeout 90 $B
cat $O/val_summary.txt
for f in base uniform uniform_v2 pdl all; do python - <<PY
import json
try:
    d = json.loads(open("$O/val_bench_$f.out").read().strip().splitlines()[-1])
    print("$f", round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["achieved"])
except Exception as e:
    print("$f", "no bench line:", e)
PY
done
tail -3 $O/val_tests_uniform.out; tail -3 $O/val_tests_pdl.out
