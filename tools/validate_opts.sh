#!/bin/bash
# One-shot B200 validation of the opt-in paths (run through gpurun from the repo root), most
# valuable step first; every step is bounded by its own timeout and writes to gpurun_out/ as it goes.
#   A = AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2   uniform-datapath tcgen05 issue loops, unrolled wgrad reduce
#   P = AVC_PDL=1                                    programmatic-dependent-launch build of the library
#   F = AVC_FUSED_DENSE=1                            fused speaker dense stack + batched AdaIN affine layers
set -u
O=gpurun_out
mkdir -p $O
: > $O/val_summary.txt
A="AVC_TC_ISSUE=uniform AVC_WGRAD_REDUCE=v2"
P="AVC_PDL=1"
F="AVC_FUSED_DENSE=1"
X="AVC_TEST_EXPERIMENTAL=1"
BENCH="python bench.py --steps 20 --warmup 5 --skip-cpu"
TESTS="python -m pytest -q -m gpu -p no:cacheprovider"
run() {  # run <name> <timeout> <env...> -- <cmd...>
  local name=$1 to=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  local t0=$(date +%s%N)
  env "${envs[@]}" timeout "$to" "$@" > $O/$name.out 2> $O/$name.err
  local rc=$?
  echo "$name rc=$rc $(( ($(date +%s%N) - t0) / 1000000 ))ms" | tee -a $O/val_summary.txt
}
run t_APF  120 $A $P $F $X -- $TESTS tests
run b_APF   60 $A $P $F    -- $BENCH
run t_A     90 $A $X       -- $TESTS tests/test_gpu_tc.py tests/test_gpu_tc_conv.py tests/test_gpu_model.py
run b_A     60 $A          -- $BENCH
run t_AF    90 $A $F $X    -- $TESTS tests/test_gpu_fused_dense.py tests/test_gpu_model.py
run b_AF    60 $A $F       -- $BENCH
run b_AP    60 $A $P       -- $BENCH
run b_base  60             -- $BENCH
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/b_*.out")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "seg/s  e2e", round(d["e2e"]["value"]), " ms/step", round(d["ms_per_step"], 3),
              " conv_tc", d["roofline"].get("achieved"), d["roofline"].get("unit"), " launches", d.get("gpu_launches"))
    except Exception as e:
        print(os.path.basename(f), "no bench line:", repr(e)[:80])
for f in sorted(glob.glob("gpurun_out/t_*.out")):
    lines = open(f).read().strip().splitlines()
    print(os.path.basename(f), lines[-1] if lines else "(empty)")
PY
