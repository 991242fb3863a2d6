#!/bin/bash
# Round-2 first GPU call: validate what was written without a GPU at the end of round 1
# (gpurun --timeout 200 -- 'bash tools/validate_opts_r2.sh').  Same recipe as tools/validate_opts.sh.
#   W = AVC_WGRAD_ACC=1   conv weight gradients accumulated in place (vector atomics) + one flush launch
#   D = AVC_FOLD_FUSED=1  reflect-padding / residual adjoint inside the data-gradient conv epilogue
set -u
O=gpurun_out
mkdir -p $O
: > $O/val2_summary.txt
W="AVC_WGRAD_ACC=1"
D="AVC_FOLD_FUSED=1"
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
  echo "$name rc=$rc $(( ($(date +%s%N) - t0) / 1000000 ))ms" | tee -a $O/val2_summary.txt
}
run t2_W    120 $W $X -- $TESTS tests/test_gpu_wgrad_acc.py tests/test_gpu_tc_conv.py tests/test_gpu_model.py
run b2_W     60 $W    -- $BENCH
run b2_base  60       -- $BENCH
run t2_D    120 $D $X -- $TESTS tests/test_gpu_fold_fused.py tests/test_gpu_model.py
run b2_D     60 $D    -- $BENCH
run b2_WD    60 $W $D -- $BENCH
run t2_props 120 $X   -- $TESTS tests/test_gpu_properties.py
run b2_prefetch 60    -- $BENCH --e2e-api run_steps
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/b2_*.out")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(os.path.basename(f), round(d["value"]), "seg/s  e2e", round(d["e2e"]["value"]), " ms/step", round(d["ms_per_step"], 3), " launches", d.get("gpu_launches"))
    except Exception as e:
        print(os.path.basename(f), "no bench line:", repr(e)[:80])
for f in sorted(glob.glob("gpurun_out/t2_*.out")):
    lines = open(f).read().strip().splitlines()
    print(os.path.basename(f), lines[-1] if lines else "(empty)")
PY
