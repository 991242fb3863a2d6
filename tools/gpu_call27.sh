#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py -k "infer" > $O/c27_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -3 $O/c27_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/c27_tests.out | head -20; exit 1; fi
for cfg in "1 1" "1 0" "0 0"; do
  set -- $cfg
  AVC_INFER_GRAPH=$1 AVC_OVERLAP=$2 timeout 120 python bench.py --workload inference --steps 20 --warmup 3 --skip-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graph=$1 overlap=$2', round(d['value']), 'utt/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value']))"
done
