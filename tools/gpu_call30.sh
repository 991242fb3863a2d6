#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
for ws in 2 1; do
  AVC_WGRAD_STREAM=$ws timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('wgrad_stream=$ws', round(d['value']), 'seg/s', round(d['ms_per_step'],3), 'ms', d['timing']['window_ms'], d['last_losses'])"
done
AVC_WGRAD_STREAM=2 timeout 400 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_dp.py tests/test_gpu_properties.py > $O/c30_tests.out 2>&1; echo "tests(ws=2) rc=$?"; tail -2 $O/c30_tests.out
