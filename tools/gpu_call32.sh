#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 400 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_model.py tests/test_gpu_dp.py tests/test_gpu_properties.py > $O/c32_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -2 $O/c32_tests.out
if [ $rc -ne 0 ]; then grep -n "Error\|assert" $O/c32_tests.out | head -20; exit 1; fi
timeout 200 python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('affine on side stream:', round(d['value']), 'seg/s', round(d['ms_per_step'],3), 'ms', d['timing']['window_ms'], d['last_losses'])"
