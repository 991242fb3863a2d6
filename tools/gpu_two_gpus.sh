#!/bin/bash
# 2-GPU sanity of the round's final tree: NCCL DP test + bench at N=2 (torchrun, as the driver launches it)
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest -q -x -m gpu -p no:cacheprovider tests/test_gpu_dp.py > $O/n2_tests.out 2>&1; echo "dp tests (nccl) rc=$?"; tail -2 $O/n2_tests.out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --skip-cpu --skip-extras > $O/n2_bench_n2.json 2> $O/n2_bench_n2.err; echo "bench N=2 rc=$?"; tail -3 $O/n2_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n2_bench_n2.json").read().strip().splitlines()[-1])
print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "n_gpus", d["n_gpus"], d["last_losses"], d["timing"]["window_ms"])
PY
