#!/bin/bash
# Round-2 profiling in ONE gpurun call (1 GPU; ncu replays kernels, never run it multi-rank):
#   gpurun --timeout 1200 -- 'bash tools/profile_r2.sh'
# .ncu-rep files stay in /tmp on the box (they exceed the 64 MiB return limit); reduced CSV / JSON extracts come back
# in gpurun_out/ and are copied to profiles/ by hand.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-extras --windows 1"
# 1. launch list: per-launch durations of the first eager steps (cold-cache, serialised: compare SHARES)
timeout 300 $NCU --metrics gpu__time_duration.sum -c 1000 --csv --log-file $O/r2_launches.csv $BENCH > $O/r2_launches_bench.out 2> $O/r2_launches_bench.err
echo "launch list rc=$?"
# 2. full capture of the dominant kernel at the roofline shape (B=256, 128->128, k5, T=128, IN+ReLU, saves c), rotating > L2 inputs
timeout 200 $NCU --set full --import-source on -k regex:conv_block_tc2 -s 3 -c 2 -o /tmp/r2_conv_roofline -f python tools/ncu_conv_tc.py 128 > $O/r2_ncu_conv.out 2>&1
echo "roofline capture rc=$?"
ncu -i /tmp/r2_conv_roofline.ncu-rep --page raw --csv > $O/r2_conv_roofline_raw.csv 2>/dev/null
# 3. full captures of IN-STEP launches of the three dominant kernels: skip two eager steps of them (2 x 205), then one
#    forward pass of conv blocks and the first part of the backward pass (data gradients, norm backward, weight gradients)
timeout 600 $NCU --set full -k regex:"conv_block_tc2|conv_wgrad_tc|norm_bwd_cached" -s 410 -c 130 -o /tmp/r2_instep -f $BENCH > $O/r2_ncu_instep.out 2>&1
echo "in-step capture rc=$?"
ncu -i /tmp/r2_instep.ncu-rep --page raw --csv > $O/r2_instep_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep
python tools/profile_reduce.py
du -sh $O
