#!/bin/bash
# Round-2 profiling in ONE gpurun call (1 GPU; ncu replays kernels, never run it multi-rank):
#   gpurun --timeout 1500 -- 'bash tools/profile_r2.sh'
# .ncu-rep files go to /tmp on the box (they exceed the 64 MiB return limit); only CSV extracts come back in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-extras --windows 1"
# 1. launch list: per-launch durations of ~5 eager steps (cold-cache, serialised: compare SHARES)
timeout 400 $NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/r2_launches.csv $BENCH > $O/r2_launches_bench.out 2> $O/r2_launches_bench.err
# 2. full capture of the dominant kernel at the roofline shape (B=256, 128->128, k5, T=128, IN+ReLU, saves c), rotating > L2 inputs
timeout 300 $NCU --set full -k regex:conv_block_tc2 -s 3 -c 2 -o /tmp/r2_conv_roofline -f python tools/ncu_conv_tc.py 128 > $O/r2_ncu_conv.out 2>&1
ncu -i /tmp/r2_conv_roofline.ncu-rep --page raw --csv > $O/r2_conv_roofline_raw.csv 2>/dev/null
# 3. full captures of IN-STEP launches (third eager step: skip two steps of launches of each kernel)
timeout 400 $NCU --set full -k regex:conv_block_tc2 -s 250 -c 40 -o /tmp/r2_conv_instep -f $BENCH > $O/r2_ncu_conv_instep.out 2>&1
ncu -i /tmp/r2_conv_instep.ncu-rep --page raw --csv > $O/r2_conv_instep_raw.csv 2>/dev/null
timeout 300 $NCU --set full -k regex:"conv_wgrad|norm_bwd" -s 200 -c 16 -o /tmp/r2_wgrad_norm -f $BENCH > $O/r2_ncu_wgrad.out 2>&1
ncu -i /tmp/r2_wgrad_norm.ncu-rep --page raw --csv > $O/r2_wgrad_norm_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep
# 4. the number itself, not under a profiler
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err
python - <<'PY'
import csv, collections, re
rows = list(csv.reader(l for l in open("gpurun_out/r2_launches.csv") if l.startswith('"')))
if rows:
    h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ik])[:90]
        tot[k][0] += v; tot[k][1] += 1
    s = sum(v[0] for v in tot.values())
    for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:18]:
        print(f"{100 * v / s:5.1f}%  {v / 1e6:8.3f} ms  {n:5d} launches  {v / n / 1e3:7.1f} us  {k}")
PY
tail -1 $O/r2_bench.json | cut -c1-300
du -sh $O
