#!/bin/bash
# Round-2 profiling in ONE gpurun call (1 GPU; ncu replays kernels, never run it multi-rank):
#   gpurun --timeout 900 -- 'bash tools/profile_r2.sh'
# Produces under gpurun_out/: the launch list of ~4 train steps (shares of the step), an ncu --set full capture of
# in-step launches of the persistent conv kernel / weight-gradient kernel / norm backward, and a clean bench line.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
BENCH="python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-extras --windows 1"
# 1. launch list: per-launch durations of ~4 eager steps (cold-cache, serialised: compare SHARES)
timeout 400 $NCU --metrics gpu__time_duration.sum -c 1300 --csv --log-file $O/r2_launches.csv $BENCH > $O/r2_launches_bench.out 2> $O/r2_launches_bench.err
# 2. full captures of in-step launches (skip the first two steps: ~252 launches each)
timeout 400 $NCU --set full --import-source on -k regex:conv_block_tc2 -s 230 -c 60 -o $O/r2_conv_tc2 -f $BENCH > $O/r2_ncu_conv.out 2>&1
timeout 300 $NCU --set full --import-source on -k regex:"conv_wgrad_tc|norm_bwd" -s 190 -c 12 -o $O/r2_wgrad_norm -f $BENCH > $O/r2_ncu_wgrad.out 2>&1
for f in r2_conv_tc2 r2_wgrad_norm; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
done
# 3. the number itself, not under a profiler
timeout 200 python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err
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
    for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:16]:
        print(f"{100 * v / s:5.1f}%  {v / 1e6:8.3f} ms  {n:5d} launches  {v / n / 1e3:7.1f} us  {k}")
PY
tail -1 $O/r2_bench.json | cut -c1-300
