#!/bin/bash
# Round-2 profiling refresh in ONE gpurun call (1 GPU; ncu replays kernels, never run it multi-rank):
#   gpurun --timeout 600 -- 'bash tools/profile_r2.sh'
# Produces under gpurun_out/: the launch list of ~5 train steps (shares of the step), an ncu --set full
# capture of the dominant conv block and of the weight-gradient kernel, the clock64 phase counters,
# and a clean (un-profiled) bench line to quote beside them.  Copy the summaries to profiles/.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# 1. launch list: per-launch durations of 5 eager steps (cold-cache, serialised: compare SHARES)
timeout 300 $NCU --metrics gpu__time_duration.sum -c 1800 --csv --log-file $O/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu > $O/r2_launches_bench.out 2> $O/r2_launches_bench.err
# 2. full capture of the fused conv block at the roofline shape (B=256, 128->128, k=5, T=128, IN+ReLU, saves c)
timeout 200 $NCU --set full --import-source on -k regex:conv_block_tc -c 3 -o $O/r2_conv_block_tc -f \
    python tools/ncu_conv_tc.py > $O/r2_ncu_conv.out 2>&1
ncu -i $O/r2_conv_block_tc.ncu-rep --page raw --csv > $O/r2_conv_block_tc_raw.csv 2>/dev/null
# 3. phase counters of the conv block (main loop / MMA issue / data wait / epilogue)
timeout 120 python tools/diag_phases.py > $O/r2_phases.out 2>&1
# 4. the number itself, not under a profiler
timeout 120 python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err
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
    for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:14]:
        print(f"{100 * v / s:5.1f}%  {v / 1e6:8.3f} ms  {n:5d} launches  {v / n / 1e3:7.1f} us  {k}")
PY
tail -1 $O/r2_bench.json | cut -c1-400
