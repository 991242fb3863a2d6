#!/bin/bash
# final validation of the round: full GPU suite, smoke, ncu capture of in-step weight-gradient launches, default bench line
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests > $O/v_tests.out 2>&1; rc=$?; echo "tests rc=$rc"; tail -4 $O/v_tests.out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/v_smoke.out 2>&1; echo "smoke rc=$?"; tail -2 $O/v_smoke.out
timeout 200 ncu --clock-control none --set full -k regex:"conv_wgrad_split|wgrad_acc_flush" -s 116 -c 16 -o /tmp/r2_wgrad -f python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu --skip-extras --windows 1 > $O/r2_ncu_wgrad.out 2>&1; echo "wgrad capture rc=$?"
ncu -i /tmp/r2_wgrad.ncu-rep --page raw --csv > $O/r2_wgrad_raw.csv 2>/dev/null
python - <<'PY'
import sys
sys.argv = ["x", "gpurun_out"]
sys.path.insert(0, "tools")
import profile_reduce as pr, os, json, csv, re, collections
h, u, rows = pr.read_raw("gpurun_out/r2_wgrad_raw.csv")
out = ["| kernel | grid | avg us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | warps active % | regs |", "|---|---|---|---|---|---|---|---|---|"]
if rows:
    ik = h.index("Kernel Name")
    for r in rows:
        d = {k: pr.num(*v) for k, v in pr.pick(h, u, r).items()}
        out.append(f"| `{re.sub(r'[(<].*', '', r[ik])[:40]}` | {int(d.get('launch__grid_size', 0))} | {d['gpu__time_duration.sum']:.1f} | {d['dram__bytes_read.sum'] / 1e6:.1f} | {d['dram__bytes_write.sum'] / 1e6:.1f} | "
                   f"{d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:.1f} | {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']:.1f} | {d['sm__warps_active.avg.pct_of_peak_sustained_active']:.1f} | {d['launch__registers_per_thread']:.0f} |")
open("gpurun_out/r2_ncu_wgrad_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[:8]))
PY
timeout 500 python bench.py > $O/v_bench.json 2> $O/v_bench.err; echo "bench rc=$?"; tail -2 $O/v_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/v_bench.json").read().strip().splitlines()[-1])
print(round(d["value"]), "seg/s e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "launches", d["launches_per_step"], "roof", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"]*1e3, 2), "us pre", d["roofline"].get("avg_launch_ms_prerounded_input"), d["last_losses"], {k: (round(v.get("value", 0)), round(v.get("e2e", {}).get("value", 0))) for k, v in d.get("extras", {}).items()}, d.get("cpu_baseline"), d["clocks"])
PY
