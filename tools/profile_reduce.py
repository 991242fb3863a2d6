"""Reduce the ncu CSV exports of tools/profile_r2.sh to the small summaries kept under profiles/.

  gpurun_out/r2_launches.csv          -> r2_launches_summary.md      (share of step time per kernel)
  gpurun_out/r2_conv_roofline_raw.csv -> r2_ncu_full_conv_block_tc.json (the metrics bench.py's roofline.traffic reads)
  gpurun_out/r2_instep_raw.csv        -> r2_ncu_instep_summary.md / .json (per-launch duration, DRAM bytes, tensor pipe, occupancy)
"""
import collections
import csv
import json
import os
import re
import sys

O = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def read_raw(path):
    """ncu --page raw --csv: header row, units row, one row per launch."""
    if not os.path.exists(path):
        return [], [], []
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def launches_summary():
    path = os.path.join(O, "r2_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    if not rows:
        return
    h = rows[0]
    ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0) if r[iu] in ("ns", "us", "ms", "s") else float(r[iv].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ik])[:100]
        tot[k][0] += v
        tot[k][1] += 1
    s = sum(v[0] for v in tot.values())
    out = ["| share | total us | launches | avg us | kernel |", "|---|---|---|---|---|"]
    for k, (v, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
        out.append(f"| {100 * v / s:.1f}% | {v:.0f} | {n} | {v / n:.1f} | `{k}` |")
    open(os.path.join(O, "r2_launches_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


def pick(h, u, row):
    d = {}
    for name in KEEP:
        if name in h:
            i = h.index(name)
            d[name] = [row[i], u[i]]
    return d


def roofline_json():
    h, u, rows = read_raw(os.path.join(O, "r2_conv_roofline_raw.csv"))
    if not rows:
        return
    d = pick(h, u, rows[-1])
    d["_kernel"] = rows[-1][h.index("Kernel Name")][:80] if "Kernel Name" in h else ""
    d["_how"] = "ncu --set full --clock-control none, tools/ncu_conv_tc.py 128 (B=256, 128->128, k=5, T=128, IN+ReLU, saves c; 10 rotating inputs > L2, pre-rounded), last captured launch"
    json.dump(d, open(os.path.join(O, "r2_ncu_full_conv_block_tc.json"), "w"), indent=1)
    print("roofline capture:", {k: v for k, v in d.items() if not k.startswith("_")})


def num(v, unit):
    try:
        return float(v.replace(",", "")) * SCALE.get(unit, 1.0)
    except ValueError:
        return float("nan")


def instep():
    h, u, rows = read_raw(os.path.join(O, "r2_instep_raw.csv"))
    if not rows:
        return
    ik = h.index("Kernel Name")
    recs = []
    for r in rows:
        d = pick(h, u, r)
        recs.append(dict(kernel=re.sub(r"\(.*", "", r[ik])[:60], **{k: num(*v) for k, v in d.items()}))
    json.dump(recs, open(os.path.join(O, "r2_ncu_instep.json"), "w"))
    groups = collections.defaultdict(list)
    for i, r in enumerate(recs):
        groups[(r["kernel"], int(r.get("launch__grid_size", 0)))].append(r)
    out = ["| kernel | grid | launches | avg us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | warps active % | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
    for (k, g), rs in sorted(groups.items(), key=lambda kv: -sum(r["gpu__time_duration.sum"] for r in kv[1])):
        n = len(rs)
        avg = lambda key: sum(r.get(key, float("nan")) for r in rs) / n
        out.append(f"| `{k}` | {g} | {n} | {avg('gpu__time_duration.sum'):.1f} | {avg('dram__bytes_read.sum') / 1e6:.1f} | {avg('dram__bytes_write.sum') / 1e6:.1f} | "
                   f"{avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {avg('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                   f"{avg('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {avg('launch__registers_per_thread'):.0f} |")
    open(os.path.join(O, "r2_ncu_instep_summary.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    launches_summary()
    roofline_json()
    instep()
