"""Turns the ncu outputs in gpurun_out/ into the tracked summaries under profiles/:
launch-list tables (share of the step per kernel) and the key raw metrics of each full capture."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def launches(name):
    p = os.path.join(G, f"launches_{name}.csv")
    if not os.path.exists(p):
        return
    rows = [r for r in csv.reader(open(p)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = os.path.join(ROOT, "profiles", f"{TAG}_launches_{name}.md")
    with open(out, "w") as f:
        f.write(f"# ncu launch list — bench.py --workload {name} (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1] / a[0] / 1e3:.2f} | {a[1] / tot * 100:.1f}% |\n")
    print("wrote", out)


def full(rep, label):
    p = os.path.join(G, rep)
    if not os.path.exists(p):
        return None
    raw = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = rows[0]
    d = {"kernel": rows[2][H.index("Kernel Name")], "report": rep}
    for k in KEYS:
        if k in H:
            d[k] = f"{rows[2][H.index(k)]} {rows[1][H.index(k)]}".strip()
    out = os.path.join(ROOT, "profiles", f"{TAG}_{label}.json")
    json.dump(d, open(out, "w"), indent=1)
    print("wrote", out)
    return d


if __name__ == "__main__":
    for n in ("h1", "v1"):
        launches(n)
    traffic = {}
    for rep, label, wl in (("prof_gemm_h1.ncu-rep", "ncu_emb_gemm_h1", "h1"), ("prof_bm25_h1.ncu-rep", "ncu_bm25_warp_h1", None),
                           ("prof_merge_h1.ncu-rep", "ncu_emb_merge_h1", None), ("prof_scan_v1.ncu-rep", "ncu_emb_scan_v1", "v1")):
        d = full(rep, label)
        if d and wl:
            def gb(s):
                v, u = s.split()[:2]
                return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            traffic[wl] = gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        print("wrote profiles/traffic.json", traffic)
