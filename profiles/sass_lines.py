#!/usr/bin/env python
"""Attribute an ncu `--page source --csv` dump (SASS rows: executed instructions, stall samples) to the CUDA source
lines of this repo, using the line table of the SAME build (`nvdisasm -g -c` of the cubin inside the .so).

    ncu -i gpurun_out/prof_bm25_h1.ncu-rep --page source --csv > /tmp/src.csv
    python profiles/sass_lines.py /tmp/src.csv bm25_tile2_kernelILb0ELb0E [--top 40]

The kernel is matched by a substring of its MANGLED name.  Rows are joined by instruction offset (ncu address minus
the first address of the dump)."""
from __future__ import annotations

import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "oramacore_b200", "liboramacore_b200.so")


def line_table(kernel_substr: str):
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(SO)], cwd=d, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], cwd=d, check=True, capture_output=True, text=True).stdout
    table, cur, on = {}, None, False
    for ln in txt.splitlines():
        if ln.startswith("//-") and ".text." in ln:
            on = kernel_substr in ln
            cur = None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m and cur:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src_csv, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    table = line_table(kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed")}
    stall_cols = [(n, i) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    base = None
    per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
    tot_inst = tot_samp = 0
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        key = table.get(a - base, (("?", 0), ""))[0]
        inst, samp = int(r[ci["Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
        e = per_line[key]
        e[0] += inst
        e[1] += samp
        for n, i in stall_cols:
            if r[i] and r[i] != "0":
                e[2][n] += int(r[i])
        tot_inst += inst
        tot_samp += samp
    src_cache = {}

    def text(f, n):
        p = os.path.join(HERE, "..", "oramacore_b200", "csrc", f)
        if f not in src_cache:
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        s = src_cache[f]
        return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""

    print(f"kernel {kernel}: {tot_inst} warp instructions, {tot_samp} stall samples")
    print("| file:line | inst % | samples % | top stalls | source |")
    print("|---|---|---|---|---|")
    for (f, n), (inst, samp, st) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        tops = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"| {f}:{n} | {100.0 * inst / max(tot_inst, 1):.1f} | {100.0 * samp / max(tot_samp, 1):.1f} | {tops} | `{text(f, n)}` |")


if __name__ == "__main__":
    main()
