#!/usr/bin/env python
"""Top source lines by warp-stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python profiles/ncu_top_lines.py gpurun_out/prof.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, cur_fn, hdr = None, None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1][:40]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            s = int(d.get("# Samples", "0") or 0)
            ins = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        key = (cur_fn, cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, d.get("Source", "")[:110]])
        a[0] += s
        a[1] += ins
tot = {}
for (fn, f, l), (s, ins, src) in agg.items():
    tot[fn] = tot.get(fn, 0) + s
for fn in tot:
    print(f"== {fn}: {tot[fn]} samples")
    items = sorted(((s, ins, f, l, src) for (fn2, f, l), (s, ins, src) in agg.items() if fn2 == fn), reverse=True)[:topn]
    for s, ins, f, l, src in items:
        print(f"{100.0 * s / max(tot[fn], 1):5.1f}%  {ins:>10d} inst  {f}:{l}  {src.strip()}")
