#!/usr/bin/env python
"""Stall samples of one profiled launch aggregated per CUDA source line (needs -lineinfo): tools/ncu_lines.py <rep> [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, agg = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ci = hdr.index("Warp Stall Sampling (All Samples)")
        names = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
        idx = {n: hdr.index(n) for n in names}
    elif hdr and r and r[0] not in ("", "Line No") and len(r) > ci:
        try:
            v = float(r[ci])
        except ValueError:
            continue
        if v > 0:
            topstall = max(idx, key=lambda n: float(r[idx[n]] or 0))
            agg.append((v, cur_file, r[0], r[1].strip()[:110], topstall.replace("stall_", "")))
tot = sum(a[0] for a in agg)
print("| share | file:line | dominant stall | source |\n|---:|---|---|---|")
for v, f, ln, src, st in sorted(agg, reverse=True)[:top]:
    print("| %.2f%% | %s:%s | %s | `%s` |" % (100 * v / tot, f, ln, st, src))
