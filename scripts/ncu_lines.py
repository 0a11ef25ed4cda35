#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line (the rows of the source page that carry a line number).
usage: ncu_lines.py REPORT KERNEL_REGEX [launch_index] [top]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx, "--launch-skip", str(which), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
f, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        f = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or not r[0].isdigit():
        continue
    d = {}
    for k, v in zip(hdr, r):
        d.setdefault(k, v)
    ns = d.get("# Samples", "0")
    n = int(ns) if ns.isdigit() else 0
    if n:
        stalls = sorted(((int(v), k) for k, v in zip(hdr[6:], r[6:]) if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0), reverse=True)[:4]
        out.append((n, f, int(r[0]), r[1].strip()[:100], ", ".join(f"{k[6:]}={v}" for v, k in stalls), d.get("Instructions Executed", "")))
tot = sum(o[0] for o in out) or 1
print(f"kernel {rx} launch {which}: total samples {tot}")
for n, f, line, src, why, inst in sorted(out, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}%  {f}:{line:<4} inst={inst:<9} {src}   [{why}]")
