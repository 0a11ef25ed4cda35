#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line.
usage: ncu_lines.py REPORT KERNEL_REGEX [launch_index] [top]"""
import csv, io, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + rx, "--launch-skip", str(which), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
f = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter()); src = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": f = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or f is None: continue
    d = dict(zip(hdr, r))
    # columns: Line No, Source (cuda), Address, Source(sass) ... duplicated 'Source' key -> sass wins; keep cuda by index
    line = r[0]
    if line.isdigit():
        cur = (f, int(line)); src.setdefault(cur, r[1])
    n = d.get("# Samples", "")
    if n.isdigit() and int(n) > 0:
        a = agg[cur]
        a["samples"] += int(n)
        for k in ("stall_long_sb", "stall_barrier", "stall_lg", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_membar", "stall_sleep", "stall_no_inst", "stall_math", "stall_mio", "stall_not_selected", "stall_selected"):
            v = d.get(k, "")
            if v.isdigit(): a[k] += int(v)
tot = sum(a["samples"] for a in agg.values()) or 1
print(f"total samples {tot}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    why = ", ".join(f"{k[6:]}={v}" for k, v in a.most_common(4) if k != "samples")
    print(f"{100 * a['samples'] / tot:5.1f}%  {key[0]}:{key[1]:<4} {src.get(key, '')[:90].strip()}   [{why}]")
