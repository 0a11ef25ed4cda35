#!/usr/bin/env python
"""Turn one GPU visit's ncu output (gpurun_out/) into the tracked summaries under profiles/.
usage: summarise_profiles.py TAG LAUNCHES_CSV RAW_CSV "COMMAND" [BENCH_JSON]"""
import csv, json, sys, collections

tag, launches, raw, cmd = sys.argv[1:5]
bench = sys.argv[5] if len(sys.argv) > 5 else None


def rows_of(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


# ---- launch list ----
agg = collections.OrderedDict()
for r in rows_of(launches):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = r["Kernel Name"]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
ours = {k: v for k, v in agg.items() if "<unnamed>" in k or "lg::" in k or "seed_local" in k or "relabel" in k}
step = {k: v for k, v in ours.items() if any(s in k for s in ("gather_", "sample_hop", "rank_kernel", "relabel_kernel",
                                                                  "batch_generate", "release_kernel", "seed_local"))}
tot = sum(v[1] for v in step.values()) or 1.0
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write(f"# {tag} — launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n`{cmd}`; "
            "cold-cache, serialised: compare shares, not absolutes.\n\n| kernel | launches | total us | avg us | share of step kernels |\n|---|---|---|---|---|\n")
    for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k[:110]}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
    if bench:
        j = json.loads([l for l in open(bench) if l.startswith("{")][-1])
        b = j["breakdown_ms"]
        g = sum(v for k, v in b.items() if k.startswith("gather"))
        f.write(f"\nbench.py (CUDA events, no profiler), ms per op: " + ", ".join(f"{k} {v:.3f}" for k, v in b.items()) +
                f" -> gather share {g:.3f}/{sum(b.values()):.3f} = {100 * g / sum(b.values()):.0f}% "
                f"(value {j['value'] / 1e6:.2f} M seeds/s, {j['ms_per_step']:.4f} ms/step, roofline.frac {j['roofline']['frac']:.3f}).\n")
# ---- full capture ----
rr = rows_of(raw)
cols = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]
units = rr[0]
with open(f"profiles/{tag}_ncu_full_summary.md", "w") as f:
    f.write(f"# {tag} — ncu --set full --clock-control none --import-source on\n\n`{cmd}` (B200)\n\n")
    f.write("| " + " | ".join(f"{c} [{units.get(c, '')}]" if units.get(c) else c for c in cols) + " |\n|" + "---|" * len(cols) + "\n")
    gather = None
    for r in rr[1:]:
        f.write("| " + " | ".join(r.get(c, "")[:90] for c in cols) + " |\n")
        if "gather_" in r["Kernel Name"]:
            gather = r
    if gather:
        t = (float(gather["dram__bytes_read.sum"]) + float(gather["dram__bytes_write.sum"])) * 1e6
        f.write(f"\nGather (one fused launch per step): DRAM traffic {t / 1e6:.1f} MB per launch "
                f"(read {gather['dram__bytes_read.sum']} MB + write {gather['dram__bytes_write.sum']} MB).\n")
        # keyed by workload / D / Kg; per-row bytes let bench.py scale the figure to the rows of its own launch
        try:
            table = json.load(open("profiles/roofline_traffic.json"))
        except Exception:
            table = {}
        if bench:
            j = json.loads([l for l in open(bench) if l.startswith("{")][-1])
            rows = j["roofline"]["rows_per_step"]
            key = j["roofline"]["traffic_key"]
            table[key] = {"bytes_per_launch": t, "rows_per_launch": rows, "bytes_per_row": t / rows,
                          "source": f"profiles/{tag}_ncu_full_summary.md: dram__bytes_read.sum+dram__bytes_write.sum of the fused gather launch "
                                    f"(ncu --set full, `{cmd}`): {t / rows:.1f} B per gathered row"}
            json.dump(table, open("profiles/roofline_traffic.json", "w"), indent=1)
print("ok")
