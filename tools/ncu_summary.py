#!/usr/bin/env python
"""Summarises one kernel of an .ncu-rep (ncu --set full --import-source on): headline metrics, stall
reasons, and the instruction / stall-sample share per source line.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines] > profiles/rNN_ncu_summary_<kernel>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
dcol = hdr.index("gpu__time_duration.sum") if "gpu__time_duration.sum" in hdr else None
cands = [r for r in rows[2:] if len(r) == len(hdr)]
vals = max(cands, key=lambda r: float(r[dcol].replace(",", "")) if dcol is not None and r[dcol] else 0.0)  # longest launch
print("launches in report:", len(cands))
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
print("kernel", d.get("Kernel Name", ("", "?"))[1])
for k in want:
    if k in d:
        print(k, d[k][0], d[k][1])
print("-- warp stall reasons (smsp__average_warps_issue_stalled_*_per_issue_active / pcsamp) --")
st = [(h, float(v.replace(",", ""))) for h, (u, v) in d.items()
      if h.startswith("smsp__average_warp_latency_issue_stalled") and h.endswith(".ratio") and v not in ("", "n/a")]
for h, v in sorted(st, key=lambda x: -x[1])[:10]:
    print("  %-90s %.3f" % (h, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None
inst, samp, text = collections.Counter(), collections.Counter(), {}
col = {}
for x in csv.reader(io.StringIO(src)):
    if x and x[0] == "Line No":
        col = {n: i for i, n in enumerate(x)}
        continue
    if not col or len(x) < 10:
        continue
    if x[0] != "":
        cur = int(x[0])
        text[cur] = x[1]
        continue
    try:
        inst[cur] += int(x[col["Instructions Executed"]])
        samp[cur] += int(x[col["# Samples"]])
    except (ValueError, KeyError):
        pass
ti, ts = sum(inst.values()), max(1, sum(samp.values()))
print("-- source lines: %d warp instructions, %d stall samples --" % (ti, ts))
for l, n in inst.most_common(top):
    print("%5d %5.1f%% inst %5.1f%% samples  %s" % (l, 100.0 * n / max(1, ti), 100.0 * samp[l] / ts, text[l].strip()[:110]))
print("-- lines by stall samples --")
for l, n in samp.most_common(12):
    print("%5d %5.1f%% samples %5.1f%% inst  %s" % (l, 100.0 * n / ts, 100.0 * inst[l] / max(1, ti), text[l].strip()[:110]))
