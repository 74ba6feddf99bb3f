"""Per-kernel table of one frame from an `ncu --metrics ... --csv` log (tools/profile_r03.sh step 2).  usage: passes_table.py log.csv"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
I = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Value")}
k = collections.OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    d = k.setdefault(r[I["ID"]], {"name": r[I["Kernel Name"]]})
    d[r[I["Metric Name"]]] = float(r[I["Metric Value"]].replace(",", ""))
print("%-24s %9s %9s %6s %7s %8s %12s %6s %10s %10s" % ("kernel", "time us", "inst M", "lanes", "fp64 %", "issue %", "act/elapsed", "grid", "dram rd MB", "dram wr MB"))
tot = 0.0
for d in k.values():
    t = d["gpu__time_duration.sum"] / 1e3
    tot += t
    print("%-24s %9.1f %9.1f %6.2f %7.1f %8.1f %12.2f %6d %10.2f %10.2f" % (
        d["name"], t, d["smsp__inst_executed.sum"] / 1e6, d["smsp__thread_inst_executed_per_inst_executed.ratio"],
        d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"], d["smsp__issue_active.avg.pct_of_peak_sustained_active"],
        d["sm__cycles_active.avg"] / d["sm__cycles_elapsed.max"], int(d["launch__grid_size"]),
        d["dram__bytes_read.sum"] / 1e6, d["dram__bytes_write.sum"] / 1e6))
print("sum of kernel times %.1f us" % tot)
