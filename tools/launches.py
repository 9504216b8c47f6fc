"""Print the tail of an ncu --csv launch list (gpu__time_duration etc.) as a table."""
import csv, sys
f = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 14
rows = [r for r in csv.reader(open(f)) if len(r) > 10]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = {}; order = []
for r in rows[1:]:
    key = (int(r[iid]), r[ik][:46])
    if key not in d: d[key] = {}; order.append(key)
    d[key][r[im]] = r[iv]
tot = 0.0
for key in order[-last:]:
    m = d[key]; t = float(m.get("gpu__time_duration.sum", "0").replace(",", "")); tot += t
    print("%3d %-46s t=%9.1f us inst=%11s act/inst=%5s warps%%=%6s issue%%=%6s" % (key[0], key[1], t / 1e3, m.get("smsp__inst_executed.sum"), m.get("smsp__thread_inst_executed_per_inst_executed.ratio"), m.get("sm__warps_active.avg.pct_of_peak_sustained_active"), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active")))
print("sum of listed: %.1f us" % (tot / 1e3))
