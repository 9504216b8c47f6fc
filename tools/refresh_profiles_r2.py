"""Turn gpurun_out/r2_bench_n1.json, r2_launches.csv and r2_full.ncu-rep (tools/make_profiles_r2.sh) into the committed summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
line = [l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][0]
json.dump(json.loads(line), open("profiles/r2_bench_n1.json", "w"), indent=1)
subprocess.run([sys.executable, "tools/ncu_summary.py", "gpurun_out/r2_full.ncu-rep", "profiles/r2_ncu_full_world_step.txt",
                "ncu --set full --clock-control none --cache-control none --graph-profiling node -k regex:k_gjk2|k_epa|k_integrate|k_resolve|k_link -s 18 -c 9 python bench.py --steps 3 --warmup 3 : "
                "the nine kernels of ONE step of the headline world (1 000 000 mixed bodies, exact search), graph nodes profiled in place, L2 left as the preceding kernels left it "
                "(dram__bytes = what the step really moves)"], stdout=subprocess.DEVNULL)
rows = [r for r in csv.reader(open("gpurun_out/r2_launches.csv")) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
L = [(r[ik].split("(")[0].replace("void ", ""), float(r[iv].replace(",", ""))) for r in rows[1:]]
idx = [i for i, (n, v) in enumerate(L) if n.startswith("k_gjk2<1>")]
s = idx[4] - 1; e = idx[5] - 1
with open("profiles/r2_launches_bench.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 400 : python bench.py --steps 3 --warmup 3\n")
    f.write("# (per-launch times are serialised and cold for the first kernels; the SHARES are what to compare with bench.py's phase_ms)\n")
    tot = sum(v for n, v in L[s:e])
    f.write("## one step of the headline world (1 000 000 mixed bodies, exact search): %d launches, %.1f us\n" % (e - s, tot / 1e3))
    for n, v in L[s:e]: f.write("%-44s %9.1f us  %5.1f%%\n" % (n[:44], v / 1e3, 100 * v / tot))
    agg = collections.OrderedDict()
    for n, v in L: agg.setdefault(n, [0, 0.0]); agg[n][0] += 1; agg[n][1] += v
    f.write("## the first %d launches of the command, by kernel\n" % len(L))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): f.write("%-44s n=%4d total %10.1f us\n" % (k[:44], v[0], v[1] / 1e3))
out = subprocess.run(["ncu", "-i", "gpurun_out/r2_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out))); hdr = rows[0]; units = rows[1]
ik = hdr.index("Kernel Name"); ir = hdr.index("dram__bytes_read.sum"); iw = hdr.index("dram__bytes_write.sum")
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = 0; per = {}
for r in rows[2:]:
    n = r[ik].split("(")[0].replace("void ", "")
    b = float(r[ir].replace(",", "")) * mult[units[ir]] + float(r[iw].replace(",", "")) * mult[units[iw]]; per[n] = b
    if n.startswith("k_gjk") or n.startswith("k_epa"): tot += b
json.dump({"bodies": 1000000, "narrow_dram_bytes_per_step": tot, "per_kernel_bytes": per,
           "source": "profiles/r2_ncu_full_world_step.txt (ncu --set full --cache-control none --graph-profiling node on `python bench.py --steps 3 --warmup 3`, dram__bytes_read.sum + "
                     "dram__bytes_write.sum of k_gjk2 + k_epa_first + k_epa2 + the k_epa tiers of one step)"}, open("profiles/r2_traffic.json", "w"), indent=1)
print(open("profiles/r2_launches_bench.txt").read()[:1100]); print("narrow DRAM traffic %.1f MB/step" % (tot / 1e6))
d = json.load(open("profiles/r2_bench_n1.json"))
print("value %.4g ms %.4f e2e %.4g frac %.3f traffic %s pairs %.4g box %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["pair_tests"]["value"], d["pair_tests"]["box_only"]["value"]))
for k, v in d["other_configs"].items(): print(k, v.get("value"), v.get("ms_per_step"))
