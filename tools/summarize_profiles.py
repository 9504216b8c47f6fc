"""Turn gpurun_out/prof_launches.csv + prof_full.ncu-rep into the committed text summaries under profiles/."""
import csv, io, subprocess, sys, os, collections
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rows = [r for r in csv.reader(open("gpurun_out/prof_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
L = [(int(r[iid]), r[ik].split("(")[0].replace("void ", ""), float(r[iv].replace(",", ""))) for r in rows[1:]]
# one timed world step = launches between two consecutive k_integrate after the warm-up
starts = [i for i, l in enumerate(L) if l[1].startswith("k_integrate")]
with open("profiles/%s_launches_bench.txt" % tag, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : python bench.py --steps 3 --warmup 3  (per-launch times are cold-cache and serialised)\n")
    if len(starts) > 5:
        s, e = starts[4], starts[5]
        tot = sum(l[2] for l in L[s:e])
        f.write("## one world step (C2, 10k bodies): %d launches, %.1f us\n" % (e - s, tot / 1e3))
        for l in L[s:e]:
            f.write("%-40s %9.1f us  %5.1f%%\n" % (l[1], l[2] / 1e3, 100 * l[2] / tot))
    agg = collections.OrderedDict()
    for l in L:
        agg.setdefault(l[1], [0, 0.0]); agg[l[1]][0] += 1; agg[l[1]][1] += l[2]
    f.write("## all %d launches of the command, by kernel\n" % len(L))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-40s n=%4d total %10.1f us\n" % (k, v[0], v[1] / 1e3))
out = subprocess.run(["ncu", "-i", "gpurun_out/prof_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
with open("profiles/%s_ncu_full_world_step.txt" % tag, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on : kernels of one C2 world step inside python bench.py --steps 3 --warmup 3\n")
    for r in rows[2:]:
        f.write("\n")
        for kname in keep:
            if kname in hdr:
                i = hdr.index(kname); f.write("%-92s %s %s\n" % (kname, r[i][:110], units[i]))
print(open("profiles/%s_launches_bench.txt" % tag).read()[:3000])
