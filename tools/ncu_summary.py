"""Summarise an .ncu-rep (read here, without a GPU): headline metrics + hot SASS."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("== kernel:", name[:100])
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print("  %-62s %-10s %s" % (h, u, v))
    stalls = [(h, float(v)) for h, v in zip(hdr, r) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    stalls.sort(key=lambda t: -t[1])
    print("  stalls/issue:", ", ".join("%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for h, v in stalls[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
# may hold several kernels: split on 'Kernel Name'
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        hdr = rows[i + 1]
        ix = {h: k for k, h in enumerate(hdr)}
        j = i + 2
        data = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            data.append(rows[j])
            j += 1
        tot = sum(int(r[ix["Instructions Executed"]]) for r in data)
        samp = sum(int(r[ix["# Samples"]]) for r in data)
        print("== SASS: %d warp-instr, %d samples; top %d by samples" % (tot, samp, top))
        order = sorted(range(len(data)), key=lambda k: -int(data[k][ix["# Samples"]]))[:top]
        for k in sorted(order):
            r = data[k]
            print("  %5d %-58s exec %9s thr %4s samp %6s lsb %5s wait %5s br %5s" % (
                k, r[ix["Source"]].strip()[:58], r[ix["Instructions Executed"]], r[ix["Avg. Threads Executed"]],
                r[ix["# Samples"]], r[ix["stall_long_sb"]], r[ix["stall_wait"]], r[ix["stall_branch_resolving"]]))
        i = j
    else:
        i += 1
