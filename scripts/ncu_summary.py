"""Summarise an .ncu-rep (read on a box without a GPU): headline metrics, stall mix, dynamic opcode mix."""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
cell_steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k][0]} {m[k][1]}")
print("-- stalls per issue")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        v = float(m[h][0])
        if v > 0.02:
            print(f"   {h.split('stalled_')[1].split('_per_issue')[0]:22s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
iS, iE = h2.index("Source"), h2.index("Instructions Executed")
byop, tot = collections.Counter(), 0
for r in rows[2:]:
    mm = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    if mm:
        byop[mm.group(1)] += int(r[iE]); tot += int(r[iE])
print(f"-- dynamic warp instructions: {tot}" + (f" = {tot / (cell_steps / 32):.1f} per warp-step" if cell_steps else ""))
for op, c in byop.most_common(18):
    print(f"   {op:10s} {100.0 * c / tot:5.1f} %" + (f"  {c / (cell_steps / 32):7.1f} per warp-step" if cell_steps else ""))
print("   static SASS instructions:", len(rows) - 2)
