"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/.
Usage: python profiles/ncu_summarise.py gpurun_out/r01h_solve_kernel.ncu-rep > profiles/r01h_solve_kernel.ncu.txt"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
print("# %s" % os.path.basename(rep))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %s %s" % (w, vals[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, h, ci, lines, sass, cur = None, None, None, [], [], None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if len(r) > 5 and r[0] == "Line No":
        h, ci = r, {}
        for i, n in enumerate(r):
            ci.setdefault(n, i)
        continue
    if h is None or len(r) < len(h):
        continue

    def g(n):
        try:
            return float(r[ci[n]])
        except Exception:
            return 0.0
    if r[0].strip().isdigit():
        st = {k[6:]: g(k) for k in h if k.startswith("stall_") and "Not Issued" not in k}
        cur = (cur_file, int(r[0]), r[1].strip()[:110])
        lines.append((g("# Samples"), cur, g("Instructions Executed"), st))
tot = sum(l[0] for l in lines) or 1.0
agg = {}
for s_, _, _, st in lines:
    for k, v in st.items():
        agg[k] = agg.get(k, 0.0) + v
print("\nwarp-stall samples by reason (%d samples): " % tot + ", ".join("%s %.1f%%" % (k, 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
print("\ntop source lines by stall samples")
for s_, (f, l, t), ie, st in sorted(lines, key=lambda t: -t[0])[:top_n]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%5.1f%% %-18s L%-4d inst=%-10d %-28s | %s" % (100 * s_ / tot, f, l, ie, " ".join("%s=%d" % (k, v) for k, v in top if v > 0), t))
