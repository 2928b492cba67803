"""Per-kernel text summaries of a multi-kernel .ncu-rep (ncu --set full --import-source on): launch shape, occupancy limiters, DRAM bytes,
pipe utilisation, warp-stall reasons and the top source lines by stall samples.
Usage: python profiles/ncu_report.py REPORT.ncu-rep OUT_PREFIX [top_n]      -> OUT_PREFIX_<kernel>[_<k>].ncu.txt + OUT_PREFIX_traffic.json"""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, prefix = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {n: i for i, n in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
seen, traffic = {}, {}


def to_bytes(v, u):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


for r in rows[2:]:
    kid, name = r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0]
    k = seen.get(name, 0)
    seen[name] = k + 1
    out = "%s_%s%s.ncu.txt" % (prefix, name, "" if k == 0 else "_%d" % k)
    lines_out = ["# %s  kernel id %s  %s (launch %d of this kernel in the report)" % (os.path.basename(rep), kid, name, k)]
    for w in want:
        if w in ix:
            lines_out.append("%-72s %s %s" % (w, r[ix[w]], units[ix[w]]))
    if "dram__bytes_read.sum" in ix:
        tb = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]]) + to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        traffic.setdefault(name, []).append({"id": int(kid), "grid": r[ix["launch__grid_size"]], "dram_bytes": tb, "ms": float(r[ix["gpu__time_duration.sum"]]) * (1e-3 if units[ix["gpu__time_duration.sum"]] == "us" else 1.0)})
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::%d" % (int(kid) + 1)], capture_output=True, text=True).stdout
    cur_file, h, ci, lines = None, None, None, []
    for rr in csv.reader(io.StringIO(src)):
        if len(rr) >= 2 and rr[0] == "File Path":
            cur_file = os.path.basename(rr[1]); continue
        if len(rr) > 5 and rr[0] == "Line No":
            h, ci = rr, {}
            for i, n in enumerate(rr):
                ci.setdefault(n, i)
            continue
        if h is None or len(rr) < len(h):
            continue

        def g(n):
            try:
                return float(rr[ci[n]])
            except Exception:
                return 0.0
        if rr[0].strip().isdigit():
            st = {kk[6:]: g(kk) for kk in h if kk.startswith("stall_") and "Not Issued" not in kk}
            lines.append((g("# Samples"), (cur_file, int(rr[0]), rr[1].strip()[:120]), g("Instructions Executed"), st))
    tot = sum(l[0] for l in lines) or 1.0
    agg = {}
    for s_, _, _, st in lines:
        for kk, v in st.items():
            agg[kk] = agg.get(kk, 0.0) + v
    lines_out.append("\nwarp-stall samples by reason (%d samples): " % tot + ", ".join("%s %.1f%%" % (kk, 100 * v / tot) for kk, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    lines_out.append("\ntop source lines by stall samples")
    for s_, (f, l, t), ie, st in sorted(lines, key=lambda t: -t[0])[:top_n]:
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        lines_out.append("%5.1f%% %-18s L%-4d inst=%-10d %-28s | %s" % (100 * s_ / tot, f, l, ie, " ".join("%s=%d" % (kk, v) for kk, v in top if v > 0), t))
    open(out, "w").write("\n".join(lines_out) + "\n")
    print(out)
json.dump(traffic, open(prefix + "_traffic.json", "w"), indent=1)
