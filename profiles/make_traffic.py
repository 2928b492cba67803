"""profiles/ncu_traffic.json (what bench.py scales to its batch for roofline.traffic / kernels[*].dram_traffic_bytes) from the per-kernel DRAM byte list that
profiles/ncu_report.py writes next to its summaries.  Units = the windows (camera streams for lk_*) of the captured launch: profiles/ncu_target.py runs 592 of each.
Usage: python profiles/make_traffic.py PREFIX_traffic.json UNITS "source note" > profiles/ncu_traffic.json"""
import json
import sys

src, units, note = json.load(open(sys.argv[1])), float(sys.argv[2]), sys.argv[3]
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum of the FIRST launch of each kernel in one ncu --set full capture (%s), divided by the %d windows / camera "
                "streams that launch processed; bench.py scales it to its batch" % (note, int(units))}
for name, launches in sorted(src.items()):
    base = name.replace("_kernel", "").replace("lk_track_tasks", "lk_track").replace("pyr_down_tasks", "lk_pyr_down")
    first = launches[0]
    out[base] = {"bytes_per_unit": first["dram_bytes"] / units, "unit": "camera stream (one of the two launches of a tick)" if base.startswith("lk_") else "window",
                 "ms_under_ncu": first["ms"], "launches_in_capture": len(launches)}
json.dump(out, sys.stdout, indent=1)
print()
