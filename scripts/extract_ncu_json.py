"""profiles/<tag>_ncu_kernels_dram.json from the raw pages of scripts/ncu_capture.sh: the per-kernel metrics bench.py (roofline.traffic)
and DESIGN.md quote.  usage: python scripts/extract_ncu_json.py <tag-in-gpurun_out> <out-tag> <cells>"""
import csv, json, re, subprocess, sys
src, out, cells = sys.argv[1], sys.argv[2], int(sys.argv[3])
keys = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
res = {"cells": cells, "commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip(),
       "source": "gpurun_out/%s_{rev,fwd}_raw.csv (ncu --set full --clock-control none, one launch per kernel, 1440x720 bench mesh)" % src}
for part in ("rev", "fwd"):
    rows = list(csv.reader(open("gpurun_out/%s_%s_raw.csv" % (src, part))))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        m = re.search(r"(Rev[ABC]|Fwd[ABC])", r[ki])
        if not m:
            continue
        d = {}
        for k in keys:
            if k in hdr:
                v = float(r[hdr.index(k)].replace(",", ""))
                u = units[hdr.index(k)]
                if k.startswith("dram__bytes"):
                    v *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
                if k == "gpu__time_duration.sum":
                    v *= {"us": 1e3, "ms": 1e6, "ns": 1.0, "usecond": 1e3, "msecond": 1e6, "nsecond": 1.0}.get(u, 1.0)  # -> ns
                d[k] = v
        res[m.group(1)] = d
json.dump(res, open("profiles/%s_ncu_kernels_dram.json" % out, "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, dict) else {a: round(b, 2) for a, b in list(v.items())[:3]}) for k, v in res.items()}, indent=0)[:900])
