#!/usr/bin/env python
"""Summarise .ncu-rep captures (read here, without a GPU) into the markdown table committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    ("time (us)", "gpu__time_duration.sum", 1.0),
    ("grid", "launch__grid_size", 1.0),
    ("block", "launch__block_size", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active", 1.0),
    ("issue active %", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1.0),
    ("FP64 pipe %", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 1.0),
    ("dram read (MB)", "dram__bytes_read.sum", None),
    ("dram write (MB)", "dram__bytes_write.sum", None),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 %", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 1.0),
    ("warp insts", "smsp__inst_executed.sum", 1.0),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0,
        "nsecond": 1e-3, "second": 1e6}


def load(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        recs.append({h: (u, v) for h, u, v in zip(head, units, r)})
    return recs


def traffic(paths, walks, out):
    """--traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of each captured kernel -> profiles/ncu_traffic.json
    (what bench.py reports as roofline.traffic; `walks` = the --walks of the captured command, bench.py refuses another size)."""
    import json
    import os
    res = {}
    for path in paths:
        for rec in load(path):
            name = rec.get("Kernel Name", ("", "?"))[1].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                u, v = rec[key]
                tot += float(v.replace(",", "")) * UNIT[u] * 1e6
            res[name] = {"dram_bytes_per_launch": tot, "walks": walks,
                         "source": f"ncu --set full capture {os.path.basename(path)} (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


def main():
    if sys.argv[1] == "--traffic":   # --traffic <walks> <out.json> captures...
        return traffic(sys.argv[4:], int(sys.argv[2]), sys.argv[3])
    lines = []
    for path in sys.argv[1:]:
        for rec in load(path):
            name = rec.get("Kernel Name", ("", "?"))[1].split("(")[0]
            vals = []
            for label, key, _ in KEYS:
                u, v = rec.get(key, ("", ""))
                try:
                    x = float(v.replace(",", ""))
                except ValueError:
                    vals.append("-"); continue
                if u in UNIT and ("byte" in u or "second" in u or u in ("ns", "us", "ms")):
                    x *= UNIT[u]
                vals.append(f"{x:,.1f}" if abs(x) < 1e6 else f"{x:,.0f}")
            stalls = sorted(((float(v[1]), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                             for k, v in rec.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and v[1]),
                            reverse=True)[:4]
            lines.append((name, vals, ", ".join(f"{k} {x:.2f}" for x, k in stalls)))
    print("| kernel | " + " | ".join(k[0] for k in KEYS) + " | top stalls (warps per issue) |")
    print("|---|" + "---|" * (len(KEYS) + 1))
    for name, vals, st in lines:
        print(f"| {name} | " + " | ".join(vals) + f" | {st} |")


if __name__ == "__main__":
    main()
