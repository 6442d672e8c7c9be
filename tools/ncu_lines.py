#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu capture (--import-source on, -lineinfo) to CUDA source lines, offline.
usage: python tools/ncu_lines.py <capture.ncu-rep> <object.o> <kernel-name-substring> [top]
Joins `ncu --page source --csv` (SASS rows in address order) with `nvdisasm -g` line markers of the same kernel."""
import csv
import re
import subprocess
import sys
import tempfile
import os


def main():
    rep, obj, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    h = rows[hi]
    ia, isrc, iinst, ismp = h.index("Address"), h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    body = []
    for r in rows[hi + 1:]:          # a capture with several launches holds one table per launch: use the first
        if "Address" in r and "Source" in r:
            break
        body.append(r)
    sass = [(int(r[ia], 16), r[isrc].strip(), int(r[iinst] or 0), int(r[ismp] or 0)) for r in body if len(r) > ismp]
    base = sass[0][0]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    line_of = {}
    cur = None
    inside = False
    for ln in dis.splitlines():
        if ln.startswith("\t.section") or ln.startswith(".section"):
            inside = kname in ln and ".text." in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    agg = {}
    for a, s, n, smp in sass:
        key = line_of.get(a - base, ("?", 0))
        e = agg.setdefault(key, [0, 0])
        e[0] += n; e[1] += smp
    ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    srcs = {}
    print(f"total warp instructions {ti}, samples {ts}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        f, l = key
        text = ""
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), f)
        if os.path.exists(path):
            srcs.setdefault(path, open(path).read().splitlines())
            if 0 < l <= len(srcs[path]):
                text = srcs[path][l - 1].strip()[:100]
        print(f"{v[1] / max(1, ts) * 100:5.1f}% smp {v[0] / max(1, ti) * 100:5.1f}% inst  {f}:{l:<4} {text}")


if __name__ == "__main__":
    main()
