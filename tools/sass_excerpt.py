#!/usr/bin/env python
"""SASS evidence for profiles/: opcode histogram of one kernel of an object file and the longest run of one opcode with its
neighbourhood (the inner loop).  usage: python tools/sass_excerpt.py <object.o> <mangled-name-substring> <opcode> [context]"""
import collections
import re
import subprocess
import sys


def main():
    obj, kname, op = sys.argv[1:4]
    ctx = int(sys.argv[4]) if len(sys.argv) > 4 else 6
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(out) if "Function :" in l and kname in l)
    end = next((i for i in range(start + 1, len(out)) if "Function :" in out[i]), len(out))
    ins = []
    for l in out[start:end]:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(1), m.group(2).strip()))
    def opcode(t):
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        return t.split()[0]
    hist = collections.Counter(opcode(t).split(".")[0] for _, t in ins)
    full = collections.Counter(opcode(t) for _, t in ins)
    print(f"kernel {out[start].split(':', 1)[1].strip()}")
    print(f"{len(ins)} SASS instructions; by opcode: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(14)))
    wide = {k: v for k, v in full.items() if any(s in k for s in ("256", "128", "LDGSTS", "DFMA", "SHFL", "BAR", "WARPSYNC"))}
    print("of which: " + ", ".join(f"{k} {v}" for k, v in sorted(wide.items(), key=lambda kv: -kv[1])[:14]))
    # densest window of 48 instructions for the opcode
    hits = [1 if opcode(t).startswith(op) else 0 for _, t in ins]
    W = 48
    best, bi = -1, 0
    s = sum(hits[:W])
    for i in range(0, max(1, len(ins) - W)):
        if s > best:
            best, bi = s, i
        s += (hits[i + W] if i + W < len(ins) else 0) - hits[i]
    print(f"densest {W}-instruction window for {op}: {best} of {W}, at /*{ins[bi][0]}*/")
    for a, t in ins[max(0, bi - ctx):bi + W + ctx]:
        print(f"  /*{a}*/  {t}")


if __name__ == "__main__":
    main()
