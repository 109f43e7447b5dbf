#!/usr/bin/env python
"""Per-function breakdown of one profiled element kernel: joins the SASS page of an .ncu-rep with the cubin's line
table (scripts/ncu_lines.py) and groups source lines by the enclosing function of elem_phases.cuh / by 20-line
bands of kernels.cu.  usage: python scripts/ncu_phases.py REPORT.ncu-rep MANGLED_SUBSTR [NWARPS_ITER]"""
import csv
import re
import subprocess
import sys
import os
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines


def function_ranges(path):
    """(name, first_line, last_line) of every TB2_HD function in a header (brace matching from the signature)."""
    src = open(path).read().splitlines()
    out = []
    i = 0
    while i < len(src):
        m = re.match(r"\s*TB2_HD\s+[\w:<>\s\*&]+?\s+(\w+)\s*\(", src[i])
        if m:
            name, depth, j, seen = m.group(1), 0, i, False
            while j < len(src):
                depth += src[j].count("{") - src[j].count("}")
                if "{" in src[j]:
                    seen = True
                if seen and depth == 0:
                    break
                j += 1
            out.append((name, i + 1, j + 1))
            i = j
        i += 1
    return out


def main():
    rep, mang = sys.argv[1], sys.argv[2]
    nw = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    obj = os.path.join(root, "tacs_b200/csrc/_build/kernels.o")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    body = []
    for r in rows[hi + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        body.append(r)
    base = int(body[0][0], 16)
    table = ncu_lines.sass_lines(obj, mang)
    ranges = function_ranges(os.path.join(root, "tacs_b200/csrc/elem_phases.cuh"))
    agg = defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    for r in body:
        off = int(r[0], 16) - base
        (f, l), _ = table.get(off, (("?", 0), r[1]))
        ni = int(r[col["Instructions Executed"]])
        ns = int(r[col["Warp Stall Sampling (All Samples)"]])
        g = f
        if f == "elem_phases.cuh":
            for name, a, b in ranges:
                if a <= l <= b:
                    g = name
        elif f == "kernels.cu":
            g = "kernels.cu:%d-%d" % ((l // 20) * 20, (l // 20) * 20 + 19)
        e = agg[g]
        e[0] += ni
        e[1] += ns
        toks = r[1].split()
        opc = toks[1] if toks[0].startswith("@") else toks[0]
        if opc.startswith(("DFMA", "DMUL", "DADD", "DMMA")):
            e[2] += ni
        if opc.startswith(("LDS", "STS")):
            e[3] += ni
        if "L1 Wavefronts Shared" in col:
            e[4] += int(r[col["L1 Wavefronts Shared"]] or 0)
            e[5] += int(r[col["L1 Wavefronts Shared Ideal"]] or 0)
    ti = sum(e[0] for e in agg.values())
    ts = sum(e[1] for e in agg.values())
    if nw <= 0:
        nw = 1.0
    print(f"total warp instructions {ti} ({ti / nw:.1f} per unit), samples {ts}")
    tw = sum(e[4] for e in agg.values())
    print(f"shared-memory wavefronts {tw} ({tw / nw:.1f} per unit)")
    print(f"{'group':34s} {'inst/unit':>10s} {'inst%':>7s} {'time%':>7s} {'fp64/unit':>10s} {'lds+sts':>8s} {'smem wf':>8s} {'ideal':>8s}")
    for g, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if e[1] * 500 < ts and e[0] * 500 < ti:
            continue
        print(f"{g:34s} {e[0] / nw:10.1f} {100 * e[0] / ti:7.2f} {100 * e[1] / max(ts, 1):7.2f} {e[2] / nw:10.1f} {e[3] / nw:8.1f} {e[4] / nw:8.1f} {e[5] / nw:8.1f}")


if __name__ == "__main__":
    main()
