#!/usr/bin/env python
"""Per-source-line / per-phase breakdown of one profiled kernel.

Joins the SASS page of an .ncu-rep (instructions executed, stall samples per SASS instruction) with the
line table of the cubin the report was taken from (`nvdisasm -g`), because `ncu --page source --csv` only
prints the SASS view.

usage: python scripts/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTR [OBJ=tacs_b200/csrc/_build/kernels.o]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj, mangled_substr):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    out = {}
    active = False
    cur = ("?", 0)
    for line in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):", line)
        if m:
            active = mangled_substr in m.group(1)
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    obj = sys.argv[3] if len(sys.argv) > 3 else "tacs_b200/csrc/_build/kernels.o"
    mang = sys.argv[4] if len(sys.argv) > 4 else kern
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], stdout=subprocess.PIPE,
                         text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # first kernel block only
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[hdr_i + 1:]:
        if not r or not r[0].startswith("0x"):
            break
        body.append(r)
    base = int(body[0][0], 16)
    table = sass_lines(obj, mang)
    per_line = defaultdict(lambda: [0, 0, 0, 0])  # warp insts, samples, fp64 insts, lds/sts
    tot_i = tot_s = 0
    for r in body:
        off = int(r[0], 16) - base
        (fl, op) = table.get(off, (("?", 0), r[1]))
        ni = int(r[col["Instructions Executed"]])
        ns = int(r[col["Warp Stall Sampling (All Samples)"]])
        e = per_line[fl]
        e[0] += ni
        e[1] += ns
        opc = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        if opc.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")):
            e[2] += ni
        if opc.startswith(("LDS", "STS")):
            e[3] += ni
        tot_i += ni
        tot_s += ns
    print(f"kernel {kern}: {tot_i} warp instructions, {tot_s} stall samples")
    print(f"{'file:line':28s} {'insts':>12s} {'%':>6s} {'samples':>9s} {'%':>6s} {'fp64':>11s} {'lds/sts':>10s}")
    for fl, e in sorted(per_line.items(), key=lambda kv: -kv[1][1]):
        if e[0] * 200 < tot_i and e[1] * 200 < tot_s:
            continue
        print(f"{fl[0] + ':' + str(fl[1]):28s} {e[0]:12d} {100.0 * e[0] / tot_i:6.2f} {e[1]:9d} {100.0 * e[1] / max(tot_s, 1):6.2f} "
              f"{e[2]:11d} {e[3]:10d}")


if __name__ == "__main__":
    main()
