#!/usr/bin/env python
"""Condensed view of a bench.py JSON line (file argument)."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("N=%d value %.1f M elem/s  %.3f ms/step  e2e %.1f M (%.3f ms)  res %.3f ms  spmv %.3f ms %.0f GB/s (%.2f)" % (
    d["n_gpus"], d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"].get("ms_per_step", 0),
    d["assemble_res"]["ms"], d["spmv"]["ms"], d["spmv"]["achieved"], d["spmv"]["frac"]))
def show_kernels(ks):
    for k in ks:
        if k.get("achieved") is None:
            print("   %-28s %8.3f ms" % (k["kernel"], k["ms"]))
        else:
            print("   %-28s %8.3f ms  %9.2f %s  frac %.3f%s" % (k["kernel"], k["ms"], k["achieved"], k["unit"], k["frac"],
                                                              "  (overlapped)" if k.get("overlapped") else ""))


show_kernels(d["kernels"])
print("   step", d.get("step_accounting"))
print("   plan", d.get("plan"))
for key in ("parity", "fullsize", "gmres", "clocks", "setup_s"):
    if d.get(key) is not None:
        print("  ", key, json.dumps(d[key])[:400])
for c in ("c4", "c3", "c5"):
    x = d.get(c)
    if not x:
        continue
    if "error" in x:
        print(c, x)
        continue
    print("%s N=%d jac %.3f ms (%.1f M elem/s) res %.3f ms spmv %.3f ms %.0f GB/s (%.2f)  setup %s" % (
        c, x["n_gpus"], x["jac_ms"], x["elements_per_s"] / 1e6, x["res_ms"], x["spmv_ms"], x["spmv_gbs_aggregate"],
        x["spmv_frac_of_hbm_peak"], {k: round(v, 1) for k, v in x["setup_s"].items()}))
    show_kernels(x["kernels"])
    print("   step", x.get("step_accounting"))
    print("   plan", x.get("plan"))
    for key in ("fullsize", "gmres"):
        if x.get(key) is not None:
            print("  ", key, json.dumps(x[key])[:300])
