#!/usr/bin/env python
"""Per-kernel device times of assembleJacobian / assembleRes / SpMV for one mesh per element family
(mid-size by default; `full` = the BASELINE sizes). TACSB200_DIRECT_KINDS variants can be listed with --direct."""
import argparse
import ctypes as C
import gc
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import tacs_b200  # noqa: E402
from tacs_b200 import TACS as T, meshgen  # noqa: E402

lib = tacs_b200.load()
assert lib.init(0) == 0


def run(name, mesh_f, elem_f, reps=5):
    t0 = time.time()
    cr, a = meshgen.build_model(T, lib, mesh_f(), [elem_f()])
    A = a.createMat()
    t1 = time.time()
    res, x, y = a.createVec(), a.createVec(), a.createVec()
    x.setArray(meshgen.hash_vector(x.getSize()))
    a.applyBCs(x)
    a.setVariables(x)
    ne = a.getNumElements()
    bs, nr, nc, nnzb = A.getSizes()
    lib.time_assemble_jacobian(a.h, 1.0, 0.0, 0.0, res.h, A.h, 2)
    mk, ck = np.zeros(8), np.zeros(8, np.int64)
    lib.profile_enable(1)
    lib.profile_collect(mk.ctypes.data_as(C.POINTER(C.c_double)), ck.ctypes.data_as(C.POINTER(C.c_long)))
    ms = lib.time_assemble_jacobian(a.h, 1.0, 0.0, 0.0, res.h, A.h, reps) / reps
    lib.profile_collect(mk.ctypes.data_as(C.POINTER(C.c_double)), ck.ctypes.data_as(C.POINTER(C.c_long)))
    lib.profile_enable(0)
    named = {ln.split("|")[0]: round(float(ln.split("|")[2]) / reps, 3) for ln in lib.profile_named().decode().splitlines()}
    msr = lib.time_assemble_res(a.h, res.h, reps) / reps
    lib.time_mat_mult(A.h, x.h, y.h, 3)
    mss = lib.time_mat_mult(A.h, x.h, y.h, 20) / 20
    bytes_ = nnzb * (8 * bs * bs + 4) + 4 * (nr + 1) + 16 * bs * nr
    A.mult(x, y)
    print(json.dumps(dict(name=name, direct=os.environ.get("TACSB200_DIRECT_KINDS", "all"), elems=ne, nnzb=nnzb,
                          setup_s=round(t1 - t0, 2), jac_ms=round(ms, 3), kernels=named,
                          M_elem_per_s=round(ne / ms * 1e-3, 1), res_ms=round(msr, 3), spmv_ms=round(mss, 4),
                          spmv_gbs=round(bytes_ / mss * 1e-6), ynorm=y.norm(), resnorm=res.norm())), flush=True)
    del A, a, cr, res, x, y
    gc.collect()


CASES = {
    "quad4": (lambda: meshgen.plate(2, 1000, 1000), lambda: meshgen.iso_shell_element(T, lib, 2)),
    "quad9": (lambda: meshgen.cylinder(3, 300, 300), lambda: meshgen.composite_shell_element(T, lib, 3)),
    "hex8": (lambda: meshgen.cube(2, 100), lambda: meshgen.solid_element(T, lib, 2)),
    "hex27": (lambda: meshgen.cube(3, 40), lambda: meshgen.solid_element(T, lib, 3)),
}
FULL = {
    "quad9": (lambda: meshgen.cylinder(3, 1000, 2000), lambda: meshgen.composite_shell_element(T, lib, 3)),
    "hex8": (lambda: meshgen.cube(2, 200), lambda: meshgen.solid_element(T, lib, 2)),
    "hex27": (lambda: meshgen.cube(3, 100), lambda: meshgen.solid_element(T, lib, 3)),
}

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*", default=["quad4", "quad9", "hex8", "hex27"])
    ap.add_argument("--full", action="store_true")
    ap.add_argument("--direct", default="", help="comma list of TACSB200_DIRECT_KINDS values to compare, e.g. 0,15")
    args = ap.parse_args()
    variants = [v for v in args.direct.split(",") if v] or [None]
    for name in args.cases:
        mesh_f, elem_f = (FULL if args.full and name in FULL else CASES)[name]
        for v in variants:
            if v is not None:
                os.environ["TACSB200_DIRECT_KINDS"] = v
            run(name + ("_full" if args.full and name in FULL else ""), mesh_f, elem_f)
