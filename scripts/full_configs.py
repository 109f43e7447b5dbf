#!/usr/bin/env python
"""BASELINE configs[2] and [4] at full size on one B200: Quad9 composite cylinder (~2M elements) and
100^3 hex27 solid. Reports timings and size-independent invariants (linearity of the operator, res = K u,
symmetry on vectors satisfying the BCs)."""
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tacs_b200
from tacs_b200 import TACS as T, meshgen

lib = tacs_b200.load()
assert lib.init(0) == 0


def run(name, mesh, elem, reps=3):
    t0 = time.time()
    cr, a = meshgen.build_model(T, lib, mesh, [elem])
    A = a.createMat()
    t1 = time.time()
    res, x, y, z, w = a.createVec(), a.createVec(), a.createVec(), a.createVec(), a.createVec()
    n = x.getSize()
    x.setArray(meshgen.hash_vector(n))
    a.applyBCs(x)
    a.setVariables(x)
    lib.time_assemble_jacobian(a.h, 1.0, 0.0, 0.0, res.h, A.h, 1)
    ms = lib.time_assemble_jacobian(a.h, 1.0, 0.0, 0.0, res.h, A.h, reps) / reps
    msr = lib.time_assemble_res(a.h, res.h, reps) / reps
    lib.time_mat_mult(A.h, x.h, y.h, 2)
    mss = lib.time_mat_mult(A.h, x.h, y.h, 10) / 10
    bs, nr, nc, nnzb = A.getSizes()
    bytes_ = nnzb * (8 * bs * bs + 4) + 4 * (nr + 1) + 16 * bs * nr
    # invariants
    A.mult(x, y)
    yn = y.norm()
    z.copyValues(x); z.scale(-2.0); A.mult(z, w); w.axpy(2.0, y)
    lin = w.norm() / yn
    a.assembleRes(res); res.axpy(-1.0, y); a.applyBCs(res)
    resid = res.norm() / yn
    z.setArray(meshgen.hash_vector(n)[::-1].copy()); a.applyBCs(z); A.mult(z, w)
    sym = abs(x.dot(w) - z.dot(y)) / abs(z.dot(y))
    print(json.dumps(dict(name=name, elems=a.getNumElements(), dof=n, nnzb=nnzb, setup_s=round(t1 - t0, 1),
                          jac_ms=round(ms, 3), jac_elem_per_s=a.getNumElements() / ms * 1e3, res_ms=round(msr, 3),
                          spmv_ms=round(mss, 4), spmv_gbs=bytes_ / mss * 1e-6, ynorm=yn, linearity=lin,
                          res_minus_Ku=resid, symmetry=sym)), flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "c3"):
    run("C3 quad9 composite cylinder 1000x2000", meshgen.cylinder(3, 1000, 2000), meshgen.composite_shell_element(T, lib, 3))
if which in ("all", "c5"):
    run("C5 hex27 100^3", meshgen.cube(3, 100), meshgen.solid_element(T, lib, 3))
