#!/usr/bin/env python
"""ms per GMRES(m) iteration on the assembled C2 (or a cube) operator: python scripts/gmres_perf.py [quad4|hex8] [m]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacs_b200  # noqa: E402
from tacs_b200 import TACS as T, meshgen  # noqa: E402

lib = tacs_b200.load()
assert lib.init(0) == 0
case = sys.argv[1] if len(sys.argv) > 1 else "quad4"
m = int(sys.argv[2]) if len(sys.argv) > 2 else 30
if case == "quad4":
    mesh, elem = meshgen.plate(2, 1000, 1000), meshgen.iso_shell_element(T, lib, 2)
else:
    mesh, elem = meshgen.cube(2, 200), meshgen.solid_element(T, lib, 2)
cr, a = meshgen.build_model(T, lib, mesh, [elem])
A, res, x, sol = a.createMat(), a.createVec(), a.createVec(), a.createVec()
x.setArray(meshgen.hash_vector(x.getSize()))
a.applyBCs(x)
a.setVariables(x)
a.assembleJacobian(1.0, 0.0, 0.0, res, A)
for classical in (False, True):
    ksm = T.KSM(lib, A, m, 0)
    ksm.setTolerances(1e-30, 1e-300)
    ksm.setOrthoType(classical)
    for _ in range(2):
        ksm.solve(res, sol)
    lib.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        ksm.solve(res, sol)
    lib.synchronize()
    dt = (time.perf_counter() - t0) / reps
    n = x.getSize()
    bs, nr, nc, nnzb = A.getSizes()
    spmv = lib.time_mat_mult(A.h, x.h, sol.h, 20) / 20
    streams = sum(4 * i + 7 for i in range(m)) / m
    ideal = spmv + streams * n * 8 / 6.5469e12 * 1e3
    print(f"{case} n={n} m={m} {'CGS' if classical else 'MGS'}: {dt*1e3/ksm.getIterCount():.4f} ms/iter "
          f"({ksm.getIterCount()} iters), spmv {spmv:.4f} ms, MGS streaming bound {ideal:.4f} ms "
          f"(SpMV + {streams:.1f} vector passes at the HBM peak), resnorm {ksm.getResidualNorm():.6e}", flush=True)
