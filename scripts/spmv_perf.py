#!/usr/bin/env python
"""SpMV-only timing of the 3x3 (hex8 / hex27) and 6x6 (Quad4 / Quad9) matrices: ms, GB/s of algorithmic bytes, |y|."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacs_b200  # noqa: E402
from tacs_b200 import TACS as T, meshgen  # noqa: E402

lib = tacs_b200.load()
assert lib.init(0) == 0
CASES = {
    "hex8": (lambda n: meshgen.cube(2, n), lambda: meshgen.solid_element(T, lib, 2), 160),
    "hex27": (lambda n: meshgen.cube(3, n), lambda: meshgen.solid_element(T, lib, 3), 70),
    "quad4": (lambda n: meshgen.plate(2, n, n), lambda: meshgen.iso_shell_element(T, lib, 2), 1000),
    "quad9": (lambda n: meshgen.cylinder(3, n, n), lambda: meshgen.composite_shell_element(T, lib, 3), 400),
}
for name in sys.argv[1:] or ["hex8", "hex27"]:
    mesh_f, elem_f, n = CASES[name]
    cr, a = meshgen.build_model(T, lib, mesh_f(n), [elem_f()])
    A, res, x, y = a.createMat(), a.createVec(), a.createVec(), a.createVec()
    x.setArray(meshgen.hash_vector(x.getSize()))
    a.applyBCs(x)
    a.assembleJacobian(1.0, 0.0, 0.0, res, A)
    bs, nr, nc, nnzb = A.getSizes()
    lib.time_mat_mult(A.h, x.h, y.h, 3)
    ms = lib.time_mat_mult(A.h, x.h, y.h, 20) / 20
    bytes_ = nnzb * (8 * bs * bs + 4) + 4 * (nr + 1) + 16 * bs * nr
    A.mult(x, y)
    print(json.dumps(dict(name=name, n=n, variant=os.environ.get("TACSB200_SPMV3", "default"), nnzb=nnzb,
                          spmv_ms=round(ms, 4), gbs=round(bytes_ / ms * 1e-6), ynorm=repr(y.norm()))), flush=True)
    del A, a, cr, res, x, y
