"""Every element / gather / SpMV / Krylov kernel once on a tiny mesh (for compute-sanitizer racecheck / memcheck
runs; TACSB200_SPMV3=stream|rows selects the 3x3 product)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tacs_b200
from tacs_b200 import TACS as T, meshgen

lib = tacs_b200.load()
assert lib.init(0) == 0
cases = [("quad4-iso", meshgen.plate(2, 6, 5), meshgen.iso_shell_element(T, lib, 2)),
         ("quad4-composite", meshgen.cylinder(2, 4, 6, defect=0.1), meshgen.composite_shell_element(T, lib, 2)),
         ("quad9-iso", meshgen.plate(3, 4, 3), meshgen.iso_shell_element(T, lib, 3)),
         ("quad4-offset", meshgen.plate(2, 4, 3), meshgen.iso_shell_element(T, lib, 2, t=0.02, transform="natural")),
         ("hex8", meshgen.cube(2, 3), meshgen.solid_element(T, lib, 2)),
         ("hex27", meshgen.cube(3, 2), meshgen.solid_element(T, lib, 3))]
for name, mesh, elem in cases:
    cr, a = meshgen.build_model(T, lib, mesh, [elem])
    A, res, x, y = a.createMat(), a.createVec(), a.createVec(), a.createVec()
    x.setArray(meshgen.hash_vector(x.getSize()))
    a.applyBCs(x)
    a.setVariables(x, None, x)
    a.assembleJacobian(1.0, 0.0, 0.5, res, A)
    a.setVariables(x)
    a.zeroVariables()
    a.setVariables(x)
    a.assembleJacobian(1.0, 0.0, 0.0, res, A)
    a.assembleRes(res)
    A.mult(x, y)
    A.multTranspose(x, y)
    # host-buffer entry point (state pieces / element chunks), smoother with fused SpMV epilogues, device GMRES
    q, out = x.getArray(), np.zeros(x.getSize())
    a.assembleJacobianHost(1.0, 0.0, 0.0, q, out, A)
    lib.synchronize()
    pc = T.ChebyshevSmoother(lib, A, 3, iters=2)
    pc.factor()
    ksm = T.KSM(lib, A, 8, 1, pc=pc, isFlexible=1)
    ksm.setTolerances(1e-8, 1e-30)
    for _ in range(3):   # direct launch, graph capture, graph replay
        ksm.solve(res, y)
    if name.startswith("hex"):
        a.assembleMatType(T.GEOMETRIC_STIFFNESS_MATRIX, A)
    print(name, a.getNumElements(), "%.6e" % y.norm(), flush=True)
