"""Every element kernel once on a tiny mesh (for compute-sanitizer racecheck / memcheck runs)."""
import sys
sys.path.insert(0, '/root/repo')
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
    print(name, a.getNumElements(), "%.6e" % y.norm(), flush=True)
