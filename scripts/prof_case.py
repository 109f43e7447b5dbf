"""One small workload per element family for ncu captures: python scripts/prof_case.py quad4|quad9|hex8|hex27 [n]"""
import sys
sys.path.insert(0, '/root/repo')
import tacs_b200
from tacs_b200 import TACS as T, meshgen

lib = tacs_b200.load()
assert lib.init(0) == 0
case = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if case == 'quad4':
    mesh, elem = meshgen.plate(2, n or 500, n or 500), meshgen.iso_shell_element(T, lib, 2)
elif case == 'quad9':
    mesh, elem = meshgen.plate(3, n or 150, n or 150), meshgen.iso_shell_element(T, lib, 3)
elif case == 'quad9c':
    mesh, elem = meshgen.cylinder(3, n or 150, n or 150), meshgen.composite_shell_element(T, lib, 3)
elif case == 'hex8':
    mesh, elem = meshgen.cube(2, n or 60), meshgen.solid_element(T, lib, 2)
else:
    mesh, elem = meshgen.cube(3, n or 24), meshgen.solid_element(T, lib, 3)
cr, a = meshgen.build_model(T, lib, mesh, [elem])
A = a.createMat()
res, x, y = a.createVec(), a.createVec(), a.createVec()
x.setArray(meshgen.hash_vector(x.getSize()))
a.applyBCs(x)
a.setVariables(x)
for _ in range(3):
    a.assembleJacobian(1.0, 0.0, 0.0, res, A)
    A.mult(x, y)
    a.assembleRes(res)
print(case, a.getNumElements(), y.norm())
