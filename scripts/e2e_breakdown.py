"""Wall-clock breakdown of bench.py's end-to-end step (host buffers in pinned memory)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import tacs_b200
from tacs_b200 import TACS as T, meshgen

lib = tacs_b200.load(); assert lib.init(0) == 0
mesh = meshgen.plate(2, 1000, 1000)
creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.iso_shell_element(T, lib, 2)])
A, res, x = asm.createMat(), asm.createVec(), asm.createVec()
n = x.getSize()
state = torch.empty(n, dtype=torch.float64).pin_memory(); out = torch.empty(n, dtype=torch.float64).pin_memory()
s_np, o_np = state.numpy(), out.numpy(); s_np[:] = meshgen.hash_vector(n)
t = np.zeros(5)
for it in range(12):
    a = time.perf_counter(); x.setArray(s_np)
    b = time.perf_counter(); asm.setVariables(x)
    c = time.perf_counter(); asm.assembleJacobian(1.0, 0.0, 0.0, res, A, wait=False)
    d = time.perf_counter(); lib.vec_get_array(res.h, tacs_b200.binding.dptr(o_np))
    e = time.perf_counter()
    if it >= 2: t += [b - a, c - b, d - c, e - d, e - a]
lib.synchronize()
print("ms per step: setArray %.3f setVariables %.3f assembleJacobian %.3f getArray %.3f total %.3f" % tuple(t / 10 * 1e3))
