"""GPU: the reference-side binding (shim/) driven by a C++ application written against the reference's own API.

tests/shim/shim_driver.cpp builds models the way examples/plate/plate.cpp and examples/tutorial/tutorial.cpp do
(TACSCreator / hand-built TACSAssembler, the reference's element and constitutive classes), runs the reference's CPU
classes and the device classes TACSB200Assembler / TACSB200Mat / TACSB200Vec / TACSB200ChebyshevPc side by side --
including the reference's own GMRES class running unchanged on the device objects -- and compares sparsity pattern
(bit-exact), Jacobian, residual, A*x (1e-12) and linear-static displacements (1e-10). The binary is linked against the
unmodified reference (oracle/_ref) in the build container (`__graft_entry__.build()`) and travels to the GPU box."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "shim", "_build", "shim_driver")


@pytest.mark.parametrize("case", ["plate", "plate9", "direct", "cube", "cube27", "schur"])
def test_reference_driver_on_device_objects(lib, case):
    if not os.path.exists(DRIVER):
        pytest.skip("tests/shim/_build/shim_driver not built (needs /root/reference at build time)")
    proc = subprocess.run([DRIVER, case], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(proc.stdout)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert "ALL OK" in proc.stdout and "FAIL" not in proc.stdout.replace("FAILED", "")
