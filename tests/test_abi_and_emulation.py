"""CPU: the C-ABI library loads and exports every symbol the header declares; the device element
phases replayed on the host (tests/emul) match the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from tests import common, oracle_port
from tests.common import TOL, relerr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import tacs_b200

    lib = tacs_b200.load()
    header = open(os.path.join(ROOT, "include", "tacs_b200.h")).read()
    names = sorted(set(re.findall(r"\b(tacsb200_[a-z0-9_]+)\s*\(", header)))
    assert len(names) > 80
    missing = [n for n in names if not hasattr(lib.dll, n)]
    assert not missing, missing
    assert lib.abi_version() == 1
    # the Python binding only refers to declared entry points
    for name in lib.signatures():
        assert "tacsb200_" + name in names, name


def test_no_compute_without_gpu_is_loud():
    """On a box without a GPU the library must refuse, not fall back."""
    import tacs_b200
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = tacs_b200.load()
    assert lib.init(0) != 0


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(ROOT, "tests", "emul", "_libemul.so")
    src = os.path.join(ROOT, "tests", "emul", "emul_elements.cpp")
    deps = [src] + [os.path.join(ROOT, "tacs_b200", "csrc", f) for f in ("elem_phases.cuh", "elem_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    DP = C.POINTER(C.c_double)
    L.emul_element.argtypes = [C.c_int, DP, DP, DP, DP, C.c_double, C.c_double, DP, DP]
    L.emul_residual.argtypes = [C.c_int, DP, DP, DP, DP]
    return L


@pytest.mark.parametrize("kind", [1, 2, 3, 4])
def test_device_phases_replayed_on_host_match_oracle(emul, kind):
    order = 2 if kind in (1, 3) else 3
    X, u, a = (common.shell_batch if kind <= 2 else common.solid_batch)(order, 4, seed=10 + kind)
    descs = [oracle_port.solid_desc()] if kind > 2 else [
        oracle_port.iso_shell_desc(t=0.02, tOffset=0.3, transform=0), oracle_port.composite_shell_desc(axis=(1, .3, .2)),
        oracle_port.iso_shell_desc(t=0.01, tOffset=0.0, transform=1, axis=(1, .3, .2)),
        oracle_port.iso_shell_desc(t=0.03, tOffset=0.0, transform=0)]
    DP = C.POINTER(C.c_double)
    for desc in descs:
        for e in range(X.shape[0]):
            res, mat = oracle_port.element(kind, desc, X[e], u[e], a[e], alpha=1.3, gamma=0.7)
            nv = res.size
            r2, m2 = np.zeros(nv), np.zeros(nv * nv)
            args = [np.ascontiguousarray(v) for v in (X[e], u[e], a[e], desc)]
            rc = emul.emul_element(kind, *[v.ctypes.data_as(DP) for v in args], 1.3, 0.7, r2.ctypes.data_as(DP),
                                   m2.ctypes.data_as(DP))
            assert rc == 0
            assert relerr(m2, mat.ravel()) < TOL
            assert relerr(r2, res) < TOL


@pytest.mark.parametrize("kind", [1, 2])
def test_residual_only_phases_match_oracle(emul, kind):
    """assembleRes fast path of the tensor-core shell kernels (state pushed through the tying space)."""
    X, u, a = common.shell_batch(kind + 1, 6, seed=77)
    DP = C.POINTER(C.c_double)
    for desc in (oracle_port.iso_shell_desc(t=0.01, tOffset=0.0, transform=1, axis=(1, .3, .2)),
                 oracle_port.iso_shell_desc(t=0.03, tOffset=0.0, transform=0),
                 oracle_port.composite_shell_desc(axis=(1, .3, .2))):
        for e in range(X.shape[0]):
            res, _ = oracle_port.element(kind, desc, X[e], u[e], 0.0 * a[e], alpha=1.0, gamma=0.0)
            r2 = np.zeros(res.size)
            args = [np.ascontiguousarray(v) for v in (X[e], u[e], desc)]
            assert emul.emul_residual(kind, *[v.ctypes.data_as(DP) for v in args], r2.ctypes.data_as(DP)) == 0
            assert relerr(r2, res) < TOL
