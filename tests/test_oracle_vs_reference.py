"""CPU: pin the plain-C oracle restatement (oracle/tacs_oracle.c) to the compiled reference
(oracle/_ref/libtacs_ref.so) and to the committed golden fixtures generated from it."""
import os

import numpy as np
import pytest

from tacs_b200 import TACS as T
from tests import common, oracle_port
from tests.common import TOL, relerr

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("order", [2, 3])
def test_constitutive_matches_reference(ref, order):
    from tacs_b200 import meshgen

    iso = meshgen.iso_shell_element(T, ref, order, t=0.02).con
    d = oracle_port.iso_shell_desc(t=0.02)
    assert np.array_equal(iso.evalTangentStiffness(), d[:22])
    assert np.array_equal(iso.evalMassMoments(), d[22:25])
    comp = meshgen.composite_shell_element(T, ref, order).con
    d = oracle_port.composite_shell_desc()
    assert np.array_equal(comp.evalTangentStiffness(), d[:22])
    assert np.array_equal(comp.evalMassMoments(), d[22:25])
    solid = meshgen.solid_element(T, ref, order).model.con
    assert np.array_equal(solid.evalTangentStiffness(), oracle_port.solid_desc()[:21])


def _desc_for(name):
    if "composite" in name:
        return oracle_port.composite_shell_desc(axis=(1.0, 0.3, 0.2))
    if name.startswith("hex"):
        return oracle_port.solid_desc()
    return oracle_port.iso_shell_desc(t=0.02, transform=0 if "natural" in name else 1, axis=(1.0, 0.3, 0.2))


def test_elements_match_reference(ref):
    for name, kind, elem in common.element_cases(ref):
        X, u, a = (common.shell_batch if kind <= 2 else common.solid_batch)(2 if kind in (1, 3) else 3, 3, seed=kind)
        res_ref, mat_ref = elem.addJacobian(1.3, 0.0, 0.7, X, u, None, a)
        desc = _desc_for(name)
        for e in range(X.shape[0]):
            res, mat = oracle_port.element(kind, desc, X[e], u[e], a[e], alpha=1.3, gamma=0.7)
            assert relerr(mat, mat_ref[e]) < TOL, name
            assert relerr(res, res_ref[e]) < TOL, name


@pytest.mark.parametrize("name", ["quad4_plate", "quad9_plate", "hex8_cube", "hex27_cube"])
def test_assembled_model_matches_reference(ref, name):
    r = common.run_model(ref, name)
    o = oracle_port.assemble(r["mesh"], r["kind"], vars=r["u"], x=r["x"])
    assert np.array_equal(o["new_nodes"], r["new_nodes"])  # first-touch numbering, bit exact
    assert np.array_equal(o["rowp"], r["rowp"]) and np.array_equal(o["cols"], r["cols"])
    assert relerr(o["A"], r["A"]) < TOL
    assert relerr(o["res"], r["res"]) < TOL
    assert relerr(o["res"], r["res_only"]) < TOL
    assert relerr(o["y"], r["y"]) < TOL


@pytest.mark.parametrize("name", sorted(common.SMALL_MODELS))
def test_oracle_matches_golden(name):
    """Fixtures were produced by tests/golden/make_golden.py from the compiled reference."""
    path = os.path.join(GOLDEN, name + ".npz")
    g = np.load(path)
    mesh_f, kind, _ = common.SMALL_MODELS[name]
    desc = None
    if "cylinder" in name:
        desc = oracle_port.composite_shell_desc()
    o = oracle_port.assemble(mesh_f(), kind, desc=desc, vars=g["u"], x=g["x"])
    assert np.array_equal(o["new_nodes"], g["new_nodes"])
    assert np.array_equal(o["rowp"], g["rowp"]) and np.array_equal(o["cols"], g["cols"])
    assert relerr(o["A"], g["A"]) < TOL
    assert relerr(o["res"], g["res"]) < TOL
    assert relerr(o["y"], g["y"]) < TOL
