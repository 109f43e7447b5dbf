"""GPU: the CUDA path through the C ABI against the oracle restatement, the compiled reference
(when oracle/_ref travelled to the box) and the committed golden fixtures."""
import os

import numpy as np
import pytest

from tacs_b200 import TACS as T
from tacs_b200 import meshgen
from tests import common, oracle_port
from tests.common import TOL, relerr

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _desc_for(name):
    if "composite" in name:
        return oracle_port.composite_shell_desc(axis=(1.0, 0.3, 0.2))
    if name.startswith("hex"):
        return oracle_port.solid_desc()
    return oracle_port.iso_shell_desc(t=0.02, transform=0 if "natural" in name else 1, axis=(1.0, 0.3, 0.2))


def test_element_kernels_match_oracle(lib):
    for name, kind, elem in common.element_cases(lib):
        order = 2 if kind in (1, 3) else 3
        X, u, a = (common.shell_batch if kind <= 2 else common.solid_batch)(order, 37, seed=kind)
        res, mat = elem.addJacobian(1.3, 0.0, 0.7, X, u, None, a)
        desc = _desc_for(name)
        for e in range(X.shape[0]):
            r0, m0 = oracle_port.element(kind, desc, X[e], u[e], a[e], alpha=1.3, gamma=0.7)
            assert relerr(mat[e], m0) < TOL, (name, e)
            assert relerr(res[e], r0) < TOL, (name, e)
        # residual-only entry point
        res2 = elem.addResidual(X, u, None, a)
        assert relerr(res2, res) < TOL, name


def test_element_kernels_match_reference(lib, ref):
    for (name, kind, elem), (_, _, relem) in zip(common.element_cases(lib), common.element_cases(ref)):
        order = 2 if kind in (1, 3) else 3
        X, u, a = (common.shell_batch if kind <= 2 else common.solid_batch)(order, 11, seed=100 + kind)
        res, mat = elem.addJacobian(0.9, 0.0, 0.25, X, u, None, a)
        res_ref, mat_ref = relem.addJacobian(0.9, 0.0, 0.25, X, u, None, a)
        assert relerr(mat, mat_ref) < TOL, name
        assert relerr(res, res_ref) < TOL, name


def test_constitutive_matches_oracle(lib):
    for order in (2, 3):
        con = meshgen.iso_shell_element(T, lib, order, t=0.02).con
        d = oracle_port.iso_shell_desc(t=0.02)
        assert np.array_equal(con.evalTangentStiffness(), d[:22]) and np.array_equal(con.evalMassMoments(), d[22:25])
        con = meshgen.composite_shell_element(T, lib, order).con
        d = oracle_port.composite_shell_desc()
        assert np.array_equal(con.evalTangentStiffness(), d[:22]) and np.array_equal(con.evalMassMoments(), d[22:25])
        con = meshgen.solid_element(T, lib, order).model.con
        assert np.array_equal(con.evalTangentStiffness(), oracle_port.solid_desc()[:21])


@pytest.mark.parametrize("name", sorted(common.SMALL_MODELS))
def test_assembled_model_matches_golden_and_oracle(lib, name):
    r = common.run_model(lib, name)
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert np.array_equal(r["new_nodes"], g["new_nodes"])  # node map, bit exact
    assert np.array_equal(r["rowp"], g["rowp"]) and np.array_equal(r["cols"], g["cols"])  # sparsity, bit exact
    assert relerr(r["A"], g["A"]) < TOL
    assert relerr(r["res"], g["res"]) < TOL
    assert relerr(r["res_only"], g["res"]) < TOL
    assert relerr(r["y"], g["y"]) < TOL
    desc = oracle_port.composite_shell_desc() if "cylinder" in name else None
    o = oracle_port.assemble(r["mesh"], r["kind"], desc=desc, vars=r["u"], x=r["x"])
    assert relerr(r["A"], o["A"]) < TOL and relerr(r["res"], o["res"]) < TOL and relerr(r["y"], o["y"]) < TOL


@pytest.mark.parametrize("name", ["quad4_plate", "hex8_cube"])
def test_assembled_model_matches_reference(lib, ref, name):
    r, q = common.run_model(lib, name), common.run_model(ref, name)
    assert np.array_equal(r["new_nodes"], q["new_nodes"])
    assert np.array_equal(r["conn"][0], q["conn"][0]) and np.array_equal(r["conn"][1], q["conn"][1])
    assert np.array_equal(r["rowp"], q["rowp"]) and np.array_equal(r["cols"], q["cols"])
    assert relerr(r["A"], q["A"]) < TOL and relerr(r["res"], q["res"]) < TOL and relerr(r["y"], q["y"]) < TOL


def test_jacobian_with_mass_and_scaling(lib):
    """assembleJacobian(alpha, beta, gamma) with a non-zero acceleration state."""
    mesh = meshgen.plate(2, 6, 5)
    elem = meshgen.iso_shell_element(T, lib, 2)
    creator, asm = meshgen.build_model(T, lib, mesh, [elem])
    A, res, u, a = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec()
    n = u.getSize()
    u.setArray(meshgen.hash_vector(n))
    a.setArray(3.0 * meshgen.hash_vector(n)[::-1].copy())
    asm.applyBCs(u)
    asm.setVariables(u, None, a)
    asm.assembleJacobian(0.7, 0.0, 2.5, res, A)
    o = oracle_port.assemble(mesh, 1, vars=u.getArray(), ddvars=a.getArray(), alpha=0.7, gamma=2.5)
    assert relerr(A.getValues(), o["A"]) < TOL and relerr(res.getArray(), o["res"]) < TOL


def test_vector_kernels(lib):
    mesh = meshgen.cube(2, 6)
    creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.solid_element(T, lib, 2)])
    rng = np.random.default_rng(3)
    vecs = [asm.createVec() for _ in range(5)]
    host = [rng.standard_normal(vecs[0].getSize()) for _ in vecs]
    for v, h in zip(vecs, host):
        v.setArray(h)
    x, y = vecs[0], vecs[1]
    assert abs(x.norm() - np.linalg.norm(host[0])) < 1e-12 * np.linalg.norm(host[0])
    assert abs(x.dot(y) - host[0] @ host[1]) < 1e-12 * np.linalg.norm(host[0]) * np.linalg.norm(host[1])
    md = x.mdot(vecs[1:])
    assert np.allclose(md, [host[0] @ h for h in host[1:]], rtol=0, atol=1e-11)
    y.axpy(0.37, x)
    host[1] = host[1] + 0.37 * host[0]
    assert relerr(y.getArray(), host[1]) < 1e-15
    y.axpby(-1.2, 0.4, x)
    host[1] = -1.2 * host[0] + 0.4 * host[1]
    assert relerr(y.getArray(), host[1]) < 1e-15
    y.scale(2.5)
    assert relerr(y.getArray(), 2.5 * host[1]) < 1e-15
    y.copyValues(x)
    assert np.array_equal(y.getArray(), host[0])
    y.zeroEntries()
    assert not y.getArray().any()


def test_gmres_matches_numpy(lib):
    """Unpreconditioned GMRES(m) on the clamped hex8 cube against a dense solve of the same matrix."""
    mesh = meshgen.cube(2, 3)
    creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.solid_element(T, lib, 2)])
    A, res, b, x = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec()
    asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    n = b.getSize()
    f = np.zeros(n)
    f[2::3] = 1.0
    b.setArray(f)
    asm.applyBCs(b)
    ksm = T.KSM(lib, A, m=n, nrestart=0)
    ksm.setTolerances(1e-14, 1e-30)
    assert ksm.solve(b, x) == 1
    rowp, cols = A.getPattern()
    vals = A.getValues()
    K = np.zeros((n, n))
    for r in range(rowp.size - 1):
        for k in range(rowp[r], rowp[r + 1]):
            K[3 * r:3 * r + 3, 3 * cols[k]:3 * cols[k] + 3] = vals[k]
    want = np.linalg.solve(K, b.getArray())
    assert relerr(x.getArray(), want) < 1e-10  # north_star: displacements within 1e-10


def test_known_answers_and_invariants_at_scale(lib):
    """300x300 Quad4 plate: |A x| of the reference build and size-independent properties."""
    mesh = meshgen.plate(2, 300, 300)
    creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.iso_shell_element(T, lib, 2)])
    A, res, x, y, z = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    x.setArray(meshgen.hash_vector(x.getSize()))
    asm.applyBCs(x)
    A.mult(x, y)
    assert abs(y.norm() - 2.328079041928202e+02) < 1e-12 * 2.328079041928202e+02
    # linearity of the operator
    z.copyValues(x)
    z.scale(-2.0)
    w = asm.createVec()
    A.mult(z, w)
    w.axpy(2.0, y)
    assert w.norm() < 1e-12 * y.norm()
    # symmetry of the constrained operator on vectors that satisfy the BCs: x.(A z) == z.(A x)
    z.setArray(meshgen.hash_vector(z.getSize())[::-1].copy())
    asm.applyBCs(z)
    A.mult(z, w)
    assert abs(x.dot(w) - z.dot(y)) < 1e-11 * abs(z.dot(y))
    # residual of a linear model is K u: assembleRes(u) == A u on unconstrained dofs
    asm.setVariables(x)
    asm.assembleRes(res)
    res.axpy(-1.0, y)
    asm.applyBCs(res)
    assert res.norm() < 1e-12 * y.norm()
