"""Shared inputs of the parity tests: seeded element batches and small models."""
import numpy as np

from tacs_b200 import TACS as T
from tacs_b200 import meshgen

TOL = 1e-12  # north_star: residual / Jacobian / SpMV within 1e-12 relative in the max-norm


def relerr(got, want):
    got, want = np.asarray(got, float), np.asarray(want, float)
    scale = np.abs(want).max()
    return np.abs(got - want).max() / (scale if scale > 0 else 1.0)


def shell_batch(order, count, seed):
    """Curved, mildly distorted shell elements with random states."""
    rng = np.random.default_rng(seed)
    n = order * order
    u = np.linspace(-1, 1, order)
    U, V = np.meshgrid(u, u, indexing="xy")
    X = np.zeros((count, n, 3))
    for e in range(count):
        a, b, c = rng.uniform(-0.2, 0.2, 3)
        base = np.stack([0.5 * U + 0.1 * V + 0.05 * U * V, 0.45 * V - 0.05 * U, a * U * U + b * V * V + c * U * V], -1)
        X[e] = base.reshape(n, 3) + 0.01 * rng.standard_normal((n, 3)) + rng.uniform(-1, 1, 3)
    vars = 1e-3 * rng.standard_normal((count, 6 * n))
    ddvars = 1e-2 * rng.standard_normal((count, 6 * n))
    return X.reshape(count, 3 * n), vars, ddvars


def solid_batch(order, count, seed):
    rng = np.random.default_rng(seed)
    n = order ** 3
    u = np.linspace(0, 1, order)
    W, V, U = np.meshgrid(u, u, u, indexing="ij")
    base = np.stack([U.ravel(), V.ravel(), W.ravel()], -1)
    X = base[None] * rng.uniform(0.5, 1.5, (count, 1, 3)) + 0.04 * rng.standard_normal((count, n, 3))
    vars = 1e-3 * rng.standard_normal((count, 3 * n))
    ddvars = 1e-2 * rng.standard_normal((count, 3 * n))
    return X.reshape(count, 3 * n), vars, ddvars


def element_cases(lib):
    """(name, kind, element object) for every supported family / transform / constitutive combination."""
    cases = []
    for order in (2, 3):
        for tr in ("natural", "refaxis"):
            cases.append((f"quad{order*order}-iso-{tr}", order - 1,
                          meshgen.iso_shell_element(T, lib, order, t=0.02, transform=tr, axis=(1.0, 0.3, 0.2))))
        cases.append((f"quad{order*order}-composite", order - 1,
                      meshgen.composite_shell_element(T, lib, order, axis=(1.0, 0.3, 0.2))))
    for order in (2, 3):
        cases.append((f"hex{order**3}", order + 1, meshgen.solid_element(T, lib, order)))
    return cases


SMALL_MODELS = {
    # name: (mesh factory, kind, element factory)
    "quad4_plate": (lambda: meshgen.plate(2, 13, 9), 1, lambda lib: meshgen.iso_shell_element(T, lib, 2)),
    "quad9_plate": (lambda: meshgen.plate(3, 7, 5), 2, lambda lib: meshgen.iso_shell_element(T, lib, 3)),
    "quad4_cylinder": (lambda: meshgen.cylinder(2, 9, 16, defect=0.1), 1,
                       lambda lib: meshgen.composite_shell_element(T, lib, 2)),
    "quad9_cylinder": (lambda: meshgen.cylinder(3, 5, 8, defect=0.1), 2,
                       lambda lib: meshgen.composite_shell_element(T, lib, 3)),
    "hex8_cube": (lambda: meshgen.cube(2, 5), 3, lambda lib: meshgen.solid_element(T, lib, 2)),
    "hex27_cube": (lambda: meshgen.cube(3, 3), 4, lambda lib: meshgen.solid_element(T, lib, 3)),
}


def run_model(lib, name, with_state=True):
    """Assemble a small model through a bound library; returns host copies of everything compared."""
    mesh_f, kind, elem_f = SMALL_MODELS[name]
    mesh = mesh_f()
    elem = elem_f(lib)
    creator, asm = meshgen.build_model(T, lib, mesh, [elem])
    A, res, x, y, u = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    n = u.getSize()
    if with_state:
        u.setArray(meshgen.hash_vector(n))
        asm.applyBCs(u)
        asm.setVariables(u)
    asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    x.setArray(meshgen.hash_vector(n)[::-1].copy())
    asm.applyBCs(x)
    A.mult(x, y)
    res2 = asm.createVec()
    asm.assembleRes(res2)
    rowp, cols = A.getPattern()
    out = dict(mesh=mesh, kind=kind, new_nodes=creator.getNodeNums(), rowp=rowp, cols=cols, A=A.getValues(),
               res=res.getArray(), res_only=res2.getArray(), y=y.getArray(), x=x.getArray(), u=u.getArray(),
               conn=asm.getElementConnectivity(), nowned=asm.getNumOwnedNodes())
    out["_keep"] = (creator, asm, A, elem)
    return out
