"""Parity check of the distributed (one process per GPU) assembly + SpMV against the serial oracle.

Test infrastructure: used by tests/nccl_worker.py and, as the checker in front of the timed region, by
bench.py at N > 1 GPUs.  Every rank assembles its METIS part of a small mesh with the product library, and
every owned row of the distributed matrix / residual / A*x is compared with the serial oracle
(oracle/tacs_oracle.c) evaluated in the same global numbering."""
import numpy as np

from tacs_b200 import TACS as T
from tacs_b200 import meshgen
from tests.test_distributed_plan import CASES, serial_reference


def check_case(lib, name):
    """Returns dict(A=, res=, y=, res_only=, norm=, dot=, pattern_exact=) of relative max-norm errors on this
    rank's owned rows (pattern_exact: the owned rows' column sets equal the serial sparsity pattern)."""
    mesh_f, kind, elem_f, desc_f = CASES[name]
    mesh = mesh_f()
    bs = mesh["vars_per_node"]
    creator, asm = meshgen.build_model(T, lib, mesh, [elem_f(lib)])
    new_nodes = creator.getNodeNums()
    u, x, bc, serial = serial_reference(mesh, kind, desc_f(), new_nodes)
    lo, hi = asm.getOwnerRange()
    A, res, uv, xv, yv = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    uv.setArray(u[bs * lo:bs * hi])
    xv.setArray(x[bs * lo:bs * hi])
    asm.setVariables(uv)
    asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    A.mult(xv, yv)
    ar, ac = A.getPattern(0)
    av = A.getValues(0)
    br, bcs = A.getPattern(1)
    bv = A.getValues(1)
    ext_cols = A.getExtColNodes()
    npr = (hi - lo) - (br.size - 1)
    srow, scol, sA = serial["rowp"], serial["cols"], serial["A"]
    scale = np.abs(sA).max()
    worst, pattern_exact = 0.0, True
    for r in range(hi - lo):
        cols = [lo + c for c in ac[ar[r]:ar[r + 1]]]
        vals = [av[k] for k in range(ar[r], ar[r + 1])]
        if r >= npr:
            cols += [ext_cols[c] for c in bcs[br[r - npr]:br[r - npr + 1]]]
            vals += [bv[k] for k in range(br[r - npr], br[r - npr + 1])]
        order = np.argsort(cols)
        g = lo + r
        if not np.array_equal(np.asarray(cols)[order], scol[srow[g]:srow[g + 1]]):
            pattern_exact = False
            continue
        for m, o in enumerate(order):
            worst = max(worst, np.abs(vals[o] - sA[srow[g] + m]).max() / scale)
    out = {"A": float(worst), "pattern_exact": bool(pattern_exact)}
    out["res"] = float(np.abs(res.getArray() - serial["res"][bs * lo:bs * hi]).max() / np.abs(serial["res"]).max())
    out["y"] = float(np.abs(yv.getArray() - serial["y"][bs * lo:bs * hi]).max() / np.abs(serial["y"]).max())
    res2 = asm.createVec()
    asm.assembleRes(res2)
    out["res_only"] = float(np.abs(res2.getArray() - serial["res"][bs * lo:bs * hi]).max() /
                            np.abs(serial["res"]).max())
    want = float(np.linalg.norm(serial["y"]))
    out["norm"] = abs(yv.norm() - want) / want
    wdot = float(serial["y"] @ x)
    out["dot"] = abs(yv.dot(xv) - wdot) / abs(wdot)
    if name == "hex8_cube":
        # distributed GMRES (Gram-Schmidt reductions across the ranks, graphs replayed on the third solve) against a
        # direct solve of the serial oracle matrix
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        nb = srow.size - 1
        S = sp.bsr_matrix((sA.reshape(-1, bs, bs), scol, srow), shape=(bs * nb, bs * nb)).tocsc()
        b = meshgen.hash_vector(bs * nb) - 0.3e-3
        for g in bc:
            b[bs * g:bs * g + bs] = 0.0
        xs = spla.spsolve(S, b)
        bv, sol = asm.createVec(), asm.createVec()
        bv.setArray(b[bs * lo:bs * hi])
        ksm = T.KSM(lib, A, 40, 10)
        ksm.setTolerances(1e-13, 1e-30)
        for _ in range(3):
            flag = ksm.solve(bv, sol)
        out["gmres"] = float(np.abs(sol.getArray() - xs[bs * lo:bs * hi]).max() / np.abs(xs).max())
        out["gmres_converged"] = bool(flag == 1)
    return out


def check_all(lib, names=None):
    """{case: errors}; the worst relative error over all cases under "max"."""
    out = {}
    for name in (names or sorted(CASES)):
        out[name] = check_case(lib, name)
    keys = ("A", "res", "y", "res_only", "norm")
    out["max"] = {k: max(out[n][k] for n in out if n != "max") for k in keys}
    out["max"]["dot"] = max(out[n]["dot"] for n in out if n != "max")
    out["pattern_exact"] = all(out[n]["pattern_exact"] for n in out if n not in ("max",))
    out["max"]["gmres"] = max(out[n].get("gmres", 0.0) for n in out if n not in ("max", "pattern_exact"))
    out["gmres_converged"] = all(out[n].get("gmres_converged", True) for n in out if n not in ("max", "pattern_exact"))
    return out
