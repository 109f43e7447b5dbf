"""ctypes front end of the plain-C oracle restatement (oracle/tacs_oracle.c) -- test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libtacs_oracle.so")

I, D = C.c_int, C.c_double
IP, DP = C.POINTER(C.c_int), C.POINTER(C.c_double)
_lib = None


def _p(a):
    if a is None:
        return None
    return a.ctypes.data_as(IP if a.dtype == np.int32 else DP)


def load():
    global _lib
    if _lib is None:
        src = os.path.join(ROOT, "oracle", "tacs_oracle.c")
        if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "port"], cwd=os.path.join(ROOT, "oracle"), stdout=subprocess.DEVNULL)
        L = C.CDLL(SO)
        L.oracle_shell_element.argtypes = [I, DP, DP, DP, I, DP, DP, DP, D, D, D, DP, DP]
        L.oracle_shell_element.restype = None
        L.oracle_solid_element.argtypes = [I, DP, DP, DP, DP, D, D, D, D, DP, DP]
        L.oracle_solid_element.restype = None
        L.oracle_iso_shell_stiffness.argtypes = [D, D, D, D, D, D, D, DP, DP]
        L.oracle_composite_shell_stiffness.argtypes = [I, DP, DP, DP, D, D, D, DP, DP]
        L.oracle_solid_stiffness.argtypes = [D, D, D, D, DP, DP]
        L.oracle_first_touch_numbering.argtypes = [I, I, IP, IP, IP, I, IP, IP, IP]
        L.oracle_node_to_node_csr.argtypes = [I, I, IP, IP, IP, IP]
        L.oracle_bcsr_mult.argtypes = [I, I, IP, IP, DP, DP, DP]
        L.oracle_assemble_jacobian.argtypes = [I, I, I, IP, IP, DP, DP, DP, DP, D, D, D, IP, IP, I, IP, IP, DP, D,
                                               DP, DP]
        _lib = L
    return _lib


def iso_shell_desc(rho=2700.0, E=70e3, nu=0.3, t=0.01, tOffset=0.0, kcorr=5.0 / 6.0, kdrill=0.1, transform=1,
                   axis=(1.0, 0.0, 0.0)):
    L = load()
    d = np.zeros(32)
    Cs, mom = np.zeros(22), np.zeros(3)
    L.oracle_iso_shell_stiffness(rho, E, nu, t, tOffset, kcorr, kdrill, _p(Cs), _p(mom))
    ax = np.asarray(axis, float)
    ax = ax / np.linalg.norm(ax)
    d[:22], d[22:25], d[25], d[26:29] = Cs, mom, transform, ax
    return d


def composite_shell_desc(kdrill=0.1, transform=1, axis=(1.0, 0.0, 0.0)):
    L = load()
    ply = np.tile(np.array([1550.0, 54e3, 18e3, 0.25, 9e3, 9e3, 9e3]), 6)
    th = np.full(6, 1.25e-4)
    ang = np.array([0.0, 45.0, 30.0, 30.0, 45.0, 0.0]) * np.pi / 180.0
    Cs, mom = np.zeros(22), np.zeros(3)
    L.oracle_composite_shell_stiffness(6, _p(ply), _p(th), _p(ang), 5.0 / 6.0, 0.0, kdrill, _p(Cs), _p(mom))
    d = np.zeros(32)
    ax = np.asarray(axis, float)
    ax = ax / np.linalg.norm(ax)
    d[:22], d[22:25], d[25], d[26:29] = Cs, mom, transform, ax
    return d


def solid_desc(rho=2700.0, E=70e3, nu=0.3, t=1.0):
    L = load()
    C21, dens = np.zeros(21), np.zeros(1)
    L.oracle_solid_stiffness(rho, E, nu, t, _p(C21), _p(dens))
    d = np.zeros(32)
    d[:21], d[21] = C21, dens[0]
    return d


def element(kind, desc, Xpts, vars, ddvars=None, alpha=1.0, gamma=0.0):
    """One element through the restatement; returns (res, mat)."""
    L = load()
    shell = kind <= 2
    order = 2 if kind in (1, 3) else 3
    nn = order ** 2 if shell else order ** 3
    nv = nn * (6 if shell else 3)
    X = np.ascontiguousarray(Xpts, float).ravel()
    u = np.ascontiguousarray(vars, float).ravel()
    a = None if ddvars is None else np.ascontiguousarray(ddvars, float).ravel()
    res, mat = np.zeros(nv), np.zeros(nv * nv)
    if shell:
        Cs, mom, ax = desc[:22].copy(), desc[22:25].copy(), desc[26:29].copy()
        L.oracle_shell_element(order, _p(X), _p(u), _p(a), int(desc[25]), _p(ax), _p(Cs), _p(mom), alpha, 0.0, gamma,
                               _p(res), _p(mat))
    else:
        C21 = desc[:21].copy()
        L.oracle_solid_element(order, _p(X), _p(u), _p(a), _p(C21), float(desc[21]), alpha, 0.0, gamma, _p(res),
                               _p(mat))
    return res, mat.reshape(nv, nv)


def first_touch(mesh, partition=None, nparts=1):
    L = load()
    ne = mesh["elem_ids"].size
    part = np.zeros(ne, np.int32) if partition is None else np.ascontiguousarray(partition, np.int32)
    new_nodes = np.zeros(mesh["num_nodes"], np.int32)
    on, oe = np.zeros(nparts, np.int32), np.zeros(nparts, np.int32)
    L.oracle_first_touch_numbering(mesh["num_nodes"], ne, _p(mesh["ptr"]), _p(mesh["conn"]), _p(part), nparts,
                                   _p(new_nodes), _p(on), _p(oe))
    return new_nodes, on, oe


def assemble(mesh, kind, new_nodes=None, desc=None, vars=None, ddvars=None, x=None, alpha=1.0, gamma=0.0,
             lam=1.0):
    """Single-rank assembleJacobian + SpMV through the restatement, in the creator's node numbering."""
    L = load()
    if new_nodes is None:
        new_nodes, _, _ = first_touch(mesh)
    new_nodes = np.ascontiguousarray(new_nodes, np.int32)
    shell = kind <= 2
    bs = 6 if shell else 3
    nnodes, ne = mesh["num_nodes"], mesh["elem_ids"].size
    conn = np.ascontiguousarray(new_nodes[mesh["conn"]], np.int32)
    X = np.zeros((nnodes, 3))
    X[new_nodes] = mesh["Xpts"]
    if desc is None:
        desc = iso_shell_desc() if shell else solid_desc()
    desc = np.ascontiguousarray(np.atleast_2d(desc), float)
    rowp = np.zeros(nnodes + 1, np.int32)
    nnz = L.oracle_node_to_node_csr(nnodes, ne, _p(mesh["ptr"]), _p(conn), _p(rowp), None)
    cols = np.zeros(nnz, np.int32)
    L.oracle_node_to_node_csr(nnodes, ne, _p(mesh["ptr"]), _p(conn), _p(rowp), _p(cols))
    bc_nodes = np.ascontiguousarray(new_nodes[mesh["bc_nodes"]], np.int32)
    bc_vals = None
    if mesh.get("bc_ptr") is None:
        bc_vars = np.full(bc_nodes.size, (1 << bs) - 1, np.int32)
    else:
        bc_vars = np.zeros(bc_nodes.size, np.int32)
        bc_vals = np.zeros((bc_nodes.size, bs))
        for k in range(bc_nodes.size):
            for j in range(mesh["bc_ptr"][k], mesh["bc_ptr"][k + 1]):
                bc_vars[k] |= 1 << int(mesh["bc_vars"][j])
                if mesh.get("bc_vals") is not None:
                    bc_vals[k, int(mesh["bc_vars"][j])] = mesh["bc_vals"][j]
    res, A = np.zeros(bs * nnodes), np.zeros(bs * bs * nnz)
    u = None if vars is None else np.ascontiguousarray(vars, float)
    a = None if ddvars is None else np.ascontiguousarray(ddvars, float)
    edesc = np.ascontiguousarray(mesh["elem_ids"], np.int32)
    rc = L.oracle_assemble_jacobian(kind, nnodes, ne, _p(conn), _p(edesc), _p(desc.ravel()), _p(X.ravel()), _p(u),
                                    _p(a), alpha, 0.0, gamma, _p(rowp), _p(cols), bc_nodes.size, _p(bc_nodes),
                                    _p(bc_vars), _p(bc_vals), lam, _p(res), _p(A))
    assert rc == 0
    out = dict(rowp=rowp, cols=cols, A=A.reshape(nnz, bs, bs), res=res, new_nodes=new_nodes)
    if x is not None:
        xx = np.ascontiguousarray(x, float)
        y = np.zeros(bs * nnodes)
        L.oracle_bcsr_mult(bs, nnodes, _p(rowp), _p(cols), _p(A), _p(xx), _p(y))
        out["y"] = y
    return out
