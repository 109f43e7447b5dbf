"""Synthetic structured meshes of the benchmark configurations, generated exactly like the
reference drivers so the reference and tacs_b200 consume identical arrays:

  plate     examples/plate/plate.cpp:29-86          Quad4 / Quad9 flat plate, edges clamped
  cylinder  examples/shell/cylinder.cpp:300-410     Quad4 / Quad9 cylinder, periodic in the hoop direction
  cube      tests/integration_tests/test_elast_linhexa_element_3d.py:76-107   hex8 / hex27 unit cube

Each generator returns a dict with num_nodes, ptr, conn, elem_ids (all zero), Xpts [num_nodes,3],
bc_nodes (+ optional bc_ptr / bc_vars) and vars_per_node.  `build_model` turns one into an
assembler through a bound library (product or reference).
"""
import numpy as np


def plate(order, nx, ny, lx=1.0, ly=1.0):
    nnx, nny = (order - 1) * nx + 1, (order - 1) * ny + 1
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    base = ((order - 1) * i + (order - 1) * j * nnx).ravel()
    ii, jj = np.meshgrid(np.arange(order), np.arange(order), indexing="xy")
    off = (ii + jj * nnx).ravel()
    conn = (base[:, None] + off[None, :]).astype(np.int32)
    ne = nx * ny
    gi, gj = np.meshgrid(np.arange(nnx), np.arange(nny), indexing="xy")
    X = np.zeros((nnx * nny, 3))
    X[:, 0] = (lx * gi / (nnx - 1)).ravel()
    X[:, 1] = (ly * gj / (nny - 1)).ravel()
    bc = []
    k = np.arange(nnx)
    bc.append(np.stack([k, k + nnx * (nny - 1)], 1).ravel())
    k = np.arange(nny)
    bc.append(np.stack([k * nnx, (k + 1) * nnx - 1], 1).ravel())
    return dict(vars_per_node=6, num_nodes=nnx * nny, ptr=(order * order * np.arange(ne + 1)).astype(np.int32),
                conn=conn.ravel(), elem_ids=np.zeros(ne, np.int32), Xpts=X,
                bc_nodes=np.concatenate(bc).astype(np.int32))


def cylinder(order, nx, ny, L=2.0, R=1.0, defect=0.0):
    """nx elements along the axis, ny around the hoop (periodic: the last ring reuses the first)."""
    nnx, nny = (order - 1) * nx + 1, (order - 1) * ny
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    i, j = i.ravel(), j.ravel()
    ii, jj = np.meshgrid(np.arange(order), np.arange(order), indexing="xy")
    ii, jj = ii.ravel(), jj.ravel()
    col = (order - 1) * i[:, None] + ii[None, :]
    row = ((order - 1) * j[:, None] + jj[None, :]) % nny
    conn = (col + row * nnx).astype(np.int32)
    ne = nx * ny
    gi, gj = np.meshgrid(np.arange(nnx), np.arange(nny), indexing="xy")
    u = gi / (nnx - 1)
    v = -np.pi + (2.0 * np.pi * gj) / nny
    theta = v + defect * np.sin(v) * np.cos(2 * np.pi * u)
    x = L * (u + defect * np.cos(v) * np.sin(2 * np.pi * u))
    X = np.stack([x.ravel(), (R * np.cos(theta)).ravel(), (-R * np.sin(theta)).ravel()], 1)
    k = np.arange(nny)
    bc = np.stack([k * nnx, nnx - 1 + k * nnx], 1).ravel()
    return dict(vars_per_node=6, num_nodes=nnx * nny, ptr=(order * order * np.arange(ne + 1)).astype(np.int32),
                conn=conn.ravel(), elem_ids=np.zeros(ne, np.int32), Xpts=X, bc_nodes=bc.astype(np.int32))


def cube(order, n):
    m = (order - 1) * n + 1
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    # element index e = i + n*(j + n*k): i fastest
    i, j, k = i.transpose(2, 1, 0).ravel(), j.transpose(2, 1, 0).ravel(), k.transpose(2, 1, 0).ravel()
    base = (order - 1) * i + m * ((order - 1) * j + m * (order - 1) * k)
    ii, jj, kk = np.meshgrid(np.arange(order), np.arange(order), np.arange(order), indexing="ij")
    ii, jj, kk = ii.transpose(2, 1, 0).ravel(), jj.transpose(2, 1, 0).ravel(), kk.transpose(2, 1, 0).ravel()
    off = ii + m * (jj + m * kk)
    conn = (base[:, None].astype(np.int64) + off[None, :]).astype(np.int32)
    ne, npe = n ** 3, order ** 3
    gi, gj, gk = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    gi, gj, gk = gi.transpose(2, 1, 0).ravel(), gj.transpose(2, 1, 0).ravel(), gk.transpose(2, 1, 0).ravel()
    X = np.stack([gi / (m - 1), gj / (m - 1), gk / (m - 1)], 1).astype(np.float64)
    jj2, kk2 = np.meshgrid(np.arange(m), np.arange(m), indexing="xy")
    bc = (m * (jj2 + m * kk2)).ravel()
    return dict(vars_per_node=3, num_nodes=m ** 3, ptr=(npe * np.arange(ne + 1)).astype(np.int32),
                conn=conn.ravel(), elem_ids=np.zeros(ne, np.int32), Xpts=X, bc_nodes=bc.astype(np.int32))


def hash_vector(n):
    """Deterministic state vector of SURVEY 8d: 1e-3*((i*2654435761 mod 2^32) mod 1000)/1000."""
    i = np.arange(n, dtype=np.uint64)
    return 1e-3 * (((i * np.uint64(2654435761)) % np.uint64(2 ** 32)) % np.uint64(1000)).astype(np.float64) / 1000.0


# ---------------------------------------------------------------------------------- model builders
def iso_shell_element(T, lib, order, t=0.01, transform="refaxis", axis=(1.0, 0.0, 0.0)):
    props = T.MaterialProperties(lib, rho=2700.0, specific_heat=921.096, E=70e3, nu=0.3, ys=270.0, alpha=24e-6,
                                 kappa=230.0)
    con = T.IsoShellConstitutive(lib, props, t=t)
    tr = T.ShellRefAxisTransform(lib, axis) if transform == "refaxis" else T.ShellNaturalTransform(lib)
    return (T.Quad4Shell if order == 2 else T.Quad9Shell)(lib, tr, con)


def composite_shell_element(T, lib, order, axis=(1.0, 0.0, 0.0)):
    """[0/45/30]s laminate, 6 plies x 1.25e-4 (tests/integration_tests/input_files/comp_plate.bdf:58-62)."""
    props = T.MaterialProperties(lib, rho=1550.0, specific_heat=0.0, E1=54e3, E2=18e3, nu12=0.25, G12=9e3, G13=9e3,
                                 G23=9e3)
    ply = T.OrthotropicPly(lib, 1.25e-4, props)
    angles = np.array([0.0, 45.0, 30.0, 30.0, 45.0, 0.0]) * np.pi / 180.0
    con = T.CompositeShellConstitutive(lib, [ply] * 6, np.full(6, 1.25e-4), angles, kcorr=5.0 / 6.0, tOffset=0.0)
    tr = T.ShellRefAxisTransform(lib, axis)
    return (T.Quad4Shell if order == 2 else T.Quad9Shell)(lib, tr, con)


def solid_element(T, lib, order):
    props = T.MaterialProperties(lib, rho=2700.0, specific_heat=921.096, E=70e3, nu=0.3, ys=270.0, alpha=24e-6,
                                 kappa=230.0)
    con = T.SolidConstitutive(lib, props, t=1.0)
    model = T.LinearElasticity3D(lib, con)
    basis = (T.LinearHexaBasis if order == 2 else T.QuadraticHexaBasis)(lib)
    return T.Element3D(lib, model, basis)


def build_model(T, lib, mesh, elements, part=None, split_size=0):
    """Creator -> Assembler for a generated mesh; `elements` is the list indexed by elem_ids."""
    creator = T.Creator(lib, mesh["vars_per_node"])
    creator.setGlobalConnectivity(mesh["num_nodes"], mesh["ptr"], mesh["conn"], mesh["elem_ids"])
    creator.setBoundaryConditions(mesh["bc_nodes"], mesh.get("bc_ptr"), mesh.get("bc_vars"), mesh.get("bc_vals"))
    creator.setNodes(mesh["Xpts"])
    creator.setElements(elements)
    if part is not None or split_size:
        creator.partitionMesh(split_size, part)
    assembler = creator.createTACS()
    return creator, assembler
