"""Python mirror of the reference's `tacs.TACS` / `tacs.elements` / `tacs.constitutive` classes
for the assembly + Krylov-operator path (tacs/TACS.pyx:594-845 Vec, :960-1100 Mat, :1256 KSM,
:1595+ Assembler, Creator; tacs/elements.pyx; tacs/constitutive.pyx), bound to a C ABI library.

Every class takes the bound library (`binding.Lib`) so the same code drives libtacs_b200.so and,
in tests, the reference build.  Method names and argument meaning follow the Cython layer.
"""
import ctypes as C

import numpy as np

from . import binding as B


class _Obj:
    def __init__(self, lib, handle, what):
        if not handle:
            raise RuntimeError(f"{lib.prefix}{what} failed (see stderr)")
        self.lib = lib
        self.h = handle

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.release(self.h)
                self.h = None
        except Exception:
            pass


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc} (see stderr)")


# ----------------------------------------------------------------------------- constitutive
class MaterialProperties(_Obj):
    def __init__(self, lib, rho=2700.0, specific_heat=921.096, E=70e3, nu=0.3, ys=270.0, alpha=24e-6,
                 kappa=230.0, E1=None, E2=None, E3=None, nu12=None, nu13=None, nu23=None, G12=None, G13=None,
                 G23=None):
        if E1 is not None:
            E3 = E2 if E3 is None else E3
            nu13 = nu12 if nu13 is None else nu13
            nu23 = nu12 if nu23 is None else nu23
            h = lib.material_properties_create_ortho(rho, specific_heat, E1, E2, E3, nu12, nu13, nu23, G12, G13,
                                                     G23)
        else:
            h = lib.material_properties_create(rho, specific_heat, E, nu, ys, alpha, kappa)
        super().__init__(lib, h, "material_properties_create")


class OrthotropicPly(_Obj):
    def __init__(self, lib, ply_thickness, props):
        self.props = props
        super().__init__(lib, lib.orthotropic_ply_create(ply_thickness, props.h), "orthotropic_ply_create")


class _Constitutive(_Obj):
    nstiff = 22

    def evalTangentStiffness(self):
        out = np.zeros(self.nstiff)
        _check(self.lib.constitutive_eval_tangent_stiffness(self.h, B.dptr(out)), "evalTangentStiffness")
        return out


class IsoShellConstitutive(_Constitutive):
    def __init__(self, lib, props, t=1.0, tOffset=0.0, kcorr=5.0 / 6.0):
        self.props = props
        super().__init__(lib, lib.iso_shell_constitutive_create(props.h, t, tOffset, kcorr),
                         "iso_shell_constitutive_create")

    def evalMassMoments(self):
        out = np.zeros(3)
        _check(self.lib.shell_constitutive_eval_mass_moments(self.h, B.dptr(out)), "evalMassMoments")
        return out


class CompositeShellConstitutive(IsoShellConstitutive):
    def __init__(self, lib, plies, thickness, angles, kcorr=5.0 / 6.0, tOffset=0.0):
        self.plies = list(plies)
        arr = (C.c_void_p * len(plies))(*[p.h for p in plies])
        th, an = B.as_f64(thickness), B.as_f64(angles)
        h = lib.composite_shell_constitutive_create(len(plies), arr, B.dptr(th), B.dptr(an), kcorr, tOffset)
        _Obj.__init__(self, lib, h, "composite_shell_constitutive_create")


class SolidConstitutive(_Constitutive):
    nstiff = 21

    def __init__(self, lib, props, t=1.0):
        self.props = props
        super().__init__(lib, lib.solid_constitutive_create(props.h, t), "solid_constitutive_create")


# ----------------------------------------------------------------------------- elements
class ShellNaturalTransform(_Obj):
    def __init__(self, lib):
        super().__init__(lib, lib.shell_natural_transform_create(), "shell_natural_transform_create")


class ShellRefAxisTransform(_Obj):
    def __init__(self, lib, axis):
        ax = B.as_f64(axis)
        super().__init__(lib, lib.shell_ref_axis_transform_create(B.dptr(ax)), "shell_ref_axis_transform_create")


class _Element(_Obj):
    def getNumNodes(self):
        return self.lib.element_num_nodes(self.h)

    def getVarsPerNode(self):
        return self.lib.element_vars_per_node(self.h)

    def addJacobian(self, alpha, beta, gamma, Xpts, vars, dvars=None, ddvars=None):
        """Batch form of TACSElement::addJacobian: arrays are [count, ...]; returns (res, mat)."""
        nn, nv = self.getNumNodes(), self.getNumNodes() * self.getVarsPerNode()
        X = B.as_f64(Xpts).reshape(-1, 3 * nn)
        count = X.shape[0]
        u = B.as_f64(vars).reshape(count, nv)
        dv = None if dvars is None else B.as_f64(dvars).reshape(count, nv)
        dd = None if ddvars is None else B.as_f64(ddvars).reshape(count, nv)
        res, mat = np.zeros((count, nv)), np.zeros((count, nv, nv))
        _check(self.lib.element_add_jacobian(self.h, count, alpha, beta, gamma, B.dptr(X), B.dptr(u), B.dptr(dv),
                                             B.dptr(dd), B.dptr(res), B.dptr(mat)), "addJacobian")
        return res, mat

    def addResidual(self, Xpts, vars, dvars=None, ddvars=None):
        nn, nv = self.getNumNodes(), self.getNumNodes() * self.getVarsPerNode()
        X = B.as_f64(Xpts).reshape(-1, 3 * nn)
        count = X.shape[0]
        u = B.as_f64(vars).reshape(count, nv)
        dv = None if dvars is None else B.as_f64(dvars).reshape(count, nv)
        dd = None if ddvars is None else B.as_f64(ddvars).reshape(count, nv)
        res = np.zeros((count, nv))
        _check(self.lib.element_add_residual(self.h, count, B.dptr(X), B.dptr(u), B.dptr(dv), B.dptr(dd),
                                             B.dptr(res)), "addResidual")
        return res


class Quad4Shell(_Element):
    def __init__(self, lib, transform, con):
        self.transform, self.con = transform, con
        super().__init__(lib, lib.quad4_shell_create(transform.h, con.h), "quad4_shell_create")


class Quad9Shell(_Element):
    def __init__(self, lib, transform, con):
        self.transform, self.con = transform, con
        super().__init__(lib, lib.quad9_shell_create(transform.h, con.h), "quad9_shell_create")


class LinearHexaBasis(_Obj):
    def __init__(self, lib):
        super().__init__(lib, lib.linear_hexa_basis_create(), "linear_hexa_basis_create")


class QuadraticHexaBasis(_Obj):
    def __init__(self, lib):
        super().__init__(lib, lib.quadratic_hexa_basis_create(), "quadratic_hexa_basis_create")


class LinearElasticity3D(_Obj):
    def __init__(self, lib, con):
        self.con = con
        super().__init__(lib, lib.linear_elasticity3d_create(con.h), "linear_elasticity3d_create")


class Element3D(_Element):
    def __init__(self, lib, model, basis):
        self.model, self.basis = model, basis
        super().__init__(lib, lib.element3d_create(model.h, basis.h), "element3d_create")


# ----------------------------------------------------------------------------- vectors / matrices
class Vec(_Obj):
    def __init__(self, lib, handle):
        super().__init__(lib, handle, "create_vec")

    def getSize(self):
        return self.lib.vec_get_size(self.h)

    def getArray(self):
        """Copy of the owned entries (the reference returns a view of host memory)."""
        out = np.zeros(self.getSize())
        _check(self.lib.vec_get_array(self.h, B.dptr(out)), "vec_get_array")
        return out

    def setArray(self, values):
        v = B.as_f64(values).ravel()
        assert v.size == self.getSize()
        _check(self.lib.vec_set_array(self.h, B.dptr(v)), "vec_set_array")

    def norm(self):
        return self.lib.vec_norm(self.h)

    def dot(self, other):
        return self.lib.vec_dot(self.h, other.h)

    def mdot(self, others):
        arr = (C.c_void_p * len(others))(*[o.h for o in others])
        out = np.zeros(len(others))
        _check(self.lib.vec_mdot(self.h, len(others), arr, B.dptr(out)), "vec_mdot")
        return out

    def axpy(self, alpha, x):
        _check(self.lib.vec_axpy(self.h, alpha, x.h), "vec_axpy")

    def axpby(self, alpha, beta, x):
        _check(self.lib.vec_axpby(self.h, alpha, beta, x.h), "vec_axpby")

    def scale(self, alpha):
        _check(self.lib.vec_scale(self.h, alpha), "vec_scale")

    def copyValues(self, x):
        _check(self.lib.vec_copy_values(self.h, x.h), "vec_copy_values")

    def zeroEntries(self):
        _check(self.lib.vec_zero_entries(self.h), "vec_zero_entries")


class Mat(_Obj):
    def __init__(self, lib, handle):
        super().__init__(lib, handle, "create_mat")

    def getSizes(self, which=0):
        v = [C.c_int() for _ in range(4)]
        _check(self.lib.mat_get_sizes(self.h, which, *[C.byref(x) for x in v]), "mat_get_sizes")
        return tuple(x.value for x in v)  # bsize, nrows, ncols, nnzb

    def getPattern(self, which=0):
        bs, nrows, ncols, nnzb = self.getSizes(which)
        rowp, cols = np.zeros(nrows + 1, np.int32), np.zeros(max(nnzb, 1), np.int32)
        _check(self.lib.mat_get_pattern(self.h, which, B.iptr(rowp), B.iptr(cols)), "mat_get_pattern")
        return rowp, cols[:nnzb]

    def getValues(self, which=0):
        bs, nrows, ncols, nnzb = self.getSizes(which)
        out = np.zeros((max(nnzb, 1), bs, bs))
        if nnzb:
            _check(self.lib.mat_get_values(self.h, which, B.dptr(out)), "mat_get_values")
        return out[:nnzb]

    def getExtColNodes(self):
        n = self.lib.mat_get_ext_col_nodes(self.h, None)
        out = np.zeros(max(n, 1), np.int32)
        self.lib.mat_get_ext_col_nodes(self.h, B.iptr(out))
        return out[:n]

    def zeroEntries(self):
        _check(self.lib.mat_zero_entries(self.h), "mat_zero_entries")

    def mult(self, x, y):
        _check(self.lib.mat_mult(self.h, x.h, y.h), "mat_mult")

    def multTranspose(self, x, y):
        _check(self.lib.mat_mult_transpose(self.h, x.h, y.h), "mat_mult_transpose")

    def createVec(self):
        return Vec(self.lib, self.lib.mat_create_vec(self.h))


class ChebyshevSmoother(_Obj):
    """TACSChebyshevSmoother (src/bpmat/TACSParallelMat.h:180): polynomial smoother used as a preconditioner."""

    def __init__(self, lib, mat, degree, lower_factor=1.0 / 30.0, upper_factor=1.1, iters=1):
        self.mat = mat
        super().__init__(lib, lib.chebyshev_create(mat.h, degree, lower_factor, upper_factor, iters),
                         "chebyshev_create")

    def factor(self):
        _check(self.lib.chebyshev_factor(self.h), "factor")

    def applyFactor(self, x, y):
        _check(self.lib.chebyshev_apply_factor(self.h, x.h, y.h), "applyFactor")

    def getSpectralRadius(self):
        return self.lib.chebyshev_get_spectral_radius(self.h)


class KSM(_Obj):
    """GMRES (tacs/TACS.pyx:1256): KSM(mat, pc, m, nrestart, isFlexible)."""

    def __init__(self, lib, mat, m, nrestart=0, pc=None, isFlexible=0):
        self.mat, self.pc = mat, pc
        if pc is None:
            super().__init__(lib, lib.gmres_create(mat.h, m, nrestart), "gmres_create")
        else:
            super().__init__(lib, lib.gmres_create_pc(mat.h, pc.h, m, nrestart, isFlexible), "gmres_create_pc")

    def setTolerances(self, rtol, atol):
        _check(self.lib.gmres_set_tolerances(self.h, rtol, atol), "gmres_set_tolerances")

    def solve(self, b, x, zero_guess=1):
        return self.lib.gmres_solve(self.h, b.h, x.h, zero_guess)

    def setOrthoType(self, classical):
        """GMRES::setOrthoType: classical Gram-Schmidt (True) or modified (False, the default)."""
        _check(self.lib.gmres_set_ortho_type(self.h, 1 if classical else 0), "gmres_set_ortho_type")

    def setMonitor(self, descript="GMRES", freq=1):
        """KSMPrintStdout(descript, rank, freq) attached with setMonitor."""
        _check(self.lib.gmres_set_monitor(self.h, descript.encode(), freq), "gmres_set_monitor")

    def setTimeMonitor(self):
        _check(self.lib.gmres_set_time_monitor(self.h), "gmres_set_time_monitor")

    def getIterCount(self):
        return self.lib.gmres_get_iter_count(self.h)

    def getResidualNorm(self):
        return self.lib.gmres_get_residual_norm(self.h)


# ----------------------------------------------------------------------------- creator / assembler
class Assembler(_Obj):
    def __init__(self, lib, handle):
        super().__init__(lib, handle, "creator_create_tacs")

    def getVarsPerNode(self):
        return self.lib.assembler_get_vars_per_node(self.h)

    def getNumNodes(self):
        return self.lib.assembler_get_num_nodes(self.h)

    def getNumOwnedNodes(self):
        return self.lib.assembler_get_num_owned_nodes(self.h)

    def getNumElements(self):
        return self.lib.assembler_get_num_elements(self.h)

    def getOwnerRange(self):
        lo, hi = C.c_int(), C.c_int()
        _check(self.lib.assembler_get_owner_range(self.h, C.byref(lo), C.byref(hi)), "get_owner_range")
        return lo.value, hi.value

    def getElementConnectivity(self):
        ne = self.getNumElements()
        n = self.lib.assembler_get_element_connectivity(self.h, None, None)
        ptr, conn = np.zeros(ne + 1, np.int32), np.zeros(max(n, 1), np.int32)
        self.lib.assembler_get_element_connectivity(self.h, B.iptr(ptr), B.iptr(conn))
        return ptr, conn[:n]

    def getLocalToGlobal(self):
        out = np.zeros(max(self.getNumNodes(), 1), np.int32)
        n = self.lib.assembler_get_local_to_global(self.h, B.iptr(out))
        return out[:n]

    def createVec(self):
        return Vec(self.lib, self.lib.assembler_create_vec(self.h))

    def createNodeVec(self):
        return Vec(self.lib, self.lib.assembler_create_node_vec(self.h))

    def createMat(self):
        return Mat(self.lib, self.lib.assembler_create_mat(self.h))

    def getNodes(self, X):
        _check(self.lib.assembler_get_nodes(self.h, X.h), "getNodes")

    def setNodes(self, X):
        _check(self.lib.assembler_set_nodes(self.h, X.h), "setNodes")

    def setVariables(self, q, qdot=None, qddot=None):
        _check(self.lib.assembler_set_variables(self.h, q.h, qdot.h if qdot else None,
                                                qddot.h if qddot else None), "setVariables")

    def zeroVariables(self):
        _check(self.lib.assembler_zero_variables(self.h), "zeroVariables")

    def setAuxElements(self, aux):
        self._aux = aux
        _check(self.lib.assembler_set_aux_elements(self.h, aux.h if aux is not None else None), "setAuxElements")

    def applyBCs(self, vec):
        _check(self.lib.assembler_apply_bcs_vec(self.h, vec.h), "applyBCs")

    def applyMatBCs(self, mat):
        _check(self.lib.assembler_apply_bcs_mat(self.h, mat.h), "applyMatBCs")

    def setBCs(self, vec):
        _check(self.lib.assembler_set_bcs(self.h, vec.h), "setBCs")

    def setNumThreads(self, t):
        _check(self.lib.assembler_set_num_threads(self.h, t), "setNumThreads")

    def assembleRes(self, res):
        _check(self.lib.assembler_assemble_res(self.h, res.h), "assembleRes")

    def assembleJacobian(self, alpha, beta, gamma, res, mat, wait=True):
        """wait=False enqueues only (tacsb200_assembler_assemble_jacobian_async): a following res.getArray() copies
        the residual to the host while the matrix is still being gathered."""
        f = self.lib.assembler_assemble_jacobian if wait else self.lib.assembler_assemble_jacobian_async
        _check(f(self.h, alpha, beta, gamma, res.h if res else None, mat.h), "assembleJacobian")


    def assembleJacobianHost(self, alpha, beta, gamma, q_host, res_host, mat):
        """setVariables(q) + assembleJacobian with the state taken from / the residual returned to host arrays
        (float64, C-contiguous, ideally pinned): transfers pipelined against the kernels. The matrix gather may still
        be running on return (lib.synchronize())."""
        assert q_host.dtype == np.float64 and res_host.dtype == np.float64
        _check(self.lib.assembler_assemble_jacobian_host(self.h, alpha, beta, gamma, B.dptr(q_host), B.dptr(res_host),
                                                         mat.h), "assembleJacobianHost")

    def assembleMatType(self, matType, mat, applyBCs=True):
        """matType: STIFFNESS_MATRIX (1) or MASS_MATRIX (2) (tacs/TACS.pyx ElementMatrixType)."""
        _check(self.lib.assembler_assemble_mat_type(self.h, int(matType), mat.h, 1 if applyBCs else 0),
               "assembleMatType")

    def addJacobianVecProduct(self, scale, alpha, beta, gamma, x, y, applyBCs=True):
        _check(self.lib.assembler_add_jacobian_vec_product(self.h, scale, alpha, beta, gamma, x.h, y.h,
                                                           1 if applyBCs else 0), "addJacobianVecProduct")


JACOBIAN_MATRIX, STIFFNESS_MATRIX, MASS_MATRIX, GEOMETRIC_STIFFNESS_MATRIX = 0, 1, 2, 3


class AuxElements(_Obj):
    """tacs.TACS.AuxElements restricted to the shell load elements of the path: addElement(num, ShellTraction(...)) /
    addElement(num, ShellPressure(...)) become addShellTraction / addShellPressure."""

    def __init__(self, lib):
        super().__init__(lib, lib.aux_elements_create(), "aux_elements_create")

    def addShellTraction(self, num, order, t):
        t = B.as_f64(t).ravel()
        _check(self.lib.aux_elements_add_shell_traction(self.h, int(num), order, B.dptr(t), 1 if t.size == 3 else 0),
               "addShellTraction")

    def addShellPressure(self, num, order, p):
        p = B.as_f64(np.atleast_1d(p)).ravel()
        _check(self.lib.aux_elements_add_shell_pressure(self.h, int(num), order, B.dptr(p), 1 if p.size == 1 else 0),
               "addShellPressure")


class Creator(_Obj):
    def __init__(self, lib, vars_per_node):
        super().__init__(lib, lib.creator_create(vars_per_node), "creator_create")
        self._elements = []
        self.num_nodes = 0
        self.num_elements = 0

    def setGlobalConnectivity(self, num_nodes, ptr, conn, elem_ids):
        ptr, conn, ids = B.as_i32(ptr), B.as_i32(conn), B.as_i32(elem_ids)
        self.num_nodes, self.num_elements = num_nodes, ids.size
        _check(self.lib.creator_set_global_connectivity(self.h, num_nodes, ids.size, B.iptr(ptr), B.iptr(conn),
                                                        B.iptr(ids)), "setGlobalConnectivity")

    def setBoundaryConditions(self, nodes, ptr=None, bcvars=None, bcvals=None):
        nodes, ptr, bcvars, bcvals = B.as_i32(nodes), B.as_i32(ptr), B.as_i32(bcvars), B.as_f64(bcvals)
        _check(self.lib.creator_set_boundary_conditions(self.h, nodes.size, B.iptr(nodes), B.iptr(ptr),
                                                        B.iptr(bcvars), B.dptr(bcvals)), "setBoundaryConditions")

    def setNodes(self, Xpts):
        X = B.as_f64(Xpts).ravel()
        _check(self.lib.creator_set_nodes(self.h, B.dptr(X)), "setNodes")

    def setElements(self, elements):
        self._elements = list(elements)
        arr = (C.c_void_p * len(elements))(*[e.h for e in elements])
        _check(self.lib.creator_set_elements(self.h, len(elements), arr), "setElements")

    def partitionMesh(self, split_size=0, part=None):
        part = B.as_i32(part)
        _check(self.lib.creator_partition_mesh(self.h, split_size, B.iptr(part)), "partitionMesh")

    def getNodeNums(self):
        out = np.zeros(max(self.num_nodes, 1), np.int32)
        n = self.lib.creator_get_node_nums(self.h, B.iptr(out))
        return out[:n]

    def getElementPartition(self):
        out = np.zeros(max(self.num_elements, 1), np.int32)
        n = self.lib.creator_get_element_partition(self.h, B.iptr(out))
        return out[:n]

    def createTACS(self):
        return Assembler(self.lib, self.lib.creator_create_tacs(self.h))

    def createPlan(self, rank, size):
        """Host-only plan of `rank` out of `size` ranks (product library only; needs no GPU)."""
        return Plan(self.lib, self.lib.creator_create_plan(self.h, rank, size))


class Plan(_Obj):
    def __init__(self, lib, handle):
        super().__init__(lib, handle, "creator_create_plan")

    def array(self, name):
        n = self.lib.plan_get_array(self.h, name.encode(), None)
        if n < 0:
            raise KeyError(name)
        out = np.zeros(max(n, 1), np.int32)
        self.lib.plan_get_array(self.h, name.encode(), B.iptr(out))
        return out[:n]

    def scalars(self):
        names = ["nelems", "nowned", "nlocal", "ext_before", "ext_after", "np", "local_blocks", "recv_blocks",
                 "local_node_slots", "recv_node_slots", "direct_blocks", "local_gather_end"]
        return dict(zip(names, (int(v) for v in self.array("scalars"))))
