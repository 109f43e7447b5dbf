"""ctypes binding of the C ABI declared in include/tacs_b200.h.

`Lib(path, prefix)` resolves `<prefix><name>` for every entry point in `SIGNATURES`.  The product
library is loaded with `load()` (prefix ``tacsb200_``); it raises if libtacs_b200.so has not been
built -- there is no Python or CPU fallback for any compute entry point.  Tests bind the reference
build (oracle/_ref/libtacs_ref.so, prefix ``ref_``) through the same class, so both sides of a
parity test are driven by identical Python code.
"""
import ctypes as C
import os

import numpy as np

H = C.c_void_p
I = C.c_int
D = C.c_double
L = C.c_long
IP = C.POINTER(C.c_int)
DP = C.POINTER(C.c_double)
HP = C.POINTER(C.c_void_p)
UP = C.POINTER(C.c_ubyte)

# name -> (restype, [argtypes]); names are relative to the prefix
SIGNATURES = {
    "abi_version": (I, []),
    "release": (None, [H]),
    "material_properties_create": (H, [D] * 7),
    "material_properties_create_ortho": (H, [D] * 11),
    "orthotropic_ply_create": (H, [D, H]),
    "iso_shell_constitutive_create": (H, [H, D, D, D]),
    "composite_shell_constitutive_create": (H, [I, HP, DP, DP, D, D]),
    "solid_constitutive_create": (H, [H, D]),
    "shell_set_drilling_regularization": (None, [D]),
    "constitutive_eval_tangent_stiffness": (I, [H, DP]),
    "shell_constitutive_eval_mass_moments": (I, [H, DP]),
    "shell_natural_transform_create": (H, []),
    "shell_ref_axis_transform_create": (H, [DP]),
    "quad4_shell_create": (H, [H, H]),
    "quad9_shell_create": (H, [H, H]),
    "linear_hexa_basis_create": (H, []),
    "quadratic_hexa_basis_create": (H, []),
    "linear_elasticity3d_create": (H, [H]),
    "element3d_create": (H, [H, H]),
    "element_num_nodes": (I, [H]),
    "element_vars_per_node": (I, [H]),
    "element_add_jacobian": (I, [H, I, D, D, D, DP, DP, DP, DP, DP, DP]),
    "element_add_residual": (I, [H, I, DP, DP, DP, DP, DP]),
    "creator_create": (H, [I]),
    "comm_rank": (I, []),
    "comm_size": (I, []),
    "creator_set_global_connectivity": (I, [H, I, I, IP, IP, IP]),
    "creator_set_boundary_conditions": (I, [H, I, IP, IP, IP, DP]),
    "creator_set_nodes": (I, [H, DP]),
    "creator_set_elements": (I, [H, I, HP]),
    "creator_partition_mesh": (I, [H, I, IP]),
    "creator_get_node_nums": (I, [H, IP]),
    "creator_get_element_partition": (I, [H, IP]),
    "creator_create_tacs": (H, [H]),
    "assembler_get_vars_per_node": (I, [H]),
    "assembler_get_num_nodes": (I, [H]),
    "assembler_get_num_owned_nodes": (I, [H]),
    "assembler_get_num_elements": (I, [H]),
    "assembler_get_owner_range": (I, [H, IP, IP]),
    "assembler_get_element_connectivity": (I, [H, IP, IP]),
    "assembler_get_local_to_global": (I, [H, IP]),
    "assembler_create_vec": (H, [H]),
    "assembler_create_node_vec": (H, [H]),
    "assembler_create_mat": (H, [H]),
    "assembler_get_nodes": (I, [H, H]),
    "assembler_set_nodes": (I, [H, H]),
    "assembler_set_variables": (I, [H, H, H, H]),
    "assembler_zero_variables": (I, [H]),
    "assembler_apply_bcs_vec": (I, [H, H]),
    "assembler_apply_bcs_mat": (I, [H, H]),
    "assembler_set_bcs": (I, [H, H]),
    "assembler_set_num_threads": (I, [H, I]),
    "assembler_assemble_res": (I, [H, H]),
    "assembler_assemble_jacobian": (I, [H, D, D, D, H, H]),
    "assembler_assemble_jacobian_async": (I, [H, D, D, D, H, H]),
    "assembler_assemble_mat_type": (I, [H, I, H, I]),
    "assembler_add_jacobian_vec_product": (I, [H, D, D, D, D, H, H, I]),
    "vec_get_size": (I, [H]),
    "vec_get_array": (I, [H, DP]),
    "vec_set_array": (I, [H, DP]),
    "vec_norm": (D, [H]),
    "vec_dot": (D, [H, H]),
    "vec_mdot": (I, [H, I, HP, DP]),
    "vec_axpy": (I, [H, D, H]),
    "vec_axpby": (I, [H, D, D, H]),
    "vec_scale": (I, [H, D]),
    "vec_copy_values": (I, [H, H]),
    "vec_zero_entries": (I, [H]),
    "mat_get_sizes": (I, [H, I, IP, IP, IP, IP]),
    "mat_get_pattern": (I, [H, I, IP, IP]),
    "mat_get_values": (I, [H, I, DP]),
    "mat_get_ext_col_nodes": (I, [H, IP]),
    "mat_zero_entries": (I, [H]),
    "mat_mult": (I, [H, H, H]),
    "mat_mult_transpose": (I, [H, H, H]),
    "chebyshev_create": (H, [H, I, D, D, I]),
    "chebyshev_factor": (I, [H]),
    "chebyshev_apply_factor": (I, [H, H, H]),
    "chebyshev_get_spectral_radius": (D, [H]),
    "aux_elements_create": (H, []),
    "aux_elements_add_shell_traction": (I, [H, I, I, DP, I]),
    "aux_elements_add_shell_pressure": (I, [H, I, I, DP, I]),
    "assembler_set_aux_elements": (I, [H, H]),
    "gmres_create": (H, [H, I, I]),
    "gmres_create_pc": (H, [H, H, I, I, I]),
    "gmres_set_tolerances": (I, [H, D, D]),
    "gmres_solve": (I, [H, H, H, I]),
    "gmres_get_iter_count": (I, [H]),
    "gmres_get_residual_norm": (D, [H]),
}

# entry points that exist only in the product library
PRODUCT_ONLY = {
    "init": (I, [I]),
    "comm_unique_id": (I, [UP]),
    "comm_init": (I, [I, I, UP]),
    "synchronize": (I, []),
    "kernel_launches": (L, [I]),
    "vec_device_ptr": (C.c_void_p, [H]),
    "mat_device_values": (C.c_void_p, [H, I]),
    "mat_mult_async": (I, [H, H, H]),
    "mat_create_vec": (H, [H]),
    "time_assemble_jacobian": (D, [H, D, D, D, H, H, I]),
    "time_assemble_res": (D, [H, H, I]),
    "time_mat_mult": (D, [H, H, H, I]),
    "creator_create_plan": (H, [H, I, I]),
    "plan_get_array": (I, [H, C.c_char_p, IP]),
    "assembler_get_plan_stats": (I, [H, C.POINTER(C.c_long)]),
    "profile_enable": (I, [I]),
    "profile_collect": (I, [DP, C.POINTER(C.c_long)]),
    "profile_named": (C.c_char_p, []),
    "assembler_assemble_jacobian_host": (I, [H, D, D, D, DP, DP, H]),
    "mat_copy_values": (I, [H, H]),
    "mat_scale": (I, [H, D]),
    "mat_axpy": (I, [H, D, H]),
    "shell_constitutive_create_raw": (H, [DP, DP]),
    "solid_constitutive_create_raw": (H, [DP, D]),
    "creator_set_keep_numbering": (I, [H, I]),
    "schur_mat_create": (H, [H, I, IP, I, IP, IP, IP, IP, IP, IP, IP, IP, IP]),
    "schur_mat_update": (I, [H]),
    "schur_mat_get_values": (I, [H, I, DP]),
    "schur_mat_mult": (I, [H, H, H]),
    "gmres_set_ortho_type": (I, [H, I]),
    "gmres_set_monitor": (I, [H, C.c_char_p, I]),
    "gmres_set_time_monitor": (I, [H]),
    "measure_fp64_tflops": (D, []),
    "measure_copy_gbs": (D, []),
}

# entry points that exist only in the reference interface (oracle/ref_capi.cpp)
def iptr(a):
    return None if a is None else a.ctypes.data_as(IP)


def dptr(a):
    return None if a is None else a.ctypes.data_as(DP)


def as_i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def as_f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Lib:
    """A shared library that exports `<prefix><name>` for the entry points of SIGNATURES plus `extra` (name ->
    (restype, argtypes)): the product library, or -- in the tests -- any other implementation of the same flat
    interface."""

    def __init__(self, path, prefix, extra=None):
        if not os.path.exists(path):
            raise OSError(
                f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                "tacs_b200 has no CPU fallback")
        self.path = path
        self.prefix = prefix
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        self._cache = {}
        self.extra = dict(extra or {})

    def signatures(self):
        sigs = dict(SIGNATURES)
        sigs.update(self.extra)
        return sigs

    def __getattr__(self, name):
        cache = self.__dict__.get("_cache")
        if cache is None:
            raise AttributeError(name)
        fn = cache.get(name)
        if fn is None:
            sigs = self.signatures()
            if name not in sigs:
                raise AttributeError(f"{self.prefix}{name} is not part of the interface")
            fn = getattr(self.dll, self.prefix + name)
            fn.restype, fn.argtypes = sigs[name][0], list(sigs[name][1])
            cache[name] = fn
        return fn


_PRODUCT = None


def product_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtacs_b200.so")


def load():
    """The product library (libtacs_b200.so next to this file)."""
    global _PRODUCT
    if _PRODUCT is None:
        _PRODUCT = Lib(product_path(), "tacsb200_", PRODUCT_ONLY)
    return _PRODUCT
