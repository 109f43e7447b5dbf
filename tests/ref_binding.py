"""Test infrastructure: the compiled reference (oracle/_ref/libtacs_ref.so, interface oracle/ref_capi.cpp) behind the
same ctypes class as the product library, so that one piece of Python drives both sides of a parity test."""
import ctypes as C
import os

from tacs_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libtacs_ref.so")
# entry points that exist only in the reference's flat interface
REFERENCE_ONLY = {
    "wtime": (C.c_double, []),
}


def load_reference(path=REF_SO):
    return binding.Lib(path, "ref_", REFERENCE_ONLY)
