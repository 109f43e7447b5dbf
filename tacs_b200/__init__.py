"""tacs_b200: B200-native (sm_100a) assembly + Krylov-operator hot path of TACS.

The compute lives in libtacs_b200.so (hand-written CUDA behind the C ABI of include/tacs_b200.h);
this package is the thin Python mirror of the reference's `tacs.TACS` interface for that path.
"""
from . import binding, TACS, meshgen  # noqa: F401
from .binding import load  # noqa: F401
