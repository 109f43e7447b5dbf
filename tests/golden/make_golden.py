"""Generate the golden fixtures from the compiled reference (oracle/_ref/libtacs_ref.so).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Each fixture holds, for one small model, the reference's node renumbering, sparsity pattern,
assembled Jacobian, residual and A*x for the deterministic state / input vectors of tests/common.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tacs_b200 import binding  # noqa: E402
from tests import common  # noqa: E402

if __name__ == "__main__":
    from tests import ref_binding
    ref = ref_binding.load_reference()
    out = os.path.dirname(os.path.abspath(__file__))
    for name in sorted(common.SMALL_MODELS):
        r = common.run_model(ref, name)
        np.savez_compressed(os.path.join(out, name + ".npz"), new_nodes=r["new_nodes"], rowp=r["rowp"],
                            cols=r["cols"], A=r["A"], res=r["res"], y=r["y"], x=r["x"], u=r["u"])
        print(name, r["A"].shape, "|y| = %.15e" % np.linalg.norm(r["y"]))
