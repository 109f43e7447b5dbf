"""Known answers of the full-size benchmark configurations, from the compiled reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref; minutes of CPU, up to ~45 GB of RAM):
    python tests/golden/make_fullsize_norms.py [c2 c4 c3max c5max ...]
For every configuration the unmodified reference assembles the Jacobian and residual for the deterministic
state of SURVEY 8d (hash vector, BC rows zeroed), multiplies with the reversed hash vector, and the script
records size-independent scalars of the result (norms, a weighted checksum and a handful of sampled entries)
in tests/golden/fullsize_norms.json.  C3 / C5 cannot be assembled by the reference at full size (its
`int` index bs^2*nnzb overflows, BCSRMat.cpp:234), so they are pinned at the largest size that fits.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tacs_b200 import TACS as T  # noqa: E402
from tacs_b200 import binding, meshgen  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fullsize_norms.json")

CONFIGS = {
    # name: (mesh factory, element factory, description)
    "c2": (lambda: meshgen.plate(2, 1000, 1000), lambda lib: meshgen.iso_shell_element(T, lib, 2),
           "1000x1000 Quad4 plate (BASELINE configs[1])"),
    "c4": (lambda: meshgen.cube(2, 200), lambda lib: meshgen.solid_element(T, lib, 2),
           "200^3 hex8 solid (BASELINE configs[3])"),
    "c3max": (lambda: meshgen.cylinder(3, 600, 1200), lambda lib: meshgen.composite_shell_element(T, lib, 3),
              "600x1200 Quad9 composite cylinder (largest C3-like size the reference's int indexing allows)"),
    "c5max": (lambda: meshgen.cube(3, 70), lambda lib: meshgen.solid_element(T, lib, 3),
              "70^3 hex27 solid (largest C5-like size the reference's int indexing allows comfortably)"),
    # small versions for a quick check of the script itself
    "c2small": (lambda: meshgen.plate(2, 100, 100), lambda lib: meshgen.iso_shell_element(T, lib, 2),
                "100x100 Quad4 plate"),
}


def checksum_weights(n):
    """Deterministic weights in [0.5, 1.5) so that a permutation or sign error cannot cancel in the checksum."""
    i = np.arange(n, dtype=np.uint64)
    return 0.5 + (((i * np.uint64(40503)) % np.uint64(65536)).astype(np.float64) / 65536.0)


def summarise(vec):
    n = vec.size
    idx = np.unique(np.linspace(0, n - 1, 16).astype(np.int64))
    return {"norm2": float(np.linalg.norm(vec)), "max": float(np.abs(vec).max()),
            "checksum": float(np.dot(checksum_weights(n), vec)), "sample_idx": idx.tolist(),
            "sample": [float(v) for v in vec[idx]]}


def run(lib, name):
    mesh_f, elem_f, what = CONFIGS[name]
    t0 = time.time()
    mesh = mesh_f()
    creator, asm = meshgen.build_model(T, lib, mesh, [elem_f(lib)])
    A, res, x, y, u = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    n = u.getSize()
    u.setArray(meshgen.hash_vector(n))
    asm.applyBCs(u)
    asm.setVariables(u)
    if lib.prefix == "ref_":
        asm.setNumThreads(min(16, len(os.sched_getaffinity(0))))
    t1 = time.time()
    asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    t2 = time.time()
    x.setArray(meshgen.hash_vector(n)[::-1].copy())
    asm.applyBCs(x)
    A.mult(x, y)
    bs, nrows, ncols, nnzb = A.getSizes()
    out = {"what": what, "elements": int(mesh["elem_ids"].size), "dof": int(n), "nnzb": int(nnzb),
           "res": summarise(res.getArray()), "y": summarise(y.getArray()),
           "seconds": {"setup": t1 - t0, "assemble_jacobian": t2 - t1}}
    return out


if __name__ == "__main__":
    names = sys.argv[1:] or ["c2", "c4", "c3max", "c5max"]
    from tests import ref_binding
    ref = ref_binding.load_reference()
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    for name in names:
        data[name] = run(ref, name)
        print(name, json.dumps(data[name]["y"])[:200], data[name]["seconds"], flush=True)
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
