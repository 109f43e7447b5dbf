"""GPU: the full-size benchmark configurations against known answers of the unmodified reference.

tests/golden/fullsize_norms.json holds, for C2 (1000x1000 Quad4), C4 (200^3 hex8) and the largest C3- / C5-like
sizes the reference's `int` indexing allows (600x1200 Quad9 composite cylinder, 70^3 hex27), the 2-norm, a weighted
checksum and 16 sampled entries of the reference's residual and A*x (tests/golden/make_fullsize_norms.py, run once in
the build container against oracle/_ref). The product library assembles the same models on the GPU through the C ABI.
Tolerance: 1e-12 relative (north_star) on the sampled entries (max-norm scale), the norm and the checksum."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

TOL = 1e-12


@pytest.mark.parametrize("name", ["c2", "c4", "c3max", "c5max"])
def test_fullsize_against_reference_known_answers(lib, name):
    import make_fullsize_norms as M

    with open(os.path.join(HERE, "golden", "fullsize_norms.json")) as f:
        gold = json.load(f)[name]
    got = M.run(lib, name)
    assert got["elements"] == gold["elements"] and got["dof"] == gold["dof"] and got["nnzb"] == gold["nnzb"]
    for key in ("res", "y"):
        g, h = gold[key], got[key]
        assert abs(h["norm2"] - g["norm2"]) <= TOL * g["norm2"], (name, key, "norm2", h["norm2"], g["norm2"])
        assert abs(h["checksum"] - g["checksum"]) <= TOL * g["norm2"] * np.sqrt(gold["dof"]), (name, key, "checksum")
        assert h["sample_idx"] == g["sample_idx"]
        err = np.abs(np.asarray(h["sample"]) - np.asarray(g["sample"])).max() / g["max"]
        assert err <= TOL, (name, key, "samples", err)
