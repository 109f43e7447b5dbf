"""CPU: the product's N-rank integer pipeline against the unmodified reference run on N ranks.

The reference runs as N forked ranks (oracle/_ref/ref_driver); METIS is the same binary on both sides.
Bit-exact: element partition, node renumbering, owner ranges, local node order, local element
connectivity, Aloc / Bext rowp+cols, np, external column nodes. Values of the reference's distributed
A, residual and A*x are compared with the serial oracle in the same numbering (1e-12)."""
import numpy as np
import pytest

from oracle import ref_mpi
from tests import oracle_port
from tests.test_distributed_plan import CASES, build_plans

CON_KIND = {"quad4_plate": 0, "hex8_cube": 2, "quad9_cylinder": 1}

pytestmark = pytest.mark.skipif(not ref_mpi.available(), reason="oracle/_ref/ref_driver not built")


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("size", [2, 4, 8])
def test_maps_match_multirank_reference(name, size):
    import tacs_b200

    lib = tacs_b200.load()
    mesh, kind, desc, creator, plans, _ = build_plans(lib, name, size)
    summary, ranks, root = ref_mpi.run(mesh, kind, CON_KIND[name], size)
    assert np.array_equal(root["partition"], creator.getElementPartition())   # same METIS binary, same graph
    assert np.array_equal(root["new_nodes"], creator.getNodeNums())           # first-touch renumbering
    bs = mesh["vars_per_node"]
    for r in range(size):
        P, R = plans[r], ranks[r]
        s = P.scalars()
        assert np.array_equal(R["owner_range"], P.array("owner_range"))
        l2g = np.array([*P.array("ext_nodes")[:s["ext_before"]],
                        *range(P.array("owner_range")[r], P.array("owner_range")[r + 1]),
                        *P.array("ext_nodes")[s["ext_before"]:]], dtype=np.int32)
        assert np.array_equal(R["local_to_global"], l2g)                       # [ext< | owned | ext>=]
        assert np.array_equal(R["elem_conn"], P.array("elem_conn_global"))     # local element order + numbering
        assert np.array_equal(R["Aloc_rowp"], P.array("Aloc_rowp"))
        assert np.array_equal(R["Aloc_cols"], P.array("Aloc_cols"))
        assert np.array_equal(R["Bext_rowp"], P.array("Bext_rowp"))
        assert np.array_equal(R["Bext_cols"], P.array("Bext_cols"))
        assert R["Bext_rowp"].size - 1 == s["nowned"] - s["np"]                # np
        assert np.array_equal(R["ext_col_nodes"], P.array("ext_col_nodes"))
    # the reference's distributed values against the serial oracle (ties the N-rank reference to the
    # serial assembly every other test is pinned to)
    new_nodes = root["new_nodes"]
    n = bs * mesh["num_nodes"]
    u = np.concatenate([R["u"] for R in ranks])
    x = np.concatenate([R["x"] for R in ranks])
    assert u.size == n
    serial = oracle_port.assemble(mesh, kind, new_nodes=new_nodes, desc=desc, vars=u, x=x)
    res = np.concatenate([R["res"] for R in ranks])
    y = np.concatenate([R["y"] for R in ranks])
    assert np.abs(res - serial["res"]).max() <= 1e-12 * np.abs(serial["res"]).max()
    assert np.abs(y - serial["y"]).max() <= 1e-12 * np.abs(serial["y"]).max()
    assert abs(summary["ynorm"] - np.linalg.norm(serial["y"])) <= 1e-12 * np.linalg.norm(serial["y"])
