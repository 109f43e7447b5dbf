"""CPU: the host-side distributed plans (partition, node maps, Aloc/Bext, gather plans, exchange lists).

In-process: all ranks' plans are built in one process and the data flow is replayed with numpy.
gloo: two real processes (world_size 2) build their own plan and exchange through torch.distributed."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tacs_b200 import TACS as T
from tacs_b200 import meshgen
from tests import dist_emul, oracle_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "quad4_plate": (lambda: meshgen.plate(2, 9, 7), 1, lambda lib: meshgen.iso_shell_element(T, lib, 2),
                    lambda: oracle_port.iso_shell_desc()),
    "hex8_cube": (lambda: meshgen.cube(2, 4), 3, lambda lib: meshgen.solid_element(T, lib, 2),
                  lambda: oracle_port.solid_desc()),
    "quad9_cylinder": (lambda: meshgen.cylinder(3, 4, 6, defect=0.1), 2,
                       lambda lib: meshgen.composite_shell_element(T, lib, 3), lambda: oracle_port.composite_shell_desc()),
}


def build_plans(lib, name, size, ranks=None):
    mesh_f, kind, elem_f, desc_f = CASES[name]
    mesh = mesh_f()
    elem = elem_f(lib)
    creator = T.Creator(lib, mesh["vars_per_node"])
    creator.setGlobalConnectivity(mesh["num_nodes"], mesh["ptr"], mesh["conn"], mesh["elem_ids"])
    creator.setBoundaryConditions(mesh["bc_nodes"])
    creator.setNodes(mesh["Xpts"])
    creator.setElements([elem])
    plans = {r: creator.createPlan(r, size) for r in (range(size) if ranks is None else ranks)}
    return mesh, kind, desc_f(), creator, plans, elem


def serial_reference(mesh, kind, desc, new_nodes):
    bs = 6 if kind <= 2 else 3
    n = bs * mesh["num_nodes"]
    u = meshgen.hash_vector(n)
    x = meshgen.hash_vector(n)[::-1].copy()
    bc = new_nodes[mesh["bc_nodes"]]
    for g in bc:
        u[bs * g:bs * g + bs] = 0.0
        x[bs * g:bs * g + bs] = 0.0
    serial = oracle_port.assemble(mesh, kind, new_nodes=new_nodes, desc=desc, vars=u, x=x)
    return u, x, bc, serial


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("size", [1, 2, 3, 4])
def test_plans_replay_matches_serial(name, size):
    import tacs_b200

    lib = tacs_b200.load()
    mesh, kind, desc, creator, plans, _ = build_plans(lib, name, size)
    new_nodes = creator.getNodeNums()
    part = creator.getElementPartition()
    assert sorted(set(part.tolist())) == list(range(size))
    # METIS partition + first-touch numbering agree with the oracle restatement given the same partition
    nn_oracle, owned_nodes, owned_elems = oracle_port.first_touch(mesh, part, size)
    assert np.array_equal(nn_oracle, new_nodes)
    # local element lists: ascending global ids, a partition of all elements
    all_elems = np.concatenate([plans[r].array("elem_global") for r in range(size)])
    assert np.array_equal(np.sort(all_elems), np.arange(mesh["elem_ids"].size))
    for r in range(size):
        eg = plans[r].array("elem_global")
        assert np.all(np.diff(eg) > 0) and np.all(part[eg] == r)
        s = plans[r].scalars()
        assert s["nowned"] == owned_nodes[r] and s["nelems"] == owned_elems[r]
        lo, hi = plans[r].array("owner_range")[[r, r + 1]]
        conn = plans[r].array("elem_conn_global")
        ext = np.unique(conn[(conn < lo) | (conn >= hi)])
        assert np.array_equal(ext, plans[r].array("ext_nodes"))
    u, x, bc, serial = serial_reference(mesh, kind, desc, new_nodes)
    tr = dist_emul.LocalTransport()
    st = {r: dist_emul.run_rank_phase1(plans, r, tr, mesh, kind, desc, new_nodes, u) for r in range(size)}
    for r in range(size):
        dist_emul.run_rank_phase2(plans, r, tr, st[r], bc, u, x)
    for r in range(size):
        dist_emul.run_rank_phase3(plans, r, st[r])
        eA, eR, eY = dist_emul.check_against_serial(plans, r, st[r], serial)
        assert eA < 1e-12 and eR < 1e-12 and eY < 1e-12, (r, eA, eR, eY)


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("size", [2, 4])
def test_gather_prefix_reads_local_staging_only(name, size):
    """The multi-GPU assembly gathers the blocks below `local_gather_end` while the off-rank staging rows are still in
    flight (tb2_host.cpp: assembleJacobianImpl): none of them may read a received slot, the blocks with a received
    source must all lie behind it, and the prefix has to be most of the plan for the overlap to be worth anything."""
    import tacs_b200

    lib = tacs_b200.load()
    mesh, kind, desc, creator, plans, _ = build_plans(lib, name, size)
    for r in range(size):
        P = plans[r]
        s = P.scalars()
        ptr, src = P.array("gb_ptr"), P.array("gb_src")
        nblk = P.array("gb_blk").size
        assert ptr.size == nblk + 1 and 0 <= s["local_gather_end"] <= nblk
        slots = src[:ptr[-1]] >> 1
        counts = np.diff(ptr)
        last = np.maximum.reduceat(slots, ptr[:-1][counts > 0]) if nblk else np.zeros(0, np.int64)
        last_all = np.full(nblk, -1, np.int64)
        last_all[counts > 0] = last
        end = s["local_gather_end"]
        assert np.all(last_all[:end] < s["local_blocks"]), "a block of the prefix reads a received staging slot"
        # blocks are ordered by the bucket (64 slots) of their last source
        assert np.all(np.diff(last_all >> 6) >= 0)
        if s["recv_blocks"] > 0:
            local_only = int((last_all < s["local_blocks"]).sum())
            assert end >= local_only - 64 * 81, (end, local_only)   # at most one bucket of blocks is held back
        else:
            assert end >= nblk - 64 * 81


def test_single_rank_plan_is_the_serial_pattern():
    import tacs_b200

    lib = tacs_b200.load()
    mesh, kind, desc, creator, plans, _ = build_plans(lib, "hex8_cube", 1)
    o = oracle_port.assemble(mesh, kind, new_nodes=creator.getNodeNums())
    assert np.array_equal(plans[0].array("Aloc_rowp"), o["rowp"]) and np.array_equal(plans[0].array("Aloc_cols"), o["cols"])
    assert plans[0].array("Bext_cols").size == 0 and plans[0].array("ext_nodes").size == 0


def test_two_ranks_over_gloo():
    """world_size 2, gloo backend: each process plans only its own rank and ships rows through torch.distributed."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "gloo_worker.py")]
    proc = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:]
    assert "rank 0 ok" in proc.stdout and "rank 1 ok" in proc.stdout, proc.stdout[-3000:]
