"""Worker of tests/test_gpu_distributed.py (torch.distributed.run, one process per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import tacs_b200  # noqa: E402
from tacs_b200 import TACS as T  # noqa: E402
from tacs_b200 import binding, meshgen  # noqa: E402
from tests import oracle_port  # noqa: E402
from tests.test_distributed_plan import CASES, serial_reference  # noqa: E402


def init_comm(lib):
    rank, size, local = dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", "0"))
    assert lib.init(local) == 0
    torch.cuda.set_device(local)
    buf = np.zeros(128, np.uint8)
    if rank == 0:
        assert lib.comm_unique_id(buf.ctypes.data_as(binding.UP)) == 0
    t = torch.from_numpy(buf).cuda()
    dist.broadcast(t, 0)
    buf = t.cpu().numpy().copy()
    assert lib.comm_init(rank, size, buf.ctypes.data_as(binding.UP)) == 0
    return rank, size


def main():
    dist.init_process_group("nccl")
    lib = tacs_b200.load()
    rank, size = init_comm(lib)
    for name in sorted(CASES):
        mesh_f, kind, elem_f, desc_f = CASES[name]
        mesh = mesh_f()
        bs = mesh["vars_per_node"]
        creator, asm = meshgen.build_model(T, lib, mesh, [elem_f(lib)])
        new_nodes = creator.getNodeNums()
        u, x, bc, serial = serial_reference(mesh, kind, desc_f(), new_nodes)
        lo, hi = asm.getOwnerRange()
        A, res, uv, xv, yv = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
        uv.setArray(u[bs * lo:bs * hi])
        xv.setArray(x[bs * lo:bs * hi])
        asm.setVariables(uv)
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
        A.mult(xv, yv)
        # owned rows of the distributed matrix against the serial assembly in the same numbering
        ar, ac = A.getPattern(0)
        av = A.getValues(0)
        br, bcs = A.getPattern(1)
        bv = A.getValues(1)
        ext_cols = A.getExtColNodes()
        npr = (hi - lo) - (br.size - 1)
        srow, scol, sA = serial["rowp"], serial["cols"], serial["A"]
        scale = np.abs(sA).max()
        worst = 0.0
        for r in range(hi - lo):
            cols = [lo + c for c in ac[ar[r]:ar[r + 1]]]
            vals = [av[k] for k in range(ar[r], ar[r + 1])]
            if r >= npr:
                cols += [ext_cols[c] for c in bcs[br[r - npr]:br[r - npr + 1]]]
                vals += [bv[k] for k in range(br[r - npr], br[r - npr + 1])]
            order = np.argsort(cols)
            g = lo + r
            assert np.array_equal(np.asarray(cols)[order], scol[srow[g]:srow[g + 1]]), (name, "pattern")
            for m, o in enumerate(order):
                worst = max(worst, np.abs(vals[o] - sA[srow[g] + m]).max() / scale)
        assert worst < 1e-12, (name, "A", worst)
        e_res = np.abs(res.getArray() - serial["res"][bs * lo:bs * hi]).max() / np.abs(serial["res"]).max()
        e_y = np.abs(yv.getArray() - serial["y"][bs * lo:bs * hi]).max() / np.abs(serial["y"]).max()
        assert e_res < 1e-12 and e_y < 1e-12, (name, e_res, e_y)
        # residual-only path and global reductions
        res2 = asm.createVec()
        asm.assembleRes(res2)
        assert np.abs(res2.getArray() - res.getArray()).max() <= 1e-12 * np.abs(serial["res"]).max()
        want = float(np.linalg.norm(serial["y"]))
        assert abs(yv.norm() - want) < 1e-12 * want
        assert abs(yv.dot(xv) - float(serial["y"] @ x)) < 1e-11 * abs(float(serial["y"] @ x))
    dist.barrier()
    print(f"rank {rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
