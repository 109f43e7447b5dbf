"""Worker of tests/test_distributed_plan.py::test_two_ranks_over_gloo (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import dist_emul  # noqa: E402
from tests.test_distributed_plan import build_plans, serial_reference  # noqa: E402


class GlooTransport:
    def __init__(self):
        self.pending = []

    def send(self, src, dst, tag, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64).ravel().copy())
        self.pending.append(dist.isend(t, dst))

    def recv(self, src, dst, tag, shape):
        t = torch.empty(int(np.prod(shape)), dtype=torch.float64)
        dist.recv(t, src)
        return t.numpy().reshape(shape)

    def wait(self):
        for p in self.pending:
            p.wait()
        self.pending = []


def main():
    dist.init_process_group("gloo")
    rank, size = dist.get_rank(), dist.get_world_size()
    import tacs_b200

    lib = tacs_b200.load()
    for name in ("quad4_plate", "hex8_cube"):
        mesh, kind, desc, creator, plans, _ = build_plans(lib, name, size, ranks=[rank])
        new_nodes = creator.getNodeNums()
        u, x, bc, serial = serial_reference(mesh, kind, desc, new_nodes)
        tr = GlooTransport()
        st = dist_emul.run_rank_phase1(plans, rank, tr, mesh, kind, desc, new_nodes, u)
        dist_emul.run_rank_phase2(plans, rank, tr, st, bc, u, x)
        dist_emul.run_rank_phase3(plans, rank, st)
        tr.wait()
        eA, eR, eY = dist_emul.check_against_serial(plans, rank, st, serial)
        assert eA < 1e-12 and eR < 1e-12 and eY < 1e-12, (name, rank, eA, eR, eY)
        # a global reduction like TACSBVec::dot: sum of owned partial dot products == serial dot
        part = torch.tensor([float(st["y"].ravel() @ st["x_owned"].ravel())], dtype=torch.float64)
        dist.all_reduce(part)
        want = float(serial["y"] @ x)
        assert abs(part.item() - want) < 1e-11 * abs(want)
    dist.barrier()
    print(f"rank {rank} ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
