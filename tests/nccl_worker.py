"""Worker of tests/test_gpu_distributed.py (torch.distributed.run, one process per GPU, NCCL)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import tacs_b200  # noqa: E402
from tacs_b200 import binding  # noqa: E402
from tests import dist_check  # noqa: E402


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
    errs = dist_check.check_all(lib)
    assert errs["pattern_exact"], errs
    for key in ("A", "res", "y", "res_only", "norm"):
        assert errs["max"][key] < 1e-12, (key, errs)
    assert errs["max"]["dot"] < 1e-11, errs
    assert errs["gmres_converged"] and errs["max"]["gmres"] < 1e-10, errs  # north_star: displacements within 1e-10
    dist.barrier()
    print(f"rank {rank} ok {json.dumps(errs['max'])}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
