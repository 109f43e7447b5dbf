#!/usr/bin/env python
"""Strong-scaling run of the n^3 hex8 solid (BASELINE configs[3]: 200^3, 24M dof) on WORLD_SIZE GPUs.

    python -m torch.distributed.run --nproc-per-node N scripts/scale_hex.py --edge 200 [--part metis|blocks]

Prints one JSON line (rank 0): aggregate elements/s of assembleJacobian(1,0,0,res,A), SpMV time and
aggregate GB/s, and the set-up times. Device timing with CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def block_partition(n, parts):
    """parts = px*py*pz structured blocks of the n^3 element grid (for comparison with METIS)."""
    p = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[parts]
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    i, j, k = i.transpose(2, 1, 0).ravel(), j.transpose(2, 1, 0).ravel(), k.transpose(2, 1, 0).ravel()
    bi, bj, bk = (i * p[0]) // n, (j * p[1]) // n, (k * p[2]) // n
    return (bi + p[0] * (bj + p[1] * bk)).astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=200)
    ap.add_argument("--order", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--part", default="metis")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import tacs_b200
    from tacs_b200 import TACS as T
    from tacs_b200 import binding, meshgen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    lib = tacs_b200.load()
    assert lib.init(local) == 0
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = np.zeros(128, np.uint8)
        if rank == 0:
            assert lib.comm_unique_id(buf.ctypes.data_as(binding.UP)) == 0
        t = torch.from_numpy(buf).cuda()
        dist.broadcast(t, 0)
        buf = t.cpu().numpy().copy()
        assert lib.comm_init(rank, world, buf.ctypes.data_as(binding.UP)) == 0
    t0 = time.time()
    mesh = meshgen.cube(args.order, args.edge)
    t1 = time.time()
    part = block_partition(args.edge, world) if (args.part == "blocks" and world > 1) else None
    creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.solid_element(T, lib, args.order)], part=part,
                                       split_size=world if part is not None else 0)
    t2 = time.time()
    A = asm.createMat()
    t3 = time.time()
    res, x, y = asm.createVec(), asm.createVec(), asm.createVec()
    lo, hi = asm.getOwnerRange()
    x.setArray(meshgen.hash_vector(3 * creator.num_nodes)[3 * lo:3 * hi])
    asm.applyBCs(x)
    asm.setVariables(x)

    def maxr(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sumr(v):
        if world == 1:
            return v
        tt = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return float(tt.item())

    for _ in range(3):
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    lib.synchronize()
    if world > 1:
        dist.barrier()
    ms = maxr(lib.time_assemble_jacobian(asm.h, 1.0, 0.0, 0.0, res.h, A.h, args.steps)) / args.steps
    lib.time_mat_mult(A.h, x.h, y.h, 3)
    if world > 1:
        dist.barrier()
    nsp = 30
    ms_sp = maxr(lib.time_mat_mult(A.h, x.h, y.h, nsp)) / nsp
    bsA = A.getSizes(0)
    bsB = A.getSizes(1)
    bytes_local = (bsA[3] + bsB[3]) * (8 * 9 + 4) + 4 * (bsA[1] + 1) + 16 * 3 * bsA[1]
    bytes_total = sumr(float(bytes_local))
    ne_total = args.edge ** 3
    ynorm = y.norm()
    s = asm.getNumElements()
    out = dict(workload=f"{args.edge}^3 hex{8 if args.order == 2 else 27} solid", n_gpus=world, partition=args.part if world > 1 else "single",
               elements=ne_total, local_elements_rank0=s, jac_ms=ms, elements_per_s=ne_total / ms * 1e3,
               spmv_ms=ms_sp, spmv_gbs_aggregate=bytes_total / ms_sp * 1e-6, ynorm=ynorm,
               setup_s=dict(mesh=t1 - t0, create_tacs=t2 - t1, create_mat=t3 - t2),
               nnzb_rank0=[bsA[3], bsB[3]], ext_cols_rank0=bsB[2])
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
