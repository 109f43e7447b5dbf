"""Run the unmodified reference on N ranks (oracle/_ref/ref_driver over oracle/mpistub/mpi_procs.c) and
load what every rank dumped -- test infrastructure (used by tests/ and by bench.py's CPU-baseline legs)."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

INT_ARRAYS = ["owner_range", "local_to_global", "elem_conn", "Aloc_rowp", "Aloc_cols", "Bext_rowp", "Bext_cols",
              "ext_col_nodes"]
F64_ARRAYS = ["Aloc_vals", "Bext_vals", "res", "y", "u", "x"]


def available():
    return os.path.exists(DRIVER)


def run(mesh, kind, con_kind, nranks, reps=0, timeout=900, load=True):
    """Returns (summary dict, per-rank dict of arrays, root arrays new_nodes/partition)."""
    with tempfile.TemporaryDirectory() as tmp:
        ind, outd = os.path.join(tmp, "in"), os.path.join(tmp, "out")
        os.makedirs(ind)
        os.makedirs(outd)
        ne = mesh["elem_ids"].size
        npe = mesh["conn"].size // ne
        with open(os.path.join(ind, "meta.txt"), "w") as f:
            f.write(f"{mesh['vars_per_node']} {mesh['num_nodes']} {ne} {npe} {mesh['bc_nodes'].size} {kind} {con_kind}\n")
        for name, arr, dt in (("ptr", mesh["ptr"], np.int32), ("conn", mesh["conn"], np.int32),
                              ("ids", mesh["elem_ids"], np.int32), ("bc", mesh["bc_nodes"], np.int32),
                              ("X", mesh["Xpts"], np.float64)):
            np.ascontiguousarray(arr, dtype=dt).tofile(os.path.join(ind, name + ".bin"))
        env = dict(os.environ, TACSB200_MPI_NP=str(nranks))
        proc = subprocess.run([DRIVER, ind, outd, str(reps)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, timeout=timeout)
        if proc.returncode != 0:
            raise RuntimeError(f"ref_driver failed ({proc.returncode}): {proc.stderr[-2000:]}")
        summary = None
        for line in proc.stdout.splitlines():
            if line.startswith("{"):
                summary = json.loads(line)
        ranks = []
        for r in range(nranks if load else 0):
            d = {}
            for name in INT_ARRAYS:
                d[name] = np.fromfile(os.path.join(outd, f"r{r}_{name}.bin"), dtype=np.int32)
            for name in F64_ARRAYS:
                d[name] = np.fromfile(os.path.join(outd, f"r{r}_{name}.bin"), dtype=np.float64)
            ranks.append(d)
        root = None if not load else dict(new_nodes=np.fromfile(os.path.join(outd, "r0_new_nodes.bin"), dtype=np.int32),
                    partition=np.fromfile(os.path.join(outd, "r0_partition.bin"), dtype=np.int32))
    return summary, ranks, root
