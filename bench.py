#!/usr/bin/env python
"""Benchmark of the TACS assembly + Krylov-operator hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # tacs_b200 (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path

Headline workload (BASELINE.json configs[1]): synthetic 1000x1000 Quad4 MITC shell plate, 6 dof/node
(~6M dof), isotropic, all edges clamped.  At N > 1 GPUs the plate grows to (1000*N) x 1000 elements
(weak scaling: 1M elements per GPU), partitioned by the reference's METIS call.

A step is one `assembleJacobian(1, 0, 0, res, A)`: element residuals + tangents, atomic-free gather into the
BCSR matrix and the residual, boundary conditions.  `value` is elements/s with everything resident in HBM
(CUDA events on the library's stream); `e2e` repeats the step through the C ABI with the state vector
arriving from pinned host memory and the residual going back to it.  The BCSR SpMV (the other half of the
metric) is timed in its own loop and reported under `spmv`.

Besides the headline the line carries
  parity   (N > 1) the distributed path checked against the serial oracle on small METIS-partitioned meshes
           before anything is timed; the run fails when an error exceeds 1e-12
  fullsize the full-size result of this run against known answers of the unmodified reference
           (tests/golden/fullsize_norms.json)
  c4       BASELINE configs[3], 200^3 hex8 solid, a FIXED problem partitioned over the N GPUs (strong scaling:
           the north_star's >= 6x target is c4.jac_ms at N=1 over c4.jac_ms at N=8)
  c3, c5   BASELINE configs[2] / [4] (Quad9 composite cylinder 2M elements, 100^3 hex27), same treatment
  gmres    time per GMRES(m) iteration on the assembled C2 / C4 operators
"""
import argparse
import ctypes as C
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# Keep stdout to the one JSON line: libraries below us write there from C (NCCL prints its version banner on stdout
# whenever NCCL_DEBUG is VERSION or above, the reference prints a banner per assembler). File descriptor 1 is pointed at
# stderr for the whole run and the result line goes to the saved original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
sys.stdout = os.fdopen(os.dup(2), "w")


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "elements/sec for Jacobian+residual assembly (BCSR SpMV GB/s vs HBM peak under 'spmv')"
UNIT = "elements/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, read from the committed `ncu --set full` capture of this
# same command at the default workload (TRAFFIC_SOURCE); reported only when the run uses that workload on one GPU.
TRAFFIC_SOURCE = "profiles/r2_v_kernels_ncu.txt"
NCU_TRAFFIC_BYTES = {"shell4_mma_kernel": 0.197762e9 + 3.599907e9,        # 10 of 16 node-pair blocks staged / direct
                     "gather_blocks36_kernel": 2.645015e9 + 1.414355e9,   # reads the staged blocks once, writes 5.0 M blocks
                     "spmv6_kernel<0>": 2.689503e9 + 0.049784e9}
# minimal-algorithm flops per element used for the FP64 roofline
FLOPS_PER_ELEMENT = {1: 57e3, 2: 551e3, 3: 69e3, 4: 2.28e6}
KIND_NN = {1: 4, 2: 9, 3: 8, 4: 27}
KIND_BS = {1: 6, 2: 6, 3: 3, 4: 3}
PARITY_TOL = 1e-12


def spmv_bytes(bs, nrows, nnzb):
    """Algorithmic bytes of one SpMV: values + cols + rowp + x read once + y written once."""
    return nnzb * (8 * bs * bs + 4) + 4 * (nrows + 1) + 16 * bs * nrows


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled while the timed region runs: NVML polled in-process every few
    milliseconds (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as fallback."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device = device
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop = threading.Event()
        self.thread = None
        self.nvml = None
        self.error = None

    def _poll_nvml(self):
        n = self.nvml
        h = n.nvmlDeviceGetHandleByIndex(self.device)
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        errors = 0
        while not self.stop.is_set() and errors < 50:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mx.append(mx)
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception as exc:  # keep polling: one failed query must not end the sampling
                errors += 1
                self.error = str(exc)[:120]
            time.sleep(0.003)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while True:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout
                r = [c.strip() for c in out.strip().split(",")]
                self.sm.append(float(r[1]))
                self.mx.append(float(r[2]))
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:
                break
            if self.stop.is_set():
                break

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            target = self._poll_nvml
        except Exception:
            self.nvml = None
            target = self._poll_smi
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        if self.thread:
            self.thread.join(timeout=6)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.mx)), "reasons": sorted(self.reasons),
                "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi", "poll_error": self.error,
                "region": "device-timed assembleJacobian / assembleRes / SpMV loops and the end-to-end loop"}


def host_threads():
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


# ------------------------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref = the unmodified reference compiled where it lies)
# ------------------------------------------------------------------------------------------------------------
def _quiet_stdout():
    class Q:
        def __enter__(self):
            self.devnull = os.open(os.devnull, os.O_WRONLY)
            self.saved = os.dup(1)
            os.dup2(self.devnull, 1)  # the reference prints a banner on stdout; keep ours one JSON line

        def __exit__(self, *exc):
            os.dup2(self.saved, 1)
            os.close(self.devnull)
            os.close(self.saved)

    return Q()


def time_reference(nx, ny, steps, warmup, gmres_m=0):
    """Reference CPU implementation on an nx x ny Quad4 plate. The reference's intra-rank parallelism is its pthread
    work queue, capped at 16 threads (src/TACSObject.h:150)."""
    from tacs_b200 import TACS as T
    from tacs_b200 import binding, meshgen

    so = os.path.join(ROOT, "oracle", "_ref", "libtacs_ref.so")
    from tests import ref_binding
    ref = ref_binding.load_reference(so)
    mesh = meshgen.plate(2, nx, ny)
    with _quiet_stdout():
        creator, asm = meshgen.build_model(T, ref, mesh, [meshgen.iso_shell_element(T, ref, 2)])
    A, res, x, y = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec()
    x.setArray(meshgen.hash_vector(x.getSize()))
    asm.applyBCs(x)
    asm.setVariables(x)
    threads = min(host_threads(), 16)
    asm.setNumThreads(threads)
    for _ in range(warmup):
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    t0 = time.perf_counter()
    for _ in range(steps):
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    dt = (time.perf_counter() - t0) / steps
    # bs=6 SpMV is threaded in the reference as well
    xr = asm.createVec()
    xr.setArray(meshgen.hash_vector(x.getSize())[::-1].copy())
    asm.applyBCs(xr)
    A.mult(xr, y)
    t1 = time.perf_counter()
    nsp = 5
    for _ in range(nsp):
        A.mult(xr, y)
    dts = (time.perf_counter() - t1) / nsp
    bs, nrows, ncols, nnzb = A.getSizes()
    out = dict(elements=nx * ny, seconds_per_step=dt, threads=threads,
               spmv_gbs=spmv_bytes(bs, nrows, nnzb) / dts * 1e-9, ynorm=y.norm(), resnorm=res.norm())
    if gmres_m > 0:
        ksm = T.KSM(ref, A, gmres_m, 0)
        ksm.setTolerances(1e-30, 1e-300)
        sol = asm.createVec()
        t2 = time.perf_counter()
        ksm.solve(res, sol)
        out["gmres_ms_per_iter"] = (time.perf_counter() - t2) * 1e3 / max(ksm.getIterCount(), 1)
        out["gmres_iters"] = ksm.getIterCount()
    return out


def time_reference_mpi(nx, ny, nranks, reps):
    """The reference's MPI path: N ranks (one per host core) through TACSCreator + METIS, run by
    oracle/_ref/ref_driver over the forked-rank MPI stand-in (no mpirun in this image)."""
    from oracle import ref_mpi
    from tacs_b200 import meshgen

    mesh = meshgen.plate(2, nx, ny)
    summary, _, _ = ref_mpi.run(mesh, 1, 0, nranks, reps=reps, timeout=1500, load=False)
    return summary


def best_cpu_baseline(n, steps, warmup, gmres_m=0):
    """Fastest of the reference's two CPU parallel modes on this box: pthreads (<= 16) and MPI ranks."""
    cores = host_threads()
    r = time_reference(n, n, steps, warmup, gmres_m)
    best = {"value": r["elements"] / r["seconds_per_step"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
            "sample": f"{n}x{n} Quad4 plate ({r['elements']} elements per step), assembleJacobian(1,0,0), "
                      f"oracle/_ref, 1 rank x {r['threads']} pthreads (setNumThreads)",
            "spmv_gbs": r["spmv_gbs"], "seconds_per_step": r["seconds_per_step"], "ynorm": r["ynorm"],
            "resnorm": r["resnorm"], "gmres_ms_per_iter": r.get("gmres_ms_per_iter"), "gmres_iters": r.get("gmres_iters")}
    try:
        from oracle import ref_mpi

        if ref_mpi.available() and cores > 1:
            nranks = min(cores, 64)
            m = time_reference_mpi(n, n, nranks, max(steps, 2))
            if m and m["elements_per_s"] > best["value"]:
                best.update({"value": m["elements_per_s"], "cores": nranks,
                             "sample": f"{n}x{n} Quad4 plate ({m['elements']} elements per step), "
                                       f"assembleJacobian(1,0,0), oracle/_ref ref_driver, {nranks} MPI ranks "
                                       f"(forked-rank stand-in, METIS partition)",
                             "seconds_per_step": m["jac_s"], "pthreads_value": best["value"],
                             "pthreads_cores": best["cores"], "spmv_gbs_mpi": m.get("spmv_gbs"),
                             "spmv_note": "spmv_gbs: the reference's threaded bs=6 product on 1 rank x pthreads"})
            else:
                best["mpi_value"] = m["elements_per_s"] if m else None
                best["mpi_ranks"] = nranks
    except Exception as exc:
        best["mpi_error"] = str(exc)[:200]
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx = ny = args.ref_n
    cpu = best_cpu_baseline(nx, args.steps, args.warmup, gmres_m=args.gmres_m)
    value = cpu["value"]
    same = nx == args.nx and ny == args.ny
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cpu["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 1000x1000 Quad4Shell plate (BASELINE configs[1])",
                   "sample": f"{nx}x{ny} Quad4 plate of the same generator ({nx * ny} elements per step)"
                             + (" = the full configuration" if same else ""),
                   "same_config": bool(same), "timing": "host wall clock"},
        "cpu_baseline": cpu,
        "spmv": {"gbs": cpu.get("spmv_gbs"), "ynorm": cpu.get("ynorm")},
        "gmres": {"ms_per_iter": cpu.get("gmres_ms_per_iter"), "m": args.gmres_m, "iters": cpu.get("gmres_iters")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
# tacs_b200 arm
# ------------------------------------------------------------------------------------------------------------
def gpu_numa_node(device):
    """NUMA node of a GPU from its PCI address (sysfs), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        with open(f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


class PreferNumaNode:
    """Allocate host memory on the NUMA node of this rank's GPU while the block is active (set_mempolicy
    MPOL_PREFERRED): the pinned state / residual buffers of the end-to-end step then sit next to the GPU's PCIe root
    even when the process itself may only run on the CPUs of another socket."""

    SYS_SET_MEMPOLICY, MPOL_DEFAULT, MPOL_PREFERRED = 238, 0, 1  # x86_64

    def __init__(self, node):
        self.node, self.ok = node, False

    def _call(self, mode, node):
        libc = C.CDLL(None, use_errno=True)
        if node is None:
            return libc.syscall(self.SYS_SET_MEMPOLICY, mode, None, 0) == 0
        mask = (C.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        return libc.syscall(self.SYS_SET_MEMPOLICY, mode, mask, 16 * 64) == 0

    def __enter__(self):
        if self.node is not None:
            try:
                self.ok = self._call(self.MPOL_PREFERRED, self.node)
            except Exception:
                self.ok = False
        return self

    def __exit__(self, *exc):
        if self.ok:
            try:
                self._call(self.MPOL_DEFAULT, None)
            except Exception:
                pass


class Dist:
    """torch.distributed is the bootstrap only (broadcasts of the ncclUniqueId and of rank 0's METIS partition,
    the closing max-over-ranks); the data path of the library uses its own NCCL communicator."""

    def __init__(self, lib):
        import torch

        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.lib = lib
        if lib.init(self.local_rank) != 0:
            raise SystemExit("tacs_b200: no usable GPU")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            import tacs_b200

            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            buf = np.zeros(128, np.uint8)
            if self.rank == 0:
                assert lib.comm_unique_id(buf.ctypes.data_as(tacs_b200.binding.UP)) == 0
            uid = torch.from_numpy(buf.copy()).cuda()
            dist.broadcast(uid, 0)
            buf = uid.cpu().numpy().copy()
            assert lib.comm_init(self.rank, self.world, buf.ctypes.data_as(tacs_b200.binding.UP)) == 0

    def barrier(self):
        self.lib.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, v, op):
        if not self.dist:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, v):
        return self._reduce(v, self.dist.ReduceOp.MAX) if self.dist else v

    def sum(self, v):
        return self._reduce(v, self.dist.ReduceOp.SUM) if self.dist else v

    def bcast_i32(self, arr, n):
        """rank 0's int32 array of length n on every rank"""
        if not self.dist:
            return arr
        t = self.torch.from_numpy(np.ascontiguousarray(arr, np.int32)).cuda() if self.rank == 0 else \
            self.torch.empty(n, dtype=self.torch.int32, device="cuda")
        self.dist.broadcast(t, 0)
        return t.cpu().numpy()

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def build_partitioned(D, T, meshgen, lib, mesh, elements):
    """Creator -> Assembler; on several ranks rank 0 alone runs METIS (the reference's root does the same,
    TACSCreator.cpp:1104-1125) and the partition is broadcast, instead of every rank partitioning the global mesh."""
    t0 = time.time()
    part = None
    if D.world > 1:
        ne = int(mesh["elem_ids"].size)
        if D.rank == 0:
            c0 = T.Creator(lib, mesh["vars_per_node"])
            c0.setGlobalConnectivity(mesh["num_nodes"], mesh["ptr"], mesh["conn"], mesh["elem_ids"])
            c0.partitionMesh()
            part = c0.getElementPartition()
            del c0
        part = D.bcast_i32(part, ne)
    t1 = time.time()
    creator, asm = meshgen.build_model(T, lib, mesh, elements, part=part, split_size=D.world if part is not None else 0)
    return creator, asm, {"partition": t1 - t0, "create_tacs": time.time() - t1}


def collect_profile(lib, steps):
    """{kernel name as launched: (launches per step, ms per step)} from the library's event log."""
    import tacs_b200

    ms_k, cnt_k = np.zeros(8), np.zeros(8, np.int64)
    lib.profile_collect(tacs_b200.binding.dptr(ms_k), cnt_k.ctypes.data_as(C.POINTER(C.c_long)))
    out = {}
    for line in (lib.profile_named() or b"").decode().splitlines():
        name, cnt, ms = line.rsplit("|", 2)
        out[name] = (int(cnt) / steps, float(ms) / steps)
    return out


def plan_stats(lib, asm):
    out = (C.c_long * 8)()
    lib.assembler_get_plan_stats(asm.h, out)
    names = ["local_slots", "recv_slots", "direct_blocks", "staged_blocks", "gather_blocks", "gather_sources",
             "blocks", "node_pairs"]
    return dict(zip(names, (int(v) for v in out)))


def kernel_rooflines(kind, nelem_local, stats, prof, fp64_peak, hbm_peak, world, default_workload):
    """Roofline entries of the kernels of one assembleJacobian step, named by what was launched."""
    nn, bs = KIND_NN[kind], KIND_BS[kind]
    b2 = bs * bs
    # algorithmic HBM bytes per launch: the element kernel reads X/u/conn/direct map and writes the upper
    # node-pair blocks that need staging, the blocks it owns alone straight into the matrix, and the residual slots;
    # the block gather reads its sources (staging blocks + 4-byte source codes) and writes each remaining block once.
    elem_bytes = (nelem_local * (nn * (3 + bs) * 8 + nn * 4 + 4 + nn * bs * 8) + stats["node_pairs"] * 4 +
                  (stats["staged_blocks"] + stats["direct_blocks"]) * b2 * 8)
    gather_bytes = stats["gather_sources"] * (b2 * 8 + 4) + stats["gather_blocks"] * (b2 * 8 + 8)
    kernels = []
    chunked = any(cnt > 1.5 for name, (cnt, ms) in prof.items() if "element_kernel" in name or "_mma_kernel" in name)
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
        entry = {"kernel": name, "ms": ms, "launches_per_step": cnt}
        t = ms * 1e-3
        if "element_kernel" in name or "_mma_kernel" in name:
            entry.update({"bound": "fp64", "achieved": FLOPS_PER_ELEMENT[kind] * nelem_local / t * 1e-12,
                          "peak": fp64_peak, "unit": "TFLOP/s", "hbm_gbs": elem_bytes / t * 1e-9,
                          "algorithmic_bytes": elem_bytes,
                          "note": "flops = SURVEY 8d minimal-algorithm count; peak = live DFMA microbenchmark "
                                  "(tacsb200_measure_fp64_tflops; DMMA m8n8k4 shares that FP64 peak on B200, "
                                  "profiles/r1_probe_fp64.txt)"})
        elif name.startswith("gather_blocks") or name.startswith("gather_rows"):
            entry.update({"bound": "hbm", "achieved": gather_bytes / t * 1e-9, "peak": hbm_peak, "unit": "GB/s",
                          "algorithmic_bytes": gather_bytes})
        else:
            entry.update({"bound": None})   # small kernels of the step (residual gather, BCs, pack / unpack of the
            kernels.append(entry)           # exchanges): listed so that the step is accounted for, no roofline
            continue
        entry["frac"] = entry["achieved"] / entry["peak"] if entry["peak"] else None
        if chunked and ("gather_blocks" in name or "element_kernel" in name):
            entry["overlapped"] = ("element chunks and the gather of the previous chunk run concurrently on two streams: "
                                   "each kernel's ms is its own launch-to-end time while sharing the GPU, so the two "
                                   "overlap and their fractions understate what either reaches alone "
                                   "(TACSB200_OVERLAP_KINDS=0: 22.3 ms / 0.67 and 11.6 ms / 0.53 on C4, step 34.5 ms)")
        elif world > 1 and "gather_blocks" in name:
            entry["overlapped"] = ("on several ranks the gather of the blocks that read local staging only runs on a "
                                   "low-priority stream while the off-rank rows are packed, sent and received: its ms "
                                   "includes the time it shares the GPU with the exchange (one rank: 0.85 ms / 0.90)")
        entry["traffic"] = NCU_TRAFFIC_BYTES.get(name) if default_workload else None
        entry["traffic_source"] = TRAFFIC_SOURCE if entry["traffic"] else None
        kernels.append(entry)
    return kernels


def time_config(D, lib, asm, A, res, x, y, kind, nelem_total, steps, fp64_peak, hbm_peak, default_workload=False,
                with_res=True):
    """Device-timed assembleJacobian / assembleRes / SpMV of one assembled configuration (max over ranks)."""
    nelem_local = asm.getNumElements()
    for _ in range(3):
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    lib.profile_enable(1)
    collect_profile(lib, 1)
    D.barrier()
    lib.kernel_launches(1)
    ms = lib.time_assemble_jacobian(asm.h, 1.0, 0.0, 0.0, res.h, A.h, steps)
    launches = lib.kernel_launches(0)
    prof = collect_profile(lib, steps)
    lib.profile_enable(0)
    ms_res = lib.time_assemble_res(asm.h, res.h, steps) / steps if with_res else None
    lib.time_mat_mult(A.h, x.h, y.h, 3)
    D.barrier()
    nsp = 30
    ms_spmv = lib.time_mat_mult(A.h, x.h, y.h, nsp) / nsp
    D.barrier()
    assert ms > 0 and ms_spmv > 0, "device timing failed"
    ms_local = ms / steps
    ms = D.max(ms) / steps
    ms_spmv = D.max(ms_spmv)
    if with_res:
        ms_res = D.max(ms_res)
    sizes_a, sizes_b = A.getSizes(0), A.getSizes(1)
    bs = sizes_a[0]
    sp_bytes_local = spmv_bytes(bs, sizes_a[1], sizes_a[3] + sizes_b[3])
    sp_bytes = D.sum(float(sp_bytes_local))
    stats = plan_stats(lib, asm)
    kernels = kernel_rooflines(kind, nelem_local, stats, prof, fp64_peak, hbm_peak, D.world, default_workload)
    spmv_names = sorted(n for n in prof if n.startswith("spmv")) or [f"spmv{bs}_kernel<0>"]
    spmv = {"kernel": "spmv6_kernel<0>" if bs == 6 else "spmv3_kernel<0>", "ms": ms_spmv, "bound": "hbm",
            "achieved": sp_bytes / (ms_spmv * 1e-3) * 1e-9, "peak": hbm_peak * D.world, "unit": "GB/s",
            "bytes_per_launch": sp_bytes, "note": "aggregate over all ranks; peak = ranks x measured HBM peak"}
    spmv["frac"] = spmv["achieved"] / spmv["peak"]
    # what the kernel list does not explain: NCCL send / recv of the exchanges, launch gaps, and -- the step is the max
    # over ranks, the kernel times are this rank's -- the imbalance of the partition. Negative when kernels overlap
    # (the gather on its own stream runs beside the exchange / the next element chunk).
    accounted = sum(k["ms"] for k in kernels if not (k.get("overlapped") and "gather" in k["kernel"]))
    step_rest = {"kernels_ms_this_rank": accounted, "step_minus_kernels_ms": ms - accounted,
                 "step_ms_min_over_ranks": -D.max(-ms_local), "step_ms_max_over_ranks": ms}
    return {"ms": ms, "ms_res": ms_res, "ms_spmv": ms_spmv, "launches": int(launches), "kernels": kernels,
            "step_accounting": step_rest,
            "plan": stats,
            "spmv": spmv, "value": nelem_total / (ms * 1e-3), "spmv_kernel_names": spmv_names}


def serial_dof_index(D, T, lib, mesh, creator, lo, hi):
    """Index of every owned dof of this rank in the SERIAL (one-rank, first-touch) numbering of the reference, so that
    the deterministic state / input vectors are the same physical vectors on any number of ranks (and the ones the
    known answers in tests/golden/fullsize_norms.json were computed for)."""
    bs = mesh["vars_per_node"]
    if D.world == 1:
        return None
    c1 = T.Creator(lib, bs)
    c1.setGlobalConnectivity(mesh["num_nodes"], mesh["ptr"], mesh["conn"], mesh["elem_ids"])
    c1.partitionMesh(1, np.zeros(mesh["elem_ids"].size, np.int32))
    serial = c1.getNodeNums().astype(np.int64)       # original node -> serial number
    dist_nodes = creator.getNodeNums().astype(np.int64)  # original node -> number on this partition
    inv = np.empty(creator.num_nodes, np.int64)
    inv[dist_nodes] = np.arange(creator.num_nodes)
    own = serial[inv[lo:hi]]
    return (bs * own[:, None] + np.arange(bs)[None, :]).ravel()


def fullsize_check(D, name, res, y, dof_index):
    """This run's residual / A*x against the known answers of the unmodified reference for the same configuration
    (tests/golden/fullsize_norms.json, generated by tests/golden/make_fullsize_norms.py). On several ranks the norms
    and the weighted checksum are reduced over the ranks; the sampled entries are compared on one rank only."""
    path = os.path.join(ROOT, "tests", "golden", "fullsize_norms.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        gold = json.load(f).get(name)
    if not gold:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_fullsize_norms import checksum_weights

    out = {"against": f"tests/golden/fullsize_norms.json[{name}] (unmodified reference, oracle/_ref)"}
    w = checksum_weights(gold["dof"])
    for key, v in (("res", res), ("y", y)):
        g = gold[key]
        vec = v.getArray()
        if D.world == 1 and vec.size != gold["dof"]:
            return {"error": f"size mismatch {vec.size} vs {gold['dof']}"}
        wl = w if dof_index is None else w[dof_index]
        out[key] = {"norm2_rel": abs(v.norm() - g["norm2"]) / g["norm2"],
                    "checksum_rel": abs(D.sum(float(np.dot(wl, vec))) - g["checksum"]) / (g["norm2"] * np.sqrt(gold["dof"]))}
        if D.world == 1:
            idx = np.asarray(g["sample_idx"])
            out[key]["sample_rel_max"] = float(np.abs(vec[idx] - np.asarray(g["sample"])).max() / g["max"])
    out["tol"] = 1e-10
    out["ok"] = all(val < 1e-10 for k in ("res", "y") for val in out[k].values())
    return out


def gmres_streaming_bound_ms(m, n_local_max, spmv_ms, hbm_gbs):
    """SpMV + the vector passes of modified Gram-Schmidt at the HBM peak: sweep j of iteration i reads w, the previous
    and the next basis vector and writes w (first sweep 2 passes, last 3), plus the normalisation (2): 4 i + 7 passes."""
    passes = sum(4 * i + 7 for i in range(m)) / m
    return spmv_ms + passes * n_local_max * 8 / (hbm_gbs * 1e9) * 1e3, passes


def time_gmres(D, lib, T, asm, A, b, m):
    """ms per GMRES(m) iteration on the assembled operator: one cycle of m iterations, tolerances that cannot be met."""
    ksm = T.KSM(lib, A, m, 0)
    ksm.setTolerances(1e-30, 1e-300)
    sol = asm.createVec()
    ksm.solve(b, sol)  # warm-up: the first solve sizes the buffers,
    ksm.solve(b, sol)  # the second captures the iteration graphs, the timed one replays them
    D.barrier()
    t0 = time.perf_counter()
    ksm.solve(b, sol)
    lib.synchronize()
    dt = D.max(time.perf_counter() - t0)
    it = max(ksm.getIterCount(), 1)
    return {"ms_per_iter": dt * 1e3 / it, "m": m, "iters": it, "orthogonalisation": "modified Gram-Schmidt",
            "timing": "host wall clock around solve(), max over ranks", "dof_per_rank_max": D.max(float(b.getSize()))}


EXTRA_CONFIGS = {
    # name: (kind, mesh factory, element factory, description)
    "c3": (2, lambda mg: mg.cylinder(3, 1000, 2000), lambda mg, T, lib: mg.composite_shell_element(T, lib, 3),
           "BASELINE configs[2]: Quad9Shell composite-laminate cylinder, 1000x2000 = 2M elements, 48M dof"),
    "c4": (3, lambda mg: mg.cube(2, 200), lambda mg, T, lib: mg.solid_element(T, lib, 2),
           "BASELINE configs[3]: 200^3 hex8 solid, 8M elements, 24.4M dof"),
    "c5": (4, lambda mg: mg.cube(3, 100), lambda mg, T, lib: mg.solid_element(T, lib, 3),
           "BASELINE configs[4]: 100^3 hex27 solid, 1M elements, 24.4M dof"),
}


def run_extra(D, lib, T, meshgen, name, steps, fp64_peak, hbm_peak, gmres_m):
    kind, mesh_f, elem_f, what = EXTRA_CONFIGS[name]
    t0 = time.time()
    mesh = mesh_f(meshgen)
    creator, asm, setup = build_partitioned(D, T, meshgen, lib, mesh, [elem_f(meshgen, T, lib)])
    t1 = time.time()
    A = asm.createMat()
    setup["create_mat"] = time.time() - t1
    setup["mesh"] = t1 - t0 - setup["partition"] - setup["create_tacs"]
    res, u, x, y = asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    bs = mesh["vars_per_node"]
    lo, hi = asm.getOwnerRange()
    h = meshgen.hash_vector(bs * creator.num_nodes)
    idx = serial_dof_index(D, T, lib, mesh, creator, lo, hi)
    u.setArray(h if idx is None else h[idx])
    x.setArray(h[::-1].copy() if idx is None else h[::-1][idx])
    asm.applyBCs(u)
    asm.applyBCs(x)
    asm.setVariables(u)
    nelem_total = int(mesh["elem_ids"].size)
    r = time_config(D, lib, asm, A, res, x, y, kind, nelem_total, steps, fp64_peak, hbm_peak)
    A.mult(x, y)
    out = {"workload": what, "n_gpus": D.world, "scaling": "strong (fixed problem partitioned over the GPUs)",
           "partition": "METIS (rank 0, broadcast)" if D.world > 1 else "single rank",
           "elements": nelem_total, "jac_ms": r["ms"], "elements_per_s": r["value"], "res_ms": r["ms_res"],
           "spmv_ms": r["ms_spmv"], "spmv_gbs_aggregate": r["spmv"]["achieved"], "spmv_frac_of_hbm_peak": r["spmv"]["frac"],
           "kernels": r["kernels"], "step_accounting": r["step_accounting"], "plan": r["plan"], "ynorm": y.norm(),
           "resnorm": res.norm(), "setup_s": setup,
           "local_elements_rank0": asm.getNumElements()}
    if name == "c4":
        out["fullsize"] = fullsize_check(D, "c4", res, y, idx)
    if gmres_m > 0 and name == "c4":
        out["gmres"] = time_gmres(D, lib, T, asm, A, res, gmres_m)
        bound, passes = gmres_streaming_bound_ms(gmres_m, out["gmres"]["dof_per_rank_max"], r["ms_spmv"], hbm_peak)
        out["gmres"].update({"streaming_bound_ms": bound, "vector_passes_per_iter": passes,
                             "ratio_to_bound": out["gmres"]["ms_per_iter"] / bound})
    del A, asm, creator, res, u, x, y
    gc.collect()
    return out


def run_b200(args):
    import torch

    import tacs_b200
    from tacs_b200 import TACS as T
    from tacs_b200 import meshgen

    lib = tacs_b200.load()  # raises when libtacs_b200.so is missing: there is no fallback
    D = Dist(lib)
    world, rank = D.world, D.rank

    # ---- multi-GPU parity gate: distributed assembly + SpMV against the serial oracle, before anything is timed ----
    parity = None
    if world > 1:
        from tests import dist_check  # checker only (oracle/tacs_oracle.c through tests/oracle_port.py)

        errs = dist_check.check_all(lib)
        worst = {k: D.max(v) for k, v in errs["max"].items()}
        exact = D.max(0.0 if errs["pattern_exact"] else 1.0) == 0.0
        parity = {"A": worst["A"], "res": worst["res"], "y": worst["y"], "res_only": worst["res_only"],
                  "norm": worst["norm"], "dot": worst["dot"], "gmres_displacements": worst["gmres"],
                  "gmres_tol": 1e-10, "pattern_exact": exact, "tol": PARITY_TOL,
                  "cases": sorted(k for k in errs if k not in ("max", "pattern_exact", "gmres_converged")),
                  "what": "every owned row of the METIS-partitioned matrix / residual / A*x on every rank (NCCL halo and "
                          "off-rank staging rows included) vs the serial oracle in the same numbering; max over ranks"}
        parity["ok"] = bool(exact and all(worst[k] < PARITY_TOL for k in ("A", "res", "y", "res_only", "norm")) and
                            worst["gmres"] < 1e-10 and D.max(0.0 if errs["gmres_converged"] else 1.0) == 0.0)
        if not parity["ok"]:
            if rank == 0:
                emit({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "parity": parity,
                      "error": "multi-GPU parity check failed; nothing was timed"})
            D.close()
            raise SystemExit(2)

    hbm_peak, hbm_src = measured_peaks()
    fp64_peak = lib.measure_fp64_tflops()

    # ---- headline: C2 plate, 1M elements per GPU -------------------------------------------------------------------
    nx, ny = args.nx * world, args.ny
    mesh = meshgen.plate(2, nx, ny)
    creator, asm, setup = build_partitioned(D, T, meshgen, lib, mesh, [meshgen.iso_shell_element(T, lib, 2)])
    A, res, x, y, xr = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec(), asm.createVec()
    n = x.getSize()
    numa = PreferNumaNode(gpu_numa_node(D.local_rank))
    with numa:
        state = torch.empty(n, dtype=torch.float64).pin_memory()
        out = torch.empty(n, dtype=torch.float64).pin_memory()
    state_np, out_np = state.numpy(), out.numpy()
    lo, hi = asm.getOwnerRange()
    h = meshgen.hash_vector(6 * creator.num_nodes)
    idx = serial_dof_index(D, T, lib, mesh, creator, lo, hi)
    state_np[:] = h if idx is None else h[idx]
    x.setArray(state_np)
    asm.applyBCs(x)
    state_np[:] = x.getArray()  # the host copy of the state satisfies the boundary conditions as well
    asm.setVariables(x)
    xr.setArray(h[::-1].copy() if idx is None else h[::-1][idx])
    asm.applyBCs(xr)
    nelem_total = nx * ny
    bs, nrows, ncols, nnzb = A.getSizes()
    default_workload = world == 1 and args.nx == 1000 and args.ny == 1000

    clocks = ClockSampler(D.local_rank)
    clocks.__enter__()
    r = time_config(D, lib, asm, A, res, xr, y, 1, nelem_total, args.steps, fp64_peak, hbm_peak, default_workload)
    ms_per_step, value = r["ms"], r["value"]

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    def e2e_step():
        # one call of the public host-buffer entry point (C ABI tacsb200_assembler_assemble_jacobian_host): state from
        # pinned host memory (H2D inside), residual back into pinned host memory (D2H inside); returns when the
        # residual has arrived, the block gather of the matrix may still be running
        asm.assembleJacobianHost(1.0, 0.0, 0.0, state_np, out_np, A)

    for _ in range(2):
        e2e_step()
    D.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    lib.synchronize()  # the last step's block gather and matrix BCs are inside the timed region
    D.barrier()
    e2e_s = D.max((time.perf_counter() - t0) / args.steps)
    e2e_value = nelem_total / e2e_s
    clocks.__exit__()
    # where an end-to-end step spends its time on this rank's host thread (outside the timed region): the pinned
    # H2D copy of the state, the enqueue of setVariables + assembleJacobian, the wait for the residual (D2H), the rest
    # of the device work. Max over ranks: the ranks of one node share the host's copy engines and memory channels.
    phases = np.zeros(6)
    for _ in range(args.steps):
        # the same work as separate, blocking calls: what each transfer costs when nothing overlaps it
        t_a = time.perf_counter()
        x.setArray(state_np)
        t_b = time.perf_counter()
        asm.setVariables(x)
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A, wait=False)
        t_c = time.perf_counter()
        lib.vec_get_array(res.h, tacs_b200.binding.dptr(out_np))
        t_d = time.perf_counter()
        lib.synchronize()
        t_e = time.perf_counter()
        # ... and the pipelined entry point the end-to-end number uses
        asm.assembleJacobianHost(1.0, 0.0, 0.0, state_np, out_np, A)
        t_f = time.perf_counter()
        lib.synchronize()
        t_g = time.perf_counter()
        phases += [t_b - t_a, t_c - t_b, t_d - t_c, t_e - t_d, t_f - t_e, t_g - t_f]
    e2e_breakdown = {k: D.max(float(v) / args.steps * 1e3) for k, v in
                     zip(("separate_calls_h2d_state_ms", "separate_calls_enqueue_ms",
                          "separate_calls_residual_d2h_wait_ms", "separate_calls_matrix_tail_ms",
                          "pipelined_call_ms", "pipelined_matrix_tail_ms"), phases)}
    e2e_breakdown["pinned_buffers_numa_node"] = numa.node if numa.ok else None
    try:
        e2e_breakdown["host_cpus_visible_to_this_rank"] = len(os.sched_getaffinity(0))
    except AttributeError:
        pass

    fullsize = None
    if default_workload:
        A.mult(xr, y)
        fullsize = fullsize_check(D, "c2", res, y, idx)
    gmres = time_gmres(D, lib, T, asm, A, res, args.gmres_m) if args.gmres_m > 0 else None
    if gmres:
        bound, passes = gmres_streaming_bound_ms(args.gmres_m, gmres["dof_per_rank_max"], r["ms_spmv"], hbm_peak)
        gmres.update({"streaming_bound_ms": bound, "vector_passes_per_iter": passes,
                      "ratio_to_bound": gmres["ms_per_iter"] / bound})
    del A, asm, creator, res, x, y, xr
    gc.collect()

    # ---- the other BASELINE configurations: fixed problems partitioned over the N GPUs (strong scaling) ------------
    extra = {}
    for name in [c for c in args.configs.split(",") if c]:
        try:
            extra[name] = run_extra(D, lib, T, meshgen, name, max(3, args.steps // 3), fp64_peak, hbm_peak, args.gmres_m)
        except Exception as exc:  # a configuration that does not fit must not take the headline down
            extra[name] = {"error": str(exc)[:300]}
            if world > 1:
                raise

    if rank != 0:
        D.close()
        return

    kernels = r["kernels"]
    dominant = max(kernels, key=lambda k: k["ms"]) if kernels else None
    roofline = None
    if dominant:
        roofline = {"kernel": dominant["kernel"], "bound": dominant["bound"], "achieved": dominant["achieved"],
                    "peak": dominant["peak"], "unit": dominant["unit"], "frac": dominant["frac"],
                    "traffic": dominant["traffic"], "traffic_source": dominant["traffic_source"],
                    "algorithmic_bytes": dominant["algorithmic_bytes"],
                    "peak_source": hbm_src if dominant["bound"] == "hbm" else "live DFMA microbenchmark",
                    "share_of_step": dominant["ms"] / ms_per_step}
    spmv = r["spmv"]
    spmv["peak_source"] = hbm_src
    spmv["traffic"] = NCU_TRAFFIC_BYTES.get(spmv["kernel"]) if default_workload else None
    spmv["traffic_source"] = TRAFFIC_SOURCE if spmv["traffic"] else None

    # ---- CPU baseline: the reference's own implementation on this box's host cores ---------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu = best_cpu_baseline(args.ref_n, 2, 1)
        except Exception as exc:  # the reference build did not travel
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {exc}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic {nx}x{ny} Quad4Shell plate (6 dof/node), isotropic, edges clamped: "
                               f"assembleJacobian(1,0,0,res,A) per step; BASELINE configs[1] at 1 GPU, "
                               f"{args.nx}x{args.ny} elements per GPU",
                   "elements": nelem_total, "dof": 6 * (nx + 1) * (ny + 1), "nnzb": int(nnzb),
                   "l2": "inputs larger than L2 (staging + matrix of several GB per step)",
                   "partition": "METIS element partition (TACSCreator::partitionMesh on rank 0, broadcast)"
                   if world > 1 else "single rank",
                   "strong_scaling": "see c4 / c3 / c5: fixed problems partitioned over the N GPUs"},
        "roofline": roofline, "kernels": kernels, "step_accounting": r["step_accounting"], "plan": r["plan"],
        "spmv": spmv,
        "assemble_res": {"ms": r["ms_res"], "value": nelem_total / (r["ms_res"] * 1e-3), "unit": UNIT,
                         "note": "assembleRes alone (SURVEY 8d metric i), same mesh and state"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(n * 8),
                "ms_per_step": e2e_s * 1e3, "breakdown": e2e_breakdown,
                "note": "tacsb200_assembler_assemble_jacobian_host per step: state vector from pinned host memory (H2D in "
                        "pieces on a copy stream, each chunk of elements starts when the piece with its last node has "
                        "arrived; on several ranks: upload, halo, then the kernels) -> assembleJacobian -> residual to "
                        "pinned host memory on the copy stream while the block gather still runs; the region ends with "
                        "a device synchronize; the BCSR matrix stays in HBM for the device-side Krylov solver"},
        "gpu_launches": r["launches"], "clocks": clocks.summary(),
        "fp64_peak_tflops": fp64_peak, "setup_s": setup,
    }
    if parity is not None:
        line["parity"] = parity
    if fullsize is not None:
        line["fullsize"] = fullsize
    if gmres is not None:
        line["gmres"] = gmres
    line.update(extra)
    emit(line)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=1000, help="plate elements per GPU along x")
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--ref-n", type=int, default=1000, help="edge of the CPU-baseline plate (1000 = the full config)")
    ap.add_argument("--configs", default="c4,c3,c5", help="extra BASELINE configurations to run after the headline")
    ap.add_argument("--gmres-m", type=int, default=30, help="GMRES subspace size of the per-iteration timing (0: skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
