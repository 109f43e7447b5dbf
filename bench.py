#!/usr/bin/env python
"""Benchmark of the TACS assembly + Krylov-operator hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # tacs_b200 (CUDA, sm_100a)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path

Workload (BASELINE.json configs[1]): synthetic 1000x1000 Quad4 MITC shell plate, 6 dof/node
(~6M dof), isotropic, all edges clamped.  At N > 1 GPUs the plate grows to (1000*N) x 1000 elements
(weak scaling: 1M elements per GPU), partitioned by the reference's METIS call.

A step is one `assembleJacobian(1, 0, 0, res, A)`: element residuals + tangents, atomic-free
gather into the BCSR matrix and the residual, boundary conditions.  `value` is elements/s with
everything resident in HBM (CUDA events on the library's stream); `e2e` repeats the step through the
C ABI with the state vector arriving from pinned host memory and the residual going back to it.
The BCSR SpMV (the other half of the metric) is timed in its own loop and reported under `spmv`.
"""
import argparse
import ctypes as C
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
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures at the default workload
# (profiles/r1_n_kernels_ncu.txt); reported only when the run uses that workload on one GPU
NCU_TRAFFIC_BYTES = {"shell4_mma_kernel": 0.142545e9 + 4.742443e9, "gather_blocks36_kernel": 4.744081e9 + 2.584352e9,
                     "spmv6_kernel<0>": 2.689707e9 + 0.050150e9}
# minimal-algorithm flops per element used for the FP64 roofline
FLOPS_PER_ELEMENT = {"quad4": 57e3, "quad9": 551e3, "hex8": 69e3, "hex27": 2.28e6}


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

    def _poll_nvml(self):
        n = self.nvml
        h = n.nvmlDeviceGetHandleByIndex(self.device)
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown,
                "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown,
                "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        mx = float(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM))
        while not self.stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                self.mx.append(mx)
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                break
            time.sleep(0.004)

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
                "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def host_threads():
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


def time_reference(nx, ny, steps, warmup):
    """Reference CPU implementation (oracle/_ref, compiled from the unmodified sources) on an nx x ny plate.
    The reference's intra-rank parallelism is its pthread work queue, capped at 16 threads
    (src/TACSObject.h:150); MPI is not available in this image."""
    from tacs_b200 import TACS as T
    from tacs_b200 import binding, meshgen

    so = os.path.join(ROOT, "oracle", "_ref", "libtacs_ref.so")
    ref = binding.Lib(so, "ref_")
    mesh = meshgen.plate(2, nx, ny)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)  # the reference prints a banner on stdout; keep ours one JSON line
    try:
        creator, asm = meshgen.build_model(T, ref, mesh, [meshgen.iso_shell_element(T, ref, 2)])
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
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
    A.mult(x, y)
    t1 = time.perf_counter()
    nsp = 10
    for _ in range(nsp):
        A.mult(x, y)
    dts = (time.perf_counter() - t1) / nsp
    bs, nrows, ncols, nnzb = A.getSizes()
    return dict(elements=nx * ny, seconds_per_step=dt, threads=threads,
                spmv_gbs=spmv_bytes(bs, nrows, nnzb) / dts * 1e-9, ynorm=y.norm())


def time_reference_mpi(nx, ny, nranks, reps):
    """The reference's MPI path: N ranks (one per host core) through TACSCreator + METIS, run by
    oracle/_ref/ref_driver over the forked-rank MPI stand-in (no mpirun in this image)."""
    from oracle import ref_mpi
    from tacs_b200 import meshgen

    mesh = meshgen.plate(2, nx, ny)
    summary, _, _ = ref_mpi.run(mesh, 1, 0, nranks, reps=reps, timeout=1500, load=False)
    return summary


def best_cpu_baseline(n, steps, warmup):
    """Fastest of the reference's two CPU parallel modes on this box: pthreads (<= 16) and MPI ranks."""
    cores = host_threads()
    r = time_reference(n, n, steps, warmup)
    best = {"value": r["elements"] / r["seconds_per_step"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
            "sample": f"{n}x{n} Quad4 plate ({r['elements']} elements per step), assembleJacobian(1,0,0), "
                      f"oracle/_ref, 1 rank x {r['threads']} pthreads (setNumThreads)",
            "spmv_gbs": r["spmv_gbs"], "seconds_per_step": r["seconds_per_step"]}
    try:
        from oracle import ref_mpi

        if ref_mpi.available() and cores > 1:
            nranks = min(cores, 64)
            m = time_reference_mpi(n, n, nranks, max(steps, 2))
            if m and m["elements_per_s"] > best["value"]:
                best = {"value": m["elements_per_s"], "unit": UNIT, "cores": nranks, "kind": "reference",
                        "sample": f"{n}x{n} Quad4 plate ({m['elements']} elements per step), assembleJacobian(1,0,0), "
                                  f"oracle/_ref ref_driver, {nranks} MPI ranks (forked-rank stand-in, METIS partition)",
                        "spmv_gbs": None, "seconds_per_step": m["jac_s"],
                        "pthreads_value": best["value"], "pthreads_cores": best["cores"]}
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
    cpu = best_cpu_baseline(nx, args.steps, args.warmup)
    value = cpu["value"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cpu["seconds_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic 1000x1000 Quad4Shell plate (BASELINE configs[1])",
                   "sample": f"{nx}x{ny} Quad4 plate of the same generator ({nx * ny} elements per step)",
                   "timing": "host wall clock"},
        "cpu_baseline": cpu,
        "spmv": {"gbs": cpu.get("spmv_gbs")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import tacs_b200
    from tacs_b200 import TACS as T
    from tacs_b200 import meshgen

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    lib = tacs_b200.load()  # raises when libtacs_b200.so is missing: there is no fallback
    if lib.init(local_rank) != 0:
        raise SystemExit("tacs_b200: no usable GPU")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (np.zeros(128, np.uint8))
            assert lib.comm_unique_id(buf.ctypes.data_as(tacs_b200.binding.UP)) == 0
            uid = torch.from_numpy(buf.copy())
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        buf = uid.cpu().numpy().copy()
        assert lib.comm_init(rank, world, buf.ctypes.data_as(tacs_b200.binding.UP)) == 0

    nx, ny = args.nx * world, args.ny
    mesh = meshgen.plate(2, nx, ny)
    creator, asm = meshgen.build_model(T, lib, mesh, [meshgen.iso_shell_element(T, lib, 2)])
    A, res, x, y = asm.createMat(), asm.createVec(), asm.createVec(), asm.createVec()
    n = x.getSize()
    state = torch.empty(n, dtype=torch.float64).pin_memory()
    out = torch.empty(n, dtype=torch.float64).pin_memory()
    state_np, out_np = state.numpy(), out.numpy()
    lo, hi = asm.getOwnerRange()
    state_np[:] = meshgen.hash_vector(6 * creator.num_nodes)[6 * lo:6 * hi] if world > 1 else meshgen.hash_vector(n)
    x.setArray(state_np)
    asm.applyBCs(x)
    asm.setVariables(x)
    nelem_local = asm.getNumElements()
    nelem_total = nx * ny
    bs, nrows, ncols, nnzb = A.getSizes()

    def barrier():
        lib.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A)
    lib.profile_enable(1)
    ms_k, cnt_k = np.zeros(8), np.zeros(8, np.int64)
    lib.profile_collect(tacs_b200.binding.dptr(ms_k), cnt_k.ctypes.data_as(C.POINTER(C.c_long)))
    barrier()
    lib.kernel_launches(1)
    with ClockSampler(local_rank) as clocks:
        ms = lib.time_assemble_jacobian(asm.h, 1.0, 0.0, 0.0, res.h, A.h, args.steps)
        launches = lib.kernel_launches(0)
        lib.profile_collect(tacs_b200.binding.dptr(ms_k),
                            cnt_k.ctypes.data_as(C.POINTER(C.c_long)))
        lib.profile_enable(0)
        ms_res = lib.time_assemble_res(asm.h, res.h, args.steps) / args.steps
        # SpMV loop inside the same clock window
        lib.time_mat_mult(A.h, x.h, y.h, 3)
        nsp = 50
        ms_spmv = lib.time_mat_mult(A.h, x.h, y.h, nsp) / nsp
    lib.profile_enable(0)
    barrier()
    assert ms > 0, "device timing failed"
    ms = max_over_ranks(ms)
    ms_spmv = max_over_ranks(ms_spmv)
    ms_res = max_over_ranks(ms_res)
    ms_per_step = ms / args.steps
    value = nelem_total / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with host buffers -------------------------------------------
    def e2e_step():
        x.setArray(state_np)           # H2D from pinned memory
        asm.setVariables(x)
        asm.assembleJacobian(1.0, 0.0, 0.0, res, A, wait=False)   # enqueue only (C ABI: ..._assemble_jacobian_async)
        lib.vec_get_array(res.h, tacs_b200.binding.dptr(out_np))  # D2H into pinned memory, behind the residual only

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    e2e_value = nelem_total / e2e_s

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the kernels in the step ----------------------------------------------------------
    hbm_peak, hbm_src = measured_peaks()
    fp64_peak = lib.measure_fp64_tflops()
    names = ["element", "gather_residual", "gather_blocks", "boundary_conditions", "spmv", "vector", "dot", "halo"]
    # per-step device time of each kernel family (a family may launch more than once per step, e.g. the
    # Aloc and Bext gathers on several ranks)
    per_launch = {names[k]: (ms_k[k] / args.steps if cnt_k[k] else None) for k in range(8)}
    nn, b2 = 4, 36
    # algorithmic HBM bytes per launch: element kernel reads X/u/conn and writes the staging
    # blocks + residual slots; block gather reads the staging blocks and the plan, writes A once.
    elem_bytes = nelem_local * (nn * (3 + 6) * 8 + nn * 4 + 4 + nn * nn * b2 * 8 + nn * 6 * 8)
    gather_bytes = nelem_local * nn * nn * (b2 * 8 + 4) + nnzb * (b2 * 8 + 4)
    kernels = []
    if per_launch["element"]:
        t = per_launch["element"] * 1e-3
        kernels.append({"kernel": "shell4_mma_kernel", "ms": per_launch["element"], "bound": "fp64",
                        "achieved": FLOPS_PER_ELEMENT["quad4"] * nelem_local / t * 1e-12, "peak": fp64_peak,
                        "unit": "TFLOP/s", "hbm_gbs": elem_bytes / t * 1e-9,
                        "note": "flops = SURVEY 8d minimal-algorithm count (57 kFLOP/element); peak = live DFMA "
                                "microbenchmark (tacsb200_measure_fp64_tflops); the kernel issues its two matrix "
                                "products as DMMA m8n8k4, which shares that FP64 peak on B200 "
                                "(profiles/r1_probe_fp64.txt: 36.5 DFMA vs 37.0 DMMA TFLOP/s)"})
    if per_launch["gather_blocks"]:
        t = per_launch["gather_blocks"] * 1e-3
        kernels.append({"kernel": "gather_blocks36_kernel", "ms": per_launch["gather_blocks"], "bound": "hbm",
                        "achieved": gather_bytes / t * 1e-9, "peak": hbm_peak, "unit": "GB/s"})
    default_workload = world == 1 and args.nx == 1000 and args.ny == 1000
    for k in kernels:
        k["frac"] = k["achieved"] / k["peak"] if k["peak"] else None
        k["traffic"] = NCU_TRAFFIC_BYTES.get(k["kernel"]) if default_workload else None
    dominant = max(kernels, key=lambda k: k["ms"]) if kernels else None
    roofline = None
    if dominant:
        roofline = {"kernel": dominant["kernel"], "bound": dominant["bound"], "achieved": dominant["achieved"],
                    "peak": dominant["peak"], "unit": dominant["unit"], "frac": dominant["frac"],
                    "traffic": dominant["traffic"], "algorithmic_bytes": elem_bytes if dominant["bound"] == "fp64" else gather_bytes,
                    "peak_source": hbm_src if dominant["bound"] == "hbm" else "live DFMA microbenchmark",
                    "share_of_step": dominant["ms"] / ms_per_step}
    sp_bytes = spmv_bytes(bs, nrows, nnzb)
    spmv = {"kernel": "spmv6_kernel<0>", "ms": ms_spmv, "bound": "hbm", "achieved": sp_bytes / (ms_spmv * 1e-3) * 1e-9,
            "peak": hbm_peak, "unit": "GB/s", "peak_source": hbm_src, "bytes_per_launch": sp_bytes}
    spmv["frac"] = spmv["achieved"] / spmv["peak"]
    spmv["traffic"] = NCU_TRAFFIC_BYTES["spmv6_kernel<0>"] if default_workload else None

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
                   "elements": nelem_total, "dof": 6 * creator.num_nodes, "nnzb": int(nnzb),
                   "l2": "inputs larger than L2 (staging 4.6 GB + matrix 2.6 GB per step)",
                   "partition": "METIS element partition (TACSCreator::partitionMesh)" if world > 1 else "single rank"},
        "roofline": roofline, "kernels": kernels, "spmv": spmv,
        "assemble_res": {"ms": ms_res, "value": nelem_total / (ms_res * 1e-3), "unit": UNIT,
                         "note": "assembleRes alone (SURVEY 8d metric i), same mesh and state"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * 8), "d2h_bytes_per_step": int(n * 8),
                "note": "state vector from pinned host memory -> setVariables -> assembleJacobian (enqueue-only C ABI "
                        "entry) -> residual to pinned host memory on the copy stream while the block gather still runs; "
                        "the region ends with a device synchronize; the BCSR matrix stays in HBM for the device-side "
                        "Krylov solver"},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
        "fp64_peak_tflops": fp64_peak,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=1000, help="plate elements per GPU along x")
    ap.add_argument("--ny", type=int, default=1000)
    ap.add_argument("--ref-n", type=int, default=300, help="edge of the bounded CPU-baseline sample plate")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
