#!/usr/bin/env python
"""bench.py -- the contract benchmark of the assembly hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--impl reference]
  (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one fp64 Jacobian-type evaluation that fills both f and A for the 3-D Poisson Q1-hex block:
AssemblyEngine<Jacobian>::evaluate(in, All) = halo import -> volume fill -> Dirichlet rows -> halo export
(BASELINE.json configs[1]).  Weak scaling: every GPU owns an n^3 brick
(n=256 -> 16.7 M elements per GPU) of a CubeHexMeshFactory mesh split like the reference splits it.

One JSON line on rank 0: value = Melem/s over all GPUs with x, coordinates, LIDs, graph resident in HBM;
e2e = the same evaluate called with HOST (pinned) x and f buffers, copies inside the timed region;
roofline = algorithmic bytes (288 B/element) / measured duration of the fill kernel vs the measured HBM
peak; cpu_baseline = the CPU restatement of the reference algorithm (oracle/) on the host cores.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_PER_ELEM = 288   # x 8 + coords 24 + LIDs 32 + f 8 + A 27*8
METRIC = "Jacobian+residual assembly Melem/s (Q1 hex 256^3)"


def measured_traffic(n_cells, mode):
    """dram__bytes_read+write of the fill kernel from the committed ncu capture, if it is for this workload."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if t["n_cells"] == n_cells and mode == 1:
            return t["traffic_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
            while not self._halt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.02)
        except Exception as e:  # NVML missing: report it instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------------------------------------
def cpu_reference(n_sample, threads, steps=1, warmup=0):
    """Time the CPU restatement of the reference algorithm (oracle/) on n_sample^3 elements."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    (d,), _ = orc.poisson_problem(n_sample)
    t = orc.tables_build(d["cell_coords"])                  # cached per-workset tables: setup, untimed (as in the reference)
    x = orc.state_by_gid(np.arange(d["n_local"]))
    tm = orc.make_terms(nthreads=threads)
    f = np.zeros(d["n_local"]); A = np.zeros(d["rowptr"][-1])
    times = []
    for it in range(warmup + steps):
        f[:] = 0.0; A[:] = 0.0
        t0 = time.perf_counter()
        orc.evaluate_volume(tm, d["lids"], t, x, None, d["rowptr"], d["colind"], f, A)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ne = d["lids"].shape[0]
    return ne, times


def pick_cpu_sample(threads, budget_s=12.0):
    ne, (t0,) = cpu_reference(24, threads)
    rate = ne / t0
    n = int(round((rate * budget_s) ** (1.0 / 3.0)))
    return max(24, min(n, 96))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = args.cpu_n or pick_cpu_sample(threads, budget_s=max(4.0, 60.0 / max(1, args.steps + args.warmup)))
    ne, times = cpu_reference(n, threads, steps=args.steps, warmup=args.warmup)
    t = sum(times) / len(times)
    val = ne / t / 1e6
    sample = f"{n}^3 = {ne} elements per step (bounded sample of the 256^3 workload), volume fill, OpenMP over worksets of 20"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Melem/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "poisson_q1hex_256^3_residual+jacobian", "elements_per_gpu": args.n ** 3, "timing": "host wall clock"},
           "cpu_baseline": {"value": val, "unit": "Melem/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "CPU restatement of the reference algorithm (oracle/); the reference itself needs Trilinos+MPI and cannot be built here"}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import numpy as np
    import torch
    from tianxin_b200 import capi, host
    from tianxin_b200.assembly_engine import (AssemblyEngine, AssemblyEngineInArgs, LinearObjContainer,
                                              build_poisson_problem)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the assembly path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    comm = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        comm = host.TorchComm(device=dev)
        box = [capi.Handle.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    # weak scaling: n^3 elements per GPU, processor grid as CubeHexMeshFactory's default (:89-133)
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    px, py, pz = grids.get(world, (world, 1, 1))
    n = args.n
    stream = torch.cuda.Stream(device=dev)
    t_setup = time.time()
    prob = build_poisson_problem((n * px, n * py, n * pz), rank=rank, nranks=world, comm=comm, procs=(px, py, pz),
                                 device=local, nccl_uid=uid, stream=stream.cuda_stream,
                                 scatter_mode={"auto": capi.SCATTER_AUTO, "rowtile": capi.SCATTER_ROWTILE,
                                               "atomic": capi.SCATTER_ATOMIC, "rowgather": capi.SCATTER_ROWGATHER}[args.mode])
    t_setup = time.time() - t_setup
    h = prob.handle
    info = h.info()
    gids = prob.dof.getOwnedAndGhostedIndices()
    x_host = host.state_by_gid(gids)
    x = torch.from_numpy(x_host).to(dev)
    f = torch.empty(prob.n_local, dtype=torch.float64, device=dev)
    A = torch.empty(prob.nnz, dtype=torch.float64, device=dev)
    ghosted = LinearObjContainer(x=x, f=f, A=A)
    inargs = AssemblyEngineInArgs(ghostedContainer_=ghosted, container_=ghosted, alpha=0.0, beta=1.0, time=0.0)
    ae = AssemblyEngine(h, capi.JACOBIAN)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms.item())

    fill_ms = []

    def step_full():
        ae.evaluate(inargs, 15)

    def step_full_probe():
        ae.evaluate(inargs, 15)
        fill_ms.append(h.last_fill_ms())        # synchronises: only used outside the headline timing

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total = timed(step_full, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    launches = h.info().kernel_launches_last_evaluate * args.steps
    ms_step = ms_total / args.steps
    n_elems_total = prob.n_cells * world
    value = n_elems_total / ms_step / 1e3

    # volume fill only (flags = 2) and the fill kernel alone (CUDA events inside the library, same stream)
    ms_vol = timed(lambda: ae.evaluate(inargs, 2), args.steps, 1) / args.steps
    for _ in range(max(3, min(args.steps, 10))):
        step_full_probe()
    k_ms = sum(fill_ms) / len(fill_ms)
    peak, peak_src = measured_peak()
    achieved = ALG_BYTES_PER_ELEM * prob.n_cells / (k_ms * 1e-3) / 1e9

    # e2e: host (pinned) x in, f out, through the same public call; A stays on the device for the solver
    xh = torch.from_numpy(x_host).pin_memory()
    fh = torch.empty(prob.n_local, dtype=torch.float64).pin_memory()
    g_host = LinearObjContainer(x=xh, f=fh, A=A)
    in_host = AssemblyEngineInArgs(ghostedContainer_=g_host, container_=g_host, alpha=0.0, beta=1.0, time=0.0)
    ms_e2e = timed(lambda: ae.evaluate(in_host, 15), args.steps, 2) / args.steps
    e2e_val = n_elems_total / ms_e2e / 1e3
    checksum = float(fh.double().abs().sum())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = os.cpu_count() or 1
        ncpu = args.cpu_n or pick_cpu_sample(threads)
        ne_c, times = cpu_reference(ncpu, threads, steps=args.cpu_steps, warmup=1)
        t_c = sum(times) / len(times)
        cpu = {"value": ne_c / t_c / 1e6, "unit": "Melem/s", "cores": threads, "kind": "port",
               "sample": f"{ncpu}^3 = {ne_c} elements per volume fill, mean of {len(times)} fills ({sum(times):.1f} s of CPU work; "
                         "OpenMP over worksets of 20); CPU restatement of the reference algorithm (oracle/)"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": f"poisson_q1hex_{n}^3_per_gpu_residual+jacobian_evaluate_all",
                          "elements_per_gpu": prob.n_cells, "rows_per_gpu": prob.n_local, "nnz_per_gpu": prob.nnz,
                          "proc_grid": [px, py, pz], "scatter_mode": {1: "rowtile", 2: "atomic", 3: "rowgather"}[info.scatter_mode],
                          "flags": "Initialize|VolumetricFill|BoundaryFill|Scatter",
                          "l2": "inputs+outputs (%.1f GB) larger than L2, no flush needed" % ((prob.nnz * 8 + prob.n_local * 40 + prob.n_cells * 32) / 1e9),
                          "tiles": info.n_tiles, "tile_rows": info.tile_rows_max, "tile_cells_max": info.tile_cells_max,
                          "smem_bytes": info.smem_bytes, "ctas_per_sm": info.ctas_per_sm, "affine_cells": info.n_affine_cells,
                          "setup_s": round(t_setup, 2)},
               "volume_fill_only": {"value": n_elems_total / ms_vol / 1e3, "unit": "Melem/s", "ms_per_step": ms_vol},
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": measured_traffic(prob.n_cells, info.scatter_mode), "kernel": "k_fill_uniform+k_fill_rowtile (one fill: uniform tiles, then the tiles on the boundary)" if info.scatter_mode == 1 else "fill",
                            "kernel_ms": k_ms, "bytes_per_element": ALG_BYTES_PER_ELEM, "peak_source": peak_src},
               "e2e": {"value": e2e_val, "unit": "Melem/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": prob.n_local * 8 * world,
                       "d2h_bytes_per_step": prob.n_local * 8 * world,
                       "note": "host pinned x -> evaluate(All) -> host pinned f; Jacobian values stay device resident for the solver",
                       "f_abs_sum": checksum},
               "gpu_launches": launches, "clocks": clocks}
        if cpu:
            out["cpu_baseline"] = cpu
        if args.full_d2h and world == 1:
            Ah = torch.empty(prob.nnz, dtype=torch.float64).pin_memory()
            g2 = LinearObjContainer(x=xh, f=fh, A=Ah)
            in2 = AssemblyEngineInArgs(ghostedContainer_=g2, container_=g2, alpha=0.0, beta=1.0, time=0.0)
            ms_full = timed(lambda: ae.evaluate(in2, 15), max(1, args.steps // 4), 1) / max(1, args.steps // 4)
            out["e2e_full_matrix_d2h"] = {"value": n_elems_total / ms_full / 1e3, "unit": "Melem/s", "ms_per_step": ms_full,
                                          "d2h_bytes_per_step": prob.nnz * 8 + prob.n_local * 8}
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.barrier()
        h.close()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=256, help="elements per axis PER GPU")
    ap.add_argument("--impl", default="txasm", choices=["txasm", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "rowtile", "atomic", "rowgather"])
    ap.add_argument("--cpu-n", type=int, default=0, help="edge of the CPU baseline sample (0 = sized for ~10-20 s)")
    ap.add_argument("--cpu-steps", type=int, default=120, help="volume fills timed for cpu_baseline (about 10-20 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--full-d2h", action="store_true", help="also time e2e with the whole Jacobian copied to the host")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "txasm":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
