#!/usr/bin/env python
"""bench.py -- the contract benchmark of the assembly hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--impl reference]
  (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one fp64 Jacobian-type evaluation that fills both f and A for the 3-D Poisson Q1-hex block:
AssemblyEngine<Jacobian>::evaluate(in, All) = halo import -> volume fill -> Dirichlet rows -> halo export
(BASELINE.json configs[1]).  Weak scaling: every GPU owns an n^3 brick
(n=256 -> 16.7 M elements per GPU) of a CubeHexMeshFactory mesh split like the reference splits it.

One JSON line on rank 0:
  value        Melem/s over all GPUs with x, coordinates, LIDs, graph resident in HBM (K steps, CUDA events, max over ranks)
  e2e          the same evaluate called with HOST (pinned) x and f buffers, copies inside the timed region
  roofline     algorithmic bytes (288 B/element) / measured duration of the fill kernels vs the measured HBM peak
  stage_timers the five AssemblyEngine stage timers of one evaluate (names of Panzer_AssemblyEngine_impl.hpp:74-118)
  parity_max_rel_err       (N = 1) sampled rows of the 256^3 result against the closed-form stencil (tools/stencil_check.py)
  halo_parity_max_rel_err  (N > 1) a small brick per rank through the same NCCL halo, every owned row against the same closed form
  general_hex  (N = 1) the same mesh with every interior node perturbed by 0.2 h: the general (FP64-bound) path
  cpu_baseline the CPU restatement of the reference algorithm (oracle/) on the host cores
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
GENERAL_FLOP_PER_ELEM = 5.0e3   # general trilinear hex, 2x2x2 Gauss (SURVEY.md sections 7, 8d)
METRIC = "Jacobian+residual assembly Melem/s (Q1 hex 256^3)"
TRAFFIC_FILE = "r2_traffic.json"


def measured_traffic(n_cells, mode):
    """dram__bytes_read+write of the fill kernels from the committed ncu capture (not measured in this run)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
        if t["n_cells"] == n_cells and mode == 1:
            return t["traffic_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


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
                time.sleep(0.005)
        except Exception as e:  # NVML missing: report it instead of inventing clocks
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def bind_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned buffers) on the CPUs NVML reports as local to the GPU: the e2e leg
    moves 136 MB per step and direction over PCIe, and with eight ranks remote-socket pinned memory halves that."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def _median(v):
    s = sorted(v)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference(n_sample, threads, steps=1, warmup=0):
    """Time the CPU restatement of the reference algorithm (oracle/) on n_sample^3 elements: volume fill + Dirichlet
    rows on the six faces, the stages evaluate(All) runs on one rank."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    (d,), _ = orc.poisson_problem(n_sample)
    t = orc.tables_build(d["cell_coords"])                  # cached per-workset tables: setup, untimed (as in the reference)
    x = orc.state_by_gid(np.arange(d["n_local"]))
    xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    ddofs = np.where(np.any((xyz < 1e-12) | (xyz > 1 - 1e-12), axis=1))[0].astype(np.int32)
    dvals = np.zeros(len(ddofs))
    tm = orc.make_terms(nthreads=threads)
    f = np.zeros(d["n_local"]); A = np.zeros(d["rowptr"][-1])
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        f[:] = 0.0; A[:] = 0.0                               # the reference zeroes its containers inside evaluate
        orc.evaluate_volume(tm, d["lids"], t, x, None, d["rowptr"], d["colind"], f, A)
        orc.dirichlet(1, ddofs, dvals, x, f, d["rowptr"], d["colind"], A)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    ne = d["lids"].shape[0]
    return ne, times


def cpu_sizes():
    """32^3 always; 128^3 when the cached tables (5.5 KB/element) fit comfortably, else 96^3."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    return [32, 128 if avail > 48e9 else 96]


def cpu_arm(threads, steps, warmup, sizes):
    res = []
    for n in sizes:
        ne, times = cpu_reference(n, threads, steps=steps, warmup=warmup)
        res.append({"n": n, "elements": ne, "median_ms": _median(times) * 1e3, "min_ms": min(times) * 1e3,
                    "Melem_per_s": ne / _median(times) / 1e6})
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sizes = [args.cpu_n] if args.cpu_n else cpu_sizes()
    res = cpu_arm(threads, max(1, args.steps), args.warmup, sizes)
    top = res[-1]
    n = top["n"]
    sample = (f"{n}^3 = {top['elements']} elements per step (bounded sample of the 256^3 workload), volume fill + Dirichlet rows, "
              f"OpenMP over worksets of 20, median of {max(1, args.steps)} steps")
    out = {"impl": "reference", "metric": METRIC, "value": top["Melem_per_s"], "unit": "Melem/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": top["median_ms"], "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"poisson_q1hex_{n}^3_residual+jacobian_volume+dirichlet (CPU sample; the GPU arm runs 256^3 per GPU)",
                      "elements_per_gpu": top["elements"], "same_config": False, "timing": "host wall clock, median",
                      "sizes": res},
           "cpu_baseline": {"value": top["Melem_per_s"], "unit": "Melem/s", "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": top["Melem_per_s"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "CPU restatement of the reference algorithm (oracle/); the reference itself needs Trilinos+MPI and cannot be built here"}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ closed-form checks
def stencil_parity(prob, x_host, f, A, rows, world, rank, procs, dims_global, comm_gather=None):
    """Sampled owned rows of this rank against tools/stencil_check.py.  Returns max relative errors (A, f)."""
    import numpy as np
    from tools import stencil_check as sc
    h = prob.handle
    NX, NY, NZ = dims_global
    n_local, n_owned = prob.n_local, prob.n_owned
    gids = prob.dof.getOwnedAndGhostedIndices()
    rowptr = np.empty(n_local + 1, np.int64); colind = np.empty(prob.nnz, np.int32)
    h.graph_get(rowptr, colind)                              # (N > 1: the fill graph, merged on the device)
    if world == 1:
        col_gid = gids
        node_of_gid = None                                   # one rank: GID = stk node id - 1 (lexicographic)
    else:
        col_gid = prob.plan["col_gids"]
        node_of_gid = comm_gather()
    def pos_of_gid(g):
        nid = g if node_of_gid is None else node_of_gid[g]
        return np.stack([nid % (NX + 1), (nid // (NX + 1)) % (NY + 1), nid // ((NX + 1) * (NY + 1))], axis=-1).astype(np.int64)
    from tianxin_b200 import host
    px, py, pz = procs
    def ranks_touching(p, n, pr):
        # number of rank intervals (in cells) along one axis whose closed node range contains lattice coordinate p
        base, rem = n // pr, n % pr
        starts = np.array([r * base + min(r, rem) for r in range(pr + 1)])
        return ((p[:, None] >= starts[None, :-1]) & (p[:, None] <= starts[None, 1:])).sum(axis=1)
    def dir_mult(pos):
        bnd = ((pos == 0) | (pos == np.array([NX, NY, NZ])[None, :])).any(axis=1)
        k = ranks_touching(pos[:, 0], NX, px) * ranks_touching(pos[:, 1], NY, py) * ranks_touching(pos[:, 2], NZ, pz)
        return np.where(bnd, k, 0).astype(np.int64)
    import torch
    rows = np.asarray(rows, np.int64)
    # pull only the sampled rows of A and f off the device
    beg = rowptr[rows]; ln = rowptr[rows + 1] - beg
    L = int(ln.max())
    idx = np.minimum(beg[:, None] + np.arange(L)[None, :], rowptr[-1] - 1)
    A_rows = A[torch.from_numpy(idx.ravel()).to(A.device)].cpu().numpy().reshape(len(rows), L)
    f_rows = f[torch.from_numpy(rows).to(f.device)].cpu().numpy()
    # compact views for the checker: row r of the sample occupies [r*L, r*L+len)
    rp = np.arange(len(rows) + 1, dtype=np.int64) * L
    ci = colind[idx].ravel()
    okm = (np.arange(L)[None, :] < ln[:, None]).ravel()
    Av = A_rows.ravel()
    # shrink rows to their true length by marking the padding as absent through a sentinel column of the row itself
    ci = np.where(okm, ci, np.repeat(rows, L))
    Av = np.where(okm, Av, 0.0)
    class _F:                                                # f indexed by sample position
        def __getitem__(self, r):
            return f_rows
    pos_cache = {}
    def pos_of_col(cols):
        return pos_of_gid(col_gid[cols])
    def x_of_col(cols):
        g = col_gid[cols]
        return host.state_by_gid(g)
    # the checker expects LIDs as row ids; feed it per-sample arrays instead
    pos_row = pos_of_col(rows)
    pos_col = np.where(okm.reshape(len(rows), L)[..., None], pos_of_col(ci).reshape(len(rows), L, 3), -1)
    x_col = x_of_col(ci).reshape(len(rows), L)
    A_exp, f_exp = sc.expected_rows(pos_row, pos_col, (NX, NY, NZ), x_col, dir_mult(pos_row))
    A_exp = np.where(okm.reshape(len(rows), L), A_exp, 0.0)
    eA = float(np.abs(Av.reshape(len(rows), L) - A_exp).max() / np.abs(A_exp).max())
    ef = float(np.abs(f_rows - f_exp).max() / np.abs(f_exp).max())
    return eA, ef


def halo_parity_check(args, world, rank, local, comm, uid, procs, stream):
    """N > 1: a small brick per rank through the same code path (NCCL import / fill / Dirichlet / export); every owned
    row of every rank against the closed form; returns the max relative error over ranks."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from tianxin_b200 import capi, host
    from tianxin_b200.assembly_engine import AssemblyEngine, AssemblyEngineInArgs, LinearObjContainer, build_poisson_problem
    px, py, pz = procs
    m = args.check_n
    dims = (m * px, m * py, m * pz)
    dev = torch.device(f"cuda:{local}")
    prob = build_poisson_problem(dims, rank=rank, nranks=world, comm=comm, procs=procs, device=local, nccl_uid=uid,
                                 stream=stream.cuda_stream)
    gids = prob.dof.getOwnedAndGhostedIndices()
    xh = np.full(prob.n_local, np.nan)                       # owned part only: the ghost tail must come from the import
    xh[:prob.n_owned] = host.state_by_gid(gids[:prob.n_owned])
    with torch.cuda.stream(stream):
        x = torch.from_numpy(xh).to(dev)
        f = torch.full((prob.n_local,), float("nan"), dtype=torch.float64, device=dev)
        A = torch.full((prob.nnz,), float("nan"), dtype=torch.float64, device=dev)
        c = LinearObjContainer(x=x, f=f, A=A)
        AssemblyEngine(prob.handle, capi.JACOBIAN).evaluate(AssemblyEngineInArgs(c, c, alpha=0.0, beta=1.0, time=0.0), 15)
    prob.handle.sync()
    torch.cuda.synchronize()

    def gather_nodes():
        mine = np.stack([prob.dof.getElementGIDs().ravel(), prob.mesh.elem_nodes().ravel() - 1], axis=1)
        box = [None] * world
        dist.all_gather_object(box, mine)
        tab = np.concatenate(box, axis=0)
        node_of_gid = np.full(int(tab[:, 0].max()) + 1, -1, np.int64)
        node_of_gid[tab[:, 0]] = tab[:, 1]
        return node_of_gid
    eA, ef = stencil_parity(prob, None, f, A, np.arange(prob.n_owned), world, rank, procs, dims, comm_gather=gather_nodes)
    info = prob.handle.info()
    err = torch.tensor([eA, ef], dtype=torch.float64, device=dev)
    dist.all_reduce(err, op=dist.ReduceOp.MAX)
    prob.handle.close()
    return {"A": float(err[0]), "f": float(err[1]), "elements_per_gpu": m ** 3, "export_overlapped": info.export_overlapped,
            "dirichlet_fused": info.dirichlet_fused, "uniform_kernel_used": info.uniform_kernel_used}


# ------------------------------------------------------------------------------------------------ GPU arm
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
    numa = bind_to_gpu_numa_node(local)
    # the host mirror (mesh, DOF manager, plan negotiation: setup only) gets this rank's share of the cores
    # (torchrun exports OMP_NUM_THREADS=1)
    host_threads = host.set_num_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
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

    halo_parity = None
    if world > 1 and not args.no_check:
        box = [capi.Handle.comm_unique_id() if rank == 0 else None]
        torch.distributed.broadcast_object_list(box, src=0)
        halo_parity = halo_parity_check(args, world, rank, local, comm, box[0], (px, py, pz), stream)

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

    def timed(fn, steps, warmup, per_step=False):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            for i in range(steps):
                fn()
                if per_step or i == steps - 1:
                    ev[i + 1].record(stream)
        barrier()
        ms = torch.tensor([ev[0].elapsed_time(ev[steps])], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        each = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)] if per_step else None
        return float(ms.item()), each

    def step_full():
        ae.evaluate(inargs, 15)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_total, _ = timed(step_full, args.steps, args.warmup)       # the headline: library defaults, one event at either end
    clocks = sampler.stop() if sampler else None
    # The roofline's kernel duration: CUDA events around the fill of every evaluate of a SECOND loop of K back-to-back
    # steps (no sync between them).  Not the headline loop: event records between the kernels cost 1-2 % of a step at
    # N = 1 and, between the forked streams of the overlapped export, 0.1 ms at N = 2 -- so multi-GPU runs take the
    # duration from separate evaluates instead.
    use_ring = world == 1
    fill_hist = []
    if use_ring:
        h.option_set("fill_event_ring", args.steps)
    _, each = timed(step_full, args.steps, 1, per_step=True)      # per-step events as well: median / min of the steps
    if use_ring:
        fill_hist = h.fill_ms_history(args.steps)
        h.option_set("fill_event_ring", 0)
    info_run = h.info()
    launches = info_run.kernel_launches_last_evaluate * args.steps
    ms_step = ms_total / args.steps
    n_elems_total = prob.n_cells * world
    value = n_elems_total / ms_step / 1e3

    # volume fill only (flags = 2), the fill kernels alone (CUDA events inside the library, same stream), stage timers
    ms_vol = timed(lambda: ae.evaluate(inargs, 2), args.steps, 1)[0] / args.steps
    fill_ms, stage = [], None
    h.option_set("stage_timers", 1)                 # (off by default: the stage events cost ~2 % of a step)
    for _ in range(max(3, min(args.steps, 10))):
        ae.evaluate(inargs, 15)
        fill_ms.append(h.last_fill_ms())        # synchronises: only used outside the headline timing
    tm = h.timers()
    stage = {"evaluate_gather": tm.evaluate_gather, "evaluate_volume": tm.evaluate_volume, "evaluate_neumannbcs": tm.evaluate_neumannbcs,
             "evaluate_interfacebcs": tm.evaluate_interfacebcs, "evaluate_dirichletbcs": tm.evaluate_dirichletbcs,
             "evaluate_scatter": tm.evaluate_scatter, "unit": "ms (device time of one evaluate; a fused Dirichlet stage reads 0)"}
    h.option_set("stage_timers", 0)
    k_ms_isolated = _median(fill_ms)
    k_ms = sum(fill_hist) / len(fill_hist) if fill_hist else k_ms_isolated     # average fill duration over the timed region
    peak, peak_src = measured_peak()
    achieved = ALG_BYTES_PER_ELEM * prob.n_cells / (k_ms * 1e-3) / 1e9

    # parity of the graded workload itself (N = 1): sampled rows against the closed-form stencil
    parity = None
    if world == 1 and not args.no_check and info.scatter_mode == 1:
        ae.evaluate(inargs, 15); h.sync()
        rng = np.random.default_rng(7)
        pos_all = None
        rows = rng.integers(0, prob.n_local, size=args.check_rows)
        bnd = prob.dirichlet_dofs[rng.integers(0, len(prob.dirichlet_dofs), size=args.check_rows // 8)]
        rows = np.unique(np.concatenate([rows, bnd, np.arange(64), np.arange(prob.n_local - 64, prob.n_local)]))
        eA, ef = stencil_parity(prob, x_host, f, A, rows, 1, 0, (1, 1, 1), (n, n, n))
        parity = {"A": eA, "f": ef, "rows_checked": int(len(rows)), "against": "closed-form Q1 stencil + 2x2x2 Gauss source load (tools/stencil_check.py)",
                  "tolerance": 1e-12}

    # e2e: host (pinned) x in, f out, through the same public call; A stays on the device for the solver
    xh = torch.from_numpy(x_host).pin_memory()
    fh = torch.empty(prob.n_local, dtype=torch.float64).pin_memory()
    g_host = LinearObjContainer(x=xh, f=fh, A=A)
    in_host = AssemblyEngineInArgs(ghostedContainer_=g_host, container_=g_host, alpha=0.0, beta=1.0, time=0.0)
    ms_e2e = timed(lambda: ae.evaluate(in_host, 15), args.steps, 2)[0] / args.steps
    e2e_val = n_elems_total / ms_e2e / 1e3
    checksum = float(fh.double().abs().sum())
    e2e_full = None
    if world == 1 and not args.no_full_d2h:
        Ah = torch.empty(prob.nnz, dtype=torch.float64).pin_memory()
        g2 = LinearObjContainer(x=xh, f=fh, A=Ah)
        in2 = AssemblyEngineInArgs(ghostedContainer_=g2, container_=g2, alpha=0.0, beta=1.0, time=0.0)
        k2 = max(2, args.steps // 5)
        ms_full = timed(lambda: ae.evaluate(in2, 15), k2, 1)[0] / k2
        e2e_full = {"value": n_elems_total / ms_full / 1e3, "unit": "Melem/s", "ms_per_step": ms_full,
                    "d2h_bytes_per_step": prob.nnz * 8 + prob.n_local * 8, "steps": k2}
        del Ah, g2, in2

    fp64_peak = h.measure_fp64_peak() if rank == 0 else None

    # general hexahedra (N = 1): the same mesh, interior nodes perturbed by 0.2 h
    general = None
    if world == 1 and not args.no_general:
        h.close()
        del A, f, ghosted, inargs, ae, g_host, in_host
        torch.cuda.empty_cache()
        prob2 = build_poisson_problem((n, n, n), device=local, stream=stream.cuda_stream, perturb=0.2)
        h2 = prob2.handle
        f2 = torch.empty(prob2.n_local, dtype=torch.float64, device=dev)
        A2 = torch.empty(prob2.nnz, dtype=torch.float64, device=dev)
        c2 = LinearObjContainer(x=x, f=f2, A=A2)
        in_g = AssemblyEngineInArgs(ghostedContainer_=c2, container_=c2, alpha=0.0, beta=1.0, time=0.0)
        ae2 = AssemblyEngine(h2, capi.JACOBIAN)
        kg = max(3, args.steps // 2)
        ms_g = timed(lambda: ae2.evaluate(in_g, 15), kg, 2)[0] / kg
        i2 = h2.info()
        tf = GENERAL_FLOP_PER_ELEM * prob2.n_cells / (ms_g * 1e-3) / 1e12
        general = {"value": prob2.n_cells / ms_g / 1e3, "unit": "Melem/s", "ms_per_step": ms_g, "steps": kg,
                   "workload": f"poisson_q1hex_{n}^3 perturbed 0.2 h (general trilinear hexahedra, full 2x2x2 rule per cell)",
                   "affine_cells": i2.n_affine_cells, "bound": "fp64", "flop_per_element": GENERAL_FLOP_PER_ELEM,
                   "fp64_tflops_achieved": tf, "fp64_tflops_peak_measured": fp64_peak, "fp64_frac": tf / fp64_peak if fp64_peak else None,
                   "hbm_frac": ALG_BYTES_PER_ELEM * prob2.n_cells / (ms_g * 1e-3) / 1e9 / peak}
        h2.close()

    # the other BASELINE.json configs (3-5) through the general element blocks: one volume fill each (tools/bench_blocks.py)
    other = None
    if world == 1 and not args.no_blocks:
        from tools.bench_blocks import run_blocks
        torch.cuda.empty_cache()
        other = run_blocks(steps=max(3, args.steps // 4), peak_gbs=peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))     # the CPU leg uses every core of the box
        except Exception:
            pass
        threads = os.cpu_count() or 1
        sizes = [args.cpu_n] if args.cpu_n else cpu_sizes()
        res = cpu_arm(threads, 5, 1, sizes)
        top = res[-1]
        cpu = {"value": top["Melem_per_s"], "unit": "Melem/s", "cores": threads, "kind": "port",
               "sample": f"{top['n']}^3 = {top['elements']} elements per evaluate (volume fill + Dirichlet), median of 5 after 1 warm-up; "
                         "OpenMP over worksets of 20; CPU restatement of the reference algorithm (oracle/)",
               "sizes": res}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": "Melem/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "ms_per_step_median": _median(each), "ms_per_step_min": min(each),
               "ms_each_step_rank0": [round(v, 4) for v in each],
               "per_step_note": "median / min / each: a second loop of K steps with an event after every step (and around every fill at N = 1); ms_per_step is the uninstrumented loop",
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": f"poisson_q1hex_{n}^3_per_gpu_residual+jacobian_evaluate_all",
                          "elements_per_gpu": prob.n_cells, "rows_per_gpu": prob.n_local, "nnz_per_gpu": prob.nnz,
                          "proc_grid": [px, py, pz], "scatter_mode": {1: "rowtile", 2: "atomic", 3: "rowgather"}[info.scatter_mode],
                          "flags": "Initialize|VolumetricFill|BoundaryFill|Scatter",
                          "l2": "inputs+outputs (%.1f GB) larger than L2, no flush needed" % ((prob.nnz * 8 + prob.n_local * 40 + prob.n_cells * 32) / 1e9),
                          "tiles": info.n_tiles, "uniform_tiles": info.n_uniform_tiles, "brick_tiles": info.n_brick_tiles,
                          "tile_rows": info.tile_rows_max, "tile_cells_max": info.tile_cells_max,
                          "uniform_kernel_used": info_run.uniform_kernel_used, "dirichlet_fused": info_run.dirichlet_fused,
                          "export_overlapped": info_run.export_overlapped, "ctas_per_sm": info_run.ctas_per_sm,
                          "affine_cells": info.n_affine_cells, "setup_s": round(t_setup, 2), "txasm_setup_ms": round(info.setup_ms, 1),
                          "cpus_bound_to_gpu_numa_node": numa},
               "volume_fill_only": {"value": n_elems_total / ms_vol / 1e3, "unit": "Melem/s", "ms_per_step": ms_vol},
               "stage_timers": stage,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": measured_traffic(prob.n_cells, info.scatter_mode),
                            "traffic_source": f"profiles/{TRAFFIC_FILE} (ncu capture of the same kernels and workload, not measured in this run)",
                            "kernel": "k_fill_edge (lattice tiles on the domain boundary) + k_fill_brick (interior lattice tiles) + k_fill_rowtile (remaining tiles): one fill"
                                      if info_run.uniform_kernel_used == 2 else "fill kernels of one evaluate",
                            "kernel_ms": k_ms, "kernel_ms_isolated": k_ms_isolated,
                            "kernel_ms_source": "mean of the fill's CUDA-event span over the timed steps" if fill_hist else "median of separate evaluates (each followed by a sync)", "bytes_per_element": ALG_BYTES_PER_ELEM, "peak_source": peak_src,
                            "frac_of_nominal_8TBs": achieved / 8000.0},
               "e2e": {"value": e2e_val, "unit": "Melem/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": prob.n_local * 8 * world,
                       "d2h_bytes_per_step": prob.n_local * 8 * world,
                       "note": "host pinned x -> evaluate(All) -> host pinned f; Jacobian values stay device resident for the solver",
                       "f_abs_sum": checksum},
               "gpu_launches": launches, "clocks": clocks, "fp64_peak_tflops_measured": fp64_peak}
        if parity:
            out["parity_max_rel_err"] = parity
        if halo_parity:
            out["halo_parity_max_rel_err"] = halo_parity
        if e2e_full:
            out["e2e_full_matrix_d2h"] = e2e_full
        if general:
            out["general_hex"] = general
        if other:
            out["other_configs"] = other
        if cpu:
            out["cpu_baseline"] = cpu
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
    ap.add_argument("--cpu-n", type=int, default=0, help="edge of the CPU baseline sample (0 = 32^3 and 128^3 / 96^3)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-general", action="store_true", help="skip the perturbed-mesh (general hexahedra) block")
    ap.add_argument("--no-blocks", action="store_true", help="skip the general element blocks (configs 3-5)")
    ap.add_argument("--no-check", action="store_true", help="skip the closed-form parity checks")
    ap.add_argument("--no-full-d2h", action="store_true", help="skip e2e with the whole Jacobian copied to the host")
    ap.add_argument("--check-rows", type=int, default=40000, help="rows sampled for parity_max_rel_err")
    ap.add_argument("--check-n", type=int, default=32, help="elements per axis per GPU of the N>1 halo parity brick")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "txasm":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
