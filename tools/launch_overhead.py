"""How long does the HOST take to issue one evaluate(All) (all launches, no synchronisation) vs the device time of the step?
torchrun --nproc-per-node N tools/launch_overhead.py   (or plain python for one GPU)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch


def main():
    from tianxin_b200 import capi, host
    from tianxin_b200.assembly_engine import AssemblyEngine, AssemblyEngineInArgs, LinearObjContainer, build_poisson_problem
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
    comm = uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        comm = host.TorchComm(device=dev)
        box = [capi.Handle.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0); uid = box[0]
    host.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    px, py, pz = grids[world]
    stream = torch.cuda.Stream(device=dev)
    prob = build_poisson_problem((n * px, n * py, n * pz), rank=rank, nranks=world, comm=comm, procs=(px, py, pz), device=local,
                                 nccl_uid=uid, stream=stream.cuda_stream)
    h = prob.handle
    x = torch.from_numpy(host.state_by_gid(prob.dof.getOwnedAndGhostedIndices())).to(dev)
    f = torch.zeros(prob.n_local, dtype=torch.float64, device=dev); A = torch.empty(prob.nnz, dtype=torch.float64, device=dev)
    ae = AssemblyEngine(h, 1)
    g = LinearObjContainer(x=x, f=f, A=A)
    ia = AssemblyEngineInArgs(ghostedContainer_=g, container_=g, alpha=0.0, beta=1.0, time=0.0)
    h.option_set("stage_timers", int(os.environ.get("STAGE_TIMERS", "1")))
    for _ in range(5):
        ae.evaluate(ia, 15)
    h.sync(); torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    K = 50
    with torch.cuda.stream(stream):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(K):
            ae.evaluate(ia, 15)
        t1 = time.perf_counter()
        e1.record(stream)
    h.sync(); torch.cuda.synchronize()
    print(f"rank {rank}: host issue {1e3 * (t1 - t0) / K:.3f} ms per evaluate, device {e0.elapsed_time(e1) / K:.3f} ms per step, launches {h.info().kernel_launches_last_evaluate}", flush=True)


if __name__ == "__main__":
    main()
