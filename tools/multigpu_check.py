"""Multi-GPU parity check, one process per GPU (torchrun).  Every rank assembles its brick through the
C ABI with the NCCL halo import/export; rank 0 assembles the same global mesh on ONE simulated rank with
the CPU oracle and compares the OWNED rows of f and A, GID by GID (values to 1e-12, structure exactly).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/multigpu_check.py --size 6 [--perturb 0.2] [--no-dirichlet]
"""
import argparse, os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tianxin_b200 import capi, host
from tianxin_b200.assembly_engine import AssemblyEngine, AssemblyEngineInArgs, LinearObjContainer, build_poisson_problem


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", dest="n", type=int, default=6, help="elements per axis per GPU")
    ap.add_argument("--perturb", type=float, default=0.0)
    ap.add_argument("--no-dirichlet", action="store_true")
    ap.add_argument("--p2p", type=int, default=1, help="1: halo over peer memory (cudaIpc), 0: NCCL send/recv")
    ap.add_argument("--overlap", type=int, default=1, help="export on a side stream under the uniform-tile kernel")
    ap.add_argument("--edge", type=int, default=1, help="1: boundary lattice tiles through k_fill_edge")
    ap.add_argument("--repeat", type=int, default=3, help="evaluate this many times (epoch flags, buffer reuse)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    comm = host.TorchComm(device=dev)
    box = [capi.Handle.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    grids = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    px, py, pz = grids[world]
    N = (a.n * px, a.n * py, a.n * pz)
    prob = build_poisson_problem(N, rank=rank, nranks=world, comm=comm, procs=(px, py, pz), device=local, nccl_uid=box[0],
                                 perturb=a.perturb, dirichlet=not a.no_dirichlet, p2p=bool(a.p2p))
    h = prob.handle
    h.option_set("export_overlap", a.overlap)
    h.option_set("edge_kernel", a.edge)
    gids = prob.dof.getOwnedAndGhostedIndices()
    # owned part of x only: the ghost tail must come from the halo import
    xh = np.full(prob.n_local, np.nan)
    xh[:prob.n_owned] = host.state_by_gid(gids[:prob.n_owned])
    x = torch.from_numpy(xh).to(dev)
    f = torch.full((prob.n_local,), np.nan, dtype=torch.float64, device=dev)
    A = torch.full((prob.nnz,), np.nan, dtype=torch.float64, device=dev)
    c = LinearObjContainer(x=x, f=f, A=A)
    for rep in range(a.repeat):
        x.copy_(torch.from_numpy(xh).to(dev)); f.fill_(float("nan")); A.fill_(float("nan"))
        torch.cuda.synchronize()
        AssemblyEngine(h, capi.JACOBIAN).evaluate(AssemblyEngineInArgs(c, c, alpha=0.0, beta=1.0, time=0.0), 15)
        h.sync()
    assert h.halo_p2p_status() == 0, "a neighbour timed out"
    ok_import = bool(np.array_equal(x.cpu().numpy(), host.state_by_gid(gids)))
    # functional response: rank-local cell integrals + ncclAllReduce (x now holds the imported ghosts)
    resp = h.response_functional(capi.RESP_L2_ERROR, x, cubature_degree=4)
    pl = dict(prob.plan)
    pl["rowptr"] = np.empty(prob.n_local + 1, np.int64); pl["colind"] = np.empty(prob.nnz, np.int32)
    h.graph_get(pl["rowptr"], pl["colind"])                  # the fill graph as merged on the device
    no = prob.n_owned
    fo = f.cpu().numpy()[:no]
    Av = A.cpu().numpy()
    rows = np.repeat(np.arange(no), np.diff(pl["rowptr"][:no + 1]))
    trip = (gids[rows], pl["col_gids"][pl["colind"][:pl["rowptr"][no]]], Av[:pl["rowptr"][no]])
    node_of_gid = dict(zip(prob.dof.getElementGIDs().ravel().tolist(), prob.mesh.elem_nodes().ravel().tolist()))
    ddof_gids = gids[prob.dirichlet_dofs] if prob.dirichlet_dofs is not None else np.zeros(0, np.int64)
    out = [None] * world
    dist.gather_object(dict(rank=rank, owned=gids[:no], f=fo, trip=trip, node_of_gid=node_of_gid, ok_import=ok_import,
                            ddof=ddof_gids, resp=resp, info=(h.info().scatter_mode, h.info().n_tiles, h.info().uniform_kernel_used, h.info().export_overlapped, a.p2p)), out if rank == 0 else None, dst=0)
    status = 0
    if rank == 0:
        from oracle import oracle as orc
        orc.build()
        (s,), _ = orc.poisson_problem(N, perturb=a.perturb)
        ser_lid_of_node = dict(zip(s["elem_nodes"].ravel().tolist(), s["lids"].ravel().tolist()))
        nodemap = {}
        for o in out:
            nodemap.update(o["node_of_gid"])
        ntot = s["n_local"]
        perm = np.array([ser_lid_of_node[nodemap[g]] for g in range(ntot)])      # gid -> serial lid
        xs = np.zeros(ntot); xs[perm] = host.state_by_gid(np.arange(ntot))
        t = orc.tables_build(s["cell_coords"])
        fs = np.zeros(ntot); As = np.zeros(s["rowptr"][-1])
        orc.evaluate_volume(orc.make_terms(), s["lids"], t, xs, None, s["rowptr"], s["colind"], fs, As)
        # Dirichlet as the reference applies it: on each rank's GHOSTED container before the Export ADD, so a
        # Dirichlet row shared by k ranks ends with diagonal k and f = k (x - v)  
        mult = np.zeros(ntot)
        for o in out:
            mult[o["ddof"]] += 1
        import scipy.sparse as sp
        M = sp.csr_matrix((As, s["colind"], s["rowptr"]), shape=(ntot, ntot)).tolil()
        for g in np.nonzero(mult)[0]:
            l = perm[g]
            M.rows[l] = [l]; M.data[l] = [mult[g]]
            fs[l] = mult[g] * xs[l]
        M = M.tocsr()
        worst_f = worst_A = 0.0
        fscale, Ascale = np.abs(fs).max(), np.abs(As).max()
        for o in out:
            status |= 0 if o["ok_import"] else 1
            worst_f = max(worst_f, np.abs(o["f"] - fs[perm[o["owned"]]]).max() / fscale)
            r, cidx, v = o["trip"]
            ref = np.asarray(M[perm[r], perm[cidx]]).ravel()
            worst_A = max(worst_A, np.abs(v - ref).max() / Ascale)
            # every non-zero of the global rows is present in the fill graph
            nnz_ref = np.diff(M[perm[o["owned"]]].indptr)
            have = np.bincount(np.searchsorted(o["owned"], r, sorter=np.argsort(o["owned"])), weights=(v != 0), minlength=len(o["owned"]))
            if (np.sort(have) < np.sort(nnz_ref) - 1e-9).any():
                status |= 4
        if worst_f > 1e-12 or worst_A > 1e-12:
            status |= 2
        resp_ref = orc.response_functional(2, 1, 4, s["lids"], s["cell_coords"], xs)
        resp_err = max(abs(o["resp"] - resp_ref) for o in out) / abs(resp_ref)
        if resp_err > 1e-12:
            status |= 8
        print(f"multigpu_check world={world} grid={px}x{py}x{pz} N={N} perturb={a.perturb} modes={[o['info'] for o in out]} "
              f"import_ok={all(o['ok_import'] for o in out)} rel_err_f={worst_f:.2e} rel_err_A={worst_A:.2e} rel_err_response={resp_err:.2e} -> {'OK' if status == 0 else 'FAIL %d' % status}",
              flush=True)
    st = torch.tensor([status], device=dev)
    dist.broadcast(st, src=0)
    h.close()
    dist.destroy_process_group()
    sys.exit(int(st.item()))


if __name__ == "__main__":
    main()
