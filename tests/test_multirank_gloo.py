"""N>1 host logic on CPU: two processes, torch.distributed gloo, 127.0.0.1 rendezvous.
Each rank builds its CubeHexMeshFactory brick, runs the distributed DOFManager (GUN) and the
TpetraLinearObjFactory plan through TorchComm, then performs the Import (INSERT) and Export (ADD) with
point-to-point messages exactly as the plan prescribes."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, procs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tianxin_b200 import host
        comm = host.TorchComm()
        fac = host.CubeHexMeshFactory(**{"X Elements": n[0], "Y Elements": n[1], "Z Elements": n[2],
                                         "X Procs": procs[0], "Y Procs": procs[1], "Z Procs": procs[2]})
        mesh = fac.buildMesh(rank, world)
        dm = host.DOFManager(rank, world)
        dm.setConnManager(mesh.getConnectivity()); dm.addField("TEMPERATURE")
        dm.buildGlobalUnknowns(comm)
        lof = host.TpetraLinearObjFactory(dm)
        lof.buildPlans(comm)
        pl = lof.plan()
        no, nl = dm.num_owned, dm.num_local
        gids = dm.getOwnedAndGhostedIndices()
        # Import: owned values travel to the ranks that ghost them
        x = np.full(nl, np.nan); x[:no] = host.state_by_gid(gids[:no])
        reqs, rbufs = [], []
        for k, nb in enumerate(pl["nbr_rank"]):
            sl = pl["send_lids"][pl["send_off"][k]:pl["send_off"][k + 1]]
            if len(sl):
                reqs.append(dist.isend(torch.from_numpy(x[sl].copy()), int(nb)))
            nr = int(pl["recv_off"][k + 1] - pl["recv_off"][k])
            if nr:
                b = torch.empty(nr, dtype=torch.float64); rbufs.append((k, b)); reqs.append(dist.irecv(b, int(nb)))
        for r in reqs: r.wait()
        for k, b in rbufs:
            x[pl["recv_lids"][pl["recv_off"][k]:pl["recv_off"][k + 1]]] = b.numpy()
        ok_import = bool(np.array_equal(x, host.state_by_gid(gids)))
        # Export ADD: every rank contributes 1 per local DOF -> owned entries count the sharing ranks
        f = np.ones(nl)
        reqs, rbufs = [], []
        for k, nb in enumerate(pl["nbr_rank"]):
            gl = pl["recv_lids"][pl["recv_off"][k]:pl["recv_off"][k + 1]]
            if len(gl):
                reqs.append(dist.isend(torch.from_numpy(f[gl].copy()), int(nb)))
            ns = int(pl["send_off"][k + 1] - pl["send_off"][k])
            if ns:
                b = torch.empty(ns, dtype=torch.float64); rbufs.append((k, b)); reqs.append(dist.irecv(b, int(nb)))
        for r in reqs: r.wait()
        for k, b in rbufs:
            np.add.at(f, pl["send_lids"][pl["send_off"][k]:pl["send_off"][k + 1]], b.numpy())
        allg = [None] * world
        dist.all_gather_object(allg, gids.tolist())
        mult = np.zeros(int(max(max(g) for g in allg)) + 1)
        for g in allg:
            mult[np.array(g)] += 1
        ok_export = bool(np.array_equal(f[:no], mult[gids[:no]]))
        # owned sets partition 0..N-1 and are rank-contiguous (GUN offsets)
        owned_all = [None] * world
        dist.all_gather_object(owned_all, gids[:no].tolist())
        flat = sorted(sum(owned_all, []))
        ok_part = flat == list(range(len(flat))) and all(min(o) > max(owned_all[i - 1]) for i, o in enumerate(owned_all) if i)
        q.put((rank, ok_import, ok_export, ok_part, int(no), int(nl)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,procs", [((4, 3, 3), (2, 1, 1)), ((3, 4, 2), (1, 2, 1))])
def test_world_size_2_gloo(n, procs):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, procs, q)) for r in range(2)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps: p.join(timeout=60)
    assert all(p.exitcode == 0 for p in ps)
    for rank, ok_i, ok_e, ok_p, no, nl in res:
        assert ok_i and ok_e and ok_p, (rank, ok_i, ok_e, ok_p)
    nodes = (n[0] + 1) * (n[1] + 1) * (n[2] + 1)
    assert res[0][4] + res[1][4] == nodes
