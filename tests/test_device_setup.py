"""Device-side mesh tables and DOF numbering (tianxin_b200/device_setup.py) against the host mirror, bit for bit.
The code is device-agnostic, so these run on the CPU: one rank directly, several ranks over gloo."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from tianxin_b200 import host
from tianxin_b200.device_setup import DeviceDOFManager, DeviceMesh


def _host_problem(n, procs, nf):
    P = procs[0] * procs[1] * procs[2]
    fac = host.CubeHexMeshFactory(**{"X Elements": n[0], "Y Elements": n[1], "Z Elements": n[2],
                                     "X Procs": procs[0], "Y Procs": procs[1], "Z Procs": procs[2]})
    meshes = [fac.buildMesh(r, P) for r in range(P)]
    dms = []
    for r, m in enumerate(meshes):
        dm = host.DOFManager(r, P)
        dm.setConnManager(m.getConnectivity())
        for f in range(nf):
            dm.addField(f"f{f}")
        dms.append(dm)
    host.DOFManager.buildGlobalUnknownsSim(dms)
    return fac, meshes, dms


def _same(dev_mesh, dev_dof, m, dm):
    assert np.array_equal(dev_mesh.elem_ids(), m.elem_ids())
    assert np.array_equal(dev_mesh.elem_nodes(), m.elem_nodes())
    assert np.array_equal(dev_mesh.getConnectivity(), m.getConnectivity())
    assert np.array_equal(dev_mesh.cell_vertex_coordinates(), m.cell_vertex_coordinates())       # bit-exact coordinates
    assert dev_mesh.proc_grid() == m.proc_grid()
    assert np.array_equal(dev_dof.getElementGIDs(), dm.getElementGIDs())
    assert np.array_equal(dev_dof.getOwnedIndices(), dm.getOwnedIndices())
    assert np.array_equal(dev_dof.getGhostedIndices(), dm.getGhostedIndices())
    assert np.array_equal(dev_dof.getGhostedOwners(), dm.getGhostedOwners())
    assert np.array_equal(dev_dof.getLIDs(), dm.getLIDs())
    assert dev_dof.getLIDs().dtype == np.int32 and dev_dof.num_owned == dm.num_owned and dev_dof.num_local == dm.num_local


@pytest.mark.parametrize("n,nf", [((4, 4, 4), 1), ((7, 3, 5), 1), ((3, 4, 2), 3)])
def test_one_rank_matches_host_mirror(n, nf):
    fac, meshes, dms = _host_problem(n, (1, 1, 1), nf)
    dm_ = DeviceMesh(fac, 0, 1, device="cpu")
    dd = DeviceDOFManager(0, 1)
    dd.setConnManager(dm_.connectivity_t())
    for f in range(nf):
        dd.addField(f"f{f}")
    dd.buildGlobalUnknowns()
    _same(dm_, dd, meshes[0], dms[0])
    with pytest.raises(host.TxhostError):
        dd.buildGlobalUnknowns()
    with pytest.raises(host.TxhostError):
        dd.addField("late")


def test_coordinates_snap_to_zero_like_the_reference():
    """Panzer_STK_MeshFactory.hpp:161-168: a coordinate that cancels against x0 within rounding is exactly 0."""
    fac = host.CubeHexMeshFactory(**{"X Elements": 10, "Y Elements": 3, "Z Elements": 3, "X0": -0.3, "Xf": 0.7, "Y0": -1.0, "Yf": 1.0})
    m = fac.buildMesh(0, 1)
    d = DeviceMesh(fac, 0, 1, device="cpu")
    assert np.array_equal(d.cell_vertex_coordinates(), m.cell_vertex_coordinates())
    assert (d.cell_vertex_coordinates() == 0.0).any()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, procs, nf, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fac, meshes, dms = _host_problem(n, procs, nf)            # every rank of the host mirror, in this process
        dmesh = DeviceMesh(fac, rank, world, device="cpu")
        dd = DeviceDOFManager(rank, world)
        dd.setConnManager(dmesh.connectivity_t())
        for f in range(nf):
            dd.addField(f"f{f}")
        dd.buildGlobalUnknowns()
        _same(dmesh, dd, meshes[rank], dms[rank])
        # the plan negotiation runs off the device-built numbering (compact mode) and matches the host path
        lof_h = host.TpetraLinearObjFactory(dms[rank])
        rp, ci = lof_h.getGhostedGraph()
        lof_d = host.TpetraLinearObjFactory(dd.host_manager())
        no, nl = dd.num_owned, dd.num_local
        lof_d.setGhostRows(rp[no:] - rp[no], ci[rp[no]:rp[nl]])
        lof_d.buildPlans(host.TorchComm())
        lof_h.buildPlans(host.TorchComm())
        pd, ph = lof_d.plan(), lof_h.plan()
        for k in ("nbr_rank", "send_off", "send_lids", "recv_off", "recv_lids", "col_gids", "mat_recv_off", "pair_rows", "pair_cols"):
            assert np.array_equal(pd[k], ph[k]), k
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "".join(traceback.format_exception(e))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,procs,nf", [((5, 4, 3), (2, 1, 1), 1), ((4, 5, 6), (2, 2, 1), 2), ((4, 4, 4), (2, 2, 2), 1)])
def test_several_ranks_over_gloo_match_host_mirror(n, procs, nf):
    world = procs[0] * procs[1] * procs[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n, procs, nf, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    for r, msg in sorted(res):
        assert msg == "ok", f"rank {r}: {msg}"


@pytest.mark.parametrize("n", [(4, 4, 4), (6, 3, 5)])
def test_perturbation_rule_matches_host_mirror(n):
    """splitmix64 in wrap-around int64 arithmetic, conversion to double through 32-bit halves: bit-exact coordinates."""
    fac = host.CubeHexMeshFactory(**{"X Elements": n[0], "Y Elements": n[1], "Z Elements": n[2]})
    m = fac.buildMesh(0, 1); m.perturb(0.2)
    d = DeviceMesh(fac, 0, 1, device="cpu"); d.perturb(0.2)
    assert np.array_equal(d.cell_vertex_coordinates(), m.cell_vertex_coordinates())
    m0 = fac.buildMesh(0, 1)
    assert not np.array_equal(m.cell_vertex_coordinates(), m0.cell_vertex_coordinates())
    # the same node gets the same displacement from every rank's brick
    fac2 = host.CubeHexMeshFactory(**{"X Elements": n[0], "Y Elements": n[1], "Z Elements": n[2], "X Procs": 2, "Y Procs": 1, "Z Procs": 1})
    for r in range(2):
        mh = fac2.buildMesh(r, 2); mh.perturb(0.2)
        dh = DeviceMesh(fac2, r, 2, device="cpu"); dh.perturb(0.2)
        assert np.array_equal(dh.cell_vertex_coordinates(), mh.cell_vertex_coordinates())
