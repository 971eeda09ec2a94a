"""Host-side mirror of the reference classes that feed the assembly hot path (ctypes over
libtxhost.so, see include/txhost.h): CubeHexMeshFactory (+STKConnManager), DOFManager,
TpetraLinearObjFactory.  Names, argument meaning and error behaviour follow the reference so the
tests read like the reference's own (adapters-stk/test/stk_connmngr/tCubeHexMeshDOFManager.cpp, ...).

Distributed steps are state machines around one primitive, an all-to-all of int64 records.  Three
movers: `LocalComm` (one rank), `SimComm` (all ranks of a small problem in one process, used by the
CPU tests) and `TorchComm` (torch.distributed, gloo on CPU / nccl on GPU: one process per GPU).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtxhost.so")
_lib = None


class TxhostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        P, I, I64, D = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.txhost_last_error.restype = C.c_char_p
        L.txhost_set_num_threads.argtypes = [I]
        L.txhost_cube_hex_mesh.restype = P
        L.txhost_cube_hex_mesh.argtypes = [I, I, I, I, I, I, D, D, D, D, D, D, I, I]
        L.txhost_mesh_destroy.argtypes = [P]
        L.txhost_cube_hex_brick.argtypes = [I, I, I, I, I, I, I, I, P]
        L.txhost_dofmgr_from_arrays.restype = P
        L.txhost_dofmgr_from_arrays.argtypes = [I, I, I, I, I64, P, I64, P, P, I64]
        L.txhost_mesh_num_elems.restype = I64
        L.txhost_mesh_num_elems.argtypes = [P]
        L.txhost_mesh_proc_grid.argtypes = [P, C.POINTER(I), C.POINTER(I), C.POINTER(I)]
        L.txhost_mesh_get.argtypes = [P, P, P, P]
        L.txhost_mesh_connectivity.argtypes = [P, P]
        L.txhost_mesh_perturb.argtypes = [P, D]
        L.txhost_mesh_sideset_nodes.restype = I64
        L.txhost_mesh_sideset_nodes.argtypes = [P, C.c_char_p, P]
        L.txhost_dofmgr_create.restype = P
        L.txhost_dofmgr_create.argtypes = [I, I, I, I]
        L.txhost_dofmgr_destroy.argtypes = [P]
        L.txhost_dofmgr_set_connectivity.argtypes = [P, I64, P]
        L.txhost_dofmgr_step.argtypes = [P, P, P, C.POINTER(P), C.POINTER(P), C.POINTER(I)]
        for n in ("txhost_dofmgr_num_owned", "txhost_dofmgr_num_ghosted"):
            getattr(L, n).restype = I64
            getattr(L, n).argtypes = [P]
        for n in ("txhost_dofmgr_get_owned", "txhost_dofmgr_get_ghosted", "txhost_dofmgr_get_ghosted_owner",
                  "txhost_dofmgr_get_elem_gids", "txhost_dofmgr_get_elem_lids"):
            getattr(L, n).argtypes = [P, P]
        L.txhost_dofmgr_field_offsets.argtypes = [P, I, P]
        L.txhost_lof_create.restype = P
        L.txhost_lof_create.argtypes = [P]
        L.txhost_lof_destroy.argtypes = [P]
        L.txhost_lof_ghosted_graph.argtypes = [P, C.POINTER(I64)]
        L.txhost_lof_get_ghosted_graph.argtypes = [P, P, P]
        L.txhost_lof_set_ghosted_graph.argtypes = [P, P, P]
        L.txhost_lof_set_ghost_rows.argtypes = [P, P, P]
        L.txhost_lof_num_pairs.restype = I64
        L.txhost_lof_num_pairs.argtypes = [P]
        L.txhost_lof_get_pairs.argtypes = [P, P, P]
        L.txhost_lof_step.argtypes = [P, P, P, C.POINTER(P), C.POINTER(P), C.POINTER(I)]
        L.txhost_lof_num_neighbors.argtypes = [P]
        L.txhost_lof_get_halo.argtypes = [P, P, P, P, P, P]
        L.txhost_lof_halo_sizes.argtypes = [P] + [C.POINTER(I64)] * 5
        L.txhost_lof_get_fill_graph.argtypes = [P, P, P, P]
        L.txhost_lof_get_matrix_plan.argtypes = [P, P, P]
        _lib = L
    return _lib


def _err():
    return TxhostError(lib().txhost_last_error().decode())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the host mirror (torchrun exports OMP_NUM_THREADS=1 to every rank); returns the count in use."""
    return int(lib().txhost_set_num_threads(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------------ comms
class LocalComm:
    rank, size = 0, 1

    def alltoallv(self, counts, buf):
        return counts.copy(), buf.copy()


class TorchComm:
    """torch.distributed mover (backend gloo or nccl)."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.device = device

    def alltoallv(self, counts, buf):
        import torch
        dev = self.device or "cpu"
        sc = torch.from_numpy(np.ascontiguousarray(counts)).to(dev)
        rc = torch.empty_like(sc)
        self.dist.all_to_all_single(rc, sc, group=self.group)
        rcl = rc.cpu().tolist()
        sb = torch.from_numpy(np.ascontiguousarray(buf)).to(dev)
        rb = torch.empty(int(sum(rcl)), dtype=torch.int64, device=dev)
        self.dist.all_to_all_single(rb, sb, output_split_sizes=rcl, input_split_sizes=counts.tolist(), group=self.group)
        return np.asarray(rcl, np.int64), rb.cpu().numpy()


def _step(fn, handle, rc, rb):
    sc, sb, done = C.c_void_p(), C.c_void_p(), C.c_int()
    if fn(handle, _p(rc), _p(rb), C.byref(sc), C.byref(sb), C.byref(done)) != 0:
        raise _err()
    return sc, sb, bool(done.value)


def _records(sc, sb, P):
    counts = np.ctypeslib.as_array(C.cast(sc, C.POINTER(C.c_int64)), (P,)).copy()
    n = int(counts.sum())
    buf = np.ctypeslib.as_array(C.cast(sb, C.POINTER(C.c_int64)), (n,)).copy() if n else np.zeros(0, np.int64)
    return counts, buf


def _drive(fn, handle, comm):
    """Run one state machine on this rank, moving records with `comm`."""
    P = comm.size
    rc, rb = np.zeros(P, np.int64), np.zeros(0, np.int64)
    while True:
        sc, sb, done = _step(fn, handle, rc, rb)
        if done:
            return
        counts, buf = _records(sc, sb, P)
        rc, rb = comm.alltoallv(counts, buf)


def _drive_sim(fn, handles):
    """Run the state machines of all ranks in lock-step inside one process (CPU tests)."""
    P = len(handles)
    rcs = [np.zeros(P, np.int64) for _ in range(P)]
    rbs = [np.zeros(0, np.int64) for _ in range(P)]
    while True:
        sends, dones = [], []
        for r in range(P):
            sc, sb, done = _step(fn, handles[r], rcs[r], rbs[r])
            dones.append(done)
            if not done:
                sends.append(_records(sc, sb, P))
        if all(dones):
            return
        assert not any(dones)
        offs = [np.concatenate([[0], np.cumsum(c)]) for c, _ in sends]
        for r in range(P):
            rcs[r] = np.array([sends[s][0][r] for s in range(P)], np.int64)
            rbs[r] = np.concatenate([sends[s][1][offs[s][r]:offs[s][r + 1]] for s in range(P)]) if P else np.zeros(0, np.int64)


# ------------------------------------------------------------------------------------ mesh
class CubeHexMeshFactory:
    """panzer_stk::CubeHexMeshFactory: parameter names as in its ParameterList
    (adapters-stk/src/stk_interface/Panzer_STK_CubeHexMeshFactory.cpp:200-260)."""

    def __init__(self, **pl):
        self.pl = {"X Elements": 5, "Y Elements": 5, "Z Elements": 5, "X Procs": -1, "Y Procs": 1, "Z Procs": 1,
                   "X0": 0.0, "Xf": 1.0, "Y0": 0.0, "Yf": 1.0, "Z0": 0.0, "Zf": 1.0}
        for k, v in pl.items():
            k = k.replace("_", " ")
            if k not in self.pl:
                raise TxhostError(f"CubeHexMeshFactory: unknown parameter \"{k}\"")
            self.pl[k] = v

    def buildMesh(self, rank=0, nranks=1):
        p = self.pl
        h = lib().txhost_cube_hex_mesh(p["X Elements"], p["Y Elements"], p["Z Elements"], p["X Procs"], p["Y Procs"],
                                       p["Z Procs"], p["X0"], p["Xf"], p["Y0"], p["Yf"], p["Z0"], p["Zf"], rank, nranks)
        if not h:
            raise _err()
        return Mesh(h, self.pl)


class Mesh:
    """STK_Interface + STKConnManager view of one rank's part of the mesh."""

    SIDESETS = ("left", "right", "bottom", "top", "back", "front")

    def __init__(self, h, pl):
        self._h, self.pl = h, dict(pl)

    def __del__(self):
        try:
            lib().txhost_mesh_destroy(self._h)
        except Exception:
            pass

    @property
    def num_elems(self):
        return lib().txhost_mesh_num_elems(self._h)

    def proc_grid(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        lib().txhost_mesh_proc_grid(self._h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def perturb(self, amp):
        lib().txhost_mesh_perturb(self._h, amp)

    def elem_ids(self):
        o = np.empty(self.num_elems, np.int64); lib().txhost_mesh_get(self._h, _p(o), None, None); return o

    def elem_nodes(self):
        o = np.empty((self.num_elems, 8), np.int64); lib().txhost_mesh_get(self._h, None, _p(o), None); return o

    def cell_vertex_coordinates(self):
        o = np.empty((self.num_elems, 8, 3), np.float64); lib().txhost_mesh_get(self._h, None, None, _p(o)); return o

    def getConnectivity(self):
        """STKConnManager::buildConnectivity(nodal pattern) -> int64[ne][8]"""
        o = np.empty((self.num_elems, 8), np.int64); lib().txhost_mesh_connectivity(self._h, _p(o)); return o

    def sideset_nodes(self, name):
        n = lib().txhost_mesh_sideset_nodes(self._h, name.encode(), None)
        if n < 0:
            raise _err()
        o = np.empty(n, np.int64)
        lib().txhost_mesh_sideset_nodes(self._h, name.encode(), _p(o))
        return o


# ------------------------------------------------------------------------------------ dofs
class DOFManager:
    """panzer::DOFManager for nodal CG fields (dof-mgr/src/Panzer_DOFManager.cpp)."""

    def __init__(self, rank=0, nranks=1):
        self.rank, self.nranks = rank, nranks
        self._fields, self._conn, self._h = [], None, None

    def __del__(self):
        try:
            if self._h:
                lib().txhost_dofmgr_destroy(self._h)
        except Exception:
            pass

    def setConnManager(self, conn):
        self._conn = np.ascontiguousarray(conn, np.int64)

    def addField(self, name):
        if self._h:
            raise TxhostError("DOFManager::addField: buildGlobalUnknowns has already been called")
        self._fields.append(name)
        return len(self._fields) - 1

    def getFieldNum(self, name):
        return self._fields.index(name)

    def _create(self):
        if self._conn is None or not self._fields:
            raise TxhostError("DOFManager::buildGlobalUnknowns needs a ConnManager and at least one field")
        if self._h:
            raise TxhostError("DOFManager::buildGlobalUnknowns cannot be called again")      # Panzer_DOFManager.cpp:492-494
        self._h = lib().txhost_dofmgr_create(self.rank, self.nranks, self._conn.shape[1], len(self._fields))
        if not self._h:
            raise _err()
        lib().txhost_dofmgr_set_connectivity(self._h, self._conn.shape[0], _p(self._conn))

    def buildGlobalUnknowns(self, comm=None):
        comm = comm or LocalComm()
        assert comm.size == self.nranks and comm.rank == self.rank
        self._create()
        _drive(lib().txhost_dofmgr_step, self._h, comm)

    @staticmethod
    def buildGlobalUnknownsSim(managers):
        """All ranks of a small problem in one process."""
        for m in managers:
            m._create()
        _drive_sim(lib().txhost_dofmgr_step, [m._h for m in managers])

    @property
    def gids_per_elem(self):
        return self._conn.shape[1] * len(self._fields)

    def getOwnedIndices(self):
        o = np.empty(lib().txhost_dofmgr_num_owned(self._h), np.int64); lib().txhost_dofmgr_get_owned(self._h, _p(o)); return o

    def getGhostedIndices(self):
        o = np.empty(lib().txhost_dofmgr_num_ghosted(self._h), np.int64); lib().txhost_dofmgr_get_ghosted(self._h, _p(o)); return o

    def getOwnedAndGhostedIndices(self):
        return np.concatenate([self.getOwnedIndices(), self.getGhostedIndices()])

    def getGhostedOwners(self):
        o = np.empty(lib().txhost_dofmgr_num_ghosted(self._h), np.int32); lib().txhost_dofmgr_get_ghosted_owner(self._h, _p(o)); return o

    def getElementGIDs(self):
        o = np.empty((self._conn.shape[0], self.gids_per_elem), np.int64); lib().txhost_dofmgr_get_elem_gids(self._h, _p(o)); return o

    def getLIDs(self):
        """GlobalIndexer::getLIDs(): int[ne][gpe], LayoutRight"""
        o = np.empty((self._conn.shape[0], self.gids_per_elem), np.int32); lib().txhost_dofmgr_get_elem_lids(self._h, _p(o)); return o

    def getGIDFieldOffsets(self, field):
        o = np.empty(self._conn.shape[1], np.int32)
        lib().txhost_dofmgr_field_offsets(self._h, field if isinstance(field, int) else self.getFieldNum(field), _p(o))
        return o

    @property
    def num_owned(self):
        return lib().txhost_dofmgr_num_owned(self._h)

    @property
    def num_local(self):
        return self.num_owned + lib().txhost_dofmgr_num_ghosted(self._h)


# ------------------------------------------------------------------------------------ lof
class TpetraLinearObjFactory:
    """panzer::TpetraLinearObjFactory: ghosted graph and the Import/Export plans
    (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:124-219, 534-650)."""

    def __init__(self, dofManager: DOFManager):
        self.dof = dofManager
        self._h = lib().txhost_lof_create(dofManager._h)
        if not self._h:
            raise _err()

    def __del__(self):
        try:
            lib().txhost_lof_destroy(self._h)
        except Exception:
            pass

    def getGhostedGraph(self):
        nnz = C.c_int64()
        lib().txhost_lof_ghosted_graph(self._h, C.byref(nnz))
        rowptr = np.empty(self.dof.num_local + 1, np.int64); colind = np.empty(nnz.value, np.int32)
        lib().txhost_lof_get_ghosted_graph(self._h, _p(rowptr), _p(colind))
        return rowptr, colind

    def setGhostedGraph(self, rowptr, colind):
        lib().txhost_lof_set_ghosted_graph(self._h, _p(np.ascontiguousarray(rowptr, np.int64)), _p(np.ascontiguousarray(colind, np.int32)))

    def setGhostRows(self, rowptr, colind):
        """Compact mode: only the ghost rows of the ghosted graph (rowptr rebased to 0); the graph itself stays where it
        was built (the device).  plan() then returns the received (row, column) pairs instead of the fill graph."""
        lib().txhost_lof_set_ghost_rows(self._h, _p(np.ascontiguousarray(rowptr, np.int64)), _p(np.ascontiguousarray(colind, np.int32)))
        self._compact = True

    def buildPlans(self, comm=None):
        _drive(lib().txhost_lof_step, self._h, comm or LocalComm())

    @staticmethod
    def buildPlansSim(lofs):
        _drive_sim(lib().txhost_lof_step, [l._h for l in lofs])

    def plan(self):
        """dict with the halo lists, the fill graph and the matrix export positions."""
        ns, nr, nm, fnnz, ncol = (C.c_int64() for _ in range(5))
        if lib().txhost_lof_halo_sizes(self._h, C.byref(ns), C.byref(nr), C.byref(nm), C.byref(fnnz), C.byref(ncol)) != 0:
            raise _err()
        nn = lib().txhost_lof_num_neighbors(self._h)
        compact = getattr(self, "_compact", False)
        npairs = lib().txhost_lof_num_pairs(self._h)
        out = dict(nbr_rank=np.empty(nn, np.int32), send_off=np.empty(nn + 1, np.int64), send_lids=np.empty(ns.value, np.int32),
                   recv_off=np.empty(nn + 1, np.int64), recv_lids=np.empty(nr.value, np.int32),
                   col_gids=np.empty(ncol.value, np.int64), mat_recv_off=np.empty(nn + 1, np.int64),
                   pair_rows=np.empty(npairs, np.int32), pair_cols=np.empty(npairs, np.int32))
        lib().txhost_lof_get_halo(self._h, _p(out["nbr_rank"]), _p(out["send_off"]), _p(out["send_lids"]),
                                  _p(out["recv_off"]), _p(out["recv_lids"]))
        lib().txhost_lof_get_pairs(self._h, _p(out["pair_rows"]), _p(out["pair_cols"]))
        if compact:
            lib().txhost_lof_get_fill_graph(self._h, None, None, _p(out["col_gids"]))
            lib().txhost_lof_get_matrix_plan(self._h, _p(out["mat_recv_off"]), None)
        else:
            out["rowptr"] = np.empty(self.dof.num_local + 1, np.int64); out["colind"] = np.empty(fnnz.value, np.int32)
            out["mat_recv_pos"] = np.empty(nm.value, np.int64)
            lib().txhost_lof_get_fill_graph(self._h, _p(out["rowptr"]), _p(out["colind"]), _p(out["col_gids"]))
            lib().txhost_lof_get_matrix_plan(self._h, _p(out["mat_recv_off"]), _p(out["mat_recv_pos"]))
        return out


# ------------------------------------------------------------------------------------ synthetic state
def state_by_gid(gids):
    """x[g] = sin(0.37 g) + 1e-3 (g mod 7)"""
    g = np.asarray(gids)
    return np.sin(0.37 * g.astype(np.float64)) + 1e-3 * (g % 7)
