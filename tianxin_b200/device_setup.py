"""Mesh tables and DOF numbering built on the device (the step BEFORE the hot path).

`DeviceMesh` forms the element ids, the nodal connectivity and the vertex coordinates of one rank's brick of a
CubeHexMeshFactory mesh directly in device memory (closed forms of Panzer_STK_CubeHexMeshFactory.cpp:401-461 and
Panzer_STK_MeshFactory.hpp:161-168); `DeviceDOFManager` restates DOFManager::buildGlobalUnknowns
(dof-mgr/src/Panzer_DOFManager.cpp:474-714, GUN :719-865) with device-wide primitives -- sort/unique, searchsorted,
scatter-min, scans -- and moves the directory / owner exchanges with torch.distributed all-to-all (NCCL on GPUs, gloo on the
CPU for the tests).  torch is plumbing here (device arrays, library primitives, collectives); the result is the LID table
the assembly handle reads, left where it was built: the 537 MB table of a 256^3 block never crosses PCIe.

Everything is device-agnostic (device="cpu" works), so `tests/test_device_setup.py` checks it bit for bit against the host
mirror (`tianxin_b200/host.py`, itself checked against the oracle and the reference's golden vectors) without a GPU.
Both classes offer the accessors of their host counterparts (numpy copies on demand), so drivers and checkers do not care
which one built the problem.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import host


def _np(t):
    return t.detach().cpu().numpy()


class DeviceMesh:
    """One rank's part of a CubeHexMeshFactory mesh, arrays on `device`."""

    SIDESETS = host.Mesh.SIDESETS

    def __init__(self, factory: host.CubeHexMeshFactory, rank=0, nranks=1, device="cuda"):
        p = factory.pl
        self.pl = dict(p)
        out = np.zeros(9, np.int64)
        if host.lib().txhost_cube_hex_brick(p["X Elements"], p["Y Elements"], p["Z Elements"], p["X Procs"], p["Y Procs"], p["Z Procs"],
                                            rank, nranks, out.ctypes.data_as(C.c_void_p)) != 0:
            raise host._err()
        xs, xn, ys, yn, zs, zn, px, py, pz = (int(v) for v in out)
        self._grid = (px, py, pz)
        self.brick = (xs, xn, ys, yn, zs, zn)
        self.device = torch.device(device)
        NX, NY, NZ = p["X Elements"], p["Y Elements"], p["Z Elements"]
        dev, i64 = self.device, torch.int64
        # :401-461 buildBlock; local element order = ascending element id (x fastest)
        ex = (xs + torch.arange(xn, device=dev, dtype=i64)).view(1, 1, xn)
        ey = (ys + torch.arange(yn, device=dev, dtype=i64)).view(1, yn, 1)
        ez = (zs + torch.arange(zn, device=dev, dtype=i64)).view(zn, 1, 1)
        self._elem_ids = (NX * NY * ez + NX * ey + ex + 1).reshape(-1)
        n0 = (ex + 1 + ey * (NX + 1) + ez * (NY + 1) * (NX + 1)).reshape(-1)
        plane = (NY + 1) * (NX + 1)
        base = torch.stack([n0, n0 + 1, n0 + 1 + (NX + 1), n0 + (NX + 1)], dim=1)
        self._elem_nodes = torch.cat([base, base + plane], dim=1).contiguous()           # 1-based stk node ids
        self._n = (NX, NY, NZ)
        self._coords = None

    @property
    def num_elems(self):
        return int(self._elem_ids.numel())

    def proc_grid(self):
        return self._grid

    # ---- device tensors
    def elem_ids_t(self):
        return self._elem_ids

    def elem_nodes_t(self):
        return self._elem_nodes

    def connectivity_t(self):
        """STKConnManager nodal pattern: stk node id - 1 (Panzer_STKConnManager.cpp:201-226)"""
        return self._elem_nodes - 1

    def cell_vertex_coordinates_t(self):
        if self._coords is None:
            p = self.pl
            NX, NY, NZ = self._n
            id0 = self._elem_nodes - 1
            ijk = (id0 % (NX + 1), (id0 // (NX + 1)) % (NY + 1), id0 // ((NX + 1) * (NY + 1)))
            out = torch.empty(self.num_elems, 8, 3, dtype=torch.float64, device=self.device)
            for d, (n, lo, hi) in enumerate(((NX, p["X0"], p["Xf"]), (NY, p["Y0"], p["Yf"]), (NZ, p["Z0"], p["Zf"]))):
                delta = (hi - lo) / float(n)
                # Panzer_STK_MeshFactory.hpp:161-168: x = i * delta; val = x + x0, snapped to 0 where the two cancel
                x = ijk[d].to(torch.float64) * delta
                val = x + lo
                snap = (lo * x < 0.0) & ((x.abs() - abs(lo)).abs() < np.finfo(np.float64).eps * abs(lo))
                out[:, :, d] = torch.where(snap, torch.zeros_like(val), val)
            self._coords = out
        return self._coords

    def release_coordinates(self):
        self._coords = None

    def perturb(self, amp):
        """Interior nodes move by amp * h * (u - 1/2) per axis, u = splitmix64(0x5EED + 3 (id - 1) + axis) / 2^64 (SURVEY.md 8d;
        the rule of txhost_mesh_perturb, so the same node gets the same displacement on every rank).  64-bit unsigned
        arithmetic is done in int64 with wrap-around and logical shifts; the value converts to double through its 32-bit halves
        (one rounding, like the C conversion)."""
        NX, NY, NZ = self._n
        p = self.pl
        h = ((p["Xf"] - p["X0"]) / NX, (p["Yf"] - p["Y0"]) / NY, (p["Zf"] - p["Z0"]) / NZ)
        X = self.cell_vertex_coordinates_t()
        id0 = self._elem_nodes - 1
        ix, iy, iz = id0 % (NX + 1), (id0 // (NX + 1)) % (NY + 1), id0 // ((NX + 1) * (NY + 1))
        interior = ~((ix == 0) | (ix == NX) | (iy == 0) | (iy == NY) | (iz == 0) | (iz == NZ))

        def i64(v):                                   # the int64 with the bit pattern of the unsigned constant
            return v - (1 << 64) if v >= (1 << 63) else v

        def lsr(z, n):
            return (z >> n) & ((1 << (64 - n)) - 1)

        for c in range(3):
            z = id0 * 3 + (0x5EED + c)
            z = z + i64(0x9E3779B97F4A7C15)
            z = (z ^ lsr(z, 30)) * i64(0xBF58476D1CE4E5B9)
            z = (z ^ lsr(z, 27)) * i64(0x94D049BB133111EB)
            z = z ^ lsr(z, 31)
            hi, lo = lsr(z, 32), z & 0xFFFFFFFF
            u = (hi.to(torch.float64) * 4294967296.0 + lo.to(torch.float64)) / 18446744073709551616.0
            d = (amp * h[c]) * (u - 0.5)
            X[:, :, c] = torch.where(interior, X[:, :, c] + d, X[:, :, c])

    # ---- the host mirror's accessors (numpy copies)
    def elem_ids(self):
        return _np(self._elem_ids)

    def elem_nodes(self):
        return _np(self._elem_nodes)

    def getConnectivity(self):
        return _np(self.connectivity_t())

    def cell_vertex_coordinates(self):
        return _np(self.cell_vertex_coordinates_t())


class DeviceDOFManager:
    """panzer::DOFManager::buildGlobalUnknowns for nodal CG fields, on the device.

    conn: int64 [ne][ids_per_elem] tensor of global ids (the ConnManager's output).  group: a torch.distributed process group
    (None = the default one) when nranks > 1; the tensors it moves live on conn's device (NCCL) or on the CPU (gloo)."""

    def __init__(self, rank=0, nranks=1):
        self.rank, self.nranks = rank, nranks
        self._fields, self._conn, self._built = [], None, False
        self._host = None

    def setConnManager(self, conn):
        self._conn = conn.to(torch.int64).contiguous()

    def addField(self, name):
        if self._built:
            raise host.TxhostError("DOFManager::addField: buildGlobalUnknowns has already been called")
        self._fields.append(name)
        return len(self._fields) - 1

    def getFieldNum(self, name):
        return self._fields.index(name)

    # ------------------------------------------------------------------ exchange
    def _alltoallv(self, counts, buf, group):
        """counts[P] (python ints), buf int64 tensor grouped by destination -> (recv counts, recv buffer by source)."""
        import torch.distributed as dist
        dev = buf.device
        sc = torch.tensor(counts, dtype=torch.int64, device=dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc, group=group)
        rcl = [int(v) for v in rc.tolist()]
        rb = torch.empty(sum(rcl), dtype=torch.int64, device=dev)
        dist.all_to_all_single(rb, buf.contiguous(), output_split_sizes=rcl, input_split_sizes=list(counts), group=group)
        return rcl, rb

    def buildGlobalUnknowns(self, group=None):
        if self._built:
            raise host.TxhostError("DOFManager::buildGlobalUnknowns cannot be called again")      # Panzer_DOFManager.cpp:492-494
        if self._conn is None or not self._fields:
            raise host.TxhostError("DOFManager::buildGlobalUnknowns needs a ConnManager and at least one field")
        P, me, nf = self.nranks, self.rank, len(self._fields)
        conn = self._conn
        dev = conn.device
        ne, ipe = conn.shape
        i64 = torch.int64
        # overlap map: ascending unique ids of my elements (std::set order, :1261-1290); position of every element id in it
        ov = torch.unique(conn)                                    # sorted
        eov = torch.searchsorted(ov, conn.reshape(-1))
        n_ov = ov.numel()
        if P == 1:
            owner = torch.zeros(n_ov, dtype=i64, device=dev)
        else:
            # ask the directory (rank id % P) who owns each id -- Tpetra::createOneToOne with GreedyTieBreak (:108-133):
            # the smallest rank holding the id.  Requests travel grouped by directory rank, ascending id inside a group.
            dr = ov % P
            order = torch.argsort(dr, stable=True)
            counts = torch.bincount(dr, minlength=P).tolist()
            rcl, rb = self._alltoallv(counts, ov[order], group)
            src = torch.repeat_interleave(torch.arange(P, device=dev, dtype=i64), torch.tensor(rcl, device=dev, dtype=i64))
            uid, inv = torch.unique(rb, return_inverse=True)
            own = torch.full((uid.numel(),), P, dtype=i64, device=dev)
            own.scatter_reduce_(0, inv, src, reduce="amin")
            _, reply = self._alltoallv(rcl, own[inv], group)       # answers come back in request order
            owner = torch.empty(n_ov, dtype=i64, device=dev)
            owner[order] = reply
        mine = owner == me
        n_own_ids = int(mine.sum())
        # my first GID = exclusive scan of the owned counts over the ranks (Teuchos::scan, :794-813)
        if P == 1:
            my_offset = 0
        else:
            import torch.distributed as dist
            cnt = torch.tensor([n_own_ids * nf], dtype=i64, device=dev)
            allc = [torch.empty_like(cnt) for _ in range(P)]
            dist.all_gather(allc, cnt, group=group)
            my_offset = int(sum(int(c) for c in allc[:me]))
        # owned ids are numbered in overlap-map order, fields inner (:823-843)
        gid0 = torch.full((n_ov,), -1, dtype=i64, device=dev)
        gid0[mine] = my_offset + nf * torch.arange(n_own_ids, device=dev, dtype=i64)
        if P > 1:
            # the GIDs of the ids I do not own come from their owners (Import REPLACE, :858)
            theirs = torch.nonzero(~mine).reshape(-1)
            o2 = torch.argsort(owner[theirs], stable=True)
            ask = theirs[o2]
            counts = torch.bincount(owner[theirs], minlength=P).tolist()
            rcl, rb = self._alltoallv(counts, ov[ask], group)
            idx = torch.searchsorted(ov, rb)
            if rb.numel() and (bool((idx >= n_ov).any()) or bool((ov[idx.clamp(max=n_ov - 1)] != rb).any()) or bool((owner[idx.clamp(max=n_ov - 1)] != me).any())):
                raise host.TxhostError("GID request for an id this rank does not own")
            _, reply = self._alltoallv(rcl, gid0[idx], group)
            gid0[ask] = reply
        # elementGIDs_ (:1293-1349): gid0[eov] + field, formed on demand (getElementGIDs)
        f_ar = torch.arange(nf, device=dev, dtype=i64)
        self._gid0, self._eov, self._ne = gid0, eov, ne
        # LIDs: first touch in element order, owned ids first (:580-636), then everything else (:650-695)
        first = torch.full((n_ov,), ne * ipe, dtype=i64, device=dev)
        first.scatter_reduce_(0, eov, torch.arange(ne * ipe, device=dev, dtype=i64), reduce="amin")
        own_idx = torch.nonzero(mine).reshape(-1)
        gh_idx = torch.nonzero(~mine).reshape(-1)
        own_idx = own_idx[torch.argsort(first[own_idx])]
        gh_idx = gh_idx[torch.argsort(first[gh_idx])]
        lid0 = torch.empty(n_ov, dtype=i64, device=dev)
        lid0[own_idx] = nf * torch.arange(own_idx.numel(), device=dev, dtype=i64)
        lid0[gh_idx] = nf * (own_idx.numel() + torch.arange(gh_idx.numel(), device=dev, dtype=i64))
        self._owned = (gid0[own_idx].view(-1, 1) + f_ar.view(1, nf)).reshape(-1)
        self._ghosted = (gid0[gh_idx].view(-1, 1) + f_ar.view(1, nf)).reshape(-1)
        self._ghosted_owner = owner[gh_idx].view(-1, 1).expand(-1, nf).reshape(-1).to(torch.int32)
        self._elids = (lid0[eov].view(ne, ipe, 1) + f_ar.view(1, 1, nf)).reshape(ne, ipe * nf).to(torch.int32).contiguous()
        self._my_offset = my_offset
        self._ipe, self._nf = ipe, nf
        self._conn = None                                          # (1 GB at 256^3: not needed any more)
        self._built = True

    # ------------------------------------------------------------------ results
    @property
    def gids_per_elem(self):
        return self._ipe * self._nf

    @property
    def num_owned(self):
        return int(self._owned.numel())

    @property
    def num_ghosted(self):
        return int(self._ghosted.numel())

    @property
    def num_local(self):
        return self.num_owned + self.num_ghosted

    def lids_t(self):
        """GlobalIndexer::getLIDs() as the device tensor the assembly handle takes (int32 [ne][gpe], LayoutRight)."""
        return self._elids

    def getLIDs(self):
        return _np(self._elids)

    def getElementGIDs(self):
        f_ar = torch.arange(self._nf, device=self._gid0.device, dtype=torch.int64)
        return _np((self._gid0[self._eov].view(self._ne, self._ipe, 1) + f_ar.view(1, 1, self._nf)).reshape(self._ne, self._ipe * self._nf))

    def getOwnedIndices(self):
        return _np(self._owned)

    def getGhostedIndices(self):
        return _np(self._ghosted)

    def getOwnedAndGhostedIndices(self):
        return np.concatenate([self.getOwnedIndices(), self.getGhostedIndices()])

    def getGhostedOwners(self):
        return _np(self._ghosted_owner)

    def getGIDFieldOffsets(self, field):
        f = field if isinstance(field, int) else self.getFieldNum(field)
        return (np.arange(self._ipe) * self._nf + f).astype(np.int32)

    def host_manager(self):
        """A host.DOFManager-compatible object holding the results (owned / ghosted lists only): what
        host.TpetraLinearObjFactory needs for the plan negotiation in compact mode."""
        if self._host is None:
            self._host = _HostView(self)
        return self._host


class _HostView:
    """Finished txhost_dofmgr made from a DeviceDOFManager's results (txhost_dofmgr_from_arrays)."""

    def __init__(self, d: DeviceDOFManager):
        self.rank, self.nranks = d.rank, d.nranks
        owned, ghosted, gown = d.getOwnedIndices(), d.getGhostedIndices(), np.ascontiguousarray(d.getGhostedOwners(), np.int32)
        self.num_owned, self.num_ghosted = len(owned), len(ghosted)
        self.num_local = self.num_owned + self.num_ghosted
        self._h = host.lib().txhost_dofmgr_from_arrays(d.rank, d.nranks, d._ipe, d._nf, len(owned), host._p(owned), len(ghosted),
                                                       host._p(ghosted), host._p(gown), d._my_offset)
        if not self._h:
            raise host._err()

    def __del__(self):
        try:
            host.lib().txhost_dofmgr_destroy(self._h)
        except Exception:
            pass
