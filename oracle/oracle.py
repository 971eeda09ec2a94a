"""ctypes binding of the CPU oracle (oracle/txoracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module; nothing under tianxin_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libtxoracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/txoracle.c -> oracle/_build/libtxoracle.so (gcc, OpenMP)."""
    src = [os.path.join(_HERE, f) for f in ("txoracle.c", "txblocks.c", "txoracle.h", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class MeshParams(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("px", C.c_int), ("py", C.c_int), ("pz", C.c_int),
                ("x0", C.c_double), ("xf", C.c_double), ("y0", C.c_double),
                ("yf", C.c_double), ("z0", C.c_double), ("zf", C.c_double)]


class Tables(C.Structure):
    _fields_ = [("ne", C.c_int64)] + [(n, C.c_void_p) for n in
                ("jac", "jac_inv", "jac_det", "wm", "ip", "basis", "wbasis", "gbasis", "wgbasis")]


class Terms(C.Structure):
    _fields_ = [("eval_type", C.c_int), ("workset_size", C.c_int),
                ("alpha", C.c_double), ("beta", C.c_double), ("kappa", C.c_double),
                ("mass_dot", C.c_double), ("react", C.c_double), ("source_mult", C.c_double),
                ("source_id", C.c_int), ("nthreads", C.c_int), ("gamma", C.c_double), ("mass_dotdot", C.c_double),
                ("fm_grad", C.c_void_p), ("fm_mass", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.orc_mesh_num_elems.restype = C.c_int64
        L.orc_dofs_create.restype = C.c_void_p
        for n in ("orc_dofs_num_elems", "orc_dofs_num_owned", "orc_dofs_num_ghosted"):
            getattr(L, n).restype = C.c_int64
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.orc_dofs_destroy.argtypes = [C.c_void_p]
        L.orc_dofs_set_conn.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
        L.orc_dofs_build.argtypes = [C.c_void_p]
        L.orc_dofs_gids_per_elem.argtypes = [C.c_void_p]
        for n in ("orc_dofs_get_elem_gids", "orc_dofs_get_elem_lids", "orc_dofs_get_owned", "orc_dofs_get_ghosted"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_dofs_field_offsets.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_ghosted_graph.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_tables_build.argtypes = [C.c_int64, C.c_void_p, C.POINTER(Tables)]
        L.orc_evaluate_volume.argtypes = [C.POINTER(Terms), C.c_int64, C.c_void_p, C.POINTER(Tables),
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.orc_evaluate_volume2.argtypes = [C.POINTER(Terms), C.c_int64, C.c_void_p, C.POINTER(Tables),
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p]
        L.orc_dirichlet.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_dirichlet_rows_and_columns.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_cload.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_neumann_flux.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_response_functional.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_gauss_legendre.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        L.orc_global_to_ghost.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_ghost_to_global_vec.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- mesh
def default_proc_grid(nranks: int):
    px, py, pz = C.c_int(), C.c_int(), C.c_int()
    lib().orc_default_proc_grid(nranks, C.byref(px), C.byref(py), C.byref(pz))
    return px.value, py.value, pz.value


def mesh_params(n, procs=(1, 1, 1), box=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)) -> MeshParams:
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    return MeshParams(nx, ny, nz, procs[0], procs[1], procs[2], *box)


def mesh_build(p: MeshParams, rank: int = 0):
    """-> elem_ids[ne], elem_nodes[ne,8] (stk ids, 1-based), cell_coords[ne,8,3]"""
    ne = lib().orc_mesh_num_elems(C.byref(p), rank)
    ids = np.empty(ne, np.int64)
    nodes = np.empty((ne, 8), np.int64)
    coords = np.empty((ne, 8, 3), np.float64)
    lib().orc_mesh_build(C.byref(p), rank, _p(ids), _p(nodes), _p(coords))
    return ids, nodes, coords


# --------------------------------------------------------------------------- dofs
class Dofs:
    """DOFManager::buildGlobalUnknowns for `nranks` simulated ranks."""

    def __init__(self, conns, nfields: int = 1):
        self.nranks = len(conns)
        self.ipe = conns[0].shape[1]
        self.nfields = nfields
        self._h = lib().orc_dofs_create(self.nranks, self.ipe, nfields)
        for r, c in enumerate(conns):
            c = np.ascontiguousarray(c, np.int64)
            lib().orc_dofs_set_conn(self._h, r, c.shape[0], _p(c))
        rc = lib().orc_dofs_build(self._h)
        if rc != 0:
            raise RuntimeError(f"orc_dofs_build failed rc={rc}")

    def __del__(self):
        try:
            lib().orc_dofs_destroy(self._h)
        except Exception:
            pass

    @property
    def gpe(self):
        return self.ipe * self.nfields

    def n_owned(self, r): return lib().orc_dofs_num_owned(self._h, r)
    def n_ghosted(self, r): return lib().orc_dofs_num_ghosted(self._h, r)
    def n_local(self, r): return self.n_owned(r) + self.n_ghosted(r)

    def elem_gids(self, r):
        ne = lib().orc_dofs_num_elems(self._h, r)
        out = np.empty((ne, self.gpe), np.int64)
        lib().orc_dofs_get_elem_gids(self._h, r, _p(out))
        return out

    def elem_lids(self, r):
        ne = lib().orc_dofs_num_elems(self._h, r)
        out = np.empty((ne, self.gpe), np.int32)
        lib().orc_dofs_get_elem_lids(self._h, r, _p(out))
        return out

    def owned(self, r):
        out = np.empty(self.n_owned(r), np.int64)
        lib().orc_dofs_get_owned(self._h, r, _p(out))
        return out

    def ghosted(self, r):
        out = np.empty(self.n_ghosted(r), np.int64)
        lib().orc_dofs_get_ghosted(self._h, r, _p(out))
        return out

    def owned_and_ghosted(self, r):
        return np.concatenate([self.owned(r), self.ghosted(r)])

    def field_offsets(self, field):
        out = np.empty(self.ipe, np.int32)
        lib().orc_dofs_field_offsets(self._h, field, _p(out))
        return out

    def global_to_ghost(self, x_owned, r):
        arr = (C.c_void_p * self.nranks)(*[x.ctypes.data for x in x_owned])
        out = np.empty(self.n_local(r), np.float64)
        rc = lib().orc_global_to_ghost(self._h, arr, r, _p(out))
        assert rc == 0
        return out

    def ghost_to_global_vec(self, f_ghosted, r):
        arr = (C.c_void_p * self.nranks)(*[x.ctypes.data for x in f_ghosted])
        out = np.empty(self.n_owned(r), np.float64)
        lib().orc_ghost_to_global_vec(self._h, arr, r, _p(out))
        return out


# --------------------------------------------------------------------------- graph
def ghosted_graph(lids: np.ndarray, n_rows: int):
    lids = np.ascontiguousarray(lids, np.int32)
    ne, npe = lids.shape
    rowptr = np.empty(n_rows + 1, np.int64)
    lib().orc_ghosted_graph(ne, npe, _p(lids), n_rows, _p(rowptr), None)
    colind = np.empty(rowptr[-1], np.int32)
    lib().orc_ghosted_graph(ne, npe, _p(lids), n_rows, _p(rowptr), _p(colind))
    return rowptr, colind


# --------------------------------------------------------------------------- tables / pipeline
@dataclass
class TableArrays:
    jac: np.ndarray
    jac_inv: np.ndarray
    jac_det: np.ndarray
    wm: np.ndarray
    ip: np.ndarray
    basis: np.ndarray
    wbasis: np.ndarray
    gbasis: np.ndarray
    wgbasis: np.ndarray
    _c: Tables = None


def ref_cubature():
    pts = np.empty((8, 3)); wts = np.empty(8)
    lib().orc_ref_cubature(_p(pts), _p(wts))
    return pts, wts


def ref_basis(pt):
    pt = np.ascontiguousarray(pt, np.float64)
    val = np.empty(8); grad = np.empty((8, 3))
    lib().orc_ref_basis(_p(pt), _p(val), _p(grad))
    return val, grad


def tables_build(cell_coords: np.ndarray) -> TableArrays:
    X = np.ascontiguousarray(cell_coords, np.float64)
    ne = X.shape[0]
    t = TableArrays(np.empty((ne, 8, 3, 3)), np.empty((ne, 8, 3, 3)), np.empty((ne, 8)), np.empty((ne, 8)),
                    np.empty((ne, 8, 3)), np.empty((ne, 8, 8)), np.empty((ne, 8, 8)),
                    np.empty((ne, 8, 8, 3)), np.empty((ne, 8, 8, 3)))
    c = Tables(ne, *[getattr(t, n).ctypes.data for n in
                     ("jac", "jac_inv", "jac_det", "wm", "ip", "basis", "wbasis", "gbasis", "wgbasis")])
    lib().orc_tables_build(ne, _p(X), C.byref(c))
    t._c = c
    return t


def make_terms(eval_type=1, alpha=0.0, beta=1.0, kappa=1.0, mass_dot=0.0, react=0.0,
               source_mult=-1.0, source_id=1, workset_size=20, nthreads=1, gamma=0.0, mass_dotdot=0.0,
               fm_grad=None, fm_mass=None) -> Terms:
    """fm_grad / fm_mass: numpy [ne][8] field multipliers at the integration points (kept alive by the caller)."""
    return Terms(eval_type, workset_size, alpha, beta, kappa, mass_dot, react, source_mult, source_id, nthreads,
                 gamma, mass_dotdot, None if fm_grad is None else fm_grad.ctypes.data, None if fm_mass is None else fm_mass.ctypes.data)


def evaluate_volume(terms: Terms, lids, tables: TableArrays, x, xdot, rowptr, colind, f, A, xdotdot=None):
    """Accumulates into f (and A when eval_type==1).  All arrays numpy, C-contiguous."""
    lids = np.ascontiguousarray(lids, np.int32)
    assert lids.shape[1] == 8
    n_rows = rowptr.shape[0] - 1
    rc = lib().orc_evaluate_volume2(C.byref(terms), lids.shape[0], _p(lids), C.byref(tables._c),
                                    _p(x), _p(xdot), _p(xdotdot), n_rows, _p(rowptr), _p(colind), _p(f), _p(A))
    assert rc == 0


def dirichlet(eval_type, local_dofs, values, x, f, rowptr, colind, A):
    local_dofs = np.ascontiguousarray(local_dofs, np.int32)
    values = np.ascontiguousarray(values, np.float64)
    lib().orc_dirichlet(eval_type, local_dofs.shape[0], _p(local_dofs), _p(values), _p(x), _p(f),
                        _p(rowptr), _p(colind), _p(A))


def dirichlet_rows_and_columns(local_dofs, rowptr, colind, A):
    """Jacobian with f == null: rows to identity and columns zeroed (the eigenvalue path)."""
    local_dofs = np.ascontiguousarray(local_dofs, np.int32)
    assert lib().orc_dirichlet_rows_and_columns(local_dofs.shape[0], _p(local_dofs), rowptr.shape[0] - 1, _p(rowptr), _p(colind), _p(A)) == 0


def cload(eval_type, local_dofs, values, f):
    local_dofs = np.ascontiguousarray(local_dofs, np.int32)
    values = np.ascontiguousarray(values, np.float64)
    lib().orc_cload(eval_type, local_dofs.shape[0], _p(local_dofs), _p(values), _p(f))


def gauss_legendre(n):
    x = np.empty(n); w = np.empty(n)
    assert lib().orc_gauss_legendre(n, _p(x), _p(w)) == 0
    return x, w


def response_functional(kind, solution_id, cub_degree, lids, cell_coords, x):
    lids = np.ascontiguousarray(lids, np.int32); cc = np.ascontiguousarray(cell_coords, np.float64)
    x = np.ascontiguousarray(x, np.float64)
    out = C.c_double()
    rc = lib().orc_response_functional(kind, solution_id, cub_degree, lids.shape[0], _p(lids), _p(cc), _p(x), C.byref(out))
    assert rc == 0
    return out.value


def response_integral(cellvalue, wm, response_vector):
    """TianXin::Response_Integral<Residual>: returns the value and adds it to response_vector[0]."""
    cellvalue = np.ascontiguousarray(cellvalue, np.float64); wm = np.ascontiguousarray(wm, np.float64)
    out = C.c_double()
    lib().orc_response_integral.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = lib().orc_response_integral(cellvalue.shape[0], cellvalue.shape[1], _p(cellvalue), _p(wm), _p(response_vector), C.byref(out))
    assert rc == 0
    return out.value


def neumann_flux(cells, sides, values, lids, cell_coords, f):
    cells = np.ascontiguousarray(cells, np.int32); sides = np.ascontiguousarray(sides, np.int32)
    values = np.ascontiguousarray(values, np.float64); lids = np.ascontiguousarray(lids, np.int32)
    cc = np.ascontiguousarray(cell_coords, np.float64)
    rc = lib().orc_neumann_flux(cells.shape[0], _p(cells), _p(sides), _p(values), _p(lids), _p(cc), _p(f))
    assert rc == 0


def sideset_sides(p: MeshParams, elem_ids, name):
    """(local cell index, Shards side ordinal) of the inline mesh's sideset `name`
    (Panzer_STK_CubeHexMeshFactory.cpp:615-750: back/front = sides 4/5 at z0/zf, bottom/top = 0/2 at y0/yf,
    left/right = 3/1 at x0/xf)."""
    e = np.asarray(elem_ids, np.int64) - 1
    ix, iy, iz = e % p.nx, (e // p.nx) % p.ny, e // (p.nx * p.ny)
    sel, side = {"left": (ix == 0, 3), "right": (ix == p.nx - 1, 1), "bottom": (iy == 0, 0), "top": (iy == p.ny - 1, 2),
                 "back": (iz == 0, 4), "front": (iz == p.nz - 1, 5)}[name]
    cells = np.nonzero(sel)[0].astype(np.int32)
    return cells, np.full(cells.shape[0], side, np.int32)


# --------------------------------------------------------------------------- convenience
def poisson_problem(n, nranks=1, procs=None, perturb=0.0, box=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)):
    """Mesh + DOFs + graph for the inline cube on `nranks` simulated ranks.

    Returns a list (one dict per rank) with keys: elem_ids, elem_nodes, cell_coords, lids, gids,
    owned, ghosted, rowptr, colind, n_local, n_owned; plus the Dofs object.
    `perturb` moves interior nodes by perturb*h*(u-1/2), u = splitmix64(0x5EED, 3*node+comp)/2^64
    (SURVEY.md section 8d), forcing the general (non-affine) geometry path.
    """
    if procs is None:
        procs = (nranks, 1, 1)
    assert procs[0] * procs[1] * procs[2] == nranks
    p = mesh_params(n, procs, box)
    ranks = []
    conns = []
    for r in range(nranks):
        ids, nodes, coords = mesh_build(p, r)
        if perturb:
            coords = perturb_coords(p, nodes, coords, perturb)
        ranks.append(dict(elem_ids=ids, elem_nodes=nodes, cell_coords=coords))
        conns.append(nodes - 1)        # STKConnManager: one id per node = stk id - 1
    dofs = Dofs(conns, 1)
    for r in range(nranks):
        d = ranks[r]
        d["lids"] = dofs.elem_lids(r)
        d["gids"] = dofs.elem_gids(r)
        d["owned"] = dofs.owned(r)
        d["ghosted"] = dofs.ghosted(r)
        d["n_owned"] = dofs.n_owned(r)
        d["n_local"] = dofs.n_local(r)
        d["rowptr"], d["colind"] = ghosted_graph(d["lids"], d["n_local"])
    return ranks, dofs


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def perturb_coords(p: MeshParams, elem_nodes, cell_coords, amp=0.2):
    NX, NY, NZ = p.nx, p.ny, p.nz
    id0 = (elem_nodes - 1).astype(np.int64)
    ix = id0 % (NX + 1); iy = (id0 // (NX + 1)) % (NY + 1); iz = id0 // ((NX + 1) * (NY + 1))
    interior = (ix > 0) & (ix < NX) & (iy > 0) & (iy < NY) & (iz > 0) & (iz < NZ)
    h = np.array([(p.xf - p.x0) / NX, (p.yf - p.y0) / NY, (p.zf - p.z0) / NZ])
    out = cell_coords.copy()
    with np.errstate(over="ignore"):
        for comp in range(3):
            key = (np.uint64(0x5EED) + (id0.astype(np.uint64) * np.uint64(3) + np.uint64(comp)))
            u = splitmix64(key).astype(np.float64) / 2.0**64
            out[..., comp] += np.where(interior, amp * h[comp] * (u - 0.5), 0.0)
    return out


def state_by_gid(gids: np.ndarray) -> np.ndarray:
    """x[g] = sin(0.37 g) + 1e-3 (g mod 7)   """
    g = gids.astype(np.float64)
    return np.sin(0.37 * g) + 1e-3 * (gids % 7)


# --------------------------------------------------------------------------- element blocks (txblocks.c)
HEX8_C1, HEX27_C2, TET4_C1, TET10_C2, HEX8_HCURL = 1, 2, 3, 4, 5
OP_DIFFUSION, OP_ELASTICITY, OP_CURLCURL = 1, 2, 3


class BlockSpec(C.Structure):
    _fields_ = [("elem", C.c_int), ("op", C.c_int), ("cub_degree", C.c_int), ("eval_type", C.c_int),
                ("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double), ("p", C.c_double * 8)]


def _blocks_lib():
    L = lib()
    if not getattr(L, "_blocks_ready", False):
        L.orb_ref_basis.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_cubature.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orb_evaluate.argtypes = [C.POINTER(BlockSpec), C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orb_q2_hex_lids.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orb_hcurl_hex_lids.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L._blocks_ready = True
    return L


def block_num_basis(elem):
    return _blocks_lib().orb_num_basis(elem)


def block_ref_basis(elem, pt):
    nb = block_num_basis(elem)
    val = np.zeros(nb * (3 if elem == HEX8_HCURL else 1)); der = np.zeros(nb * 3)
    pt = np.ascontiguousarray(pt, np.float64)
    _blocks_lib().orb_ref_basis(elem, _p(pt), _p(val), _p(der))
    return (val.reshape(nb, 3) if elem == HEX8_HCURL else val), der.reshape(nb, 3)


def block_cubature(elem, deg):
    pts = np.zeros((64, 3)); wts = np.zeros(64)
    n = _blocks_lib().orb_cubature(elem, deg, _p(pts), _p(wts))
    assert n > 0
    return pts[:n].copy(), wts[:n].copy()


def block_evaluate(elem, op, cub_degree, params, cell_coords, lids, rowptr, colind, x, xdot=None, xdotdot=None, eval_type=1,
                   alpha=0.0, beta=1.0, gamma=0.0, field_offsets=None, signs=None, f=None, A=None):
    """One element block through the evaluator chain; accumulates into f and A (allocated when None)."""
    sp = BlockSpec(elem, op, cub_degree, eval_type, alpha, beta, gamma, (C.c_double * 8)(*(list(params) + [0.0] * (8 - len(params)))))
    cc = np.ascontiguousarray(cell_coords, np.float64); lids = np.ascontiguousarray(lids, np.int32)
    n_rows = rowptr.shape[0] - 1
    if f is None:
        f = np.zeros(n_rows)
    if A is None and eval_type == 1:
        A = np.zeros(rowptr[-1])
    fo = None if field_offsets is None else np.ascontiguousarray(field_offsets, np.int32)
    sg = None if signs is None else np.ascontiguousarray(signs, np.int8)
    rc = _blocks_lib().orb_evaluate(C.byref(sp), lids.shape[0], _p(cc), lids.shape[1], _p(lids), _p(fo), _p(sg), _p(x), _p(xdot), _p(xdotdot),
                                    _p(rowptr), _p(colind), _p(f), _p(A))
    assert rc == 0, rc
    return f, A


def q2_hex_lids(n):
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    out = np.empty((nx * ny * nz, 27), np.int32)
    _blocks_lib().orb_q2_hex_lids(nx, ny, nz, _p(out))
    return out


def hcurl_hex_lids(n):
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    lids = np.empty((nx * ny * nz, 12), np.int32); signs = np.empty((nx * ny * nz, 12), np.int8)
    _blocks_lib().orb_hcurl_hex_lids(nx, ny, nz, _p(lids), _p(signs))
    return lids, signs


def cube_tet_mesh(n, box=(0.0, 1.0, 0.0, 1.0, 0.0, 1.0)):
    """CubeTetMeshFactory (adapters-stk/src/stk_interface/Panzer_STK_CubeTetMeshFactory.cpp:397-465): every hexahedron of the
    inline cube becomes 12 tetrahedra around a centroid node; tet id = 12 (hex_id - 1) + 1 + i, centroid node id =
    hex_id + (NX+1)(NY+1)(NZ+1).  Returns vertex node ids [ne][4] (0-based) and coordinates [ne][4][3]."""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    p = mesh_params((nx, ny, nz), (1, 1, 1), box)
    ids, nodes, coords = mesh_build(p, 0)
    nn = (nx + 1) * (ny + 1) * (nz + 1)
    faces = [(0, 1, 2, 3), (4, 7, 6, 5), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)]   # outward quads, split in two
    tn, tc = [], []
    cen_xyz = coords.mean(axis=1)
    for (a, b, c, d) in faces:
        for tri in ((a, b, c), (a, c, d)):
            v = np.stack([nodes[:, tri[0]] - 1, nodes[:, tri[1]] - 1, nodes[:, tri[2]] - 1, ids - 1 + nn], axis=1)
            xyz = np.stack([coords[:, tri[0]], coords[:, tri[1]], coords[:, tri[2]], cen_xyz], axis=1)
            tn.append(v); tc.append(xyz)
    tn = np.stack(tn, axis=1).reshape(-1, 4); tc = np.stack(tc, axis=1).reshape(-1, 4, 3)
    # positive orientation (the reference's tets are built with positive Jacobians)
    J = tc[:, 1:] - tc[:, :1]
    neg = np.linalg.det(J) < 0
    tn[neg] = tn[neg][:, [0, 2, 1, 3]]; tc[neg] = tc[neg][:, [0, 2, 1, 3]]
    return tn.astype(np.int64), tc, nn + nx * ny * nz
