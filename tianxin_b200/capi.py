"""ctypes binding of include/txasm.h (libtxasm.so).  No torch types cross this boundary: arrays are
passed as raw addresses (numpy host arrays or torch/CUDA device pointers via .data_ptr())."""
from __future__ import annotations

import ctypes as C
import numpy as np
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TXASM_LIB") or os.path.join(_HERE, "libtxasm.so")   # TXASM_LIB: kernel-variant builds (tools/)

# enums (include/txasm.h)
OK, EINVAL, ECUDA, ENOMEM, ESTATE, ENCCL, EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
TOPO_HEX8, TOPO_HEX27, TOPO_TET4, TOPO_TET10 = 8, 27, 4, 10
BASIS_HGRAD_C1, BASIS_HGRAD_C2, BASIS_HCURL_I1 = 1, 2, 3
OP_DIFFUSION, OP_ELASTICITY, OP_CURLCURL = 1, 2, 3
RESIDUAL, JACOBIAN = 0, 1
FLAG_INITIALIZE, FLAG_VOLUMETRIC_FILL, FLAG_BOUNDARY_FILL, FLAG_SCATTER, FLAG_ALL = 1, 2, 4, 8, 15
SCATTER_AUTO, SCATTER_ROWTILE, SCATTER_ATOMIC, SCATTER_ROWGATHER, SCATTER_GENERIC = 0, 1, 2, 3, 4
TERM_GRADGRAD, TERM_MASS, TERM_SOURCE, TERM_TRANSIENT_MASS = 1, 2, 3, 4
VEC_X, VEC_XDOT, VEC_XDOTDOT = 0, 1, 2
SOURCE_SIN3, SOURCE_CONSTANT, SOURCE_IP_ARRAY = 1, 2, 100
RESP_INTEGRAL, RESP_L2_ERROR, RESP_H1_ERROR = 1, 2, 3

EXPORTS = [
    "txasm_version", "txasm_create", "txasm_destroy", "txasm_last_error", "txasm_block_add",
    "txasm_graph_set", "txasm_graph_build", "txasm_graph_get", "txasm_terms_set", "txasm_dirichlet_set",
    "txasm_setup", "txasm_info_get", "txasm_evaluate", "txasm_sync", "txasm_timers_get",
    "txasm_last_fill_ms", "txasm_fill_ms_history", "txasm_graph_get_rows", "txasm_graph_merge_columns", "txasm_comm_unique_id", "txasm_comm_init", "txasm_halo_set",
    "txasm_halo_set_matrix", "txasm_tile_get", "txasm_cload_set", "txasm_neumann_set", "txasm_response_functional",
    "txasm_option_set", "txasm_option_get", "txasm_measure_fp64_peak",
    "txasm_gblock_add", "txasm_gblock_terms_set", "txasm_response_integral", "txasm_debug_timeline",
    "txasm_halo_p2p_blob_size", "txasm_halo_p2p_export", "txasm_halo_p2p_connect", "txasm_halo_p2p_status",
]


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("stream", C.c_void_p), ("scatter_mode", C.c_int),
                ("affine_tol", C.c_double), ("reserved", C.c_int * 8)]


class BlockDesc(C.Structure):
    _fields_ = [("topology", C.c_int), ("basis", C.c_int), ("cubature_degree", C.c_int), ("n_cells", C.c_int64),
                ("cell_vertex_coords", C.c_void_p), ("n_fields", C.c_int), ("dofs_per_cell", C.c_int), ("lids", C.c_void_p),
                ("field_offsets", C.c_void_p), ("orientation_signs", C.c_void_p)]


class Term(C.Structure):
    _fields_ = [("kind", C.c_int), ("vec", C.c_int), ("multiplier", C.c_double),
                ("source_id", C.c_int), ("ip_values", C.c_void_p), ("gather_seed_index1", C.c_int), ("reserved", C.c_int),
                ("field_multiplier_ip", C.c_void_p)]


class InArgs(C.Structure):
    _fields_ = [("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("time", C.c_double), ("step_size", C.c_double), ("stage_number", C.c_double),
                ("evaluate_transient_terms", C.c_int), ("zero_outputs", C.c_int),
                ("n_gather_seeds", C.c_int), ("gather_seeds", C.c_void_p)]


class Timers(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("evaluate_gather", "evaluate_volume", "evaluate_neumannbcs",
                                          "evaluate_interfacebcs", "evaluate_dirichletbcs", "evaluate_scatter")]


class Info(C.Structure):
    _fields_ = [("n_cells", C.c_int64), ("n_rows", C.c_int64), ("nnz", C.c_int64),
                ("n_affine_cells", C.c_int64), ("n_regular_rows", C.c_int64),
                ("scatter_mode", C.c_int), ("n_tiles", C.c_int), ("tile_rows_max", C.c_int),
                ("tile_cells_max", C.c_int), ("smem_bytes", C.c_int), ("threads_per_cta", C.c_int),
                ("ctas_per_sm", C.c_int), ("kernel_launches_last_evaluate", C.c_int), ("n_sm", C.c_int),
                ("n_uniform_tiles", C.c_int), ("n_brick_tiles", C.c_int), ("uniform_kernel_used", C.c_int),
                ("dirichlet_fused", C.c_int), ("export_overlapped", C.c_int), ("n_edge_tiles", C.c_int), ("reserved_i", C.c_int * 2),
                ("setup_ms", C.c_double)]


class TxasmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"txasm error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load libtxasm.so.  Fails loudly when the extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        P, I, I64, D = C.c_void_p, C.c_int, C.c_int64, C.c_double
        L.txasm_last_error.restype = C.c_char_p
        L.txasm_last_error.argtypes = [P]
        L.txasm_version.argtypes = [C.POINTER(I), C.POINTER(I)]
        L.txasm_create.argtypes = [C.POINTER(Config), C.POINTER(P)]
        L.txasm_destroy.argtypes = [P]
        L.txasm_block_add.argtypes = [P, I, I, I, I64, I, P, P, P, I64]
        L.txasm_graph_set.argtypes = [P, I64, P, P]
        L.txasm_graph_build.argtypes = [P, C.POINTER(I64)]
        L.txasm_graph_get.argtypes = [P, P, P]
        L.txasm_graph_get_rows.argtypes = [P, C.c_int64, C.c_int64, P, P]
        L.txasm_graph_merge_columns.argtypes = [P, C.c_int64, P, P, P, C.POINTER(C.c_int64)]
        L.txasm_terms_set.argtypes = [P, C.POINTER(Term), I]
        L.txasm_dirichlet_set.argtypes = [P, I, P, P]
        L.txasm_cload_set.argtypes = [P, I, P, P]
        L.txasm_neumann_set.argtypes = [P, I, P, P, P]
        L.txasm_response_functional.argtypes = [P, I, I, I, P, P]
        L.txasm_setup.argtypes = [P]
        L.txasm_info_get.argtypes = [P, C.POINTER(Info)]
        L.txasm_evaluate.argtypes = [P, I, I, C.POINTER(InArgs), P, P, P, P, P]
        L.txasm_tile_get.argtypes = [P, I, P, P, P, C.POINTER(I)]
        L.txasm_sync.argtypes = [P]
        L.txasm_timers_get.argtypes = [P, C.POINTER(Timers)]
        L.txasm_last_fill_ms.argtypes = [P, C.POINTER(D)]
        L.txasm_fill_ms_history.argtypes = [P, C.POINTER(D), C.c_int, C.POINTER(C.c_int)]
        L.txasm_comm_unique_id.argtypes = [P]
        L.txasm_comm_init.argtypes = [P, I, I, P]
        L.txasm_halo_set.argtypes = [P, I64, I, P, P, P, P, P]
        L.txasm_halo_set_matrix.argtypes = [P, P, P]
        L.txasm_option_set.argtypes = [P, C.c_char_p, I]
        L.txasm_option_get.argtypes = [P, C.c_char_p, C.POINTER(I)]
        L.txasm_measure_fp64_peak.argtypes = [P, C.POINTER(D)]
        L.txasm_response_integral.argtypes = [P, I, P, P, P]
        L.txasm_debug_timeline.argtypes = [P, P]
        L.txasm_gblock_add.argtypes = [P, C.POINTER(BlockDesc), I64, C.POINTER(I)]
        L.txasm_gblock_terms_set.argtypes = [P, I, I, P, I]
        L.txasm_halo_p2p_export.argtypes = [P, P]
        L.txasm_halo_p2p_connect.argtypes = [P, I, P]
        L.txasm_halo_p2p_status.argtypes = [P, C.POINTER(I)]
        _lib = L
    return _lib


def addr(a):
    """Raw address of a numpy array / torch tensor / int / None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class Handle:
    """RAII wrapper of txasm_handle."""

    def __init__(self, device=0, stream=None, scatter_mode=SCATTER_AUTO, affine_tol=0.0):
        self._h = C.c_void_p()
        cfg = Config(device, stream, scatter_mode, affine_tol)
        rc = lib().txasm_create(C.byref(cfg), C.byref(self._h))
        if rc != OK:
            raise TxasmError(rc, lib().txasm_last_error(None).decode())
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            lib().txasm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != OK:
            raise TxasmError(rc, lib().txasm_last_error(self._h).decode())

    def block_add(self, lids, cell_coords=None, node_coords=None, n_rows=None, n_cells=None,
                  topology=TOPO_HEX8, basis=BASIS_HGRAD_C1, cubature_degree=2, dofs_per_cell=8):
        if n_cells is None:
            n_cells = lids.shape[0]
        self._keep += [lids, cell_coords, node_coords]
        self._ck(lib().txasm_block_add(self._h, topology, basis, cubature_degree, n_cells, dofs_per_cell,
                                       addr(lids), addr(cell_coords), addr(node_coords), n_rows))

    def gblock_add(self, topology, basis, cubature_degree, cell_vertex_coords, lids, n_rows, n_fields=1, field_offsets=None,
                   orientation_signs=None):
        """txasm_gblock_add: one general element block; returns its id."""
        import numpy as np
        fo = None if field_offsets is None else np.ascontiguousarray(field_offsets, np.int32)
        self._keep += [cell_vertex_coords, lids, fo, orientation_signs]
        d = BlockDesc(topology, basis, cubature_degree, lids.shape[0], addr(cell_vertex_coords), n_fields, lids.shape[1], addr(lids),
                      addr(fo), addr(orientation_signs))
        bid = C.c_int()
        self._ck(lib().txasm_gblock_add(self._h, C.byref(d), n_rows, C.byref(bid)))
        return bid.value

    def gblock_terms_set(self, block_id, op, params):
        arr = (C.c_double * len(params))(*params)
        self._ck(lib().txasm_gblock_terms_set(self._h, block_id, op, arr, len(params)))

    def graph_set(self, rowptr, colind):
        self._keep += [rowptr, colind]
        self._ck(lib().txasm_graph_set(self._h, rowptr.shape[0] - 1, addr(rowptr), addr(colind)))

    def graph_build(self):
        nnz = C.c_int64()
        self._ck(lib().txasm_graph_build(self._h, C.byref(nnz)))
        return nnz.value

    def graph_get(self, rowptr, colind):
        self._ck(lib().txasm_graph_get(self._h, addr(rowptr), addr(colind)))

    def graph_get_rows(self, first_row, n_rows):
        """(rowptr rebased to 0, colind) of rows [first_row, first_row + n_rows) as numpy arrays."""
        rp = np.empty(n_rows + 1, np.int64)
        self._ck(lib().txasm_graph_get_rows(self._h, first_row, n_rows, addr(rp), None))
        ci = np.empty(int(rp[-1]), np.int32)
        self._ck(lib().txasm_graph_get_rows(self._h, first_row, n_rows, addr(rp), addr(ci)))
        return rp, ci

    def graph_merge_columns(self, rows, cols):
        """Insert (row, col) pairs into the device graph; returns (their positions in the new A_values, new nnz)."""
        rows = np.ascontiguousarray(rows, np.int32); cols = np.ascontiguousarray(cols, np.int32)
        pos = np.empty(len(rows), np.int64)
        nnz = C.c_int64()
        self._ck(lib().txasm_graph_merge_columns(self._h, len(rows), addr(rows), addr(cols), addr(pos), C.byref(nnz)))
        return pos, nnz.value

    def terms_set(self, terms):
        arr = (Term * len(terms))(*terms)
        self._keep.append(arr)
        self._ck(lib().txasm_terms_set(self._h, arr, len(terms)))

    def dirichlet_set(self, local_dofs, values):
        n = 0 if local_dofs is None else local_dofs.shape[0]
        self._ck(lib().txasm_dirichlet_set(self._h, n, addr(local_dofs), addr(values)))

    def response_functional(self, kind, x, solution_id=SOURCE_SIN3, cubature_degree=10):
        """Integrator_Scalar + Response_Functional: sum over cells (and ranks) of the cell integrals; returns a float."""
        out = C.c_double()
        self._ck(lib().txasm_response_functional(self._h, kind, solution_id, cubature_degree, addr(x), C.byref(out)))
        return out.value

    def response_integral(self, cell_ip_values, response_vector, cubature_degree=2):
        """TianXin::Response_Integral: returns the global value and adds it to response_vector[0] (numpy, host)."""
        out = C.c_double()
        self._ck(lib().txasm_response_integral(self._h, cubature_degree, addr(cell_ip_values), addr(response_vector), C.byref(out)))
        return out.value

    def neumann_set(self, cells, local_sides, values):
        n = 0 if cells is None else cells.shape[0]
        self._ck(lib().txasm_neumann_set(self._h, n, addr(cells), addr(local_sides), addr(values)))

    def cload_set(self, local_dofs, values):
        n = 0 if local_dofs is None else local_dofs.shape[0]
        self._ck(lib().txasm_cload_set(self._h, n, addr(local_dofs), addr(values)))

    def setup(self):
        self._ck(lib().txasm_setup(self._h))

    def info(self) -> Info:
        i = Info()
        self._ck(lib().txasm_info_get(self._h, C.byref(i)))
        return i

    def option_set(self, name, value):
        self._ck(lib().txasm_option_set(self._h, name.encode(), int(value)))

    def option_get(self, name):
        v = C.c_int()
        self._ck(lib().txasm_option_get(self._h, name.encode(), C.byref(v)))
        return v.value

    def evaluate(self, eval_type, x, f, A=None, xdot=None, xdotdot=None, flags=FLAG_ALL,
                 alpha=0.0, beta=1.0, gamma=0.0, time=0.0, zero_outputs=1, evaluate_transient_terms=None,
                 gather_seeds=None):
        if evaluate_transient_terms is None:
            evaluate_transient_terms = xdot is not None
        seeds = None
        if gather_seeds is not None and len(gather_seeds):
            seeds = (C.c_double * len(gather_seeds))(*gather_seeds)
        ia = InArgs(alpha, beta, gamma, time, 0.0, 1.0, 1 if evaluate_transient_terms else 0, zero_outputs,
                    0 if seeds is None else len(gather_seeds), None if seeds is None else C.cast(seeds, C.c_void_p))
        self._ck(lib().txasm_evaluate(self._h, eval_type, flags, C.byref(ia), addr(x), addr(xdot), addr(xdotdot),
                                      addr(f), addr(A)))

    def tile_get(self, tile):
        import numpy as np
        i = self.info()
        rows = np.empty(i.tile_rows_max, np.int32); cells = np.full(i.tile_cells_max, -2, np.int32)
        adjl = np.empty((i.tile_rows_max, 8), np.uint16); n = C.c_int()
        self._ck(lib().txasm_tile_get(self._h, tile, addr(rows), addr(cells), addr(adjl), C.byref(n)))
        return rows, cells[:n.value], adjl

    def sync(self):
        self._ck(lib().txasm_sync(self._h))

    def timers(self) -> Timers:
        t = Timers()
        self._ck(lib().txasm_timers_get(self._h, C.byref(t)))
        return t

    def debug_timeline(self):
        out = (C.c_double * 6)()
        self._ck(lib().txasm_debug_timeline(self._h, out))
        return list(out)

    def measure_fp64_peak(self) -> float:
        d = C.c_double()
        self._ck(lib().txasm_measure_fp64_peak(self._h, C.byref(d)))
        return d.value

    def last_fill_ms(self) -> float:
        d = C.c_double()
        self._ck(lib().txasm_last_fill_ms(self._h, C.byref(d)))
        return d.value

    def fill_ms_history(self, cap: int = 1024):
        """Fill times (ms) of the last evaluates, oldest first (needs option_set("fill_event_ring", R))."""
        buf = (C.c_double * cap)()
        n = C.c_int()
        self._ck(lib().txasm_fill_ms_history(self._h, buf, cap, C.byref(n)))
        return [buf[i] for i in range(n.value)]

    # multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().txasm_comm_unique_id(buf)
        if rc != OK:
            raise TxasmError(rc, lib().txasm_last_error(None).decode())
        return buf.raw

    def comm_init(self, nranks, rank, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        self._ck(lib().txasm_comm_init(self._h, nranks, rank, buf))

    def halo_set(self, n_owned, nbr_rank, send_off, send_lids, recv_off, recv_lids):
        self._ck(lib().txasm_halo_set(self._h, n_owned, len(nbr_rank), addr(nbr_rank), addr(send_off),
                                      addr(send_lids), addr(recv_off), addr(recv_lids)))

    def halo_set_matrix(self, mat_recv_off, mat_recv_pos):
        self._ck(lib().txasm_halo_set_matrix(self._h, addr(mat_recv_off), addr(mat_recv_pos)))


def _p2p_methods():
    def halo_p2p_export(self) -> bytes:
        buf = C.create_string_buffer(lib().txasm_halo_p2p_blob_size())
        self._ck(lib().txasm_halo_p2p_export(self._h, buf))
        return buf.raw

    def halo_p2p_connect(self, blobs):
        """blobs: list of every rank's blob, in rank order."""
        raw = b"".join(blobs)
        buf = C.create_string_buffer(raw, len(raw))
        self._ck(lib().txasm_halo_p2p_connect(self._h, len(blobs), buf))

    def halo_p2p_status(self) -> int:
        v = C.c_int()
        self._ck(lib().txasm_halo_p2p_status(self._h, C.byref(v)))
        return v.value
    Handle.halo_p2p_export, Handle.halo_p2p_connect, Handle.halo_p2p_status = halo_p2p_export, halo_p2p_connect, halo_p2p_status


_p2p_methods()


def poisson_terms(kappa=1.0, source_mult=-1.0, source_id=SOURCE_SIN3, mass_dot=0.0, react=0.0, mass_dotdot=0.0):
    """The Poisson equation set's term list (Example_PoissonEquationSet_impl.hpp:150-195); mass_dotdot adds the
    second-order-in-time mass term on D2XDT2 (an extension, seed gamma: SURVEY.md section 8a quirk)."""
    t = []
    if mass_dotdot:
        t.append(Term(TERM_MASS, VEC_XDOTDOT, mass_dotdot, 0, None))
    if mass_dot:
        t.append(Term(TERM_MASS, VEC_XDOT, mass_dot, 0, None))
    if kappa:
        t.append(Term(TERM_GRADGRAD, VEC_X, kappa, 0, None))
    if react:
        t.append(Term(TERM_MASS, VEC_X, react, 0, None))
    if source_mult and source_id:
        t.append(Term(TERM_SOURCE, VEC_X, source_mult, source_id, None))
    return t
