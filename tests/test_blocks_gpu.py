"""GPU parity of the general element blocks (BASELINE.json configs 3-5) against the oracle's restatement of the evaluator chain
(oracle/txblocks.c): Q2 hexahedra (generic kernel and the DMMA kernel), P1 / P2 tetrahedra, a mixed hex + tet handle,
three interleaved fields (elastodynamics, second order in time), HCURL edge elements with orientation signs.
PARITY UNPINNED by the reference for all of them; the oracle itself is pinned by identities
(tests/test_oracle_blocks.py).  Tolerance 1e-12 relative to max|A| resp. max|f| (atomic summation order differs)."""
import numpy as np
import pytest
import scipy.sparse as sp

torch = pytest.importorskip("torch")
from tianxin_b200 import capi

pytestmark = pytest.mark.gpu
RTOL = 1e-12
DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["owner", "atomic"])
def block_scatter_mode(request, monkeypatch):
    """every test runs twice: owner-computes gather (default) and the searched-atomic scatter (the literal ScatterResidual)"""
    orig = capi.Handle.setup

    def setup(self):
        orig(self)
        self.option_set("block_atomic", 1 if request.param == "atomic" else 0)
    monkeypatch.setattr(capi.Handle, "setup", setup)
    return request.param


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _run(h, n_rows, nnz, x, xdot=None, xdotdot=None, eval_type=capi.JACOBIAN, **kw):
    t = lambda v: None if v is None else torch.from_numpy(v).to(DEV)
    f = torch.full((n_rows,), np.nan, dtype=torch.float64, device=DEV)
    A = torch.full((nnz,), np.nan, dtype=torch.float64, device=DEV) if eval_type == capi.JACOBIAN else None
    h.evaluate(eval_type, t(x), f, A, xdot=t(xdot), xdotdot=t(xdotdot), flags=capi.FLAG_VOLUMETRIC_FILL, **kw)
    h.sync()
    return f.cpu().numpy(), (None if A is None else A.cpu().numpy())


def _perturbed(cc, amp, seed=0):
    """move every vertex by a smooth field so that shared vertices stay shared: general (non-affine) cells"""
    return cc + amp * np.sin(3.0 * cc[..., [1, 2, 0]] + seed) * np.cos(2.0 * cc[..., [2, 0, 1]])


@pytest.mark.parametrize("dmma", [1, 0])
@pytest.mark.parametrize("n,amp", [(3, 0.0), ((4, 3, 2), 0.03)])
def test_q2_hex_diffusion(oracle, n, amp, dmma):
    (d,), _ = oracle.poisson_problem(n)
    lids = oracle.q2_hex_lids(n)
    n_rows = int(lids.max()) + 1
    rp, ci = oracle.ghosted_graph(lids, n_rows)
    cc = _perturbed(d["cell_coords"], amp)
    rng = np.random.default_rng(1)
    x, xd, xdd = (rng.standard_normal(n_rows) for _ in range(3))
    params = [1.7, 0.3, 0.9, 2.0, -1.25]
    fo, Ao = oracle.block_evaluate(oracle.HEX27_C2, oracle.OP_DIFFUSION, 4, params, cc, lids, rp, ci, x, xd, xdd, alpha=1.5, beta=0.75, gamma=3.0)
    h = capi.Handle()
    b = h.gblock_add(capi.TOPO_HEX27, capi.BASIS_HGRAD_C2, 4, torch.from_numpy(cc).to(DEV), torch.from_numpy(lids).to(DEV), n_rows)
    h.gblock_terms_set(b, capi.OP_DIFFUSION, params)
    h.graph_set(torch.from_numpy(rp).to(DEV), torch.from_numpy(ci).to(DEV))
    h.setup()
    h.option_set("dmma", dmma)
    assert h.info().scatter_mode == capi.SCATTER_GENERIC
    fg, Ag = _run(h, n_rows, len(ci), x, xd, xdd, alpha=1.5, beta=0.75, gamma=3.0)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    fr, _ = _run(h, n_rows, len(ci), x, xd, xdd, eval_type=capi.RESIDUAL, alpha=1.5, beta=0.75, gamma=3.0)
    assert _relerr(fr, fo) < RTOL
    h.close()


def test_tets_and_mixed_hex_tet_handle(oracle):
    """P1 and P2 tetrahedra on the CubeTetMeshFactory split, then Q1 hexahedra (x < 1/2) and P1 tetrahedra (x > 1/2) as two
    blocks of ONE handle sharing the interface nodes (the reference loops over blocks, Panzer_AssemblyEngine_impl.hpp:152)."""
    n = 4
    tn, tc, nnodes = oracle.cube_tet_mesh(n)
    tn = tn.astype(np.int32)
    tc = _perturbed(tc, 0.02)
    rng = np.random.default_rng(2)
    x = rng.standard_normal(nnodes)
    rp, ci = oracle.ghosted_graph(tn, nnodes)
    for deg in (1, 2):
        fo, Ao = oracle.block_evaluate(oracle.TET4_C1, oracle.OP_DIFFUSION, deg, [1.0, 0.5, 0, 0, 2.0], tc, tn, rp, ci, x)
        h = capi.Handle()
        b = h.gblock_add(capi.TOPO_TET4, capi.BASIS_HGRAD_C1, deg, tc, tn, nnodes)                # host arrays
        h.gblock_terms_set(b, capi.OP_DIFFUSION, [1.0, 0.5, 0, 0, 2.0])
        h.graph_set(rp, ci); h.setup()
        fg, Ag = _run(h, nnodes, len(ci), x)
        assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
        h.close()
    # P2
    edge_id = {}
    l10 = np.zeros((len(tn), 10), np.int32); l10[:, :4] = tn
    for c, t in enumerate(tn):
        for e, (i, j) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
            l10[c, 4 + e] = edge_id.setdefault((min(t[i], t[j]), max(t[i], t[j])), nnodes + len(edge_id))
    n10 = nnodes + len(edge_id)
    rp3, ci3 = oracle.ghosted_graph(l10, n10)
    x10 = rng.standard_normal(n10)
    fo, Ao = oracle.block_evaluate(oracle.TET10_C2, oracle.OP_DIFFUSION, 2, [1.0, 0.5], tc, l10, rp3, ci3, x10)
    h = capi.Handle()
    b = h.gblock_add(capi.TOPO_TET10, capi.BASIS_HGRAD_C2, 2, tc, l10, n10)
    h.gblock_terms_set(b, capi.OP_DIFFUSION, [1.0, 0.5]); h.graph_set(rp3, ci3); h.setup()
    fg, Ag = _run(h, n10, len(ci3), x10)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    h.close()
    # mixed handle
    (d,), _ = oracle.poisson_problem(n)
    hex_nodes = (d["elem_nodes"] - 1).astype(np.int32)
    hc = _perturbed(d["cell_coords"], 0.02)
    left = d["cell_coords"].mean(axis=1)[:, 0] < 0.5
    tn0, tc0, _ = oracle.cube_tet_mesh(n)
    right = tc0.mean(axis=1)[:, 0] > 0.5
    hl, tl = np.ascontiguousarray(hex_nodes[left]), np.ascontiguousarray(tn[right])
    rows = np.concatenate([np.repeat(hl, 8, axis=1).ravel(), np.repeat(tl, 4, axis=1).ravel()])
    cols = np.concatenate([np.tile(hl, (1, 8)).ravel(), np.tile(tl, (1, 4)).ravel()])
    G = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(nnodes, nnodes)); G.sum_duplicates(); G.sort_indices()
    rp2, ci2 = G.indptr.astype(np.int64), G.indices.astype(np.int32)
    fo = np.zeros(nnodes); Ao = np.zeros(rp2[-1])
    oracle.block_evaluate(oracle.HEX8_C1, oracle.OP_DIFFUSION, 2, [2.0], hc[left], hl, rp2, ci2, x, f=fo, A=Ao)
    oracle.block_evaluate(oracle.TET4_C1, oracle.OP_DIFFUSION, 1, [0.5], tc[right], tl, rp2, ci2, x, f=fo, A=Ao)
    h = capi.Handle()
    b0 = h.gblock_add(capi.TOPO_HEX8, capi.BASIS_HGRAD_C1, 2, np.ascontiguousarray(hc[left]), hl, nnodes)
    b1 = h.gblock_add(capi.TOPO_TET4, capi.BASIS_HGRAD_C1, 1, np.ascontiguousarray(tc[right]), tl, nnodes)
    h.gblock_terms_set(b0, capi.OP_DIFFUSION, [2.0]); h.gblock_terms_set(b1, capi.OP_DIFFUSION, [0.5])
    h.graph_set(rp2, ci2); h.setup()
    fg, Ag = _run(h, nnodes, len(ci2), x)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    # Dirichlet rows work on any block layout
    dofs = np.array([0, 7, 31], np.int32); vals = np.array([1.0, -2.0, 0.5])
    h.dirichlet_set(dofs, vals)
    f = torch.zeros(nnodes, dtype=torch.float64, device=DEV); A = torch.zeros(len(ci2), dtype=torch.float64, device=DEV)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(DEV), f, A, flags=capi.FLAG_ALL); h.sync()
    oracle.dirichlet(1, dofs, vals, x, fo, rp2, ci2, Ao)
    assert _relerr(f.cpu().numpy(), fo) < RTOL and _relerr(A.cpu().numpy(), Ao) < RTOL
    h.close()


@pytest.mark.parametrize("layout", ["interleaved", "blocked"])
def test_elastodynamics_three_fields(oracle, layout):
    """3 DOF per node: rho M d2u/dt2 + c M du/dt + K(lambda, mu) u; J = gamma rho M + alpha c M + beta K.  `blocked` passes an
    explicit getGIDFieldOffsets table (field-major element DOF order) instead of FieldAggPattern's interleaving."""
    n = (4, 3, 3)
    (d,), _ = oracle.poisson_problem(n)
    nn = d["n_local"]
    cc = _perturbed(d["cell_coords"], 0.03)
    if layout == "interleaved":
        l3 = (3 * d["lids"][:, :, None] + np.arange(3)[None, None, :]).reshape(-1, 24).astype(np.int32)
        fo_tab = None
    else:
        l3 = np.concatenate([3 * d["lids"] + i for i in range(3)], axis=1).astype(np.int32)          # field-major
        fo_tab = np.arange(24, dtype=np.int32).reshape(3, 8)
    rp, ci = oracle.ghosted_graph(l3, 3 * nn)
    rng = np.random.default_rng(3)
    x, xd, xdd = (rng.standard_normal(3 * nn) for _ in range(3))
    params = [1.3, 0.7, 2.5, 0.4, 0.1, -0.2, 9.81]
    fo, Ao = oracle.block_evaluate(oracle.HEX8_C1, oracle.OP_ELASTICITY, 2, params, cc, l3, rp, ci, x, xd, xdd, alpha=0.6, beta=1.1, gamma=4.0,
                                   field_offsets=fo_tab)
    h = capi.Handle()
    b = h.gblock_add(capi.TOPO_HEX8, capi.BASIS_HGRAD_C1, 2, torch.from_numpy(cc).to(DEV), torch.from_numpy(l3).to(DEV), 3 * nn, n_fields=3,
                     field_offsets=fo_tab)
    h.gblock_terms_set(b, capi.OP_ELASTICITY, params)
    h.graph_set(torch.from_numpy(rp).to(DEV), torch.from_numpy(ci).to(DEV)); h.setup()
    fg, Ag = _run(h, 3 * nn, len(ci), x, xd, xdd, alpha=0.6, beta=1.1, gamma=4.0)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    K = sp.csr_matrix((Ag, ci, rp))
    assert abs(K - K.T).max() < 1e-12 * np.abs(Ag).max()
    h.close()


def test_hcurl_curlcurl_and_mass_with_orientations(oracle):
    n = (4, 3, 2)
    (d,), _ = oracle.poisson_problem(n)
    lids, signs = oracle.hcurl_hex_lids(n)
    ne = int(lids.max()) + 1
    rp, ci = oracle.ghosted_graph(lids, ne)
    cc = _perturbed(d["cell_coords"], 0.03)
    rng = np.random.default_rng(5)
    x, xd = rng.standard_normal(ne), rng.standard_normal(ne)
    # flip some orientation signs arbitrarily as well: the kernel must apply whatever table it is given
    signs2 = signs.copy(); signs2[::3, 5] *= -1
    params = [-1.0, -1.0, 1.0, 0.0, 0.3, -0.1, 0.2]               # the CurlLaplacian example's multipliers + transient
    for sg in (signs, signs2):
        fo, Ao = oracle.block_evaluate(oracle.HEX8_HCURL, oracle.OP_CURLCURL, 2, params, cc, lids, rp, ci, x, xd, alpha=2.0, beta=1.0, signs=sg)
        h = capi.Handle()
        b = h.gblock_add(capi.TOPO_HEX8, capi.BASIS_HCURL_I1, 2, torch.from_numpy(cc).to(DEV), torch.from_numpy(lids).to(DEV), ne,
                         orientation_signs=torch.from_numpy(sg).to(DEV))
        h.gblock_terms_set(b, capi.OP_CURLCURL, params)
        h.graph_set(torch.from_numpy(rp).to(DEV), torch.from_numpy(ci).to(DEV)); h.setup()
        fg, Ag = _run(h, ne, len(ci), x, xd, alpha=2.0, beta=1.0)
        assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
        h.close()
    with pytest.raises(capi.TxasmError):
        capi.Handle().gblock_add(capi.TOPO_HEX8, capi.BASIS_HCURL_I1, 2, cc, lids, ne)           # signs are required
