"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): sparsity graph / LIDs bit-exact; residual and Jacobian values
within 1e-12 relative in fp64 (summation order differs).  "Relative" is taken against the largest
magnitude of the row-space quantity (max |A| resp. max |f|): entries that are analytically zero
(e.g. the face-neighbour couplings of a cube mesh) only carry rounding noise in both codes.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from tianxin_b200 import capi

pytestmark = pytest.mark.gpu

RTOL = 1e-12
MODES = {"atomic": capi.SCATTER_ATOMIC, "rowgather": capi.SCATTER_ROWGATHER, "rowtile": capi.SCATTER_ROWTILE}


def _close(a, b, what):
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err < RTOL, f"{what}: max rel err {err:.3e}"


def _oracle_eval(orc, d, terms, x, xdot=None):
    t = orc.tables_build(d["cell_coords"])
    f = np.zeros(d["n_local"])
    A = np.zeros(d["rowptr"][-1]) if terms.eval_type == 1 else None
    orc.evaluate_volume(terms, d["lids"], t, x, xdot, d["rowptr"], d["colind"], f, A)
    return f, A


def _gpu_handle(d, mode, terms, use_device_arrays=True, build_graph=False):
    h = capi.Handle(scatter_mode=mode)
    dev = torch.device("cuda:0")
    if use_device_arrays:
        lids = torch.from_numpy(d["lids"]).to(dev)
        cc = torch.from_numpy(d["cell_coords"]).to(dev)
        h.block_add(lids, cell_coords=cc, n_rows=d["n_local"])
    else:
        h.block_add(d["lids"], cell_coords=d["cell_coords"], n_rows=d["n_local"])
    if build_graph:
        h.graph_build()
    elif use_device_arrays:
        h.graph_set(torch.from_numpy(d["rowptr"]).to(dev), torch.from_numpy(d["colind"]).to(dev))
    else:
        h.graph_set(d["rowptr"], d["colind"])
    h.terms_set(terms)
    h.setup()
    return h


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("n,perturb", [(8, 0.0), (7, 0.2), ((5, 3, 4), 0.2)])
def test_poisson_jacobian_residual_parity(oracle, mode, n, perturb):
    (d,), _ = oracle.poisson_problem(n, perturb=perturb)
    x = oracle.state_by_gid(d["gids"].max() - np.arange(d["n_local"]))
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    h = _gpu_handle(d, MODES[mode], capi.poisson_terms())
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    f = torch.full((d["n_local"],), 7.0, dtype=torch.float64, device=dev)     # garbage in: must be overwritten / zeroed
    A = torch.full((int(d["rowptr"][-1]),), 7.0, dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, xd, f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
    h.sync()
    _close(f.cpu().numpy(), fo, "f")
    _close(A.cpu().numpy(), Ao, "A")
    # Residual evaluation type: same f, A untouched
    f2 = torch.zeros_like(f)
    h.evaluate(capi.RESIDUAL, xd, f2, None, flags=capi.FLAG_VOLUMETRIC_FILL)
    h.sync()
    fr, _ = _oracle_eval(oracle, d, oracle.make_terms(eval_type=0), x)
    _close(f2.cpu().numpy(), fr, "f (residual)")
    h.close()


@pytest.mark.parametrize("mode", list(MODES))
def test_transient_and_reaction_terms(oracle, mode):
    (d,), _ = oracle.poisson_problem(6, perturb=0.2)
    rng = np.random.default_rng(1)
    x, xdot = rng.standard_normal(d["n_local"]), rng.standard_normal(d["n_local"])
    alpha, beta = 3.25, 0.5
    tm = oracle.make_terms(alpha=alpha, beta=beta, mass_dot=1.0, react=0.3, kappa=2.0)
    fo, Ao = _oracle_eval(oracle, d, tm, x, xdot)
    h = _gpu_handle(d, MODES[mode], capi.poisson_terms(kappa=2.0, mass_dot=1.0, react=0.3))
    dev = torch.device("cuda:0")
    f = torch.zeros(d["n_local"], dtype=torch.float64, device=dev)
    A = torch.zeros(int(d["rowptr"][-1]), dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(dev), f, A, xdot=torch.from_numpy(xdot).to(dev),
               alpha=alpha, beta=beta, flags=capi.FLAG_VOLUMETRIC_FILL)
    h.sync()
    _close(f.cpu().numpy(), fo, "f")
    _close(A.cpu().numpy(), Ao, "A")
    h.close()


def test_host_arrays_and_device_graph_build(oracle):
    """Host pointers are staged by the library; the device-built graph equals the oracle's
    (= CrsGraph::fillComplete's) bit for bit."""
    (d,), _ = oracle.poisson_problem(6, perturb=0.2)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    h = _gpu_handle(d, capi.SCATTER_AUTO, capi.poisson_terms(), use_device_arrays=False, build_graph=True)
    rp = np.empty_like(d["rowptr"]); ci = np.empty_like(d["colind"])
    assert h.info().nnz == d["rowptr"][-1]
    h.graph_get(rp, ci)
    assert np.array_equal(rp, d["rowptr"]) and np.array_equal(ci, d["colind"])
    f = np.full(d["n_local"], np.nan); A = np.full(d["rowptr"][-1], np.nan)
    h.evaluate(capi.JACOBIAN, x, f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
    _close(f, fo, "f"); _close(A, Ao, "A")
    h.close()


@pytest.mark.parametrize("mode", list(MODES))
def test_dirichlet_on_device(oracle, mode):
    (d,), _ = oracle.poisson_problem(5)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    # the six faces: nodes with a coordinate on the box boundary
    xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    dofs = np.where(np.any((xyz < 1e-12) | (xyz > 1 - 1e-12), axis=1))[0].astype(np.int32)
    vals = np.linspace(-1, 1, len(dofs))
    oracle.dirichlet(1, dofs, vals, x, fo, d["rowptr"], d["colind"], Ao)
    h = _gpu_handle(d, MODES[mode], capi.poisson_terms())
    h.dirichlet_set(dofs, vals)
    dev = torch.device("cuda:0")
    f = torch.zeros(d["n_local"], dtype=torch.float64, device=dev)
    A = torch.zeros(int(d["rowptr"][-1]), dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(dev), f, A, flags=capi.FLAG_ALL)
    h.sync()
    fg, Ag = f.cpu().numpy(), A.cpu().numpy()
    _close(fg, fo, "f"); _close(Ag, Ao, "A")
    for l in dofs:        # identity rows exactly
        row = Ag[d["rowptr"][l]:d["rowptr"][l + 1]]
        assert row.sum() == 1.0 and np.count_nonzero(row) == 1
    assert np.array_equal(fg[dofs], x[dofs] - vals)
    h.close()


@pytest.mark.parametrize("n,perturb", [(20, 0.0), (17, 0.2), ((33, 9, 5), 0.0)])
def test_rowtile_many_tiles(oracle, n, perturb):
    """Meshes spanning many row tiles (Morton-ordered chunks of 256 / 128 rows)."""
    (d,), _ = oracle.poisson_problem(n, perturb=perturb)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    info = h.info()
    assert info.scatter_mode == capi.SCATTER_ROWTILE and info.n_tiles > 4
    assert info.n_regular_rows == d["n_local"]
    assert info.n_affine_cells == (d["lids"].shape[0] if perturb == 0.0 else 0) or perturb != 0.0
    dev = torch.device("cuda:0")
    f = torch.full((d["n_local"],), np.nan, dtype=torch.float64, device=dev)
    A = torch.full((int(d["rowptr"][-1]),), np.nan, dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(dev), f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
    h.sync()
    _close(f.cpu().numpy(), fo, "f"); _close(A.cpu().numpy(), Ao, "A")
    h.close()


def test_irregular_connectivity_falls_back_per_row(oracle):
    """Cells whose local vertex numbering is rotated break the canonical 27-point pattern of the rows
    around them; those rows go through the general row-gather kernel, the rest through the tiles."""
    (d,), _ = oracle.poisson_problem(9, perturb=0.1)
    lids, cc = d["lids"].copy(), d["cell_coords"].copy()
    rot = [1, 2, 3, 0, 5, 6, 7, 4]                      # rotate the cell about its zeta axis
    for e in (100, 333, 334, 600):
        lids[e] = lids[e][rot]; cc[e] = cc[e][rot]
    d2 = dict(d, lids=lids, cell_coords=cc)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle_eval(oracle, d2, oracle.make_terms(), x)
    f_ref, A_ref = _oracle_eval(oracle, d, oracle.make_terms(), x)     # same operator, other numbering
    _close(fo, f_ref, "oracle invariance f"); _close(Ao, A_ref, "oracle invariance A")
    h = _gpu_handle(d2, capi.SCATTER_ROWTILE, capi.poisson_terms())
    info = h.info()
    assert 0 < info.n_regular_rows < d["n_local"]
    dev = torch.device("cuda:0")
    f = torch.full((d["n_local"],), np.nan, dtype=torch.float64, device=dev)
    A = torch.full((int(d["rowptr"][-1]),), np.nan, dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(dev), f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
    h.sync()
    _close(f.cpu().numpy(), fo, "f"); _close(A.cpu().numpy(), Ao, "A")
    h.close()


def test_reproducible_owner_computes(oracle):
    """The atomics-free paths are bitwise reproducible run to run."""
    (d,), _ = oracle.poisson_problem(9, perturb=0.2)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    h = _gpu_handle(d, capi.SCATTER_AUTO, capi.poisson_terms())
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    outs = []
    for _ in range(3):
        f = torch.empty(d["n_local"], dtype=torch.float64, device=dev)
        A = torch.empty(int(d["rowptr"][-1]), dtype=torch.float64, device=dev)
        h.evaluate(capi.JACOBIAN, xd, f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
        h.sync()
        outs.append((f.cpu().numpy().copy(), A.cpu().numpy().copy()))
    assert all(np.array_equal(outs[0][0], o[0]) and np.array_equal(outs[0][1], o[1]) for o in outs[1:])
    h.close()


def test_tile_tables_are_consistent(oracle):
    """Row tiles partition the rows; every (row, a) entry points at a staged cell that has the row as local
    vertex a; the staged cells of a tile are in ascending cell-id order without repeats."""
    (d,), _ = oracle.poisson_problem(12)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    info = h.info()
    seen = np.zeros(d["n_local"], np.int32)
    for t in range(info.n_tiles):
        rows, cells, adjl = h.tile_get(t)
        assert len(cells) <= info.tile_cells_max
        assert np.all(np.diff(cells) > 0)
        for slot, r in enumerate(rows):
            if r < 0:
                assert np.all(adjl[slot] == 0xFFFF)
                continue
            seen[r] += 1
            for a in range(8):
                pos = int(adjl[slot, a])
                if pos == 0xFFFF:
                    assert not np.any(d["lids"][:, a] == r)          # no cell has this row as vertex a
                    continue
                c = int(cells[pos])
                assert c >= 0 and d["lids"][c, a] == r
    assert np.all(seen == 1)
    h.close()


def test_multi_gpu_parity_if_available():
    """NCCL halo import/export against the oracle on 2 GPUs (tools/multigpu_check.py); skipped on one GPU."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for extra in (["--size", "5", "--perturb", "0.2"], ["--size", "20"], ["--size", "20", "--p2p", "0"], ["--size", "20", "--overlap", "0"], ["--size", "20", "--edge", "1"]):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
               "--master-port", "29531", os.path.join(root, "tools", "multigpu_check.py")] + extra
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
        assert "-> OK" in out.stdout


def test_concentrated_load_and_dirichlet_residual(oracle):
    """Residual evaluation with all stages: volume fill, concentrated loads (f += v), then Dirichlet rows
    (f = x - value); the Jacobian-type evaluation ignores concentrated loads like the reference's evaluator."""
    (d,), _ = oracle.poisson_problem(6, perturb=0.1)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    cl_dofs = np.array([3, 77, 200, 77], np.int32); cl_vals = np.array([1.5, -2.0, 0.25, 4.0])
    dr_dofs = np.array([0, 5, 200], np.int32); dr_vals = np.array([0.5, -1.0, 2.0])
    fo, _ = _oracle_eval(oracle, d, oracle.make_terms(eval_type=0), x)
    oracle.cload(0, cl_dofs, cl_vals, fo)
    oracle.dirichlet(0, dr_dofs, dr_vals, x, fo, d["rowptr"], d["colind"], None)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    h.cload_set(cl_dofs, cl_vals); h.dirichlet_set(dr_dofs, dr_vals)
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    f = torch.zeros(d["n_local"], dtype=torch.float64, device=dev)
    h.evaluate(capi.RESIDUAL, xd, f, None, flags=capi.FLAG_ALL); h.sync()
    _close(f.cpu().numpy(), fo, "f")
    assert f[200].item() == x[200] - 2.0                       # Dirichlet after the load wins
    fj, Aj = _oracle_eval(oracle, d, oracle.make_terms(), x)
    oracle.dirichlet(1, dr_dofs, dr_vals, x, fj, d["rowptr"], d["colind"], Aj)
    A = torch.zeros(int(d["rowptr"][-1]), dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, xd, f, A, flags=capi.FLAG_ALL); h.sync()
    _close(f.cpu().numpy(), fj, "f (jacobian type, no cload)"); _close(A.cpu().numpy(), Aj, "A")
    h.close()


def _fill_and_compare(oracle, d, x, terms_gpu=None, terms_orc=None, xdot=None, alpha=0.0):
    terms_orc = terms_orc or oracle.make_terms()
    fo, Ao = _oracle_eval(oracle, d, terms_orc, x, xdot)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, terms_gpu or capi.poisson_terms())
    dev = torch.device("cuda:0")
    f = torch.full((d["n_local"],), np.nan, dtype=torch.float64, device=dev)
    A = torch.full((int(d["rowptr"][-1]),), np.nan, dtype=torch.float64, device=dev)
    xd = torch.from_numpy(x).to(dev)
    xdd = None if xdot is None else torch.from_numpy(xdot).to(dev)
    for _ in range(2):                       # twice: the second evaluate reuses the constant row image
        f.fill_(float("nan")); A.fill_(float("nan"))
        h.evaluate(capi.JACOBIAN, xd, f, A, xdot=xdd, alpha=alpha, flags=capi.FLAG_VOLUMETRIC_FILL)
        h.sync()
        _close(f.cpu().numpy(), fo, "f"); _close(A.cpu().numpy(), Ao, "A")
    info = h.info()
    h.close()
    return info


def test_congruent_tiles_with_renumbered_nodes(oracle):
    """Uniform mesh whose LIDs are shuffled: the cells stay congruent (interior rows take the tile's stiffness row)
    but the CSR column order is no longer canonical, so rows are placed through their permutation, not the image."""
    (d,), _ = oracle.poisson_problem(11)
    rng = np.random.default_rng(5)
    P = rng.permutation(d["n_local"]).astype(np.int32)            # old lid -> new lid
    lids = P[d["lids"]]
    rowptr, colind = oracle.ghosted_graph(lids, d["n_local"])
    d2 = dict(d, lids=lids, rowptr=rowptr, colind=colind)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    _fill_and_compare(oracle, d2, x)


def test_two_cell_sizes_rebuild_the_row_image(oracle):
    """Rectilinear mesh with two spacings along x: two congruence classes (the constant row image is rebuilt when a CTA
    moves from one to the other) and non-congruent tiles across the interface."""
    (d,), _ = oracle.poisson_problem((16, 9, 9))
    cc = d["cell_coords"].copy()
    xs = cc[..., 0]
    cc[..., 0] = np.where(xs <= 0.5, xs, 0.5 + 2.0 * (xs - 0.5))  # cells right of x = 1/2 are twice as long
    d2 = dict(d, cell_coords=cc)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    info = _fill_and_compare(oracle, d2, x)
    assert info.n_affine_cells == d["lids"].shape[0]


def test_transient_terms_on_congruent_tiles(oracle):
    """Mass terms switch the interior-row shortcut off (A = cK K + cM M); the metric broadcast stays."""
    (d,), _ = oracle.poisson_problem(9)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    xdot = np.cos(0.11 * np.arange(d["n_local"]))
    terms_orc = oracle.make_terms(alpha=3.0, beta=1.0, mass_dot=2.0, react=0.5)
    terms_gpu = capi.poisson_terms(mass_dot=2.0, react=0.5)
    _fill_and_compare(oracle, d, x, terms_gpu=terms_gpu, terms_orc=terms_orc, xdot=xdot, alpha=3.0)


def test_neumann_flux_then_dirichlet(oracle):
    """BoundaryFill order of the reference: Neumann flux on side sets (f += val * int_side phi), then Dirichlet rows.
    The flux is added by both evaluation types and never touches A."""
    n = (6, 5, 4)
    (d,), _ = oracle.poisson_problem(n, perturb=0.15)
    p = oracle.mesh_params(n)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    cr, sr = oracle.sideset_sides(p, d["elem_ids"], "right")
    ct, st = oracle.sideset_sides(p, d["elem_ids"], "top")
    cells = np.concatenate([cr, ct]); sides = np.concatenate([sr, st])
    vals = np.concatenate([np.full(len(cr), 1.75), np.linspace(-1.0, 2.0, len(ct))])
    dr_dofs = np.array([0, 3, 40], np.int32); dr_vals = np.array([0.5, -1.0, 2.0])
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    A_vol = Ao.copy()
    oracle.neumann_flux(cells, sides, vals, d["lids"], d["cell_coords"], fo)
    oracle.dirichlet(1, dr_dofs, dr_vals, x, fo, d["rowptr"], d["colind"], Ao)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    h.neumann_set(cells, sides, vals); h.dirichlet_set(dr_dofs, dr_vals)
    dev = torch.device("cuda:0")
    xd = torch.from_numpy(x).to(dev)
    f = torch.zeros(d["n_local"], dtype=torch.float64, device=dev)
    A = torch.zeros(int(d["rowptr"][-1]), dtype=torch.float64, device=dev)
    h.evaluate(capi.JACOBIAN, xd, f, A, flags=capi.FLAG_ALL); h.sync()
    _close(f.cpu().numpy(), fo, "f"); _close(A.cpu().numpy(), Ao, "A")
    # residual type: same flux, no matrix
    fr, _ = _oracle_eval(oracle, d, oracle.make_terms(eval_type=0), x)
    oracle.neumann_flux(cells, sides, vals, d["lids"], d["cell_coords"], fr)
    oracle.dirichlet(0, dr_dofs, dr_vals, x, fr, d["rowptr"], d["colind"], None)
    h.evaluate(capi.RESIDUAL, xd, f, None, flags=capi.FLAG_ALL); h.sync()
    _close(f.cpu().numpy(), fr, "f (residual type)")
    with pytest.raises(Exception):
        h.neumann_set(cells[:1], np.array([7], np.int32), vals[:1])
    h.close()
    del A_vol


@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_functional_responses(oracle, perturb):
    """L2 / H1 error functionals and the plain integral against the oracle (sums of positive terms: 1e-12 relative)."""
    (d,), _ = oracle.poisson_problem((9, 7, 6), perturb=perturb)
    x = oracle.state_by_gid(np.arange(d["n_local"])) * 0.1
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    xd = torch.from_numpy(x).to("cuda:0")
    for kind, sol, deg in ((capi.RESP_INTEGRAL, 1, 2), (capi.RESP_L2_ERROR, 1, 10), (capi.RESP_H1_ERROR, 1, 10),
                           (capi.RESP_L2_ERROR, 3, 4), (capi.RESP_H1_ERROR, 3, 6)):
        ref = oracle.response_functional(kind, sol, deg, d["lids"], d["cell_coords"], x)
        got = h.response_functional(kind, xd, solution_id=sol, cubature_degree=deg)
        assert abs(got - ref) <= 1e-12 * max(abs(ref), 1e-300), (kind, sol, deg, got, ref)
        assert h.response_functional(kind, xd, solution_id=sol, cubature_degree=deg) == got     # fixed-order reduction
        assert abs(h.response_functional(kind, x, solution_id=sol, cubature_degree=deg) - got) <= 1e-15 * abs(got)  # host x
    with pytest.raises(Exception):
        h.response_functional(7, xd)
    h.close()


def test_tianxin_response_integral(oracle):
    """TianXin::Response_Integral<Residual>: an arbitrary field at the integration points times weighted_measure, summed;
    the value is also accumulated into entry 0 of the response vector (sumIntoLocalValue)."""
    (d,), _ = oracle.poisson_problem((6, 5, 4), perturb=0.2)
    t = oracle.tables_build(d["cell_coords"])
    rng = np.random.default_rng(8)
    cv = rng.standard_normal((d["lids"].shape[0], 8))
    rv_ref = np.array([2.5]); ref = oracle.response_integral(cv, t.wm, rv_ref)
    h = _gpu_handle(d, capi.SCATTER_ROWTILE, capi.poisson_terms())
    rv = np.array([2.5])
    got = h.response_integral(torch.from_numpy(cv).to("cuda:0"), rv)
    assert abs(got - ref) <= 1e-12 * abs(ref) and abs(rv[0] - rv_ref[0]) <= 1e-12 * abs(rv_ref[0])
    assert abs(h.response_integral(cv, rv) - ref) <= 1e-12 * abs(ref)          # host array; accumulates again
    assert abs(rv[0] - (2.5 + 2 * ref)) <= 1e-12 * abs(ref)
    ones = np.ones_like(cv)
    assert abs(h.response_integral(ones, np.zeros(1)) - 1.0) < 1e-13            # the volume of the unit cube
    with pytest.raises(capi.TxasmError):
        h.response_integral(cv, None)
    h.close()


@pytest.mark.parametrize("own_graph", [True, False])
def test_graph_merge_columns_on_device(oracle, own_graph):
    """txasm_graph_merge_columns (the fill graph on the device): pairs that exist, repeat, are new local
    columns or remote-only columns (index >= n_rows) -- rows stay sorted, positions index the new A_values; then an
    evaluate on the merged graph leaves the inserted entries 0 and the others equal to the oracle's."""
    (d,), _ = oracle.poisson_problem((6, 5, 4))
    nl = d["n_local"]
    rp, ci = d["rowptr"], d["colind"]
    rng = np.random.default_rng(11)
    rows = rng.integers(0, nl, size=400).astype(np.int32)
    cols = rng.integers(0, nl + 37, size=400).astype(np.int32)            # some exist, some are new, some are remote-only
    rows[:50] = rows[50:100]; cols[:50] = cols[50:100]                     # repeats
    rows[100:160] = np.arange(60); cols[100:160] = ci[rp[:60]]             # existing entries
    dev = torch.device("cuda:0")
    h = capi.Handle(scatter_mode=capi.SCATTER_ROWTILE)
    lids = torch.from_numpy(d["lids"]).to(dev); cc = torch.from_numpy(d["cell_coords"]).to(dev)
    h.block_add(lids, cell_coords=cc, n_rows=nl)
    if own_graph:
        h.graph_build()
    else:
        rpt, cit = torch.from_numpy(rp).to(dev), torch.from_numpy(ci).to(dev)
        h.graph_set(rpt, cit)
    pos, nnz = h.graph_merge_columns(rows, cols)
    sets = [set(ci[rp[i]:rp[i + 1]].tolist()) for i in range(nl)]
    for r, c in zip(rows, cols):
        sets[r].add(int(c))
    erp = np.concatenate([[0], np.cumsum([len(s) for s in sets])]).astype(np.int64)
    eci = np.concatenate([np.array(sorted(s), np.int32) for s in sets])
    grp = np.empty(nl + 1, np.int64); gci = np.empty(nnz, np.int32)
    h.graph_get(grp, gci)
    assert nnz == erp[-1] and np.array_equal(grp, erp) and np.array_equal(gci, eci)
    assert np.array_equal(gci[pos], cols) and np.all((pos >= erp[rows]) & (pos < erp[rows + 1]))
    if not own_graph:
        assert np.array_equal(cit.cpu().numpy(), ci)                      # the caller's arrays are left alone
    # ghost rows through graph_get_rows
    g_rp, g_ci = h.graph_get_rows(nl - 17, 17)
    assert np.array_equal(g_rp, erp[nl - 17:] - erp[nl - 17]) and np.array_equal(g_ci, eci[erp[nl - 17]:])
    # the merged graph assembles: old entries as before, inserted entries 0
    h.terms_set(capi.poisson_terms()); h.setup()
    x = oracle.state_by_gid(np.arange(nl))
    fo, Ao = _oracle_eval(oracle, d, oracle.make_terms(), x)
    f = torch.zeros(nl, dtype=torch.float64, device=dev); A = torch.full((nnz,), float("nan"), dtype=torch.float64, device=dev)
    h.evaluate(1, torch.from_numpy(x).to(dev), f, A)
    h.sync()
    Ag = A.cpu().numpy()
    old = np.concatenate([erp[i] + np.searchsorted(eci[erp[i]:erp[i + 1]], ci[rp[i]:rp[i + 1]]) for i in range(nl)])
    _close(Ag[old], Ao, "A on the old entries")
    mask = np.ones(nnz, bool); mask[old] = False
    assert np.all(Ag[mask] == 0.0)
    _close(f.cpu().numpy(), fo, "f")
