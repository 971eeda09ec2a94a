"""GPU parity on the graded configurations and on the kernels that carry the headline number.

BASELINE.json configs[0] (32^3, C1) through the whole evaluate(All); a 96^3 mesh (thousands of tiles: the persistent
CTAs of k_fill_brick / k_fill_uniform / k_fill_rowtile walk many tiles each) against the oracle at the north-star
tolerance (1e-12 relative to max|A| resp. max|f|); the three GPU paths for uniform tiles (brick, uniform, general
row-tile kernel) cross-checked against each other; the image rebuild when the cell shape changes inside a kernel;
fused versus separate Dirichlet; host arrays through split stages; the second-order-in-time extension.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
from tianxin_b200 import capi

pytestmark = pytest.mark.gpu
RTOL = 1e-12
DEV = "cuda:0"


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def _oracle(orc, d, terms, x, xdot=None, xdotdot=None, nthreads=1):
    t = orc.tables_build(d["cell_coords"])
    f = np.zeros(d["n_local"])
    A = np.zeros(d["rowptr"][-1]) if terms.eval_type == 1 else None
    terms.nthreads = nthreads
    orc.evaluate_volume(terms, d["lids"], t, x, xdot, d["rowptr"], d["colind"], f, A, xdotdot=xdotdot)
    return f, A


def _handle(d, terms, mode=capi.SCATTER_ROWTILE, **opts):
    h = capi.Handle(scatter_mode=mode)
    lids = torch.from_numpy(d["lids"]).to(DEV)
    cc = torch.from_numpy(d["cell_coords"]).to(DEV)
    h.block_add(lids, cell_coords=cc, n_rows=d["n_local"])
    h.graph_set(torch.from_numpy(d["rowptr"]).to(DEV), torch.from_numpy(d["colind"]).to(DEV))
    h.terms_set(terms)
    h.setup()
    for k, v in opts.items():
        h.option_set(k, v)
    return h


def _boundary_dofs(d, lo=0.0, hi=1.0):
    xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    return np.where(np.any((xyz < lo + 1e-12) | (xyz > hi - 1e-12), axis=1))[0].astype(np.int32)


def _evaluate(h, d, x, flags=capi.FLAG_VOLUMETRIC_FILL, eval_type=capi.JACOBIAN, **kw):
    xd = torch.from_numpy(x).to(DEV)
    f = torch.full((d["n_local"],), np.nan, dtype=torch.float64, device=DEV)
    A = torch.full((int(d["rowptr"][-1]),), np.nan, dtype=torch.float64, device=DEV) if eval_type == capi.JACOBIAN else None
    kw = {k: (torch.from_numpy(v).to(DEV) if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
    h.evaluate(eval_type, xd, f, A, flags=flags, **kw)
    h.sync()
    return f.cpu().numpy(), (A.cpu().numpy() if A is not None else None)


def test_config1_32cubed_evaluate_all(oracle):
    """BASELINE.json configs[0]: 3-D Poisson Q1 hex on the 32^3 inline mesh, residual + Jacobian, all four stages
    (Dirichlet on the six side sets, value 0), against the oracle."""
    (d,), _ = oracle.poisson_problem(32)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle(oracle, d, oracle.make_terms(), x)
    dofs = _boundary_dofs(d); vals = np.zeros(len(dofs))
    oracle.dirichlet(1, dofs, vals, x, fo, d["rowptr"], d["colind"], Ao)
    for cap, edge in ((0, 0), (3, 0), (0, 1)):          # 3 CTAs: every kernel walks many tiles per CTA; edge: k_fill_edge on the boundary tiles
        h = _handle(d, capi.poisson_terms(), grid_cap=cap, edge_kernel=edge)
        h.dirichlet_set(dofs, vals)
        fg, Ag = _evaluate(h, d, x, flags=capi.FLAG_ALL)
        info = h.info()
        assert info.n_uniform_tiles > 0 and info.n_brick_tiles == info.n_uniform_tiles
        assert info.uniform_kernel_used == 2 and info.dirichlet_fused == 1
        assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
        assert np.array_equal(fg[dofs], x[dofs] - vals)
        h.close()


@pytest.mark.parametrize("n", [96])
def test_96cubed_many_tiles_per_cta(oracle, n):
    """Thousands of tiles: n_brick_tiles exceeds ctas_per_sm * n_sm, so persistent CTAs walk several tiles each
    (record prefetch two tiles ahead, mbarrier parity, image reuse).  f and A against the oracle."""
    (d,), _ = oracle.poisson_problem(n)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle(oracle, d, oracle.make_terms(), x, nthreads=8)
    h = _handle(d, capi.poisson_terms())
    fg, Ag = _evaluate(h, d, x)
    info = h.info()
    assert info.uniform_kernel_used == 2
    assert info.n_brick_tiles > info.ctas_per_sm * info.n_sm, (info.n_brick_tiles, info.ctas_per_sm, info.n_sm)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    # second evaluate on the same handle (tables warm), residual type
    fr, _ = _evaluate(h, d, x, eval_type=capi.RESIDUAL)
    assert _relerr(fr, fo) < RTOL
    h.close()


@pytest.mark.parametrize("n", [40, (37, 21, 50)])
def test_three_gpu_paths_agree(oracle, n):
    """The same uniform tiles through k_fill_brick, k_fill_uniform and the general k_fill_rowtile.  The two cell-based
    kernels must agree bit for bit in A (the same constant row image) and to rounding in f; the lattice kernel sums in
    another order (1e-13)."""
    (d,), _ = oracle.poisson_problem(n)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle(oracle, d, oracle.make_terms(), x, nthreads=4)
    out = {}
    for name, opts in (("brick", {"edge_kernel": 1}), ("brick_noedge", {"edge_kernel": 0}), ("uniform", {"brick_kernel": 0, "edge_kernel": 0}),
                       ("rowtile", {"uniform_kernel": 0})):
        h = _handle(d, capi.poisson_terms(), grid_cap=5, **opts)
        out[name] = _evaluate(h, d, x)
        assert h.info().uniform_kernel_used == {"brick": 2, "brick_noedge": 2, "uniform": 1, "rowtile": 0}[name]
        assert h.info().n_edge_tiles > 0                   # the tiles on the boundary of an inline mesh are lattice tiles too
        assert _relerr(out[name][0], fo) < RTOL and _relerr(out[name][1], Ao) < RTOL, name
        h.close()
    assert np.array_equal(out["uniform"][1], out["rowtile"][1])
    assert _relerr(out["uniform"][0], out["rowtile"][0]) < 1e-14
    assert _relerr(out["brick_noedge"][1], out["uniform"][1]) < 1e-14
    assert _relerr(out["brick_noedge"][0], out["uniform"][0]) < 1e-13
    assert _relerr(out["brick"][1], out["brick_noedge"][1]) < 1e-13     # k_fill_edge vs k_fill_rowtile on the boundary tiles
    assert _relerr(out["brick"][0], out["brick_noedge"][0]) < 1e-13


def test_two_cell_sizes_inside_the_uniform_kernels(oracle):
    """Rectilinear mesh with two spacings along x, large enough that both halves hold brick tiles: a CTA (grid capped at
    2) meets both cell shapes and rebuilds its row image in flight; the tiles across the interface are not congruent."""
    (d,), _ = oracle.poisson_problem((48, 20, 20))
    cc = d["cell_coords"].copy()
    xs = cc[..., 0]
    cc[..., 0] = np.where(xs <= 0.5, xs, 0.5 + 2.0 * (xs - 0.5))
    d2 = dict(d, cell_coords=cc)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    fo, Ao = _oracle(oracle, d2, oracle.make_terms(), x)
    for opts in ({}, {"brick_kernel": 0}):
        h = _handle(d2, capi.poisson_terms(), grid_cap=2, **opts)
        info = h.info()
        assert info.n_uniform_tiles >= 8 and info.n_affine_cells == d["lids"].shape[0]
        for _ in range(2):
            fg, Ag = _evaluate(h, d2, x)
            assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
        h.close()


def test_mass_terms_through_the_brick_kernel(oracle):
    """Transient + reaction terms: A = cK K + cM M; the lattice kernel carries the node mass stencil, the uniform
    cell kernel does not take mass terms (those tiles fall to the general kernel)."""
    (d,), _ = oracle.poisson_problem(24)
    rng = np.random.default_rng(3)
    x, xdot = rng.standard_normal(d["n_local"]), rng.standard_normal(d["n_local"])
    alpha, beta = 2.5, 0.75
    tm = oracle.make_terms(alpha=alpha, beta=beta, mass_dot=1.5, react=0.3, kappa=2.0)
    fo, Ao = _oracle(oracle, d, tm, x, xdot)
    for opts, used in (({}, 2), ({"edge_kernel": 0}, 2), ({"brick_kernel": 0}, 0)):
        h = _handle(d, capi.poisson_terms(kappa=2.0, mass_dot=1.5, react=0.3), grid_cap=4, **opts)
        fg, Ag = _evaluate(h, d, x, xdot=xdot, alpha=alpha, beta=beta)
        assert h.info().uniform_kernel_used == used
        assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
        h.close()


@pytest.mark.parametrize("mode", ["atomic", "rowgather", "rowtile"])
@pytest.mark.parametrize("n,perturb", [(10, 0.0), (7, 0.2)])
def test_second_order_in_time_extension(oracle, mode, n, perturb):
    """TXASM_VEC_XDOTDOT with seed gamma (extension, parity unpinned by the reference): against the oracle's restatement
    and against identities -- J(gamma) - J(0) = gamma * rho * M with M row sums = nodal volumes, f linear in xdotdot."""
    (d,), _ = oracle.poisson_problem(n, perturb=perturb)
    rng = np.random.default_rng(11)
    x, xd, xdd = (rng.standard_normal(d["n_local"]) for _ in range(3))
    alpha, beta, gamma, rho = 1.25, 0.5, 4.0, 3.0
    tm = oracle.make_terms(alpha=alpha, beta=beta, gamma=gamma, mass_dot=0.7, mass_dotdot=rho)
    fo, Ao = _oracle(oracle, d, tm, x, xd, xdd)
    mode_id = {"atomic": capi.SCATTER_ATOMIC, "rowgather": capi.SCATTER_ROWGATHER, "rowtile": capi.SCATTER_ROWTILE}[mode]
    h = _handle(d, capi.poisson_terms(mass_dot=0.7, mass_dotdot=rho), mode=mode_id)
    fg, Ag = _evaluate(h, d, x, xdot=xd, xdotdot=xdd, alpha=alpha, beta=beta, gamma=gamma)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    f0, A0 = _evaluate(h, d, x, xdot=xd, xdotdot=xdd, alpha=alpha, beta=beta, gamma=0.0)
    M = (Ag - A0) / (gamma * rho)                      # the mass matrix
    rowsum = np.add.reduceat(M, d["rowptr"][:-1])
    vol = np.zeros(d["n_local"])
    t = oracle.tables_build(d["cell_coords"])
    np.add.at(vol, d["lids"].ravel(), t.wbasis.sum(axis=2).ravel())          # int phi_b over each cell
    assert np.abs(rowsum - vol).max() < 1e-12 * vol.max()
    assert abs(rowsum.sum() - 1.0) < 1e-12             # the unit cube
    f2, _ = _evaluate(h, d, x, xdot=xd, xdotdot=2.0 * xdd, alpha=alpha, beta=beta, gamma=gamma)
    fz, _ = _evaluate(h, d, x, xdot=xd, xdotdot=0.0 * xdd, alpha=alpha, beta=beta, gamma=gamma)
    assert _relerr(f2 - fg, fg - fz) < 1e-10           # f is linear in xdotdot: both differences are rho M xdotdot
    h.close()


def test_fused_dirichlet_equals_separate_launch(oracle):
    (d,), _ = oracle.poisson_problem(20)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    dofs = _boundary_dofs(d); vals = np.linspace(-1, 1, len(dofs))
    out = []
    for fuse in (1, 0):
        h = _handle(d, capi.poisson_terms(), fuse_dirichlet=fuse)
        h.dirichlet_set(dofs, vals)
        out.append(_evaluate(h, d, x, flags=capi.FLAG_ALL))
        info = h.info()
        assert info.dirichlet_fused == fuse
        h.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    # a Dirichlet DOF in the interior (a uniform row) cannot be fused: the library falls back to the separate launch
    h = _handle(d, capi.poisson_terms())
    xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    interior = np.nonzero((np.abs(xyz[:, 1] - 0.5) < 1e-9) & (np.abs(xyz[:, 2] - 0.5) < 1e-9) & (np.abs(xyz[:, 0] - 0.5) < 0.06))[0].astype(np.int32)
    assert len(interior) == 3                 # nodes (9..11, 10, 10): rows of a uniform tile
    dd = np.concatenate([dofs, interior]); vv = np.concatenate([vals, [0.5, -0.5, 2.0]])
    h.dirichlet_set(dd, vv)
    fg, Ag = _evaluate(h, d, x, flags=capi.FLAG_ALL)
    assert h.info().dirichlet_fused == 0
    fo, Ao = _oracle(oracle, d, oracle.make_terms(), x)
    oracle.dirichlet(1, dd, vv, x, fo, d["rowptr"], d["colind"], Ao)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    h.close()


def test_dirichlet_without_residual_zeroes_columns(oracle):
    """Jacobian evaluation with f == NULL (the eigenvalue path): rows AND columns of the Dirichlet DOFs
    (applyDirichletBoundaryConditionToLocalMatrixRowsAndColumns, lof/Panzer_TpetraLinearObjContainer.hpp:223-226)."""
    (d,), _ = oracle.poisson_problem(8, perturb=0.1)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    _, Ao = _oracle(oracle, d, oracle.make_terms(), x)
    dofs = _boundary_dofs(d)
    oracle.dirichlet_rows_and_columns(dofs, d["rowptr"], d["colind"], Ao)
    h = _handle(d, capi.poisson_terms())
    h.dirichlet_set(dofs, np.zeros(len(dofs)))
    A = torch.full((int(d["rowptr"][-1]),), np.nan, dtype=torch.float64, device=DEV)
    h.evaluate(capi.JACOBIAN, torch.from_numpy(x).to(DEV), None, A, flags=capi.FLAG_ALL)
    h.sync()
    Ag = A.cpu().numpy()
    assert _relerr(Ag, Ao) < RTOL
    import scipy.sparse as sp
    M = sp.csr_matrix((Ag, d["colind"], d["rowptr"]))
    assert abs(M - M.T).max() < 1e-12 * np.abs(Ag).max()       # stays symmetric
    h.close()


def test_split_stages_on_host_arrays(oracle):
    """evaluate(VolumetricFill) then evaluate(BoundaryFill) on numpy arrays == evaluate(All): the staging buffers of
    host outputs start from the caller's contents when a call does not overwrite them."""
    (d,), _ = oracle.poisson_problem(7, perturb=0.1)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    dofs = _boundary_dofs(d); vals = np.linspace(0.5, 1.5, len(dofs))
    h = capi.Handle(scatter_mode=capi.SCATTER_AUTO)
    h.block_add(d["lids"], cell_coords=d["cell_coords"], n_rows=d["n_local"])
    h.graph_set(d["rowptr"], d["colind"])
    h.terms_set(capi.poisson_terms()); h.dirichlet_set(dofs, vals); h.setup()
    f_all = np.full(d["n_local"], np.nan); A_all = np.full(d["rowptr"][-1], np.nan)
    h.evaluate(capi.JACOBIAN, x, f_all, A_all, flags=capi.FLAG_ALL)
    f = np.full(d["n_local"], np.nan); A = np.full(d["rowptr"][-1], np.nan)
    h.evaluate(capi.JACOBIAN, x, f, A, flags=capi.FLAG_VOLUMETRIC_FILL)
    # scribble over the library's staging buffers with another evaluate, so that only a copy-in can be right
    h.evaluate(capi.JACOBIAN, 2.0 * x, np.empty_like(f), np.empty_like(A), flags=capi.FLAG_VOLUMETRIC_FILL)
    h.evaluate(capi.JACOBIAN, x, f, A, flags=capi.FLAG_BOUNDARY_FILL)
    assert np.array_equal(f, f_all) and np.array_equal(A, A_all)
    h.close()
    # accumulate mode of the atomic path: zero_outputs = 0 adds to what the caller holds
    h = capi.Handle(scatter_mode=capi.SCATTER_ATOMIC)
    h.block_add(d["lids"], cell_coords=d["cell_coords"], n_rows=d["n_local"])
    h.graph_set(d["rowptr"], d["colind"]); h.terms_set(capi.poisson_terms()); h.setup()
    f1 = np.zeros(d["n_local"]); A1 = np.zeros(d["rowptr"][-1])
    h.evaluate(capi.JACOBIAN, x, f1, A1, flags=capi.FLAG_VOLUMETRIC_FILL)
    f2 = np.ones(d["n_local"]); A2 = np.full(d["rowptr"][-1], 2.0)
    h.evaluate(capi.JACOBIAN, x, f2, A2, flags=capi.FLAG_VOLUMETRIC_FILL, zero_outputs=0)
    assert _relerr(f2 - 1.0, f1) < 1e-12 and _relerr(A2 - 2.0, A1) < 1e-12
    h.close()


def test_gather_seeds_and_transient_flag(oracle):
    """gather_seeds[i] replaces beta for a DOF gathered with "Gather Seed Index" i
    (Panzer_GatherSolution_Tpetra_impl.hpp:554-572); Integrator_TransientBasisTimesScalar contributes only when
    evaluate_transient_terms is set."""
    (d,), _ = oracle.poisson_problem(6, perturb=0.2)
    rng = np.random.default_rng(2)
    x, xdot = rng.standard_normal(d["n_local"]), rng.standard_normal(d["n_local"])
    terms = [capi.Term(capi.TERM_GRADGRAD, capi.VEC_X, 1.0, 0, None, 2, 0),            # seeded with gather_seeds[1]
             capi.Term(capi.TERM_TRANSIENT_MASS, capi.VEC_XDOT, 1.0, 0, None, 0, 0),
             capi.Term(capi.TERM_SOURCE, capi.VEC_X, -1.0, capi.SOURCE_SIN3, None, 0, 0)]
    h = _handle(d, terms)
    fg, Ag = _evaluate(h, d, x, xdot=xdot, alpha=2.0, beta=123.0, gather_seeds=[9.0, 0.25], evaluate_transient_terms=True)
    fo, Ao = _oracle(oracle, d, oracle.make_terms(alpha=2.0, beta=0.25, mass_dot=1.0), x, xdot)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    fg, Ag = _evaluate(h, d, x, xdot=xdot, alpha=2.0, beta=123.0, gather_seeds=[9.0, 0.25], evaluate_transient_terms=False)
    fo, Ao = _oracle(oracle, d, oracle.make_terms(alpha=2.0, beta=0.25), x)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    with pytest.raises(capi.TxasmError):
        _evaluate(h, d, x, xdot=xdot, alpha=2.0, beta=1.0)          # term wants gather_seeds[1], none given
    h.close()


@pytest.mark.parametrize("mode", ["atomic", "rowgather", "rowtile"])
def test_integrator_field_multipliers(oracle, mode):
    """Field multipliers of the integrators (Panzer_Integrator_GradBasisDotVector_impl.hpp:257-294): a conductivity and a
    density given at the integration points.  Even on the uniform mesh every cell then takes the general path."""
    (d,), _ = oracle.poisson_problem((7, 6, 5))
    ne = d["lids"].shape[0]
    rng = np.random.default_rng(6)
    kq, rq = 1.0 + rng.random((ne, 8)), 0.5 + rng.random((ne, 8))
    x, xdot = rng.standard_normal(d["n_local"]), rng.standard_normal(d["n_local"])
    fo, Ao = _oracle(oracle, d, oracle.make_terms(alpha=2.0, beta=0.5, kappa=1.5, mass_dot=0.8, react=0.2, fm_grad=kq, fm_mass=rq), x, xdot)
    kd, rd = torch.from_numpy(kq).to(DEV), torch.from_numpy(rq).to(DEV)
    terms = [capi.Term(capi.TERM_MASS, capi.VEC_XDOT, 0.8, 0, None, 0, 0, rd.data_ptr()),
             capi.Term(capi.TERM_GRADGRAD, capi.VEC_X, 1.5, 0, None, 0, 0, kd.data_ptr()),
             capi.Term(capi.TERM_MASS, capi.VEC_X, 0.2, 0, None, 0, 0, rd.data_ptr()),
             capi.Term(capi.TERM_SOURCE, capi.VEC_X, -1.0, capi.SOURCE_SIN3, None, 0, 0, None)]
    mode_id = {"atomic": capi.SCATTER_ATOMIC, "rowgather": capi.SCATTER_ROWGATHER, "rowtile": capi.SCATTER_ROWTILE}[mode]
    h = _handle(d, terms, mode=mode_id)
    assert h.info().n_affine_cells == 0
    fg, Ag = _evaluate(h, d, x, xdot=xdot, alpha=2.0, beta=0.5)
    assert _relerr(fg, fo) < RTOL and _relerr(Ag, Ao) < RTOL
    h.close()


def test_stage_timers_are_an_option_and_do_not_change_results(oracle):
    """The CUDA events between the stages are off by default (they cost ~2 % of a 1 ms step): txasm_timers_get /
    txasm_last_fill_ms then refuse; with option stage_timers = 1 they report, and the fill-event ring returns one span per
    evaluate.  Results are bitwise the same either way."""
    (d,), _ = oracle.poisson_problem(16)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    h = _handle(d, capi.poisson_terms())
    assert h.option_get("stage_timers") == 0
    f0, A0 = _evaluate(h, d, x, flags=capi.FLAG_ALL)
    with pytest.raises(capi.TxasmError):
        h.timers()
    with pytest.raises(capi.TxasmError):
        h.last_fill_ms()
    h.option_set("stage_timers", 1)
    h.option_set("fill_event_ring", 4)
    for _ in range(6):
        f1, A1 = _evaluate(h, d, x, flags=capi.FLAG_ALL)
    tm = h.timers()
    assert tm.evaluate_volume > 0.0 and h.last_fill_ms() > 0.0
    hist = h.fill_ms_history(16)
    assert len(hist) == 4 and all(v > 0.0 for v in hist)
    assert np.array_equal(f0, f1) and np.array_equal(A0, A1)
    h.option_set("fill_event_ring", 0)
    assert h.fill_ms_history(16) == []
