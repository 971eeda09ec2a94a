"""The reference stores no golden residual/Jacobian for any mesh: element
matrix VALUES are pinned through closed forms and identities instead.  These tests apply them to
the oracle; tests/test_parity_gpu.py then holds the CUDA path to the oracle."""
import numpy as np
import pytest


def _assemble(orc, prob, terms, x=None, xdot=None):
    d = prob
    t = orc.tables_build(d["cell_coords"])
    n = d["n_local"]
    if x is None:
        x = np.zeros(n)
    f = np.zeros(n)
    A = np.zeros(d["rowptr"][-1]) if terms.eval_type == 1 else None
    orc.evaluate_volume(terms, d["lids"], t, x, xdot, d["rowptr"], d["colind"], f, A)
    return f, A


def _dense(d, A):
    n = d["n_local"]
    M = np.zeros((n, n))
    for i in range(n):
        M[i, d["colind"][d["rowptr"][i]:d["rowptr"][i + 1]]] = A[d["rowptr"][i]:d["rowptr"][i + 1]]
    return M


def test_unit_cube_stencil(oracle):
    # h = 1 cubes: K_e = 1/3 (diag), 0 (edge), -1/12 (face diag), -1/12 (body diag)
    (d,), _ = oracle.poisson_problem(4, box=(0, 4, 0, 4, 0, 4))
    f, A = _assemble(oracle, d, oracle.make_terms(source_id=0))
    M = _dense(d, A)
    # locate the centre node via the gids (gid = node id - 1 on one rank)
    gid_of_lid = np.empty(d["n_local"], np.int64)
    gid_of_lid[d["lids"].ravel()] = d["gids"].ravel()
    lid_of_gid = np.argsort(gid_of_lid)
    node = lambda i, j, k: lid_of_gid[i + 5 * (j + 5 * k)]
    c = node(2, 2, 2)
    assert M[c, c] == pytest.approx(8 / 3, rel=1e-14)
    for (di, dj, dk) in [(a, b, e) for a in (-1, 0, 1) for b in (-1, 0, 1) for e in (-1, 0, 1)]:
        nz = abs(di) + abs(dj) + abs(dk)
        want = {0: 8 / 3, 1: 0.0, 2: -1 / 6, 3: -1 / 12}[nz]
        assert M[c, node(2 + di, 2 + dj, 2 + dk)] == pytest.approx(want, rel=1e-13, abs=1e-15)
    assert d["rowptr"][c + 1] - d["rowptr"][c] == 27


@pytest.mark.parametrize("perturb", [0.0, 0.2])
def test_symmetry_nullspace_linear(oracle, perturb):
    (d,), _ = oracle.poisson_problem(5, perturb=perturb)
    f, A = _assemble(oracle, d, oracle.make_terms(source_id=0))
    M = _dense(d, A)
    scale = np.abs(M).max()
    assert np.abs(M - M.T).max() < 1e-14 * scale                   # symmetry
    assert np.abs(M.sum(axis=1)).max() < 1e-13 * scale             # K * 1 = 0
    # R(u) = K u for the source-free operator, and J = beta K: residual of a run equals A x
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    f2, A2 = _assemble(oracle, d, oracle.make_terms(source_id=0), x=x)
    assert np.allclose(A2, A, rtol=0, atol=1e-14 * scale)
    assert np.allclose(f2, M @ x, rtol=0, atol=1e-12 * scale)
    # linear field: K u_lin vanishes at every interior node (consistency of the gradients)
    node_xyz = np.zeros((d["n_local"], 3))
    node_xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    u = 1.0 + 2.0 * node_xyz[:, 0] - 3.0 * node_xyz[:, 1] + 0.5 * node_xyz[:, 2]
    r = M @ u
    interior = np.all((node_xyz > 1e-12) & (node_xyz < 1 - 1e-12), axis=1)
    assert interior.sum() == 4 ** 3
    assert np.abs(r[interior]).max() < 1e-13 * scale * np.abs(u).max()


def test_beta_alpha_seeds_and_mass(oracle):
    (d,), _ = oracle.poisson_problem(4, perturb=0.2)
    n = d["n_local"]
    rng = np.random.default_rng(0)
    x, xd = rng.standard_normal(n), rng.standard_normal(n)
    fK, AK = _assemble(oracle, d, oracle.make_terms(source_id=0, beta=1.0), x=x)
    # mass matrix via the reaction term alone
    fM, AM = _assemble(oracle, d, oracle.make_terms(source_id=0, kappa=0.0, react=1.0, beta=1.0), x=x)
    M = _dense(d, AM)
    assert M.sum() == pytest.approx(1.0, rel=1e-13)                 # total volume of [0,1]^3
    assert np.abs(M - M.T).max() < 1e-16
    # transient Jacobian = alpha*M + beta*K (GatherSolution seeds, Panzer_GatherSolution_Tpetra_impl.hpp:554-572)
    a, b = 2.5, 0.75
    tm = oracle.make_terms(source_id=0, alpha=a, beta=b, mass_dot=1.0)
    f, A = _assemble(oracle, d, tm, x=x, xdot=xd)
    assert np.allclose(A, a * AM + b * AK, rtol=0, atol=1e-14 * np.abs(AK).max())
    assert np.allclose(f, _dense(d, AK) @ x + M @ xd, rtol=0, atol=1e-12)
    # Residual evaluation type gives the same f and touches no matrix
    tr = oracle.make_terms(eval_type=0, source_id=0, alpha=a, beta=b, mass_dot=1.0)
    fr, _ = _assemble(oracle, d, tr, x=x, xdot=xd)
    assert np.array_equal(fr, f)


def test_source_term_and_threads(oracle):
    (d,), _ = oracle.poisson_problem(6)
    # constant source 1 with multiplier -1: f_i = -(nodal volume); sum = -volume
    f, _ = _assemble(oracle, d, oracle.make_terms(eval_type=0, kappa=0.0, source_id=2))
    assert f.sum() == pytest.approx(-1.0, rel=1e-13)
    # sin source integrates to ~0 over the periodic box and is antisymmetric
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    f1, A1 = _assemble(oracle, d, oracle.make_terms(), x=x)
    f4, A4 = _assemble(oracle, d, oracle.make_terms(nthreads=4), x=x)
    assert np.allclose(f1, f4, rtol=0, atol=1e-13 * np.abs(f1).max())
    assert np.allclose(A1, A4, rtol=0, atol=1e-14 * np.abs(A1).max())
    # workset size does not change the result (chunking only)
    f7, A7 = _assemble(oracle, d, oracle.make_terms(workset_size=7), x=x)
    assert np.array_equal(f1, f7) and np.array_equal(A1, A7)


def test_dirichlet_rows(oracle):
    (d,), _ = oracle.poisson_problem(3)
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    f, A = _assemble(oracle, d, oracle.make_terms(), x=x)
    dofs = np.array([0, 5, 17], np.int32); vals = np.array([0.0, 1.5, -2.0])
    A0 = A.copy()
    oracle.dirichlet(1, dofs, vals, x, f, d["rowptr"], d["colind"], A)
    M = _dense(d, A)
    for l, v in zip(dofs, vals):
        assert f[l] == x[l] - v
        row = M[l].copy(); assert row[l] == 1.0; row[l] = 0; assert not row.any()
    keep = np.ones(d["n_local"], bool); keep[dofs] = False
    assert np.array_equal(_dense(d, A0)[keep], M[keep])      # columns untouched


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_multirank_export_matches_serial(oracle, procs):
    """Owned-row sums over ranks (Export ADD, Panzer_TpetraLinearObjFactory_impl.hpp:170-205)
    reproduce the one-rank assembly GID by GID; ghost import reproduces the owner's x."""
    n = 4
    P = procs[0] * procs[1] * procs[2]
    ranks, dofs = oracle.poisson_problem(n, nranks=P, procs=procs, perturb=0.2)
    (s,), sd = oracle.poisson_problem(n, perturb=0.2)
    # map GIDs of the P-rank numbering to the serial numbering through the stk node ids
    tm = oracle.make_terms()
    node_of_gid = {}
    for d in ranks:
        for g, nd in zip(d["gids"].ravel(), d["elem_nodes"].ravel()):
            node_of_gid[int(g)] = int(nd)
    assert len(node_of_gid) == (n + 1) ** 3
    assert sorted(node_of_gid) == list(range((n + 1) ** 3))       # GIDs are 0..N-1, rank-contiguous
    ser_lid_of_node = dict(zip(s["elem_nodes"].ravel().tolist(), s["lids"].ravel().tolist()))
    x_owned = [oracle.state_by_gid(np.array([node_of_gid[int(g)] for g in d["owned"]])) for d in ranks]
    xs = np.zeros(s["n_local"])
    for d, xo in zip(ranks, x_owned):
        for g, v in zip(d["owned"], xo):
            xs[ser_lid_of_node[node_of_gid[int(g)]]] = v
    fs, As = _assemble(oracle, s, tm, x=xs)
    Ms = _dense(s, As)
    fg, Mg = [], []
    for r, d in enumerate(ranks):
        xg = dofs.global_to_ghost(x_owned, r)
        # ghost entries equal the owner's values
        for l, g in enumerate(np.concatenate([d["owned"], d["ghosted"]])):
            assert xg[l] == xs[ser_lid_of_node[node_of_gid[int(g)]]]
        f, A = _assemble(oracle, d, tm, x=xg)
        fg.append(f); Mg.append(_dense(d, A))
    for r, d in enumerate(ranks):
        fo = dofs.ghost_to_global_vec(fg, r)
        for i, g in enumerate(d["owned"]):
            assert fo[i] == pytest.approx(fs[ser_lid_of_node[node_of_gid[int(g)]]], rel=1e-12, abs=1e-13)
    # matrix: sum of all ranks' ghosted rows, GID by GID
    N = (n + 1) ** 3
    G = np.zeros((N, N))
    for d, M in zip(ranks, Mg):
        g = np.concatenate([d["owned"], d["ghosted"]])
        G[np.ix_(g, g)] += M
    perm = np.array([ser_lid_of_node[node_of_gid[g]] for g in range(N)])
    assert np.allclose(G, Ms[np.ix_(perm, perm)], rtol=0, atol=1e-13 * np.abs(Ms).max())


def test_neumann_flux_integrates_the_side_area(oracle):
    """TianXin::Flux on a side workset adds val * int_side phi_b; summed over the nodes that is val * area, whatever
    the interior node positions (perturbing interior nodes leaves the boundary faces planar rectangles)."""
    import numpy as np
    box = (0.0, 2.0, 0.0, 1.0, 0.0, 3.0)
    (d,), _ = oracle.poisson_problem((4, 3, 2), perturb=0.2, box=box)
    p = oracle.mesh_params((4, 3, 2), (1, 1, 1), box)
    for name, area in (("left", 3.0), ("right", 3.0), ("bottom", 6.0), ("top", 6.0), ("back", 2.0), ("front", 2.0)):
        cells, sides = oracle.sideset_sides(p, d["elem_ids"], name)
        f = np.zeros(d["n_local"])
        oracle.neumann_flux(cells, sides, np.full(len(cells), 2.5), d["lids"], d["cell_coords"], f)
        assert abs(f.sum() - 2.5 * area) < 1e-12
        assert np.count_nonzero(f) == {"left": 12, "right": 12, "bottom": 15, "top": 15, "back": 20, "front": 20}[name]


def test_functional_response_identities(oracle):
    """Integrator_Scalar + Response_Functional: the integral of 1 is the volume on any mesh, a trilinear field is
    integrated exactly by the 2-point rule, and the L2 / H1 errors of the nodal interpolant of the exact solution
    fall like h^2 / h (the acceptance check of the reference's Poisson example)."""
    import numpy as np
    xg, wg = oracle.gauss_legendre(6)
    xr, wr = np.polynomial.legendre.leggauss(6)
    assert np.abs(xg - xr).max() < 5e-16 and np.abs(wg - wr).max() < 5e-16
    box = (0.0, 2.0, -1.0, 1.0, 0.0, 0.5)
    (d,), _ = oracle.poisson_problem((5, 4, 3), perturb=0.2, box=box)
    ones = np.ones(d["n_local"])
    for deg in (2, 10):
        assert abs(oracle.response_functional(1, 1, deg, d["lids"], d["cell_coords"], ones) - 2.0) < 1e-13
    errs = []
    for n in (4, 8, 16):
        (d,), _ = oracle.poisson_problem(n)
        xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
        ue = np.sin(2 * np.pi * xyz[:, 0]) * np.sin(2 * np.pi * xyz[:, 1]) * np.sin(2 * np.pi * xyz[:, 2])
        errs.append((np.sqrt(oracle.response_functional(2, 1, 10, d["lids"], d["cell_coords"], ue)),
                     np.sqrt(oracle.response_functional(3, 1, 10, d["lids"], d["cell_coords"], ue))))
    assert 3.0 < errs[0][0] / errs[1][0] < 4.5 and 3.5 < errs[1][0] / errs[2][0] < 4.5
    assert 1.8 < errs[0][1] / errs[1][1] < 2.6 and 1.8 < errs[1][1] / errs[2][1] < 2.3


def test_functional_response_partitions_over_ranks(oracle):
    """Response_Functional sums rank-local cell integrals (each element is owned by exactly one rank) and reduces:
    the per-rank values of a 2x2x1 decomposition add up to the serial value."""
    import numpy as np
    n = (6, 4, 3)
    (s,), _ = oracle.poisson_problem(n, perturb=0.15)
    node_val = {}
    xs = oracle.state_by_gid(np.arange(s["n_local"])) * 0.05
    for lid, node in zip(s["lids"].ravel(), s["elem_nodes"].ravel()):
        node_val[int(node)] = xs[lid]
    ref = oracle.response_functional(3, 1, 6, s["lids"], s["cell_coords"], xs)
    ranks, _ = oracle.poisson_problem(n, nranks=4, procs=(2, 2, 1), perturb=0.15)
    total, ncells = 0.0, 0
    for d in ranks:
        x = np.zeros(d["n_local"])
        for lid, node in zip(d["lids"].ravel(), d["elem_nodes"].ravel()):
            x[lid] = node_val[int(node)]
        total += oracle.response_functional(3, 1, 6, d["lids"], d["cell_coords"], x)
        ncells += d["lids"].shape[0]
    assert ncells == s["lids"].shape[0]
    assert abs(total - ref) <= 1e-12 * abs(ref)
