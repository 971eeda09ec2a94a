"""Identities that pin the oracle's element blocks beyond the scalar Q1 hexahedron (oracle/txblocks.c): the third-party
conventions (Shards node orders, Intrepid2 bases, cubature) are not in the reference tree, so they are checked the way
SURVEY.md appendix C asks for -- Kronecker deltas, partition of unity, exactness, null spaces, energies."""
import numpy as np
import pytest
import scipy.sparse as sp


def _csr(orc, lids, n_rows):
    return orc.ghosted_graph(lids, n_rows)


def _mat(A, rowptr, colind):
    n = len(rowptr) - 1
    return sp.csr_matrix((A, colind, rowptr), shape=(n, n))


HEX27 = np.array([[-1,-1,-1],[1,-1,-1],[1,1,-1],[-1,1,-1],[-1,-1,1],[1,-1,1],[1,1,1],[-1,1,1],[0,-1,-1],[1,0,-1],[0,1,-1],[-1,0,-1],
                  [-1,-1,0],[1,-1,0],[1,1,0],[-1,1,0],[0,-1,1],[1,0,1],[0,1,1],[-1,0,1],[0,0,0],[0,0,-1],[0,0,1],[-1,0,0],[1,0,0],[0,-1,0],[0,1,0]], float)
TET10 = np.array([[0,0,0],[1,0,0],[0,1,0],[0,0,1],[.5,0,0],[.5,.5,0],[0,.5,0],[0,0,.5],[.5,0,.5],[0,.5,.5]])


def test_nodal_bases_are_kronecker_and_sum_to_one(oracle):
    for elem, nodes in ((oracle.HEX8_C1, HEX27[:8]), (oracle.HEX27_C2, HEX27), (oracle.TET4_C1, TET10[:4]), (oracle.TET10_C2, TET10)):
        V = np.array([oracle.block_ref_basis(elem, p)[0] for p in nodes])
        assert np.allclose(V, np.eye(len(nodes)), atol=1e-14)
        rng = np.random.default_rng(0)
        for p in rng.uniform(0.05, 0.3, (5, 3)):
            v, g = oracle.block_ref_basis(elem, p)
            assert abs(v.sum() - 1) < 1e-14 and np.abs(g.sum(axis=0)).max() < 1e-13
            # gradients by central differences
            for d in range(3):
                e = np.zeros(3); e[d] = 1e-6
                fd = (oracle.block_ref_basis(elem, p + e)[0] - oracle.block_ref_basis(elem, p - e)[0]) / 2e-6
                assert np.allclose(fd, g[:, d], atol=1e-8)


def test_hcurl_basis_traces_and_curls(oracle):
    S = HEX27[:8]
    edges = [(0,1),(1,2),(2,3),(3,0),(4,5),(5,6),(6,7),(7,4),(0,4),(1,5),(2,6),(3,7)]
    for e, (a, b) in enumerate(edges):
        mid, t = 0.5 * (S[a] + S[b]), 0.5 * (S[b] - S[a])
        v, _ = oracle.block_ref_basis(oracle.HEX8_HCURL, mid)
        trace = v @ t
        for e2, (a2, b2) in enumerate(edges):
            mid2, t2 = 0.5 * (S[a2] + S[b2]), 0.5 * (S[b2] - S[a2])
            assert abs(oracle.block_ref_basis(oracle.HEX8_HCURL, mid2)[0][e] @ t2 - (1.0 if e2 == e else 0.0)) < 1e-14
        assert abs(trace[e] - 1) < 1e-14
    p = np.array([0.13, -0.4, 0.27])
    v, c = oracle.block_ref_basis(oracle.HEX8_HCURL, p)
    h = 1e-6
    dv = np.zeros((12, 3, 3))                      # dv[e][i][j] = d v_i / d x_j
    for j in range(3):
        d = np.zeros(3); d[j] = h
        dv[:, :, j] = (oracle.block_ref_basis(oracle.HEX8_HCURL, p + d)[0] - oracle.block_ref_basis(oracle.HEX8_HCURL, p - d)[0]) / (2 * h)
    curl = np.stack([dv[:, 2, 1] - dv[:, 1, 2], dv[:, 0, 2] - dv[:, 2, 0], dv[:, 1, 0] - dv[:, 0, 1]], axis=1)
    assert np.allclose(curl, c, atol=1e-8)


def test_cubature_exactness(oracle):
    from math import factorial
    for deg in (1, 2, 3):
        pts, w = oracle.block_cubature(oracle.TET4_C1, deg)
        assert abs(w.sum() - 1 / 6) < 1e-15
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                for c in range(deg + 1 - a - b):
                    exact = factorial(a) * factorial(b) * factorial(c) / factorial(a + b + c + 3)
                    assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum() - exact) < 1e-15, (deg, a, b, c)
    for deg in (2, 4, 5):
        pts, w = oracle.block_cubature(oracle.HEX27_C2, deg)
        assert len(w) == (deg // 2 + 1) ** 3 and abs(w.sum() - 8) < 1e-14
        assert abs((w * pts[:, 0] ** (2 * (deg // 2)) * pts[:, 1] ** 2).sum() - (2 / (2 * (deg // 2) + 1)) * (2 / 3) * 2) < 1e-13


def _hex_mesh(oracle, n):
    (d,), _ = oracle.poisson_problem(n)
    return d


def test_q2_hex_diffusion_energy_and_null_space(oracle):
    n = 3
    d = _hex_mesh(oracle, n)
    lids = oracle.q2_hex_lids(n)
    M = 2 * n + 1
    nrows = M ** 3
    rp, ci = _csr(oracle, lids, nrows)
    g = np.arange(nrows); X = (g % M) / (M - 1.0); Y = ((g // M) % M) / (M - 1.0); Z = (g // (M * M)) / (M - 1.0)
    u = X ** 2 + Y * Z
    f, A = oracle.block_evaluate(oracle.HEX27_C2, oracle.OP_DIFFUSION, 4, [1.0], d["cell_coords"], lids, rp, ci, u)
    K = _mat(A, rp, ci)
    assert abs(K - K.T).max() < 1e-13
    assert np.abs(K @ np.ones(nrows)).max() < 1e-12
    assert abs(u @ (K @ u) - 2.0) < 1e-12                        # int |grad u|^2 = 4/3 + 1/3 + 1/3, exact for Q2
    assert np.allclose(f, K @ u, atol=1e-12)                     # residual of a linear operator
    interior = (X > 0) & (X < 1) & (Y > 0) & (Y < 1) & (Z > 0) & (Z < 1)
    ulin = 2 * X - Y + 0.5 * Z
    assert np.abs((K @ ulin)[interior]).max() < 1e-12


def test_tet_blocks_and_mixed_mesh(oracle):
    """P1 tets on the CubeTetMeshFactory split; then a mixed mesh: Q1 hexahedra for x < 1/2 and P1 tets for x > 1/2 sharing
    the interface nodes -- the configuration the reference's DOFManager refuses -- assembled as two
    blocks into one matrix.  Linear fields are reproduced by both, so the energy is exact and K u_lin vanishes inside."""
    n = 4
    tn, tc, nnodes = oracle.cube_tet_mesh(n)
    rp, ci = _csr(oracle, tn.astype(np.int32), nnodes)
    xyz = np.zeros((nnodes, 3)); xyz[tn.ravel()] = tc.reshape(-1, 3)
    u = 1.5 * xyz[:, 0] - 2 * xyz[:, 1] + xyz[:, 2]
    f, A = oracle.block_evaluate(oracle.TET4_C1, oracle.OP_DIFFUSION, 1, [1.0], tc, tn.astype(np.int32), rp, ci, u)
    K = _mat(A, rp, ci)
    assert abs(u @ (K @ u) - (1.5 ** 2 + 4 + 1)) < 1e-12 and abs(K - K.T).max() < 1e-13
    # mixed: hexes of the left half (cells with centroid x < 1/2) + tets of the right half
    d = _hex_mesh(oracle, n)
    hex_nodes = (d["elem_nodes"] - 1).astype(np.int32)
    left = d["cell_coords"].mean(axis=1)[:, 0] < 0.5
    right = tc.mean(axis=1)[:, 0] > 0.5
    hl, tl = hex_nodes[left], tn[right].astype(np.int32)
    import itertools
    rows = np.concatenate([np.repeat(hl, 8, axis=1).ravel(), np.repeat(tl, 4, axis=1).ravel()])
    cols = np.concatenate([np.tile(hl, (1, 8)).ravel(), np.tile(tl, (1, 4)).ravel()])
    G = sp.csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(nnodes, nnodes)); G.sum_duplicates(); G.sort_indices()
    rp2, ci2 = G.indptr.astype(np.int64), G.indices.astype(np.int32)
    f = np.zeros(nnodes); A = np.zeros(rp2[-1])
    oracle.block_evaluate(oracle.HEX8_C1, oracle.OP_DIFFUSION, 2, [1.0], d["cell_coords"][left], hl, rp2, ci2, u, f=f, A=A)
    oracle.block_evaluate(oracle.TET4_C1, oracle.OP_DIFFUSION, 1, [1.0], tc[right], tl, rp2, ci2, u, f=f, A=A)
    K = _mat(A, rp2, ci2)
    used = np.zeros(nnodes, bool); used[hl.ravel()] = True; used[tl.ravel()] = True
    inside = used & np.all((xyz > 1e-9) & (xyz < 1 - 1e-9), axis=1)
    inside[nnodes - n ** 3:] = used[nnodes - n ** 3:]              # centroid nodes of the tets are interior
    assert abs(u @ (K @ u) - (1.5 ** 2 + 4 + 1)) < 1e-12
    assert np.abs((K @ u)[inside]).max() < 1e-12
    # P2 tets: quadratic energy exact with the degree-2 rule
    from collections import OrderedDict
    edge_id = OrderedDict()
    l10 = np.zeros((len(tn), 10), np.int32); l10[:, :4] = tn
    for c, t in enumerate(tn):
        for e, (i, j) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
            key = (min(t[i], t[j]), max(t[i], t[j]))
            l10[c, 4 + e] = edge_id.setdefault(key, nnodes + len(edge_id))
    n10 = nnodes + len(edge_id)
    xyz10 = np.zeros((n10, 3)); xyz10[:nnodes] = xyz
    for (a, b), k in edge_id.items():
        xyz10[k] = 0.5 * (xyz[a] + xyz[b])
    rp3, ci3 = _csr(oracle, l10, n10)
    u2 = xyz10[:, 0] ** 2 + xyz10[:, 1] * xyz10[:, 2]
    _, A = oracle.block_evaluate(oracle.TET10_C2, oracle.OP_DIFFUSION, 2, [1.0], tc, l10, rp3, ci3, u2)
    K = _mat(A, rp3, ci3)
    assert abs(u2 @ (K @ u2) - 2.0) < 1e-12 and np.abs(K @ np.ones(n10)).max() < 1e-12


def test_elastodynamics_identities(oracle):
    """Three interleaved HGRAD fields (FieldAggPattern order): rigid-body modes span the null space of K, uniform strain
    gives the exact energy, mass rows sum to rho times the nodal volume, J = gamma rho M + beta K."""
    n = (3, 2, 2)
    d = _hex_mesh(oracle, n)
    nn = d["n_local"]
    l3 = (3 * d["lids"][:, :, None] + np.arange(3)[None, None, :]).reshape(len(d["lids"]), 24).astype(np.int32)
    rp, ci = _csr(oracle, l3, 3 * nn)
    xyz = np.zeros((nn, 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    lam, mu, rho = 1.3, 0.7, 2.5
    def K_of(u, **kw):
        return oracle.block_evaluate(oracle.HEX8_C1, oracle.OP_ELASTICITY, 2, [lam, mu, 0.0, 0.0], d["cell_coords"], l3, rp, ci, u.ravel().copy(), **kw)
    modes = [np.tile(e, (nn, 1)) for e in np.eye(3)] + [np.cross(w, xyz) for w in np.eye(3)]
    f, A = K_of(modes[0])
    K = _mat(A, rp, ci)
    assert abs(K - K.T).max() < 1e-13
    for m in modes:
        assert np.abs(K @ m.ravel()).max() < 1e-12
    u = np.zeros((nn, 3)); u[:, 0] = xyz[:, 0]                    # eps_xx = 1
    assert abs(u.ravel() @ (K @ u.ravel()) - (lam + 2 * mu)) < 1e-12
    u = np.zeros((nn, 3)); u[:, 0] = xyz[:, 1]                    # simple shear: eps_xy = 1/2 -> energy mu
    assert abs(u.ravel() @ (K @ u.ravel()) - mu) < 1e-12
    # mass and the Jacobian seeds
    rng = np.random.default_rng(4)
    x, xdd = rng.standard_normal(3 * nn), rng.standard_normal(3 * nn)
    gamma, beta = 4.0, 0.5
    f, A = oracle.block_evaluate(oracle.HEX8_C1, oracle.OP_ELASTICITY, 2, [lam, mu, rho, 0.0], d["cell_coords"], l3, rp, ci, x,
                                 xdotdot=xdd, beta=beta, gamma=gamma)
    Mm = (_mat(A, rp, ci) - beta * K) / (gamma * rho)
    vol = np.zeros(nn); t = oracle.tables_build(d["cell_coords"]); np.add.at(vol, d["lids"].ravel(), t.wbasis.sum(axis=2).ravel())
    assert np.allclose(np.asarray(Mm.sum(axis=1)).ravel(), np.repeat(vol, 3), atol=1e-13)
    assert np.allclose(f, K @ x + rho * (Mm @ xdd), atol=1e-12)


def test_hcurl_curlcurl_null_space_and_mass(oracle):
    """Curl-curl annihilates gradients of Q1 functions (dof of a global edge = half the potential difference along it),
    the mass matrix integrates constant fields exactly, both with the edge orientation signs."""
    n = (3, 2, 2)
    nx, ny, nz = n
    d = _hex_mesh(oracle, n)
    lids, signs = oracle.hcurl_hex_lids(n)
    ne = nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz
    assert lids.max() == ne - 1 and len(np.unique(lids)) == ne
    rp, ci = _csr(oracle, lids, ne)
    x0 = np.zeros(ne)
    _, C = oracle.block_evaluate(oracle.HEX8_HCURL, oracle.OP_CURLCURL, 2, [1.0, 0.0, 0.0], d["cell_coords"], lids, rp, ci, x0, signs=signs)
    _, Mv = oracle.block_evaluate(oracle.HEX8_HCURL, oracle.OP_CURLCURL, 2, [0.0, 1.0, 0.0], d["cell_coords"], lids, rp, ci, x0, signs=signs)
    C, Mv = _mat(C, rp, ci), _mat(Mv, rp, ci)
    assert abs(C - C.T).max() < 1e-13 and abs(Mv - Mv.T).max() < 1e-13
    # potential at the nodes -> edge dofs through the element tables (edge e of a cell runs between its Shards vertices)
    nodes = (d["elem_nodes"] - 1)
    xyz = np.zeros((d["n_local"], 3)); xyz[d["lids"].ravel()] = d["cell_coords"].reshape(-1, 3)
    node_xyz = np.zeros((nodes.max() + 1, 3)); node_xyz[nodes.ravel()] = d["cell_coords"].reshape(-1, 3)
    phi = np.sin(node_xyz[:, 0] * 2) + node_xyz[:, 1] * node_xyz[:, 2] - node_xyz[:, 0] * node_xyz[:, 1] * node_xyz[:, 2]
    edges = [(0,1),(1,2),(2,3),(3,0),(4,5),(5,6),(6,7),(7,4),(0,4),(1,5),(2,6),(3,7)]
    dof = np.zeros(ne)
    for e, (a, b) in enumerate(edges):
        dof[lids[:, e]] = signs[:, e] * 0.5 * (phi[nodes[:, b]] - phi[nodes[:, a]])
    assert np.abs(C @ dof).max() < 1e-12
    # constant field (1, 2, 3): dof of an x-edge = h_x / 2 * 1 ..., energy = |E|^2 * volume
    E = np.array([1.0, 2.0, 3.0]); dofc = np.zeros(ne)
    for e, (a, b) in enumerate(edges):
        dofc[lids[:, e]] = signs[:, e] * 0.5 * ((node_xyz[nodes[:, b]] - node_xyz[nodes[:, a]]) @ E)
    assert abs(dofc @ (Mv @ dofc) - 14.0) < 1e-12
    assert np.abs(C @ dofc).max() < 1e-12
