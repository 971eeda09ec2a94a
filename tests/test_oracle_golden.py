"""Pin the CPU oracle against every golden vector / known-answer test the reference's own test
suite holds for the assembly hot path.  Each test names the reference
test file:line whose assertions it replays."""
import numpy as np
import pytest


# --------------------------------------------------------------------------------------------
# adapters-stk/test/stk_connmngr/tCubeHexMeshDOFManager.cpp:96-211  (2x2x2 hex, fields ux,uy,p)
# --------------------------------------------------------------------------------------------
def _hex_conns(orc, n, nranks):
    p = orc.mesh_params(n, (nranks, 1, 1))     # "X Procs"=-1 -> x-only decomposition (:128-133)
    out = []
    for r in range(nranks):
        ids, nodes, _ = orc.mesh_build(p, r)
        out.append((ids, nodes - 1))
    return out


def test_cubehex_dofmanager_one_rank_gids(oracle):
    (ids, conn), = _hex_conns(oracle, 2, 1)
    d = oracle.Dofs([conn], nfields=3)
    g = d.elem_gids(0)
    assert g.shape == (8, 24)
    # :147-157 element 0
    assert list(g[0, :12]) == [0, 1, 2, 3, 4, 5, 12, 13, 14, 9, 10, 11]
    assert list(g[0, 12:]) == [27, 28, 29, 30, 31, 32, 39, 40, 41, 36, 37, 38]
    # :160-172 element with stk id 5 (elementLocalId(5))
    e = int(np.where(ids == 5)[0][0])
    assert list(g[e, :12]) == [27, 28, 29, 30, 31, 32, 39, 40, 41, 36, 37, 38]
    assert list(g[e, 12:]) == [54, 55, 56, 57, 58, 59, 66, 67, 68, 63, 64, 65]


def test_cubehex_two_rank_connectivity(oracle):
    (ids0, c0), (ids1, c1) = _hex_conns(oracle, 2, 2)
    # :174-183 rank 0, element id 7
    e = int(np.where(ids0 == 7)[0][0])
    assert list(c0[e]) == [12, 13, 16, 15, 21, 22, 25, 24]
    # :184-193 rank 1, element id 2
    e = int(np.where(ids1 == 2)[0][0])
    assert list(c1[e]) == [1, 2, 5, 4, 10, 11, 14, 13]
    # :196-211 owned is a prefix of owned_and_ghosted, ghosted disjoint from owned
    d = oracle.Dofs([c0, c1], nfields=3)
    for r in range(2):
        o, og = d.owned(r), d.owned_and_ghosted(r)
        assert len(o) <= len(og) and np.array_equal(o, og[:len(o)])
        assert not set(og[len(o):]) & set(o)


# --------------------------------------------------------------------------------------------
# adapters-stk/test/stk_connmngr/tSquareQuadMeshDOFManager.cpp   (2x2 quads, 2 ranks, x slabs)
# mesh numbering: Panzer_STK_SquareQuadMeshFactory.cpp:369-390
# --------------------------------------------------------------------------------------------
def _quad_conns(nx, ny, nranks):
    conns = []
    per = nx // nranks
    for r in range(nranks):
        c = []
        for j in range(ny):
            for i in range(r * per, (r + 1) * per):
                n0 = i + 1 + j * (nx + 1)
                c.append([n0 - 1, n0, n0 + nx + 1, n0 + nx])      # nodes[0..3] - 1
        conns.append(np.array(c, np.int64))
    return conns


def test_squarequad_three_fields_gids(oracle):
    d = oracle.Dofs(_quad_conns(2, 2, 2), nfields=3)
    g0, g1 = d.elem_gids(0), d.elem_gids(1)
    # :145-172 rank 0
    assert list(g0[0]) == [0, 1, 2, 3, 4, 5, 9, 10, 11, 6, 7, 8]
    assert list(g0[1]) == [6, 7, 8, 9, 10, 11, 15, 16, 17, 12, 13, 14]
    # :174-201 rank 1
    assert list(g1[0]) == [3, 4, 5, 18, 19, 20, 21, 22, 23, 9, 10, 11]
    assert list(g1[1]) == [9, 10, 11, 21, 22, 23, 24, 25, 26, 15, 16, 17]
    # :140-142,155-159 offsets: field f of basis b sits at b*3+f, p<ux<uy
    p, ux, uy = d.field_offsets(0), d.field_offsets(1), d.field_offsets(2)
    for i in range(4):
        assert g0[0][p[i]] < g0[0][ux[i]] < g0[0][uy[i]]


def test_squarequad_owned_ghosted(oracle):
    d = oracle.Dofs(_quad_conns(2, 2, 2), nfields=1)
    # :328-351 rank 0
    assert sorted(d.owned(0)) == [0, 1, 2, 3, 4, 5]
    assert sorted(d.owned_and_ghosted(0)) == [0, 1, 2, 3, 4, 5]
    # :352-373 rank 1
    assert sorted(d.owned(1)) == [6, 7, 8]
    assert sorted(d.owned_and_ghosted(1)) == [1, 3, 5, 6, 7, 8]


def test_squarequad_single_field_gids(oracle):
    # :412-446 (dofManager_temp: one Q1 field "T")
    d = oracle.Dofs(_quad_conns(2, 2, 2), nfields=1)
    g0, g1 = d.elem_gids(0), d.elem_gids(1)
    assert list(g0[0]) == [0, 1, 3, 2] and list(g0[1]) == [2, 3, 5, 4]
    assert list(g1[0]) == [1, 6, 7, 3] and list(g1[1]) == [3, 7, 8, 5]


# --------------------------------------------------------------------------------------------
# disc-fe/test/core_tests/basis_values2.cpp:264-286 (Q1 identities, replayed in 3-D on 1/2-cubes)
# disc-fe/test/core_tests/integration_values2.cpp:106-115
# --------------------------------------------------------------------------------------------
def test_basis_values_identities(oracle):
    # four cells of edge 1/2 laid out like :213-236, extruded in z
    cells = []
    for c in range(4):
        xl, yl = c % 2, c // 2
        X = np.array([[xl, yl, 0], [xl + 1, yl, 0], [xl + 1, yl + 1, 0], [xl, yl + 1, 0],
                      [xl, yl, 1], [xl + 1, yl, 1], [xl + 1, yl + 1, 1], [xl, yl + 1, 1]], float) * 0.5
        cells.append(X)
    t = oracle.tables_build(np.array(cells))
    pts, wts = oracle.ref_cubature()
    rel_vol = 0.25 ** 3                      # relCellVol (:262), 3-D
    for q in range(8):
        x, y, z = pts[q]
        val, grad = oracle.ref_basis(pts[q])
        # :268-270 reference values of phi_0
        assert val[0] == pytest.approx(-0.125 * (x - 1) * (y - 1) * (z - 1), abs=1e-16)
        assert grad[0, 0] == pytest.approx(-0.125 * (y - 1) * (z - 1), abs=1e-16)
        # The reference asserts these with TEST_EQUALITY in 2-D; the 3-D sums over 8 nodes at the
        # irrational Gauss abscissae round differently, so they are held to a few ulp here.
        eq = lambda a, b: a == pytest.approx(b, rel=4e-15, abs=1e-17)
        for c in range(4):
            assert eq(t.jac_det[c, q], rel_vol)                                            # :275
            assert t.basis[c, 0, q] == val[0]                                              # :278
            assert eq(t.wbasis[c, 0, q], rel_vol * wts[q] * t.basis[c, 0, q])              # :279
            for d in range(3):
                assert eq(t.gbasis[c, 0, q, d], 4.0 * grad[0, d])                          # :281-282
                assert eq(t.wgbasis[c, 0, q, d], rel_vol * wts[q] * t.gbasis[c, 0, q, d])  # :284-285


def test_integration_values_point(oracle):
    # unit cell (:97-105): a cubature point sits at ((1/sqrt3 + 1)/2, ...) to 1e-8 (:109-115)
    X = np.array([[[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]]], float)
    t = oracle.tables_build(X)
    assert t.ip.shape == (1, 8, 3)
    target = (1.0 / np.sqrt(3.0) + 1.0) / 2.0
    assert np.isclose(t.ip[0], target, rtol=1e-8).all(axis=1).any()
    assert set(np.round(t.ip[0].ravel(), 12)) == {round(target, 12), round(1 - target, 12)}
    pts, wts = oracle.ref_cubature()
    assert wts.sum() == 8.0 and np.allclose(np.abs(pts), 1 / np.sqrt(3))
    # partition of unity / Kronecker delta of the C1 hex basis at the Shards vertices
    S = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    for i in range(8):
        v, g = oracle.ref_basis(S[i])
        assert np.array_equal(v, np.eye(8)[i])
    v, g = oracle.ref_basis(np.array([0.3, -0.2, 0.7]))
    assert v.sum() == pytest.approx(1.0, abs=1e-15) and np.allclose(g.sum(axis=0), 0, atol=1e-15)


# --------------------------------------------------------------------------------------------
# mesh factory facts: Panzer_STK_CubeHexMeshFactory.cpp:89-133, 463-535, 918-927
# adapters-stk/test/stk_interface_test/tCubeHexMeshFactory.cpp (counts)
# --------------------------------------------------------------------------------------------
def test_mesh_factory_counts_and_partition(oracle):
    assert oracle.default_proc_grid(8) == (2, 2, 2)
    assert oracle.default_proc_grid(1) == (1, 1, 1)
    assert sorted(oracle.default_proc_grid(4)) == [1, 2, 2]
    assert sorted(oracle.default_proc_grid(2)) == [1, 1, 2]
    p = oracle.mesh_params((5, 4, 3), (2, 2, 1))
    tot, seen = 0, set()
    for r in range(4):
        ids, nodes, X = oracle.mesh_build(p, r)
        assert np.all(np.diff(ids) > 0)
        tot += len(ids); seen |= set(ids)
        # positive Jacobian, element volume = hx*hy*hz
        t = oracle.tables_build(X)
        assert np.allclose(t.wm.sum(axis=1), (1 / 5) * (1 / 4) * (1 / 3), rtol=1e-14)
    assert tot == 60 and seen == set(range(1, 61))
    # first "extra" procs get the extra layer (:474-484): nx=5 on 2 procs -> 3 + 2
    assert len(oracle.mesh_build(p, 0)[0]) == 3 * 2 * 3 and len(oracle.mesh_build(p, 1)[0]) == 2 * 2 * 3
