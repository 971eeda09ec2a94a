"""Host mirror (libtxhost.so) vs the oracle and the reference's golden vectors.
Bar: GIDs, LIDs, owned/ghosted order and the sparsity graph are bit-exact."""
import numpy as np
import pytest

from tianxin_b200 import host


def _build(n, procs, nfields=1):
    P = procs[0] * procs[1] * procs[2]
    fac = host.CubeHexMeshFactory(**{"X Elements": n[0], "Y Elements": n[1], "Z Elements": n[2],
                                     "X Procs": procs[0], "Y Procs": procs[1], "Z Procs": procs[2]})
    meshes = [fac.buildMesh(r, P) for r in range(P)]
    dms = []
    for r, m in enumerate(meshes):
        dm = host.DOFManager(r, P)
        dm.setConnManager(m.getConnectivity())
        for f in range(nfields):
            dm.addField(f"f{f}")
        dms.append(dm)
    host.DOFManager.buildGlobalUnknownsSim(dms)
    return meshes, dms


@pytest.mark.parametrize("n,procs,nf", [((4, 4, 4), (1, 1, 1), 1), ((5, 4, 3), (2, 1, 1), 1), ((4, 5, 6), (2, 2, 1), 3),
                                        ((4, 4, 4), (2, 2, 2), 1), ((7, 3, 5), (3, 1, 2), 2)])
def test_mesh_dofs_graph_match_oracle_bit_exact(oracle, n, procs, nf):
    P = procs[0] * procs[1] * procs[2]
    meshes, dms = _build(n, procs, nf)
    p = oracle.mesh_params(n, procs)
    conns = []
    for r in range(P):
        ids, nodes, X = oracle.mesh_build(p, r)
        assert np.array_equal(ids, meshes[r].elem_ids())
        assert np.array_equal(nodes, meshes[r].elem_nodes())
        assert np.array_equal(X, meshes[r].cell_vertex_coordinates())
        conns.append(nodes - 1)
    od = oracle.Dofs(conns, nf)
    for r in range(P):
        assert np.array_equal(od.elem_gids(r), dms[r].getElementGIDs())
        assert np.array_equal(od.owned(r), dms[r].getOwnedIndices())
        assert np.array_equal(od.ghosted(r), dms[r].getGhostedIndices())
        assert np.array_equal(od.elem_lids(r), dms[r].getLIDs())
        lof = host.TpetraLinearObjFactory(dms[r])
        rp, ci = lof.getGhostedGraph()
        orp, oci = oracle.ghosted_graph(od.elem_lids(r), od.n_local(r))
        assert np.array_equal(rp, orp) and np.array_equal(ci, oci)


def test_reference_golden_gids_through_host_mirror():
    # adapters-stk/test/stk_connmngr/tCubeHexMeshDOFManager.cpp:147-198 (fields ux, uy, p)
    meshes, dms = _build((2, 2, 2), (1, 1, 1), 3)
    g = dms[0].getElementGIDs()
    assert list(g[0, :12]) == [0, 1, 2, 3, 4, 5, 12, 13, 14, 9, 10, 11]
    e = int(np.where(meshes[0].elem_ids() == 5)[0][0])
    assert list(g[e, 12:]) == [54, 55, 56, 57, 58, 59, 66, 67, 68, 63, 64, 65]
    meshes, dms = _build((2, 2, 2), (2, 1, 1), 3)
    e = int(np.where(meshes[0].elem_ids() == 7)[0][0])
    assert list(meshes[0].getConnectivity()[e]) == [12, 13, 16, 15, 21, 22, 25, 24]
    e = int(np.where(meshes[1].elem_ids() == 2)[0][0])
    assert list(meshes[1].getConnectivity()[e]) == [1, 2, 5, 4, 10, 11, 14, 13]
    for dm in dms:   # :196-211
        o, og = dm.getOwnedIndices(), dm.getOwnedAndGhostedIndices()
        assert np.array_equal(o, og[:len(o)]) and not set(og[len(o):]) & set(o)
    assert list(dms[0].getGIDFieldOffsets(2)) == [2, 5, 8, 11, 14, 17, 20, 23]


def test_api_errors_mirror_reference():
    with pytest.raises(host.TxhostError):      # product of procs must equal the communicator size (:134-136)
        host.CubeHexMeshFactory(**{"X Procs": 3, "Y Procs": 1, "Z Procs": 1}).buildMesh(0, 2)
    with pytest.raises(host.TxhostError):
        host.CubeHexMeshFactory(**{"Bogus": 1})
    dm = host.DOFManager()
    dm.setConnManager(np.zeros((1, 8), np.int64) + np.arange(8))
    dm.addField("T")
    dm.buildGlobalUnknowns()
    with pytest.raises(host.TxhostError):      # Panzer_DOFManager.cpp:492-494
        dm.buildGlobalUnknowns()
    with pytest.raises(host.TxhostError):
        dm.addField("late")


def test_sidesets():
    fac = host.CubeHexMeshFactory(**{"X Elements": 3, "Y Elements": 2, "Z Elements": 2, "X Procs": 2})
    m0, m1 = fac.buildMesh(0, 2), fac.buildMesh(1, 2)
    assert len(m0.sideset_nodes("left")) == 9 and len(m1.sideset_nodes("left")) == 0
    assert len(m1.sideset_nodes("right")) == 9 and len(m0.sideset_nodes("right")) == 0
    assert len(m0.sideset_nodes("top")) == 3 * 3 and len(m1.sideset_nodes("top")) == 2 * 3
    # ids follow nz*(NY+1)*(NX+1)+ny*(NX+1)+nx+1
    assert list(m0.sideset_nodes("left")) == [1 + 4 * j + 12 * k for k in range(3) for j in range(3)]


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 1), (2, 2, 2)])
def test_import_export_plans(oracle, procs):
    """Plans reproduce Tpetra Import(INSERT)/Export(ADD): applied in numpy to the oracle's per-rank
    assembly they give the one-rank matrix and vector, GID by GID, including the remote-only columns
    of shared rows (buildGraph's Export INSERT)."""
    n = (4, 4, 4)
    P = procs[0] * procs[1] * procs[2]
    meshes, dms = _build(n, procs)
    lofs = [host.TpetraLinearObjFactory(d) for d in dms]
    host.TpetraLinearObjFactory.buildPlansSim(lofs)
    plans = [l.plan() for l in lofs]
    N = 5 ** 3
    rng = np.random.default_rng(3)
    xg = rng.standard_normal(N)                      # global solution by GID
    # --- import: pack on the owner, unpack on the ghosting rank
    x_loc = []
    for r in range(P):
        x = np.full(dms[r].num_local, np.nan)
        x[:dms[r].num_owned] = xg[dms[r].getOwnedIndices()]
        x_loc.append(x)
    for r in range(P):
        pr = plans[r]
        for k, s in enumerate(pr["nbr_rank"]):
            ps = plans[s]; ks = list(ps["nbr_rank"]).index(r)
            sent = x_loc[s][ps["send_lids"][ps["send_off"][ks]:ps["send_off"][ks + 1]]]
            x_loc[r][pr["recv_lids"][pr["recv_off"][k]:pr["recv_off"][k + 1]]] = sent
    for r in range(P):
        assert np.array_equal(x_loc[r], xg[dms[r].getOwnedAndGhostedIndices()])
    # --- assemble per rank with the oracle on the FILL graph, then export
    tm = oracle.make_terms(source_id=0)
    f_loc, A_loc = [], []
    for r in range(P):
        t = oracle.tables_build(meshes[r].cell_vertex_coordinates())
        pr = plans[r]
        f = np.zeros(dms[r].num_local); A = np.zeros(pr["rowptr"][-1])
        oracle.evaluate_volume(tm, dms[r].getLIDs(), t, x_loc[r], None, pr["rowptr"], pr["colind"], f, A)
        f_loc.append(f); A_loc.append(A)
    f_pre = [f.copy() for f in f_loc]; A_pre = [A.copy() for A in A_loc]
    for r in range(P):
        pr = plans[r]
        for k, s in enumerate(pr["nbr_rank"]):
            ps = plans[s]; ks = list(ps["nbr_rank"]).index(r)
            ghost_rows = ps["recv_lids"][ps["recv_off"][ks]:ps["recv_off"][ks + 1]]          # rows s ghosts from me
            mine = pr["send_lids"][pr["send_off"][k]:pr["send_off"][k + 1]]
            assert len(ghost_rows) == len(mine)
            np.add.at(f_loc[r], mine, f_pre[s][ghost_rows])
            vals = np.concatenate([A_pre[s][ps["rowptr"][g]:ps["rowptr"][g + 1]] for g in ghost_rows]) if len(ghost_rows) else np.zeros(0)
            pos = pr["mat_recv_pos"][pr["mat_recv_off"][k]:pr["mat_recv_off"][k + 1]]
            assert len(vals) == len(pos) and (pos >= 0).all()
            np.add.at(A_loc[r], pos, vals)
    # --- one-rank reference, mapped through the stk node ids
    (s,), _ = oracle.poisson_problem(n[0])
    node_of_gid = {}
    for r in range(P):
        for g, nd in zip(dms[r].getElementGIDs().ravel(), meshes[r].elem_nodes().ravel()):
            node_of_gid[int(g)] = int(nd)
    ser_lid = dict(zip(s["elem_nodes"].ravel().tolist(), s["lids"].ravel().tolist()))
    perm = np.array([ser_lid[node_of_gid[g]] for g in range(N)])
    xs = np.zeros(N); xs[perm] = xg
    ts = oracle.tables_build(s["cell_coords"])
    fs = np.zeros(N); As = np.zeros(s["rowptr"][-1])
    oracle.evaluate_volume(tm, s["lids"], ts, xs, None, s["rowptr"], s["colind"], fs, As)
    Ms = np.zeros((N, N))
    for i in range(N):
        Ms[i, s["colind"][s["rowptr"][i]:s["rowptr"][i + 1]]] = As[s["rowptr"][i]:s["rowptr"][i + 1]]
    for r in range(P):
        pr = plans[r]; no = dms[r].num_owned
        owned = dms[r].getOwnedIndices()
        assert np.allclose(f_loc[r][:no], fs[perm[owned]], rtol=0, atol=1e-13)
        for i in range(no):
            cols = pr["col_gids"][pr["colind"][pr["rowptr"][i]:pr["rowptr"][i + 1]]]
            row = A_loc[r][pr["rowptr"][i]:pr["rowptr"][i + 1]]
            want = Ms[perm[owned[i]], perm[cols]]
            assert np.allclose(row, want, rtol=0, atol=1e-13)
            # the fill-graph row holds every non-zero of the global row
            assert np.count_nonzero(Ms[perm[owned[i]]]) <= len(cols)


@pytest.mark.parametrize("procs", [(2, 1, 1), (2, 2, 2)])
def test_compact_plans_match_full_plans(procs):
    """Compact mode (only the ghost rows of the graph reach the host; txasm_graph_merge_columns inserts the received
    pairs on the device): same halo lists, and the pairs merged into the ghosted graph -- here in numpy -- give the
    fill graph and the matrix positions of the full host path."""
    n = (4, 4, 3)
    _, dms = _build(n, procs)
    full = [host.TpetraLinearObjFactory(d) for d in dms]
    host.TpetraLinearObjFactory.buildPlansSim(full)
    comp = [host.TpetraLinearObjFactory(d) for d in dms]
    ghosted = [l.getGhostedGraph() for l in full]
    for l, d, (rp, ci) in zip(comp, dms, ghosted):
        no, nl = d.num_owned, d.num_local
        l.setGhostRows(rp[no:] - rp[no], ci[rp[no]:rp[nl]])
    host.TpetraLinearObjFactory.buildPlansSim(comp)
    for lf, lc, d, (rp, ci) in zip(full, comp, dms, ghosted):
        pf, pc = lf.plan(), lc.plan()
        for k in ("nbr_rank", "send_off", "send_lids", "recv_off", "recv_lids", "col_gids", "mat_recv_off", "pair_rows", "pair_cols"):
            assert np.array_equal(pf[k], pc[k]), k
        assert "rowptr" not in pc
        rows = [set(ci[rp[i]:rp[i + 1]].tolist()) for i in range(d.num_local)]
        for r, c in zip(pc["pair_rows"], pc["pair_cols"]):
            assert r < d.num_owned
            rows[r].add(int(c))
        frp = np.concatenate([[0], np.cumsum([len(s) for s in rows])])
        fci = np.concatenate([np.array(sorted(s), np.int32) for s in rows])
        assert np.array_equal(frp, pf["rowptr"]) and np.array_equal(fci, pf["colind"])
        pos = np.array([frp[r] + sorted(rows[r]).index(int(c)) for r, c in zip(pc["pair_rows"], pc["pair_cols"])], np.int64)
        assert np.array_equal(pos, pf["mat_recv_pos"])


@pytest.mark.parametrize("procs,nf", [((2, 1, 1), 1), ((2, 2, 1), 2)])
def test_dof_manager_sparse_ids_take_the_sort_path_and_agree(oracle, procs, nf):
    """The overlap map and the directory use a bitmap / a direct table when the ids are dense (a mesh) and fall back to
    sort + unique + binary search when they are not.  Ids stretched by 1000 force the fall-back: same owners, same
    LIDs, and GIDs that map one to one (the numbering follows the id ORDER, which the stretch keeps), and both paths match
    the oracle's GUN restatement."""
    n = (5, 4, 3)
    P = procs[0] * procs[1] * procs[2]
    meshes, dense = _build(n, procs, nf)
    sparse = []
    for r, m in enumerate(meshes):
        dm = host.DOFManager(r, P)
        dm.setConnManager(m.getConnectivity() * 1000 + 7)
        for f in range(nf):
            dm.addField(f"f{f}")
        sparse.append(dm)
    host.DOFManager.buildGlobalUnknownsSim(sparse)
    od = oracle.Dofs([m.elem_nodes() * 1000 - 993 for m in meshes], nf)
    for r in range(P):
        assert np.array_equal(dense[r].getLIDs(), sparse[r].getLIDs())
        assert dense[r].num_owned == sparse[r].num_owned and dense[r].num_local == sparse[r].num_local
        assert np.array_equal(od.elem_gids(r), sparse[r].getElementGIDs())
        assert np.array_equal(od.owned(r), sparse[r].getOwnedIndices())
        assert np.array_equal(od.ghosted(r), sparse[r].getGhostedIndices())


def test_host_thread_count_is_settable():
    n0 = host.set_num_threads(0)            # query
    assert n0 >= 1
    assert host.set_num_threads(2) == 2
    host.set_num_threads(n0)
