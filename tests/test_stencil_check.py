"""The closed-form stencil checker bench.py uses at 256^3 (tools/stencil_check.py) against the oracle on small meshes:
the two restatements are independent (quadrature pipeline vs. 1-D factor products) and must agree to rounding."""
import numpy as np
import pytest

from tools import stencil_check as sc


@pytest.mark.parametrize("n", [5, (4, 7, 3)])
def test_closed_form_matches_oracle(oracle, n):
    (d,), _ = oracle.poisson_problem(n)
    dims = (n, n, n) if isinstance(n, int) else n
    x = oracle.state_by_gid(np.arange(d["n_local"]))
    t = oracle.tables_build(d["cell_coords"])
    f = np.zeros(d["n_local"]); A = np.zeros(d["rowptr"][-1])
    oracle.evaluate_volume(oracle.make_terms(), d["lids"], t, x, None, d["rowptr"], d["colind"], f, A)
    # node lattice positions by LID from the stk node ids (id - 1 = I + (NX+1) (J + (NY+1) K))
    pos = np.zeros((d["n_local"], 3), np.int64)
    ids = d["elem_nodes"].ravel() - 1
    pos[d["lids"].ravel()] = np.stack([ids % (dims[0] + 1), (ids // (dims[0] + 1)) % (dims[1] + 1), ids // ((dims[0] + 1) * (dims[1] + 1))], axis=1)
    on_bnd = lambda p: ((p == 0) | (p == np.asarray(dims)[None, :])).any(axis=1)
    rows = np.arange(d["n_local"])
    # 1. volume fill only
    eA, sA, ef, sf = sc.check_rows(rows, d["rowptr"], d["colind"], A, f, lambda c: pos[c], lambda c: x[c], dims,
                                   lambda p: np.zeros(len(p), np.int64))
    assert eA < 1e-13 * sA and ef < 1e-12 * sf
    # 2. with the six Dirichlet faces (value 0)
    dofs = np.nonzero(on_bnd(pos))[0].astype(np.int32)
    oracle.dirichlet(1, dofs, np.zeros(len(dofs)), x, f, d["rowptr"], d["colind"], A)
    eA, sA, ef, sf = sc.check_rows(rows, d["rowptr"], d["colind"], A, f, lambda c: pos[c], lambda c: x[c], dims,
                                   lambda p: on_bnd(p).astype(np.int64))
    assert eA < 1e-13 * sA and ef < 1e-12 * sf
    # 3. the checker notices a wrong entry
    A[d["rowptr"][len(rows) // 2] + 3] *= 1.0 + 1e-9
    eA, sA, _, _ = sc.check_rows(rows, d["rowptr"], d["colind"], A, f, lambda c: pos[c], lambda c: x[c], dims,
                                 lambda p: on_bnd(p).astype(np.int64))
    assert eA > 1e-12 * sA or A[d["rowptr"][len(rows) // 2] + 3] == 0.0
