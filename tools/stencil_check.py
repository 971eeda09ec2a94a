"""Closed-form check of an assembled Poisson system on an inline (uniform, axis-aligned) cube mesh.

Independent of the oracle and of the GPU kernels: for trilinear elements on a box grid the element integrals
factorise, so every matrix entry and every residual entry of evaluate(All) has a closed form in the lattice
positions of its row and column nodes:

    A[p, q] = k1x m1y m1z + m1x k1y m1z + m1x m1y k1z
    1-D factors between lattice nodes at offset o (cells of length h; nL, nR in {0, 1}: cell present on that side):
        k1(0) = (nL + nR) / h,  k1(+-1) = -1 / h;      m1(0) = (nL + nR) h / 3,  m1(+-1) = h / 6
    f[p]    = sum_q A[p, q] x[q] - b[p],  b[p] = sum over the cells around p of the 2x2x2 Gauss quadrature of phi_p s,
              s = 12 pi^2 sin 2 pi x sin 2 pi y sin 2 pi z  (the example's source with multiplier -1)
    Dirichlet rows (the six faces, value 0): the reference applies them on each rank's ghosted container before the
    Export ADD, so a row shared by k ranks ends as k * identity and f = k * x.

(unit-cube interior check: 8h/3 on the diagonal, 0 to face neighbours, -h/6 to edge and -h/12 to corner neighbours).
bench.py uses it on sampled rows of the 256^3 workload and, for N > 1 GPUs, on a small brick per rank; the result is
printed as parity_max_rel_err / halo_parity_max_rel_err.
"""
import numpy as np

G = 0.57735026918962576451


def _k1(o, nl, nr, h):
    return np.where(o == 0, (nl + nr) / h, -1.0 / h)


def _m1(o, nl, nr, h):
    return np.where(o == 0, (nl + nr) * h / 3.0, h / 6.0)


def _load_1d(i, n, h, x0=0.0):
    """sum over the (up to two) cells around lattice coordinate i of  int phi_i(x) sin(2 pi x)  by 2-point Gauss"""
    i = np.asarray(i, np.float64)
    out = np.zeros_like(i)
    for side in (-1, 1):                      # cell to the left / right of the node
        present = (i > 0) if side < 0 else (i < n)
        xc = x0 + (i + 0.5 * side) * h        # cell centre
        for g in (-G, G):
            xq = xc + 0.5 * h * g
            N = 0.5 * (1.0 - side * g)        # the node's hat function at the Gauss point (node at the cell's -side end)
            out += np.where(present, 0.5 * h * N * np.sin(2.0 * np.pi * xq), 0.0)
    return out


def expected_rows(pos_row, pos_col, dims, x_col, dir_mult, lengths=(1.0, 1.0, 1.0)):
    """pos_row [R,3], pos_col [R,L,3] lattice positions (-1 rows in pos_col = absent slot), dims = (NX, NY, NZ) cells,
    x_col [R,L] solution at the columns, dir_mult [R] (0: not a Dirichlet row, k: shared by k ranks).
    Returns (A_expected [R,L], f_expected [R])."""
    n = np.asarray(dims, np.int64)
    h = np.asarray(lengths, np.float64) / n
    valid = pos_col[..., 0] >= 0
    o = np.where(valid[..., None], pos_col - pos_row[:, None, :], 0)
    inside = valid & (np.abs(o).max(axis=-1) <= 1)
    k1, m1 = [], []
    for d in range(3):
        nl = (pos_row[:, d] > 0).astype(np.float64)[:, None]
        nr = (pos_row[:, d] < n[d]).astype(np.float64)[:, None]
        k1.append(_k1(o[..., d], nl, nr, h[d]))
        m1.append(_m1(o[..., d], nl, nr, h[d]))
    A = k1[0] * m1[1] * m1[2] + m1[0] * k1[1] * m1[2] + m1[0] * m1[1] * k1[2]
    A = np.where(inside, A, 0.0)
    b = 12.0 * np.pi ** 2 * _load_1d(pos_row[:, 0], n[0], h[0]) * _load_1d(pos_row[:, 1], n[1], h[1]) * _load_1d(pos_row[:, 2], n[2], h[2])
    f = (A * np.where(valid, x_col, 0.0)).sum(axis=1) - b
    isd = dir_mult > 0
    diag = valid & (np.abs(o).max(axis=-1) == 0)
    A = np.where(isd[:, None], np.where(diag, dir_mult[:, None].astype(np.float64), 0.0), A)
    x_self = (np.where(diag, x_col, 0.0)).sum(axis=1)
    f = np.where(isd, dir_mult * x_self, f)
    return A, f


def check_rows(rows, rowptr, colind, A, f, pos_of_col, x_of_col, dims, dir_mult_of_row):
    """rows: LIDs to check; rowptr/colind/A/f: numpy views (only the sampled rows are touched); pos_of_col(cols) -> [.,3]
    lattice positions, x_of_col(cols) -> solution values, dir_mult_of_row(pos) -> multiplicity.
    Returns (max |A - A_exp|, max |A_exp|, max |f - f_exp|, max |f_exp|)."""
    rows = np.asarray(rows, np.int64)
    beg = rowptr[rows]
    ln = (rowptr[rows + 1] - beg).astype(np.int64)
    L = int(ln.max())
    idx = beg[:, None] + np.arange(L)[None, :]
    ok = np.arange(L)[None, :] < ln[:, None]
    idx = np.where(ok, idx, beg[:, None])
    cols = colind[idx]
    pos_col = np.where(ok[..., None], pos_of_col(cols.ravel()).reshape(len(rows), L, 3), -1)
    x_col = x_of_col(cols.ravel()).reshape(len(rows), L)
    pos_row = pos_of_col(rows)
    A_exp, f_exp = expected_rows(pos_row, pos_col, dims, x_col, dir_mult_of_row(pos_row))
    A_got = np.where(ok, A[idx], 0.0)
    return (float(np.abs(A_got - A_exp).max()), float(np.abs(A_exp).max()),
            float(np.abs(f[rows] - f_exp).max()), float(np.abs(f_exp).max()))
