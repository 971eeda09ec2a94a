// fill_rowgather.cu -- TXASM_SCATTER_ROWGATHER: owner-computes, general fallback.
//
// One thread owns one matrix row (one local DOF).  It walks the node->cell adjacency, recomputes
// the geometry of every cell around the node with the full 2x2x2 rule, forms only row a of each
// element matrix and adds it into its own CSR row with plain read-modify-writes: no atomics, no
// zero-fill pass, bitwise reproducible, any connectivity (irregular valence, rows of any length,
// columns missing from the graph are skipped like KokkosSparse sumIntoValues does).  It pays ~8x
// the geometry flops of the element-parallel form, so it is only the fallback for rows the
// row-tile kernel cannot take.
#include "txasm_internal.hpp"
#include "elem_q1hex.cuh"

namespace txasm {

template <bool JAC>
__global__ void __launch_bounds__(128) k_fill_rowgather(FillArgs A, const int64_t *__restrict__ adj_ptr,
                                                         const int *__restrict__ adj,
                                                         const int *__restrict__ row_list, int64_t n_list)
{
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= n_list) return;
  const int64_t row = row_list ? row_list[t] : t;
  const int64_t b0 = A.rowptr[row];
  const int len = (int)(A.rowptr[row + 1] - b0);
  if (JAC)
    for (int i = 0; i < len; ++i) A.A[b0 + i] = 0.0;
  double fr = 0.0;
  for (int64_t k = adj_ptr[row]; k < adj_ptr[row + 1]; ++k) {
    const int packed = adj[k];
    const int64_t e = packed >> 3;
    const int a = packed & 7;
    int lid[8];
    double X[8][3], ug[8], um[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      lid[n] = A.lids[e * 8 + n];
      const int64_t l = lid[n];
      X[n][0] = A.xyz[l * 3]; X[n][1] = A.xyz[l * 3 + 1]; X[n][2] = A.xyz[l * 3 + 2];
      double g = 0.0, m = 0.0;
#pragma unroll
      for (int v = 0; v < 3; ++v)
        if (A.c.has_vec[v]) {
          const double xv = A.x[v][l];
          g = fma(A.c.kg[v], xv, g);
          m = fma(A.c.km[v], xv, m);
        }
      ug[n] = g; um[n] = m;
    }
    const double sax = hex_sx(a), say = hex_sy(a), saz = hex_sz(a);
    double Krow[8];
#pragma unroll
    for (int b = 0; b < 8; ++b) Krow[b] = 0.0;
#pragma unroll 1
    for (int q = 0; q < 8; ++q) {
      const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      const double et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      const double ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
      double fx[2] = {1.0 - xi, 1.0 + xi}, fy[2] = {1.0 - et, 1.0 + et}, fz[2] = {1.0 - ze, 1.0 + ze};
      double N[8], dN[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const double ax = fx[hex_sx(n) > 0], ay = fy[hex_sy(n) > 0], az = fz[hex_sz(n) > 0];
        N[n] = 0.125 * ax * ay * az;
        dN[n][0] = 0.125 * hex_sx(n) * ay * az;
        dN[n][1] = 0.125 * ax * hex_sy(n) * az;
        dN[n][2] = 0.125 * ax * ay * hex_sz(n);
      }
      double J[3][3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          double s = 0.0;
#pragma unroll
          for (int n = 0; n < 8; ++n) s = fma(X[n][d], dN[n][c], s);
          J[d][c] = s;
        }
      const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
      const double c1 = J[2][0] * J[1][2] - J[1][0] * J[2][2];
      const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
      const double idet = 1.0 / det;
      double Ji[3][3];
      Ji[0][0] = c0 * idet; Ji[1][0] = c1 * idet; Ji[2][0] = c2 * idet;
      Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet;
      Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
      Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet;
      Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
      Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
      Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
      // the row's own basis function at this point
      const double aax = 1.0 + sax * xi, aay = 1.0 + say * et, aaz = 1.0 + saz * ze;
      const double Na = 0.125 * aax * aay * aaz;
      const double dNa[3] = {0.125 * sax * aay * aaz, 0.125 * aax * say * aaz, 0.125 * aax * aay * saz};
      double Ga[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) Ga[d] = Ji[0][d] * dNa[0] + Ji[1][d] * dNa[1] + Ji[2][d] * dNa[2];
      const double wK = A.c.fmK ? det * A.c.fmK[e * 8 + q] : det, wM = A.c.fmM ? det * A.c.fmM[e * 8 + q] : det;
      double gu[3] = {0.0, 0.0, 0.0};
      double sq = 0.0, xq = 0.0, yq = 0.0, zq = 0.0;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        double Gn[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          Gn[d] = Ji[0][d] * dN[n][0] + Ji[1][d] * dN[n][1] + Ji[2][d] * dN[n][2];
          gu[d] = fma(ug[n], Gn[d], gu[d]);
        }
        if (JAC) {
          const double gg = Ga[0] * Gn[0] + Ga[1] * Gn[1] + Ga[2] * Gn[2];
          Krow[n] = fma(wK * A.c.cK, gg, Krow[n]);
          if (A.c.has_mass) Krow[n] = fma(wM * A.c.cM * Na, N[n], Krow[n]);
        }
        sq = fma(N[n], um[n], sq);
        xq = fma(N[n], X[n][0], xq); yq = fma(N[n], X[n][1], yq); zq = fma(N[n], X[n][2], zq);
      }
      sq = A.c.has_mass ? sq * wM : 0.0;
      for (int s = 0; s < A.c.n_src; ++s) {
        const double v = (A.c.src_id[s] == TXASM_SOURCE_IP_ARRAY) ? A.c.src_ip[s][e * 8 + q] : source_eval(A.c.src_id[s], xq, yq, zq);
        sq = fma(A.c.src_mult[s] * det, v, sq);
      }
      fr = fma(wK, Ga[0] * gu[0] + Ga[1] * gu[1] + Ga[2] * gu[2], fr);
      fr = fma(sq, Na, fr);
    }
    if (JAC) {
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const int col = lid[b];
        int lo = 0, hi = len - 1, at = -1;
        while (lo <= hi) {
          const int mid = (lo + hi) >> 1;
          const int c = A.colind[b0 + mid];
          if (c == col) { at = mid; break; }
          if (c < col) lo = mid + 1; else hi = mid - 1;
        }
        if (at >= 0) A.A[b0 + at] += Krow[b];
      }
    }
  }
  if (A.f) A.f[row] = fr;
}

int launch_fill_rowgather(txasm_handle h, const FillArgs &a)
{
  const unsigned grid = (unsigned)((a.n_rows + 127) / 128);
  if (a.jacobian) k_fill_rowgather<true><<<grid, 128, 0, h->stream>>>(a, h->d_adj_ptr, h->d_adj, nullptr, a.n_rows);
  else k_fill_rowgather<false><<<grid, 128, 0, h->stream>>>(a, h->d_adj_ptr, h->d_adj, nullptr, a.n_rows);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

// rows listed in row_list only (the irregular rows the row-tile kernel leaves out)
int launch_fill_rowgather_list(txasm_handle h, const FillArgs &a, const int *row_list, int64_t n)
{
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (a.jacobian) k_fill_rowgather<true><<<grid, 128, 0, h->stream>>>(a, h->d_adj_ptr, h->d_adj, row_list, n);
  else k_fill_rowgather<false><<<grid, 128, 0, h->stream>>>(a, h->d_adj_ptr, h->d_adj, row_list, n);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return TXASM_OK;
}

}  // namespace txasm
