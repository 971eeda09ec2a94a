// fill_general.cu -- element matrices of general (non-affine) trilinear hexahedra, one thread per cell.
//
// The full 2x2x2 Gauss rule (IntegrationValues2 / BasisValues2 / DOFGradient / Integrator_* of the reference, fused in
// elem_general(), elem_q1hex.cuh) costs ~5 kflop per cell and is bound by the FP64 pipe, not by HBM.  The row-tile
// kernel used to run it inside phase 1 for every tile cell, halo cells included: 1.8x the work at 128-row tiles, 255
// registers and 8 warps per SM for the whole kernel (18.2 ms at 256^3, round 1).  Here every cell is computed exactly
// once into a 352-byte record (K upper triangle [36] | r[8]); k_fill_rowtile<.., AFFINE = false, ..> then only gathers the matrix
// rows of its DOFs from the records of their 8 cells, so the second pass has no halo and no staging.
#include "txasm_internal.hpp"
#include "tiles.hpp"

namespace txasm {

// Reference basis at the 2x2x2 Gauss points (Basis_HGRAD_HEX_C1 on Shards Hexahedron<8>, tensor Gauss-Legendre rule:
// SURVEY.md appendix C): filled once by the host; the kernel reads them as constant-bank operands instead of rebuilding
// them from (1 +- xi)(1 +- eta)(1 +- zeta) for every cell and point (72 DMUL per point).
__constant__ double c_N[8][8];          // [q][n]
__constant__ double c_dN[8][8][3];      // [q][n][e]

// One cell, every Gauss point: same integrals as elem_general() (elem_q1hex.cuh), arranged for the FP64 pipe:
//  * the Jacobian comes from the 8 coefficient vectors of the trilinear map x(xi) = sum_m C_m xi^m1 eta^m2 zeta^m3
//    (a butterfly over the vertices, once per cell), 9 FMA per point instead of 72;
//  * N, dN are constants; the closure model uses the polynomial sin2pi_fast.
template <bool JAC>
__device__ __forceinline__ void elem_general_fast(const double (&X)[8][3], const double (&ug)[8], const double (&um)[8],
                                                  const FillCoef &c, int64_t cell, double (&K)[36], double (&r)[8])
{
#pragma unroll
  for (int a = 0; a < 8; ++a) r[a] = 0.0;
  if (JAC) {
#pragma unroll
    for (int i = 0; i < 36; ++i) K[i] = 0.0;
  }
  // C[m][d], m = bx + 2 by + 4 bz: coefficient of xi^bx eta^by zeta^bz (times 8)
  double C[8][3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double P[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) P[m] = X[hex_vertex(m & 1, (m >> 1) & 1, m >> 2)][d];
#pragma unroll
    for (int st = 1; st < 8; st <<= 1)
#pragma unroll
      for (int m = 0; m < 8; ++m)
        if (!(m & st)) { const double lo = P[m], hi = P[m | st]; P[m] = hi + lo; P[m | st] = hi - lo; }
#pragma unroll
    for (int m = 0; m < 8; ++m) C[m][d] = 0.125 * P[m];
  }
  const bool do_grad = (c.kg[0] != 0.0 || c.kg[1] != 0.0 || c.kg[2] != 0.0) || (JAC && c.cK != 0.0);
#pragma unroll 1
  for (int q = 0; q < 8; ++q) {
    const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
    const double et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
    const double ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
    const double ez = et * ze, xz = xi * ze, xe = xi * et;
    double J[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      J[d][0] = fma(C[7][d], ez, fma(C[5][d], ze, fma(C[3][d], et, C[1][d])));
      J[d][1] = fma(C[7][d], xz, fma(C[6][d], ze, fma(C[3][d], xi, C[2][d])));
      J[d][2] = fma(C[7][d], xe, fma(C[6][d], et, fma(C[5][d], xi, C[4][d])));
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
    const double w0 = det;  // weighted_measure = detJ * w_q, w_q = 1
    const double w = c.fmK ? w0 * c.fmK[cell * 8 + q] : w0;          // ... times the GRADGRAD field multipliers
    const double wmass = c.fmM ? w0 * c.fmM[cell * 8 + q] : w0;
    if (do_grad) {
      double G[8][3];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const double d0 = c_dN[q][n][0], d1 = c_dN[q][n][1], d2 = c_dN[q][n][2];
#pragma unroll
        for (int d = 0; d < 3; ++d) G[n][d] = fma(Ji[2][d], d2, fma(Ji[1][d], d1, Ji[0][d] * d0));
      }
      double gu[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int d = 0; d < 3; ++d) gu[d] = fma(ug[n], G[n][d], gu[d]);
#pragma unroll
      for (int d = 0; d < 3; ++d) gu[d] *= w;
#pragma unroll
      for (int a = 0; a < 8; ++a) r[a] = fma(G[a][2], gu[2], fma(G[a][1], gu[1], fma(G[a][0], gu[0], r[a])));
      if (JAC) {
        const double wk = w * c.cK;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const double h0 = wk * G[a][0], h1 = wk * G[a][1], h2 = wk * G[a][2];
#pragma unroll
          for (int b = a; b < 8; ++b)
            K[sym_idx(a, b)] = fma(h2, G[b][2], fma(h1, G[b][1], fma(h0, G[b][0], K[sym_idx(a, b)])));
        }
      }
    }
    double sq = 0.0;
    if (c.has_mass) {
#pragma unroll
      for (int n = 0; n < 8; ++n) sq = fma(c_N[q][n], um[n], sq);
      sq *= wmass;
      if (JAC) {
        const double wm = wmass * c.cM;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const double na = wm * c_N[q][a];
#pragma unroll
          for (int b = a; b < 8; ++b) K[sym_idx(a, b)] = fma(na, c_N[q][b], K[sym_idx(a, b)]);
        }
      }
    }
    if (c.n_src > 0) {
      double pq[3];
#pragma unroll
      for (int d = 0; d < 3; ++d)
        pq[d] = fma(C[7][d], xe * ze, fma(C[6][d], ez, fma(C[5][d], xz, fma(C[3][d], xe, fma(C[4][d], ze, fma(C[2][d], et, fma(C[1][d], xi, C[0][d])))))));
      for (int s = 0; s < c.n_src; ++s) {
        double v;
        if (c.src_id[s] == TXASM_SOURCE_IP_ARRAY) v = c.src_ip[s][cell * 8 + q];
        else if (c.src_id[s] == TXASM_SOURCE_SIN3) v = 118.43525281307230 * sin2pi_fast(pq[0]) * sin2pi_fast(pq[1]) * sin2pi_fast(pq[2]);
        else v = source_eval(c.src_id[s], pq[0], pq[1], pq[2]);
        sq = fma(c.src_mult[s] * w0, v, sq);
      }
    }
    if (c.has_mass || c.n_src > 0) {
#pragma unroll
      for (int a = 0; a < 8; ++a) r[a] = fma(sq, c_N[q][a], r[a]);
    }
  }
}

constexpr int EG_THREADS = 128;

template <bool JAC, int MINB>
__global__ void __launch_bounds__(EG_THREADS, MINB) k_elem_general(FillArgs A, double *__restrict__ elem)
{
  const int64_t e = (int64_t)blockIdx.x * EG_THREADS + threadIdx.x;
  if (e >= A.n_cells) return;
  double X[8][3], ug[8], um[8];
  int lid[8];
  {
    const int4 *p = reinterpret_cast<const int4 *>(A.lids + e * 8);
    const int4 v0 = __ldg(p), v1 = __ldg(p + 1);
    lid[0] = v0.x; lid[1] = v0.y; lid[2] = v0.z; lid[3] = v0.w; lid[4] = v1.x; lid[5] = v1.y; lid[6] = v1.z; lid[7] = v1.w;
  }
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    const int64_t l = lid[n];
    X[n][0] = __ldg(A.xyz + l * 3); X[n][1] = __ldg(A.xyz + l * 3 + 1); X[n][2] = __ldg(A.xyz + l * 3 + 2);
    double g = 0.0, m = 0.0;
#pragma unroll
    for (int v = 0; v < 3; ++v)
      if (A.c.has_vec[v]) {
        const double xv = __ldg(A.x[v] + l);
        g = fma(A.c.kg[v], xv, g);
        m = fma(A.c.km[v], xv, m);
      }
    ug[n] = g; um[n] = m;
  }
  double K[36], r[8];
  elem_general_fast<JAC>(X, ug, um, A.c, e, K, r);
  // the record: the upper triangle as accumulated (352 bytes instead of 576: the element matrix is symmetric, and the row
  // pass reads the 8 rows of a cell from the same CTA, so the mirrored half only cost DRAM traffic in both passes)
  double2 *o = reinterpret_cast<double2 *>(elem + e * ELEM_REC);
  if (JAC) {
#pragma unroll
    for (int i = 0; i < 36; i += 2) o[i / 2] = make_double2(K[i], K[i + 1]);
  }
#pragma unroll
  for (int k = 0; k < 8; k += 2) o[18 + k / 2] = make_double2(r[k], r[k + 1]);
}

static int upload_reference_basis(txasm_handle h)
{
  static bool done[64] = {};
  if (h->device >= 0 && h->device < 64 && done[h->device]) return TXASM_OK;
  double N[8][8], dN[8][8][3];
  for (int q = 0; q < 8; ++q) {
    const double xi = (q & 1) ? TX_INV_SQRT3 : -TX_INV_SQRT3, et = (q & 2) ? TX_INV_SQRT3 : -TX_INV_SQRT3, ze = (q & 4) ? TX_INV_SQRT3 : -TX_INV_SQRT3;
    for (int n = 0; n < 8; ++n) {
      const double ax = 1.0 + hex_sx(n) * xi, ay = 1.0 + hex_sy(n) * et, az = 1.0 + hex_sz(n) * ze;
      N[q][n] = 0.125 * ax * ay * az;
      dN[q][n][0] = 0.125 * hex_sx(n) * ay * az;
      dN[q][n][1] = 0.125 * ax * hex_sy(n) * az;
      dN[q][n][2] = 0.125 * ax * ay * hex_sz(n);
    }
  }
  TX_CUDA(h, cudaMemcpyToSymbolAsync(c_N, N, sizeof(N), 0, cudaMemcpyHostToDevice, h->stream));
  TX_CUDA(h, cudaMemcpyToSymbolAsync(c_dN, dN, sizeof(dN), 0, cudaMemcpyHostToDevice, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  if (h->device >= 0 && h->device < 64) done[h->device] = true;
  return TXASM_OK;
}

int launch_elem_general(txasm_handle h, FillArgs &a, cudaStream_t st)
{
  { int rc = upload_reference_basis(h); if (rc) return rc; }
  if (!h->d_elem) {
    int rc = dev_alloc(h, &h->d_elem, (size_t)h->n_cells * ELEM_REC);
    if (rc) return rc;
  }
  const unsigned grid = (unsigned)((h->n_cells + EG_THREADS - 1) / EG_THREADS);
  // 2 CTAs per SM (244 registers).  TXASM_ELEM_MINB=3 selects the 168-register build (3 CTAs per SM, a few loop invariants
  // spilled): same time in the bench, lower FP64-pipe utilisation under ncu
  static const int minb = [] { const char *e = getenv("TXASM_ELEM_MINB"); return e ? atoi(e) : 2; }();
  if (a.jacobian) {
    if (minb == 2) k_elem_general<true, 2><<<grid, EG_THREADS, 0, st>>>(a, h->d_elem);
    else k_elem_general<true, 3><<<grid, EG_THREADS, 0, st>>>(a, h->d_elem);
  } else k_elem_general<false, 2><<<grid, EG_THREADS, 0, st>>>(a, h->d_elem);
  TX_CUDA(h, cudaGetLastError());
  h->launches += 1;
  a.elem = h->d_elem;
  return TXASM_OK;
}

}  // namespace txasm
