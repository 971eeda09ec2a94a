// gblock.cu -- general element blocks: every (topology, basis, field layout) the scalar Q1-hexahedron fast path does not take.
//
// BASELINE.json configs 3-5: Q2 hexahedra and P1 / P2 tetrahedra (scalar diffusion, several blocks in one handle), three
// interleaved HGRAD fields on Q1 hexahedra (linear elastodynamics, second order in time), HCURL edge elements on
// hexahedra (curl-curl + mass, with edge orientations).  One kernel per block replaces the reference's workset loop
//
//   GatherSolution_Tpetra -> DOF / DOFGradient / DOFCurl -> closure models -> Integrator_GradBasisDotVector /
//   _BasisTimesScalar / _BasisTimesVector / _CurlBasisDotVector -> ScatterResidual_Tpetra
//   (disc-fe/src/evaluators/Panzer_*_impl.hpp; BasisValues2 transforms disc-fe/src/Panzer_BasisValues2_impl.hpp:1036-1190,
//    1264-1275, 1376-1521, 1727-1735; IntegrationValues2 disc-fe/src/Panzer_IntegrationValues2.cpp:946-1221; interleaved
//    field layout dof-mgr/src/Panzer_FieldAggPattern.cpp:201-276; orientations disc-fe/src/Panzer_IntrepidOrientation.cpp:96-99)
//
// with: one thread per (cell, element DOF row).  Per integration point the threads of a cell build the Jacobian and
// the physical basis quantities of their own basis function once into shared memory; then every thread accumulates ITS
// row of the element matrix in registers (NDOF accumulators) together with its residual entry -- the forward-mode
// derivative of these linear integrands is the element matrix times the gather seed, so it is formed directly.  The
// rows leave like ScatterResidual's sumIntoValues does (red.global.add.f64), but the searched column positions come from
// a plan built at setup (one byte per element-matrix entry; 0xFF = column absent from the row, skipped as KokkosSparse does).
//
// The Q2 hexahedron has a second kernel, k_gblock_q2_dmma: its element matrix  K = sum_q w_q G_q G_q^T  (G_q: 27 x 3)
// is a dense 27 x 81 by 81 x 27 contraction and goes to the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA), one warp per cell.
#include "txasm_internal.hpp"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace txasm {

constexpr int GB_THREADS = 128;
constexpr int GB_MAXQ = 64;

enum { GE_HEX8_C1 = 1, GE_HEX27_C2 = 2, GE_TET4_C1 = 3, GE_TET10_C2 = 4, GE_HEX8_HCURL = 5 };

struct GBlock {
  int elem = 0, nb = 0, nv = 0, nfld = 1, ndof = 0, nq = 0, deg = 0;
  int64_t n_cells = 0;
  const double *d_coords = nullptr;
  const int *d_lids = nullptr;
  const signed char *d_signs = nullptr;
  int *d_dofmap = nullptr;          // [ndof] packed basis | field << 8 of the element DOF at position j
  double *d_tab = nullptr;          // weights[nq] | geo grads [nq][nv][3] | values [nq][nb][vdim] | derivatives [nq][nb][3]
  unsigned char *d_plan = nullptr;  // [n_cells][ndof][ndof] CSR offset of column lid[k] in row lid[j] (0xFF absent)
  double *d_scratch = nullptr;      // owner-computes mode: element rows [n_cells][ndof][ndof + 1], allocated at the first evaluate
  int op = 0;
  double p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int use_dmma = 1;
};

struct GBlocks {
  std::vector<GBlock> b;
  int64_t *d_adj_ptr = nullptr, *d_adj = nullptr;      // row -> (block << 56 | local dof << 48 | cell), sorted
};

// ------------------------------------------------------------------ reference elements (host): tables at the cubature points
static const double kHex27[27][3] = {
  {-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1},
  {0,-1,-1},{1,0,-1},{0,1,-1},{-1,0,-1},{-1,-1,0},{1,-1,0},{1,1,0},{-1,1,0},{0,-1,1},{1,0,1},{0,1,1},{-1,0,1},
  {0,0,0},{0,0,-1},{0,0,1},{-1,0,0},{1,0,0},{0,-1,0},{0,1,0}};
static const int kHexEdge[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};
static const int kTetEdge[6][2] = {{0,1},{1,2},{0,2},{0,3},{1,3},{2,3}};

static int elem_nb(int e) { return e == GE_HEX8_C1 ? 8 : e == GE_HEX27_C2 ? 27 : e == GE_TET4_C1 ? 4 : e == GE_TET10_C2 ? 10 : e == GE_HEX8_HCURL ? 12 : -1; }
static int elem_nv(int e) { return (e == GE_TET4_C1 || e == GE_TET10_C2) ? 4 : 8; }

// quadratic Lagrange polynomial on {-1, 0, 1} attached to node coordinate c, and its derivative
static double q2(double c, double t) { return c < -0.5 ? 0.5 * t * (t - 1) : (c > 0.5 ? 0.5 * t * (t + 1) : (1 - t) * (1 + t)); }
static double dq2(double c, double t) { return c < -0.5 ? t - 0.5 : (c > 0.5 ? t + 0.5 : -2 * t); }

static void ref_point(int elem, const double pt[3], double *val, double *der)
{
  const double x = pt[0], y = pt[1], z = pt[2];
  switch (elem) {
    case GE_HEX8_C1:
      for (int n = 0; n < 8; ++n) {
        const double sx = kHex27[n][0], sy = kHex27[n][1], sz = kHex27[n][2];
        val[n] = (1 + sx * x) * (1 + sy * y) * (1 + sz * z) / 8;
        der[3 * n] = sx * (1 + sy * y) * (1 + sz * z) / 8; der[3 * n + 1] = (1 + sx * x) * sy * (1 + sz * z) / 8; der[3 * n + 2] = (1 + sx * x) * (1 + sy * y) * sz / 8;
      }
      break;
    case GE_HEX27_C2:
      for (int n = 0; n < 27; ++n) {
        const double *c = kHex27[n];
        val[n] = q2(c[0], x) * q2(c[1], y) * q2(c[2], z);
        der[3 * n] = dq2(c[0], x) * q2(c[1], y) * q2(c[2], z); der[3 * n + 1] = q2(c[0], x) * dq2(c[1], y) * q2(c[2], z); der[3 * n + 2] = q2(c[0], x) * q2(c[1], y) * dq2(c[2], z);
      }
      break;
    case GE_TET4_C1: case GE_TET10_C2: {
      const double L[4] = {1 - x - y - z, x, y, z}, dL[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      for (int n = 0; n < 4; ++n)
        for (int d = 0; d < 3; ++d) {
          if (elem == GE_TET4_C1) { val[n] = L[n]; der[3 * n + d] = dL[n][d]; }
          else { val[n] = L[n] * (2 * L[n] - 1); der[3 * n + d] = (4 * L[n] - 1) * dL[n][d]; }
        }
      if (elem == GE_TET10_C2)
        for (int e = 0; e < 6; ++e) {
          const int i = kTetEdge[e][0], j = kTetEdge[e][1];
          val[4 + e] = 4 * L[i] * L[j];
          for (int d = 0; d < 3; ++d) der[3 * (4 + e) + d] = 4 * (dL[i][d] * L[j] + L[i] * dL[j][d]);
        }
      break;
    }
    case GE_HEX8_HCURL:
      for (int e = 0; e < 12; ++e) {
        const double *a = kHex27[kHexEdge[e][0]], *b = kHex27[kHexEdge[e][1]];
        int dir = 0;
        for (int d = 0; d < 3; ++d) if (a[d] != b[d]) dir = d;
        const int d1 = (dir + 1) % 3, d2 = (dir + 2) % 3;
        const double s = (b[dir] - a[dir]) / 2, f1 = 1 + a[d1] * pt[d1], f2 = 1 + a[d2] * pt[d2];
        for (int d = 0; d < 3; ++d) { val[3 * e + d] = 0; der[3 * e + d] = 0; }
        val[3 * e + dir] = s * f1 * f2 / 4;          // unit tangential trace on the own edge
        der[3 * e + d1] = s * f1 * a[d2] / 4;        // curl (phi e_dir) = grad phi x e_dir
        der[3 * e + d2] = -s * a[d1] * f2 / 4;
      }
      break;
  }
}

static int gauss_rule(int n, double *x, double *w)
{
  if (n < 1 || n > 4) return -1;
  static const double X[4][4] = {{0}, {-0.57735026918962576451, 0.57735026918962576451}, {-0.77459666924148337704, 0, 0.77459666924148337704},
                                 {-0.86113631159405257522, -0.33998104358485626480, 0.33998104358485626480, 0.86113631159405257522}};
  static const double W[4][4] = {{2}, {1, 1}, {5.0 / 9, 8.0 / 9, 5.0 / 9}, {0.34785484513745385737, 0.65214515486254614263, 0.65214515486254614263, 0.34785484513745385737}};
  for (int i = 0; i < n; ++i) { x[i] = X[n - 1][i]; w[i] = W[n - 1][i]; }
  return 0;
}

static int cubature(int elem, int deg, double (*pts)[3], double *w)
{
  if (elem == GE_TET4_C1 || elem == GE_TET10_C2) {
    if (deg <= 1) { pts[0][0] = pts[0][1] = pts[0][2] = 0.25; w[0] = 1.0 / 6; return 1; }
    if (deg == 2) {
      const double a = 0.58541019662496845446, b = 0.13819660112501051518;
      for (int q = 0; q < 4; ++q) { for (int d = 0; d < 3; ++d) pts[q][d] = (q == d) ? a : b; w[q] = 1.0 / 24; }
      return 4;
    }
    if (deg == 3) {
      pts[0][0] = pts[0][1] = pts[0][2] = 0.25; w[0] = -2.0 / 15;
      for (int q = 1; q < 5; ++q) { for (int d = 0; d < 3; ++d) pts[q][d] = (q - 1 == d) ? 0.5 : 1.0 / 6; w[q] = 3.0 / 40; }
      return 5;
    }
    return -1;
  }
  const int n = deg / 2 + 1;
  double gx[4], gw[4];
  if (gauss_rule(n, gx, gw)) return -1;
  int q = 0;
  for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i, ++q) {
    pts[q][0] = gx[i]; pts[q][1] = gx[j]; pts[q][2] = gx[k]; w[q] = gw[i] * gw[j] * gw[k];
  }
  return q;
}

// ------------------------------------------------------------------ kernels
struct GArgs {
  int64_t n_cells;
  const double *coords;
  const int *lids;
  const signed char *signs;
  const int *dofmap;
  const double *tab;
  const unsigned char *plan;
  const int64_t *rowptr;
  const double *x[3];
  double *f, *A;
  double *scratch;          // owner-computes mode: element rows [n_cells][ND][ND + 1] (row of K | residual entry), else NULL
  int nq, jacobian;
  double cK, cM;            // Jacobian = cK * K0 + cM * M0
  double kx, mx[3];         // residual = K0 (kx x) + M0 (mx[0] x + mx[1] xdot + mx[2] xdotdot) + source
  double lam, mu;           // elasticity
  double src[3];
};

enum { GOP_DIFFUSION = 1, GOP_ELASTICITY = 2, GOP_CURLCURL = 3 };

template <int OP, int NB, int NV, int NFLD>
__global__ void __launch_bounds__(GB_THREADS) k_gblock(GArgs G)
{
  constexpr int ND = NB * NFLD, CPC = GB_THREADS / ND, VD = (OP == GOP_CURLCURL) ? 3 : 1;
  __shared__ double sX[CPC][NV * 3];
  __shared__ double sJ[CPC][9];
  __shared__ double sP[CPC][NB][3];          // physical gradients (HGRAD) or curls (HCURL) at the current point
  __shared__ double sV[CPC][NB][VD];         // physical values
  __shared__ double sU[CPC][ND], sM[CPC][ND];
  __shared__ int sL[CPC][ND];
  const int tid = threadIdx.x, cl = tid / ND, j = tid - cl * ND;
  const int64_t cell = (int64_t)blockIdx.x * CPC + cl;
  const bool on = cl < CPC && cell < G.n_cells;
  int mybas = 0, myfld = 0;
  if (on) {
    const int dm = G.dofmap[j];
    mybas = dm & 0xFF; myfld = dm >> 8;
    const int lid = G.lids[cell * ND + j];
    sL[cl][j] = lid;
    double u = 0.0, m = 0.0;
    if (G.x[0]) { const double v = G.x[0][lid]; u = G.kx * v; m = G.mx[0] * v; }
    if (G.x[1]) m = fma(G.mx[1], G.x[1][lid], m);
    if (G.x[2]) m = fma(G.mx[2], G.x[2][lid], m);
    sU[cl][j] = u; sM[cl][j] = m;
    for (int i = j; i < NV * 3; i += ND) sX[cl][i] = G.coords[cell * NV * 3 + i];
  }
  const double *wts = G.tab, *geo = wts + G.nq, *val = geo + (size_t)G.nq * NV * 3, *der = val + (size_t)G.nq * NB * VD;
  double row[ND];
#pragma unroll
  for (int k = 0; k < ND; ++k) row[k] = 0.0;
  double r = 0.0;
  __syncthreads();
  for (int q = 0; q < G.nq; ++q) {
    if (on)                                  // IntegrationValues2: J[d][e] = sum_n x_n[d] dN_n/dxi_e
      for (int i = j; i < 9; i += ND) {
        const int d = i / 3, e = i - 3 * d;
        double a = 0.0;
#pragma unroll
        for (int n = 0; n < NV; ++n) a = fma(sX[cl][n * 3 + d], __ldg(geo + (q * NV + n) * 3 + e), a);
        sJ[cl][i] = a;
      }
    __syncthreads();
    double w = 0.0;
    if (on) {
      const double J00 = sJ[cl][0], J01 = sJ[cl][1], J02 = sJ[cl][2], J10 = sJ[cl][3], J11 = sJ[cl][4], J12 = sJ[cl][5],
                   J20 = sJ[cl][6], J21 = sJ[cl][7], J22 = sJ[cl][8];
      const double c0 = J11 * J22 - J21 * J12, c1 = J20 * J12 - J10 * J22, c2 = J10 * J21 - J20 * J11;
      const double det = J00 * c0 + J01 * c1 + J02 * c2, id = 1.0 / det;
      w = det * __ldg(wts + q);
      if (myfld == 0) {                      // BasisValues2: the physical quantities of basis function `mybas`
        const double *dr = der + (q * NB + mybas) * 3;
        const double r0 = __ldg(dr), r1 = __ldg(dr + 1), r2 = __ldg(dr + 2);
        if (OP != GOP_CURLCURL) {
          // HGRADtransformGRAD: J^-T grad_ref
          sP[cl][mybas][0] = (c0 * r0 + c1 * r1 + c2 * r2) * id;
          sP[cl][mybas][1] = ((J02 * J21 - J01 * J22) * r0 + (J00 * J22 - J02 * J20) * r1 + (J01 * J20 - J00 * J21) * r2) * id;
          sP[cl][mybas][2] = ((J01 * J12 - J02 * J11) * r0 + (J02 * J10 - J00 * J12) * r1 + (J00 * J11 - J01 * J10) * r2) * id;
          sV[cl][mybas][0] = __ldg(val + q * NB + mybas);
        } else {
          const double s = G.signs ? (double)G.signs[cell * NB + mybas] : 1.0;      // applyOrientations
          // HCURLtransformCURL: J curl_ref / det;  HCURLtransformVALUE: J^-T phi_ref
          sP[cl][mybas][0] = s * (J00 * r0 + J01 * r1 + J02 * r2) * id;
          sP[cl][mybas][1] = s * (J10 * r0 + J11 * r1 + J12 * r2) * id;
          sP[cl][mybas][2] = s * (J20 * r0 + J21 * r1 + J22 * r2) * id;
          const double *vr = val + (q * NB + mybas) * 3;
          const double v0 = __ldg(vr), v1 = __ldg(vr + 1), v2 = __ldg(vr + 2);
          sV[cl][mybas][0] = s * (c0 * v0 + c1 * v1 + c2 * v2) * id;
          sV[cl][mybas][VD > 1 ? 1 : 0] = s * ((J02 * J21 - J01 * J22) * v0 + (J00 * J22 - J02 * J20) * v1 + (J01 * J20 - J00 * J21) * v2) * id;
          sV[cl][mybas][VD > 2 ? 2 : 0] = s * ((J01 * J12 - J02 * J11) * v0 + (J02 * J10 - J00 * J12) * v1 + (J00 * J11 - J01 * J10) * v2) * id;
        }
      }
    }
    __syncthreads();
    if (on) {
      const double ga0 = sP[cl][mybas][0], ga1 = sP[cl][mybas][1], ga2 = sP[cl][mybas][2];
      const double va0 = sV[cl][mybas][0], va1 = sV[cl][mybas][VD > 1 ? 1 : 0], va2 = sV[cl][mybas][VD > 2 ? 2 : 0];
      if (OP == GOP_DIFFUSION) {
#pragma unroll
        for (int k = 0; k < ND; ++k) {       // (one field: element DOF k is basis function dofmap[k]; identity in practice)
          const int b = G.dofmap[k] & 0xFF;
          const double k0 = w * (ga0 * sP[cl][b][0] + ga1 * sP[cl][b][1] + ga2 * sP[cl][b][2]);
          const double m0 = w * va0 * sV[cl][b][0];
          row[k] = fma(G.cK, k0, fma(G.cM, m0, row[k]));
          r = fma(k0, sU[cl][k], fma(m0, sM[cl][k], r));
        }
        r = fma(w * va0, G.src[0], r);
      } else if (OP == GOP_ELASTICITY) {
        const double gi = myfld == 0 ? ga0 : (myfld == 1 ? ga1 : ga2);
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const int dm = G.dofmap[k], b = dm & 0xFF, kk = dm >> 8;
          const double g0 = sP[cl][b][0], g1 = sP[cl][b][1], g2 = sP[cl][b][2];
          const double gk = kk == 0 ? g0 : (kk == 1 ? g1 : g2);            // d phi_b / d x_kk
          const double gak = kk == 0 ? ga0 : (kk == 1 ? ga1 : ga2);         // d phi_a / d x_kk
          const double gbi = myfld == 0 ? g0 : (myfld == 1 ? g1 : g2);      // d phi_b / d x_i
          double k0 = G.lam * gi * gk + G.mu * gak * gbi;
          double m0 = 0.0;
          if (kk == myfld) { k0 = fma(G.mu, ga0 * g0 + ga1 * g1 + ga2 * g2, k0); m0 = w * va0 * sV[cl][b][0]; }
          k0 *= w;
          row[k] = fma(G.cK, k0, fma(G.cM, m0, row[k]));
          r = fma(k0, sU[cl][k], fma(m0, sM[cl][k], r));
        }
        r = fma(w * va0, G.src[myfld], r);
      } else {
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const int b = G.dofmap[k] & 0xFF;
          const double k0 = w * (ga0 * sP[cl][b][0] + ga1 * sP[cl][b][1] + ga2 * sP[cl][b][2]);
          const double m0 = w * (va0 * sV[cl][b][0] + va1 * sV[cl][b][VD > 1 ? 1 : 0] + va2 * sV[cl][b][VD > 2 ? 2 : 0]);
          row[k] = fma(G.cK, k0, fma(G.cM, m0, row[k]));
          r = fma(k0, sU[cl][k], fma(m0, sM[cl][k], r));
        }
        r += w * (va0 * G.src[0] + va1 * G.src[1] + va2 * G.src[2]);
      }
    }
    __syncthreads();
  }
  if (!on) return;
  if (G.scratch) {                           // owner-computes mode: k_gblock_rows gathers the element rows
    double *o = G.scratch + (cell * ND + j) * (ND + 1);
    if (G.jacobian) {
#pragma unroll
      for (int k = 0; k < ND; ++k) o[k] = row[k];
    }
    o[ND] = r;
    return;
  }
  // ScatterResidual_Tpetra: atomic add of the residual entry; sumIntoValues of the row with planned positions
  const int lid = sL[cl][j];
  if (G.f) atomicAdd(G.f + lid, r);
  if (G.jacobian && G.A) {
    const int64_t base = G.rowptr[lid];
    const unsigned char *pl = G.plan + (cell * ND + j) * ND;
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const unsigned p = pl[k];
      if (p != 0xFFu) atomicAdd(G.A + base + p, row[k]);
    }
  }
}

// CSR offset of every element-matrix entry: row lid[j], column lid[k]
__global__ void k_gblock_plan(int64_t n_cells, int nd, const int *__restrict__ lids, const int64_t *__restrict__ rowptr,
                              const int *__restrict__ colind, unsigned char *__restrict__ plan, int *__restrict__ too_long)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;          // (cell, j)
  if (i >= n_cells * nd) return;
  const int64_t cell = i / nd;
  const int row = lids[i];
  const int64_t b = rowptr[row];
  const int len = (int)(rowptr[row + 1] - b);
  if (len > 255) { atomicExch(too_long, 1); return; }
  for (int k = 0; k < nd; ++k) {
    const int col = lids[cell * nd + k];
    int lo = 0, hi = len - 1, at = 0xFF;
    while (lo <= hi) {
      const int mid = (lo + hi) >> 1, v = colind[b + mid];
      if (v == col) { at = mid; break; }
      if (v < col) lo = mid + 1; else hi = mid - 1;
    }
    plan[i * nd + k] = (unsigned char)at;
  }
}

// ------------------------------------------------------------------ Q2 hexahedron on the FP64 tensor cores
// One warp per cell.  Per integration point the lanes 0..26 hold the physical gradient of "their" basis function scaled
// by sqrt(w_q) ... written as a 32 x 4 panel (3 components + a zero column) in shared memory; the 27 x 27 stiffness
// matrix accumulates as 16 tiles of 8 x 8 through mma.sync.m8n8k4.f64: K += P P^T.  Fragment layout of m8n8k4 (PTX ISA):
// A (8x4, row): lane holds A[lane / 4][lane % 4]; B (4x8, col): lane holds B[lane % 4][lane / 4]; C (8x8): lane holds
// C[lane / 4][2 (lane % 4) + {0, 1}].  With B = P^T the two operand fragments of tile (ti, tj) are P[8 ti + lane/4][lane%4]
// and P[8 tj + lane/4][lane%4]: one shared-memory load each.
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int Q2_WARPS = 4;
__global__ void __launch_bounds__(Q2_WARPS * 32) k_gblock_q2_dmma(GArgs G)
{
  constexpr int NB = 27, NV = 8;
  __shared__ double sX[Q2_WARPS][24];
  __shared__ double sP[Q2_WARPS][32][4];     // sqrt(w) * physical gradient | 0
  __shared__ double sV[Q2_WARPS][32];        // sqrt(w) * value
  __shared__ double sK[Q2_WARPS][32][33];    // the element matrix, for the row-wise scatter
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t cell = (int64_t)blockIdx.x * Q2_WARPS + wl;
  if (cell >= G.n_cells) return;             // (whole warps leave; no block-wide barrier below)
  const double *wts = G.tab, *geo = wts + G.nq, *val = geo + (size_t)G.nq * NV * 3, *der = val + (size_t)G.nq * NB;
  if (lane < 24) sX[wl][lane] = G.coords[cell * 24 + lane];
  int lid = 0;
  double u = 0.0, um = 0.0;
  if (lane < NB) {
    lid = G.lids[cell * NB + lane];
    if (G.x[0]) { const double v = G.x[0][lid]; u = G.kx * v; um = G.mx[0] * v; }
    if (G.x[1]) um = fma(G.mx[1], G.x[1][lid], um);
    if (G.x[2]) um = fma(G.mx[2], G.x[2][lid], um);
  }
  __syncwarp();
  double acc[16][2], macc[16][2];
#pragma unroll
  for (int t = 0; t < 16; ++t) { acc[t][0] = acc[t][1] = 0.0; macc[t][0] = macc[t][1] = 0.0; }
  const bool mass = G.cM != 0.0 || G.mx[0] != 0.0 || G.mx[1] != 0.0 || G.mx[2] != 0.0;
  double rsrc = 0.0;
  for (int q = 0; q < G.nq; ++q) {
    double J[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = 0.0;
#pragma unroll
    for (int n = 0; n < NV; ++n)
#pragma unroll
      for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int e = 0; e < 3; ++e) J[d * 3 + e] = fma(sX[wl][n * 3 + d], __ldg(geo + (q * NV + n) * 3 + e), J[d * 3 + e]);
    const double c0 = J[4] * J[8] - J[7] * J[5], c1 = J[6] * J[5] - J[3] * J[8], c2 = J[3] * J[7] - J[6] * J[4];
    const double det = J[0] * c0 + J[1] * c1 + J[2] * c2, id = 1.0 / det;
    const double w = det * __ldg(wts + q), sw = sqrt(w);
    double p0 = 0.0, p1 = 0.0, p2 = 0.0, v = 0.0;
    if (lane < NB) {
      const double *dr = der + (q * NB + lane) * 3;
      const double r0 = __ldg(dr), r1 = __ldg(dr + 1), r2 = __ldg(dr + 2);
      p0 = sw * (c0 * r0 + c1 * r1 + c2 * r2) * id;
      p1 = sw * ((J[2] * J[7] - J[1] * J[8]) * r0 + (J[0] * J[8] - J[2] * J[6]) * r1 + (J[1] * J[6] - J[0] * J[7]) * r2) * id;
      p2 = sw * ((J[1] * J[5] - J[2] * J[4]) * r0 + (J[2] * J[3] - J[0] * J[5]) * r1 + (J[0] * J[4] - J[1] * J[3]) * r2) * id;
      v = __ldg(val + q * NB + lane);
      rsrc = fma(w * v, G.src[0], rsrc);
      v *= sw;
    }
    sP[wl][lane][0] = p0; sP[wl][lane][1] = p1; sP[wl][lane][2] = p2; sP[wl][lane][3] = 0.0;
    sV[wl][lane] = v;
    __syncwarp();
    double fa[4], fv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) { fa[t] = sP[wl][8 * t + (lane >> 2)][lane & 3]; fv[t] = ((lane & 3) == 0) ? sV[wl][8 * t + (lane >> 2)] : 0.0; }
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        dmma_m8n8k4(acc[ti * 4 + tj][0], acc[ti * 4 + tj][1], fa[ti], fa[tj]);
        if (mass) dmma_m8n8k4(macc[ti * 4 + tj][0], macc[ti * 4 + tj][1], fv[ti], fv[tj]);
      }
    __syncwarp();
  }
  // residual r_a = sum_b K0[a][b] u_b + M0[a][b] um_b: through shared memory, then the Jacobian row = cK K0 + cM M0
  double r = rsrc;
  for (int pass = 0; pass < (mass ? 2 : 1); ++pass) {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        const int rr = 8 * ti + (lane >> 2), cc = 8 * tj + 2 * (lane & 3);
        sK[wl][rr][cc] = pass ? macc[ti * 4 + tj][0] : acc[ti * 4 + tj][0];
        sK[wl][rr][cc + 1] = pass ? macc[ti * 4 + tj][1] : acc[ti * 4 + tj][1];
      }
    sV[wl][lane] = pass ? um : u;
    __syncwarp();
    if (lane < NB)
      for (int b = 0; b < NB; ++b) r = fma(sK[wl][lane][b], sV[wl][b], r);
    __syncwarp();
  }
  // the Jacobian rows: lane a owns row a
  if (mass) {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        const int rr = 8 * ti + (lane >> 2), cc = 8 * tj + 2 * (lane & 3);
        sK[wl][rr][cc] = fma(G.cK, acc[ti * 4 + tj][0], G.cM * macc[ti * 4 + tj][0]);
        sK[wl][rr][cc + 1] = fma(G.cK, acc[ti * 4 + tj][1], G.cM * macc[ti * 4 + tj][1]);
      }
  } else {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        const int rr = 8 * ti + (lane >> 2), cc = 8 * tj + 2 * (lane & 3);
        sK[wl][rr][cc] = G.cK * acc[ti * 4 + tj][0];
        sK[wl][rr][cc + 1] = G.cK * acc[ti * 4 + tj][1];
      }
  }
  __syncwarp();
  if (G.scratch) {
    double *o = G.scratch + cell * NB * (NB + 1);
    if (G.jacobian)
      for (int i = lane; i < NB * NB; i += 32) { const int a = i / NB, b = i - a * NB; o[a * (NB + 1) + b] = sK[wl][a][b]; }
    if (lane < NB) o[lane * (NB + 1) + NB] = r;
    return;
  }
  if (lane < NB) {
    if (G.f) atomicAdd(G.f + lid, r);
    if (G.jacobian && G.A) {
      const int64_t base = G.rowptr[lid];
      const unsigned char *pl = G.plan + (cell * NB + lane) * NB;
      for (int k = 0; k < NB; ++k) {
        const unsigned p = pl[k];
        if (p != 0xFFu) atomicAdd(G.A + base + p, sK[wl][lane][k]);
      }
    }
  }
}

// ------------------------------------------------------------------ ghosted graph of a multi-block handle, on the device
// TpetraLinearObjFactory::buildGhostedGraph (disc-fe/src/lof/Panzer_TpetraLinearObjFactory_impl.hpp:558-650): every element
// couples all its DOFs; fillComplete sorts each row by local column index and merges duplicates.  Row-centric: the
// transpose of the LID tables (row -> (block, cell) list), then each row merges the LID lists of its cells.
constexpr int GG_MAXROW = 255;
struct GGBlocks { int n; const int *lids[8]; int nd[8]; int64_t n_cells[8]; };

__global__ void k_gg_count(GGBlocks B, int b, int *__restrict__ cnt)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < B.n_cells[b] * B.nd[b]) atomicAdd(&cnt[B.lids[b][i]], 1);
}
__global__ void k_gg_fill(GGBlocks B, int b, const int64_t *__restrict__ ptr, int *__restrict__ cursor, int64_t *__restrict__ adj)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B.n_cells[b] * B.nd[b]) return;
  const int row = B.lids[b][i];
  adj[ptr[row] + atomicAdd(&cursor[row], 1)] = ((int64_t)b << 56) | ((int64_t)(i % B.nd[b]) << 48) | (i / B.nd[b]);
}
// sort each row's (block, local dof, cell) list: the owner-computes gather then adds in a fixed order (reproducible)
__global__ void k_gg_sort(int64_t n_rows, const int64_t *__restrict__ ptr, int64_t *__restrict__ adj)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const int64_t b = ptr[r];
  const int n = (int)(ptr[r + 1] - b);
  for (int i = 1; i < n; ++i) {
    const int64_t v = adj[b + i];
    int k = i - 1;
    while (k >= 0 && adj[b + k] > v) { adj[b + k + 1] = adj[b + k]; --k; }
    adj[b + k + 1] = v;
  }
}

// Owner-computes second pass: one warp per matrix row.  The row accumulates in shared memory from the element rows
// k_gblock wrote (positions from the scatter plan), then leaves with coalesced stores: every entry of A and f is written
// exactly once, no atomics, no zero-fill pass, bitwise reproducible.
struct GRowBlocks { int n; const double *scratch[8]; const unsigned char *plan[8]; int nd[8]; };
constexpr int GR_WARPS = 8;
__global__ void __launch_bounds__(GR_WARPS * 32) k_gblock_rows(int64_t n_rows, GRowBlocks B, const int64_t *__restrict__ adj_ptr,
                                                               const int64_t *__restrict__ adj, const int64_t *__restrict__ rowptr,
                                                               int jacobian, double *__restrict__ f, double *__restrict__ A)
{
  __shared__ double srow[GR_WARPS][256];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * GR_WARPS + wl;
  if (r >= n_rows) return;
  const int64_t base = rowptr[r];
  const int len = (int)(rowptr[r + 1] - base);
  if (jacobian) for (int i = lane; i < len; i += 32) srow[wl][i] = 0.0;
  __syncwarp();
  double fr = 0.0;
  for (int64_t k = adj_ptr[r]; k < adj_ptr[r + 1]; ++k) {
    const int64_t e = adj[k];
    const int b = (int)(e >> 56), j = (int)((e >> 48) & 0xFF);
    const int64_t cell = e & 0x0000FFFFFFFFFFFFll;
    const int nd = B.nd[b];
    const double *row = B.scratch[b] + (cell * nd + j) * (nd + 1);
    if (jacobian) {
      const unsigned char *pl = B.plan[b] + (cell * nd + j) * nd;
      for (int i = lane; i < nd; i += 32) {
        const unsigned p = pl[i];
        if (p != 0xFFu) srow[wl][p] += row[i];          // distinct positions within one element row
      }
    }
    if (lane == 0) fr += row[nd];
    __syncwarp();
  }
  if (jacobian && A) for (int i = lane; i < len; i += 32) A[base + i] = srow[wl][i];
  if (lane == 0 && f) f[r] = fr;
}

template <bool WRITE>
__global__ void k_gg_rows(int64_t n_rows, GGBlocks B, const int64_t *__restrict__ adj_ptr, const int64_t *__restrict__ adj,
                          int64_t *__restrict__ cnt_or_ptr, int *__restrict__ colind, int *__restrict__ overflow)
{
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  int cols[GG_MAXROW + 1];
  int n = 0;
  for (int64_t k = adj_ptr[r]; k < adj_ptr[r + 1]; ++k) {
    const int b = (int)(adj[k] >> 56);
    const int64_t cell = adj[k] & 0x0000FFFFFFFFFFFFll;
    const int *l = B.lids[b] + cell * B.nd[b];
    for (int i = 0; i < B.nd[b]; ++i) {
      const int c = l[i];
      int lo = 0, hi = n;                      // first position with cols[pos] >= c
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (cols[mid] < c) lo = mid + 1; else hi = mid; }
      if (lo < n && cols[lo] == c) continue;
      if (n >= GG_MAXROW) { atomicExch(overflow, 1); continue; }
      for (int m = n; m > lo; --m) cols[m] = cols[m - 1];
      cols[lo] = c;
      ++n;
    }
  }
  if (!WRITE) cnt_or_ptr[r] = n;
  else {
    const int64_t b0 = cnt_or_ptr[r];
    for (int i = 0; i < n; ++i) colind[b0 + i] = cols[i];
  }
}

}  // namespace txasm
#include <cub/cub.cuh>
namespace txasm {

static int scan_i64(txasm_handle h, const int64_t *in, int64_t *out, int64_t n)
{
  size_t tb = 0;
  TX_CUDA(h, cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, n, h->stream));
  void *tmp = nullptr;
  TX_CUDA(h, cudaMalloc(&tmp, tb ? tb : 1));
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, in, out, n, h->stream);
  cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  TX_CUDA(h, e);
  return TXASM_OK;
}
__global__ void k_gg_widen(int64_t n, const int *__restrict__ in, int64_t *__restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

static int gg_blocks(txasm_handle h, GGBlocks &B, int64_t *n_entries)
{
  if (!h->gblocks || h->gblocks->b.empty()) return set_err(h, TXASM_ESTATE, "no element blocks");
  if (h->gblocks->b.size() > 8) return set_err(h, TXASM_EUNSUPPORTED, "at most 8 element blocks per handle");
  memset(&B, 0, sizeof(B));
  B.n = (int)h->gblocks->b.size();
  *n_entries = 0;
  for (int b = 0; b < B.n; ++b) {
    const GBlock &g = h->gblocks->b[b];
    B.lids[b] = g.d_lids; B.nd[b] = g.ndof; B.n_cells[b] = g.n_cells;
    *n_entries += g.n_cells * g.ndof;
  }
  return TXASM_OK;
}

// transpose of the LID tables: row -> its (block, local dof, cell) entries, sorted
static int gblocks_adjacency(txasm_handle h)
{
  GBlocks *GB = h->gblocks;
  if (GB->d_adj_ptr) return TXASM_OK;
  GGBlocks B;
  int64_t n_entries = 0;
  int rc = gg_blocks(h, B, &n_entries);
  if (rc) return rc;
  const int64_t nr = h->n_rows;
  int *cnt = nullptr;
  int64_t *cnt64 = nullptr;
  TX_CUDA(h, cudaMalloc(&cnt, sizeof(int) * (nr + 1)));
  TX_CUDA(h, cudaMalloc(&cnt64, sizeof(int64_t) * (nr + 1)));
  if ((rc = dev_alloc(h, &GB->d_adj_ptr, (size_t)nr + 1))) return rc;
  if ((rc = dev_alloc(h, &GB->d_adj, (size_t)std::max<int64_t>(n_entries, 1)))) return rc;
  TX_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int) * (nr + 1), h->stream));
  for (int b = 0; b < B.n; ++b) {
    const int64_t n = B.n_cells[b] * B.nd[b];
    k_gg_count<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(B, b, cnt);
  }
  k_gg_widen<<<(unsigned)((nr + 1 + 255) / 256), 256, 0, h->stream>>>(nr + 1, cnt, cnt64);
  if ((rc = scan_i64(h, cnt64, GB->d_adj_ptr, nr + 1))) return rc;
  TX_CUDA(h, cudaMemsetAsync(cnt, 0, sizeof(int) * (nr + 1), h->stream));
  for (int b = 0; b < B.n; ++b) {
    const int64_t n = B.n_cells[b] * B.nd[b];
    k_gg_fill<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(B, b, GB->d_adj_ptr, cnt, GB->d_adj);
  }
  k_gg_sort<<<(unsigned)((nr + 255) / 256), 256, 0, h->stream>>>(nr, GB->d_adj_ptr, GB->d_adj);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(cnt); cudaFree(cnt64);
  return TXASM_OK;
}

int gblocks_graph_build(txasm_handle h, int64_t *nnz_out)
{
  GGBlocks B;
  int64_t n_entries = 0;
  int rc = gg_blocks(h, B, &n_entries);
  if (rc) return rc;
  if ((rc = gblocks_adjacency(h))) return rc;
  const int64_t nr = h->n_rows;
  const int64_t *adj_ptr = h->gblocks->d_adj_ptr, *adj = h->gblocks->d_adj;
  int *d_over = nullptr, over = 0;
  int64_t *rcnt = nullptr, *rowptr = nullptr;
  TX_CUDA(h, cudaMalloc(&rcnt, sizeof(int64_t) * (nr + 1)));
  TX_CUDA(h, cudaMalloc(&d_over, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_over, 0, sizeof(int), h->stream));
  TX_CUDA(h, cudaMemsetAsync(rcnt, 0, sizeof(int64_t) * (nr + 1), h->stream));
  k_gg_rows<false><<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, B, adj_ptr, adj, rcnt, nullptr, d_over);
  if ((rc = dev_alloc(h, &rowptr, (size_t)nr + 1))) return rc;
  if ((rc = scan_i64(h, rcnt, rowptr, nr + 1))) return rc;
  int64_t nnz = 0;
  TX_CUDA(h, copy_to_device_sync(h, &nnz, rowptr + nr, sizeof(int64_t)));
  TX_CUDA(h, copy_to_device_sync(h, &over, d_over, sizeof(int)));
  if (over) { cudaFree(rcnt); cudaFree(d_over); return set_err(h, TXASM_EUNSUPPORTED, "graph_build: a row has more than %d entries", GG_MAXROW); }
  int *colind = nullptr;
  if ((rc = dev_alloc(h, &colind, (size_t)nnz))) return rc;
  k_gg_rows<true><<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, B, adj_ptr, adj, rowptr, colind, d_over);
  TX_CUDA(h, cudaGetLastError());
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(rcnt); cudaFree(d_over);
  h->d_rowptr = rowptr; h->d_colind = colind; h->nnz = nnz; h->have_graph = true;
  if (nnz_out) *nnz_out = nnz;
  return TXASM_OK;
}

// ------------------------------------------------------------------ host side
void gblocks_free(txasm_handle h)
{
  if (!h->gblocks) return;
  delete h->gblocks;
  h->gblocks = nullptr;
}

int gblocks_count(txasm_handle h) { return h->gblocks ? (int)h->gblocks->b.size() : 0; }

int gblocks_setup(txasm_handle h)
{
  if (!h->gblocks) return TXASM_OK;
  { int rc = gblocks_adjacency(h); if (rc) return rc; }
  int *d_flag = nullptr, flag = 0;
  TX_CUDA(h, cudaMalloc(&d_flag, sizeof(int)));
  TX_CUDA(h, cudaMemsetAsync(d_flag, 0, sizeof(int), h->stream));
  for (GBlock &B : h->gblocks->b) {
    if (B.d_plan) { dev_free(h, B.d_plan); B.d_plan = nullptr; }
    int rc = dev_alloc(h, &B.d_plan, (size_t)B.n_cells * B.ndof * B.ndof);
    if (rc) { cudaFree(d_flag); return rc; }
    const int64_t n = B.n_cells * B.ndof;
    k_gblock_plan<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(B.n_cells, B.ndof, B.d_lids, h->d_rowptr, h->d_colind, B.d_plan, d_flag);
  }
  TX_CUDA(h, cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  TX_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_flag);
  if (flag) return set_err(h, TXASM_EUNSUPPORTED, "a matrix row of an element block has more than 255 entries");
  return TXASM_OK;
}

int launch_gblocks(txasm_handle h, int jacobian, const txasm_inargs *in, const double *const x[3], double *f, double *A)
{
  const bool owner = !h->opt_block_atomic;           // element rows to scratch, then one gather pass; else searched atomics
  for (GBlock &B : h->gblocks->b) {
    if (B.op == 0) return set_err(h, TXASM_ESTATE, "an element block has no terms (txasm_gblock_terms_set)");
    if (owner && !B.d_scratch) { int rc = dev_alloc(h, &B.d_scratch, (size_t)B.n_cells * B.ndof * (B.ndof + 1)); if (rc) return rc; }
    GArgs g;
    memset(&g, 0, sizeof(g));
    g.scratch = owner ? B.d_scratch : nullptr;
    g.n_cells = B.n_cells; g.coords = B.d_coords; g.lids = B.d_lids; g.signs = B.d_signs; g.dofmap = B.d_dofmap; g.tab = B.d_tab;
    g.plan = B.d_plan; g.rowptr = h->d_rowptr; g.f = f; g.A = A; g.nq = B.nq; g.jacobian = jacobian;
    for (int v = 0; v < 3; ++v) g.x[v] = x[v];
    // seeds as GatherSolution_Tpetra<Jacobian> chooses them: beta for x, alpha for xdot, gamma for xdotdot (extension)
    if (B.op == GOP_DIFFUSION) {             // p = kappa, react, mass_dot, mass_dotdot, constant source
      g.kx = B.p[0]; g.mx[0] = B.p[1]; g.mx[1] = B.p[2]; g.mx[2] = B.p[3]; g.src[0] = B.p[4];
      g.cK = B.p[0] * in->beta; g.cM = B.p[1] * in->beta + B.p[2] * in->alpha + B.p[3] * in->gamma;
    } else if (B.op == GOP_ELASTICITY) {     // p = lambda, mu, rho (d2u/dt2), damping (du/dt), body force[3]
      g.lam = B.p[0]; g.mu = B.p[1]; g.kx = 1.0; g.mx[0] = 0.0; g.mx[1] = B.p[3]; g.mx[2] = B.p[2];
      g.src[0] = B.p[4]; g.src[1] = B.p[5]; g.src[2] = B.p[6];
      g.cK = in->beta; g.cM = B.p[2] * in->gamma + B.p[3] * in->alpha;
    } else {                                 // p = curl multiplier, mass multiplier, mass_dot multiplier, -, source[3]
      g.kx = B.p[0]; g.mx[0] = B.p[1]; g.mx[1] = B.p[2]; g.src[0] = B.p[4]; g.src[1] = B.p[5]; g.src[2] = B.p[6];
      g.cK = B.p[0] * in->beta; g.cM = B.p[1] * in->beta + B.p[2] * in->alpha;
    }
    for (int v = 1; v < 3; ++v)
      if (g.mx[v] != 0.0 && !x[v]) return set_err(h, TXASM_EINVAL, "an element block reads solution vector %d but it is NULL", v);
    if (!x[0]) return set_err(h, TXASM_EINVAL, "x is NULL");
#define TX_GB(OPK, NBK, NVK, NFK) { constexpr int cpc = GB_THREADS / ((NBK) * (NFK)); \
      k_gblock<OPK, NBK, NVK, NFK><<<(unsigned)((B.n_cells + cpc - 1) / cpc), GB_THREADS, 0, h->stream>>>(g); }
    if (B.elem == GE_HEX27_C2 && B.op == GOP_DIFFUSION && B.use_dmma && h->opt_dmma)
      k_gblock_q2_dmma<<<(unsigned)((B.n_cells + Q2_WARPS - 1) / Q2_WARPS), Q2_WARPS * 32, 0, h->stream>>>(g);
    else if (B.op == GOP_DIFFUSION && B.elem == GE_HEX8_C1) TX_GB(GOP_DIFFUSION, 8, 8, 1)
    else if (B.op == GOP_DIFFUSION && B.elem == GE_HEX27_C2) TX_GB(GOP_DIFFUSION, 27, 8, 1)
    else if (B.op == GOP_DIFFUSION && B.elem == GE_TET4_C1) TX_GB(GOP_DIFFUSION, 4, 4, 1)
    else if (B.op == GOP_DIFFUSION && B.elem == GE_TET10_C2) TX_GB(GOP_DIFFUSION, 10, 4, 1)
    else if (B.op == GOP_ELASTICITY && B.elem == GE_HEX8_C1) TX_GB(GOP_ELASTICITY, 8, 8, 3)
    else if (B.op == GOP_CURLCURL && B.elem == GE_HEX8_HCURL) TX_GB(GOP_CURLCURL, 12, 8, 1)
    else return set_err(h, TXASM_EUNSUPPORTED, "element %d with operator %d is not implemented", B.elem, B.op);
#undef TX_GB
    TX_CUDA(h, cudaGetLastError());
    h->launches += 1;
  }
  if (owner) {
    GRowBlocks R;
    memset(&R, 0, sizeof(R));
    R.n = (int)h->gblocks->b.size();
    for (int b = 0; b < R.n; ++b) { const GBlock &B = h->gblocks->b[b]; R.scratch[b] = B.d_scratch; R.plan[b] = B.d_plan; R.nd[b] = B.ndof; }
    k_gblock_rows<<<(unsigned)((h->n_rows + GR_WARPS - 1) / GR_WARPS), GR_WARPS * 32, 0, h->stream>>>(
        h->n_rows, R, h->gblocks->d_adj_ptr, h->gblocks->d_adj, h->d_rowptr, jacobian, f, A);
    TX_CUDA(h, cudaGetLastError());
    h->launches += 1;
  }
  return TXASM_OK;
}

}  // namespace txasm

using namespace txasm;

extern "C" {

int txasm_gblock_add(txasm_handle h, const txasm_block_desc *d, int64_t n_rows, int *block_id)
{
  if (!h || !d) return TXASM_EINVAL;
  if (h->sticky) return TXASM_ECUDA;
  TX_CUDA(h, cudaSetDevice(h->device));
  if (h->have_block) return set_err(h, TXASM_ESTATE, "gblock_add: the handle already holds a Q1 fast-path block (txasm_block_add)");
  int elem = 0;
  if (d->topology == TXASM_TOPO_HEX8 && d->basis == TXASM_BASIS_HGRAD_C1) elem = GE_HEX8_C1;
  else if (d->topology == TXASM_TOPO_HEX27 && d->basis == TXASM_BASIS_HGRAD_C2) elem = GE_HEX27_C2;
  else if (d->topology == TXASM_TOPO_TET4 && d->basis == TXASM_BASIS_HGRAD_C1) elem = GE_TET4_C1;
  else if (d->topology == TXASM_TOPO_TET10 && d->basis == TXASM_BASIS_HGRAD_C2) elem = GE_TET10_C2;
  else if (d->topology == TXASM_TOPO_HEX8 && d->basis == TXASM_BASIS_HCURL_I1) elem = GE_HEX8_HCURL;
  else return set_err(h, TXASM_EUNSUPPORTED, "topology %d with basis %d is not implemented", d->topology, d->basis);
  GBlock B;
  B.elem = elem; B.nb = elem_nb(elem); B.nv = elem_nv(elem); B.nfld = d->n_fields > 0 ? d->n_fields : 1; B.ndof = B.nb * B.nfld;
  B.n_cells = d->n_cells; B.deg = d->cubature_degree;
  if (d->dofs_per_cell != B.ndof) return set_err(h, TXASM_EINVAL, "gblock_add: %d DOFs per cell, expected %d fields x %d", d->dofs_per_cell, B.nfld, B.nb);
  if (B.nfld != 1 && B.nfld != 3) return set_err(h, TXASM_EUNSUPPORTED, "gblock_add: 1 or 3 fields per block");
  if (d->n_cells <= 0 || !d->lids || !d->cell_vertex_coords || n_rows <= 0) return set_err(h, TXASM_EINVAL, "gblock_add: bad arguments");
  if (h->gblocks && h->n_rows != n_rows) return set_err(h, TXASM_EINVAL, "gblock_add: n_rows differs from the first block's");
  double pts[GB_MAXQ][3], w[GB_MAXQ];
  B.nq = cubature(elem, B.deg, pts, w);
  if (B.nq <= 0 || B.nq > GB_MAXQ) return set_err(h, TXASM_EUNSUPPORTED, "cubature degree %d is not implemented for this topology", B.deg);
  const int vd = (elem == GE_HEX8_HCURL) ? 3 : 1;
  std::vector<double> tab((size_t)B.nq * (1 + B.nv * 3 + B.nb * vd + B.nb * 3));
  double *tw = tab.data(), *tg = tw + B.nq, *tv = tg + (size_t)B.nq * B.nv * 3, *td = tv + (size_t)B.nq * B.nb * vd;
  for (int q = 0; q < B.nq; ++q) {
    tw[q] = w[q];
    double gv[8], tmpv[36];
    ref_point(B.nv == 8 ? GE_HEX8_C1 : GE_TET4_C1, pts[q], gv, tg + (size_t)q * B.nv * 3);
    ref_point(elem, pts[q], tmpv, td + (size_t)q * B.nb * 3);
    for (int i = 0; i < B.nb * vd; ++i) tv[(size_t)q * B.nb * vd + i] = tmpv[i];
  }
  int rc;
  if ((rc = dev_alloc(h, &B.d_tab, tab.size()))) return rc;
  TX_CUDA(h, copy_to_device_sync(h, B.d_tab, tab.data(), sizeof(double) * tab.size()));
  // element DOF order: position of (field, basis) = field_offsets[field][basis], default interleaved (FieldAggPattern)
  std::vector<int> dofmap(B.ndof, -1);
  for (int fl = 0; fl < B.nfld; ++fl)
    for (int b = 0; b < B.nb; ++b) {
      const int pos = d->field_offsets ? d->field_offsets[fl * B.nb + b] : b * B.nfld + fl;
      if (pos < 0 || pos >= B.ndof || dofmap[pos] != -1) return set_err(h, TXASM_EINVAL, "gblock_add: field_offsets is not a permutation");
      dofmap[pos] = b | (fl << 8);
    }
  if ((rc = dev_alloc(h, &B.d_dofmap, (size_t)B.ndof))) return rc;
  TX_CUDA(h, copy_to_device_sync(h, B.d_dofmap, dofmap.data(), sizeof(int) * B.ndof));
  if ((rc = to_device(h, d->cell_vertex_coords, (size_t)B.n_cells * B.nv * 3, &B.d_coords))) return rc;
  if ((rc = to_device(h, d->lids, (size_t)B.n_cells * B.ndof, &B.d_lids))) return rc;
  if (d->orientation_signs && (rc = to_device(h, d->orientation_signs, (size_t)B.n_cells * B.nb, &B.d_signs))) return rc;
  if (elem == GE_HEX8_HCURL && !d->orientation_signs) return set_err(h, TXASM_EINVAL, "gblock_add: HCURL blocks need orientation_signs");
  if (!h->gblocks) h->gblocks = new GBlocks();
  h->gblocks->b.push_back(B);
  h->n_rows = n_rows;
  h->n_cells += B.n_cells;
  h->is_setup = false;
  if (block_id) *block_id = (int)h->gblocks->b.size() - 1;
  return TXASM_OK;
}

int txasm_gblock_terms_set(txasm_handle h, int block_id, int op, const double *params, int n_params)
{
  if (!h || !h->gblocks || block_id < 0 || block_id >= (int)h->gblocks->b.size()) return TXASM_EINVAL;
  if (op < GOP_DIFFUSION || op > GOP_CURLCURL || n_params < 0 || n_params > 8 || (n_params && !params)) return set_err(h, TXASM_EINVAL, "gblock_terms_set: bad arguments");
  GBlock &B = h->gblocks->b[block_id];
  const bool ok = (op == GOP_DIFFUSION && B.nfld == 1 && B.elem != GE_HEX8_HCURL) || (op == GOP_ELASTICITY && B.nfld == 3 && B.elem == GE_HEX8_C1) ||
                  (op == GOP_CURLCURL && B.elem == GE_HEX8_HCURL);
  if (!ok) return set_err(h, TXASM_EUNSUPPORTED, "operator %d does not fit the block's basis / field layout", op);
  B.op = op;
  for (int i = 0; i < 8; ++i) B.p[i] = i < n_params ? params[i] : 0.0;
  return TXASM_OK;
}

}  // extern "C"
