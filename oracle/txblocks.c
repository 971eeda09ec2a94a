/*
 * txblocks.c -- CPU ORACLE for the element blocks beyond the scalar Q1 hexahedron.  TEST INFRASTRUCTURE ONLY.
 *
 * BASELINE.json configs 3-5: Q2 hexahedra and P1/P2 tetrahedra (scalar diffusion), three interleaved HGRAD fields
 * (linear elastodynamics, second order in time), HCURL edge elements (curl-curl + mass).  What the reference offers for
 * them and what it does not is in SURVEY.md appendix B; the restatement follows the reference's evaluators wherever
 * they exist and says so where the build had to define the operator:
 *
 *   IntegrationValues2 (jac, jac_inv, jac_det, weighted_measure)      disc-fe/src/Panzer_IntegrationValues2.cpp:946-1221
 *   BasisValues2: HGRAD value / grad, HCURL value / curl, orientations disc-fe/src/Panzer_BasisValues2_impl.hpp:1036-1190,
 *                                                                      1264-1275 (HCURLtransformVALUE, J^-T), 1376-1521,
 *                                                                      1727-1735 (HCURLtransformCURL, J/det), 63-109
 *   edge orientation from global vertex ids                           disc-fe/src/Panzer_IntrepidOrientation.cpp:96-99
 *   basis selection                                                    disc-fe/src/Panzer_IntrepidBasisFactory.hpp:138-235
 *   DOF order of several fields in one element (interleaved per id)   dof-mgr/src/Panzer_FieldAggPattern.cpp:201-276
 *   GatherSolution (Fad seeds), DOF, DOFGradient, DOFCurl              disc-fe/src/evaluators/Panzer_GatherSolution_Tpetra_impl.hpp:
 *                                                                      547-649, Panzer_DOF_impl.hpp, Panzer_DOFGradient_impl.hpp:83-131,
 *                                                                      Panzer_DOFCurl_impl.hpp
 *   Integrator_GradBasisDotVector, _BasisTimesScalar,                  disc-fe/src/evaluators/Panzer_Integrator_*_impl.hpp
 *   _BasisTimesVector, _CurlBasisDotVector                             (CurlBasisDotVector: :364-400, 683-776)
 *   ScatterResidual_Tpetra (searched sumIntoValues)                   disc-fe/src/evaluators/Panzer_ScatterResidual_Tpetra_impl.hpp:374-440
 *   term lists: CurlLaplacian example                                  adapters-stk/example/CurlLaplacianExample/
 *                                                                      Example_CurlLaplacianEquationSet_impl.hpp:130-225
 *
 * Third-party definitions (Intrepid2 / Shards, not in the tree; checked by identities in
 * tests/test_oracle_blocks.py): Shards Hexahedron<27> and Tetrahedron<10> node orders, Basis_HGRAD_HEX_C2 / TET_C1 /
 * TET_C2 nodal Lagrange functions, Basis_HCURL_HEX_I1 (unit tangential trace on the own edge), tensor Gauss rules,
 * the tetrahedron rules of degree 1-3.  PARITY UNPINNED for every operator here: the reference has no elasticity
 * equation set, throws on mixed topologies and stores no golden values.
 */
#include "txoracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NFMAX 32
typedef struct { double val; double dx[NFMAX]; } bfad;

static void fad_zero(bfad *a) { memset(a, 0, sizeof(*a)); }
static void fad_axpy(bfad *y, double a, const bfad *x, int n) { y->val += a * x->val; for (int k = 0; k < n; ++k) y->dx[k] += a * x->dx[k]; }

/* ------------------------------------------------------------------ reference elements */
static const double HEX27_NODE[27][3] = {
  {-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1},
  {0,-1,-1},{1,0,-1},{0,1,-1},{-1,0,-1},            /* bottom edge mids (0,1)(1,2)(2,3)(3,0) */
  {-1,-1,0},{1,-1,0},{1,1,0},{-1,1,0},              /* vertical edge mids (0,4)(1,5)(2,6)(3,7) */
  {0,-1,1},{1,0,1},{0,1,1},{-1,0,1},                /* top edge mids */
  {0,0,0},{0,0,-1},{0,0,1},{-1,0,0},{1,0,0},{0,-1,0},{0,1,0}};   /* centre, z-, z+, x-, x+, y-, y+ */

static double lag2(double node, double t) { return node < -0.5 ? 0.5 * t * (t - 1.0) : (node > 0.5 ? 0.5 * t * (t + 1.0) : 1.0 - t * t); }
static double dlag2(double node, double t) { return node < -0.5 ? t - 0.5 : (node > 0.5 ? t + 0.5 : -2.0 * t); }

static const int TET10_EDGE[6][2] = {{0,1},{1,2},{0,2},{0,3},{1,3},{2,3}};
static const int HEX_EDGE[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};
static const double HEX8_S[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};

int orb_num_basis(int elem) { switch (elem) { case 1: return 8; case 2: return 27; case 3: return 4; case 4: return 10; case 5: return 12; } return -1; }
int orb_num_vertices(int elem) { return (elem == 3 || elem == 4) ? 4 : 8; }
static int is_vector_basis(int elem) { return elem == 5; }

/* basis at one reference point: val[nb] (scalar) or val[nb][3] (HCURL); der = grad[nb][3] (HGRAD) or curl[nb][3] (HCURL) */
void orb_ref_basis(int elem, const double *pt, double *val, double *der)
{
  const double x = pt[0], y = pt[1], z = pt[2];
  if (elem == 1) {
    for (int n = 0; n < 8; ++n) {
      const double ax = 1 + HEX8_S[n][0] * x, ay = 1 + HEX8_S[n][1] * y, az = 1 + HEX8_S[n][2] * z;
      val[n] = 0.125 * ax * ay * az;
      der[n * 3 + 0] = 0.125 * HEX8_S[n][0] * ay * az; der[n * 3 + 1] = 0.125 * ax * HEX8_S[n][1] * az; der[n * 3 + 2] = 0.125 * ax * ay * HEX8_S[n][2];
    }
  } else if (elem == 2) {
    for (int n = 0; n < 27; ++n) {
      const double *c = HEX27_NODE[n];
      const double lx = lag2(c[0], x), ly = lag2(c[1], y), lz = lag2(c[2], z);
      val[n] = lx * ly * lz;
      der[n * 3 + 0] = dlag2(c[0], x) * ly * lz; der[n * 3 + 1] = lx * dlag2(c[1], y) * lz; der[n * 3 + 2] = lx * ly * dlag2(c[2], z);
    }
  } else if (elem == 3 || elem == 4) {
    const double L[4] = {1 - x - y - z, x, y, z};
    const double dL[4][3] = {{-1,-1,-1},{1,0,0},{0,1,0},{0,0,1}};
    if (elem == 3) {
      for (int n = 0; n < 4; ++n) { val[n] = L[n]; for (int d = 0; d < 3; ++d) der[n * 3 + d] = dL[n][d]; }
    } else {
      for (int n = 0; n < 4; ++n) { val[n] = L[n] * (2 * L[n] - 1); for (int d = 0; d < 3; ++d) der[n * 3 + d] = (4 * L[n] - 1) * dL[n][d]; }
      for (int e = 0; e < 6; ++e) {
        const int i = TET10_EDGE[e][0], j = TET10_EDGE[e][1];
        val[4 + e] = 4 * L[i] * L[j];
        for (int d = 0; d < 3; ++d) der[(4 + e) * 3 + d] = 4 * (dL[i][d] * L[j] + L[i] * dL[j][d]);
      }
    }
  } else if (elem == 5) {
    /* Basis_HCURL_HEX_I1: function e is tangent to edge e (from its first to its second vertex) with unit tangential
       component on it and none on the other edges */
    memset(val, 0, sizeof(double) * 36); memset(der, 0, sizeof(double) * 36);
    for (int e = 0; e < 12; ++e) {
      const double *a = HEX8_S[HEX_EDGE[e][0]], *b = HEX8_S[HEX_EDGE[e][1]];
      int dir = 0; for (int d = 0; d < 3; ++d) if (a[d] != b[d]) dir = d;
      const int d1 = (dir + 1) % 3, d2 = (dir + 2) % 3;
      const double sgn = 0.5 * (b[dir] - a[dir]);             /* +-1: edge direction along +-dir */
      const double p[3] = {x, y, z};
      const double f1 = 1 + a[d1] * p[d1], f2 = 1 + a[d2] * p[d2];
      val[e * 3 + dir] = sgn * 0.25 * f1 * f2;
      /* curl of (phi e_dir): component d1 = d(phi)/d(d2) ... with the cyclic order (dir, d1, d2) */
      der[e * 3 + d1] = sgn * 0.25 * f1 * a[d2];             /*  d phi / d x_d2 */
      der[e * 3 + d2] = -sgn * 0.25 * a[d1] * f2;            /* -d phi / d x_d1 */
    }
  }
}

/* cubature: hexahedron = tensor Gauss-Legendre with deg/2+1 points per direction (x fastest); tetrahedron: rules of
   degree 1 (1 point), 2 (4 points), 3 (5 points, Keast) */
int orb_cubature(int elem, int deg, double *pts, double *wts)
{
  if (elem == 3 || elem == 4) {
    if (deg <= 1) { pts[0] = pts[1] = pts[2] = 0.25; wts[0] = 1.0 / 6.0; return 1; }
    if (deg == 2) {
      const double a = 0.58541019662496845446, b = 0.13819660112501051518;
      for (int q = 0; q < 4; ++q) { for (int d = 0; d < 3; ++d) pts[q * 3 + d] = b; wts[q] = 1.0 / 24.0; }
      pts[0 * 3 + 0] = a; pts[1 * 3 + 1] = a; pts[2 * 3 + 2] = a;      /* the fourth point has 1-x-y-z = a */
      return 4;
    }
    if (deg == 3) {
      pts[0] = pts[1] = pts[2] = 0.25; wts[0] = -2.0 / 15.0;
      for (int q = 1; q < 5; ++q) { for (int d = 0; d < 3; ++d) pts[q * 3 + d] = 1.0 / 6.0; wts[q] = 3.0 / 40.0; }
      pts[1 * 3 + 0] = 0.5; pts[2 * 3 + 1] = 0.5; pts[3 * 3 + 2] = 0.5;
      return 5;
    }
    return -1;
  }
  const int n = deg / 2 + 1;
  double gx[16], gw[16];
  if (orc_gauss_legendre(n, gx, gw)) return -1;
  int q = 0;
  for (int k = 0; k < n; ++k) for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i, ++q) {
    pts[q * 3 + 0] = gx[i]; pts[q * 3 + 1] = gx[j]; pts[q * 3 + 2] = gx[k]; wts[q] = gw[i] * gw[j] * gw[k];
  }
  return q;
}

/* geometry basis gradients at a point: hex8 trilinear or tet4 linear in the cell vertices (cell_vertex_coordinates) */
static void geom_grad(int nv, const double *pt, double *g /*[nv][3]*/)
{
  double v[27];
  if (nv == 8) orb_ref_basis(1, pt, v, g); else orb_ref_basis(3, pt, v, g);
}

/* ------------------------------------------------------------------ one cell through the evaluator chain */
typedef struct {
  int elem, op, deg, eval_type;
  double alpha, beta, gamma;
  double p[8];
} orb_spec_i;

static void sum_into_row(const int64_t *rowptr, const int *colind, double *A, int row, const int *cols, int n, const double *vals)
{
  const int64_t b = rowptr[row], e = rowptr[row + 1];
  for (int k = 0; k < n; ++k) {
    int64_t lo = b, hi = e - 1, at = -1;                  /* sorted row: binary search; absent column -> skipped */
    while (lo <= hi) { const int64_t mid = (lo + hi) / 2; if (colind[mid] == cols[k]) { at = mid; break; } if (colind[mid] < cols[k]) lo = mid + 1; else hi = mid - 1; }
    if (at >= 0) A[at] += vals[k];
  }
}

int orb_evaluate(const orb_spec *sp_, int64_t ne, const double *cell_coords, int ndof, const int *lids, const int *field_offsets,
                 const signed char *signs, const double *x, const double *xdot, const double *xdotdot,
                 const int64_t *rowptr, const int *colind, double *f, double *A)
{
  const orb_spec_i *sp = (const orb_spec_i *)sp_;
  const int elem = sp->elem, nb = orb_num_basis(elem), nv = orb_num_vertices(elem), vec = is_vector_basis(elem);
  const int nfld = (sp->op == 2) ? 3 : 1;
  if (nb < 0 || ndof != nb * nfld || ndof > NFMAX) return -1;
  double pts[64 * 3], wts[64];
  const int nq = orb_cubature(elem, sp->deg, pts, wts);
  if (nq <= 0 || nq > 64) return -2;
  const int jac = sp->eval_type == 1;
  /* reference tables (BasisValues2 keeps them per workset) */
  double *rv = (double *)malloc(sizeof(double) * nq * nb * 3), *rd = (double *)malloc(sizeof(double) * nq * nb * 3);
  double *gg = (double *)malloc(sizeof(double) * nq * nv * 3);
  for (int q = 0; q < nq; ++q) { orb_ref_basis(elem, pts + q * 3, rv + q * nb * (vec ? 3 : 1), rd + q * nb * 3); geom_grad(nv, pts + q * 3, gg + q * nv * 3); }
  bfad *U = (bfad *)malloc(sizeof(bfad) * 3 * ndof);      /* gathered DOFs: [vector (x, xdot, xdotdot)][dof] */
  bfad *R = (bfad *)malloc(sizeof(bfad) * ndof);
  for (int64_t c = 0; c < ne; ++c) {
    const double *X = cell_coords + c * nv * 3;
    const int *L = lids + c * ndof;
    /* GatherSolution: dof j of the element is basis b of field fld at offsets(fld, b); seed beta / alpha / gamma */
    const double *vecs[3] = {x, xdot, xdotdot};
    const double seeds[3] = {sp->beta, sp->alpha, sp->gamma};
    for (int v = 0; v < 3; ++v)
      for (int j = 0; j < ndof; ++j) {
        fad_zero(&U[v * ndof + j]);
        if (!vecs[v]) continue;
        U[v * ndof + j].val = vecs[v][L[j]];
        if (jac) U[v * ndof + j].dx[j] = seeds[v];
      }
    for (int j = 0; j < ndof; ++j) fad_zero(&R[j]);
    for (int q = 0; q < nq; ++q) {
      /* IntegrationValues2 */
      double J[3][3] = {{0}}, Ji[3][3];
      for (int d = 0; d < 3; ++d) for (int e = 0; e < 3; ++e) { double a = 0; for (int n = 0; n < nv; ++n) a += X[n * 3 + d] * gg[(q * nv + n) * 3 + e]; J[d][e] = a; }
      const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2], c1 = J[2][0] * J[1][2] - J[1][0] * J[2][2], c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
      const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
      Ji[0][0] = c0 / det; Ji[1][0] = c1 / det; Ji[2][0] = c2 / det;
      Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det; Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det; Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
      Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det; Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det; Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
      const double wm = det * wts[q];
      /* BasisValues2: physical basis values / derivatives at this point */
      double val[NFMAX][3], der[NFMAX][3];                 /* scalar basis: val[b][0] */
      for (int b = 0; b < nb; ++b) {
        if (!vec) {
          val[b][0] = rv[q * nb + b];
          for (int d = 0; d < 3; ++d) der[b][d] = Ji[0][d] * rd[(q * nb + b) * 3 + 0] + Ji[1][d] * rd[(q * nb + b) * 3 + 1] + Ji[2][d] * rd[(q * nb + b) * 3 + 2];
        } else {
          const double s = signs ? (double)signs[c * nb + b] : 1.0;      /* applyOrientations */
          for (int d = 0; d < 3; ++d) {
            val[b][d] = s * (Ji[0][d] * rv[(q * nb + b) * 3 + 0] + Ji[1][d] * rv[(q * nb + b) * 3 + 1] + Ji[2][d] * rv[(q * nb + b) * 3 + 2]);
            der[b][d] = s * (J[d][0] * rd[(q * nb + b) * 3 + 0] + J[d][1] * rd[(q * nb + b) * 3 + 1] + J[d][2] * rd[(q * nb + b) * 3 + 2]) / det;
          }
        }
      }
#define OFF(fld, b) (field_offsets ? field_offsets[(fld) * nb + (b)] : (b) * nfld + (fld))
      if (sp->op == 1) {
        /* scalar diffusion: DOFGradient, DOF; Integrator_GradBasisDotVector(kappa) + Integrator_BasisTimesScalar */
        bfad gu[3], u0, u1, u2;
        for (int d = 0; d < 3; ++d) fad_zero(&gu[d]);
        fad_zero(&u0); fad_zero(&u1); fad_zero(&u2);
        for (int b = 0; b < nb; ++b) {
          const int j = OFF(0, b);
          for (int d = 0; d < 3; ++d) fad_axpy(&gu[d], der[b][d], &U[j], ndof);
          fad_axpy(&u0, val[b][0], &U[j], ndof); fad_axpy(&u1, val[b][0], &U[ndof + j], ndof); fad_axpy(&u2, val[b][0], &U[2 * ndof + j], ndof);
        }
        for (int b = 0; b < nb; ++b) {
          const int j = OFF(0, b);
          for (int d = 0; d < 3; ++d) fad_axpy(&R[j], wm * der[b][d] * sp->p[0], &gu[d], ndof);
          fad_axpy(&R[j], wm * val[b][0] * sp->p[1], &u0, ndof);
          fad_axpy(&R[j], wm * val[b][0] * sp->p[2], &u1, ndof);
          fad_axpy(&R[j], wm * val[b][0] * sp->p[3], &u2, ndof);
          R[j].val += wm * val[b][0] * sp->p[4];
        }
      } else if (sp->op == 2) {
        /* linear elastodynamics (operator defined by the build): three HGRAD fields;
           strain from DOFGradient of each field, stress = lambda tr(eps) I + 2 mu eps (closure model), rows
           Integrator_GradBasisDotVector(stress_i), mass Integrator_BasisTimesScalar(rho, D2XDT2_u_i), damping on DXDT */
        bfad gu[3][3], a2[3], a1[3];
        for (int i = 0; i < 3; ++i) { fad_zero(&a2[i]); fad_zero(&a1[i]); for (int d = 0; d < 3; ++d) fad_zero(&gu[i][d]); }
        for (int i = 0; i < 3; ++i)
          for (int b = 0; b < nb; ++b) {
            const int j = OFF(i, b);
            for (int d = 0; d < 3; ++d) fad_axpy(&gu[i][d], der[b][d], &U[j], ndof);
            fad_axpy(&a1[i], val[b][0], &U[ndof + j], ndof); fad_axpy(&a2[i], val[b][0], &U[2 * ndof + j], ndof);
          }
        bfad tr, sig[3][3];
        fad_zero(&tr);
        for (int i = 0; i < 3; ++i) fad_axpy(&tr, 1.0, &gu[i][i], ndof);
        for (int i = 0; i < 3; ++i)
          for (int d = 0; d < 3; ++d) {
            fad_zero(&sig[i][d]);
            fad_axpy(&sig[i][d], sp->p[1], &gu[i][d], ndof); fad_axpy(&sig[i][d], sp->p[1], &gu[d][i], ndof);
            if (i == d) fad_axpy(&sig[i][d], sp->p[0], &tr, ndof);
          }
        for (int i = 0; i < 3; ++i)
          for (int b = 0; b < nb; ++b) {
            const int j = OFF(i, b);
            for (int d = 0; d < 3; ++d) fad_axpy(&R[j], wm * der[b][d], &sig[i][d], ndof);
            fad_axpy(&R[j], wm * val[b][0] * sp->p[2], &a2[i], ndof);
            fad_axpy(&R[j], wm * val[b][0] * sp->p[3], &a1[i], ndof);
            R[j].val += wm * val[b][0] * sp->p[4 + i];
          }
      } else if (sp->op == 3) {
        /* CurlLaplacian term list (Example_CurlLaplacianEquationSet_impl.hpp:163-212): Integrator_CurlBasisDotVector
           (p0) on CURL_EFIELD, Integrator_BasisTimesVector (p1) on EFIELD, (p2) on DXDT_EFIELD, constant source */
        bfad cu[3], e0[3], e1[3];
        for (int d = 0; d < 3; ++d) { fad_zero(&cu[d]); fad_zero(&e0[d]); fad_zero(&e1[d]); }
        for (int b = 0; b < nb; ++b) {
          const int j = OFF(0, b);
          for (int d = 0; d < 3; ++d) { fad_axpy(&cu[d], der[b][d], &U[j], ndof); fad_axpy(&e0[d], val[b][d], &U[j], ndof); fad_axpy(&e1[d], val[b][d], &U[ndof + j], ndof); }
        }
        for (int b = 0; b < nb; ++b) {
          const int j = OFF(0, b);
          for (int d = 0; d < 3; ++d) {
            fad_axpy(&R[j], wm * der[b][d] * sp->p[0], &cu[d], ndof);
            fad_axpy(&R[j], wm * val[b][d] * sp->p[1], &e0[d], ndof);
            fad_axpy(&R[j], wm * val[b][d] * sp->p[2], &e1[d], ndof);
            R[j].val += wm * val[b][d] * sp->p[4 + d];
          }
        }
      } else { free(rv); free(rd); free(gg); free(U); free(R); return -3; }
#undef OFF
    }
    /* ScatterResidual_Tpetra */
    for (int j = 0; j < ndof; ++j) {
      if (f) f[L[j]] += R[j].val;
      if (jac && A) sum_into_row(rowptr, colind, A, L[j], L, ndof, R[j].dx);
    }
  }
  free(rv); free(rd); free(gg); free(U); free(R);
  return 0;
}

/* ------------------------------------------------------------------ mesh helpers for the tests / bench */
/* Q2 node numbering of an inline hexahedral mesh: the (2NX+1)(2NY+1)(2NZ+1) lattice, lexicographic.  The reference takes
   edge / face / cell ids from STK-generated subcell entities (Panzer_STKConnManager.cpp:160-226), which cannot be
   reproduced without STK; this is the Cartesian rule SURVEY.md appendix B proposes. */
int orb_q2_hex_lids(int nx, int ny, int nz, int *lids /*[ne][27]*/)
{
  const int64_t MX = 2 * nx + 1, MY = 2 * ny + 1;
  int64_t e = 0;
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i, ++e)
    for (int n = 0; n < 27; ++n) {
      const int64_t I = 2 * i + 1 + (int)HEX27_NODE[n][0], Jn = 2 * j + 1 + (int)HEX27_NODE[n][1], K = 2 * k + 1 + (int)HEX27_NODE[n][2];
      lids[e * 27 + n] = (int)(I + MX * (Jn + MY * K));
    }
  return 0;
}

/* Edge numbering of an inline hexahedral mesh (HCURL I1): x-edges, then y-edges, then z-edges, each lexicographic; the
   sign of local edge e is +1 when its first vertex has the smaller global vertex id (Panzer_IntrepidOrientation.cpp:96-99). */
int orb_hcurl_hex_lids(int nx, int ny, int nz, int *lids /*[ne][12]*/, signed char *signs /*[ne][12]*/)
{
  const int64_t NXe = (int64_t)nx * (ny + 1) * (nz + 1), NYe = (int64_t)(nx + 1) * ny * (nz + 1);
  int64_t e = 0;
  for (int k = 0; k < nz; ++k) for (int j = 0; j < ny; ++j) for (int i = 0; i < nx; ++i, ++e)
    for (int ed = 0; ed < 12; ++ed) {
      const double *a = HEX8_S[HEX_EDGE[ed][0]], *b = HEX8_S[HEX_EDGE[ed][1]];
      int dir = 0; for (int d = 0; d < 3; ++d) if (a[d] != b[d]) dir = d;
      /* lattice position of the edge's lower end */
      const int lo[3] = {i + (a[0] > 0 && dir != 0), j + (a[1] > 0 && dir != 1), k + (a[2] > 0 && dir != 2)};
      int64_t id;
      if (dir == 0) id = lo[0] + (int64_t)nx * (lo[1] + (int64_t)(ny + 1) * lo[2]);
      else if (dir == 1) id = NXe + lo[0] + (int64_t)(nx + 1) * (lo[1] + (int64_t)ny * lo[2]);
      else id = NXe + NYe + lo[0] + (int64_t)(nx + 1) * (lo[1] + (int64_t)(ny + 1) * lo[2]);
      lids[e * 12 + ed] = (int)id;
      /* global vertex ids grow with the lattice coordinate, so the edge points "up" iff its direction sign is + */
      signs[e * 12 + ed] = (b[dir] > a[dir]) ? 1 : -1;
    }
  return 0;
}
