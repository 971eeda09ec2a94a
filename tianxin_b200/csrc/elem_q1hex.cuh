// elem_q1hex.cuh -- Q1 hexahedron element kernels (device functions).
//
// Fuses what the reference keeps in cached per-workset tables and evaluates on the fly, in
// registers: IntegrationValues2 (jac, jac_inv, jac_det, weighted_measure, ip_coordinates;
// disc-fe/src/Panzer_IntegrationValues2.cpp:946-1221), BasisValues2 (basis_scalar, grad_basis
// and their weighted forms; disc-fe/src/Panzer_BasisValues2_impl.hpp:1036-1190,1376-1521),
// DOFGradient / DOF (Panzer_DOFGradient_impl.hpp:83-131, Panzer_DOF_Functors.hpp:156-186) and
// Integrator_GradBasisDotVector / Integrator_BasisTimesScalar
// (Panzer_Integrator_GradBasisDotVector_impl.hpp:220-294, Panzer_Integrator_BasisTimesScalar_impl.hpp:
// 210-270).  The forward-mode (Sacado Fad) derivative of these linear integrands w.r.t. the
// element DOFs is the element matrix itself times the gather seed, so it is formed directly.
#pragma once
#include "txasm_internal.hpp"

namespace txasm {

// Shards Hexahedron<8> vertex signs on [-1,1]^3
__host__ __device__ constexpr int hex_sx(int n) { return ((n & 1) ^ ((n >> 1) & 1)) ? 1 : -1; }
__host__ __device__ constexpr int hex_sy(int n) { return (n & 2) ? 1 : -1; }
__host__ __device__ constexpr int hex_sz(int n) { return (n & 4) ? 1 : -1; }
__host__ __device__ constexpr int hex_s(int n, int d) { return d == 0 ? hex_sx(n) : (d == 1 ? hex_sy(n) : hex_sz(n)); }
// symmetric 8x8 storage, a<=b
__host__ __device__ constexpr int sym_idx(int a, int b) { return a <= b ? a * 8 - a * (a - 1) / 2 + (b - a) : b * 8 - b * (b - 1) / 2 + (a - b); }

#define TX_INV_SQRT3 0.57735026918962576451

// sin(2 pi x) for the built-in closure model: t = 2x - 2 rint(x) in [-1,1] is exact, fold to |t| <= 1/2 and
// evaluate the odd Taylor polynomial of sin(pi t) to t^21 (truncation 3e-16, a few ulp overall); ~half the
// instructions of sinpi(), which also serves huge / special arguments.
__device__ __forceinline__ double sin2pi_fast(double x)
{
  double t = 2.0 * (x - rint(x));                       // [-1, 1]
  t = (fabs(t) > 0.5) ? copysign(1.0, t) - t : t;       // sin(pi t) = sin(pi (sgn(t) - t))
  const double s = t * t;
  double p = 5.39266466260812895e-10;                   //  pi^21/21!   Horner in s = t^2
  p = fma(p, s, -2.29484289972698730e-08);              // -pi^19/19!
  p = fma(p, s, 7.95205400147551261e-07);               //  pi^17/17!
  p = fma(p, s, -2.19153534478302173e-05);              // -pi^15/15!
  p = fma(p, s, 4.66302805767612554e-04);               //  pi^13/13!
  p = fma(p, s, -7.37043094571435044e-03);              // -pi^11/11!
  p = fma(p, s, 8.21458866111282326e-02);               //  pi^9/9!
  p = fma(p, s, -5.99264529320792105e-01);              // -pi^7/7!
  p = fma(p, s, 2.55016403987734552e+00);               //  pi^5/5!
  p = fma(p, s, -5.16771278004997026e+00);              // -pi^3/3!
  p = fma(p, s, 3.14159265358979312e+00);               //  pi
  return p * t;
}

__device__ __forceinline__ double source_eval(int id, double x, double y, double z)
{
  switch (id) {
    case TXASM_SOURCE_SIN3:
      return 118.43525281307230 * sinpi(2.0 * x) * sinpi(2.0 * y) * sinpi(2.0 * z);  // 12 pi^2
    case TXASM_SOURCE_CONSTANT: return 1.0;
    default: return 0.0;
  }
}

// General trilinear hexahedron, 2x2x2 Gauss.  X: vertex coordinates; ug / um: the combined
// solution coefficients seen by the GRADGRAD / MASS integrands.
//   r[a]   = sum_q w_q detJ [ grad(phi_a).grad(ug) + phi_a um(q) + phi_a sum_s mult_s s_s(x_q) ]
//   K[a,b] = sum_q w_q detJ [ cK grad(phi_a).grad(phi_b) + cM phi_a phi_b ]      (a<=b, JAC only)
template <bool JAC>
__device__ __forceinline__ void elem_general(const double (&X)[8][3], const double (&ug)[8], const double (&um)[8],
                                             const FillCoef &c, int64_t cell, double (&K)[36], double (&r)[8])
{
#pragma unroll
  for (int a = 0; a < 8; ++a) r[a] = 0.0;
  if (JAC) {
#pragma unroll
    for (int i = 0; i < 36; ++i) K[i] = 0.0;
  }
  const bool do_grad = (c.kg[0] != 0.0 || c.kg[1] != 0.0 || c.kg[2] != 0.0) || (JAC && c.cK != 0.0);
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
      for (int e = 0; e < 3; ++e) {
        double s = 0.0;
#pragma unroll
        for (int n = 0; n < 8; ++n) s = fma(X[n][d], dN[n][e], s);
        J[d][e] = s;
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
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int d = 0; d < 3; ++d)
          G[n][d] = Ji[0][d] * dN[n][0] + Ji[1][d] * dN[n][1] + Ji[2][d] * dN[n][2];
      double gu[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int n = 0; n < 8; ++n)
#pragma unroll
        for (int d = 0; d < 3; ++d) gu[d] = fma(ug[n], G[n][d], gu[d]);
#pragma unroll
      for (int a = 0; a < 8; ++a) r[a] = fma(w, G[a][0] * gu[0] + G[a][1] * gu[1] + G[a][2] * gu[2], r[a]);
      if (JAC) {
        const double wk = w * c.cK;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = a; b < 8; ++b)
            K[sym_idx(a, b)] = fma(wk, G[a][0] * G[b][0] + G[a][1] * G[b][1] + G[a][2] * G[b][2], K[sym_idx(a, b)]);
      }
    }
    double sq = 0.0;
    if (c.has_mass) {
#pragma unroll
      for (int n = 0; n < 8; ++n) sq = fma(N[n], um[n], sq);
      sq *= wmass;
      if (JAC) {
        const double wm = wmass * c.cM;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = a; b < 8; ++b) K[sym_idx(a, b)] = fma(wm * N[a], N[b], K[sym_idx(a, b)]);
      }
    }
    if (c.n_src > 0) {
      double xq = 0.0, yq = 0.0, zq = 0.0;
#pragma unroll
      for (int n = 0; n < 8; ++n) { xq = fma(N[n], X[n][0], xq); yq = fma(N[n], X[n][1], yq); zq = fma(N[n], X[n][2], zq); }
      for (int s = 0; s < c.n_src; ++s) {
        const double v = (c.src_id[s] == TXASM_SOURCE_IP_ARRAY) ? c.src_ip[s][cell * 8 + q] : source_eval(c.src_id[s], xq, yq, zq);
        sq = fma(c.src_mult[s] * w0, v, sq);
      }
    }
    if (c.has_mass || c.n_src > 0) {
#pragma unroll
      for (int a = 0; a < 8; ++a) r[a] = fma(sq, N[a], r[a]);
    }
  }
}

// Parallelepiped test + constant Jacobian.  Jc[d][e] = (1/8) sum_n s_e(n) X[n][d] is the linear
// part of the trilinear map; the element is affine iff the higher-order coefficients vanish.
__device__ __forceinline__ double hex_nonaffinity(const double (&X)[8][3])
{
  // coefficients of xi*eta, eta*zeta, zeta*xi, xi*eta*zeta and the longest edge
  double dev = 0.0, len = 0.0;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double cxy = 0, cyz = 0, czx = 0, cxyz = 0, jx = 0, jy = 0, jz = 0;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      const double v = X[n][d];
      cxy += hex_sx(n) * hex_sy(n) * v; cyz += hex_sy(n) * hex_sz(n) * v; czx += hex_sz(n) * hex_sx(n) * v;
      cxyz += hex_sx(n) * hex_sy(n) * hex_sz(n) * v;
      jx += hex_sx(n) * v; jy += hex_sy(n) * v; jz += hex_sz(n) * v;
    }
    dev = fmax(dev, fmax(fmax(fabs(cxy), fabs(cyz)), fmax(fabs(czx), fabs(cxyz))));
    len = fmax(len, fmax(fabs(jx), fmax(fabs(jy), fabs(jz))));
  }
  return dev / len;   // both carry the same factor 1/8
}

// Affine geometry: Jc (constant Jacobian), detJ, Gs = detJ * Jinv Jinv^T (6 unique: xx,yy,zz,xy,yz,zx),
// Xc = centroid.
struct AffineGeom { double G[6]; double det; };

__device__ __forceinline__ void affine_geom(const double (&X)[8][3], double (&J)[3][3], AffineGeom &g)
{
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    double jx = 0, jy = 0, jz = 0;
#pragma unroll
    for (int n = 0; n < 8; ++n) { jx += hex_sx(n) * X[n][d]; jy += hex_sy(n) * X[n][d]; jz += hex_sz(n) * X[n][d]; }
    J[d][0] = 0.125 * jx; J[d][1] = 0.125 * jy; J[d][2] = 0.125 * jz;
  }
  const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
  const double c1 = J[2][0] * J[1][2] - J[1][0] * J[2][2];
  const double c2 = J[1][0] * J[2][1] - J[2][0] * J[1][1];
  const double det = J[0][0] * c0 + J[0][1] * c1 + J[0][2] * c2;
  const double idet = 1.0 / det;
  double Ji[3][3];  // Ji[e][d] = d xi_e / d x_d
  Ji[0][0] = c0 * idet; Ji[1][0] = c1 * idet; Ji[2][0] = c2 * idet;
  Ji[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * idet;
  Ji[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * idet;
  Ji[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * idet;
  Ji[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * idet;
  Ji[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * idet;
  Ji[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * idet;
  g.det = det;
  // G[e][e'] = det * sum_d Ji[e][d] Ji[e'][d]
  g.G[0] = det * (Ji[0][0] * Ji[0][0] + Ji[0][1] * Ji[0][1] + Ji[0][2] * Ji[0][2]);
  g.G[1] = det * (Ji[1][0] * Ji[1][0] + Ji[1][1] * Ji[1][1] + Ji[1][2] * Ji[1][2]);
  g.G[2] = det * (Ji[2][0] * Ji[2][0] + Ji[2][1] * Ji[2][1] + Ji[2][2] * Ji[2][2]);
  g.G[3] = det * (Ji[0][0] * Ji[1][0] + Ji[0][1] * Ji[1][1] + Ji[0][2] * Ji[1][2]);
  g.G[4] = det * (Ji[1][0] * Ji[2][0] + Ji[1][1] * Ji[2][1] + Ji[1][2] * Ji[2][2]);
  g.G[5] = det * (Ji[2][0] * Ji[0][0] + Ji[2][1] * Ji[0][1] + Ji[2][2] * Ji[0][2]);
}

// Exact reference integrals for a constant Jacobian (2x2x2 Gauss integrates them exactly, so this
// equals the quadrature sum of the reference up to rounding):
//   int grad(phi_a).grad(phi_b) = sum_e G_ee p_e c_f c_g / 8 + sum_{e<e'} G_ee' (s^e_a s^e'_b + s^e'_a s^e_b) c_f / 8
//   int phi_a phi_b             = det c_x c_y c_z / 8,        p_d = s^d_a s^d_b,  c_d = 1 + p_d/3
__host__ __device__ constexpr double aff_cd(int a, int b, int d) { return 1.0 + (hex_s(a, d) * hex_s(b, d)) / 3.0; }
__host__ __device__ constexpr double aff_kdiag(int a, int b, int e)
{
  return 0.125 * (hex_s(a, e) * hex_s(b, e)) * aff_cd(a, b, (e + 1) % 3) * aff_cd(a, b, (e + 2) % 3);
}
// pair index 3:(x,y) 4:(y,z) 5:(z,x)
__host__ __device__ constexpr double aff_koff(int a, int b, int pr)
{
  const int e = pr == 3 ? 0 : (pr == 4 ? 1 : 2);
  const int e2 = (e + 1) % 3, f = (e + 2) % 3;
  return 0.125 * (hex_s(a, e) * hex_s(b, e2) + hex_s(a, e2) * hex_s(b, e)) * aff_cd(a, b, f);
}
__host__ __device__ constexpr double aff_mass(int a, int b) { return 0.125 * aff_cd(a, b, 0) * aff_cd(a, b, 1) * aff_cd(a, b, 2); }

template <int A, int B>
__device__ __forceinline__ double aff_kab(const double (&G)[6])
{
  double t = G[0] * aff_kdiag(A, B, 0);
  t = fma(G[1], aff_kdiag(A, B, 1), t);
  t = fma(G[2], aff_kdiag(A, B, 2), t);
  if (aff_koff(A, B, 3) != 0.0) t = fma(G[3], aff_koff(A, B, 3), t);
  if (aff_koff(A, B, 4) != 0.0) t = fma(G[4], aff_koff(A, B, 4), t);
  if (aff_koff(A, B, 5) != 0.0) t = fma(G[5], aff_koff(A, B, 5), t);
  return t;
}

}  // namespace txasm
