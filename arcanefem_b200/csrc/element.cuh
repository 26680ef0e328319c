// Element-level physics: geometry -> (b x b) blocks K_ab of the element matrix.
//
// The reference expresses these as device lambdas handed to BSRFormat::assembleBilinear*
// (femutils/BSRFormat.h:218-236) or as per-format helper functions; each recomputes the
// determinant and re-gathers coordinates 3-4 times and divides 8-12 times per element
// (femutils/ArcaneFemFunctionsGpu.h:241-281,414-535).  Here every P1 element does one
// determinant, one reciprocal, cofactor gradients in registers:
//
//   Tet4 : e_i = m_i - m_0, c1 = e2 x e3, c2 = e3 x e1, c3 = e1 x e2, c0 = -(c1+c2+c3),
//          det = e1.c1, grad phi_a = c_a/det, V = |det|/6, s = V/det^2 = 1/(6|det|)
//          (c_a are the dPhi_a of modules/testlab/FemModule.h:425-428)
//   Tri3 : d0=(y1-y2,x2-x1) d1=(y2-y0,x0-x2) d2=(y0-y1,x1-x0), A2 = signed 2*area,
//          grad phi_a = d_a/A2, s = area/A2^2 = 1/(2|A2|)   (modules/testlab/FemModule.h:359-363)
//   Poisson    K_ab      = s (c_a . c_b)                                   (FemModule.h:455-462, FemModule.cc:267-299)
//   Elasticity K_ab(i,j) = s [ lambda c_a,i c_b,j + mu c_a,j c_b,i + mu (c_a.c_b) delta_ij ]
//                          (closed form of modules/elasticity/ElementMatrix.h:41-58,151-183)
//   Bilaplacian (Tri3, dofs (u1,u2)): K(2a,2b+1) = K(2a+1,2b) = s (d_a.d_b),
//                          K(2a+1,2b+1) = area (1 + delta_ab), K(2a,2b) = 0
//                          (modules/bilaplacian/ElementMatrix.h:37-45, massMatrix femutils/FemUtils.h:583-597)
//   P2 (Tri6/Tet10) Poisson: isoparametric, reference shape derivatives
//                          (femutils/ArcaneFemFunctions.h:3298-3319,3964-4005) and the order-2 Gauss
//                          rules of femutils/GaussQuadrature.h:141-176,203-243.
// Results agree with every reference formulation to rounding (tests: 1e-12 relative to the
// row's largest entry); the summation order differs, which is also true between the
// reference's own back-ends.
#pragma once

#include "afb_internal.h"

namespace afb {

struct ElemParams {
  double p0, p1;  // elasticity: lambda, mu
  double p2 = 0.0; // elastodynamics: {p0, p1, p2} = {c0, c1, c2}
  int flags;      // AFB_FLAG_*
  const double* cell_coef = nullptr; // per-cell multiplier of the Poisson element matrix (afb_set_cell_coefficient) or null
  double scale = 1.0;                // the current cell's multiplier (set by the cell-wise / node-wise kernels)
};

__device__ __forceinline__ void load3(const double* __restrict__ coords, int32_t n, double& x, double& y, double& z)
{
  const double* p = coords + 3 * (int64_t)n;
  x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
}

// ---------------------------------------------------------------------------------------------
// P1 tetrahedron
// ---------------------------------------------------------------------------------------------
struct Tet4Geom {
  double c[4][3];
  double s;    // 1/(6|det|)
  double vol;  // |det|/6
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[4])
  {
    double x0, y0, z0, x1, y1, z1, x2, y2, z2, x3, y3, z3;
    load3(coords, nd[0], x0, y0, z0);
    load3(coords, nd[1], x1, y1, z1);
    load3(coords, nd[2], x2, y2, z2);
    load3(coords, nd[3], x3, y3, z3);
    init_xyz(x0, y0, z0, x1, y1, z1, x2, y2, z2, x3, y3, z3);
  }
  __device__ __forceinline__ void init_xyz(double x0, double y0, double z0, double x1, double y1, double z1,
                                           double x2, double y2, double z2, double x3, double y3, double z3)
  {
    const double e1x = x1 - x0, e1y = y1 - y0, e1z = z1 - z0;
    const double e2x = x2 - x0, e2y = y2 - y0, e2z = z2 - z0;
    const double e3x = x3 - x0, e3y = y3 - y0, e3z = z3 - z0;
    c[1][0] = e2y * e3z - e2z * e3y; c[1][1] = e2z * e3x - e2x * e3z; c[1][2] = e2x * e3y - e2y * e3x;
    c[2][0] = e3y * e1z - e3z * e1y; c[2][1] = e3z * e1x - e3x * e1z; c[2][2] = e3x * e1y - e3y * e1x;
    c[3][0] = e1y * e2z - e1z * e2y; c[3][1] = e1z * e2x - e1x * e2z; c[3][2] = e1x * e2y - e1y * e2x;
    c[0][0] = -(c[1][0] + c[2][0] + c[3][0]);
    c[0][1] = -(c[1][1] + c[2][1] + c[3][1]);
    c[0][2] = -(c[1][2] + c[2][2] + c[3][2]);
    const double det = fabs(e1x * c[1][0] + e1y * c[1][1] + e1z * c[1][2]);
    vol = det * (1.0 / 6.0);
    s = 1.0 / (6.0 * det);
  }
  __device__ __forceinline__ double dot(int a, int b) const { return c[a][0] * c[b][0] + c[a][1] * c[b][1] + c[a][2] * c[b][2]; }
};

// ---------------------------------------------------------------------------------------------
// P1 triangle (planar mesh, z ignored: the reference's unsigned area is |(n1-n0)x(n2-n0)|/2,
// equal to |A2|/2 for planar z=const meshes)
// ---------------------------------------------------------------------------------------------
struct Tri3Geom {
  double c[3][2];
  double s;     // 1/(2*A2) with A2 signed or |A2|
  double area;  // signed or unsigned accordingly
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], bool signed_area)
  {
    double x0, y0, z0, x1, y1, z1, x2, y2, z2;
    load3(coords, nd[0], x0, y0, z0);
    load3(coords, nd[1], x1, y1, z1);
    load3(coords, nd[2], x2, y2, z2);
    (void)z0; (void)z1; (void)z2;
    init_xy(x0, y0, x1, y1, x2, y2, signed_area);
  }
  __device__ __forceinline__ void init_xy(double x0, double y0, double x1, double y1, double x2, double y2, bool signed_area)
  {
    c[0][0] = y1 - y2; c[0][1] = x2 - x1;
    c[1][0] = y2 - y0; c[1][1] = x0 - x2;
    c[2][0] = y0 - y1; c[2][1] = x1 - x0;
    double A2 = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    if (!signed_area) A2 = fabs(A2);
    area = 0.5 * A2;
    s = 1.0 / (2.0 * A2);
  }
  __device__ __forceinline__ double dot(int a, int b) const { return c[a][0] * c[b][0] + c[a][1] * c[b][1]; }
};

// ---------------------------------------------------------------------------------------------
// Element functors: NPC nodes, B dofs per node, block(a, b, out[B*B]) row-major
// ---------------------------------------------------------------------------------------------
struct Tet4Poisson {
  static constexpr int NPC = 4, B = 1, DIM = 3;
  Tet4Geom g;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[4], const ElemParams& p)
  {
    g.init(coords, nd);
    g.s *= p.scale;
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = g.dot(a, b) * g.s; }
  __device__ __forceinline__ double measure() const { return g.vol; }
};

struct Tri3Poisson {
  static constexpr int NPC = 3, B = 1, DIM = 2;
  Tri3Geom g;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], const ElemParams& p)
  {
    g.init(coords, nd, (p.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
    g.s *= p.scale;
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = g.dot(a, b) * g.s; }
  __device__ __forceinline__ double measure() const { return g.area; }
};

// alpha * stiffness + beta * consistent mass on P1 simplices (acoustics: modules/acoustics/ElementMatrix.h:14,29 with alpha = -1,
// beta = kc2; heat: modules/heat/ElementMatrix.h with alpha = lambda, beta = 1/dt): mass = meas / ((d+1)(d+2)) * (1 + delta_ab)
// (massMatrix(U,U) with U = 1, femutils/FemUtils.h:583-597).  A per-cell coefficient multiplies the stiffness part.
struct Tet4DiffReact {
  static constexpr int NPC = 4, B = 1, DIM = 3;
  Tet4Geom g;
  double mm;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[4], const ElemParams& p)
  {
    g.init(coords, nd);
    g.s *= p.p0 * p.scale;
    mm = p.p1 * g.vol * (1 / 20.);
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = g.dot(a, b) * g.s + (a == b ? 2.0 * mm : mm); }
  __device__ __forceinline__ double measure() const { return g.vol; }
};

struct Tri3DiffReact {
  static constexpr int NPC = 3, B = 1, DIM = 2;
  Tri3Geom g;
  double mm;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], const ElemParams& p)
  {
    g.init(coords, nd, false);
    g.s *= p.p0 * p.scale;
    mm = p.p1 * g.area * (1 / 12.);
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = g.dot(a, b) * g.s + (a == b ? 2.0 * mm : mm); }
  __device__ __forceinline__ double measure() const { return g.area; }
};

struct Tet4Elasticity {
  static constexpr int NPC = 4, B = 3, DIM = 3;
  Tet4Geom g;
  double ls, ms; // lambda*s, mu*s
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[4], const ElemParams& p)
  {
    g.init(coords, nd);
    ls = p.p0 * g.s;
    ms = p.p1 * g.s;
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[9]) const
  {
    const double dd = g.dot(a, b) * ms;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        o[i * 3 + j] = ls * g.c[a][i] * g.c[b][j] + ms * g.c[a][j] * g.c[b][i] + (i == j ? dd : 0.0);
  }
};

// elastodynamics (modules/elastodynamics/ElementMatrix.h): elasticity with (lambda, mu) = (c1, c2) + c0 * consistent mass on each component
struct Tet4Elastodynamics {
  static constexpr int NPC = 4, B = 3, DIM = 3;
  Tet4Elasticity e;
  double mm;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[4], const ElemParams& p)
  {
    ElemParams q = p;
    q.p0 = p.p1;
    q.p1 = p.p2;
    e.init(coords, nd, q);
    mm = p.p0 * e.g.vol * (1 / 20.);
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[9]) const
  {
    e.block(a, b, o);
    const double m = a == b ? 2.0 * mm : mm;
    o[0] += m;
    o[4] += m;
    o[8] += m;
  }
};

struct Tri3Elasticity {
  static constexpr int NPC = 3, B = 2, DIM = 2;
  Tri3Geom g;
  double ls, ms;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], const ElemParams& p)
  {
    g.init(coords, nd, false);
    ls = p.p0 * g.s;
    ms = p.p1 * g.s;
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[4]) const
  {
    const double dd = g.dot(a, b) * ms;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        o[i * 2 + j] = ls * g.c[a][i] * g.c[b][j] + ms * g.c[a][j] * g.c[b][i] + (i == j ? dd : 0.0);
  }
};

struct Tri3Elastodynamics {
  static constexpr int NPC = 3, B = 2, DIM = 2;
  Tri3Elasticity e;
  double mm;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], const ElemParams& p)
  {
    ElemParams q = p;
    q.p0 = p.p1;
    q.p1 = p.p2;
    e.init(coords, nd, q);
    mm = p.p0 * e.g.area * (1 / 12.);
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[4]) const
  {
    e.block(a, b, o);
    const double m = a == b ? 2.0 * mm : mm;
    o[0] += m;
    o[3] += m;
  }
};

struct Tri3Bilaplacian {
  static constexpr int NPC = 3, B = 2, DIM = 2;
  Tri3Geom g;
  __device__ __forceinline__ void init(const double* __restrict__ coords, const int32_t (&nd)[3], const ElemParams&) { g.init(coords, nd, false); }
  __device__ __forceinline__ void block(int a, int b, double (&o)[4]) const
  {
    const double sab = g.dot(a, b) * g.s;
    o[0] = 0.0;
    o[1] = sab;
    o[2] = sab;
    o[3] = (a == b) ? 2.0 * g.area : g.area;
  }
};

// ---------------------------------------------------------------------------------------------
// P2 simplices (Poisson): physical gradients at the Gauss points, pre-scaled by sqrt(w |J|)
// so that K_ab = sum_g G[g][a] . G[g][b]
// ---------------------------------------------------------------------------------------------
struct Tri6Poisson {
  static constexpr int NPC = 6, B = 1, DIM = 2;
  double G[3][6][2];
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[6], const ElemParams&)
  {
    double x[6], y[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) { double z; load3(coords, nd[a], x[a], y[a], z); (void)z; }
    const double gr[3] = { 0.5, 0.0, 0.5 }, gs[3] = { 0.5, 0.5, 0.0 };
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const double r = gr[q], sq = gs[q], t = 1.0 - r - sq;
      double dN[6][2];
      dN[0][0] = -3.0 + 4.0 * (r + sq); dN[0][1] = dN[0][0];
      dN[1][0] = -1.0 + 4.0 * r; dN[1][1] = 0.0;
      dN[2][0] = 0.0; dN[2][1] = -1.0 + 4.0 * sq;
      dN[3][0] = 4.0 * (t - r); dN[3][1] = -4.0 * r;
      dN[4][0] = 4.0 * sq; dN[4][1] = 4.0 * r;
      dN[5][0] = -4.0 * sq; dN[5][1] = 4.0 * (t - sq);
      double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        J00 += x[a] * dN[a][0]; J01 += x[a] * dN[a][1];
        J10 += y[a] * dN[a][0]; J11 += y[a] * dN[a][1];
      }
      const double det = J00 * J11 - J01 * J10;
      const double sc = sqrt(fabs(det) * (1.0 / 6.0)) / det;
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        G[q][a][0] = (J11 * dN[a][0] - J10 * dN[a][1]) * sc;
        G[q][a][1] = (-J01 * dN[a][0] + J00 * dN[a][1]) * sc;
      }
    }
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const
  {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 3; ++q) v += G[q][a][0] * G[q][b][0] + G[q][a][1] * G[q][b][1];
    o[0] = v;
  }
};

struct Tet10Poisson {
  static constexpr int NPC = 10, B = 1, DIM = 3;
  double G[4][10][3];
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[10], const ElemParams&)
  {
    double X[10][3];
#pragma unroll
    for (int a = 0; a < 10; ++a) load3(coords, nd[a], X[a][0], X[a][1], X[a][2]);
    const double a2 = 0.1381966011250105151795413165634361882280, b2 = 0.5854101966249684544613760503096914353161;
    for (int q = 0; q < 4; ++q) {
      const double x = (q == 3) ? b2 : a2, y = (q == 2) ? b2 : a2, z = (q == 1) ? b2 : a2;
      const double t = 1.0 - x - y - z, x4 = 4 * x, y4 = 4 * y, z4 = 4 * z, t4 = 4 * t;
      double dN[10][3] = { { 1. - t4, 1. - t4, 1. - t4 }, { x4 - 1., 0., 0. }, { 0., y4 - 1., 0. }, { 0., 0., z4 - 1. },
                           { t4 - x4, -x4, -x4 }, { y4, x4, 0. }, { -y4, t4 - y4, -y4 }, { -z4, -z4, t4 - z4 }, { z4, 0., x4 }, { 0., z4, y4 } };
      double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
      for (int a = 0; a < 10; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) J[i][j] += X[a][i] * dN[a][j];
      double C[3][3];
      C[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
      C[0][1] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
      C[0][2] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
      C[1][0] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
      C[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
      C[1][2] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
      C[2][0] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      C[2][1] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      C[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double det = J[0][0] * C[0][0] + J[0][1] * C[0][1] + J[0][2] * C[0][2];
      const double sc = sqrt(fabs(det) * (1.0 / 24.0)) / det;
#pragma unroll
      for (int a = 0; a < 10; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) G[q][a][i] = (C[i][0] * dN[a][0] + C[i][1] * dN[a][1] + C[i][2] * dN[a][2]) * sc;
    }
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const
  {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) v += G[q][a][0] * G[q][b][0] + G[q][a][1] * G[q][b][1] + G[q][a][2] * G[q][b][2];
    o[0] = v;
  }
};

// ---------------------------------------------------------------------------------------------
// slot search and value indexing
// ---------------------------------------------------------------------------------------------
// Q1 quadrilateral / hexahedron, Poisson (modules/poisson/ElementMatrixHexQuad.h:68-104, :197-236; geometry
// femutils/ArcaneFemFunctionsGpu.h:296-349, :554-586; reference gradients femutils/ShapeFunctions.h:123-129, :314-346).
// 2x2 (2x2x2) Gauss points at +-1/sqrt(3), weight 1: K_ab = sum_g detJ_g grad N_a . grad N_b.  The inverse Jacobian is applied
// as adjugate / det: G_a = adj(J)^T dN_a / det, so K_ab = sum_g (A_a . A_b) / det_g with A_a = adj-transformed reference gradient.
// ---------------------------------------------------------------------------------------------
// MASS: alpha * stiffness + beta * sum_gp N_a N_b detJ (modules/acoustics/ElementMatrixHexQuad.h), alpha = p0, beta = p1
template <bool MASS>
struct Quad4Op {
  static constexpr int NPC = 4, B = 1, DIM = 2;
  double K[4][4];
  double area;
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[4], const ElemParams& prm)
  {
    double x[4], y[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { double z; load3(coords, nd[a], x[a], y[a], z); (void)z; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) K[a][b] = 0.0;
    area = 0.0;
    const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
#pragma unroll
    for (int ixi = 0; ixi < 2; ++ixi)
#pragma unroll
      for (int ieta = 0; ieta < 2; ++ieta) {
        const double xi = gp[ixi], eta = gp[ieta];
        const double dxi[4] = { -0.25 * (1.0 - eta), 0.25 * (1.0 - eta), 0.25 * (1.0 + eta), -0.25 * (1.0 + eta) };
        const double det_[4] = { -0.25 * (1.0 - xi), -0.25 * (1.0 + xi), 0.25 * (1.0 + xi), 0.25 * (1.0 - xi) };
        double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          J00 += dxi[a] * x[a]; J01 += dxi[a] * y[a];
          J10 += det_[a] * x[a]; J11 += det_[a] * y[a];
        }
        const double det = J00 * J11 - J01 * J10;
        const double inv = (MASS ? prm.p0 * prm.scale : prm.scale) / det;
        double ax[4], ay[4]; // det * physical gradients
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          ax[a] = J11 * dxi[a] - J01 * det_[a];
          ay[a] = J00 * det_[a] - J10 * dxi[a];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = a; b < 4; ++b) K[a][b] += (ax[a] * ax[b] + ay[a] * ay[b]) * inv;
        if constexpr (MASS) {
          const double N[4] = { 0.25 * (1.0 - xi) * (1.0 - eta), 0.25 * (1.0 + xi) * (1.0 - eta), 0.25 * (1.0 + xi) * (1.0 + eta), 0.25 * (1.0 - xi) * (1.0 + eta) };
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) K[a][b] += N[a] * N[b] * (prm.p1 * det);
        }
        area += det;
      }
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < a; ++b) K[a][b] = K[b][a];
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = K[a][b]; }
  __device__ __forceinline__ double measure() const { return area; }
};
using Quad4Poisson = Quad4Op<false>;
using Quad4DiffReact = Quad4Op<true>;

template <bool MASS>
struct Hexa8Op {
  static constexpr int NPC = 8, B = 1, DIM = 3;
  double K[8][8];
  double vol;
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[8], const ElemParams& prm)
  {
    double x[8], y[8], z[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) load3(coords, nd[a], x[a], y[a], z[a]);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) K[a][b] = 0.0;
    vol = 0.0;
    const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
    const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      const double xi = gp[(g >> 2) & 1], eta = gp[(g >> 1) & 1], zeta = gp[g & 1];
      double dxi[8], det_[8], dze[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        dxi[a] = sx[a] * 0.125 * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
        det_[a] = sy[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sz[a] * zeta);
        dze[a] = sz[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta);
      }
      double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        J[0][0] += dxi[a] * x[a]; J[0][1] += dxi[a] * y[a]; J[0][2] += dxi[a] * z[a];
        J[1][0] += det_[a] * x[a]; J[1][1] += det_[a] * y[a]; J[1][2] += det_[a] * z[a];
        J[2][0] += dze[a] * x[a]; J[2][1] += dze[a] * y[a]; J[2][2] += dze[a] * z[a];
      }
      // adjugate (rows of det * J^-1)
      const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c02 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
      const double c10 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0], c12 = J[0][2] * J[1][0] - J[0][0] * J[1][2];
      const double c20 = J[1][0] * J[2][1] - J[1][1] * J[2][0], c21 = J[0][1] * J[2][0] - J[0][0] * J[2][1], c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
      const double det = J[0][0] * c00 + J[0][1] * c10 + J[0][2] * c20;
      const double inv = (MASS ? prm.p0 * prm.scale : prm.scale) / det;
      double ax[8], ay[8], az[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        ax[a] = c00 * dxi[a] + c01 * det_[a] + c02 * dze[a];
        ay[a] = c10 * dxi[a] + c11 * det_[a] + c12 * dze[a];
        az[a] = c20 * dxi[a] + c21 * det_[a] + c22 * dze[a];
      }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = a; b < 8; ++b) K[a][b] += (ax[a] * ax[b] + ay[a] * ay[b] + az[a] * az[b]) * inv;
      if constexpr (MASS) {
        double N[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) N[a] = 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = a; b < 8; ++b) K[a][b] += N[a] * N[b] * (prm.p1 * det);
      }
      vol += det;
    }
#pragma unroll
    for (int a = 1; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < a; ++b) K[a][b] = K[b][a];
  }
  __device__ __forceinline__ void block(int a, int b, double (&o)[1]) const { o[0] = K[a][b]; }
  __device__ __forceinline__ double measure() const { return vol; }
};
using Hexa8Poisson = Hexa8Op<false>;
using Hexa8DiffReact = Hexa8Op<true>;

// Quad4 / Hexa8 isotropic elasticity (modules/elasticity/ElementMatrixHexQuad.h:computeElementMatrix{Quad4,Hexa8}Base summed
// over the 2x2 / 2x2x2 Gauss rule): per Gauss point the block of nodes (a, b) is w [lambda g_a g_b^T + mu g_b g_a^T + mu (g_a.g_b) I]
// with the physical gradients g at the point and w = detJ -- the same block as on simplices, once per point.  The gradients of
// all points are kept (a 24 x 24 element matrix is never formed); these cells run through the cell-wise and node-wise variants.
template <int DIM_>
struct Q1Elasticity {
  static constexpr int DIM = DIM_, NPC = DIM_ == 2 ? 4 : 8, B = DIM_, NG = 1 << DIM_;
  double G[NG][NPC][DIM_]; // physical gradients
  double w[NG];            // detJ (Gauss weights are 1)
  double lam, mu, meas;
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[NPC], const ElemParams& p)
  {
    lam = p.p0;
    mu = p.p1;
    meas = 0.0;
    double x[NPC], y[NPC], z[NPC];
#pragma unroll
    for (int a = 0; a < NPC; ++a) load3(coords, nd[a], x[a], y[a], z[a]);
    const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      const double xi = gp[(g >> (DIM - 1)) & 1], eta = gp[(g >> (DIM - 2)) & 1], zeta = DIM == 3 ? gp[g & 1] : 0.0;
      double dxi[NPC], det_[NPC], dze[NPC];
#pragma unroll
      for (int a = 0; a < NPC; ++a) {
        const double sx = ((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0, sy = (a & 2) ? 1.0 : -1.0, sz = (a & 4) ? 1.0 : -1.0;
        const double fx = 1.0 + sx * xi, fy = 1.0 + sy * eta, fz = DIM == 3 ? 1.0 + sz * zeta : 1.0, s = DIM == 3 ? 0.125 : 0.25;
        dxi[a] = sx * s * fy * fz;
        det_[a] = sy * s * fx * fz;
        dze[a] = DIM == 3 ? sz * s * fx * fy : 0.0;
      }
      if constexpr (DIM == 2) {
        double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          J00 += dxi[a] * x[a]; J01 += dxi[a] * y[a];
          J10 += det_[a] * x[a]; J11 += det_[a] * y[a];
        }
        const double det = J00 * J11 - J01 * J10, inv = 1.0 / det;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          G[g][a][0] = (J11 * dxi[a] - J01 * det_[a]) * inv;
          G[g][a][1] = (J00 * det_[a] - J10 * dxi[a]) * inv;
        }
        w[g] = det;
        meas += det;
      }
      else {
        double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          J[0][0] += dxi[a] * x[a]; J[0][1] += dxi[a] * y[a]; J[0][2] += dxi[a] * z[a];
          J[1][0] += det_[a] * x[a]; J[1][1] += det_[a] * y[a]; J[1][2] += det_[a] * z[a];
          J[2][0] += dze[a] * x[a]; J[2][1] += dze[a] * y[a]; J[2][2] += dze[a] * z[a];
        }
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c02 = J[0][1] * J[1][2] - J[0][2] * J[1][1];
        const double c10 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0], c12 = J[0][2] * J[1][0] - J[0][0] * J[1][2];
        const double c20 = J[1][0] * J[2][1] - J[1][1] * J[2][0], c21 = J[0][1] * J[2][0] - J[0][0] * J[2][1], c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        const double det = J[0][0] * c00 + J[0][1] * c10 + J[0][2] * c20, inv = 1.0 / det;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          G[g][a][0] = (c00 * dxi[a] + c01 * det_[a] + c02 * dze[a]) * inv;
          G[g][a][1] = (c10 * dxi[a] + c11 * det_[a] + c12 * dze[a]) * inv;
          G[g][a][2] = (c20 * dxi[a] + c21 * det_[a] + c22 * dze[a]) * inv;
        }
        w[g] = det;
        meas += det;
      }
    }
  }
  __device__ void block(int a, int b, double (&o)[DIM_ * DIM_]) const
  {
#pragma unroll
    for (int k = 0; k < DIM * DIM; ++k) o[k] = 0.0;
#pragma unroll 1
    for (int g = 0; g < NG; ++g) {
      double dd = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i) dd += G[g][a][i] * G[g][b][i];
      const double lw = lam * w[g], mw = mu * w[g];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) o[i * DIM + j] += lw * G[g][a][i] * G[g][b][j] + mw * G[g][a][j] * G[g][b][i] + (i == j ? mw * dd : 0.0);
    }
  }
  __device__ __forceinline__ double measure() const { return meas; }
};
using Quad4Elasticity = Q1Elasticity<2>;
using Hexa8Elasticity = Q1Elasticity<3>;

// elastodynamics on Quad4 / Hexa8 (modules/elastodynamics/ElementMatrixHexQuad.h): Q1 elasticity with (lambda, mu) = (c1, c2) plus
// c0 * sum over the Gauss points of N_a N_b detJ on every component
template <int DIM_>
struct Q1Elastodynamics {
  static constexpr int DIM = DIM_, NPC = DIM_ == 2 ? 4 : 8, B = DIM_, NG = 1 << DIM_;
  Q1Elasticity<DIM_> e;
  double c0;
  __device__ void init(const double* __restrict__ coords, const int32_t (&nd)[NPC], const ElemParams& p)
  {
    ElemParams q = p;
    q.p0 = p.p1;
    q.p1 = p.p2;
    e.init(coords, nd, q);
    c0 = p.p0;
  }
  // shape function a at Gauss point g (same point order as Q1Elasticity::init)
  __device__ __forceinline__ static double shape(int g, int a)
  {
    const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
    const double xi = gp[(g >> (DIM - 1)) & 1], eta = gp[(g >> (DIM - 2)) & 1], zeta = DIM == 3 ? gp[g & 1] : 0.0;
    const double sx = ((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0, sy = (a & 2) ? 1.0 : -1.0, sz = (a & 4) ? 1.0 : -1.0;
    return (DIM == 3 ? 0.125 * (1.0 + sz * zeta) : 0.25) * (1.0 + sx * xi) * (1.0 + sy * eta);
  }
  __device__ void block(int a, int b, double (&o)[DIM_ * DIM_]) const
  {
    e.block(a, b, o);
    double m = 0.0;
#pragma unroll
    for (int g = 0; g < NG; ++g) m += shape(g, a) * shape(g, b) * e.w[g];
    m *= c0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) o[i * DIM + i] += m;
  }
  __device__ __forceinline__ double measure() const { return e.meas; }
};
using Quad4Elastodynamics = Q1Elastodynamics<2>;
using Hexa8Elastodynamics = Q1Elastodynamics<3>;

// ---------------------------------------------------------------------------------------------
// position of `col` in the ascending segment cols[lo,hi) (present by construction)
__device__ __forceinline__ int find_col(const int32_t* __restrict__ cols, int lo, int hi, int32_t col)
{
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(cols + mid) <= col) lo = mid; else hi = mid;
  }
  return lo;
}

// first index i in [0,n) with a[i] >= key (COO row search: femutils/CooFormatMatrix.h:308-353)
__device__ __forceinline__ int64_t lower_bound_i32(const int32_t* __restrict__ a, int64_t n, int32_t key)
{
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <int B, int LAYOUT>
__device__ __forceinline__ int64_t value_index(int rb, int nz, int p, int i, int j)
{
  if constexpr (LAYOUT == AFB_LAYOUT_PER_BLOCK) return (int64_t)p * (B * B) + i * B + j;
  else return (int64_t)rb * (B * B) + (int64_t)B * ((p - rb) + (int64_t)i * nz) + j;
}

} // namespace afb
