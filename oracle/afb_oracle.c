/*
 * afb_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of ArcaneFEM's bilinear-form assembly hot path (the path
 * named by BASELINE.json `north_star`).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / `--impl reference` legs may load this file's shared
 * object.  The product path (arcanefem_b200/csrc, include/afb200.h) never links,
 * imports or falls back to it.
 *
 * The reference itself cannot be compiled here (needs the external Arcane
 * framework >= 3.14.14, absent and not installable: see DESIGN.md), so every
 * function below restates the reference arithmetic and cites the file:line it
 * follows (paths relative to the reference root).  The oracle is pinned against
 * the reference's own golden solution vectors (modules/testlab/tests/ *.txt,
 * modules/elasticity/check/ *.txt, modules/bilaplacian/check/2d_test.txt,
 * modules/poisson/check/ *quad*.txt and *hexa.txt for Quad4 / Hexa8) by
 * tests/test_oracle_golden.py.  Pieces that no reference test pins (intra-row
 * column order, P2 tri/tet stiffness assembled into CSR) are marked
 * "parity unpinned" where they are defined.
 *
 * Conventions:
 *   coords  : AoS double[nb_node][3]  (Arcane VariableNodeReal3 layout)
 *   conn    : int32[nb_cell][npc]     (cnc.nodeId(cell,i))
 *   is_own  : uint8[nb_node] or NULL (= all owned) (ItemGenericInfoListView::isOwn)
 *   dof     : node_lid*b + component  (femutils/FemDoFsOnNodes.cc:79-111)
 *   rows    : int32[nb_row+1] (the reference keeps nb_row entries without the
 *             sentinel, femutils/CsrFormatMatrix.h:71-80; entry nb_row = nnz here)
 *   columns : ascending inside each row (canonical form; the reference's
 *             intra-row order is non-deterministic / connectivity-order dependent:
 *             parity unpinned)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* enums shared with include/afb200.h (same numeric values)                   */
/* ------------------------------------------------------------------------- */
enum { ORC_OP_POISSON = 0, ORC_OP_ELASTICITY = 1, ORC_OP_BILAPLACIAN = 2, ORC_OP_DIFFUSION_REACTION = 3, ORC_OP_ELASTODYNAMICS = 4 };
/* which reference formulation of the element matrix to follow */
enum {
  ORC_FORM_COMPACT = 0, /* modules/testlab/FemModule.h:342-463 (CSR/COO GPU back-ends)   */
  ORC_FORM_HOST = 1,    /* modules/testlab/FemModule.cc:1863-1970 (host CSR/COO/DOK)     */
  ORC_FORM_BSR = 2,     /* modules/testlab/FemModule.cc:267-299 + ArcaneFemFunctionsGpu.h */
  ORC_FORM_NODEWISE = 3 /* modules/testlab/NodeWiseCsrBiliAssembly.cc:185-296 / AF lambdas */
};
enum { ORC_LAYOUT_PER_BLOCK = 0, ORC_LAYOUT_PER_ROW = 1 };
enum { ORC_DIR_PENALTY = 0, ORC_DIR_WEAK_PENALTY = 1, ORC_DIR_ROW = 2, ORC_DIR_ROW_COLUMN = 3 };

typedef struct { double x, y, z; } r3;

static inline r3 r3_load(const double* c, int32_t n) { r3 r = { c[3 * (size_t)n], c[3 * (size_t)n + 1], c[3 * (size_t)n + 2] }; return r; }
static inline r3 r3_sub(r3 a, r3 b) { r3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
/* Arcane math::cross / math::dot (arcane/utils/Real3.h; trivial, restated) */
static inline r3 r3_cross(r3 u, r3 v) { r3 r = { u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x }; return r; }
static inline double r3_dot(r3 u, r3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }

/* ========================================================================= */
/* Geometry helpers of the BSR formulation                                    */
/* femutils/ArcaneFemFunctionsGpu.h:76-86 (area), :113-128 (volume),          */
/* :241-281 (tri gradients), :414-535 (tet gradients)                         */
/* ========================================================================= */
static double area_tri3_unsigned(r3 n0, r3 n1, r3 n2)
{
  r3 v = r3_cross(r3_sub(n1, n0), r3_sub(n2, n0));
  return sqrt(v.x * v.x + v.y * v.y + v.z * v.z) / 2.0; /* v.normL2()/2 */
}
static void grad_tri3(r3 n0, r3 n1, r3 n2, double dx[3], double dy[3])
{
  double A2 = ((n1.x - n0.x) * (n2.y - n0.y) - (n2.x - n0.x) * (n1.y - n0.y));
  dx[0] = (n1.y - n2.y) / A2; dx[1] = (n2.y - n0.y) / A2; dx[2] = (n0.y - n1.y) / A2;
  dy[0] = (n2.x - n1.x) / A2; dy[1] = (n0.x - n2.x) / A2; dy[2] = (n1.x - n0.x) / A2;
}
static double volume_tet4(r3 n0, r3 n1, r3 n2, r3 n3)
{
  r3 v0 = r3_sub(n1, n0), v1 = r3_sub(n2, n0), v2 = r3_sub(n3, n0);
  return fabs(r3_dot(v0, r3_cross(v1, v2))) / 6.0;
}
static void grad_tet4(r3 n0, r3 n1, r3 n2, r3 n3, double dx[4], double dy[4], double dz[4])
{
  r3 v0 = r3_sub(n1, n0), v1 = r3_sub(n2, n0), v2 = r3_sub(n3, n0);
  double V6 = fabs(r3_dot(v0, r3_cross(v1, v2)));
  dx[0] = (n1.y * (n3.z - n2.z) + n2.y * (n1.z - n3.z) + n3.y * (n2.z - n1.z)) / V6;
  dx[1] = (n0.y * (n2.z - n3.z) + n2.y * (n3.z - n0.z) + n3.y * (n0.z - n2.z)) / V6;
  dx[2] = (n0.y * (n3.z - n1.z) + n1.y * (n0.z - n3.z) + n3.y * (n1.z - n0.z)) / V6;
  dx[3] = (n0.y * (n1.z - n2.z) + n1.y * (n2.z - n0.z) + n2.y * (n0.z - n1.z)) / V6;
  dy[0] = (n1.z * (n3.x - n2.x) + n2.z * (n1.x - n3.x) + n3.z * (n2.x - n1.x)) / V6;
  dy[1] = (n0.z * (n2.x - n3.x) + n2.z * (n3.x - n0.x) + n3.z * (n0.x - n2.x)) / V6;
  dy[2] = (n0.z * (n3.x - n1.x) + n1.z * (n0.x - n3.x) + n3.z * (n1.x - n0.x)) / V6;
  dy[3] = (n0.z * (n1.x - n2.x) + n1.z * (n2.x - n0.x) + n2.z * (n0.x - n1.x)) / V6;
  dz[0] = (n1.x * (n3.y - n2.y) + n2.x * (n1.y - n3.y) + n3.x * (n2.y - n1.y)) / V6;
  dz[1] = (n0.x * (n2.y - n3.y) + n2.x * (n3.y - n0.y) + n3.x * (n0.y - n2.y)) / V6;
  dz[2] = (n0.x * (n3.y - n1.y) + n1.x * (n0.y - n3.y) + n3.x * (n1.y - n0.y)) / V6;
  dz[3] = (n0.x * (n1.y - n2.y) + n1.x * (n2.y - n0.y) + n2.x * (n0.y - n1.y)) / V6;
}

/* ========================================================================= */
/* P1 Poisson element matrices                                                */
/* ========================================================================= */

/* modules/testlab/FemModule.h:342-393 (_computeElementMatrixTRIA3GPU) */
static void ke_tri3_poisson_compact(r3 m0, r3 m1, r3 m2, double* K)
{
  double area = 0.5 * ((m1.x - m0.x) * (m2.y - m0.y) - (m2.x - m0.x) * (m1.y - m0.y));
  double d0x = m1.y - m2.y, d0y = m2.x - m1.x;
  double d1x = m2.y - m0.y, d1y = m0.x - m2.x;
  double d2x = m0.y - m1.y, d2y = m1.x - m0.x;
  double A2 = 2.0 * area;
  double b[2][3] = { { d0x / A2, d1x / A2, d2x / A2 }, { d0y / A2, d1y / A2, d2y / A2 } };
  for (int i = 0; i < 9; ++i) K[i] = 0.0;
  for (int i = 0; i < 3; ++i)
    for (int j = i; j < 3; ++j) {
      for (int k = 0; k < 2; ++k) K[i * 3 + j] += b[k][i] * b[k][j];
      K[i * 3 + j] *= area;
      K[j * 3 + i] = K[i * 3 + j];
    }
}

/* modules/testlab/FemModule.cc:1863-1904 (_computeElementMatrixTRIA3, host)   */
/* matrixMultiplication: femutils/FemUtils.h:303-319                           */
static void ke_tri3_poisson_host(r3 m0, r3 m1, r3 m2, double* K)
{
  double area = 0.5 * ((m1.x - m0.x) * (m2.y - m0.y) - (m2.x - m0.x) * (m1.y - m0.y));
  double b[2][3] = { { m1.y - m2.y, m2.y - m0.y, m0.y - m1.y }, { m2.x - m1.x, m0.x - m2.x, m1.x - m0.x } };
  double mul = 1.0 / (2.0 * area);
  for (int k = 0; k < 2; ++k) for (int i = 0; i < 3; ++i) b[k][i] *= mul;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double x = 0.0;
      for (int k = 0; k < 2; ++k) x += b[k][i] * b[k][j];
      K[i * 3 + j] = 0.0 + x;
    }
  for (int i = 0; i < 9; ++i) K[i] *= area;
}

/* modules/testlab/FemModule.cc:267-273 (computeElementMatrixTria3, BSR path)  */
static void ke_tri3_poisson_bsr(r3 n0, r3 n1, r3 n2, double* K)
{
  double area = area_tri3_unsigned(n0, n1, n2);
  double dx[3], dy[3];
  grad_tri3(n0, n1, n2, dx, dy);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      K[i * 3 + j] = area * (dx[i] * dx[j]) + area * (dy[i] * dy[j]);
}

/* B-matrix of the node-wise CSR back-end: modules/testlab/FemModule.h:468-492 */
static double bmat_tri3(r3 m0, r3 m1, r3 m2, double b[6])
{
  double area = 0.5 * ((m1.x - m0.x) * (m2.y - m0.y) - (m2.x - m0.x) * (m1.y - m0.y));
  double mul = 0.5 / area;
  b[0] = (m1.y - m2.y) * mul; b[1] = (m2.x - m1.x) * mul;
  b[2] = (m2.y - m0.y) * mul; b[3] = (m0.x - m2.x) * mul;
  b[4] = (m0.y - m1.y) * mul; b[5] = (m1.x - m0.x) * mul;
  return area;
}

/* modules/testlab/FemModule.h:398-463 (_computeElementMatrixTETRA4GPU) */
static void ke_tet4_poisson_compact(r3 m0, r3 m1, r3 m2, r3 m3, double* K)
{
  r3 v0 = r3_sub(m1, m0), v1 = r3_sub(m2, m0), v2 = r3_sub(m3, m0);
  double volume = fabs(r3_dot(v0, r3_cross(v1, v2))) / 6.0;
  r3 d0 = r3_cross(r3_sub(m2, m1), r3_sub(m1, m3));
  r3 d1 = r3_cross(r3_sub(m3, m0), r3_sub(m0, m2));
  r3 d2 = r3_cross(r3_sub(m1, m0), r3_sub(m0, m3));
  r3 d3 = r3_cross(r3_sub(m0, m1), r3_sub(m1, m2));
  double mul = 1.0 / (6.0 * volume);
  double b[3][4] = { { d0.x, d1.x, d2.x, d3.x }, { d0.y, d1.y, d2.y, d3.y }, { d0.z, d1.z, d2.z, d3.z } };
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) b[i][j] *= mul;
  for (int i = 0; i < 16; ++i) K[i] = 0.0;
  for (int i = 0; i < 4; ++i)
    for (int j = i; j < 4; ++j) {
      for (int k = 0; k < 3; ++k) K[i * 4 + j] += b[k][i] * b[k][j];
      K[i * 4 + j] *= volume;
      K[j * 4 + i] = K[i * 4 + j];
    }
}

/* modules/testlab/FemModule.cc:1909-1970 (_computeElementMatrixTETRA4, host) */
static void ke_tet4_poisson_host(r3 m0, r3 m1, r3 m2, r3 m3, double* K)
{
  double volume = volume_tet4(m0, m1, m2, m3);
  r3 d0 = r3_cross(r3_sub(m2, m1), r3_sub(m1, m3));
  r3 d1 = r3_cross(r3_sub(m3, m0), r3_sub(m0, m2));
  r3 d2 = r3_cross(r3_sub(m1, m0), r3_sub(m0, m3));
  r3 d3 = r3_cross(r3_sub(m0, m1), r3_sub(m1, m2));
  double b[3][4] = { { d0.x, d1.x, d2.x, d3.x }, { d0.y, d1.y, d2.y, d3.y }, { d0.z, d1.z, d2.z, d3.z } };
  double mul = 1.0 / (6.0 * volume);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) b[i][j] *= mul;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double x = 0.0;
      for (int k = 0; k < 3; ++k) x += b[k][i] * b[k][j];
      K[i * 4 + j] = 0.0 + x;
    }
  for (int i = 0; i < 16; ++i) K[i] *= volume;
}

/* modules/testlab/FemModule.cc:292-299 (computeElementMatrixTetra4, BSR path) */
static void ke_tet4_poisson_bsr(r3 n0, r3 n1, r3 n2, r3 n3, double* K)
{
  double volume = volume_tet4(n0, n1, n2, n3);
  double dx[4], dy[4], dz[4];
  grad_tet4(n0, n1, n2, n3, dx, dy, dz);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      K[i * 4 + j] = (volume * (dx[i] * dx[j]) + volume * (dy[i] * dy[j])) + volume * (dz[i] * dz[j]);
}

/* B-matrix of the node-wise CSR back-end: modules/testlab/FemModule.h:497-538 */
static double bmat_tet4(r3 m0, r3 m1, r3 m2, r3 m3, double b[12])
{
  r3 v0 = r3_sub(m1, m0), v1 = r3_sub(m2, m0), v2 = r3_sub(m3, m0);
  double volume = fabs(r3_dot(v0, r3_cross(v1, v2))) / 6.0;
  r3 d0 = r3_cross(r3_sub(m2, m1), r3_sub(m1, m3));
  r3 d1 = r3_cross(r3_sub(m3, m0), r3_sub(m0, m2));
  r3 d2 = r3_cross(r3_sub(m1, m0), r3_sub(m0, m3));
  r3 d3 = r3_cross(r3_sub(m0, m1), r3_sub(m1, m2));
  double mul = 1.0 / (6.0 * volume);
  b[0] = d0.x * mul; b[1] = d0.y * mul; b[2] = d0.z * mul;
  b[3] = d1.x * mul; b[4] = d1.y * mul; b[5] = d1.z * mul;
  b[6] = d2.x * mul; b[7] = d2.y * mul; b[8] = d2.z * mul;
  b[9] = d3.x * mul; b[10] = d3.y * mul; b[11] = d3.z * mul;
  return volume;
}

/* ========================================================================= */
/* Elasticity (interleaved DoFs ux,uy[,uz] per node)                          */
/* ========================================================================= */

/* modules/elasticity/ElementMatrix.h:41-58 (computeElementMatrixTria3Base)   */
/* outer product ^ : femutils/FemUtils.h:553-563                               */
static void ke_tri3_elasticity(r3 n0, r3 n1, r3 n2, double lambda, double mu, double* K)
{
  double dxu[3], dyu[3];
  grad_tri3(n0, n1, n2, dxu, dyu);
  double area = area_tri3_unsigned(n0, n1, n2);
  double dxUx[6] = { dxu[0], 0., dxu[1], 0., dxu[2], 0. };
  double dyUx[6] = { dyu[0], 0., dyu[1], 0., dyu[2], 0. };
  double dxUy[6] = { 0., dxu[0], 0., dxu[1], 0., dxu[2] };
  double dyUy[6] = { 0., dyu[0], 0., dyu[1], 0., dyu[2] };
  double s1[6], s2[6];
  for (int i = 0; i < 6; ++i) { s1[i] = dxUy[i] + dyUx[i]; s2[i] = dyUx[i] + dxUy[i]; }
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double nse = (lambda + 2 * mu) * ((dxUx[i] * dxUx[j]) + (dyUy[i] * dyUy[j])) * area;
      double ce = (lambda) * ((dyUy[i] * dxUx[j]) + (dxUx[i] * dyUy[j])) * area;
      double se = (mu) * (s1[i] * s2[j]) * area;
      K[i * 6 + j] = (nse + ce) + se;
    }
}

/* modules/elasticity/ElementMatrix.h:151-183 (computeElementMatrixTetra4Base) */
static void ke_tet4_elasticity(r3 n0, r3 n1, r3 n2, r3 n3, double lambda, double mu, double* K)
{
  double dxu[4], dyu[4], dzu[4];
  grad_tet4(n0, n1, n2, n3, dxu, dyu, dzu);
  double volume = volume_tet4(n0, n1, n2, n3);
  double dxUx[12], dyUx[12], dzUx[12], dxUy[12], dyUy[12], dzUy[12], dxUz[12], dyUz[12], dzUz[12];
  for (int i = 0; i < 12; ++i)
    dxUx[i] = dyUx[i] = dzUx[i] = dxUy[i] = dyUy[i] = dzUy[i] = dxUz[i] = dyUz[i] = dzUz[i] = 0.;
  for (int a = 0; a < 4; ++a) {
    dxUx[3 * a] = dxu[a]; dyUx[3 * a] = dyu[a]; dzUx[3 * a] = dzu[a];
    dxUy[3 * a + 1] = dxu[a]; dyUy[3 * a + 1] = dyu[a]; dzUy[3 * a + 1] = dzu[a];
    dxUz[3 * a + 2] = dxu[a]; dyUz[3 * a + 2] = dyu[a]; dzUz[3 * a + 2] = dzu[a];
  }
  double sxy1[12], sxy2[12], syz1[12], syz2[12], sxz1[12], sxz2[12];
  for (int i = 0; i < 12; ++i) {
    sxy1[i] = dxUy[i] + dyUx[i]; sxy2[i] = dyUx[i] + dxUy[i];
    syz1[i] = dzUy[i] + dyUz[i]; syz2[i] = dyUz[i] + dzUy[i];
    sxz1[i] = dxUz[i] + dzUx[i]; sxz2[i] = dzUx[i] + dxUz[i];
  }
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) {
      double nse = (lambda + 2 * mu) * (((dxUx[i] * dxUx[j]) + (dyUy[i] * dyUy[j])) + (dzUz[i] * dzUz[j])) * volume;
      double ce = (lambda) * ((((((dyUy[i] * dxUx[j]) + (dxUx[i] * dyUy[j])) + (dzUz[i] * dxUx[j])) + (dxUx[i] * dzUz[j])) + (dyUy[i] * dzUz[j])) + (dzUz[i] * dyUy[j])) * volume;
      double se = (mu) * (((sxy1[i] * sxy2[j]) + (syz1[i] * syz2[j])) + (sxz1[i] * sxz2[j])) * volume;
      K[i * 12 + j] = (nse + ce) + se;
    }
}

/* modules/bilaplacian/ElementMatrix.h:30-47; massMatrix femutils/FemUtils.h:583-597 */
static void ke_tri3_bilaplacian(r3 n0, r3 n1, r3 n2, double* K)
{
  double dxu[3], dyu[3];
  grad_tri3(n0, n1, n2, dxu, dyu);
  double area = area_tri3_unsigned(n0, n1, n2);
  double Uy[6] = { 0., 1., 0., 1., 0., 1. };
  double dxUx[6] = { dxu[0], 0., dxu[1], 0., dxu[2], 0. };
  double dyUx[6] = { dyu[0], 0., dyu[1], 0., dyu[2], 0. };
  double dxUy[6] = { 0., dxu[0], 0., dxu[1], 0., dxu[2] };
  double dyUy[6] = { 0., dyu[0], 0., dyu[1], 0., dyu[2] };
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) {
      double mm = Uy[i] * Uy[j];
      if (i == j) mm *= 2.;
      K[i * 6 + j] = (((dxUx[i] * dxUy[j]) + (dyUx[i] * dyUy[j])) * area + ((dxUy[i] * dxUx[j]) + (dyUy[i] * dyUx[j])) * area) + mm * area;
    }
}

/* ========================================================================= */
/* P2 simplex stiffness (Poisson)                                             */
/* Shape-function derivatives: femutils/ArcaneFemFunctions.h:3298-3319 (Tri6), */
/* :3964-4005 (Tet10); Gauss rules femutils/GaussQuadrature.h:141-176 (order 2 */
/* triangle: 3 mid-edge-style points, weight 1/6), :203-243 (order 2 tet: 4    */
/* points a2/b2, weight 1/24).  The reference has NO P2 tri/tet stiffness      */
/* assembly into CSR/BSR: "parity unpinned" — this is  */
/* the isoparametric stiffness built from the reference's own shape functions  */
/* and quadrature tables.                                                      */
/* ========================================================================= */
static void tri6_dshape(int inod, double ri, double si, double d[2])
{
  double ti = 1. - ri - si;
  switch (inod) {
  case 0: { double wi = -3. + 4. * (ri + si); d[0] = wi; d[1] = wi; break; }
  case 1: d[0] = -1. + 4. * ri; d[1] = 0.; break;
  case 2: d[0] = 0.; d[1] = -1. + 4. * si; break;
  case 3: d[0] = 4. * (ti - ri); d[1] = -4. * ri; break;
  case 4: d[0] = 4. * si; d[1] = 4. * ri; break;
  default: d[0] = -4. * si; d[1] = 4. * (ti - si); break;
  }
}
static void tet10_dshape(int inod, double x, double y, double z, double d[3])
{
  double t = 1. - x - y - z, x4 = 4 * x, y4 = 4 * y, z4 = 4 * z, t4 = 4 * t;
  switch (inod) {
  case 0: d[0] = 1. - t4; d[1] = 1. - t4; d[2] = 1. - t4; break;
  case 1: d[0] = x4 - 1.; d[1] = 0.; d[2] = 0.; break;
  case 2: d[0] = 0.; d[1] = y4 - 1.; d[2] = 0.; break;
  case 3: d[0] = 0.; d[1] = 0.; d[2] = z4 - 1.; break;
  case 4: d[0] = t4 - x4; d[1] = -x4; d[2] = -x4; break;
  case 5: d[0] = y4; d[1] = x4; d[2] = 0.; break;
  case 6: d[0] = -y4; d[1] = t4 - y4; d[2] = -y4; break;
  case 8: d[0] = z4; d[1] = 0.; d[2] = x4; break;
  case 9: d[0] = 0.; d[1] = z4; d[2] = y4; break;
  default: d[0] = -z4; d[1] = -z4; d[2] = t4 - z4; break; /* inod == 7 */
  }
}

static void ke_tri6_poisson(const r3* m, double* K)
{
  /* GaussQuadrature.h:141-176, order index 1: xg1={xh,0,xh} xg2={xh,xh,0}, wg=1/6 */
  const double xh = 0.5;
  const double gr[3] = { xh, 0., xh }, gs[3] = { xh, xh, 0. };
  const double w = 1. / 6.;
  for (int i = 0; i < 36; ++i) K[i] = 0.0;
  for (int g = 0; g < 3; ++g) {
    double dN[6][2];
    for (int a = 0; a < 6; ++a) tri6_dshape(a, gr[g], gs[g], dN[a]);
    double J[2][2] = { { 0, 0 }, { 0, 0 } }; /* J[i][j] = d x_i / d xi_j */
    for (int a = 0; a < 6; ++a) {
      J[0][0] += m[a].x * dN[a][0]; J[0][1] += m[a].x * dN[a][1];
      J[1][0] += m[a].y * dN[a][0]; J[1][1] += m[a].y * dN[a][1];
    }
    double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double inv = 1.0 / det;
    /* physical gradient g_a = J^{-T} dN_a */
    double gx[6], gy[6];
    for (int a = 0; a < 6; ++a) {
      gx[a] = (J[1][1] * dN[a][0] - J[1][0] * dN[a][1]) * inv;
      gy[a] = (-J[0][1] * dN[a][0] + J[0][0] * dN[a][1]) * inv;
    }
    double wd = w * fabs(det);
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b)
        K[a * 6 + b] += (gx[a] * gx[b] + gy[a] * gy[b]) * wd;
  }
}

static void ke_tet10_poisson(const r3* m, double* K)
{
  /* GaussQuadrature.h:203-243 order index 1: (a2,a2,a2),(a2,a2,b2)... weight 1/24 */
  const double a2 = (5. - sqrt(5.)) / 20., b2 = (5. + 3. * sqrt(5.)) / 20.;
  const double gx_[4] = { a2, a2, a2, b2 }, gy_[4] = { a2, a2, b2, a2 }, gz_[4] = { a2, b2, a2, a2 };
  const double w = 1. / 24.;
  for (int i = 0; i < 100; ++i) K[i] = 0.0;
  for (int g = 0; g < 4; ++g) {
    double dN[10][3];
    for (int a = 0; a < 10; ++a) tet10_dshape(a, gx_[g], gy_[g], gz_[g], dN[a]);
    double J[3][3] = { { 0 } };
    for (int a = 0; a < 10; ++a)
      for (int j = 0; j < 3; ++j) {
        J[0][j] += m[a].x * dN[a][j];
        J[1][j] += m[a].y * dN[a][j];
        J[2][j] += m[a].z * dN[a][j];
      }
    /* cofactors C[i][j] of J; J^{-1} = C^T/det ; g_a = J^{-T} dN_a = C dN_a / det */
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
    double det = J[0][0] * C[0][0] + J[0][1] * C[0][1] + J[0][2] * C[0][2];
    double inv = 1.0 / det;
    double g3[10][3];
    for (int a = 0; a < 10; ++a)
      for (int i = 0; i < 3; ++i)
        g3[a][i] = (C[i][0] * dN[a][0] + C[i][1] * dN[a][1] + C[i][2] * dN[a][2]) * inv;
    double wd = w * fabs(det);
    for (int a = 0; a < 10; ++a)
      for (int b = 0; b < 10; ++b)
        K[a * 10 + b] += (g3[a][0] * g3[b][0] + g3[a][1] * g3[b][1] + g3[a][2] * g3[b][2]) * wd;
  }
}


/* ---------------------------------------------------------------------------
 * Q1 quadrilateral / hexahedron, Poisson (modules/poisson/ElementMatrixHexQuad.h:33-66, :159-195):
 * 2x2 (2x2x2) Gauss points at +-1/sqrt(3), weights 1; at every point the reference gradients
 * (femutils/ShapeFunctions.h:123-129, :314-346), the Jacobian J = sum_a dN_a (x) x_a, its determinant and
 * inverse (femutils/ArcaneFemFunctionsGpu.h:296-349, :554-586), physical gradients, and
 *   ae += (dxU ^ dxU) * w + (dyU ^ dyU) * w (+ (dzU ^ dzU) * w),  w = detJ.
 * ------------------------------------------------------------------------- */
static void ke_quad4_poisson(const r3* m, double* K)
{
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  for (int i = 0; i < 16; ++i) K[i] = 0.0;
  for (int ixi = 0; ixi < 2; ++ixi)
    for (int ieta = 0; ieta < 2; ++ieta) {
      const double xi = gp[ixi], eta = gp[ieta];
      const double dxi[4] = { -0.25 * (1.0 - eta), 0.25 * (1.0 - eta), 0.25 * (1.0 + eta), -0.25 * (1.0 + eta) };
      const double det_[4] = { -0.25 * (1.0 - xi), -0.25 * (1.0 + xi), 0.25 * (1.0 + xi), 0.25 * (1.0 - xi) };
      double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
      for (int a = 0; a < 4; ++a) {
        J00 += dxi[a] * m[a].x;
        J01 += dxi[a] * m[a].y;
        J10 += det_[a] * m[a].x;
        J11 += det_[a] * m[a].y;
      }
      const double detJ = J00 * J11 - J01 * J10;
      const double i00 = J11 / detJ, i01 = -J01 / detJ, i10 = -J10 / detJ, i11 = J00 / detJ;
      double dx[4], dy[4];
      for (int a = 0; a < 4; ++a) {
        dx[a] = i00 * dxi[a] + i01 * det_[a];
        dy[a] = i10 * dxi[a] + i11 * det_[a];
      }
      const double w = detJ * 1.0 * 1.0;
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) K[a * 4 + b] += (dx[a] * dx[b]) * w + (dy[a] * dy[b]) * w;
    }
}

/* physical gradients and detJ of Quad4 / Hexa8 at one Gauss point (femutils/ArcaneFemFunctions.h computeGradientsAndJacobianQuad4 /
 * ...Hexa8: J = sum dN/dxi x, inverse Jacobian applied to the reference gradients) */
static double q1_gradients(int dim, const r3* m, double xi, double eta, double zeta, double* dx, double* dy, double* dz)
{
  static const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
  if (dim == 2) {
    double dxi[4], det_[4];
    for (int a = 0; a < 4; ++a) {
      dxi[a] = sx[a] * 0.25 * (1.0 + sy[a] * eta);
      det_[a] = sy[a] * 0.25 * (1.0 + sx[a] * xi);
    }
    double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
    for (int a = 0; a < 4; ++a) {
      J00 += dxi[a] * m[a].x; J01 += dxi[a] * m[a].y;
      J10 += det_[a] * m[a].x; J11 += det_[a] * m[a].y;
    }
    const double detJ = J00 * J11 - J01 * J10;
    const double i00 = J11 / detJ, i01 = -J01 / detJ, i10 = -J10 / detJ, i11 = J00 / detJ;
    for (int a = 0; a < 4; ++a) {
      dx[a] = i00 * dxi[a] + i01 * det_[a];
      dy[a] = i10 * dxi[a] + i11 * det_[a];
    }
    return detJ;
  }
  double dxi[8], det_[8], dze[8];
  for (int a = 0; a < 8; ++a) {
    dxi[a] = sx[a] * 0.125 * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
    det_[a] = sy[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sz[a] * zeta);
    dze[a] = sz[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta);
  }
  double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
  for (int a = 0; a < 8; ++a) {
    J[0][0] += dxi[a] * m[a].x; J[0][1] += dxi[a] * m[a].y; J[0][2] += dxi[a] * m[a].z;
    J[1][0] += det_[a] * m[a].x; J[1][1] += det_[a] * m[a].y; J[1][2] += det_[a] * m[a].z;
    J[2][0] += dze[a] * m[a].x; J[2][1] += dze[a] * m[a].y; J[2][2] += dze[a] * m[a].z;
  }
  const double detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                      J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  double inv[3][3];
  inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / detJ;
  inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / detJ;
  inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / detJ;
  inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / detJ;
  inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / detJ;
  inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / detJ;
  inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / detJ;
  inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / detJ;
  inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / detJ;
  for (int a = 0; a < 8; ++a) {
    dx[a] = inv[0][0] * dxi[a] + inv[0][1] * det_[a] + inv[0][2] * dze[a];
    dy[a] = inv[1][0] * dxi[a] + inv[1][1] * det_[a] + inv[1][2] * dze[a];
    dz[a] = inv[2][0] * dxi[a] + inv[2][1] * det_[a] + inv[2][2] * dze[a];
  }
  return detJ;
}

/* modules/elasticity/ElementMatrixHexQuad.h: computeElementMatrixQuad4Base / Hexa8Base summed over the 2x2 / 2x2x2 Gauss rule
 * (_computeElementMatrixQuad4 / Hexa8), interleaved DoFs, the three energy terms grouped as upstream */
static void ke_q1_elasticity(int dim, const r3* m, double lambda, double mu, double* K)
{
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  const int npc = dim == 2 ? 4 : 8, n = npc * dim;
  for (int i = 0; i < n * n; ++i) K[i] = 0.0;
  for (int ixi = 0; ixi < 2; ++ixi)
    for (int ieta = 0; ieta < 2; ++ieta)
      for (int izeta = 0; izeta < (dim == 3 ? 2 : 1); ++izeta) {
        double dxu[8], dyu[8], dzu[8];
        const double detJ = q1_gradients(dim, m, gp[ixi], gp[ieta], dim == 3 ? gp[izeta] : 0.0, dxu, dyu, dzu);
        const double w = detJ * 1.0 * 1.0 * 1.0;
        /* d{x,y,z}U{x,y,z}[dof]: gradient component of the shape function in the slot of the displacement component */
        double g[3][3][24]; /* [derivative][component][dof] */
        for (int d = 0; d < 3; ++d)
          for (int c = 0; c < 3; ++c)
            for (int q = 0; q < n; ++q) g[d][c][q] = 0.0;
        for (int a = 0; a < npc; ++a)
          for (int c = 0; c < dim; ++c) {
            g[0][c][dim * a + c] = dxu[a];
            g[1][c][dim * a + c] = dyu[a];
            if (dim == 3) g[2][c][dim * a + c] = dzu[a];
          }
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < n; ++j) {
            double nse, ce, se;
            if (dim == 2) {
              nse = (lambda + 2 * mu) * ((g[0][0][i] * g[0][0][j]) + (g[1][1][i] * g[1][1][j])) * w;
              ce = (lambda) * ((g[1][1][i] * g[0][0][j]) + (g[0][0][i] * g[1][1][j])) * w;
              se = (mu) * ((g[0][1][i] + g[1][0][i]) * (g[1][0][j] + g[0][1][j])) * w;
            }
            else {
              nse = (lambda + 2 * mu) * ((g[0][0][i] * g[0][0][j]) + (g[1][1][i] * g[1][1][j]) + (g[2][2][i] * g[2][2][j])) * w;
              ce = lambda * ((g[0][0][i] * (g[1][1][j] + g[2][2][j])) + (g[1][1][i] * (g[0][0][j] + g[2][2][j])) + (g[2][2][i] * (g[0][0][j] + g[1][1][j]))) * w;
              se = mu * (((g[1][0][i] + g[0][1][i]) * (g[1][0][j] + g[0][1][j])) + ((g[2][0][i] + g[0][2][i]) * (g[2][0][j] + g[0][2][j])) +
                         ((g[2][1][i] + g[1][2][i]) * (g[2][1][j] + g[1][2][j]))) * w;
            }
            K[i * n + j] += (nse + ce) + se;
          }
      }
}

static void ke_hexa8_poisson(const r3* m, double* K)
{
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  static const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
  for (int i = 0; i < 64; ++i) K[i] = 0.0;
  for (int ixi = 0; ixi < 2; ++ixi)
    for (int ieta = 0; ieta < 2; ++ieta)
      for (int izeta = 0; izeta < 2; ++izeta) {
        const double xi = gp[ixi], eta = gp[ieta], zeta = gp[izeta];
        double dxi[8], det_[8], dze[8];
        for (int a = 0; a < 8; ++a) { /* femutils/ShapeFunctions.h:317-346: +-0.125 (1 +- eta)(1 +- zeta) ... */
          dxi[a] = sx[a] * 0.125 * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
          det_[a] = sy[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sz[a] * zeta);
          dze[a] = sz[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta);
        }
        double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
        for (int a = 0; a < 8; ++a) {
          J[0][0] += dxi[a] * m[a].x; J[0][1] += dxi[a] * m[a].y; J[0][2] += dxi[a] * m[a].z;
          J[1][0] += det_[a] * m[a].x; J[1][1] += det_[a] * m[a].y; J[1][2] += det_[a] * m[a].z;
          J[2][0] += dze[a] * m[a].x; J[2][1] += dze[a] * m[a].y; J[2][2] += dze[a] * m[a].z;
        }
        const double detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                            J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        double inv[3][3]; /* adjugate / det (Arcane math::inverseMatrix) */
        inv[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / detJ;
        inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / detJ;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / detJ;
        inv[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / detJ;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / detJ;
        inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / detJ;
        inv[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / detJ;
        inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / detJ;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / detJ;
        double dx[8], dy[8], dz[8];
        for (int a = 0; a < 8; ++a) {
          dx[a] = inv[0][0] * dxi[a] + inv[0][1] * det_[a] + inv[0][2] * dze[a];
          dy[a] = inv[1][0] * dxi[a] + inv[1][1] * det_[a] + inv[1][2] * dze[a];
          dz[a] = inv[2][0] * dxi[a] + inv[2][1] * det_[a] + inv[2][2] * dze[a];
        }
        const double w = detJ * 1.0 * 1.0;
        for (int a = 0; a < 8; ++a)
          for (int b = 0; b < 8; ++b) K[a * 8 + b] += (dx[a] * dx[b]) * w + (dy[a] * dy[b]) * w + (dz[a] * dz[b]) * w;
      }
}


/* ------------------------------------------------------------------------- */
/* Element-matrix dispatcher: K is (npc*b) x (npc*b), row-major.              */
/* params: ELASTICITY -> {lambda, mu}                                         */
/* ------------------------------------------------------------------------- */
static int op_block_size(int op, int dim) { return (op == ORC_OP_POISSON || op == ORC_OP_DIFFUSION_REACTION) ? 1 : ((op == ORC_OP_ELASTICITY || op == ORC_OP_ELASTODYNAMICS) ? dim : 2); }
static double g_stiffness_scale = 1.0; /* per-cell coefficient of the stiffness part of ORC_OP_DIFFUSION_REACTION (set by the assembly loops) */

static int element_matrix(int npc, int dim, int op, int form, const double* params, const double* coords, const int32_t* cn, double* K)
{
  r3 m[10];
  for (int i = 0; i < npc; ++i) m[i] = r3_load(coords, cn[i]);
  if (op == ORC_OP_POISSON) {
    if (npc == 3 && dim == 2) {
      if (form == ORC_FORM_COMPACT) ke_tri3_poisson_compact(m[0], m[1], m[2], K);
      else if (form == ORC_FORM_HOST) ke_tri3_poisson_host(m[0], m[1], m[2], K);
      else if (form == ORC_FORM_BSR) ke_tri3_poisson_bsr(m[0], m[1], m[2], K);
      else {
        double b[6];
        double area = bmat_tri3(m[0], m[1], m[2], b);
        for (int a = 0; a < 3; ++a)
          for (int i = 0; i < 3; ++i) {
            /* modules/testlab/NodeWiseCsrBiliAssembly.cc:202,211 */
            double x = b[a * 2] * b[i * 2] + b[a * 2 + 1] * b[i * 2 + 1];
            K[a * 3 + i] = x * area;
          }
      }
      return 0;
    }
    if (npc == 4 && dim == 3) {
      if (form == ORC_FORM_COMPACT) ke_tet4_poisson_compact(m[0], m[1], m[2], m[3], K);
      else if (form == ORC_FORM_HOST) ke_tet4_poisson_host(m[0], m[1], m[2], m[3], K);
      else if (form == ORC_FORM_BSR) ke_tet4_poisson_bsr(m[0], m[1], m[2], m[3], K);
      else {
        double b[12];
        double volume = bmat_tet4(m[0], m[1], m[2], m[3], b);
        for (int a = 0; a < 4; ++a)
          for (int i = 0; i < 4; ++i) {
            /* modules/testlab/NodeWiseCsrBiliAssembly.cc:278,287 */
            double x = b[a * 3] * b[i * 3] + b[a * 3 + 1] * b[i * 3 + 1] + b[a * 3 + 2] * b[i * 3 + 2];
            K[a * 4 + i] = x * volume;
          }
      }
      return 0;
    }
    if (npc == 4 && dim == 2) { ke_quad4_poisson(m, K); return 0; }
    if (npc == 8 && dim == 3) { ke_hexa8_poisson(m, K); return 0; }
    if (npc == 6 && dim == 2) { ke_tri6_poisson(m, K); return 0; }
    if (npc == 10 && dim == 3) { ke_tet10_poisson(m, K); return 0; }
    return -1;
  }
  if (op == ORC_OP_DIFFUSION_REACTION) {
    /* alpha * stiffness + beta * consistent mass: the acoustics module (modules/acoustics/ElementMatrix.h:14,29:
     * -area (dxU^dxU) - area (dyU^dyU) + kc2 area (1/12) massMatrix(U,U); ElementMatrixHexQuad.h: (dxU^dxU) w + (dyU^dyU) w + (N^N) kc2 w
     * per Gauss point) and the heat module (modules/heat/ElementMatrix.h: lambda (area (dxU^dxU) + area (dyU^dyU)) + (1/12) massMatrix(U,U) area / dt);
     * massMatrix(U,U) with U = 1: ones with a doubled diagonal (femutils/FemUtils.h:583-597).  A per-cell coefficient multiplies alpha. */
    const double alpha = params[0] * g_stiffness_scale, beta = params[1];
    const int n = npc;
    if (npc == dim + 1) {
      double S[16];
      double meas;
      if (dim == 2) { ke_tri3_poisson_bsr(m[0], m[1], m[2], S); meas = area_tri3_unsigned(m[0], m[1], m[2]); }
      else { ke_tet4_poisson_bsr(m[0], m[1], m[2], m[3], S); meas = volume_tet4(m[0], m[1], m[2], m[3]); }
      const double mc = dim == 2 ? (1 / 12.) : (1 / 20.);
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b) K[a * n + b] = alpha * S[a * n + b] + beta * meas * mc * (a == b ? 2.0 : 1.0);
      return 0;
    }
    if (npc == (1 << dim)) {
      const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
      static const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
      for (int i = 0; i < n * n; ++i) K[i] = 0.0;
      for (int ixi = 0; ixi < 2; ++ixi)
        for (int ieta = 0; ieta < 2; ++ieta)
          for (int izeta = 0; izeta < (dim == 3 ? 2 : 1); ++izeta) {
            const double xi = gp[ixi], eta = gp[ieta], zeta = dim == 3 ? gp[izeta] : 0.0;
            double dx[8], dy[8], dz[8], N[8];
            const double w = q1_gradients(dim, m, xi, eta, zeta, dx, dy, dz);
            for (int a = 0; a < n; ++a)
              N[a] = dim == 2 ? 0.25 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) : 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
            for (int a = 0; a < n; ++a)
              for (int b = 0; b < n; ++b) {
                double st = (dx[a] * dx[b]) * w + (dy[a] * dy[b]) * w;
                if (dim == 3) st += (dz[a] * dz[b]) * w;
                K[a * n + b] += alpha * st + (N[a] * N[b]) * beta * w;
              }
          }
      return 0;
    }
    return -1;
  }
  if (op == ORC_OP_ELASTICITY) {
    if (npc == 3 && dim == 2) { ke_tri3_elasticity(m[0], m[1], m[2], params[0], params[1], K); return 0; }
    if (npc == 4 && dim == 3) { ke_tet4_elasticity(m[0], m[1], m[2], m[3], params[0], params[1], K); return 0; }
    if ((npc == 4 && dim == 2) || (npc == 8 && dim == 3)) { ke_q1_elasticity(dim, m, params[0], params[1], K); return 0; }
    return -1;
  }
  if (op == ORC_OP_BILAPLACIAN) {
    if (npc == 3 && dim == 2) { ke_tri3_bilaplacian(m[0], m[1], m[2], K); return 0; }
    return -1;
  }
  if (op == ORC_OP_ELASTODYNAMICS) {
    /* Newmark-beta / generalised-alpha matrix of the elastodynamics module, params = {c0, c1, c2}
     * (modules/elastodynamics/ElementMatrix.h:41-60 Tria3, :150-196 Tetra4; coefficients FemModule.cc:203-205):
     *   (c0/12) (massMatrix(Ux,Ux) + massMatrix(Uy,Uy)) area + c1 (cross terms) area + (2 c2 + c1) (normal terms) area + c2 (shear terms) area
     * = the elasticity element matrix with lambda = c1, mu = c2, plus c0 times the consistent mass on every component
     * ((1 + delta_ab)/12 area, /20 volume: massMatrix femutils/FemUtils.h:583-597). */
    const double c0 = params[0], c1 = params[1], c2 = params[2];
    if (npc == 3 && dim == 2) {
      double M[36];
      const double area = area_tri3_unsigned(m[0], m[1], m[2]);
      const double Ux[6] = { 1., 0., 1., 0., 1., 0. }, Uy[6] = { 0., 1., 0., 1., 0., 1. };
      ke_tri3_elasticity(m[0], m[1], m[2], c1, c2, K);
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          const double dbl = i == j ? 2. : 1.;
          M[i * 6 + j] = (c0 / 12.) * ((Ux[i] * Ux[j]) * dbl + (Uy[i] * Uy[j]) * dbl) * area;
        }
      for (int i = 0; i < 36; ++i) K[i] = M[i] + K[i];
      return 0;
    }
    if (npc == 4 && dim == 3) {
      double M[144];
      const double volume = volume_tet4(m[0], m[1], m[2], m[3]);
      ke_tet4_elasticity(m[0], m[1], m[2], m[3], c1, c2, K);
      for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j) {
          const double dbl = i == j ? 2. : 1.;
          const double same = (i % 3) == (j % 3) ? 1. : 0.; /* Ux^Ux + Uy^Uy + Uz^Uz */
          M[i * 12 + j] = (c0 / 20.) * (same * dbl) * volume;
        }
      for (int i = 0; i < 144; ++i) K[i] = M[i] + K[i];
      return 0;
    }
    if ((npc == 4 && dim == 2) || (npc == 8 && dim == 3)) {
      /* modules/elastodynamics/ElementMatrixHexQuad.h: per Gauss point c0 ((Nx^Nx) + (Ny^Ny) [+ (Nz^Nz)]) weight + the elasticity terms
       * with (c1, c2); the elasticity part is ke_q1_elasticity, the mass part is summed here over the same 2x2(x2) rule */
      const int n = npc, nd = n * dim;
      const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
      static const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
      double Mq[576];
      for (int i = 0; i < nd * nd; ++i) Mq[i] = 0.0;
      for (int ixi = 0; ixi < 2; ++ixi)
        for (int ieta = 0; ieta < 2; ++ieta)
          for (int izeta = 0; izeta < (dim == 3 ? 2 : 1); ++izeta) {
            const double xi = gp[ixi], eta = gp[ieta], zeta = dim == 3 ? gp[izeta] : 0.0;
            double dx[8], dy[8], dz[8], N[8];
            const double w = q1_gradients(dim, m, xi, eta, zeta, dx, dy, dz);
            for (int a = 0; a < n; ++a)
              N[a] = dim == 2 ? 0.25 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) : 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
            for (int a = 0; a < n; ++a)
              for (int b = 0; b < n; ++b)
                for (int k = 0; k < dim; ++k) Mq[(a * dim + k) * nd + (b * dim + k)] += (c0 * (N[a] * N[b])) * w;
          }
      ke_q1_elasticity(dim, m, c1, c2, K);
      for (int i = 0; i < nd * nd; ++i) K[i] = Mq[i] + K[i];
      return 0;
    }
    return -1;
  }
  return -1;
}

/* exported single-element entry (tests compare formulations) */
ORC_API int orc_element_matrix(int npc, int dim, int op, int form, const double* params, const double* coords, const int32_t* cell_nodes, double* K_out)
{
  return element_matrix(npc, dim, op, form, params, coords, cell_nodes, K_out);
}

/* Lamé parameters: modules/elasticity/FemModule.cc:171-172 */
ORC_API void orc_lame(double E, double nu, double* lambda, double* mu)
{
  *mu = (E / (2 * (1 + nu)));
  *lambda = E * nu / ((1 + nu) * (1 - 2 * nu));
}

/* ========================================================================= */
/* Sparsity pattern                                                           */
/* Sort-based construction from cells only:                                   */
/*   modules/testlab/CsrGpuBiliAssembly.cc:23-207 (pack/sort/unique/degree/   */
/*   scan/columns), twin femutils/BSRFormat.cc:799-1006.  P1 simplices use    */
/*   their 3/6 edges; higher-order cells use all node pairs of the cell        */
/*   (femutils/BSRFormat.cc:290-310, 848-862).  nnz = nbNode + 2*nbEdge         */
/*   (CsrGpuBiliAssembly.cc:193).  Output columns canonical (ascending).        */
/* ========================================================================= */
static int cmp_u64(const void* a, const void* b)
{
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}
static int cmp_i32(const void* a, const void* b)
{
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* Returns nnz; fills rows[nb_node+1]. If columns==NULL only counts. */
ORC_API int64_t orc_build_pattern(int npc, int32_t nb_node, int64_t nb_cell, const int32_t* conn, int32_t* rows, int32_t* columns)
{
  int pairs = npc * (npc - 1) / 2;
  int64_t nkeys = nb_cell * pairs;
  uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(nkeys > 0 ? nkeys : 1));
  int64_t k = 0;
  for (int64_t c = 0; c < nb_cell; ++c) {
    const int32_t* cn = conn + c * npc;
    for (int i = 0; i < npc; ++i)
      for (int j = i + 1; j < npc; ++j) {
        int32_t n0 = cn[i], n1 = cn[j];
        int32_t mn = n0 > n1 ? n1 : n0, mx = n0 > n1 ? n0 : n1;
        keys[k++] = ((uint64_t)(uint32_t)mn << 32) | (uint64_t)(uint32_t)mx; /* pack(): CsrGpuBiliAssembly.cc:23-28 */
      }
  }
  qsort(keys, (size_t)nkeys, sizeof(uint64_t), cmp_u64); /* GenericSorter: CsrGpuBiliAssembly.cc:88-90 */
  int32_t* deg = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nb_node + 1));
  for (int32_t n = 0; n < nb_node; ++n) deg[n] = 1; /* neighbors.fill(1): :121 (the diagonal) */
  for (int64_t i = 0; i < nkeys; ++i)
    if (i == nkeys - 1 || keys[i] != keys[i + 1]) { /* :105 */
      deg[(int32_t)(keys[i] >> 32)]++;
      deg[(int32_t)(keys[i] & 0xFFFFFFFFu)]++;
    }
  int64_t acc = 0;
  for (int32_t n = 0; n < nb_node; ++n) { rows[n] = (int32_t)acc; acc += deg[n]; } /* exclusiveSum :124-126 */
  rows[nb_node] = (int32_t)acc;
  if (columns) {
    int32_t* off = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nb_node + 1));
    for (int32_t n = 0; n < nb_node; ++n) { columns[rows[n]] = n; off[n] = 1; } /* diag seed :150-155, offsets.fill(1) :162 */
    for (int64_t i = 0; i < nkeys; ++i)
      if (i == nkeys - 1 || keys[i] != keys[i + 1]) { /* :174-179 */
        int32_t n0 = (int32_t)(keys[i] >> 32), n1 = (int32_t)(keys[i] & 0xFFFFFFFFu);
        columns[rows[n0] + off[n0]++] = n1;
        columns[rows[n1] + off[n1]++] = n0;
      }
    /* canonical form: ascending columns inside each row */
    for (int32_t n = 0; n < nb_node; ++n) qsort(columns + rows[n], (size_t)(rows[n + 1] - rows[n]), sizeof(int32_t), cmp_i32);
    free(off);
  }
  free(deg);
  free(keys);
  return acc;
}

/* ------------------------------------------------------------------------- */
/* Slot lookup.                                                               */
/* Block (row_node,col_node) position: linear scan of the row, as             */
/*   femutils/CsrFormatMatrix.h:66-87 (indexValue) and BSRFormat.h:288-302.    */
/* Value index inside the block, both layouts:                                */
/*   per-block: begin*b*b + i*b + j           (femutils/BSRFormat.h:292-296)   */
/*   per-row  : rows[r]*b*b + b*(x + i*nz_r) + j  (femutils/BSRFormat.h:356)    */
/* ------------------------------------------------------------------------- */
static inline int32_t find_block(const int32_t* rows, const int32_t* cols, int32_t r, int32_t c)
{
  for (int32_t p = rows[r]; p < rows[r + 1]; ++p)
    if (cols[p] == c) return p;
  return -1;
}
static inline int64_t value_index(const int32_t* rows, int32_t r, int32_t p, int b, int layout, int i, int j)
{
  if (layout == ORC_LAYOUT_PER_BLOCK) return (int64_t)p * b * b + i * b + j;
  int32_t nz = rows[r + 1] - rows[r];
  int32_t x = p - rows[r];
  return (int64_t)rows[r] * b * b + (int64_t)b * (x + (int64_t)i * nz) + j;
}

ORC_API int64_t orc_value_index(const int32_t* rows, const int32_t* cols, int b, int layout, int32_t dof_row, int32_t dof_col)
{
  /* femutils/BSRFormat.cc:79-106 (BSRMatrix::findValueIndex) */
  int32_t br = dof_row / b, bc = dof_col / b;
  int32_t p = find_block(rows, cols, br, bc);
  if (p < 0) return -1;
  return value_index(rows, br, p, b, layout, dof_row % b, dof_col % b);
}

/* ========================================================================= */
/* Bilinear assembly                                                          */
/* ========================================================================= */

/*
 * Cell-wise scatter: for each cell (ascending id), K_e then
 *   A[dof(n1),dof(n2)] += K_e[i1,i2] iff n1.isOwn()
 * modules/testlab/CsrGpuBiliAssembly.cc:339-372 (CSR_GPU), CsrBiliAssembly.cc:148-182
 * (CSR host: skip_zero=1 restates CsrFormat::matrixAddValue's `value == 0.0` early
 * return, femutils/CsrFormatMatrix.h:64), femutils/BSRFormat.h:257-370 (BSR),
 * modules/elasticity/FemModule.cc:311-341 (host DOK loop, same accumulation order).
 * The device back-ends add with atomics in arbitrary order; this sequential order
 * is one valid order, so device results are compared at 1e-12 (tests).
 */
/* Per-cell coefficient of the stiffness operator (m_cell_lambda of the fourier / heat modules: modules/fourier/ElementMatrix.h:11,
 * 22,45,57 `area * lambda * (dxU ^ dxU) + ...`; electrostatics epsilon): K_e of cell c is multiplied by coef[c].  Set before an
 * orc_assemble_* call, NULL switches it off (test infrastructure: a process-wide pointer). */
static const double* g_cell_coef = 0;
ORC_API void orc_set_cell_coefficient(const double* coef) { g_cell_coef = coef; }
static void scale_ke(double* K, int n, int64_t cell, int op)
{
  if (!g_cell_coef || op != ORC_OP_POISSON) return; /* (ORC_OP_DIFFUSION_REACTION took it inside, on the stiffness part) */
  for (int i = 0; i < n * n; ++i) K[i] *= g_cell_coef[cell];
}

ORC_API int orc_assemble_cellwise(int npc, int dim, int op, int form, const double* params,
                                  int32_t nb_node, int64_t nb_cell, const double* coords, const int32_t* conn, const uint8_t* is_own,
                                  const int32_t* rows, const int32_t* cols, int layout, int skip_zero, double* values)
{
  (void)nb_node;
  int b = op_block_size(op, dim);
  int n = npc * b;
  double K[576];
  for (int64_t c = 0; c < nb_cell; ++c) {
    const int32_t* cn = conn + c * npc;
    g_stiffness_scale = g_cell_coef ? g_cell_coef[c] : 1.0;
    if (element_matrix(npc, dim, op, form, params, coords, cn, K)) return -1;
    scale_ke(K, n, c, op);
    for (int a1 = 0; a1 < npc; ++a1) {
      int32_t r = cn[a1];
      if (is_own && !is_own[r]) continue;
      for (int a2 = 0; a2 < npc; ++a2) {
        int32_t p = find_block(rows, cols, r, cn[a2]);
        if (p < 0) return -2;
        for (int i = 0; i < b; ++i)
          for (int j = 0; j < b; ++j) {
            double v = K[(a1 * b + i) * n + (a2 * b + j)];
            if (skip_zero && v == 0.0) continue;
            values[value_index(rows, r, p, b, layout, i, j)] += v;
          }
      }
    }
  }
  return 0;
}

/* node -> incident cells, ascending cell id (Arcane nodeCell view; its order is
 * mesh-reader dependent, ascending is the canonical order used here) */
static void build_node_cells(int npc, int32_t nb_node, int64_t nb_cell, const int32_t* conn, int64_t** ptr_out, int32_t** list_out)
{
  int64_t* ptr = (int64_t*)calloc((size_t)nb_node + 1, sizeof(int64_t));
  for (int64_t c = 0; c < nb_cell; ++c)
    for (int i = 0; i < npc; ++i) ptr[conn[c * npc + i] + 1]++;
  for (int32_t n = 0; n < nb_node; ++n) ptr[n + 1] += ptr[n];
  int32_t* list = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ptr[nb_node] > 0 ? ptr[nb_node] : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb_node + 1));
  memcpy(fill, ptr, sizeof(int64_t) * (size_t)(nb_node + 1));
  for (int64_t c = 0; c < nb_cell; ++c)
    for (int i = 0; i < npc; ++i) list[fill[conn[c * npc + i]]++] = (int32_t)c;
  free(fill);
  *ptr_out = ptr;
  *list_out = list;
}

/*
 * Node-wise (atomic-free) assembly: for each owned node, for each incident cell,
 * the element row of that node is recomputed and added with plain `+=`.
 * modules/testlab/NodeWiseCsrBiliAssembly.cc:259-296 (form = NODEWISE, b-matrix dot
 * products), femutils/BSRFormat.h:406-537 with the AF lambdas
 * modules/testlab/FemModule.cc:278-315 and modules/elasticity/ElementMatrix.h:87-119,
 * 216-301 (form = BSR: the row of the full-matrix formula; the AF lambdas evaluate
 * the same products, association differs only in where `volume` multiplies:
 * `volume*dx[a]*dx[j]` vs `volume*(dx[a]*dx[j])`, i.e. <= 1 ulp per term).
 */
ORC_API int orc_assemble_nodewise(int npc, int dim, int op, int form, const double* params,
                                  int32_t nb_node, int64_t nb_cell, const double* coords, const int32_t* conn, const uint8_t* is_own,
                                  const int32_t* rows, const int32_t* cols, int layout, double* values)
{
  int b = op_block_size(op, dim);
  int n = npc * b;
  int64_t* ptr;
  int32_t* list;
  build_node_cells(npc, nb_node, nb_cell, conn, &ptr, &list);
  double K[576];
  int rc = 0;
  for (int32_t r = 0; r < nb_node && !rc; ++r) {
    if (is_own && !is_own[r]) continue;
    for (int64_t q = ptr[r]; q < ptr[r + 1]; ++q) {
      const int32_t* cn = conn + (int64_t)list[q] * npc;
      int a1 = -1;
      for (int i = 0; i < npc; ++i) if (cn[i] == r) { a1 = i; break; }
      if (a1 < 0) continue;
      g_stiffness_scale = g_cell_coef ? g_cell_coef[list[q]] : 1.0;
      if (element_matrix(npc, dim, op, form, params, coords, cn, K)) { rc = -1; break; }
      scale_ke(K, n, list[q], op);
      for (int a2 = 0; a2 < npc; ++a2) {
        int32_t p = find_block(rows, cols, r, cn[a2]);
        if (p < 0) { rc = -2; break; }
        for (int i = 0; i < b; ++i)
          for (int j = 0; j < b; ++j)
            values[value_index(rows, r, p, b, layout, i, j)] += K[(a1 * b + i) * n + (a2 * b + j)];
      }
    }
  }
  free(ptr);
  free(list);
  return rc;
}

/* ========================================================================= */
/* Hand-off views                                                             */
/* ========================================================================= */

/* femutils/CsrFormatMatrix.cc:161-184 (_translateCSRToCOO): expand row pointer */
ORC_API void orc_csr_to_coo_rows(int32_t nb_row, const int32_t* rows, int32_t* coo_rows)
{
  for (int32_t r = 0; r < nb_row; ++r)
    for (int32_t p = rows[r]; p < rows[r + 1]; ++p) coo_rows[p] = r;
}

/* femutils/BSRFormat.cc:110-172 (BSRMatrix::toCsr, b>1): rows[nbRow*b+1] with
 * sentinel, col = block_col*b + k, rows_nb_column = nz_r*b; values shared
 * (per-row layout is already CSR order). */
ORC_API void orc_bsr_to_csr(int32_t nb_block_row, int b, const int32_t* rows, const int32_t* cols,
                            int32_t* csr_rows, int32_t* csr_cols, int32_t* csr_rows_nb_column)
{
  csr_rows[0] = 0;
  int64_t off = 1;
  for (int32_t i = 0; i < nb_block_row; ++i)
    for (int j = 0; j < b; ++j) {
      csr_rows[off] = csr_rows[off - 1] + (rows[i + 1] - rows[i]) * b;
      off++;
    }
  off = 0;
  for (int32_t i = 0; i < nb_block_row; ++i)
    for (int j = 0; j < b; ++j)
      for (int32_t p = rows[i]; p < rows[i + 1]; ++p)
        for (int k = 0; k < b; ++k) csr_cols[off++] = cols[p] * b + k;
  off = 0;
  for (int32_t i = 0; i < nb_block_row; ++i)
    for (int j = 0; j < b; ++j) csr_rows_nb_column[off++] = (rows[i + 1] - rows[i]) * b;
}

/* ========================================================================= */
/* RHS: constant source                                                       */
/* ========================================================================= */

/* Cell-wise constant source, b components per node, accumulation in cell order:
 *   rhs[dof(n,k)] += f[k]*meas/npc   for isOwn(n) (and !dirichlet(n) when given)
 * testlab (b=1): modules/testlab/FemModule.cc:836-868 (host), :1358-1532 (device);
 *   meas = SIGNED tri area (:1762) / abs tet volume (:1748), skips Dirichlet nodes.
 * elasticity body force: modules/elasticity/BodyForce.h:93-104 (tri, area/3) and
 *   the Tetra4 branch (volume/4); bilaplacian: modules/bilaplacian/FemModule.cc:157-172
 *   (component 0 only: pass f={f,0}); meas = UNSIGNED area (cross norm). */
/* Quad4 / Hexa8: the source term is integrated with the 2x2 / 2x2x2 Gauss rule, rhs_i += N_i(gp) * qdot * w * detJ
 * (femutils/ArcaneFemFunctions.cc:222-290 applyConstantSourceToRhsQuad4, :437-483 applyConstantSourceToRhsHexa8;
 * shape functions femutils/ShapeFunctions.h, nodes counter-clockwise, bottom face then top face). */
static void q1_source_weights(int dim, const r3* m, double* w /* [npc]: integral of N_i over the cell */)
{
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  static const double sx[8] = { -1, 1, 1, -1, -1, 1, 1, -1 }, sy[8] = { -1, -1, 1, 1, -1, -1, 1, 1 }, sz[8] = { -1, -1, -1, -1, 1, 1, 1, 1 };
  const int npc = dim == 2 ? 4 : 8;
  for (int i = 0; i < npc; ++i) w[i] = 0.0;
  if (dim == 2) {
    for (int ixi = 0; ixi < 2; ++ixi)
      for (int ieta = 0; ieta < 2; ++ieta) {
        const double xi = gp[ixi], eta = gp[ieta];
        double N[4], dxi[4], det_[4];
        for (int a = 0; a < 4; ++a) {
          N[a] = 0.25 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta);
          dxi[a] = sx[a] * 0.25 * (1.0 + sy[a] * eta);
          det_[a] = sy[a] * 0.25 * (1.0 + sx[a] * xi);
        }
        double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
        for (int a = 0; a < 4; ++a) {
          J00 += dxi[a] * m[a].x;
          J01 += dxi[a] * m[a].y;
          J10 += det_[a] * m[a].x;
          J11 += det_[a] * m[a].y;
        }
        const double iw = 1.0 * (J00 * J11 - J01 * J10);
        for (int a = 0; a < 4; ++a) w[a] += N[a] * iw;
      }
    return;
  }
  for (int ixi = 0; ixi < 2; ++ixi)
    for (int ieta = 0; ieta < 2; ++ieta)
      for (int izeta = 0; izeta < 2; ++izeta) {
        const double xi = gp[ixi], eta = gp[ieta], zeta = gp[izeta];
        double N[8], dxi[8], det_[8], dze[8];
        for (int a = 0; a < 8; ++a) {
          N[a] = 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
          dxi[a] = sx[a] * 0.125 * (1.0 + sy[a] * eta) * (1.0 + sz[a] * zeta);
          det_[a] = sy[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sz[a] * zeta);
          dze[a] = sz[a] * 0.125 * (1.0 + sx[a] * xi) * (1.0 + sy[a] * eta);
        }
        double J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
        for (int a = 0; a < 8; ++a) {
          J[0][0] += dxi[a] * m[a].x; J[0][1] += dxi[a] * m[a].y; J[0][2] += dxi[a] * m[a].z;
          J[1][0] += det_[a] * m[a].x; J[1][1] += det_[a] * m[a].y; J[1][2] += det_[a] * m[a].z;
          J[2][0] += dze[a] * m[a].x; J[2][1] += dze[a] * m[a].y; J[2][2] += dze[a] * m[a].z;
        }
        const double detJ = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                            J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
        for (int a = 0; a < 8; ++a) w[a] += N[a] * detJ;
      }
}

ORC_API void orc_rhs_source_cellwise(int npc, int dim, int b, int signed_area, int32_t nb_node, int64_t nb_cell, const double* coords, const int32_t* conn,
                                     const uint8_t* is_own, const uint8_t* is_dirichlet, const double* f, double* rhs)
{
  (void)nb_node;
  if (npc == (dim == 2 ? 4 : 8)) { /* Quad4 / Hexa8 */
    for (int64_t c = 0; c < nb_cell; ++c) {
      const int32_t* cn = conn + c * npc;
      r3 m[8];
      double w[8];
      for (int a = 0; a < npc; ++a) m[a] = r3_load(coords, cn[a]);
      q1_source_weights(dim, m, w);
      for (int i = 0; i < npc; ++i) {
        const int32_t nd = cn[i];
        if ((is_dirichlet && is_dirichlet[nd]) || (is_own && !is_own[nd])) continue;
        for (int k = 0; k < b; ++k)
          if (f[k] != 0.0) rhs[(int64_t)nd * b + k] += w[i] * f[k];
      }
    }
    return;
  }
  for (int64_t c = 0; c < nb_cell; ++c) {
    const int32_t* cn = conn + c * npc;
    double meas;
    if (dim == 2) {
      r3 m0 = r3_load(coords, cn[0]), m1 = r3_load(coords, cn[1]), m2 = r3_load(coords, cn[2]);
      meas = signed_area ? 0.5 * ((m1.x - m0.x) * (m2.y - m0.y) - (m2.x - m0.x) * (m1.y - m0.y)) : area_tri3_unsigned(m0, m1, m2);
    }
    else
      meas = volume_tet4(r3_load(coords, cn[0]), r3_load(coords, cn[1]), r3_load(coords, cn[2]), r3_load(coords, cn[3]));
    for (int i = 0; i < npc; ++i) {
      int32_t nd = cn[i];
      if ((is_dirichlet && is_dirichlet[nd]) || (is_own && !is_own[nd])) continue;
      for (int k = 0; k < b; ++k)
        if (f[k] != 0.0) rhs[(int64_t)nd * b + k] += f[k] * meas / npc;
    }
  }
}

/* BC-service / production modules (node-wise, all owned nodes, b components):
 * femutils/ArcaneFemFunctionsGpu.h:675-708 (sum += qdot*domain/nbNode(cell), then
 * rhs = sum), modules/elasticity/BodyForce.h:93-104 (f[k]*area/3 per component).
 * meas = unsigned tri area (cross norm) / abs tet volume. Node-cell order ascending. */
ORC_API void orc_rhs_source_nodewise(int npc, int dim, int b, int32_t nb_node, int64_t nb_cell, const double* coords, const int32_t* conn,
                                     const uint8_t* is_own, const double* f, double* rhs)
{
  int64_t* ptr;
  int32_t* list;
  build_node_cells(npc, nb_node, nb_cell, conn, &ptr, &list);
  for (int32_t r = 0; r < nb_node; ++r) {
    if (is_own && !is_own[r]) continue;
    for (int k = 0; k < b; ++k) {
      double sum = 0.0;
      for (int64_t q = ptr[r]; q < ptr[r + 1]; ++q) {
        const int32_t* cn = conn + (int64_t)list[q] * npc;
        double meas = (dim == 2) ? area_tri3_unsigned(r3_load(coords, cn[0]), r3_load(coords, cn[1]), r3_load(coords, cn[2]))
                                 : volume_tet4(r3_load(coords, cn[0]), r3_load(coords, cn[1]), r3_load(coords, cn[2]), r3_load(coords, cn[3]));
        sum += f[k] * meas / npc;
      }
      rhs[(int64_t)r * b + k] = sum;
    }
  }
  free(ptr);
  free(list);
}

/* Boundary integrals of the RHS on P1 faces (edges of a Tri3 mesh, triangles of a Tet4 mesh).
 * faces[nb_face][dim]: the face's nodes in Arcane's order with the swap of the reference's normal helpers
 * already applied when the face is not "subdomain boundary outside" (computeNormalFace / computeNormalTriangle,
 * femutils/ArcaneFemFunctionsGpu.h:159-214; _computeEdgeNormal2Gpu, modules/testlab/FemModule.cc:1825-1841).
 * measure: edge length in the x-y plane (FemModule.cc:1782-1787, ArcaneFemFunctionsGpu.h:139-147) / triangle area
 *          |(n1-n0) x (n2-n0)| / 2 (ArcaneFemFunctionsGpu.h:92-102).
 * kind 0 (Neumann flux on DoF 0): nb_value == 1: rhs += value*measure/nn
 *          (FemModule.cc:1574-1581, ArcaneFemFunctionsGpu.cc:701-717, 1105-1120);
 *          nb_value == dim: rhs += (N.q)*measure/nn with the unit normal N of the helpers above
 *          (FemModule.cc:1611-1621, ArcaneFemFunctionsGpu.cc:718-737, 1121-1141).
 * kind 1 (traction, b DoFs per node): rhs[dof(node,k)] += t[k]*measure/nn
 *          (femutils/ArcaneFemFunctions.h:2188-2220, 2854-2885).
 * Gates: owned nodes only; testlab additionally skips Dirichlet nodes (FemModule.cc:1577). */
/* Quad4 faces of a Hexa8 mesh (femutils/ArcaneFemFunctions.h:1843-1953 applyNeumannToRhsHexa8): 2x2 Gauss rule on the
 * bilinear patch, tangents t1 = dr/dxi, t2 = dr/deta, detJ = |t1 x t2|, unit normal (t1 x t2)/detJ at every Gauss point;
 * rhs_j += value * N_j * detJ (scalar) or (normal . q) * N_j * detJ (vector).  The faces come oriented (outward t1 x t2). */
ORC_API void orc_rhs_neumann_quad4(int b, int kind, int nb_value, int64_t nb_face, const double* coords, const int32_t* faces, const double* values, const uint8_t* is_own,
                                   const uint8_t* is_dirichlet, double* rhs)
{ /* kind 1: traction, rhs[dof(j,k)] += t[k] * N_j * detJ (femutils/ArcaneFemFunctions.h:2222-2315 applyTractionToRhsHexa8) */
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  static const double sx[4] = { -1, 1, 1, -1 }, sy[4] = { -1, -1, 1, 1 };
  for (int64_t f = 0; f < nb_face; ++f) {
    const int32_t* fn = faces + f * 4;
    r3 c[4];
    for (int i = 0; i < 4; ++i) c[i] = r3_load(coords, fn[i]);
    for (int ixi = 0; ixi < 2; ++ixi)
      for (int ieta = 0; ieta < 2; ++ieta) {
        const double xi = gp[ixi], eta = gp[ieta];
        double N[4];
        r3 t1 = { 0, 0, 0 }, t2 = { 0, 0, 0 };
        for (int i = 0; i < 4; ++i) {
          N[i] = 0.25 * (1.0 + sx[i] * xi) * (1.0 + sy[i] * eta);
          const double dxi = sx[i] * 0.25 * (1.0 + sy[i] * eta), det_ = sy[i] * 0.25 * (1.0 + sx[i] * xi);
          t1.x += dxi * c[i].x; t1.y += dxi * c[i].y; t1.z += dxi * c[i].z;
          t2.x += det_ * c[i].x; t2.y += det_ * c[i].y; t2.z += det_ * c[i].z;
        }
        r3 nr = r3_cross(t1, t2);
        const double detJ = sqrt(nr.x * nr.x + nr.y * nr.y + nr.z * nr.z);
        nr.x /= detJ; nr.y /= detJ; nr.z /= detJ;
        const double iw = 1.0 * 1.0 * detJ;
        for (int j = 0; j < 4; ++j) {
          const int32_t nd = fn[j];
          if ((is_dirichlet && is_dirichlet[nd]) || (is_own && !is_own[nd])) continue;
          if (kind == 1) {
            for (int k = 0; k < b; ++k) rhs[(int64_t)nd * b + k] += values[k] * N[j] * iw;
            continue;
          }
          const double v = nb_value == 1 ? values[0] * N[j] * iw : (nr.x * values[0] + nr.y * values[1] + nr.z * values[2]) * N[j] * iw;
          rhs[(int64_t)nd * b] += v;
        }
      }
  }
}

ORC_API void orc_rhs_neumann(int dim, int b, int kind, int nb_value, int64_t nb_face, const double* coords, const int32_t* faces, const double* values,
                             const uint8_t* is_own, const uint8_t* is_dirichlet, double* rhs)
{
  const int nn = dim; /* 2 nodes per edge, 3 per triangle */
  for (int64_t f = 0; f < nb_face; ++f) {
    const int32_t* fn = faces + f * nn;
    double meas, w = 0.0;
    r3 n0 = r3_load(coords, fn[0]), n1 = r3_load(coords, fn[1]);
    if (dim == 2) {
      meas = sqrt((n1.x - n0.x) * (n1.x - n0.x) + (n1.y - n0.y) * (n1.y - n0.y));
      if (kind == 0 && nb_value > 1) {
        double norm_N = sqrt((n1.y - n0.y) * (n1.y - n0.y) + (n1.x - n0.x) * (n1.x - n0.x));
        double Nx = (n1.y - n0.y) / norm_N, Ny = (n0.x - n1.x) / norm_N;
        w = (Nx * values[0] + Ny * values[1]);
      }
    }
    else {
      r3 n2 = r3_load(coords, fn[2]);
      meas = area_tri3_unsigned(n0, n1, n2);
      if (kind == 0 && nb_value > 1) {
        r3 e1 = r3_sub(n1, n0), e2 = r3_sub(n2, n0);
        r3 nr = r3_cross(e1, e2);
        double norm = sqrt(nr.x * nr.x + nr.y * nr.y + nr.z * nr.z);
        w = ((nr.x / norm) * values[0] + (nr.y / norm) * values[1] + (nr.z / norm) * values[2]);
      }
    }
    for (int i = 0; i < nn; ++i) {
      int32_t nd = fn[i];
      if ((is_dirichlet && is_dirichlet[nd]) || (is_own && !is_own[nd])) continue;
      if (kind == 0)
        rhs[(int64_t)nd * b] += (nb_value > 1 ? w : values[0]) * meas / nn;
      else
        for (int k = 0; k < b; ++k) rhs[(int64_t)nd * b + k] += values[k] * meas / nn;
    }
  }
}

/* ========================================================================= */
/* Dirichlet                                                                  */
/* All operate on a scalar CSR view (rows[nb_dof+1], cols, values): for b>1    */
/* that is the expanded view of orc_bsr_to_csr over per-row-layout values.     */
/* ========================================================================= */

/* Penalty / weak penalty: modules/testlab/FemModule.cc:728-790 (host),
 * :1201-1313 (device K19): A[i,i] = P (weak: += P); b[i] = P*g  (set, not add) */
ORC_API int orc_dirichlet_penalty(int weak, double penalty, int32_t n, const int32_t* dof_ids, const double* g,
                                  const int32_t* rows, const int32_t* cols, double* values, double* rhs)
{
  for (int32_t k = 0; k < n; ++k) {
    int32_t d = dof_ids[k];
    int32_t p = find_block(rows, cols, d, d);
    if (p < 0) return -2;
    if (weak) values[p] += penalty; else values[p] = penalty;
    rhs[d] = penalty * g[k];
  }
  return 0;
}

/*
 * Row / row-column elimination, femutils/CsrDoFLinearSystemImpl.cc:
 *   applyMatrixTransformation (:235-242) = _fillRowColumnEliminationInfos (:187-230,
 *     saves pre-elimination A[row,col] for entries whose row or column DoF is
 *     RC-eliminated, ordered by (row,col): internal/OrderedRowColumnMap.h:49-54)
 *     -> _applyRowEliminationOnMatrix (:126-152) -> _applyRowColumnEliminationOnMatrix
 *     (:88-121, note `if (column_index > 0)`: column 0 is skipped — reference quirk,
 *     replicated when quirk_skip_col0 != 0) -> _applyForcedValuesToLhs (:50-72).
 *   applyRHSTransformation (:247-253) = DoFLinearSystemImplBase::_applyRowColumnEliminationToRHS
 *     (femutils/DoFLinearSystemImplBase.cc:55-88: for saved (row,col) in map order,
 *     row!=col, col owned, row RC-eliminated: rhs[col] -= A[row,col]*g_row)
 *     -> _applyRowOrRowColumnEliminationOnRHS (:157-182: rhs[row] = g_row).
 * elim_info[nb_dof] in {0,1,2} (femutils/FemUtilsGlobal.h:51-62), elim_value[nb_dof].
 * forced_info/forced_value may be NULL.
 */
ORC_API void orc_apply_elimination(int32_t nb_dof, const int32_t* rows, const int32_t* cols, double* values, double* rhs,
                                   const uint8_t* elim_info, const double* elim_value,
                                   const uint8_t* forced_info, const double* forced_value,
                                   const uint8_t* dof_is_own, int quirk_skip_col0)
{
  int has_rc = 0;
  for (int32_t i = 0; i < nb_dof; ++i) if (elim_info[i] == 2) { has_rc = 1; break; }
  /* saved pre-elimination values (CSR order == (row,col) order because columns ascend) */
  double* saved = NULL;
  if (has_rc) {
    saved = (double*)malloc(sizeof(double) * (size_t)(rows[nb_dof] > 0 ? rows[nb_dof] : 1));
    memcpy(saved, values, sizeof(double) * (size_t)rows[nb_dof]);
  }
  /* row elimination (ELIMINATE_ROW only) */
  for (int32_t i = 0; i < nb_dof; ++i)
    if (elim_info[i] == 1)
      for (int32_t p = rows[i]; p < rows[i + 1]; ++p) values[p] = (cols[p] == i) ? 1.0 : 0.0;
  /* row+column elimination on matrix */
  if (has_rc)
    for (int32_t i = 0; i < nb_dof; ++i) {
      int row_el = (elim_info[i] == 1) || (elim_info[i] == 2);
      for (int32_t p = rows[i]; p < rows[i + 1]; ++p) {
        int32_t c = cols[p];
        if (quirk_skip_col0 ? (c > 0) : (c >= 0)) {
          int col_el = (elim_info[c] == 1) || (elim_info[c] == 2);
          if (row_el || col_el) values[p] = (c == i) ? 1.0 : 0.0;
        }
      }
    }
  /* forced values */
  if (forced_info)
    for (int32_t i = 0; i < nb_dof; ++i)
      if (forced_info[i]) {
        int32_t p = find_block(rows, cols, i, i);
        if (p >= 0) values[p] = forced_value[i];
      }
  /* RHS: RC correction in (row,col) map order, then overwrite eliminated rows */
  if (has_rc) {
    for (int32_t i = 0; i < nb_dof; ++i)
      for (int32_t p = rows[i]; p < rows[i + 1]; ++p) {
        int32_t c = cols[p];
        if (c < 0) continue;
        if (!(elim_info[i] == 2 || elim_info[c] == 2)) continue; /* entry is in the map */
        if (c == i) continue;
        if (dof_is_own && !dof_is_own[c]) continue;
        if (elim_info[i] == 2) rhs[c] = rhs[c] - saved[p] * elim_value[i];
      }
    free(saved);
  }
  for (int32_t i = 0; i < nb_dof; ++i)
    if (elim_info[i] == 1 || elim_info[i] == 2) rhs[i] = elim_value[i];
}

/* ========================================================================= */
/* Small utilities used by tests / bench                                      */
/* ========================================================================= */

/* y = A x on a scalar CSR view (residual checks against golden solutions) */
ORC_API void orc_spmv(int32_t nb_row, const int32_t* rows, const int32_t* cols, const double* values, const double* x, double* y)
{
  for (int32_t r = 0; r < nb_row; ++r) {
    double s = 0.0;
    for (int32_t p = rows[r]; p < rows[r + 1]; ++p) s += values[p] * x[cols[p]];
    y[r] = s;
  }
}

/*
 * CPU baseline leg (bench.py cpu_baseline / --impl reference): the reference's
 * sequential CSR assembly, modules/testlab/CsrBiliAssembly.cc:23-182 =
 * _buildMatrixCsr (host pattern; here orc_build_pattern) + per-cell host K_e +
 * CsrFormat::matrixAddValue linear scan.  `owner_lo/hi` restrict written rows to
 * [owner_lo, owner_hi) which is how an MPI rank's isOwn gate acts on a slab
 * sub-domain (modules/testlab/CsrBiliAssembly.cc:174); cells [cell_lo, cell_hi)
 * are the sub-domain's own + ghost cells.  Thread-safe for disjoint owner ranges.
 */
ORC_API int orc_assemble_csr_host_range(int npc, int dim, int64_t cell_lo, int64_t cell_hi, int32_t owner_lo, int32_t owner_hi,
                                        const double* coords, const int32_t* conn, const int32_t* rows, const int32_t* cols, double* values)
{
  double K[16];
  for (int64_t c = cell_lo; c < cell_hi; ++c) {
    const int32_t* cn = conn + c * npc;
    if (element_matrix(npc, dim, ORC_OP_POISSON, ORC_FORM_HOST, NULL, coords, cn, K)) return -1;
    for (int a1 = 0; a1 < npc; ++a1) {
      int32_t r = cn[a1];
      if (r < owner_lo || r >= owner_hi) continue;
      for (int a2 = 0; a2 < npc; ++a2) {
        double v = K[a1 * npc + a2];
        if (v == 0.0) continue;
        int32_t p = find_block(rows, cols, r, cn[a2]);
        if (p < 0) return -2;
        values[p] += v;
      }
    }
  }
  return 0;
}

/*
 * Host pattern build as the reference's sequential CSR back-end does it
 * (modules/testlab/CsrBiliAssembly.cc:23-92): per node, diagonal first then the
 * node-node-via-edge neighbours.  Arcane's computeNodeNodeViaEdgeConnectivity is
 * external; here the neighbour lists come from a per-node scan of incident cells
 * (host, single pass, no global sort) and are emitted ascending.  Used only as the
 * timed CPU baseline's "BuildMatrix" leg.
 */
ORC_API int64_t orc_build_pattern_host(int npc, int32_t nb_node, int64_t nb_cell, const int32_t* conn, int32_t* rows, int32_t* columns, int64_t capacity)
{
  int64_t* ptr;
  int32_t* list;
  build_node_cells(npc, nb_node, nb_cell, conn, &ptr, &list);
  int64_t nnz = 0;
  int32_t tmp[4096];
  for (int32_t r = 0; r < nb_node; ++r) {
    int cnt = 0;
    tmp[cnt++] = r;
    for (int64_t q = ptr[r]; q < ptr[r + 1]; ++q) {
      const int32_t* cn = conn + (int64_t)list[q] * npc;
      for (int i = 0; i < npc; ++i) {
        int32_t v = cn[i];
        int found = 0;
        for (int t = 0; t < cnt; ++t) if (tmp[t] == v) { found = 1; break; }
        if (!found && cnt < 4096) tmp[cnt++] = v;
      }
    }
    qsort(tmp, (size_t)cnt, sizeof(int32_t), cmp_i32);
    rows[r] = (int32_t)nnz;
    if (nnz + cnt > capacity) { free(ptr); free(list); return -1; }
    memcpy(columns + nnz, tmp, sizeof(int32_t) * (size_t)cnt);
    nnz += cnt;
  }
  rows[nb_node] = (int32_t)nnz;
  free(ptr);
  free(list);
  return nnz;
}

/*
 * One "MPI rank" of the reference's sequential CSR back-end on a sub-domain
 * (bench.py --impl reference / cpu_baseline; run concurrently from N host threads,
 * which is how `mpirun -n N Testlab` splits the work: Arcane partitions the mesh,
 * every rank loops over its own + ghost cells and writes only the rows of the nodes
 * it owns, modules/testlab/CsrBiliAssembly.cc:23-92 (BuildMatrix) and :97-182
 * (AddAndCompute, isOwn gate :174); no communication during assembly).
 * Sub-domain = cells [cell_lo, cell_hi) (own + ghost layer) and owned nodes
 * [owner_lo, owner_hi).
 *
 * Init time (untimed, like FemModule::startInit, modules/testlab/FemModule.cc:124,136):
 * orc_reference_init builds the node-node-via-edge connectivity of the owned nodes
 * (Arcane's MeshUtils::computeNodeNodeViaEdgeConnectivity is external; the neighbour
 * lists are emitted ascending here).
 *
 * Timed, per assembly:
 *   BuildMatrix   CsrFormat::initialize (femutils/CsrFormatMatrix.cc:35-58: row / column /
 *                 value / rows_nb_column arrays allocated and filled with -1 / -1 / 0 / 0 on
 *                 every assembly) + the walk of _buildMatrixCsr (CsrBiliAssembly.cc:79-91):
 *                 per node setCoordinates(diagonal) then setCoordinates(neighbour) for
 *                 every entry of the node-node view -- no sort, no search.
 *   AddAndCompute element matrix (host form) + matrixAddValue with the linear scan of
 *                 indexValue (femutils/CsrFormatMatrix.h:58-87), exact zeros skipped.
 * Returns nnz of the owned rows; *checksum = sum of values so the work cannot be elided.
 * If out_rows/out_cols/out_vals are non-NULL they receive the rank's arrays (rows relative
 * to the rank's first entry, columns in the reference's order: diagonal first).
 * seconds[0] = BuildMatrix, seconds[1] = AddAndCompute.
 *
 * orc_reference_rank (no handle) is the first-assembly form: it builds the connectivity
 * inside the timed BuildMatrix (what a run pays once), kept for the record.
 */
#include <time.h>
static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
  int32_t owner_lo, owner_hi;
  int64_t* ptr;  /* nb_own + 1 */
  int32_t* list; /* neighbours (self excluded), ascending */
} orc_nn_t;

ORC_API void orc_reference_free(void* h)
{
  orc_nn_t* nn = (orc_nn_t*)h;
  if (!nn) return;
  free(nn->ptr);
  free(nn->list);
  free(nn);
}

ORC_API void* orc_reference_init(int npc, int64_t cell_lo, int64_t cell_hi, int32_t owner_lo, int32_t owner_hi, const int32_t* conn)
{
  const int32_t nb_own = owner_hi - owner_lo;
  orc_nn_t* nn = (orc_nn_t*)calloc(1, sizeof(orc_nn_t));
  nn->owner_lo = owner_lo;
  nn->owner_hi = owner_hi;
  /* node -> cells of the sub-domain restricted to owned nodes */
  int64_t* ptr = (int64_t*)calloc((size_t)nb_own + 1, sizeof(int64_t));
  for (int64_t c = cell_lo; c < cell_hi; ++c)
    for (int i = 0; i < npc; ++i) {
      int32_t v = conn[c * npc + i];
      if (v >= owner_lo && v < owner_hi) ptr[v - owner_lo + 1]++;
    }
  for (int32_t n = 0; n < nb_own; ++n) ptr[n + 1] += ptr[n];
  int32_t* list = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ptr[nb_own] > 0 ? ptr[nb_own] : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(nb_own + 1));
  memcpy(fill, ptr, sizeof(int64_t) * (size_t)(nb_own + 1));
  for (int64_t c = cell_lo; c < cell_hi; ++c)
    for (int i = 0; i < npc; ++i) {
      int32_t v = conn[c * npc + i];
      if (v >= owner_lo && v < owner_hi) list[fill[v - owner_lo]++] = (int32_t)c;
    }
  free(fill);
  int64_t cap = 0;
  for (int32_t n = 0; n < nb_own; ++n) cap += (ptr[n + 1] - ptr[n]) * (npc - 1); /* upper bound */
  nn->ptr = (int64_t*)malloc(sizeof(int64_t) * ((size_t)nb_own + 1));
  nn->list = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap > 0 ? cap : 1));
  int64_t tot = 0;
  int32_t tmp[4096];
  for (int32_t n = 0; n < nb_own; ++n) {
    const int32_t r = owner_lo + n;
    int cnt = 0;
    for (int64_t q = ptr[n]; q < ptr[n + 1]; ++q) {
      const int32_t* cn = conn + (int64_t)list[q] * npc;
      for (int i = 0; i < npc; ++i) {
        int32_t v = cn[i];
        if (v == r) continue;
        int found = 0;
        for (int t = 0; t < cnt; ++t) if (tmp[t] == v) { found = 1; break; }
        if (!found && cnt < 4096) tmp[cnt++] = v;
      }
    }
    qsort(tmp, (size_t)cnt, sizeof(int32_t), cmp_i32);
    nn->ptr[n] = tot;
    memcpy(nn->list + tot, tmp, sizeof(int32_t) * (size_t)cnt);
    tot += cnt;
  }
  nn->ptr[nb_own] = tot;
  free(ptr);
  free(list);
  return nn;
}

ORC_API int64_t orc_reference_rank_nn(const void* handle, int npc, int dim, int64_t cell_lo, int64_t cell_hi, int32_t owner_lo, int32_t owner_hi,
                                      const double* coords, const int32_t* conn, double* checksum, double* seconds,
                                      int32_t* out_rows, int32_t* out_cols, double* out_vals, int64_t out_capacity)
{
  const orc_nn_t* nn = (const orc_nn_t*)handle;
  if (!nn || nn->owner_lo != owner_lo || nn->owner_hi != owner_hi) return -3;
  const double t0 = now_s();
  const int32_t nb_own = owner_hi - owner_lo;
  /* BuildMatrix: CsrFormat::initialize (allocate + 4 fills), then the append walk */
  const int64_t nnz = (int64_t)nb_own + nn->ptr[nb_own]; /* nbNode + 2 nbEdge */
  int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * ((size_t)nb_own + 1));
  int32_t* cols = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  double* vals = (double*)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  int32_t* rows_nb_column = (int32_t*)malloc(sizeof(int32_t) * ((size_t)nb_own + 1));
  for (int32_t n = 0; n < nb_own; ++n) rows[n] = -1;
  for (int64_t i = 0; i < nnz; ++i) cols[i] = -1;
  for (int64_t i = 0; i < nnz; ++i) vals[i] = 0.0;
  for (int32_t n = 0; n < nb_own; ++n) rows_nb_column[n] = 0;
  int64_t last = 0;
  for (int32_t n = 0; n < nb_own; ++n) {
    if (rows[n] == -1) rows[n] = (int32_t)last; /* setCoordinates, femutils/CsrFormatMatrix.h:104-112 */
    cols[last++] = owner_lo + n;
    for (int64_t q = nn->ptr[n]; q < nn->ptr[n + 1]; ++q) cols[last++] = nn->list[q];
  }
  rows[nb_own] = (int32_t)last;
  const double t1 = now_s();
  /* AddAndCompute */
  double K[16];
  int rc = 0;
  for (int64_t c = cell_lo; c < cell_hi && !rc; ++c) {
    const int32_t* cn = conn + c * npc;
    if (element_matrix(npc, dim, ORC_OP_POISSON, ORC_FORM_HOST, NULL, coords, cn, K)) { rc = -1; break; }
    for (int a1 = 0; a1 < npc; ++a1) {
      int32_t r = cn[a1];
      if (r < owner_lo || r >= owner_hi) continue;
      const int32_t rb = rows[r - owner_lo], re = rows[r - owner_lo + 1];
      for (int a2 = 0; a2 < npc; ++a2) {
        double v = K[a1 * npc + a2];
        if (v == 0.0) continue; /* femutils/CsrFormatMatrix.h:64 */
        int32_t p = -1;
        for (int32_t q = rb; q < re; ++q) if (cols[q] == cn[a2]) { p = q; break; }
        if (p < 0) { rc = -2; break; }
        vals[p] += v;
      }
    }
  }
  const double t2 = now_s();
  double s = 0.0;
  for (int64_t i = 0; i < nnz; ++i) s += vals[i];
  if (checksum) *checksum = s;
  if (seconds) { seconds[0] = t1 - t0; seconds[1] = t2 - t1; }
  if (out_rows && out_cols && out_vals && nnz <= out_capacity) {
    memcpy(out_rows, rows, sizeof(int32_t) * ((size_t)nb_own + 1));
    memcpy(out_cols, cols, sizeof(int32_t) * (size_t)nnz);
    memcpy(out_vals, vals, sizeof(double) * (size_t)nnz);
  }
  free(rows);
  free(cols);
  free(vals);
  free(rows_nb_column);
  return rc ? rc : nnz;
}

/* first-assembly form: connectivity built inside the timed BuildMatrix */
ORC_API int64_t orc_reference_rank(int npc, int dim, int64_t cell_lo, int64_t cell_hi, int32_t owner_lo, int32_t owner_hi,
                                   const double* coords, const int32_t* conn, double* checksum, double* seconds,
                                   int32_t* out_rows, int32_t* out_cols, double* out_vals, int64_t out_capacity)
{
  const double t0 = now_s();
  void* h = orc_reference_init(npc, cell_lo, cell_hi, owner_lo, owner_hi, conn);
  const double t1 = now_s();
  const int64_t r = orc_reference_rank_nn(h, npc, dim, cell_lo, cell_hi, owner_lo, owner_hi, coords, conn, checksum, seconds, out_rows, out_cols, out_vals, out_capacity);
  if (seconds) seconds[0] += t1 - t0;
  orc_reference_free(h);
  return r;
}
