// Linear form (constant source), Dirichlet enforcement, solver hand-off views.
#include "element.cuh"

namespace afb {

// ---------------------------------------------------------------------------------------------
// measure of a P1 cell
// ---------------------------------------------------------------------------------------------
template <int NPC>
__device__ __forceinline__ double cell_measure(const double* __restrict__ coords, const int32_t* __restrict__ cn, bool signed_area)
{
  if constexpr (NPC == 3) {
    int32_t nd[3] = { __ldg(cn), __ldg(cn + 1), __ldg(cn + 2) };
    Tri3Geom g;
    g.init(coords, nd, signed_area);
    return g.area;
  }
  else {
    int4 v = __ldg(reinterpret_cast<const int4*>(cn));
    int32_t nd[4] = { v.x, v.y, v.z, v.w };
    Tet4Geom g;
    g.init(coords, nd);
    return g.vol;
  }
}

// cell-wise, atomics: modules/testlab/FemModule.cc:1392-1404,1478-1487 (K25);
// modules/elasticity/BodyForce.h:93-104; modules/bilaplacian/FemModule.cc:157-172
template <int NPC>
__global__ void __launch_bounds__(256) k_rhs_source_cellwise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own,
                                                              const uint8_t* __restrict__ dir_node, int64_t nb_cell, int b, double f0, double f1, double f2,
                                                              bool signed_area, double* __restrict__ rhs)
{
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb_cell) return;
  const int32_t* cn = conn + c * NPC;
  const double meas = cell_measure<NPC>(coords, cn, signed_area);
  const double f[3] = { f0, f1, f2 };
#pragma unroll
  for (int i = 0; i < NPC; ++i) {
    int32_t nd = __ldg(cn + i);
    if ((dir_node && dir_node[nd]) || (is_own && !is_own[nd])) continue;
    for (int k = 0; k < b; ++k)
      if (f[k] != 0.0) atomicAdd(rhs + (int64_t)nd * b + k, f[k] * meas / NPC);
  }
}

// node-wise, rhs = sum: femutils/ArcaneFemFunctionsGpu.h:675-708
template <int NPC>
__global__ void __launch_bounds__(128) k_rhs_source_nodewise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own,
                                                              const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int32_t nb_node, int b,
                                                              double f0, double f1, double f2, double* __restrict__ rhs)
{
  int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb_node) return;
  if (is_own && !is_own[r]) return;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int q = nc_ptr[r]; q < nc_ptr[r + 1]; ++q) {
    const double meas = cell_measure<NPC>(coords, conn + (int64_t)nc_list[q] * NPC, false);
    s0 += f0 * meas / NPC;
    s1 += f1 * meas / NPC;
    s2 += f2 * meas / NPC;
  }
  rhs[(int64_t)r * b] = s0;
  if (b > 1) rhs[(int64_t)r * b + 1] = s1;
  if (b > 2) rhs[(int64_t)r * b + 2] = s2;
}

// Quad4 / Hexa8: integral of every shape function over the cell by the 2x2 / 2x2x2 Gauss rule, w_i = sum_gp N_i detJ
// (femutils/ArcaneFemFunctions.cc:222-290, 437-483; device twin femutils/ArcaneFemFunctionsGpu.cc:393-470, 875-960)
template <int NPC>
__device__ __forceinline__ void q1_source_weights(const double* __restrict__ coords, const int32_t* __restrict__ cn, double (&w)[NPC])
{
  constexpr int DIM = NPC == 4 ? 2 : 3;
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
  double x[NPC], y[NPC], z[NPC];
#pragma unroll
  for (int a = 0; a < NPC; ++a) {
    load3(coords, __ldg(cn + a), x[a], y[a], z[a]);
    w[a] = 0.0;
  }
  // node a of the reference element sits at (sx, sy, sz)[a]: counter-clockwise, bottom face then top face
  auto sx = [](int a) { return ((a & 3) == 1 || (a & 3) == 2) ? 1.0 : -1.0; };
  auto sy = [](int a) { return (a & 2) ? 1.0 : -1.0; };
  auto sz = [](int a) { return (a & 4) ? 1.0 : -1.0; };
#pragma unroll
  for (int g = 0; g < (1 << DIM); ++g) {
    const double xi = gp[(g >> (DIM - 1)) & 1], eta = gp[(g >> (DIM - 2)) & 1], zeta = DIM == 3 ? gp[g & 1] : 0.0;
    double N[NPC], J[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
#pragma unroll
    for (int a = 0; a < NPC; ++a) {
      const double fx = 1.0 + sx(a) * xi, fy = 1.0 + sy(a) * eta, fz = DIM == 3 ? 1.0 + sz(a) * zeta : 1.0;
      const double s = DIM == 3 ? 0.125 : 0.25;
      N[a] = s * fx * fy * fz;
      const double dxi = sx(a) * s * fy * fz, det_ = sy(a) * s * fx * fz;
      J[0][0] += dxi * x[a]; J[0][1] += dxi * y[a];
      J[1][0] += det_ * x[a]; J[1][1] += det_ * y[a];
      if constexpr (DIM == 3) {
        const double dze = sz(a) * s * fx * fy;
        J[0][2] += dxi * z[a]; J[1][2] += det_ * z[a];
        J[2][0] += dze * x[a]; J[2][1] += dze * y[a]; J[2][2] += dze * z[a];
      }
    }
    const double detJ = DIM == 2 ? J[0][0] * J[1][1] - J[0][1] * J[1][0]
                                 : J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                                     J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
#pragma unroll
    for (int a = 0; a < NPC; ++a) w[a] += N[a] * detJ;
  }
}

template <int NPC>
__global__ void __launch_bounds__(256) k_rhs_source_q1_cellwise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own,
                                                                 const uint8_t* __restrict__ dir_node, int64_t nb_cell, int b, double f0, double f1, double f2,
                                                                 double* __restrict__ rhs)
{
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb_cell) return;
  const int32_t* cn = conn + c * NPC;
  double w[NPC];
  q1_source_weights<NPC>(coords, cn, w);
  const double f[3] = { f0, f1, f2 };
#pragma unroll
  for (int i = 0; i < NPC; ++i) {
    const int32_t nd = __ldg(cn + i);
    if ((dir_node && dir_node[nd]) || (is_own && !is_own[nd])) continue;
    for (int k = 0; k < b; ++k)
      if (f[k] != 0.0) atomicAdd(rhs + (int64_t)nd * b + k, w[i] * f[k]);
  }
}

// node-wise (no atomics): every owned node sums its share of its incident cells, ascending cell ids
template <int NPC>
__global__ void __launch_bounds__(128) k_rhs_source_q1_nodewise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own,
                                                                 const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int32_t nb_node, int b,
                                                                 double f0, double f1, double f2, double* __restrict__ rhs)
{
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb_node) return;
  if (is_own && !is_own[r]) return;
  double s = 0.0;
  for (int q = nc_ptr[r]; q < nc_ptr[r + 1]; ++q) {
    const int32_t* cn = conn + (int64_t)nc_list[q] * NPC;
    double w[NPC];
    q1_source_weights<NPC>(coords, cn, w);
#pragma unroll
    for (int i = 0; i < NPC; ++i)
      if (__ldg(cn + i) == r) s += w[i];
  }
  rhs[(int64_t)r * b] = f0 * s;
  if (b > 1) rhs[(int64_t)r * b + 1] = f1 * s;
  if (b > 2) rhs[(int64_t)r * b + 2] = f2 * s;
}

int rhs_source(afb_ctx* ctx, const double* f, int nb_f, int nodewise, int signed_area)
{
  const bool q1 = (ctx->dim == 2 && ctx->npc == 4) || (ctx->dim == 3 && ctx->npc == 8);
  AFB_REQUIRE(ctx->npc == ctx->dim + 1 || q1, AFB_ERR_UNSUPPORTED, "constant source term is implemented for P1 simplices, Quad4 and Hexa8 (%d-node cells in dimension %d)", ctx->npc,
              ctx->dim);
  if (q1) {
    AFB_REQUIRE(nb_f >= 1 && nb_f <= 3 && nb_f <= ctx->b, AFB_ERR_INVALID, "source has %d components, matrix has %d dof per node", nb_f, ctx->b);
    double ff[3] = { 0, 0, 0 };
    for (int k = 0; k < nb_f; ++k) ff[k] = f[k];
    const uint8_t* own = ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>();
    if (!nodewise) {
      if (ctx->nb_cell == 0) return AFB_OK;
      const uint8_t* dir = ctx->has_dir_nodes ? ctx->dir_node.as<uint8_t>() : nullptr;
      const int grid = grid_for(ctx->nb_cell, 256);
      if (ctx->npc == 4)
        k_rhs_source_q1_cellwise<4><<<grid, 256, 0, ctx->stream>>>(ctx->coords.as<double>(), ctx->conn.as<int32_t>(), own, dir, ctx->nb_cell, ctx->b, ff[0], ff[1], ff[2], ctx->rhs.as<double>());
      else
        k_rhs_source_q1_cellwise<8><<<grid, 256, 0, ctx->stream>>>(ctx->coords.as<double>(), ctx->conn.as<int32_t>(), own, dir, ctx->nb_cell, ctx->b, ff[0], ff[1], ff[2], ctx->rhs.as<double>());
    }
    else {
      const int grid = grid_for(ctx->nb_node, 128);
      if (ctx->npc == 4)
        k_rhs_source_q1_nodewise<4><<<grid, 128, 0, ctx->stream>>>(ctx->coords.as<double>(), ctx->conn.as<int32_t>(), own, ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(), ctx->nb_node, ctx->b, ff[0], ff[1], ff[2], ctx->rhs.as<double>());
      else
        k_rhs_source_q1_nodewise<8><<<grid, 128, 0, ctx->stream>>>(ctx->coords.as<double>(), ctx->conn.as<int32_t>(), own, ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(), ctx->nb_node, ctx->b, ff[0], ff[1], ff[2], ctx->rhs.as<double>());
    }
    AFB_LAUNCH_CHECK(ctx);
    return AFB_OK;
  }
  AFB_REQUIRE(nb_f >= 1 && nb_f <= 3 && nb_f <= ctx->b, AFB_ERR_INVALID, "source has %d components, matrix has %d dof per node", nb_f, ctx->b);
  double ff[3] = { 0, 0, 0 };
  for (int k = 0; k < nb_f; ++k) ff[k] = f[k];
  const double* coords = ctx->coords.as<double>();
  const int32_t* conn = ctx->conn.as<int32_t>();
  const uint8_t* own = ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>();
  double* rhs = ctx->rhs.as<double>();
  if (!nodewise) {
    if (ctx->nb_cell == 0) return AFB_OK;
    const uint8_t* dir = ctx->has_dir_nodes ? ctx->dir_node.as<uint8_t>() : nullptr;
    int grid = grid_for(ctx->nb_cell, 256);
    if (ctx->npc == 3)
      k_rhs_source_cellwise<3><<<grid, 256, 0, ctx->stream>>>(coords, conn, own, dir, ctx->nb_cell, ctx->b, ff[0], ff[1], ff[2], signed_area != 0, rhs);
    else
      k_rhs_source_cellwise<4><<<grid, 256, 0, ctx->stream>>>(coords, conn, own, dir, ctx->nb_cell, ctx->b, ff[0], ff[1], ff[2], signed_area != 0, rhs);
  }
  else {
    int grid = grid_for(ctx->nb_node, 128);
    const int32_t* ptr = ctx->nc_ptr.as<int32_t>();
    const int32_t* list = ctx->nc_list.as<int32_t>();
    if (ctx->npc == 3)
      k_rhs_source_nodewise<3><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, ptr, list, ctx->nb_node, ctx->b, ff[0], ff[1], ff[2], rhs);
    else
      k_rhs_source_nodewise<4><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, ptr, list, ctx->nb_node, ctx->b, ff[0], ff[1], ff[2], rhs);
  }
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// boundary integrals on P1 faces (edges / triangles): constant flux, q.n flux, traction.  One thread per face, fp64
// atomics like the reference (modules/testlab/FemModule.cc:1574-1693, femutils/ArcaneFemFunctionsGpu.cc:679-738,1082-1141;
// traction: femutils/ArcaneFemFunctions.h:2188-2220,2854-2885 -- a host loop upstream).  Arithmetic as the helpers
// computeLengthFace / computeAreaTria / computeNormalFace / computeNormalTriangle (ArcaneFemFunctionsGpu.h:92-102,139-214).
template <int DIM>
__global__ void __launch_bounds__(256) k_rhs_neumann(const double* __restrict__ coords, const int32_t* __restrict__ faces, int64_t nb_face, const uint8_t* __restrict__ is_own,
                                                      const uint8_t* __restrict__ dir_node, int b, int kind, int nb_value, double v0, double v1, double v2,
                                                      double* __restrict__ rhs)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nb_face) return;
  const int32_t* fn = faces + f * DIM;
  double x0, y0, z0, x1, y1, z1;
  load3(coords, __ldg(fn), x0, y0, z0);
  load3(coords, __ldg(fn + 1), x1, y1, z1);
  double meas, w = v0;
  if (DIM == 2) {
    meas = sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
    if (kind == AFB_NEUMANN_FLUX && nb_value > 1) {
      const double norm_n = sqrt((y1 - y0) * (y1 - y0) + (x1 - x0) * (x1 - x0));
      w = ((y1 - y0) / norm_n) * v0 + ((x0 - x1) / norm_n) * v1;
    }
  }
  else {
    double x2, y2, z2;
    load3(coords, __ldg(fn + 2), x2, y2, z2);
    const double ax = x1 - x0, ay = y1 - y0, az = z1 - z0, bx = x2 - x0, by = y2 - y0, bz = z2 - z0;
    const double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const double norm = sqrt(nx * nx + ny * ny + nz * nz);
    meas = norm / 2.0;
    if (kind == AFB_NEUMANN_FLUX && nb_value > 1) w = (nx / norm) * v0 + (ny / norm) * v1 + (nz / norm) * v2;
  }
  const double t[3] = { v0, v1, v2 };
#pragma unroll
  for (int i = 0; i < DIM; ++i) {
    const int32_t nd = __ldg(fn + i);
    if ((dir_node && dir_node[nd]) || (is_own && !is_own[nd])) continue;
    if (kind == AFB_NEUMANN_FLUX) atomicAdd(rhs + (int64_t)nd * b, w * meas / DIM);
    else
      for (int k = 0; k < b; ++k) atomicAdd(rhs + (int64_t)nd * b + k, t[k] * meas / DIM);
  }
}

// Quad4 faces of a Hexa8 mesh: 2x2 Gauss rule on the bilinear patch, detJ = |dr/dxi x dr/deta|, the unit normal taken at every
// Gauss point from the face's node order (femutils/ArcaneFemFunctions.h:1843-1953; device twin ArcaneFemFunctionsGpu.cc:1146-1260)
__global__ void __launch_bounds__(256) k_rhs_neumann_quad4(const double* __restrict__ coords, const int32_t* __restrict__ faces, int64_t nb_face,
                                                            const uint8_t* __restrict__ is_own, const uint8_t* __restrict__ dir_node, int b, int kind, int nb_value, double v0,
                                                            double v1, double v2, double* __restrict__ rhs)
{
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nb_face) return;
  const int32_t* fn = faces + f * 4;
  double x[4], y[4], z[4], acc[4] = { 0.0, 0.0, 0.0, 0.0 };
#pragma unroll
  for (int i = 0; i < 4; ++i) load3(coords, __ldg(fn + i), x[i], y[i], z[i]);
  const double gp[2] = { -0.57735026918962576451, 0.57735026918962576451 };
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const double xi = gp[g >> 1], eta = gp[g & 1];
    double N[4], t1x = 0, t1y = 0, t1z = 0, t2x = 0, t2y = 0, t2z = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double sx = (i == 1 || i == 2) ? 1.0 : -1.0, sy = (i & 2) ? 1.0 : -1.0;
      N[i] = 0.25 * (1.0 + sx * xi) * (1.0 + sy * eta);
      const double dxi = sx * 0.25 * (1.0 + sy * eta), det_ = sy * 0.25 * (1.0 + sx * xi);
      t1x += dxi * x[i]; t1y += dxi * y[i]; t1z += dxi * z[i];
      t2x += det_ * x[i]; t2y += det_ * y[i]; t2z += det_ * z[i];
    }
    double nx = t1y * t2z - t1z * t2y, ny = t1z * t2x - t1x * t2z, nz = t1x * t2y - t1y * t2x;
    const double detJ = sqrt(nx * nx + ny * ny + nz * nz);
    nx /= detJ; ny /= detJ; nz /= detJ;
    // flux: value or q.n; traction (applyTractionToRhsHexa8, femutils/ArcaneFemFunctions.h:2222-2315): the shape-function integral, times t[k] below
    const double q = kind == AFB_NEUMANN_TRACTION ? 1.0 : (nb_value == 1 ? v0 : nx * v0 + ny * v1 + nz * v2);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += q * N[j] * detJ;
  }
  const double t[3] = { v0, v1, v2 };
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t nd = __ldg(fn + j);
    if ((dir_node && dir_node[nd]) || (is_own && !is_own[nd])) continue;
    if (kind == AFB_NEUMANN_TRACTION) {
      for (int k = 0; k < b; ++k) atomicAdd(rhs + (int64_t)nd * b + k, t[k] * acc[j]);
    }
    else atomicAdd(rhs + (int64_t)nd * b, acc[j]);
  }
}

int rhs_neumann(afb_ctx* ctx, int64_t nb_face, const int32_t* faces_dev, int kind, int nb_value, const double* values, int skip_dirichlet)
{
  if (nb_face <= 0) return AFB_OK;
  double v[3] = { 0.0, 0.0, 0.0 };
  for (int k = 0; k < nb_value && k < 3; ++k) v[k] = values[k];
  const uint8_t* own = ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>();
  const uint8_t* dir = (skip_dirichlet && ctx->has_dir_nodes) ? ctx->dir_node.as<uint8_t>() : nullptr;
  const int grid = grid_for(nb_face, 256);
  if (ctx->dim == 3 && ctx->npc == 8)
    k_rhs_neumann_quad4<<<grid, 256, 0, ctx->stream>>>(ctx->coords.as<double>(), faces_dev, nb_face, own, dir, ctx->b, kind, nb_value, v[0], v[1], v[2], ctx->rhs.as<double>());
  else if (ctx->dim == 2)
    k_rhs_neumann<2><<<grid, 256, 0, ctx->stream>>>(ctx->coords.as<double>(), faces_dev, nb_face, own, dir, ctx->b, kind, nb_value, v[0], v[1], v[2], ctx->rhs.as<double>());
  else
    k_rhs_neumann<3><<<grid, 256, 0, ctx->stream>>>(ctx->coords.as<double>(), faces_dev, nb_face, own, dir, ctx->b, kind, nb_value, v[0], v[1], v[2], ctx->rhs.as<double>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// ---------------------------------------------------------------------------------------------
// scalar (dof_row, dof_col) -> index into values for the stored layout, or -1
// (BSRMatrix::findValueIndex, femutils/BSRFormat.cc:79-106)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t scalar_slot(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int layout, int32_t dr, int32_t dc)
{
  const int32_t br = dr / b, bc = dc / b;
  const int rb = __ldg(rows + br), re = __ldg(rows + br + 1);
  if (re <= rb) return -1;
  const int p = find_col(cols, rb, re, bc);
  if (__ldg(cols + p) != bc) return -1;
  const int i = dr - br * b, j = dc - bc * b;
  if (layout == AFB_LAYOUT_PER_BLOCK) return (int64_t)p * b * b + i * b + j;
  return (int64_t)rb * b * b + (int64_t)b * ((p - rb) + (int64_t)i * (re - rb)) + j;
}

__global__ void k_lookup_slots(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int layout, int64_t n,
                               const int32_t* __restrict__ dr, const int32_t* __restrict__ dc, int64_t* __restrict__ slots)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) slots[i] = scalar_slot(rows, cols, b, layout, dr[i], dc[i]);
}

__global__ void k_add_values_at(int64_t n, const int64_t* __restrict__ slots, const double* __restrict__ contrib, double* __restrict__ values)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && slots[i] >= 0) values[slots[i]] += contrib[i];
}

__global__ void k_renumber(int64_t n, const int32_t* __restrict__ cols, const int32_t* __restrict__ l2g, int32_t* __restrict__ out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const int32_t c = cols[i];
    out[i] = c >= 0 ? __ldg(l2g + c) : c;
  }
}

__global__ void k_iota(int32_t first, int32_t n, int32_t* __restrict__ out)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = first + i;
}

int fill_iota(afb_ctx* ctx, int32_t first, int32_t n, int32_t* out)
{
  if (n <= 0) return AFB_OK;
  k_iota<<<grid_for(n, 256), 256, 0, ctx->stream>>>(first, n, out);
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

int renumber_columns(afb_ctx* ctx, const int32_t* dof_local_to_global, int32_t* out)
{
  const int32_t* cols = ctx->cols.as<int32_t>();
  int64_t n = ctx->nnz;
  if (ctx->b > 1) {
    AFB_TRY(ensure_scalar_csr(ctx));
    cols = ctx->csr_cols.as<int32_t>();
    n = ctx->nnz * ctx->b * ctx->b;
  }
  if (n <= 0) return AFB_OK;
  k_renumber<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, cols, dof_local_to_global, out);
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

int lookup_value_slots(afb_ctx* ctx, int64_t n, const int32_t* dof_rows, const int32_t* dof_cols, int64_t* slots)
{
  if (n <= 0) return AFB_OK;
  k_lookup_slots<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->b, ctx->layout, n, dof_rows, dof_cols, slots);
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

int add_values_at(afb_ctx* ctx, int64_t n, const int64_t* slots, const double* contrib)
{
  if (n <= 0) return AFB_OK;
  k_add_values_at<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, slots, contrib, ctx->values.as<double>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// ---------------------------------------------------------------------------------------------
// Dirichlet: penalty (K19), flags, elimination (K20-K23)
// ---------------------------------------------------------------------------------------------
__global__ void k_penalty(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int layout, int weak, double penalty, int32_t n,
                          const int32_t* __restrict__ dofs, const double* __restrict__ g, double* __restrict__ values, double* __restrict__ rhs, int* __restrict__ err)
{
  int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int32_t d = dofs[k];
  const int64_t s = scalar_slot(rows, cols, b, layout, d, d);
  if (s < 0) { *err = 1; return; }
  if (weak) values[s] += penalty; else values[s] = penalty;
  rhs[d] = penalty * g[k];
}

int dirichlet_penalty(afb_ctx* ctx, int weak, double penalty, int32_t n, const int32_t* dof_ids, const double* g)
{
  if (n <= 0) return AFB_OK;
  AFB_TRY(ctx->tmp_flag.reserve(sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, sizeof(int), ctx->stream));
  k_penalty<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->b, ctx->layout, weak, penalty, n, dof_ids, g,
                                                       ctx->values.as<double>(), ctx->rhs.as<double>(), ctx->tmp_flag.as<int>());
  AFB_LAUNCH_CHECK(ctx);
  int err = 0;
  AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  AFB_REQUIRE(err == 0, AFB_ERR_INVALID, "Dirichlet DoF without a diagonal entry (BSRMatrix(findValueIndex): Value not found)");
  return AFB_OK;
}

__global__ void k_scatter_flags(uint8_t* __restrict__ flags, double* __restrict__ vals, uint8_t flag, int32_t n, const int32_t* __restrict__ ids, const double* __restrict__ v)
{
  int32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  flags[ids[k]] = flag;
  if (vals) vals[ids[k]] = v[k];
}

int scatter_flags(afb_ctx* ctx, uint8_t* flags, double* vals, uint8_t flag, int32_t n, const int32_t* ids, const double* v)
{
  if (n <= 0) return AFB_OK;
  k_scatter_flags<<<grid_for(n, 128), 128, 0, ctx->stream>>>(flags, vals, flag, n, ids, v);
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// One thread per scalar row.  Order of the reference (CsrDoFLinearSystemImpl.cc:235-242):
// row elimination (ELIMINATE_ROW rows only, :126-152), then the row+column pass over every
// row when any RC elimination exists (:88-121, incl. the `column_index > 0` quirk), then the
// forced diagonal values (:50-72).  All three touch only the thread's own row.
__global__ void __launch_bounds__(128) k_matrix_transformation(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int layout, int32_t nb_dof,
                                                                const uint8_t* __restrict__ elim_info, const uint8_t* __restrict__ forced_info,
                                                                const double* __restrict__ forced_value, int has_rc, int quirk, double* __restrict__ values)
{
  int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nb_dof) return;
  const int32_t br = row / b;
  const int i = row - br * b;
  const int rb = rows[br], re = rows[br + 1], nz = re - rb;
  const uint8_t info = elim_info ? elim_info[row] : 0;
  const bool forced = forced_info && forced_info[row];
  if (info == 0 && !has_rc && !forced) return;
  const bool row_el = info == AFB_ELIMINATE_ROW || info == AFB_ELIMINATE_ROW_COLUMN;
  for (int p = rb; p < re; ++p) {
    const int32_t bc = cols[p];
    for (int j = 0; j < b; ++j) {
      const int32_t c = bc * b + j;
      const int64_t idx = (layout == AFB_LAYOUT_PER_BLOCK) ? (int64_t)p * b * b + i * b + j : (int64_t)rb * b * b + (int64_t)b * ((p - rb) + (int64_t)i * nz) + j;
      if (info == AFB_ELIMINATE_ROW) values[idx] = (c == row) ? 1.0 : 0.0;
      if (has_rc && (quirk ? (c > 0) : true)) {
        const uint8_t ci = elim_info[c];
        const bool col_el = ci == AFB_ELIMINATE_ROW || ci == AFB_ELIMINATE_ROW_COLUMN;
        if (row_el || col_el) values[idx] = (c == row) ? 1.0 : 0.0;
      }
      if (forced && c == row) values[idx] = forced_value[row];
    }
  }
}

// rhs[col] -= A_saved[row,col]*g_row for RC-eliminated rows, ascending row (the (row,col)
// order of OrderedRowColumnMap restricted to one column; DoFLinearSystemImplBase.cc:55-88),
// then rhs[row] = g_row for eliminated rows (CsrDoFLinearSystemImpl.cc:157-182).
// The pattern is structurally symmetric, so the rows holding column `col` are exactly the
// columns of row `col`.
__global__ void __launch_bounds__(128) k_rhs_transformation(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int layout, int32_t nb_dof,
                                                             const uint8_t* __restrict__ elim_info, const double* __restrict__ elim_value,
                                                             const uint8_t* __restrict__ is_own, const double* __restrict__ saved, double* __restrict__ rhs)
{
  int32_t col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nb_dof) return;
  const int32_t bcn = col / b;
  if (saved && !(is_own && !is_own[bcn])) {
    double v = rhs[col];
    bool touched = false;
    for (int p = rows[bcn]; p < rows[bcn + 1]; ++p) {
      const int32_t brn = cols[p];
      for (int i = 0; i < b; ++i) {
        const int32_t row = brn * b + i;
        if (row == col || elim_info[row] != AFB_ELIMINATE_ROW_COLUMN) continue;
        const int64_t s = scalar_slot(rows, cols, b, layout, row, col);
        if (s < 0) continue;
        v = v - saved[s] * elim_value[row];
        touched = true;
      }
    }
    if (touched) rhs[col] = v;
  }
  const uint8_t info = elim_info[col];
  if (info == AFB_ELIMINATE_ROW || info == AFB_ELIMINATE_ROW_COLUMN) rhs[col] = elim_value[col];
}

int apply_matrix_transformation(afb_ctx* ctx, int quirk)
{
  const int32_t nb_dof = ctx->nb_node * ctx->b;
  if (!ctx->has_elim && !ctx->has_forced) return AFB_OK;
  ctx->saved_valid = false;
  if (ctx->has_rc) {
    // keep the pre-elimination values for the RHS correction (replaces the host
    // OrderedRowColumnMap of CsrDoFLinearSystemImpl::_fillRowColumnEliminationInfos)
    size_t bytes = sizeof(double) * (size_t)ctx->nnz * ctx->b * ctx->b;
    AFB_TRY(ctx->saved_values.reserve(bytes));
    AFB_CUDA(cudaMemcpyAsync(ctx->saved_values.p, ctx->values.p, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->saved_valid = true;
  }
  k_matrix_transformation<<<grid_for(nb_dof, 128), 128, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->b, ctx->layout, nb_dof,
                                                                           ctx->has_elim ? ctx->elim_info.as<uint8_t>() : nullptr,
                                                                           ctx->has_forced ? ctx->forced_info.as<uint8_t>() : nullptr,
                                                                           ctx->has_forced ? ctx->forced_value.as<double>() : nullptr,
                                                                           ctx->has_rc ? 1 : 0, quirk, ctx->values.as<double>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

int apply_rhs_transformation(afb_ctx* ctx)
{
  if (!ctx->has_elim) return AFB_OK;
  const int32_t nb_dof = ctx->nb_node * ctx->b;
  AFB_REQUIRE(!ctx->has_rc || ctx->saved_valid, AFB_ERR_INVALID, "applyRHSTransformation with row-column elimination needs applyMatrixTransformation first");
  k_rhs_transformation<<<grid_for(nb_dof, 128), 128, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->b, ctx->layout, nb_dof,
                                                                        ctx->elim_info.as<uint8_t>(), ctx->elim_value.as<double>(),
                                                                        ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>(),
                                                                        ctx->has_rc ? ctx->saved_values.as<double>() : nullptr, ctx->rhs.as<double>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// ---------------------------------------------------------------------------------------------
// hand-off views
// ---------------------------------------------------------------------------------------------
// _translateCSRToCOO (femutils/CsrFormatMatrix.cc:161-184); one warp per row
__global__ void __launch_bounds__(256) k_coo_rows(const int32_t* __restrict__ rows, int32_t nb_row, int32_t* __restrict__ coo_rows)
{
  const int lane = threadIdx.x & 31;
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nb_row) return;
  for (int p = rows[r] + lane; p < rows[r + 1]; p += 32) coo_rows[p] = (int32_t)r;
}

int ensure_coo_rows(afb_ctx* ctx)
{
  if (ctx->coo_rows_valid) return AFB_OK;
  AFB_TRY(ctx->coo_rows.reserve(sizeof(int32_t) * (size_t)ctx->nnz));
  if (ctx->nb_node > 0) {
    k_coo_rows<<<grid_for((int64_t)ctx->nb_node * 32, 256), 256, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->nb_node, ctx->coo_rows.as<int32_t>());
    AFB_LAUNCH_CHECK(ctx);
  }
  ctx->coo_rows_valid = true;
  return AFB_OK;
}

// BSRMatrix::toCsr for b > 1 (femutils/BSRFormat.cc:127-162), on the device; one warp per block row
__global__ void __launch_bounds__(256) k_bsr_to_csr(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int b, int32_t nb_block_row, int64_t nnz,
                                                     int32_t* __restrict__ csr_rows, int32_t* __restrict__ csr_cols, int32_t* __restrict__ csr_nbcol)
{
  const int lane = threadIdx.x & 31;
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nb_block_row) return;
  const int rb = rows[r], nz = rows[r + 1] - rb;
  for (int i = 0; i < b; ++i) {
    const int64_t start = (int64_t)rb * b * b + (int64_t)i * nz * b;
    if (lane == 0) {
      csr_rows[r * b + i] = (int32_t)start;
      csr_nbcol[r * b + i] = nz * b;
    }
    for (int t = lane; t < nz * b; t += 32) csr_cols[start + t] = cols[rb + t / b] * b + (t % b);
  }
  if (r == nb_block_row - 1 && lane == 0) csr_rows[(int64_t)nb_block_row * b] = (int32_t)(nnz * b * b);
}

int ensure_scalar_csr(afb_ctx* ctx)
{
  if (ctx->csr_valid) return AFB_OK;
  const int b = ctx->b;
  AFB_TRY(ctx->csr_rows.reserve(sizeof(int32_t) * ((size_t)ctx->nb_node * b + 1)));
  AFB_TRY(ctx->csr_cols.reserve(sizeof(int32_t) * (size_t)ctx->nnz * b * b));
  AFB_TRY(ctx->csr_nbcol.reserve(sizeof(int32_t) * ((size_t)ctx->nb_node * b + 1)));
  if (ctx->nb_node > 0) {
    k_bsr_to_csr<<<grid_for((int64_t)ctx->nb_node * 32, 256), 256, 0, ctx->stream>>>(ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), b, ctx->nb_node, ctx->nnz,
                                                                                     ctx->csr_rows.as<int32_t>(), ctx->csr_cols.as<int32_t>(), ctx->csr_nbcol.as<int32_t>());
    AFB_LAUNCH_CHECK(ctx);
  }
  ctx->csr_valid = true;
  return AFB_OK;
}

} // namespace afb
