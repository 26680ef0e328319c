// Synthetic structured-box meshes of SURVEY.md §8(d), generated directly in HBM.
// Bit-identical to arcanefem_b200/mesh.py::box_mesh (same hash, same rounding sequence).
#include "afb_internal.h"

namespace afb {

__device__ __forceinline__ double jitter_unit(uint32_t id, uint32_t comp, uint32_t seed)
{
  uint32_t h = id * 3u + comp;
  h *= 0x9E3779B1u;
  h ^= seed;
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return __dadd_rn(__dmul_rn((double)h, 1.0 / 4294967296.0), -0.5);
}

struct BoxDesc {
  int dim, n, m;       // m = n+1
  int k_lo, k_hi;      // own cube layers of the last axis
  int k_top;           // node planes k_lo..k_top are local (k_top = k_hi, or k_hi+1 with a ghost cell layer)
  int32_t plane;       // nodes per layer of the last axis: m^(dim-1)
  int32_t nb_own, nb_node;
  double jitter;
  uint32_t seed;
};

// local id of the node in plane `k` (last axis) with in-plane offset `inl`: owned planes first
// (ascending), then the bottom ghost plane (owner: lower neighbour), then the top ghost plane
// (owner: upper neighbour)
__device__ __forceinline__ int32_t local_node(const BoxDesc& d, int k, int32_t inl)
{
  if (k > d.k_hi) return d.nb_own + (d.k_lo > 0 ? d.plane : 0) + inl;
  if (d.k_lo > 0) return (k == d.k_lo) ? d.nb_own + inl : (int32_t)(k - d.k_lo - 1) * d.plane + inl;
  return (int32_t)k * d.plane + inl;
}

__global__ void __launch_bounds__(256) k_gen_nodes(BoxDesc d, double* __restrict__ xyz, uint8_t* __restrict__ is_own)
{
  int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.nb_node) return;
  int k = d.k_lo + t / d.plane;
  int32_t inl = t % d.plane;
  int i = inl % d.m, j = (d.dim == 3) ? inl / d.m : k;
  int kk = (d.dim == 3) ? k : 0;
  const uint32_t gid = (d.dim == 3) ? (uint32_t)(i + d.m * (j + d.m * kk)) : (uint32_t)(i + d.m * j);
  bool interior = i > 0 && i < d.n && j > 0 && j < d.n && (d.dim == 2 || (kk > 0 && kk < d.n));
  const int idx[3] = { i, j, kk };
  double c[3] = { 0.0, 0.0, 0.0 };
  for (int a = 0; a < d.dim; ++a) {
    double off = interior ? __dmul_rn(d.jitter, jitter_unit(gid, (uint32_t)a, d.seed)) : 0.0;
    c[a] = __ddiv_rn(__dadd_rn((double)idx[a], off), (double)d.n);
  }
  const int32_t lid = local_node(d, k, inl);
  xyz[3 * (int64_t)lid] = c[0];
  xyz[3 * (int64_t)lid + 1] = c[1];
  xyz[3 * (int64_t)lid + 2] = c[2];
  if (is_own) is_own[lid] = ((d.k_lo > 0 && k == d.k_lo) || k > d.k_hi) ? 0 : 1;
}

// 3-D: one thread per cube -> 6 Kuhn tets {v0, v0+e_p1, v0+e_p1+e_p2, v0+e_p1+e_p2+e_p3}
__global__ void __launch_bounds__(256) k_gen_tets(BoxDesc d, int64_t nb_cube, int32_t* __restrict__ conn)
{
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb_cube) return;
  const int n = d.n, m = d.m;
  int ci = (int)(q % n), cj = (int)((q / n) % n), ck = d.k_lo + (int)(q / ((int64_t)n * n));
  const int perms[6][3] = { { 0, 1, 2 }, { 0, 2, 1 }, { 1, 0, 2 }, { 1, 2, 0 }, { 2, 0, 1 }, { 2, 1, 0 } };
  int4* out = reinterpret_cast<int4*>(conn) + q * 6;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    int pos[3] = { ci, cj, ck };
    int32_t v[4];
    v[0] = local_node(d, pos[2], pos[0] + m * pos[1]);
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      pos[perms[p][s]] += 1;
      v[s + 1] = local_node(d, pos[2], pos[0] + m * pos[1]);
    }
    out[p] = make_int4(v[0], v[1], v[2], v[3]);
  }
}

// 2-D: one thread per square -> 2 CCW triangles along the (0,0)-(1,1) diagonal
__global__ void __launch_bounds__(256) k_gen_tris(BoxDesc d, int64_t nb_sq, int32_t* __restrict__ conn)
{
  int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nb_sq) return;
  int ci = (int)(q % d.n), cj = d.k_lo + (int)(q / d.n);
  int32_t v00 = local_node(d, cj, ci), v10 = local_node(d, cj, ci + 1), v11 = local_node(d, cj + 1, ci + 1), v01 = local_node(d, cj + 1, ci);
  int32_t* o = conn + q * 6;
  o[0] = v00; o[1] = v10; o[2] = v11;
  o[3] = v00; o[4] = v11; o[5] = v01;
}

int generate_box(afb_ctx* ctx, int dim, int n, double jitter, uint32_t seed, int k_lo, int k_hi, int ghost_cell_layer)
{
  AFB_REQUIRE(dim == 2 || dim == 3, AFB_ERR_INVALID, "box dimension must be 2 or 3");
  AFB_REQUIRE(n >= 1 && k_lo >= 0 && k_hi <= n && k_lo < k_hi, AFB_ERR_INVALID, "bad box slab n=%d layers [%d,%d)", n, k_lo, k_hi);
  const bool ghost = ghost_cell_layer && k_hi < n;
  BoxDesc d;
  d.dim = dim; d.n = n; d.m = n + 1; d.k_lo = k_lo; d.k_hi = k_hi; d.k_top = ghost ? k_hi + 1 : k_hi;
  d.jitter = jitter; d.seed = seed;
  int64_t plane = dim == 3 ? (int64_t)d.m * d.m : d.m;
  int64_t nb_node = plane * (d.k_top - k_lo + 1);
  // the ghost layer's cubes directly follow the own ones: same kernel, cube layers [k_lo, k_hi + ghost)
  int64_t nb_cube_own = (dim == 3 ? (int64_t)n * n : n) * (int64_t)(k_hi - k_lo);
  int64_t nb_cube = (dim == 3 ? (int64_t)n * n : n) * (int64_t)(k_hi - k_lo + (ghost ? 1 : 0));
  int64_t nb_cell = nb_cube * (dim == 3 ? 6 : 2);
  AFB_REQUIRE(nb_node < 2147483647LL && (int64_t)(dim + 1) * nb_cell < 2147483647LL, AFB_ERR_OVERFLOW, "box too large for Int32 ids");
  d.plane = (int32_t)plane;
  d.nb_node = (int32_t)nb_node;
  d.nb_own = (int32_t)(nb_node - (k_lo > 0 ? plane : 0) - (ghost ? plane : 0));
  const int npc = dim + 1;
  AFB_TRY(ctx->coords.reserve(sizeof(double) * 3 * (size_t)nb_node));
  AFB_TRY(ctx->conn.reserve(sizeof(int32_t) * (size_t)npc * (size_t)nb_cell));
  AFB_TRY(ctx->is_own.reserve((size_t)nb_node));
  k_gen_nodes<<<grid_for(nb_node, 256), 256, 0, ctx->stream>>>(d, ctx->coords.as<double>(), ctx->is_own.as<uint8_t>());
  AFB_LAUNCH_CHECK(ctx);
  if (dim == 3) k_gen_tets<<<grid_for(nb_cube, 256), 256, 0, ctx->stream>>>(d, nb_cube, ctx->conn.as<int32_t>());
  else k_gen_tris<<<grid_for(nb_cube, 256), 256, 0, ctx->stream>>>(d, nb_cube, ctx->conn.as<int32_t>());
  AFB_LAUNCH_CHECK(ctx);
  ctx->dim = dim;
  ctx->npc = npc;
  ctx->nb_node = (int32_t)nb_node;
  ctx->nb_own_node = d.nb_own;
  ctx->nb_cell = nb_cell;
  ctx->nb_own_cell = nb_cube_own * (dim == 3 ? 6 : 2);
  ctx->all_own = (k_lo == 0 && !ghost);
  ctx->has_mesh = true;
  return AFB_OK;
}

} // namespace afb
