// Tiled gather assembly: the B200 path of the atomic-free ("node-wise") back-ends.
//
// Reference behaviour replaced: _assembleNodeWiseCsrBilinearOperator{Tria3,Tetra4}
// (modules/testlab/NodeWiseCsrBiliAssembly.cc:157-297) and BSRFormat::assembleBilinearAtomicFree
// (femutils/BSRFormat.h:406-577): every matrix row is written by exactly one owner, without
// atomics.  The reference does it with one thread per node that recomputes the geometry of
// every incident cell (4x redundant fp64 work on tetrahedra, valence-divergent).  Here:
//
//   inspector (once per mesh, build_tile_plan)
//     nodes are binned into spatial bricks (tiles of <= RMAX rows); each tile gets the list of
//     cells touching it, and each matrix entry (row in tile, column) gets the list of
//     (cell, local pair) contributions as 16-bit indices into a per-tile element-matrix cache.
//     Entries are sorted by contribution count and cut in units of 32 (one warp), so a warp
//     walks 32 equally long lists.
//   executor (every assembly, k_assemble_tiled) -- one persistent CTA per SM, per tile:
//     phase A  one thread per tile cell: geometry once, the 10 (Tet4) / 6 (Tri3) distinct
//              K_e values go to the shared-memory cache (structure of arrays, conflict-free);
//     phase B  one thread per entry: sum the cached contributions in a fixed order (ascending
//              cell id => bit-reproducible) and store the value once.
//   Rows are written exactly once, so no zero fill is needed for the rows a tile owns.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "element.cuh"

namespace afb {

constexpr int TG_THREADS = 1024;           // executor CTA
constexpr int TG_CMAX = 2304;              // cells per tile
constexpr int TG_CS = TG_CMAX + 1;         // cache stride (odd: consecutive pair planes shift banks)
constexpr int TG_KMAX = 10;                // distinct K_e values of a symmetric 4x4
constexpr int TG_ZERO = TG_KMAX * TG_CS;   // cache slot that holds 0.0 (list padding)
constexpr int TG_EMAX = 5120;              // entries per tile
constexpr int TG_RALLOC = 1024;            // rows per tile, allocation bound
constexpr int TB_THREADS = 512;            // builder CTA
constexpr unsigned TG_NONE = 0xFFFFFFFFu;

struct TileDesc {
  int32_t node_off, nb_row, cell_off, nb_cell, unit_off, nb_unit;
  uint32_t list_off;
  int32_t nb_entry;
};

// index of the symmetric pair (a,b) of a 4-node cell in the cache: 00 01 02 03 11 12 13 22 23 33
__host__ __device__ __forceinline__ constexpr int sym_pair(int a, int b)
{
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * 4 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// ---------------------------------------------------------------------------------------------
// inspector step 1: spatial bricks
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_f64(double x)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double unorder_f64(unsigned long long u)
{
  u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
  double x;
  memcpy(&x, &u, sizeof(x));
  return x;
}

__global__ void __launch_bounds__(256) k_bbox(const double* __restrict__ coords, int32_t nb_node, unsigned long long* __restrict__ box /* min xyz, max xyz */)
{
  double mn[3] = { 1e300, 1e300, 1e300 }, mx[3] = { -1e300, -1e300, -1e300 };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nb_node; i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = coords[3 * i + a];
      mn[a] = fmin(mn[a], v);
      mx[a] = fmax(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], d));
      mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], d));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(box + a, order_f64(mn[a]));
      atomicMax(box + 3 + a, order_f64(mx[a]));
    }
  }
}

struct BrickGrid {
  double x0[3], inv_h[3];
  int g[3];
};

__global__ void __launch_bounds__(256) k_brick_assign(const double* __restrict__ coords, const uint8_t* __restrict__ is_own, int32_t nb_node, BrickGrid bg,
                                                       int32_t* __restrict__ brick_of, int32_t* __restrict__ count)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_node) return;
  if (is_own && !is_own[i]) {
    brick_of[i] = -1;
    return;
  }
  int id[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    int v = (int)((coords[3 * (int64_t)i + a] - bg.x0[a]) * bg.inv_h[a]);
    id[a] = min(max(v, 0), bg.g[a] - 1);
  }
  const int32_t b = id[0] + bg.g[0] * (id[1] + bg.g[1] * id[2]);
  brick_of[i] = b;
  atomicAdd(count + b, 1);
}

__global__ void __launch_bounds__(256) k_brick_fill(const int32_t* __restrict__ brick_of, int32_t nb_node, const int32_t* __restrict__ brick_ptr,
                                                     int32_t* __restrict__ cursor, int32_t* __restrict__ tnodes)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_node) return;
  const int32_t b = brick_of[i];
  if (b < 0) return;
  tnodes[brick_ptr[b] + atomicAdd(cursor + b, 1)] = i;
}

// ascending node ids inside each brick (the atomic fill order is arbitrary); rank sort in shared
// memory for bricks of <= 2048 nodes, larger ones keep the fill order (correct, just less local)
__global__ void __launch_bounds__(256) k_brick_sort(const int32_t* __restrict__ brick_ptr, int32_t nb_brick, int32_t* __restrict__ tnodes)
{
  __shared__ int32_t s[2048];
  const int b = blockIdx.x;
  if (b >= nb_brick) return;
  const int beg = brick_ptr[b], n = brick_ptr[b + 1] - beg;
  if (n <= 1 || n > 2048) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = tnodes[beg + i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t x = s[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += s[j] < x ? 1 : 0;
    tnodes[beg + rank] = x;
  }
}

// tiles of a brick: ceil(cnt / rmax) equal pieces
__global__ void __launch_bounds__(256) k_brick_tiles(const int32_t* __restrict__ brick_ptr, int32_t nb_brick, int rmax, int32_t* __restrict__ ntile_of)
{
  const int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb_brick) return;
  const int n = brick_ptr[b + 1] - brick_ptr[b];
  ntile_of[b] = (n + rmax - 1) / rmax;
}

__global__ void __launch_bounds__(256) k_tile_nodes(const int32_t* __restrict__ brick_of, const int32_t* __restrict__ brick_ptr, const int32_t* __restrict__ tile_first,
                                                     const int32_t* __restrict__ ntile_of, const int32_t* __restrict__ tnodes, int32_t nb_tnode,
                                                     int32_t* __restrict__ node_tile, int32_t* __restrict__ node_lrow, TileDesc* __restrict__ desc)
{
  const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nb_tnode) return;
  const int32_t node = tnodes[p];
  const int32_t b = brick_of[node];
  const int beg = brick_ptr[b], n = brick_ptr[b + 1] - beg;
  const int nt = ntile_of[b];
  const int chunk = (n + nt - 1) / nt;
  const int j = p - beg;
  const int s = j / chunk;
  const int lrow = j - s * chunk;
  const int32_t t = tile_first[b] + s;
  node_tile[node] = t;
  node_lrow[node] = lrow;
  if (lrow == 0) {
    desc[t].node_off = p;
    desc[t].nb_row = min(chunk, n - s * chunk);
  }
}

// ---------------------------------------------------------------------------------------------
// leader test: the incidence (row i of tile t, cell) owns the cell inside the tile iff no other
// node of the cell is a row of the same tile with a smaller row index
// ---------------------------------------------------------------------------------------------
template <int NPC>
__device__ __forceinline__ bool is_leader(const int32_t* __restrict__ conn, const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow, int32_t t, int i,
                                          int32_t r, int32_t cell)
{
  const int32_t* cn = conn + (int64_t)cell * NPC;
  bool lead = true;
#pragma unroll
  for (int a = 0; a < NPC; ++a) {
    const int32_t n = __ldg(cn + a);
    if (n != r && __ldg(node_tile + n) == t && __ldg(node_lrow + n) < i) lead = false;
  }
  return lead;
}

// per tile: number of cells, entries, largest valence
template <int NPC>
__global__ void __launch_bounds__(128) k_tile_stats(const TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ conn,
                                                     const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, const int32_t* __restrict__ rows,
                                                     const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow, int32_t* __restrict__ stats /* [nb_tile][4] */)
{
  __shared__ int s_c, s_e, s_v;
  const int32_t t = blockIdx.x;
  if (t >= nb_tile) return;
  if (threadIdx.x == 0) s_c = s_e = s_v = 0;
  __syncthreads();
  const TileDesc d = desc[t];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int c = 0, e = 0, v = 0;
  for (int i = warp; i < d.nb_row; i += 4) {
    const int32_t r = tnodes[d.node_off + i];
    const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
    if (lane == 0) {
      e += rows[r + 1] - rows[r];
      v = max(v, qe - qb);
    }
    for (int q = qb + lane; q < qe; q += 32)
      if (is_leader<NPC>(conn, node_tile, node_lrow, t, i, r, nc_list[q])) ++c;
  }
  atomicAdd(&s_c, c);
  atomicAdd(&s_e, e);
  atomicMax(&s_v, v);
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[4 * t + 0] = s_c;
    stats[4 * t + 1] = s_e;
    stats[4 * t + 2] = s_v;
    stats[4 * t + 3] = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// shared-memory helpers of the builder
// ---------------------------------------------------------------------------------------------
// ascending bitonic sort of n2 (power of two) 32-bit keys in shared memory
__device__ void smem_bitonic_sort(unsigned* s, int n2)
{
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            s[i] = b;
            s[ixj] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// in-place exclusive scan of n ints in shared memory (n <= a few thousand); returns the total
__device__ int smem_exclusive_scan(int* s, int n, int* s_tmp /* >= blockDim.x/32 + 1 ints */)
{
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int beg = min((int)threadIdx.x * per, n), end = min(beg + per, n);
  int sum = 0;
  for (int i = beg; i < end; ++i) sum += s[i];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    int w = lane < nw ? s_tmp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    if (lane < nw) s_tmp[lane] = winc - w;
    if (lane == 31) s_tmp[32] = winc;
  }
  __syncthreads();
  int run = s_tmp[warp] + inc - sum;
  const int total = s_tmp[32];
  for (int i = beg; i < end; ++i) {
    const int v = s[i];
    s[i] = run;
    run += v;
  }
  __syncthreads();
  return total;
}

// ---------------------------------------------------------------------------------------------
// inspector step 2: per-tile plan
// ---------------------------------------------------------------------------------------------
struct BuilderSmem {
  unsigned cells[4096];          // tile cells (sorted ascending), padded to a power of two
  int erow_off[TG_RALLOC + 1];   // first entry of each tile row
  int cnt[TG_EMAX];              // contributions per entry
  int eoff[TG_EMAX + 1];         // start of each entry's list in clist
  int cur[TG_EMAX];              // fill cursors
  unsigned egpos[TG_EMAX];       // entry -> index into values
  unsigned keys[8192];           // entries sorted by count (descending)
  uint16_t clist[16 * TG_CMAX];  // contribution codes
  int ulen[TG_EMAX / 32 + 1], ubase[TG_EMAX / 32 + 2];
  int tmp[40];
  int nb_cell;
};

template <int NPC>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_tile_build(TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ conn, const int32_t* __restrict__ nc_ptr,
             const int32_t* __restrict__ nc_list, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ node_tile,
             const int32_t* __restrict__ node_lrow, int32_t* __restrict__ tile_cells, uint32_t* __restrict__ unit_base, uint16_t* __restrict__ unit_len,
             uint32_t* __restrict__ gpos, uint16_t* __restrict__ lists, int* __restrict__ error)
{
  extern __shared__ unsigned char tb_raw[];
  BuilderSmem& S = *reinterpret_cast<BuilderSmem*>(tb_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    const int R = d.nb_row;
    // ---- leader cells -> S.cells (then sorted ascending) ----
    if (threadIdx.x == 0) S.nb_cell = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) S.cells[i] = 0xFFFFFFFFu;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) {
      int deg = 0;
      if (i < R) {
        const int32_t r = tnodes[d.node_off + i];
        deg = rows[r + 1] - rows[r];
      }
      S.erow_off[i] = deg;
    }
    __syncthreads();
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
      for (int q = qb + lane; q < qe; q += 32) {
        const int32_t c = nc_list[q];
        if (is_leader<NPC>(conn, node_tile, node_lrow, t, i, r, c)) {
          const int pos = atomicAdd(&S.nb_cell, 1);
          if (pos < 4096) S.cells[pos] = (unsigned)c;
        }
      }
    }
    __syncthreads();
    const int C = S.nb_cell;
    const int E = smem_exclusive_scan(S.erow_off, R + 1, S.tmp);
    if (C != d.nb_cell || E != d.nb_entry || C > TG_CMAX || E > TG_EMAX) {
      if (threadIdx.x == 0) atomicExch(error, 1);
      __syncthreads();
      continue;
    }
    int c2 = 32;
    while (c2 < C) c2 <<= 1;
    smem_bitonic_sort(S.cells, c2);
    for (int i = threadIdx.x; i < C; i += blockDim.x) tile_cells[d.cell_off + i] = (int32_t)S.cells[i];
    // ---- entries: value positions, contribution counts ----
    for (int e = threadIdx.x; e < E; e += blockDim.x) S.cnt[e] = 0;
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int rb = rows[r], deg = rows[r + 1] - rb, e0 = S.erow_off[i];
      for (int p = lane; p < deg; p += 32) S.egpos[e0 + p] = (unsigned)(rb + p);
    }
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
      for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
        const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
        int32_t nd[NPC];
#pragma unroll
        for (int a = 0; a < NPC; ++a) nd[a] = __ldg(cn + a);
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          if (__ldg(node_tile + nd[a]) != t) continue;
          const int i = __ldg(node_lrow + nd[a]);
          const int rb = __ldg(rows + nd[a]), re = __ldg(rows + nd[a] + 1);
#pragma unroll
          for (int bq = 0; bq < NPC; ++bq) {
            const int e = S.erow_off[i] + (find_col(cols, rb, re, nd[bq]) - rb);
            if (pass == 0) atomicAdd(&S.cnt[e], 1);
            else {
              const int slot = atomicAdd(&S.cur[e], 1);
              S.clist[S.eoff[e] + slot] = (uint16_t)(sym_pair(a, bq) * TG_CS + lc);
            }
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        for (int e = threadIdx.x; e <= E; e += blockDim.x) {
          S.eoff[e] = e < E ? S.cnt[e] : 0;
          if (e < E) S.cur[e] = 0;
        }
        __syncthreads();
        smem_exclusive_scan(S.eoff, E + 1, S.tmp);
      }
    }
    // ---- fixed summation order: ascending local cell index (= ascending global cell id) ----
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      uint16_t* l = S.clist + S.eoff[e];
      const int n = S.cnt[e];
      for (int i = 1; i < n; ++i) {
        const uint16_t x = l[i];
        const unsigned kx = ((unsigned)(x % TG_CS) << 16) | x;
        int j = i - 1;
        while (j >= 0) {
          const uint16_t y = l[j];
          if ((((unsigned)(y % TG_CS) << 16) | y) <= kx) break;
          l[j + 1] = y;
          --j;
        }
        l[j + 1] = x;
      }
    }
    // ---- entries by descending count, cut into units of 32 ----
    int e2 = 32;
    while (e2 < E) e2 <<= 1;
    for (int e = threadIdx.x; e < e2; e += blockDim.x)
      S.keys[e] = e < E ? (((unsigned)(0xFFFF - min(S.cnt[e], 0xFFFF)) << 16) | (unsigned)e) : 0xFFFFFFFFu;
    __syncthreads();
    smem_bitonic_sort(S.keys, e2);
    const int nunit = (E + 31) / 32;
    for (int u = threadIdx.x; u <= nunit; u += blockDim.x) {
      int len = 0;
      if (u < nunit) len = S.cnt[S.keys[u * 32] & 0xFFFFu];
      S.ulen[u] = len;
      S.ubase[u] = ((len + 7) >> 3) * 256; // chunks of 8 indices x 32 lanes
    }
    __syncthreads();
    const int list_total = smem_exclusive_scan(S.ubase, nunit + 1, S.tmp);
    if (nunit != d.nb_unit) {
      if (threadIdx.x == 0) atomicExch(error, 2);
      __syncthreads();
      continue;
    }
    (void)list_total;
    for (int u = threadIdx.x; u < nunit; u += blockDim.x) {
      unit_base[d.unit_off + u] = d.list_off + (uint32_t)S.ubase[u];
      unit_len[d.unit_off + u] = (uint16_t)S.ulen[u];
    }
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const int u = x >> 5, l = x & 31;
      const bool valid = x < E;
      const int e = valid ? (int)(S.keys[x] & 0xFFFFu) : 0;
      gpos[(size_t)(d.unit_off + u) * 32 + l] = valid ? S.egpos[e] : TG_NONE;
      const int len = S.ulen[u], n = valid ? S.cnt[e] : 0;
      uint16_t* out = lists + d.list_off + S.ubase[u] + l * 8;
      const uint16_t* src = S.clist + (valid ? S.eoff[e] : 0);
      const int padded = ((len + 7) >> 3) * 8;
      for (int k = 0; k < padded; ++k) out[(k >> 3) * 256 + (k & 7)] = k < n ? src[k] : (uint16_t)TG_ZERO;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// executor
// ---------------------------------------------------------------------------------------------
template <int NPC> struct SymK;
template <> struct SymK<4> {
  static constexpr int N = 10;
  __device__ static __forceinline__ void compute(const double* __restrict__ coords, const int32_t* __restrict__ conn, int32_t cell, const ElemParams&, double (&K)[10])
  {
    const int4 v = __ldg(reinterpret_cast<const int4*>(conn) + cell);
    const int32_t nd[4] = { v.x, v.y, v.z, v.w };
    Tet4Geom g;
    g.init(coords, nd);
    int p = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) K[p++] = g.dot(a, b) * g.s;
  }
};
template <> struct SymK<3> {
  static constexpr int N = 10; // pair indexing of a 4-node cell is reused; slots with a or b == 3 stay unused
  __device__ static __forceinline__ void compute(const double* __restrict__ coords, const int32_t* __restrict__ conn, int32_t cell, const ElemParams& prm, double (&K)[10])
  {
    const int32_t* cn = conn + 3 * (int64_t)cell;
    const int32_t nd[3] = { __ldg(cn), __ldg(cn + 1), __ldg(cn + 2) };
    Tri3Geom g;
    g.init(coords, nd, (prm.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
#pragma unroll
    for (int p = 0; p < 10; ++p) K[p] = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) K[sym_pair(a, b)] = g.dot(a, b) * g.s;
  }
};

template <int NPC>
__global__ void __launch_bounds__(TG_THREADS, 1)
k_assemble_tiled(const TileDesc* __restrict__ desc, int32_t nb_tile, const double* __restrict__ coords, const int32_t* __restrict__ conn,
                 const int32_t* __restrict__ tile_cells, const uint32_t* __restrict__ unit_base, const uint16_t* __restrict__ unit_len,
                 const uint32_t* __restrict__ gpos, const uint16_t* __restrict__ lists, double* __restrict__ values, int accumulate, ElemParams prm)
{
  extern __shared__ double Kc[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TG_THREADS / 32;
  if (threadIdx.x == 0) Kc[TG_ZERO] = 0.0;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    // phase A: element matrices of the tile's cells, once each
    for (int lc = threadIdx.x; lc < d.nb_cell; lc += TG_THREADS) {
      double K[10];
      SymK<NPC>::compute(coords, conn, __ldg(tile_cells + d.cell_off + lc), prm, K);
#pragma unroll
      for (int p = 0; p < 10; ++p)
        if (NPC == 4 || (p != 3 && p != 6 && p != 8 && p != 9)) Kc[p * TG_CS + lc] = K[p];
    }
    __syncthreads();
    // phase B: one warp per unit of 32 entries with equally long contribution lists; each lane
    // pulls its entry's indices 8 at a time (one 128-bit load), the next chunk / next unit is
    // in flight while the current one is summed
    {
      int u = warp;
      uint4 cur = make_uint4(0, 0, 0, 0);
      uint32_t g = TG_NONE;
      if (u < d.nb_unit) {
        cur = __ldg(reinterpret_cast<const uint4*>(lists + __ldg(unit_base + d.unit_off + u)) + lane);
        g = __ldg(gpos + (size_t)(d.unit_off + u) * 32 + lane);
      }
      while (u < d.nb_unit) {
        const uint4* l = reinterpret_cast<const uint4*>(lists + __ldg(unit_base + d.unit_off + u)) + lane;
        const int len = __ldg(unit_len + d.unit_off + u);
        const int un = u + NW;
        uint4 nfirst = make_uint4(0, 0, 0, 0);
        uint32_t gn = TG_NONE;
        if (un < d.nb_unit) {
          nfirst = __ldg(reinterpret_cast<const uint4*>(lists + __ldg(unit_base + d.unit_off + un)) + lane);
          gn = __ldg(gpos + (size_t)(d.unit_off + un) * 32 + lane);
        }
        double acc0 = 0.0, acc1 = 0.0;
        for (int k0 = 0; k0 < len; k0 += 8) {
          uint4 nxt = make_uint4(0, 0, 0, 0);
          if (k0 + 8 < len) nxt = __ldg(l + ((k0 >> 3) + 1) * 32);
          const int rem = len - k0;
          acc0 += Kc[cur.x & 0xFFFFu];
          if (rem > 1) acc1 += Kc[cur.x >> 16];
          if (rem > 2) acc0 += Kc[cur.y & 0xFFFFu];
          if (rem > 3) acc1 += Kc[cur.y >> 16];
          if (rem > 4) acc0 += Kc[cur.z & 0xFFFFu];
          if (rem > 5) acc1 += Kc[cur.z >> 16];
          if (rem > 6) acc0 += Kc[cur.w & 0xFFFFu];
          if (rem > 7) acc1 += Kc[cur.w >> 16];
          cur = nxt;
        }
        if (g != TG_NONE) {
          const double v = acc0 + acc1;
          if (accumulate) values[g] += v;
          else values[g] = v;
        }
        u = un;
        cur = nfirst;
        g = gn;
      }
    }
    __syncthreads();
  }
}

// values of rows that no tile owns (non-owned nodes) must read as zero after a fresh assembly
__global__ void __launch_bounds__(256) k_zero_unowned_rows(const int32_t* __restrict__ rows, const int32_t* __restrict__ node_tile, int32_t nb_node, int bb, double* __restrict__ values)
{
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= nb_node || node_tile[r] >= 0) return;
  for (int64_t p = (int64_t)rows[r] * bb + lane; p < (int64_t)rows[r + 1] * bb; p += 32) values[p] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool tiled_supported(const afb_ctx* ctx) { return ctx->b == 1 && (ctx->npc == 3 || ctx->npc == 4); }

int build_tile_plan(afb_ctx* ctx)
{
  AFB_REQUIRE(tiled_supported(ctx), AFB_ERR_UNSUPPORTED,
              "AFB_VARIANT_TILED_GATHER is not available for %d-node cells with %d dof per node (P1 scalar operators only); use AFB_VARIANT_NODEWISE", ctx->npc, ctx->b);
  TilePlan& P = ctx->plan;
  P.valid = false;
  cudaStream_t st = ctx->stream;
  const int32_t nb_node = ctx->nb_node;
  const int dim = ctx->dim, npc = ctx->npc;
  cudaEvent_t e0, e1;
  AFB_CUDA(cudaEventCreate(&e0));
  AFB_CUDA(cudaEventCreate(&e1));
  AFB_CUDA(cudaEventRecord(e0, st));

  // bounding box
  AFB_TRY(P.stats.reserve(sizeof(unsigned long long) * 8));
  unsigned long long init[6] = { ~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull }, got[6];
  AFB_CUDA(cudaMemcpyAsync(P.stats.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_bbox<<<std::min(grid_for(nb_node, 256), 4 * ctx->sm_count), 256, 0, st>>>(ctx->coords.as<double>(), nb_node, P.stats.as<unsigned long long>());
  AFB_LAUNCH_CHECK(ctx);
  AFB_CUDA(cudaMemcpyAsync(got, P.stats.p, sizeof(got), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  double lo[3], ext[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = unorder_f64(got[a]);
    ext[a] = unorder_f64(got[3 + a]) - lo[a];
    if (!(ext[a] > 0.0) || a >= dim) ext[a] = 0.0;
  }
  AFB_TRY(P.node_tile.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.node_lrow.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.tile_nodes.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.scratch_c.reserve(sizeof(int32_t) * (size_t)nb_node)); // brick_of
  int32_t* brick_of = P.scratch_c.as<int32_t>();

  // rows per tile: start from what the cache can hold on a regular mesh, halve until every tile fits
  int rtarget = dim == 3 ? 216 : 640;
  std::vector<TileDesc> hdesc;
  std::vector<int32_t> hstats;
  int32_t nb_tile = 0;
  for (int attempt = 0;; ++attempt) {
    AFB_REQUIRE(rtarget >= 1, AFB_ERR_UNSUPPORTED, "tiled gather: a single row exceeds the tile limits (%d cells / %d entries); use AFB_VARIANT_NODEWISE", TG_CMAX, TG_EMAX);
    const int rmax = std::min(TG_RALLOC, rtarget + rtarget / 2);
    // brick edge so that a brick holds ~rtarget nodes on a uniform mesh
    double vol = 1.0;
    int nd_ext = 0;
    for (int a = 0; a < 3; ++a)
      if (ext[a] > 0.0) { vol *= ext[a]; ++nd_ext; }
    BrickGrid bg;
    double h = nd_ext ? pow(vol * (double)rtarget / (double)nb_node, 1.0 / nd_ext) : 1.0;
    int64_t nb_brick = 1;
    for (int a = 0; a < 3; ++a) {
      bg.x0[a] = lo[a];
      bg.g[a] = ext[a] > 0.0 ? std::max(1, (int)ceil(ext[a] / h)) : 1;
      bg.inv_h[a] = ext[a] > 0.0 ? (double)bg.g[a] / ext[a] : 0.0;
      nb_brick *= bg.g[a];
    }
    AFB_REQUIRE(nb_brick < (1ll << 30), AFB_ERR_UNSUPPORTED, "tiled gather: brick grid too large");
    AFB_TRY(P.scratch_a.reserve(sizeof(int32_t) * (size_t)(4 * (nb_brick + 2))));
    int32_t* bcount = P.scratch_a.as<int32_t>();
    int32_t* bptr = bcount + (nb_brick + 2);
    int32_t* bntile = bptr + (nb_brick + 2);
    int32_t* bfirst = bntile + (nb_brick + 2);
    AFB_CUDA(cudaMemsetAsync(bcount, 0, sizeof(int32_t) * (size_t)(nb_brick + 2), st));
    k_brick_assign<<<grid_for(nb_node, 256), 256, 0, st>>>(ctx->coords.as<double>(), ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>(), nb_node, bg, brick_of, bcount);
    AFB_LAUNCH_CHECK(ctx);
    AFB_TRY(exclusive_scan_i32(ctx, bcount, bptr, nb_brick));
    AFB_CUDA(cudaMemsetAsync(bcount, 0, sizeof(int32_t) * (size_t)(nb_brick + 2), st));
    k_brick_fill<<<grid_for(nb_node, 256), 256, 0, st>>>(brick_of, nb_node, bptr, bcount, P.tile_nodes.as<int32_t>());
    AFB_LAUNCH_CHECK(ctx);
    k_brick_sort<<<(int)nb_brick, 256, 0, st>>>(bptr, (int32_t)nb_brick, P.tile_nodes.as<int32_t>());
    AFB_LAUNCH_CHECK(ctx);
    k_brick_tiles<<<grid_for(nb_brick, 256), 256, 0, st>>>(bptr, (int32_t)nb_brick, rmax, bntile);
    AFB_LAUNCH_CHECK(ctx);
    AFB_TRY(exclusive_scan_i32(ctx, bntile, bfirst, nb_brick));
    int32_t nb_tnode = 0;
    AFB_CUDA(cudaMemcpyAsync(&nb_tile, bfirst + nb_brick, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaMemcpyAsync(&nb_tnode, bptr + nb_brick, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    AFB_TRY(P.tile_desc.reserve(sizeof(TileDesc) * (size_t)std::max(nb_tile, 1)));
    AFB_TRY(P.scratch_b.reserve(sizeof(int32_t) * 4 * (size_t)std::max(nb_tile, 1)));
    AFB_CUDA(cudaMemsetAsync(P.node_tile.p, 0xFF, sizeof(int32_t) * (size_t)nb_node, st));
    AFB_CUDA(cudaMemsetAsync(P.tile_desc.p, 0, sizeof(TileDesc) * (size_t)std::max(nb_tile, 1), st));
    if (nb_tile == 0) break;
    k_tile_nodes<<<grid_for(nb_tnode, 256), 256, 0, st>>>(brick_of, bptr, bfirst, bntile, P.tile_nodes.as<int32_t>(), nb_tnode, P.node_tile.as<int32_t>(),
                                                           P.node_lrow.as<int32_t>(), P.tile_desc.as<TileDesc>());
    AFB_LAUNCH_CHECK(ctx);
    int32_t* stats = P.scratch_b.as<int32_t>();
    if (npc == 4)
      k_tile_stats<4><<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                               ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), P.node_tile.as<int32_t>(), P.node_lrow.as<int32_t>(), stats);
    else
      k_tile_stats<3><<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                               ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), P.node_tile.as<int32_t>(), P.node_lrow.as<int32_t>(), stats);
    AFB_LAUNCH_CHECK(ctx);
    hdesc.resize(nb_tile);
    hstats.resize(4 * (size_t)nb_tile);
    AFB_CUDA(cudaMemcpyAsync(hdesc.data(), P.tile_desc.p, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaMemcpyAsync(hstats.data(), stats, sizeof(int32_t) * 4 * (size_t)nb_tile, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    bool ok = true;
    for (int32_t t = 0; t < nb_tile && ok; ++t)
      if (hstats[4 * t] > TG_CMAX || hstats[4 * t + 1] > TG_EMAX || hdesc[t].nb_row > TG_RALLOC) ok = false;
    if (ok) break;
    rtarget /= 2;
  }

  // sizes -> offsets (host; a few thousand tiles)
  int64_t cell_off = 0, unit_off = 0, list_off = 0;
  for (int32_t t = 0; t < nb_tile; ++t) {
    TileDesc& d = hdesc[t];
    const int C = hstats[4 * t], E = hstats[4 * t + 1], V = hstats[4 * t + 2];
    d.cell_off = (int32_t)cell_off;
    d.nb_cell = C;
    d.unit_off = (int32_t)unit_off;
    d.nb_unit = (E + 31) / 32;
    d.nb_entry = E;
    d.list_off = (uint32_t)list_off;
    cell_off += C;
    unit_off += d.nb_unit;
    // capacity of the padded lists: sum over units of 32*8*ceil(len/8) <= contributions + 32*max count + 224 per unit
    int64_t cap = (int64_t)npc * npc * C + 32ll * (V + 2) + 224ll * d.nb_unit;
    cap = (cap + 7) & ~7ll;
    list_off += cap;
    AFB_REQUIRE(list_off < (1ll << 32) && cell_off < (1ll << 31), AFB_ERR_OVERFLOW, "tiled gather plan exceeds 32-bit offsets");
  }
  P.nb_tile = nb_tile;
  P.nb_tile_cell = cell_off;
  P.nb_unit = unit_off;
  P.nb_list = list_off;
  AFB_TRY(P.tile_cells.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(cell_off, 1)));
  AFB_TRY(P.unit_base.reserve(sizeof(uint32_t) * (size_t)std::max<int64_t>(unit_off, 1)));
  AFB_TRY(P.unit_len.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(unit_off, 1)));
  AFB_TRY(P.gpos.reserve(sizeof(uint32_t) * 32 * (size_t)std::max<int64_t>(unit_off, 1)));
  AFB_TRY(P.lists.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(list_off, 8)));
  if (nb_tile > 0) {
    AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc.data(), sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
    AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
    AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
    const size_t smem = sizeof(BuilderSmem);
    const int grid = std::min<int>(nb_tile, 2 * ctx->sm_count);
    if (npc == 4) {
      AFB_CUDA(cudaFuncSetAttribute(k_tile_build<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tile_build<4><<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                                      ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(),
                                                      P.node_lrow.as<int32_t>(), P.tile_cells.as<int32_t>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(),
                                                      P.gpos.as<uint32_t>(), P.lists.as<uint16_t>(), ctx->tmp_flag.as<int>());
    }
    else {
      AFB_CUDA(cudaFuncSetAttribute(k_tile_build<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tile_build<3><<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                                      ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(),
                                                      P.node_lrow.as<int32_t>(), P.tile_cells.as<int32_t>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(),
                                                      P.gpos.as<uint32_t>(), P.lists.as<uint16_t>(), ctx->tmp_flag.as<int>());
    }
    AFB_LAUNCH_CHECK(ctx);
    int err = 0;
    AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    AFB_REQUIRE(err == 0, AFB_ERR_CUDA, "tiled gather: plan builder inconsistency (code %d)", err);
  }
  AFB_CUDA(cudaEventRecord(e1, st));
  AFB_CUDA(cudaEventSynchronize(e1));
  AFB_CUDA(cudaEventElapsedTime(&P.build_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  P.mesh_gen = ctx->mesh_gen;
  P.b = ctx->b;
  P.valid = true;
  return AFB_OK;
}

int assemble_tiled(afb_ctx* ctx, int op, const double* params, int layout, int flags)
{
  (void)layout;
  AFB_REQUIRE(op == AFB_OP_POISSON && tiled_supported(ctx), AFB_ERR_UNSUPPORTED,
              "AFB_VARIANT_TILED_GATHER is not available for operator %d on %d-node cells (P1 Poisson only); use AFB_VARIANT_NODEWISE", op, ctx->npc);
  const TilePlan& P = ctx->plan;
  ElemParams prm;
  prm.p0 = params ? params[0] : 0.0;
  prm.p1 = params ? params[1] : 0.0;
  prm.flags = flags;
  if (P.nb_tile == 0) return AFB_OK;
  const size_t smem = sizeof(double) * (TG_ZERO + 1);
  const int grid = std::min<int>(P.nb_tile, ctx->sm_count);
  // values already holding contributions (a second operator added on top) are accumulated into;
  // a freshly reset matrix is simply overwritten
  const int accumulate = ctx->assembled ? 1 : 0;
  if (ctx->npc == 4) {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled<4><<<grid, TG_THREADS, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.nb_tile, ctx->coords.as<double>(), ctx->conn.as<int32_t>(),
                                                                 P.tile_cells.as<int32_t>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(),
                                                                 P.lists.as<uint16_t>(), ctx->values.as<double>(), accumulate, prm);
  }
  else {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled<3><<<grid, TG_THREADS, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.nb_tile, ctx->coords.as<double>(), ctx->conn.as<int32_t>(),
                                                                 P.tile_cells.as<int32_t>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(),
                                                                 P.lists.as<uint16_t>(), ctx->values.as<double>(), accumulate, prm);
  }
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

} // namespace afb
