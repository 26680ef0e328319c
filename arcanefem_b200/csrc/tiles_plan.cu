// Inspector of the tiled path (runs once per mesh, at the first tiled assembly).
//
//   mesh tiling   nodes are binned into spatial bricks; a brick is cut in pieces (= tiles) until
//                 each piece fits the executor's shared-memory budget.  Per tile: its rows
//                 (ascending node ids), the cells touching it, its footprint (rows + halo nodes,
//                 ascending node ids), the cells' connectivity in footprint-local 16-bit indices,
//                 and per row the incident cells' other nodes as packed 10-bit footprint indices
//                 (the localized form of Arcane's nodeCell x cellNode views that the reference's
//                 node-wise back-ends walk: modules/testlab/NodeWiseCsrBiliAssembly.cc:179-220).
//                 The steady-state BuildMatrix kernel (pattern_tiled.cu) runs on these.
//   value plan    for every matrix entry of a tile's rows the list of (cell, local pair)
//                 contributions as 16-bit indices into the tile's element-matrix cache, sorted by
//                 ascending cell id (fixed summation order => bit-reproducible), entries sorted by
//                 list length and cut in units of 32 (one warp walks 32 equally long lists).
//                 Symmetric twins inside a tile are computed once; for operators with zero row
//                 sums (Poisson) the diagonal is not listed at all: the executor derives it from
//                 the row's off-diagonals.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "element.cuh"
#include "tiles.cuh"

namespace afb {

constexpr int TB_THREADS = 512;  // builder CTA
constexpr int TB_HASH = 2048;    // footprint hash slots (>= 2 * TG_FMAX)
constexpr unsigned TB_EMPTY = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// bricks
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long order_f64(double x)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static double unorder_f64(unsigned long long u)
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

__global__ void __launch_bounds__(256) k_brick_assign(const double* __restrict__ coords, int32_t nb_node, BrickGrid bg, int32_t* __restrict__ brick_of, int32_t* __restrict__ count)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_node) return;
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

__global__ void __launch_bounds__(256) k_brick_fill(const int32_t* __restrict__ brick_of, int32_t nb_node, const int32_t* __restrict__ brick_ptr, int32_t* __restrict__ cursor,
                                                     int32_t* __restrict__ tnodes)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_node) return;
  const int32_t b = brick_of[i];
  tnodes[brick_ptr[b] + atomicAdd(cursor + b, 1)] = i;
}

// ascending node ids inside each brick (the atomic fill order is arbitrary); rank sort in shared
// memory for bricks of <= 2048 nodes, larger ones (then cut in many pieces anyway) are sorted by
// a single thread per brick with an in-place heap sort
__global__ void __launch_bounds__(256) k_brick_sort(const int32_t* __restrict__ brick_ptr, int32_t nb_brick, int32_t* __restrict__ tnodes)
{
  __shared__ int32_t s[2048];
  const int b = blockIdx.x;
  if (b >= nb_brick) return;
  const int beg = brick_ptr[b], n = brick_ptr[b + 1] - beg;
  if (n <= 1) return;
  if (n > 2048) {
    if (threadIdx.x == 0) {
      int32_t* a = tnodes + beg;
      for (int start = n / 2 - 1; start >= 0; --start) {
        int root = start;
        while (2 * root + 1 < n) {
          int ch = 2 * root + 1;
          if (ch + 1 < n && a[ch] < a[ch + 1]) ++ch;
          if (a[root] >= a[ch]) break;
          const int32_t tmp = a[root]; a[root] = a[ch]; a[ch] = tmp;
          root = ch;
        }
      }
      for (int end = n - 1; end > 0; --end) {
        int32_t tmp = a[0]; a[0] = a[end]; a[end] = tmp;
        int root = 0;
        while (2 * root + 1 < end) {
          int ch = 2 * root + 1;
          if (ch + 1 < end && a[ch] < a[ch + 1]) ++ch;
          if (a[root] >= a[ch]) break;
          tmp = a[root]; a[root] = a[ch]; a[ch] = tmp;
          root = ch;
        }
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = tnodes[beg + i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t x = s[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += s[j] < x ? 1 : 0;
    tnodes[beg + rank] = x;
  }
}

// a brick of n nodes cut in nt pieces of ceil(n/nt) consecutive (ascending id) nodes
__global__ void __launch_bounds__(256) k_tile_nodes(const int32_t* __restrict__ brick_of, const int32_t* __restrict__ brick_ptr, const int32_t* __restrict__ tile_first,
                                                     const int32_t* __restrict__ ntile_of, const int32_t* __restrict__ tnodes, int32_t nb_node,
                                                     int32_t* __restrict__ node_tile, int32_t* __restrict__ node_lrow, TileDesc* __restrict__ desc)
{
  const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nb_node) return;
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

// per tile: number of cells, entries, largest valence, footprint nodes (stops counting once the
// footprint is known to be too large)
template <int NPC>
__global__ void __launch_bounds__(128) k_tile_stats(const TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ conn,
                                                     const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, const int32_t* __restrict__ rows,
                                                     const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow, int32_t* __restrict__ stats /* [nb_tile][4] */)
{
  __shared__ int s_c, s_e, s_v, s_h;
  __shared__ unsigned s_tab[TB_HASH];
  const int32_t t = blockIdx.x;
  if (t >= nb_tile) return;
  if (threadIdx.x == 0) s_c = s_e = s_v = s_h = 0;
  for (int i = threadIdx.x; i < TB_HASH; i += blockDim.x) s_tab[i] = TB_EMPTY;
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
    for (int q = qb + lane; q < qe; q += 32) {
      const int32_t cell = nc_list[q];
      if (is_leader<NPC>(conn, node_tile, node_lrow, t, i, r, cell)) {
        ++c;
        if (s_h <= TG_FMAX) {
#pragma unroll
          for (int a = 0; a < NPC; ++a) {
            const int32_t n = conn[(int64_t)cell * NPC + a];
            if (node_tile[n] != t) {
              unsigned h = ((unsigned)n * 0x9E3779B1u) >> 21; // 11 bits = TB_HASH
              while (true) {
                const unsigned old = atomicCAS(s_tab + h, TB_EMPTY, (unsigned)n);
                if (old == TB_EMPTY) { atomicAdd(&s_h, 1); break; }
                if (old == (unsigned)n) break;
                h = (h + 1) & (TB_HASH - 1);
              }
            }
          }
        }
      }
    }
  }
  atomicAdd(&s_c, c);
  atomicAdd(&s_e, e);
  atomicMax(&s_v, v);
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[4 * t + 0] = s_c;
    stats[4 * t + 1] = s_e;
    stats[4 * t + 2] = s_v;
    stats[4 * t + 3] = s_h + d.nb_row;
  }
}

// ---------------------------------------------------------------------------------------------
// shared-memory helpers of the builders
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

// in-place exclusive scan of n ints in shared memory; returns the total
__device__ int smem_exclusive_scan(int* s, int n, int* s_tmp /* >= 33 ints */)
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

// index of `id` in the ascending array s[0,n) (present by construction)
__device__ __forceinline__ int smem_find(const unsigned* __restrict__ s, int n, unsigned id)
{
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s[mid] <= id) lo = mid; else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// mesh tiling: cells, footprint, local connectivity, incidence lists
// ---------------------------------------------------------------------------------------------
struct MeshBuilderSmem {
  unsigned cells[2048];
  unsigned htab[TB_HASH];
  unsigned sfoot[1024];
  int gmax[TG_GMAX], goff[TG_GMAX + 1];
  uint16_t rowfs[TG_RMAX]; // footprint index of every row
  int nb_cell, nb_foot;
};

template <int NPC>
__global__ void __launch_bounds__(TB_THREADS)
k_tile_mesh(const TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ conn, const int32_t* __restrict__ nc_ptr,
            const int32_t* __restrict__ nc_list, const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow, int32_t* __restrict__ tile_cells,
            int32_t* __restrict__ foot, ushort4* __restrict__ lconn, uint16_t* __restrict__ rowf, uint32_t* __restrict__ inc, uint2* __restrict__ inc_grp, int* __restrict__ error)
{
  __shared__ MeshBuilderSmem S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    const int R = d.nb_row;
    if (threadIdx.x == 0) S.nb_cell = S.nb_foot = 0;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) S.cells[i] = TB_EMPTY;
    for (int i = threadIdx.x; i < TB_HASH; i += blockDim.x) S.htab[i] = TB_EMPTY;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) S.sfoot[i] = TB_EMPTY;
    __syncthreads();
    // ---- leader cells, ascending ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
      for (int q = qb + lane; q < qe; q += 32) {
        const int32_t c = nc_list[q];
        if (is_leader<NPC>(conn, node_tile, node_lrow, t, i, r, c)) {
          const int pos = atomicAdd(&S.nb_cell, 1);
          if (pos < 2048) S.cells[pos] = (unsigned)c;
        }
      }
    }
    __syncthreads();
    const int C = S.nb_cell;
    if (C != d.nb_cell || C > 2048) {
      if (threadIdx.x == 0) atomicExch(error, 1);
      __syncthreads();
      continue;
    }
    int c2 = 32;
    while (c2 < C) c2 <<= 1;
    smem_bitonic_sort(S.cells, c2);
    // ---- footprint = rows + nodes of the cells, ascending ----
    auto insert = [&](unsigned n) {
      unsigned h = (n * 0x9E3779B1u) >> 21;
      while (true) {
        const unsigned old = atomicCAS(S.htab + h, TB_EMPTY, n);
        if (old == TB_EMPTY) {
          const int pos = atomicAdd(&S.nb_foot, 1);
          if (pos < 1024) S.sfoot[pos] = n;
          return;
        }
        if (old == n) return;
        h = (h + 1) & (TB_HASH - 1);
      }
    };
    for (int i = threadIdx.x; i < R; i += blockDim.x) insert((unsigned)tnodes[d.node_off + i]);
    for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
      tile_cells[d.cell_off + lc] = (int32_t)S.cells[lc];
      const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
#pragma unroll
      for (int a = 0; a < NPC; ++a) insert((unsigned)__ldg(cn + a));
    }
    __syncthreads();
    const int F = S.nb_foot;
    if (F != d.nb_foot || F > TG_FMAX) {
      if (threadIdx.x == 0) atomicExch(error, 2);
      __syncthreads();
      continue;
    }
    int f2 = 32;
    while (f2 < F) f2 <<= 1;
    smem_bitonic_sort(S.sfoot, f2);
    for (int f = threadIdx.x; f < F; f += blockDim.x) foot[d.foot_off + f] = (int32_t)S.sfoot[f];
    for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
      const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
      unsigned short loc[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int a = 0; a < NPC; ++a) loc[a] = (unsigned short)smem_find(S.sfoot, F, (unsigned)__ldg(cn + a));
      lconn[d.cell_off + lc] = make_ushort4(loc[0], loc[1], loc[2], loc[3]);
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
      const uint16_t f = (uint16_t)smem_find(S.sfoot, F, (unsigned)tnodes[d.node_off + i]);
      rowf[d.node_off + i] = f;
      S.rowfs[i] = f;
    }
    // ---- incidence lists: groups of 32 rows, padded to the group's largest valence ----
    const int ngroup = (R + 31) >> 5;
    for (int g = warp; g < ngroup; g += nwarp) {
      const int i = g * 32 + lane;
      int v = 0;
      if (i < R) {
        const int32_t r = tnodes[d.node_off + i];
        v = nc_ptr[r + 1] - nc_ptr[r];
      }
      v = __reduce_max_sync(0xffffffffu, v);
      if (lane == 0) S.gmax[g] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = 0;
      for (int g = 0; g < ngroup; ++g) {
        S.goff[g] = run;
        run += 32 * S.gmax[g];
      }
      S.goff[ngroup] = run;
    }
    __syncthreads();
    for (int g = threadIdx.x; g < TG_GMAX; g += blockDim.x)
      inc_grp[(size_t)t * TG_GMAX + g] = g < ngroup ? make_uint2((unsigned)S.goff[g], (unsigned)S.gmax[g]) : make_uint2(0u, 0u);
    // one thread per (row, incident-cell slot): all warps of the CTA work on the lists of all groups
    const int total = S.goff[ngroup];
    for (int x = threadIdx.x; x < total; x += blockDim.x) {
      int g = 0;
      while (g + 1 < ngroup && x >= S.goff[g + 1]) ++g;
      const int y = x - S.goff[g], k = y >> 5, i = g * 32 + (y & 31);
      unsigned w = 0u;
      if (i < R) {
        const int32_t r = tnodes[d.node_off + i];
        const int qb = nc_ptr[r], v = nc_ptr[r + 1] - qb;
        const unsigned self = S.rowfs[i];
        unsigned f[3] = { self, self, self };
        if (k < v) {
          const int32_t* cn = conn + (int64_t)nc_list[qb + k] * NPC;
          int m = 0;
#pragma unroll
          for (int a = 0; a < NPC; ++a) {
            const int32_t n = __ldg(cn + a);
            if (n != r && m < 3) f[m++] = (unsigned)smem_find(S.sfoot, F, (unsigned)n);
          }
        }
        w = f[0] | (f[1] << 10) | (f[2] << 20);
      }
      inc[d.inc_off + x] = w;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// value plan: contribution lists
// ---------------------------------------------------------------------------------------------
// Entry classes of a tile (row i = local row index, column node c):
//   computed : columns outside the tile, and columns inside the tile with a larger row index
//              ("upper": the value is also stored at the mirror position (row of c, column of i),
//              the element matrices being symmetric); the diagonal when it is listed (VEC)
//   mirror   : columns inside the tile with a smaller row index: written by their upper twin
//   derived  : the diagonal of a zero-row-sum operator: minus the sum of the row's off-diagonals
//   zero     : entries of rows the assembly does not write (non-owned nodes: the isOwn gate)
constexpr int CLS_MIRROR = -1, CLS_ZERO = -2, CLS_DERIVED = -3;
constexpr int TL_KEYS = 4096; // power of two >= TG_EMAX

struct ListBuilderSmem {
  int erow[TG_RMAX + 1];
  int cnt[TG_EMAX];
  int eoff[TG_EMAX + 1];
  unsigned keys[TL_KEYS];
  unsigned scol[TG_EMAX]; // columns of the tile's rows, in tile entry order (the searches below stay in shared memory)
  uint16_t e2[TG_EMAX];
  uint16_t clist[16 * TG_CMAX];
  int ulen[TG_UMAX + 1], ubase[TG_UMAX + 2];
  int tmp[40];
};

template <int NPC, bool VEC>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_tile_lists(TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ tile_cells, const int32_t* __restrict__ conn,
             const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow,
             const uint8_t* __restrict__ is_own, int64_t nb_own_cell, uint32_t* __restrict__ rowinfo, uint32_t* __restrict__ unit_base, uint16_t* __restrict__ unit_len,
             uint32_t* __restrict__ emap, uint32_t* __restrict__ emap_rows, uint16_t* __restrict__ lists, int list_max, int* __restrict__ error)
{
  extern __shared__ unsigned char tl_raw[];
  ListBuilderSmem& S = *reinterpret_cast<ListBuilderSmem*>(tl_raw);
  constexpr int CS = VEC ? TV_CS : TG_CS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    const int R = d.nb_row, C = d.nb_cell;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) {
      int deg = 0;
      if (i < R) {
        const int32_t r = tnodes[d.node_off + i];
        deg = rows[r + 1] - rows[r];
      }
      S.erow[i] = deg;
    }
    __syncthreads();
    const int E = smem_exclusive_scan(S.erow, R + 1, S.tmp);
    if (E != d.nb_entry || E > TG_EMAX) {
      if (threadIdx.x == 0) atomicExch(error, 1);
      __syncthreads();
      continue;
    }
    // ---- columns of the tile's rows -> shared memory; position of the diagonal ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int rb = rows[r], deg = rows[r + 1] - rb, e0 = S.erow[i];
      int pdiag = 0;
      for (int p0 = 0; p0 < deg; p0 += 32) {
        const int p = p0 + lane;
        int32_t c = -1;
        if (p < deg) {
          c = cols[rb + p];
          S.scol[e0 + p] = (unsigned)c;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, c == r);
        if (hit) pdiag = p0 + __ffs(hit) - 1;
      }
      if (lane == 0) rowinfo[d.node_off + i] = pack_rowinfo(e0, pdiag, !is_own || is_own[r]);
    }
    __syncthreads();
    // ---- entries: class, mirror ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const bool own_i = !is_own || is_own[r];
      const int e0 = S.erow[i], deg = S.erow[i + 1] - e0;
      for (int p = lane; p < deg; p += 32) {
        const int32_t c = (int32_t)S.scol[e0 + p];
        int cls = 0;
        unsigned m = TG_NONE16;
        if (!own_i) cls = CLS_ZERO;
        else if (c == r) cls = VEC ? 0 : CLS_DERIVED;
        else if (__ldg(node_tile + c) == t && (!is_own || is_own[c])) {
          const int j = __ldg(node_lrow + c);
          if (j < i) cls = CLS_MIRROR;
          else m = (unsigned)(S.erow[j] + smem_find(S.scol + S.erow[j], S.erow[j + 1] - S.erow[j], (unsigned)r));
        }
        S.cnt[e0 + p] = cls;
        S.e2[e0 + p] = (uint16_t)m;
      }
    }
    __syncthreads();
    // ---- contribution lists of the computed entries (count, scan, fill) ----
    for (int pass = 0; pass < 2; ++pass) {
      for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
        const int32_t cell = tile_cells[d.cell_off + lc];
        if ((int64_t)cell >= nb_own_cell) continue; // ghost cells contribute nothing (domain-decomposition mode B)
        const int32_t* cn = conn + (int64_t)cell * NPC;
        int32_t nd[NPC];
        int li[NPC];
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          nd[a] = __ldg(cn + a);
          li[a] = (__ldg(node_tile + nd[a]) == t && (!is_own || is_own[nd[a]])) ? __ldg(node_lrow + nd[a]) : -1;
        }
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          if (li[a] < 0) continue;
          const int ea = S.erow[li[a]], dega = S.erow[li[a] + 1] - ea;
#pragma unroll
          for (int bq = 0; bq < NPC; ++bq) {
            if (!VEC && bq == a) continue;                         // derived diagonal
            if (bq != a && li[bq] >= 0 && li[bq] < li[a]) continue; // the twin entry (li[bq], li[a]) takes it
            const int e = ea + smem_find(S.scol + ea, dega, (unsigned)nd[bq]);
            if (pass == 0) atomicAdd(&S.cnt[e], 1);
            else {
              const int slot = atomicSub(&S.cnt[e], 1) - 1; // countdown cursor, restored from eoff below
              const int plane = VEC ? (a * NPC + bq) : off_pair(NPC, a, bq);
              S.clist[S.eoff[e] + slot] = (uint16_t)(plane * CS + lc);
            }
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        for (int e = threadIdx.x; e <= E; e += blockDim.x) S.eoff[e] = e < E ? max(S.cnt[e], 0) : 0;
        __syncthreads();
        smem_exclusive_scan(S.eoff, E + 1, S.tmp);
      }
      else {
        for (int e = threadIdx.x; e < E; e += blockDim.x)
          if (S.cnt[e] >= 0) S.cnt[e] = S.eoff[e + 1] - S.eoff[e];
        __syncthreads();
      }
    }
    // ---- canonical order first: ascending local cell index (= ascending global cell id) ----
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      uint16_t* l = S.clist + S.eoff[e];
      const int n = S.cnt[e];
      for (int i = 1; i < n; ++i) {
        const uint16_t x = l[i];
        const unsigned kx = ((unsigned)(x % CS) << 16) | x;
        int j = i - 1;
        while (j >= 0) {
          const uint16_t y = l[j];
          if ((((unsigned)(y % CS) << 16) | y) <= kx) break;
          l[j + 1] = y;
          --j;
        }
        l[j + 1] = x;
      }
    }
    // ---- computed entries by descending count, cut into units of 32 ----
    int k2 = 32;
    while (k2 < E) k2 <<= 1;
    for (int e = threadIdx.x; e < k2; e += blockDim.x)
      S.keys[e] = (e < E && S.cnt[e] >= 0) ? (((unsigned)(0xFFFF - min(S.cnt[e], 0xFFFF)) << 16) | (unsigned)e) : 0xFFFFFFFFu;
    if (threadIdx.x == 0) S.tmp[34] = 0;
    __syncthreads();
    {
      int mine = 0;
      for (int e = threadIdx.x; e < E; e += blockDim.x) mine += S.cnt[e] >= 0 ? 1 : 0;
      atomicAdd(&S.tmp[34], mine);
    }
    smem_bitonic_sort(S.keys, k2);
    const int EC = S.tmp[34]; // computed entries
    const int nunit = (EC + 31) / 32;
    for (int u = threadIdx.x; u <= nunit; u += blockDim.x) {
      int len = 0;
      if (u < nunit) len = (S.cnt[S.keys[u * 32] & 0xFFFFu] + 1) & ~1;
      S.ulen[u] = len;
      S.ubase[u] = len * 32;
    }
    __syncthreads();
    const int list_total = (smem_exclusive_scan(S.ubase, nunit + 1, S.tmp) + 7) & ~7;
    if (nunit > d.nb_unit || list_total > list_max) {
      if (threadIdx.x == 0) atomicExch(error, list_total > list_max ? 3 : 2);
      __syncthreads();
      continue;
    }
    if (threadIdx.x == 0) {
      desc[t].nb_unit = nunit;
      desc[t].list_len = list_total;
    }
    for (int u = threadIdx.x; u < nunit; u += blockDim.x) {
      unit_base[d.unit_off + u] = (uint32_t)S.ubase[u]; // relative to the tile's list region
      unit_len[d.unit_off + u] = (uint16_t)S.ulen[u];
    }
    constexpr uint16_t PAD = (uint16_t)(VEC ? (CS - 1) : TG_ZERO);
    for (int x = S.ubase[nunit] + threadIdx.x; x < list_total; x += blockDim.x) lists[d.list_off + x] = PAD;
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const bool valid = x < EC;
      const int e = valid ? (int)(S.keys[x] & 0xFFFFu) : 0;
      emap[(size_t)(d.unit_off + (x >> 5)) * 32 + (x & 31)] = valid ? ((uint32_t)e | ((uint32_t)S.e2[e] << 16)) : 0xFFFFFFFFu;
      if (VEC) { // tile rows holding the entry and its mirror (the vector executor writes blocks straight to their rows)
        auto row_of = [&](int ee) {
          int lo = 0, hi = R;
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (S.erow[mid] <= ee) lo = mid; else hi = mid;
          }
          return (uint32_t)lo;
        };
        uint32_t w = 0xFFFFFFFFu;
        if (valid) w = row_of(e) | ((S.e2[e] != TG_NONE16 ? row_of((int)S.e2[e]) : (uint32_t)TG_NONE16) << 16);
        emap_rows[(size_t)(d.unit_off + (x >> 5)) * 32 + (x & 31)] = w;
      }
    }
    // the lists in canonical order ([len/2][32 lanes][2], padded); k_bank_order below fixes the order inside each list
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const int u = x >> 5, ln = x & 31, len = S.ulen[u];
      const bool valid = x < EC;
      const int e = valid ? (int)(S.keys[x] & 0xFFFFu) : 0;
      const int c = valid ? S.cnt[e] : 0;
      const uint16_t* src = S.clist + (valid ? S.eoff[e] : 0);
      uint16_t* out = lists + d.list_off + S.ubase[u] + ln * 2;
      for (int k = 0; k < len; ++k) out[(k >> 1) * 64 + (k & 1)] = k < c ? src[k] : PAD;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// Order inside the lists of a unit (second kernel of the value plan: one thread per half-unit over the whole grid).
// The executor reads contribution k of its 16 lanes (a half-warp) with one shared-memory instruction: the k-th slots of
// the 16 lists are chosen so that they fall into different banks (8-byte bank = cache index mod 16) whenever possible --
// greedy choice in lane order with one-step augmentation (a lane that finds all its banks taken may move an earlier lane
// to another free bank of that lane's remaining contributions).  The order inside an entry's list is therefore
// plan-defined (not ascending cell id), but fixed: the sums stay bit-reproducible.
//   lists: in canonical order on entry (padded with `pad`), reordered in place
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_bank_order(const TileDesc* __restrict__ desc, const uint32_t* __restrict__ unit_base, const uint16_t* __restrict__ unit_len,
                                                    uint16_t* __restrict__ lists, uint16_t pad)
{
  const TileDesc d = desc[blockIdx.x];
  for (int hx = threadIdx.x; hx < d.nb_unit * 2; hx += blockDim.x) {
    const int u = hx >> 1, l0 = (hx & 1) * 16;
    const int len = unit_len[d.unit_off + u];
    const size_t base = (size_t)d.list_off + unit_base[d.unit_off + u] + l0 * 2;
    uint16_t* out0 = lists + base; // [len/2][32 lanes][2], canonical order on entry, reordered in place
    auto cand = [&](int j, int q) { return out0[j * 2 + (q >> 1) * 64 + (q & 1)]; };
    // the banks of a list's contributions, 4 bits each (lists of up to 16: the searches below run on registers / local
    // words; a longer list keeps its canonical order)
    int cnt_l[16];
    unsigned long long bw[16];
    unsigned taken[16];
    bool simple = false;
    for (int j = 0; j < 16; ++j) {
      int c = 0;
      unsigned long long w = 0ull;
      while (c < len) {
        const unsigned cd = cand(j, c);
        if (cd == pad) break;
        if (c < 16) w |= (unsigned long long)(cd & 15u) << (4 * c);
        ++c;
      }
      cnt_l[j] = c;
      bw[j] = w;
      taken[j] = 0u;
      if (c > 16) simple = true;
    }
    if (simple) continue; // canonical order stays
    unsigned long long perm[16]; // list -> its contribution of step k, 4 bits per step
    for (int j = 0; j < 16; ++j) perm[j] = 0ull;
    for (int k = 0; k < len; ++k) {
      signed char owner[16];  // bank -> lane
      signed char choice[16]; // lane -> index into its list
      for (int j = 0; j < 16; ++j) { owner[j] = -1; choice[j] = -1; }
      {
        for (int j = 0; j < 16; ++j) {
          if (k >= cnt_l[j]) continue;
          const int n = cnt_l[j];
          const unsigned long long wj = bw[j];
          const unsigned tj = taken[j];
          int first = -1;
          for (int q = 0; q < n && choice[j] < 0; ++q) {
            if ((tj >> q) & 1u) continue;
            if (first < 0) first = q;
            const int bk = (int)((wj >> (4 * q)) & 15ull);
            if (owner[bk] < 0) { owner[bk] = (signed char)j; choice[j] = (signed char)q; }
          }
          if (choice[j] >= 0) continue;
          // all banks of this lane's remaining contributions are taken: try to move one of their owners
          for (int q = 0; q < n && choice[j] < 0; ++q) {
            if ((tj >> q) & 1u) continue;
            const int bk = (int)((wj >> (4 * q)) & 15ull);
            const int j2 = owner[bk];
            if (j2 < 0 || j2 == j) continue;
            const int n2 = cnt_l[j2];
            const unsigned long long w2 = bw[j2];
            const unsigned t2 = taken[j2];
            for (int q2 = 0; q2 < n2; ++q2) {
              if (((t2 >> q2) & 1u) || q2 == choice[j2]) continue;
              const int b2 = (int)((w2 >> (4 * q2)) & 15ull);
              if (owner[b2] < 0) {
                owner[b2] = (signed char)j2;
                choice[j2] = (signed char)q2;
                owner[bk] = (signed char)j;
                choice[j] = (signed char)q;
                break;
              }
            }
          }
          if (choice[j] < 0) choice[j] = (signed char)first; // a bank conflict remains
        }
      }
      for (int j = 0; j < 16; ++j)
        if (k < cnt_l[j]) {
          taken[j] |= 1u << choice[j];
          perm[j] |= (unsigned long long)choice[j] << (4 * k);
        }
    }
    // apply: every list is permuted in place
    for (int j = 0; j < 16; ++j) {
      const int n = cnt_l[j];
      if (n < 2) continue;
      uint16_t c[16];
      for (int q = 0; q < n; ++q) c[q] = cand(j, q);
      for (int k = 0; k < n; ++k) out0[j * 2 + (k >> 1) * 64 + (k & 1)] = c[(perm[j] >> (4 * k)) & 15ull];
    }
  }
}

// the same for the row-ordered plans (k_tile_rowlists): two gradients are read per contribution (node a and node b of the
// cell), both should fall into different banks across the 16 lanes; equal words are broadcast and do not collide; lists
// that have no step to spare choose first, a list shorter than the unit's may sit a step out (padding slot) when every
// remaining contribution of it would collide
template <int NPC>
__global__ void __launch_bounds__(128) k_bank_order_rows(const TileDesc* __restrict__ desc, const uint2* __restrict__ units, uint16_t* __restrict__ lists)
{
  constexpr int CS = VR_CS;
  constexpr uint16_t PAD = (uint16_t)(CS - 1);
  const TileDesc d = desc[blockIdx.x];
  for (int hx = threadIdx.x; hx < d.nb_unit * 2; hx += blockDim.x) {
    const int u = hx >> 1, l0 = (hx & 1) * 16;
    const uint2 U = units[d.unit_off + u];
    const int len = (int)(U.y >> 17);
    const size_t base = (size_t)d.list_off + U.x + l0 * 2;
    uint16_t* out0 = lists + base; // canonical order on entry, reordered in place
    auto cand = [&](int j, int q) { return out0[j * 2 + (q >> 1) * 64 + (q & 1)]; };
    // per list: the two banks (8-byte bank of the word of node a / node b) of every contribution, 4 bits each, and the low
    // 12 bits of the two words (equal words are broadcast): lists of up to 16; a longer list keeps its canonical order
    unsigned taken[16];
    unsigned char cnt_l[16], used[16];
    unsigned long long bwa[16], bwb[16];
    bool simple = false;
    auto words = [&](unsigned cd, int& wa, int& wb) {
      const unsigned pl = cd >> VR_LC_BITS, lc = cd & VR_LC_MASK;
      wa = (int)((pl >> 2) * (NPC - 1) * CS + lc);
      wb = (int)((pl & 3u) * (NPC - 1) * CS + lc);
    };
    for (int j = 0; j < 16; ++j) {
      int c = 0;
      unsigned long long a = 0ull, b = 0ull;
      while (c < len) {
        const unsigned cd = cand(j, c);
        if (cd == PAD) break;
        int wa, wb;
        words(cd, wa, wb);
        if (c < 16) {
          a |= (unsigned long long)(wa & 15) << (4 * c);
          b |= (unsigned long long)(wb & 15) << (4 * c);
        }
        ++c;
      }
      if (c > 16) simple = true;
      cnt_l[j] = (unsigned char)min(c, 255);
      bwa[j] = a;
      bwb[j] = b;
      taken[j] = 0u;
      used[j] = 0;
    }
    if (simple || len > 16) continue; // canonical order stays
    unsigned long long perm[16]; // list -> its contribution of step k (4 bits per step)
    unsigned short sits[16];     // list -> steps it sits out (padding slot)
    for (int j = 0; j < 16; ++j) { perm[j] = 0ull; sits[j] = 0; }
    for (int k = 0; k < len; ++k) {
      short ownA[16], ownB[16]; // 8-byte bank -> cache word read from it in this step (-1: free)
      for (int q = 0; q < 16; ++q) ownA[q] = ownB[q] = -1;
      unsigned must_mask = 0u;
      for (int j = 0; j < 16; ++j)
        if (simple || (int)cnt_l[j] - (int)used[j] >= len - k) must_mask |= 1u << j;
      for (int pass = 0; pass < 2; ++pass) {
        for (int j = 0; j < 16; ++j) {
          const bool must = (must_mask >> j) & 1u;
          if (must != (pass == 0)) continue;
          const int c = cnt_l[j], rem = c - used[j];
          bool took = false;
          if (rem > 0) {
            // a free bank is certainly no collision; an occupied one may hold the same word (checked on the few that tie)
            const unsigned long long wa_j = bwa[j], wb_j = bwb[j];
            int best = -1, best_score = -1;
            for (int q = 0; q < c; ++q) {
              if ((taken[j] >> q) & 1u) continue;
              const int ba = (int)((wa_j >> (4 * q)) & 15ull), bb = (int)((wb_j >> (4 * q)) & 15ull);
              int sa = ownA[ba] < 0 ? 1 : 0, sb = ownB[bb] < 0 ? 1 : 0, shared = 0;
              if (sa + sb < 2) {
                int wa, wb;
                words(cand(j, q), wa, wb);
                if (!sa && ownA[ba] == wa) { sa = 1; ++shared; }
                if (!sb && ownB[bb] == wb) { sb = 1; ++shared; }
              }
              // no collision first; among those, words another lane reads anyway (a cell shared by several entries of the
              // row is read once for all of them and takes no further bank)
              const int score = VR_SHARE ? 4 * (sa + sb) + shared : sa + sb;
              if (score > best_score) { best_score = score; best = q; }
              if (!VR_SHARE && score == 2) break;
            }
            if (must || best_score >= (VR_SHARE ? 8 : 2)) {
              took = true;
              taken[j] |= 1u << best;
              ++used[j];
              perm[j] |= (unsigned long long)best << (4 * k);
              int wa, wb;
              words(cand(j, best), wa, wb);
              if (ownA[wa & 15] < 0) ownA[wa & 15] = (short)wa;
              if (ownB[wb & 15] < 0) ownB[wb & 15] = (short)wb;
            }
          }
          if (!took) sits[j] |= (unsigned short)(1u << k);
        }
      }
    }
    // apply: every list is permuted in place (padding slots where it sits out)
    for (int j = 0; j < 16; ++j) {
      const int n = cnt_l[j];
      if (n == 0) continue;
      uint16_t c[16];
      for (int q = 0; q < n; ++q) c[q] = cand(j, q);
      for (int k = 0; k < len; ++k) out0[j * 2 + (k >> 1) * 64 + (k & 1)] = ((sits[j] >> k) & 1u) ? PAD : c[(perm[j] >> (4 * k)) & 15ull];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// value plan of the row-ordered vector executor (k_assemble_rows_vec)
//   entries in tile order (row by row, columns ascending); a unit = up to 32 consecutive entries made of whole rows (a
//   row longer than 32 entries is cut into units of its own); every off-diagonal entry carries the list of its cells
//   (both orientations of a pair are computed: no mirror writes, so a unit's output is a few contiguous runs); the
//   diagonal block is derived from the row's blocks (zero block row sums) unless the row was cut, then it has a list.
//   FILL = false: only the sizes (desc.nb_unit, desc.list_len) -- the host turns them into offsets.
// ---------------------------------------------------------------------------------------------
struct RowListSmem {
  int erow[TG_RMAX + 1];
  int cnt[TG_EMAX];
  int eoff[TG_EMAX + 1];
  uint16_t clist[16 * TG_CMAX];
  int ufirst[VR_UMAX + 1], ucnt[VR_UMAX + 1], ubase[VR_UMAX + 2];
  uint8_t own[TG_RMAX];
  int tmp[40];
};

template <int NPC, bool FILL>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_tile_rowlists(TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ tile_cells, const int32_t* __restrict__ conn,
                const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ node_tile, const int32_t* __restrict__ node_lrow,
                const uint8_t* __restrict__ is_own, int64_t nb_own_cell, uint32_t* __restrict__ rowinfo, uint2* __restrict__ units, uint16_t* __restrict__ lists,
                int* __restrict__ error)
{
  extern __shared__ unsigned char tl_raw[];
  RowListSmem& S = *reinterpret_cast<RowListSmem*>(tl_raw);
  constexpr int CS = VR_CS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    const int R = d.nb_row, C = d.nb_cell;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) {
      int deg = 0;
      if (i < R) {
        const int32_t r = tnodes[d.node_off + i];
        deg = rows[r + 1] - rows[r];
        S.own[i] = (!is_own || is_own[r]) ? 1 : 0;
      }
      S.erow[i] = deg;
    }
    __syncthreads();
    const int E = smem_exclusive_scan(S.erow, R + 1, S.tmp);
    if (E != d.nb_entry || E > TG_EMAX) {
      if (threadIdx.x == 0) atomicExch(error, 1);
      __syncthreads();
      continue;
    }
    for (int e = threadIdx.x; e < E; e += blockDim.x) S.cnt[e] = 0;
    // ---- row words: first entry, position of the diagonal, ownership ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int rb = rows[r], deg = rows[r + 1] - rb;
      int pdiag = 0;
      for (int p0 = 0; p0 < deg; p0 += 32) {
        const int p = p0 + lane;
        const unsigned hit = __ballot_sync(0xffffffffu, p < deg && cols[rb + p] == r);
        if (hit) pdiag = p0 + __ffs(hit) - 1;
      }
      if (FILL && lane == 0) rowinfo[d.node_off + i] = pack_rowinfo(S.erow[i], pdiag, S.own[i] != 0);
    }
    __syncthreads();
    // ---- contribution lists (count, scan, fill) ----
    for (int pass = 0; pass < (FILL ? 2 : 1); ++pass) {
      for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
        const int32_t cell = tile_cells[d.cell_off + lc];
        if ((int64_t)cell >= nb_own_cell) continue; // ghost cells contribute nothing (domain-decomposition mode B)
        const int32_t* cn = conn + (int64_t)cell * NPC;
        int32_t nd[NPC];
        int li[NPC];
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          nd[a] = __ldg(cn + a);
          li[a] = (__ldg(node_tile + nd[a]) == t) ? __ldg(node_lrow + nd[a]) : -1;
          if (li[a] >= 0 && !S.own[li[a]]) li[a] = -1;
        }
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          if (li[a] < 0) continue;
          const int rb = __ldg(rows + nd[a]), re = __ldg(rows + nd[a] + 1);
          const bool cut = re - rb > 32; // the row spans several units: its diagonal block has a list of its own
#pragma unroll
          for (int bq = 0; bq < NPC; ++bq) {
            if (bq == a && !cut) continue;
            const int e = S.erow[li[a]] + (find_col(cols, rb, re, nd[bq]) - rb);
            if (pass == 0) atomicAdd(&S.cnt[e], 1);
            else {
              const int slot = atomicSub(&S.cnt[e], 1) - 1; // countdown cursor, restored from eoff below
              S.clist[S.eoff[e] + slot] = (uint16_t)(((a * 4 + bq) << VR_LC_BITS) | lc);
            }
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        if (FILL) {
          for (int e = threadIdx.x; e <= E; e += blockDim.x) S.eoff[e] = e < E ? S.cnt[e] : 0;
          __syncthreads();
          smem_exclusive_scan(S.eoff, E + 1, S.tmp);
        }
      }
      else {
        for (int e = threadIdx.x; e < E; e += blockDim.x) S.cnt[e] = S.eoff[e + 1] - S.eoff[e];
        __syncthreads();
      }
    }
    // ---- units: whole rows packed greedily in tile order ----
    if (threadIdx.x == 0) {
      int nu = 0, first = -1, cnt = 0;
      auto close = [&]() {
        if (cnt > 0) {
          if (nu < VR_UMAX) { S.ufirst[nu] = first; S.ucnt[nu] = cnt; }
          ++nu;
        }
        cnt = 0;
      };
      for (int i = 0; i < R; ++i) {
        const int e0 = S.erow[i], nz = S.erow[i + 1] - e0;
        if (!S.own[i] || nz == 0) { close(); continue; }
        if (nz > 32) {
          close();
          for (int x = 0; x < nz; x += 32) { first = e0 + x; cnt = min(32, nz - x); close(); }
          continue;
        }
        if (cnt + nz > 32) close();
        if (cnt == 0) first = e0;
        cnt += nz;
      }
      close();
      S.tmp[34] = nu;
    }
    __syncthreads();
    const int nunit = S.tmp[34];
    if (nunit > VR_UMAX) {
      if (threadIdx.x == 0) atomicExch(error, 2);
      __syncthreads();
      continue;
    }
    for (int u = threadIdx.x; u <= nunit; u += blockDim.x) {
      int len = 0;
      if (u < nunit) {
        const int f = S.ufirst[u], n = S.ucnt[u];
        for (int x = 0; x < n; ++x) len = max(len, S.cnt[f + x]);
        if (len > 0) len = (len + VR_SLACK + 1) & ~1;
        S.ucnt[u] |= len << 8;
      }
      S.ubase[u] = len * 32;
    }
    __syncthreads();
    const int list_total = (smem_exclusive_scan(S.ubase, nunit + 1, S.tmp) + 7) & ~7;
    if (!FILL) {
      if (threadIdx.x == 0) {
        desc[t].nb_unit = nunit;
        desc[t].list_len = list_total;
      }
      __syncthreads();
      continue;
    }
    if (nunit != d.nb_unit || list_total != d.list_len) {
      if (threadIdx.x == 0) atomicExch(error, 3);
      __syncthreads();
      continue;
    }
    // canonical order inside a list: ascending local cell index (= ascending global cell id): fixed summation order
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      uint16_t* l = S.clist + S.eoff[e];
      const int n = S.cnt[e];
      for (int i = 1; i < n; ++i) {
        const uint16_t x = l[i];
        const unsigned kx = ((unsigned)(x & VR_LC_MASK) << 16) | x;
        int j = i - 1;
        while (j >= 0) {
          const uint16_t y = l[j];
          if ((((unsigned)(y & VR_LC_MASK) << 16) | y) <= kx) break;
          l[j + 1] = y;
          --j;
        }
        l[j + 1] = x;
      }
    }
    __syncthreads();
    constexpr uint16_t PAD = (uint16_t)(CS - 1);
    for (int u = threadIdx.x; u < nunit; u += blockDim.x)
      units[d.unit_off + u] = make_uint2((uint32_t)S.ubase[u], vr_pack_unit(S.ufirst[u], S.ucnt[u] & 0xFF, S.ucnt[u] >> 8));
    for (int x = S.ubase[nunit] + threadIdx.x; x < list_total; x += blockDim.x) lists[d.list_off + x] = PAD;
    // the lists in canonical order ([len/2][32 lanes][2], padded); k_bank_order_rows fixes the order inside each list
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const int u = x >> 5, ln = x & 31;
      const int f = S.ufirst[u], n = S.ucnt[u] & 0xFF, len = S.ucnt[u] >> 8;
      const int c = ln < n ? S.cnt[f + ln] : 0;
      const uint16_t* src = S.clist + (ln < n ? S.eoff[f + ln] : 0);
      uint16_t* out = lists + d.list_off + S.ubase[u] + ln * 2;
      for (int k = 0; k < len; ++k) out[(k >> 1) * 64 + (k & 1)] = k < c ? src[k] : PAD;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool tiled_cells_supported(const afb_ctx* ctx) { return ctx->npc == ctx->dim + 1; } // P1 simplices (4 nodes in 2-D is a quadrilateral)

template <class F> static int launch_by_npc(int npc, F f)
{
  if (npc == 4) return f(std::integral_constant<int, 4>());
  return f(std::integral_constant<int, 3>());
}

// list slots of a tile's value plan: contributions + per-unit padding up to the unit's longest list (<= valence + 1, even)
static inline int64_t value_plan_tile_slots(const TileDesc& d, int npc, bool vec)
{
  const int64_t cap = (int64_t)(vec ? npc * npc : npc * (npc - 1)) * d.nb_cell + 32ll * (d.max_val + 2) + 32ll * ((d.nb_entry + 31) / 32) * 2 + 8;
  return (cap + 7) & ~7ll; // a tile whose lists exceed the executor's staging buffer (TG_LMAX) is read from global memory
}
// the arrays of the value plan (entries-by-length plans: scalar executor and the "units" vector executor) in one allocation
static int reserve_value_plan(TilePlan& P, int32_t nb_node, bool vec, int64_t units, int64_t slots)
{
  return reserve_group(P.arena_lists, { { &P.rowinfo, sizeof(uint32_t) * (size_t)nb_node },
                                        { &P.unit_base, sizeof(uint32_t) * (size_t)std::max<int64_t>(units, 1) },
                                        { &P.unit_len, sizeof(uint16_t) * (size_t)std::max<int64_t>(units, 1) },
                                        { &P.emap, sizeof(uint32_t) * 32 * (size_t)std::max<int64_t>(units, 1) },
                                        { &P.emap_rows, vec ? sizeof(uint32_t) * 32 * (size_t)std::max<int64_t>(units, 1) : 16 },
                                        { &P.lists, sizeof(uint16_t) * (size_t)std::max<int64_t>(slots, 8) } });
}

// AFB_INSPECTOR_TRACE=1: host-side timeline of the inspector on stderr (microseconds since the call started)
struct InspectorTrace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  explicit InspectorTrace(const char* what) : on(getenv("AFB_INSPECTOR_TRACE") != nullptr), t0(std::chrono::steady_clock::now())
  {
    if (on) fprintf(stderr, "[inspector] %s\n", what);
  }
  void mark(const char* label) const
  {
    if (on) fprintf(stderr, "[inspector]   %8.1f us  %s\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(), label);
  }
};

int build_tile_mesh(afb_ctx* ctx, int cls)
{
  afb::NvtxRange nvtx_range("TileInspector(mesh)");
  InspectorTrace tr("mesh tiling");
  AFB_REQUIRE(tiled_cells_supported(ctx), AFB_ERR_UNSUPPORTED, "the tiled path is not available for %d-node cells (P1 simplices only); use AFB_VARIANT_NODEWISE", ctx->npc);
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "tile inspector: build the pattern first");
  TilePlan& P = ctx->plan;
  P.mesh_valid = false;
  P.lists_valid = false;
  cudaStream_t st = ctx->stream;
  const int32_t nb_node = ctx->nb_node;
  const int dim = ctx->dim, npc = ctx->npc;
  const bool vec = ctx->b > 1;
  if (cls < 0) cls = !vec ? 0 : (ctx->vec_rows() ? 2 : 1);
  const int cmax = cls == 2 ? VR_CMAX : cls == 1 ? TV_CMAX : TG_CMAX;
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
  tr.mark("bounding box (sync)");
  double lo[3], ext[3];
  for (int a = 0; a < 3; ++a) {
    lo[a] = unorder_f64(got[a]);
    ext[a] = unorder_f64(got[3 + a]) - lo[a];
    if (!(ext[a] > 0.0) || a >= dim) ext[a] = 0.0;
  }
  AFB_TRY(reserve_group(P.arena_nodes, { { &P.node_tile, sizeof(int32_t) * (size_t)nb_node }, { &P.node_lrow, sizeof(int32_t) * (size_t)nb_node },
                                         { &P.tile_nodes, sizeof(int32_t) * (size_t)nb_node }, { &P.scratch_c, sizeof(int32_t) * (size_t)nb_node } })); // scratch_c: brick_of
  int32_t* brick_of = P.scratch_c.as<int32_t>();
  tr.mark("node arrays reserved");

  // bricks holding ~rtarget nodes on a uniform mesh
  const int rtarget = dim == 3 ? (cls == 2 ? VR_RT3 : cls == 1 ? TV_RT3 : TG_RT3) : (cls == 2 ? VR_RT2 : cls == 1 ? TV_RT2 : TG_RT2);
  double vol = 1.0;
  int nd_ext = 0;
  for (int a = 0; a < 3; ++a)
    if (ext[a] > 0.0) { vol *= ext[a]; ++nd_ext; }
  BrickGrid bg;
  const double h = nd_ext ? pow(vol * (double)rtarget / (double)nb_node, 1.0 / nd_ext) : 1.0;
  int64_t nb_brick64 = 1;
  for (int a = 0; a < 3; ++a) {
    bg.x0[a] = lo[a];
    bg.g[a] = ext[a] > 0.0 ? std::max(1, (int)ceil(ext[a] / h)) : 1;
    bg.inv_h[a] = ext[a] > 0.0 ? (double)bg.g[a] / ext[a] : 0.0;
    nb_brick64 *= bg.g[a];
  }
  AFB_REQUIRE(nb_brick64 < (1ll << 28), AFB_ERR_UNSUPPORTED, "tile inspector: brick grid too large");
  const int32_t nb_brick = (int32_t)nb_brick64;
  AFB_TRY(P.scratch_a.reserve(sizeof(int32_t) * (size_t)(4 * ((size_t)nb_brick + 2))));
  int32_t* bcount = P.scratch_a.as<int32_t>();
  int32_t* bptr = bcount + (nb_brick + 2);
  int32_t* bntile = bptr + (nb_brick + 2);
  int32_t* bfirst = bntile + (nb_brick + 2);
  AFB_CUDA(cudaMemsetAsync(bcount, 0, sizeof(int32_t) * (size_t)(nb_brick + 2), st));
  k_brick_assign<<<grid_for(nb_node, 256), 256, 0, st>>>(ctx->coords.as<double>(), nb_node, bg, brick_of, bcount);
  AFB_LAUNCH_CHECK(ctx);
  AFB_TRY(exclusive_scan_i32(ctx, bcount, bptr, nb_brick));
  AFB_CUDA(cudaMemsetAsync(bcount, 0, sizeof(int32_t) * (size_t)(nb_brick + 2), st));
  k_brick_fill<<<grid_for(nb_node, 256), 256, 0, st>>>(brick_of, nb_node, bptr, bcount, P.tile_nodes.as<int32_t>());
  AFB_LAUNCH_CHECK(ctx);
  k_brick_sort<<<nb_brick, 256, 0, st>>>(bptr, nb_brick, P.tile_nodes.as<int32_t>());
  AFB_LAUNCH_CHECK(ctx);
  std::vector<int32_t> hptr((size_t)nb_brick + 1), hnt((size_t)nb_brick), hfirst((size_t)nb_brick + 1);
  AFB_CUDA(cudaMemcpyAsync(hptr.data(), bptr, sizeof(int32_t) * ((size_t)nb_brick + 1), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  tr.mark("bricks (sync)");
  // pieces per brick: start from the row limit, then refine the bricks whose tiles do not fit
  const int rmax0 = std::min(TG_RMAX, rtarget + rtarget / 2);
  for (int32_t b = 0; b < nb_brick; ++b) {
    const int n = hptr[b + 1] - hptr[b];
    hnt[b] = n > 0 ? (n + rmax0 - 1) / rmax0 : 0;
  }
  std::vector<TileDesc> hdesc;
  std::vector<int32_t> hstats;
  int32_t nb_tile = 0;
  for (int attempt = 0;; ++attempt) {
    AFB_REQUIRE(attempt < 64, AFB_ERR_UNSUPPORTED, "tile inspector: refinement did not converge");
    int64_t run = 0;
    for (int32_t b = 0; b < nb_brick; ++b) {
      hfirst[b] = (int32_t)run;
      run += hnt[b];
    }
    hfirst[nb_brick] = (int32_t)run;
    AFB_REQUIRE(run < (1ll << 30), AFB_ERR_OVERFLOW, "tile inspector: too many tiles");
    nb_tile = (int32_t)run;
    AFB_CUDA(cudaMemcpyAsync(bntile, hnt.data(), sizeof(int32_t) * (size_t)nb_brick, cudaMemcpyHostToDevice, st));
    AFB_CUDA(cudaMemcpyAsync(bfirst, hfirst.data(), sizeof(int32_t) * ((size_t)nb_brick + 1), cudaMemcpyHostToDevice, st));
    AFB_TRY(reserve_group(P.arena_desc, { { &P.tile_desc, sizeof(TileDesc) * (size_t)std::max(nb_tile, 1) }, { &P.scratch_b, sizeof(int32_t) * 4 * (size_t)std::max(nb_tile, 1) } }));
    AFB_CUDA(cudaMemsetAsync(P.tile_desc.p, 0, sizeof(TileDesc) * (size_t)std::max(nb_tile, 1), st));
    k_tile_nodes<<<grid_for(nb_node, 256), 256, 0, st>>>(brick_of, bptr, bfirst, bntile, P.tile_nodes.as<int32_t>(), nb_node, P.node_tile.as<int32_t>(),
                                                          P.node_lrow.as<int32_t>(), P.tile_desc.as<TileDesc>());
    AFB_LAUNCH_CHECK(ctx);
    int32_t* stats = P.scratch_b.as<int32_t>();
    AFB_TRY(launch_by_npc(npc, [&](auto N) {
      k_tile_stats<decltype(N)::value><<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(),
                                                                  ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), P.node_tile.as<int32_t>(),
                                                                  P.node_lrow.as<int32_t>(), stats);
      return AFB_OK;
    }));
    AFB_LAUNCH_CHECK(ctx);
    hdesc.resize(nb_tile);
    hstats.resize(4 * (size_t)nb_tile);
    AFB_CUDA(cudaMemcpyAsync(hdesc.data(), P.tile_desc.p, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaMemcpyAsync(hstats.data(), stats, sizeof(int32_t) * 4 * (size_t)nb_tile, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    tr.mark("tile statistics (sync)");
    bool ok = true;
    for (int32_t b = 0; b < nb_brick; ++b) {
      bool bad = false;
      for (int32_t t = hfirst[b]; t < hfirst[b + 1]; ++t) {
        const int C = hstats[4 * t], E = hstats[4 * t + 1], F = hstats[4 * t + 3];
        if (C > cmax || E > TG_EMAX || hdesc[t].nb_row > TG_RMAX || F > TG_FMAX) bad = true;
      }
      if (bad) {
        const int n = hptr[b + 1] - hptr[b];
        AFB_REQUIRE(hnt[b] < n, AFB_ERR_UNSUPPORTED,
                    "tiled path: a single row exceeds the tile limits (%d cells / %d entries / %d footprint nodes); use AFB_VARIANT_NODEWISE", cmax, TG_EMAX, TG_FMAX);
        hnt[b] = std::min(n, hnt[b] + std::max(1, hnt[b] / 2));
        ok = false;
      }
    }
    if (ok) break;
  }
  P.nb_tile = nb_tile;
  // sizes -> offsets
  int64_t cell_off = 0, foot_off = 0, inc_off = 0, ent_off = 0;
  int max_rows = 0;
  for (int32_t t = 0; t < nb_tile; ++t) {
    TileDesc& d = hdesc[t];
    const int C = hstats[4 * t], E = hstats[4 * t + 1], V = hstats[4 * t + 2], F = hstats[4 * t + 3];
    d.cell_off = (int32_t)cell_off;
    d.nb_cell = C;
    d.foot_off = (int32_t)foot_off;
    d.nb_foot = F;
    d.inc_off = (uint32_t)inc_off;
    d.nb_group = (d.nb_row + 31) / 32;
    d.nb_entry = E;
    d.ent_off = (uint32_t)ent_off;
    ent_off += E;
    d.max_val = V;
    d.unit_off = d.nb_unit = 0;
    d.list_off = 0;
    d.list_len = 0;
    cell_off += C;
    foot_off += F;
    inc_off += (int64_t)d.nb_group * 32 * V;
    max_rows = std::max(max_rows, d.nb_row);
    AFB_REQUIRE(inc_off < (1ll << 32) && cell_off < (1ll << 31) && foot_off < (1ll << 31), AFB_ERR_OVERFLOW, "tile inspector: plan exceeds 32-bit offsets");
  }
  P.max_rows = max_rows;
  P.nb_tile_cell = cell_off;
  P.nb_foot = foot_off;
  P.nb_inc = inc_off;
  P.nb_entry = ent_off;
  tr.mark("offsets (host)");
  AFB_TRY(reserve_group(P.arena_tiles, { { &P.tile_cells, sizeof(int32_t) * (size_t)std::max<int64_t>(cell_off, 1) },
                                         { &P.lconn, sizeof(ushort4) * (size_t)std::max<int64_t>(cell_off, 1) },
                                         { &P.foot, sizeof(int32_t) * (size_t)std::max<int64_t>(foot_off, 1) },
                                         { &P.rowf, sizeof(uint16_t) * (size_t)nb_node },
                                         { &P.inc, sizeof(uint32_t) * (size_t)std::max<int64_t>(inc_off, 1) },
                                         { &P.inc_grp, sizeof(uint2) * (size_t)TG_GMAX * (size_t)std::max(nb_tile, 1) } }));
  tr.mark("tile arrays reserved");
  AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc.data(), sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
  if (nb_tile > 0) {
    const int grid = std::min<int>(nb_tile, 4 * ctx->sm_count);
    AFB_TRY(launch_by_npc(npc, [&](auto N) {
      k_tile_mesh<decltype(N)::value><<<grid, TB_THREADS, 0, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(),
                                                                    ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(), P.node_tile.as<int32_t>(), P.node_lrow.as<int32_t>(),
                                                                    P.tile_cells.as<int32_t>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(), P.rowf.as<uint16_t>(),
                                                                    P.inc.as<uint32_t>(), P.inc_grp.as<uint2>(), ctx->tmp_flag.as<int>());
      return AFB_OK;
    }));
    AFB_LAUNCH_CHECK(ctx);
  }
  int err = 0;
  AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaEventRecord(e1, st));
  AFB_CUDA(cudaEventSynchronize(e1));
  AFB_CUDA(cudaEventElapsedTime(&P.mesh_ms, e0, e1));
  tr.mark("k_tile_mesh (sync)");
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  AFB_REQUIRE(err == 0, AFB_ERR_CUDA, "tile inspector: mesh tiling inconsistency (code %d)", err);
  P.hdesc_host.assign(reinterpret_cast<const int32_t*>(hdesc.data()), reinterpret_cast<const int32_t*>(hdesc.data()) + 16 * (size_t)nb_tile);
  P.mesh_gen = ctx->mesh_gen;
  P.mesh_b_class = cls;
  P.mesh_valid = true;
  AFB_TRY(pattern_nn_build(ctx)); // tile-local node-node connectivity of the connectivity-based BuildMatrix
  tr.mark("node-node connectivity (queued)");
  return AFB_OK;
}

int build_tile_lists(afb_ctx* ctx, int mode_flags)
{
  afb::NvtxRange nvtx_range("TileInspector(values)");
  TilePlan& P = ctx->plan;
  AFB_REQUIRE(P.mesh_valid, AFB_ERR_INVALID, "tile inspector: no mesh tiling");
  P.lists_valid = false;
  InspectorTrace tr("value plan");
  cudaStream_t st = ctx->stream;
  const int npc = ctx->npc;
  const bool vec = ctx->b > 1;
  const int32_t nb_tile = P.nb_tile;
  cudaEvent_t e0, e1;
  AFB_CUDA(cudaEventCreate(&e0));
  AFB_CUDA(cudaEventCreate(&e1));
  AFB_CUDA(cudaEventRecord(e0, st));
  TileDesc* hdesc = reinterpret_cast<TileDesc*>(P.hdesc_host.data());
  int64_t unit_off = 0, list_off = 0;
  for (int32_t t = 0; t < nb_tile; ++t) {
    TileDesc& d = hdesc[t];
    d.unit_off = (int32_t)unit_off;
    d.nb_unit = (d.nb_entry + 31) / 32; // upper bound; the builder stores the number of units of computed entries
    d.list_off = (uint32_t)list_off;
    d.list_len = 0;
    unit_off += d.nb_unit;
    list_off += value_plan_tile_slots(d, npc, vec);
    AFB_REQUIRE(list_off < (1ll << 32) && unit_off < (1ll << 26), AFB_ERR_OVERFLOW, "tile inspector: value plan exceeds 32-bit offsets");
  }
  P.nb_unit = unit_off;
  P.nb_list = list_off;
  AFB_TRY(reserve_value_plan(P, ctx->nb_node, vec, unit_off, list_off));
  tr.mark("list arrays reserved");
  AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
  const uint8_t* own = (ctx->all_own || (mode_flags & AFB_FLAG_ALL_ROWS)) ? nullptr : ctx->is_own.as<uint8_t>();
  const int64_t nb_own_cell = (mode_flags & AFB_FLAG_OWN_CELLS_ONLY) ? ctx->nb_own_cell : ctx->nb_cell;
  if (nb_tile > 0) {
    const size_t smem = sizeof(ListBuilderSmem);
    const int grid = std::min<int>(nb_tile, 2 * ctx->sm_count);
    auto go = [&](auto kernel, int list_max) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kernel<<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), P.tile_cells.as<int32_t>(), ctx->conn.as<int32_t>(),
                                             ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(), P.node_lrow.as<int32_t>(), own, nb_own_cell,
                                             P.rowinfo.as<uint32_t>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.emap.as<uint32_t>(), P.emap_rows.as<uint32_t>(), P.lists.as<uint16_t>(),
                                             list_max, ctx->tmp_flag.as<int>());
      return cudaGetLastError();
    };
    cudaError_t e;
    if (npc == 4) e = vec ? go(k_tile_lists<4, true>, 1 << 30) : go(k_tile_lists<4, false>, 1 << 30);
    else e = vec ? go(k_tile_lists<3, true>, 1 << 30) : go(k_tile_lists<3, false>, 1 << 30);
    AFB_CUDA(e);
    // order inside the lists (shared-memory banks of the executor's gathers): one thread per half-unit, whole grid
    k_bank_order<<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.lists.as<uint16_t>(),
                                          (uint16_t)(vec ? (TV_CS - 1) : TG_ZERO));
    AFB_LAUNCH_CHECK(ctx);
    ctx->launches += 2;
  }
  int err = 0;
  AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaEventRecord(e1, st));
  AFB_CUDA(cudaEventSynchronize(e1));
  AFB_CUDA(cudaEventElapsedTime(&P.lists_ms, e0, e1));
  tr.mark("k_tile_lists + k_bank_order (sync)");
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  AFB_REQUIRE(err == 0, AFB_ERR_CUDA, "tile inspector: value plan inconsistency (code %d)", err);
  P.lists_b = ctx->b;
  P.lists_mode = mode_flags & (AFB_FLAG_ALL_ROWS | AFB_FLAG_OWN_CELLS_ONLY);
  P.lists_mesh_gen = ctx->mesh_gen;
  P.lists_kind = 0;
  P.lists_valid = true;
  return AFB_OK;
}

int build_tile_rowlists(afb_ctx* ctx, int mode_flags)
{
  afb::NvtxRange nvtx_range("TileInspector(values)");
  TilePlan& P = ctx->plan;
  AFB_REQUIRE(P.mesh_valid && P.mesh_b_class == 2, AFB_ERR_INVALID, "tile inspector: no mesh tiling for the row-ordered executor");
  P.lists_valid = false;
  cudaStream_t st = ctx->stream;
  const int npc = ctx->npc;
  const int32_t nb_tile = P.nb_tile;
  cudaEvent_t e0, e1;
  AFB_CUDA(cudaEventCreate(&e0));
  AFB_CUDA(cudaEventCreate(&e1));
  AFB_CUDA(cudaEventRecord(e0, st));
  TileDesc* hdesc = reinterpret_cast<TileDesc*>(P.hdesc_host.data());
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
  const uint8_t* own = (ctx->all_own || (mode_flags & AFB_FLAG_ALL_ROWS)) ? nullptr : ctx->is_own.as<uint8_t>();
  const int64_t nb_own_cell = (mode_flags & AFB_FLAG_OWN_CELLS_ONLY) ? ctx->nb_own_cell : ctx->nb_cell;
  AFB_TRY(P.rowinfo.reserve(sizeof(uint32_t) * (size_t)ctx->nb_node));
  const size_t smem = sizeof(RowListSmem);
  const int grid = std::min<int>(std::max(nb_tile, 1), 2 * ctx->sm_count);
  auto go = [&](auto kernel) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), P.tile_cells.as<int32_t>(), ctx->conn.as<int32_t>(),
                                           ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(), P.node_lrow.as<int32_t>(), own, nb_own_cell,
                                           P.rowinfo.as<uint32_t>(), P.vr_units.as<uint2>(), P.lists.as<uint16_t>(), ctx->tmp_flag.as<int>());
    return cudaGetLastError();
  };
  int64_t unit_off = 0, list_off = 0;
  if (nb_tile > 0) {
    // pass 1: sizes
    AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
    AFB_CUDA(npc == 4 ? go(k_tile_rowlists<4, false>) : go(k_tile_rowlists<3, false>));
    ctx->launches++;
    AFB_CUDA(cudaMemcpyAsync(hdesc, P.tile_desc.p, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    for (int32_t t = 0; t < nb_tile; ++t) {
      TileDesc& d = hdesc[t];
      d.unit_off = (int32_t)unit_off;
      d.list_off = (uint32_t)list_off;
      unit_off += d.nb_unit;
      list_off += d.list_len;
      AFB_REQUIRE(list_off < (1ll << 32) && unit_off < (1ll << 30), AFB_ERR_OVERFLOW, "tile inspector: value plan exceeds 32-bit offsets");
    }
  }
  P.nb_unit = unit_off;
  P.nb_list = list_off;
  AFB_TRY(P.vr_units.reserve(sizeof(uint2) * (size_t)std::max<int64_t>(unit_off, 1)));
  AFB_TRY(P.lists.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(list_off, 8)));
  if (nb_tile > 0) {
    // pass 2: unit records and lists in canonical order; then the order inside the lists (one thread per half-unit)
    AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc, sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
    AFB_CUDA(npc == 4 ? go(k_tile_rowlists<4, true>) : go(k_tile_rowlists<3, true>));
    if (npc == 4) k_bank_order_rows<4><<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), P.vr_units.as<uint2>(), P.lists.as<uint16_t>());
    else k_bank_order_rows<3><<<nb_tile, 128, 0, st>>>(P.tile_desc.as<TileDesc>(), P.vr_units.as<uint2>(), P.lists.as<uint16_t>());
    AFB_LAUNCH_CHECK(ctx);
    ctx->launches += 2;
  }
  int err = 0;
  AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaEventRecord(e1, st));
  AFB_CUDA(cudaEventSynchronize(e1));
  AFB_CUDA(cudaEventElapsedTime(&P.lists_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  AFB_REQUIRE(err != 2, AFB_ERR_UNSUPPORTED, "tiled path: a tile has more than %d units of rows; use AFB_VARIANT_NODEWISE", VR_UMAX);
  AFB_REQUIRE(err == 0, AFB_ERR_CUDA, "tile inspector: value plan inconsistency (code %d)", err);
  P.lists_b = ctx->b;
  P.lists_mode = mode_flags & (AFB_FLAG_ALL_ROWS | AFB_FLAG_OWN_CELLS_ONLY);
  P.lists_mesh_gen = ctx->mesh_gen;
  P.lists_kind = 1;
  P.lists_valid = true;
  return AFB_OK;
}

} // namespace afb
