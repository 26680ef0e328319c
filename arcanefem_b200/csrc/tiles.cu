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
constexpr int TG_CMAX = 2080;              // cells per tile
constexpr int TG_CS = TG_CMAX + 1;         // cache stride (odd: consecutive pair planes shift banks)
constexpr int TG_KMAX = 10;                // distinct K_e values of a symmetric 4x4
constexpr int TG_ZERO = TG_KMAX * TG_CS;   // cache slot that holds 0.0 (list padding)
constexpr int TG_EMAX = 5120;              // entries per tile
constexpr int TG_RALLOC = 1024;            // rows per tile, allocation bound
constexpr int TG_FMAX = 896;               // footprint nodes per tile (rows + halo), coordinates staged in shared memory
constexpr int TG_LMAX = 20480;             // 16-bit list slots per tile staged in shared memory (40 KB)
constexpr int TG_UMAX = TG_EMAX / 32;      // units per tile
constexpr int TG_ROUNDS = (TG_CMAX + TG_THREADS - 1) / TG_THREADS; // phase-A rounds
constexpr int TB_THREADS = 512;            // builder CTA
constexpr int TB_HASH = 4096;              // halo-node hash slots of the builder
constexpr unsigned TG_NONE = 0xFFFFFFFFu;

struct TileDesc {
  int32_t node_off, nb_row, cell_off, nb_cell;
  int32_t unit_off, nb_unit;
  uint32_t list_off; // first 16-bit slot of the tile in `lists` (multiple of 8)
  int32_t list_len;  // used slots (multiple of 8)
  int32_t foot_off, nb_foot, nb_entry, pad;
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

// open-addressing set of node ids in shared memory (halo nodes of a tile); returns the slot
__device__ __forceinline__ int hash_insert(unsigned* __restrict__ tab, unsigned id)
{
  unsigned h = (id * 0x9E3779B1u) >> 20; // 12 bits = TB_HASH
  while (true) {
    const unsigned old = atomicCAS(tab + h, 0xFFFFFFFFu, id);
    if (old == 0xFFFFFFFFu || old == id) return (int)h;
    h = (h + 1) & (TB_HASH - 1);
  }
}
__device__ __forceinline__ int hash_find(const unsigned* __restrict__ tab, unsigned id)
{
  unsigned h = (id * 0x9E3779B1u) >> 20;
  while (tab[h] != id) h = (h + 1) & (TB_HASH - 1);
  return (int)h;
}

// per tile: number of cells, entries, largest valence, halo nodes
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
  for (int i = threadIdx.x; i < TB_HASH; i += blockDim.x) s_tab[i] = 0xFFFFFFFFu;
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
        // halo nodes; stop inserting once the tile is known to be too large (keeps the table from filling)
        if (s_h <= TG_FMAX) {
#pragma unroll
          for (int a = 0; a < NPC; ++a) {
            const int32_t n = conn[(int64_t)cell * NPC + a];
            if (node_tile[n] != t) {
              unsigned h = ((unsigned)n * 0x9E3779B1u) >> 20;
              while (true) {
                const unsigned old = atomicCAS(s_tab + h, 0xFFFFFFFFu, (unsigned)n);
                if (old == 0xFFFFFFFFu) { atomicAdd(&s_h, 1); break; }
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
    stats[4 * t + 3] = s_h;
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

// ---------------------------------------------------------------------------------------------
// inspector step 2: per-tile plan
// ---------------------------------------------------------------------------------------------
// Entry classes of a tile (row i = local row index, column node c):
//   computed : the diagonal, columns outside the tile, and columns inside the tile with a larger
//              row index ("upper"); the latter also store the value at the mirror position
//              (row of c, column of i) -- the element matrices are symmetric, so only one of the two
//              sums is formed (bitwise symmetric result)
//   mirror   : columns inside the tile with a smaller row index: written by their upper twin
struct BuilderSmem {
  unsigned cells[4096];          // tile cells (sorted ascending), padded to a power of two
  int erow_off[TG_RALLOC + 1];   // first entry of each tile row
  int cnt[TG_EMAX];              // contributions per entry (-1: mirror entry)
  int eoff[TG_EMAX + 1];         // start of each entry's list in clist
  unsigned egpos[TG_EMAX];       // entry -> index into values
  unsigned egpos2[TG_EMAX];      // entry -> mirror index into values, or TG_NONE
  unsigned keys[8192];           // entries sorted by count (descending); first used as the halo hash (4096 + 4096)
  uint16_t clist[16 * TG_CMAX];  // contribution codes
  int ulen[TG_UMAX + 1], ubase[TG_UMAX + 2];
  int tmp[40];
  int nb_cell;
};

template <int NPC>
__global__ void __launch_bounds__(TB_THREADS, 1)
k_tile_build(TileDesc* __restrict__ desc, int32_t nb_tile, const int32_t* __restrict__ tnodes, const int32_t* __restrict__ conn, const int32_t* __restrict__ nc_ptr,
             const int32_t* __restrict__ nc_list, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ node_tile,
             const int32_t* __restrict__ node_lrow, int32_t* __restrict__ tile_cells, int32_t* __restrict__ foot, ushort4* __restrict__ lconn,
             uint32_t* __restrict__ unit_base, uint16_t* __restrict__ unit_len, uint32_t* __restrict__ gpos, uint32_t* __restrict__ gpos2,
             uint16_t* __restrict__ lists, int64_t list_capacity_check, int* __restrict__ error)
{
  extern __shared__ unsigned char tb_raw[];
  BuilderSmem& S = *reinterpret_cast<BuilderSmem*>(tb_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  (void)list_capacity_check;
  for (int32_t t = blockIdx.x; t < nb_tile; t += gridDim.x) {
    const TileDesc d = desc[t];
    const int R = d.nb_row;
    unsigned* htab = S.keys;                            // halo hash: node ids
    int* hidx = reinterpret_cast<int*>(S.keys + TB_HASH); // slot -> halo index
    // ---- leader cells -> S.cells (then sorted ascending) ----
    if (threadIdx.x == 0) S.nb_cell = 0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
      S.cells[i] = 0xFFFFFFFFu;
      htab[i] = 0xFFFFFFFFu;
    }
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
    // ---- footprint: rows first, then the halo nodes (hash set -> dense indices) ----
    for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
      tile_cells[d.cell_off + lc] = (int32_t)S.cells[lc];
      const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
#pragma unroll
      for (int a = 0; a < NPC; ++a) {
        const int32_t n = __ldg(cn + a);
        if (__ldg(node_tile + n) != t) hash_insert(htab, (unsigned)n);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TB_HASH; i += blockDim.x) hidx[i] = htab[i] != 0xFFFFFFFFu ? 1 : 0;
    __syncthreads();
    const int H = smem_exclusive_scan(hidx, TB_HASH, S.tmp);
    if (R + H != d.nb_foot || R + H > TG_FMAX) {
      if (threadIdx.x == 0) atomicExch(error, 4);
      __syncthreads();
      continue;
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) foot[d.foot_off + i] = tnodes[d.node_off + i];
    for (int i = threadIdx.x; i < TB_HASH; i += blockDim.x)
      if (htab[i] != 0xFFFFFFFFu) foot[d.foot_off + R + hidx[i]] = (int32_t)htab[i];
    for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
      const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
      unsigned short loc[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int a = 0; a < NPC; ++a) {
        const int32_t n = __ldg(cn + a);
        loc[a] = (unsigned short)(__ldg(node_tile + n) == t ? __ldg(node_lrow + n) : R + hidx[hash_find(htab, (unsigned)n)]);
      }
      lconn[d.cell_off + lc] = make_ushort4(loc[0], loc[1], loc[2], loc[3]);
    }
    // ---- entries: class, value positions ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = tnodes[d.node_off + i];
      const int rb = rows[r], deg = rows[r + 1] - rb, e0 = S.erow_off[i];
      for (int p = lane; p < deg; p += 32) {
        const int32_t c = cols[rb + p];
        int cnt0 = 0;
        unsigned g2 = TG_NONE;
        if (c != r && __ldg(node_tile + c) == t) {
          const int j = __ldg(node_lrow + c);
          if (j < i) cnt0 = -1; // mirror entry
          else {
            const int cb = rows[c], ce = rows[c + 1];
            g2 = (unsigned)find_col(cols, cb, ce, r);
          }
        }
        S.cnt[e0 + p] = cnt0;
        S.egpos[e0 + p] = (unsigned)(rb + p);
        S.egpos2[e0 + p] = g2;
      }
    }
    __syncthreads(); // also: every reader of the halo hash is done before keys are reused
    // ---- contribution lists of the computed entries (count, scan, fill) ----
    for (int pass = 0; pass < 2; ++pass) {
      for (int lc = threadIdx.x; lc < C; lc += blockDim.x) {
        const int32_t* cn = conn + (int64_t)S.cells[lc] * NPC;
        int32_t nd[NPC];
        int li[NPC];
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          nd[a] = __ldg(cn + a);
          li[a] = __ldg(node_tile + nd[a]) == t ? __ldg(node_lrow + nd[a]) : -1;
        }
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          if (li[a] < 0) continue;
          const int rb = __ldg(rows + nd[a]), re = __ldg(rows + nd[a] + 1);
#pragma unroll
          for (int bq = 0; bq < NPC; ++bq) {
            if (bq != a && li[bq] >= 0 && li[bq] < li[a]) continue; // the twin entry (li[bq], li[a]) takes it
            const int e = S.erow_off[li[a]] + (find_col(cols, rb, re, nd[bq]) - rb);
            if (pass == 0) atomicAdd(&S.cnt[e], 1);
            else {
              const int slot = atomicSub(&S.cnt[e], 1) - 1; // countdown cursor, restored from eoff below
              S.clist[S.eoff[e] + slot] = (uint16_t)(sym_pair(a, bq) * TG_CS + lc);
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
    // ---- computed entries by descending count, cut into units of 32; mirror entries sort last ----
    int e2 = 32;
    while (e2 < E) e2 <<= 1;
    for (int e = threadIdx.x; e < e2; e += blockDim.x)
      S.keys[e] = (e < E && S.cnt[e] >= 0) ? (((unsigned)(0xFFFF - min(S.cnt[e], 0xFFFF)) << 16) | (unsigned)e) : 0xFFFFFFFFu;
    if (threadIdx.x == 0) S.tmp[34] = 0;
    __syncthreads();
    {
      int mine = 0;
      for (int e = threadIdx.x; e < E; e += blockDim.x) mine += S.cnt[e] >= 0 ? 1 : 0;
      atomicAdd(&S.tmp[34], mine);
    }
    smem_bitonic_sort(S.keys, e2);
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
    if (nunit > d.nb_unit || list_total > TG_LMAX) {
      if (threadIdx.x == 0) atomicExch(error, list_total > TG_LMAX ? 3 : 2);
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
    // padding of the last 16-byte group
    for (int x = S.ubase[nunit] + threadIdx.x; x < list_total; x += blockDim.x) lists[d.list_off + x] = (uint16_t)TG_ZERO;
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const int u = x >> 5, l = x & 31;
      const bool valid = x < EC;
      const int e = valid ? (int)(S.keys[x] & 0xFFFFu) : 0;
      gpos[(size_t)(d.unit_off + u) * 32 + l] = valid ? S.egpos[e] : TG_NONE;
      gpos2[(size_t)(d.unit_off + u) * 32 + l] = valid ? S.egpos2[e] : TG_NONE;
      const int len = S.ulen[u], n = valid ? S.cnt[e] : 0;
      uint16_t* out = lists + d.list_off + S.ubase[u] + l * 2; // [len/2][32 lanes][2]
      const uint16_t* src = S.clist + (valid ? S.eoff[e] : 0);
      for (int k = 0; k < len; ++k) out[(k >> 1) * 64 + (k & 1)] = k < n ? src[k] : (uint16_t)TG_ZERO;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// executor
// ---------------------------------------------------------------------------------------------
// K_e of one P1 cell from the coordinates staged in shared memory (AoS, 3 doubles per footprint node)
template <int NPC> struct SymK;
template <> struct SymK<4> {
  __device__ static __forceinline__ void compute(const double* __restrict__ cx, ushort4 ln, const ElemParams&, double (&K)[10])
  {
    const double* p0 = cx + 3 * ln.x;
    const double* p1 = cx + 3 * ln.y;
    const double* p2 = cx + 3 * ln.z;
    const double* p3 = cx + 3 * ln.w;
    Tet4Geom g;
    g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    int p = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) K[p++] = g.dot(a, b) * g.s;
  }
};
template <> struct SymK<3> {
  // pair indexing of a 4-node cell is reused; slots with a or b == 3 stay unused
  __device__ static __forceinline__ void compute(const double* __restrict__ cx, ushort4 ln, const ElemParams& prm, double (&K)[10])
  {
    const double* p0 = cx + 3 * ln.x;
    const double* p1 = cx + 3 * ln.y;
    const double* p2 = cx + 3 * ln.z;
    Tri3Geom g;
    g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], (prm.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
#pragma unroll
    for (int p = 0; p < 10; ++p) K[p] = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = a; b < 3; ++b) K[sym_pair(a, b)] = g.dot(a, b) * g.s;
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct ExecSmem {
  double Kc[TG_ZERO + 1];
  double cx[3 * TG_FMAX];
  __align__(16) uint16_t lists[TG_LMAX];
  uint32_t ubase[TG_UMAX];
  uint16_t ulen[TG_UMAX];
  __align__(16) TileDesc desc[3]; // ring: current, next, next-next tile of this CTA
  __align__(8) unsigned long long mbar;
};

// inputs of the next tile a thread carries in registers across phase B
struct TilePrefetch {
  double c0, c1, c2;          // coordinates of footprint node `threadIdx.x`
  ushort4 ln[TG_ROUNDS];      // local connectivity of this thread's cells
  uint32_t ubase;             // unit table entry `threadIdx.x`
  uint16_t ulen;
};

template <int NPC>
__device__ __forceinline__ void prefetch_tile(const TileDesc& d, const double* __restrict__ coords, const int32_t* __restrict__ foot, const ushort4* __restrict__ lconn,
                                              const uint32_t* __restrict__ unit_base, const uint16_t* __restrict__ unit_len, TilePrefetch& pf)
{
  if ((int)threadIdx.x < d.nb_foot) {
    const double* p = coords + 3 * (int64_t)__ldg(foot + d.foot_off + threadIdx.x);
    pf.c0 = __ldg(p);
    pf.c1 = __ldg(p + 1);
    pf.c2 = __ldg(p + 2);
  }
#pragma unroll
  for (int r = 0; r < TG_ROUNDS; ++r) {
    const int lc = r * TG_THREADS + threadIdx.x;
    if (lc < d.nb_cell) pf.ln[r] = __ldg(lconn + d.cell_off + lc);
  }
  if ((int)threadIdx.x < d.nb_unit) {
    pf.ubase = __ldg(unit_base + d.unit_off + threadIdx.x);
    pf.ulen = __ldg(unit_len + d.unit_off + threadIdx.x);
  }
}

template <int NPC>
__global__ void __launch_bounds__(TG_THREADS, 1)
k_assemble_tiled(const TileDesc* __restrict__ desc, int32_t nb_tile, const double* __restrict__ coords, const int32_t* __restrict__ foot,
                 const ushort4* __restrict__ lconn, const uint32_t* __restrict__ unit_base, const uint16_t* __restrict__ unit_len,
                 const uint32_t* __restrict__ gpos, const uint32_t* __restrict__ gpos2, const uint16_t* __restrict__ lists, double* __restrict__ values,
                 int accumulate, ElemParams prm)
{
  extern __shared__ __align__(16) unsigned char ex_raw[];
  ExecSmem& S = *reinterpret_cast<ExecSmem*>(ex_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TG_THREADS / 32;
  constexpr int DW = sizeof(TileDesc) / 4;
  const uint32_t mbar = smem_u32(&S.mbar);
  int32_t t = blockIdx.x;
  if (threadIdx.x == 0) {
    S.Kc[TG_ZERO] = 0.0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 2 * DW) { // descriptors of the first two tiles
    const int k = threadIdx.x / DW, w = threadIdx.x % DW;
    const int64_t tt = (int64_t)t + (int64_t)k * gridDim.x;
    if (tt < nb_tile) reinterpret_cast<int32_t*>(&S.desc[k])[w] = __ldg(reinterpret_cast<const int32_t*>(desc + tt) + w);
  }
  __syncthreads();
  unsigned parity = 0;
  int slot = 0;
  TilePrefetch pf;
  if (t < nb_tile) prefetch_tile<NPC>(S.desc[0], coords, foot, lconn, unit_base, unit_len, pf);
  while (t < nb_tile) {
    const TileDesc d = S.desc[slot];
    // stage: coordinates and unit tables (registers -> shared) and -- asynchronously, by the TMA
    // engine -- the tile's contribution lists, which land while phase A computes
    if ((int)threadIdx.x < d.nb_foot) {
      S.cx[3 * threadIdx.x] = pf.c0;
      S.cx[3 * threadIdx.x + 1] = pf.c1;
      S.cx[3 * threadIdx.x + 2] = pf.c2;
    }
    if ((int)threadIdx.x < d.nb_unit) {
      S.ubase[threadIdx.x] = pf.ubase;
      S.ulen[threadIdx.x] = pf.ulen;
    }
    if (threadIdx.x == 0 && d.list_len > 0) {
      const uint32_t bytes = (uint32_t)d.list_len * 2u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.lists)),
                   "l"(lists + d.list_off), "r"(bytes), "r"(mbar)
                   : "memory");
    }
    __syncthreads();
    // phase A: element matrices of the tile's cells, once each
    if (!(prm.flags & (1 << 16))) {
#pragma unroll
      for (int r = 0; r < TG_ROUNDS; ++r) {
        const int lc = r * TG_THREADS + threadIdx.x;
        if (lc < d.nb_cell) {
          double K[10];
          SymK<NPC>::compute(S.cx, pf.ln[r], prm, K);
#pragma unroll
          for (int p = 0; p < 10; ++p)
            if (NPC == 4 || (p != 3 && p != 6 && p != 8 && p != 9)) S.Kc[p * TG_CS + lc] = K[p];
        }
      }
    }
    // the next tile's inputs and the descriptor after it travel while phase B runs
    const int64_t tn = (int64_t)t + gridDim.x, tnn = tn + gridDim.x;
    const int nslot = slot == 2 ? 0 : slot + 1, nnslot = nslot == 2 ? 0 : nslot + 1;
    if (tn < nb_tile) prefetch_tile<NPC>(S.desc[nslot], coords, foot, lconn, unit_base, unit_len, pf);
    if (threadIdx.x < DW && tnn < nb_tile) reinterpret_cast<int32_t*>(&S.desc[nnslot])[threadIdx.x] = __ldg(reinterpret_cast<const int32_t*>(desc + tnn) + threadIdx.x);
    // first unit's destinations, requested before the barrier
    int u = warp;
    uint32_t g = TG_NONE, g2 = TG_NONE;
    if (u < d.nb_unit) {
      g = __ldg(gpos + (size_t)(d.unit_off + u) * 32 + lane);
      g2 = __ldg(gpos2 + (size_t)(d.unit_off + u) * 32 + lane);
    }
    __syncthreads();
    if (d.list_len > 0) {
      // wait for the bulk copy of this tile's lists
      unsigned done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(mbar), "r"(parity)
                     : "memory");
      }
      parity ^= 1u;
    }
    // phase B: one warp per unit of 32 entries with equally long contribution lists
    if (!(prm.flags & (1 << 17))) {
      const uint32_t* l32 = reinterpret_cast<const uint32_t*>(S.lists);
      while (u < d.nb_unit) {
        const int un = u + NW;
        uint32_t gn = TG_NONE, g2n = TG_NONE;
        if (un < d.nb_unit) {
          gn = __ldg(gpos + (size_t)(d.unit_off + un) * 32 + lane);
          g2n = __ldg(gpos2 + (size_t)(d.unit_off + un) * 32 + lane);
        }
        const uint32_t* l = l32 + (S.ubase[u] >> 1) + lane;
        const int len2 = S.ulen[u] >> 1;
        double acc0 = 0.0, acc1 = 0.0;
        int k = 0;
        for (; k + 4 <= len2; k += 4) {
          const uint32_t i0 = l[(k + 0) * 32], i1 = l[(k + 1) * 32], i2 = l[(k + 2) * 32], i3 = l[(k + 3) * 32];
          acc0 += S.Kc[i0 & 0xFFFFu]; acc1 += S.Kc[i0 >> 16];
          acc0 += S.Kc[i1 & 0xFFFFu]; acc1 += S.Kc[i1 >> 16];
          acc0 += S.Kc[i2 & 0xFFFFu]; acc1 += S.Kc[i2 >> 16];
          acc0 += S.Kc[i3 & 0xFFFFu]; acc1 += S.Kc[i3 >> 16];
        }
        for (; k < len2; ++k) {
          const uint32_t i0 = l[k * 32];
          acc0 += S.Kc[i0 & 0xFFFFu]; acc1 += S.Kc[i0 >> 16];
        }
        if (g != TG_NONE) {
          const double v = acc0 + acc1;
          if (accumulate) {
            values[g] += v;
            if (g2 != TG_NONE) values[g2] += v;
          }
          else {
            values[g] = v;
            if (g2 != TG_NONE) values[g2] = v;
          }
        }
        u = un;
        g = gn;
        g2 = g2n;
      }
    }
    __syncthreads();
    t = (int32_t)tn;
    slot = nslot;
    if (tn >= nb_tile) break;
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

  // rows per tile: start from what the cache can hold on a regular mesh, halve until every tile
  // fits the executor's shared-memory budget (cells, entries, footprint nodes, list slots)
  int rtarget = dim == 3 ? 216 : 640;
  std::vector<TileDesc> hdesc;
  std::vector<int32_t> hstats;
  int32_t nb_tile = 0;
  int attempts = 0;
  for (;; rtarget /= 2, ++attempts) {
    AFB_REQUIRE(rtarget >= 1, AFB_ERR_UNSUPPORTED,
                "tiled gather: a single row exceeds the tile limits (%d cells / %d entries / %d footprint nodes / %d list slots); use AFB_VARIANT_NODEWISE", TG_CMAX,
                TG_EMAX, TG_FMAX, TG_LMAX);
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
    P.nb_tile = nb_tile;
    P.nb_tile_cell = P.nb_unit = P.nb_list = P.nb_foot = 0;
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
      if (hstats[4 * t] > TG_CMAX || hstats[4 * t + 1] > TG_EMAX || hdesc[t].nb_row > TG_RALLOC || hdesc[t].nb_row + hstats[4 * t + 3] > TG_FMAX) ok = false;
    if (!ok) continue;

    // sizes -> offsets (host; a few thousand tiles)
    int64_t cell_off = 0, unit_off = 0, list_off = 0, foot_off = 0;
    for (int32_t t = 0; t < nb_tile; ++t) {
      TileDesc& d = hdesc[t];
      const int C = hstats[4 * t], E = hstats[4 * t + 1], V = hstats[4 * t + 2], H = hstats[4 * t + 3];
      d.cell_off = (int32_t)cell_off;
      d.nb_cell = C;
      d.unit_off = (int32_t)unit_off;
      d.nb_unit = (E + 31) / 32; // upper bound; the builder stores the number of units of computed entries
      d.nb_entry = E;
      d.list_off = (uint32_t)list_off;
      d.list_len = 0;
      d.foot_off = (int32_t)foot_off;
      d.nb_foot = d.nb_row + H;
      d.pad = 0;
      cell_off += C;
      unit_off += d.nb_unit;
      foot_off += d.nb_foot;
      // list slots: sum over units of 32*len <= contributions + 32*(max count) + 32 per unit (even rounding),
      // never more than the executor can stage
      int64_t cap = (int64_t)npc * npc * C + 32ll * (V + 2) + 32ll * d.nb_unit + 8;
      cap = std::min<int64_t>((cap + 7) & ~7ll, TG_LMAX);
      list_off += cap;
      AFB_REQUIRE(list_off < (1ll << 32) && cell_off < (1ll << 31) && foot_off < (1ll << 31), AFB_ERR_OVERFLOW, "tiled gather plan exceeds 32-bit offsets");
    }
    P.nb_tile_cell = cell_off;
    P.nb_unit = unit_off;
    P.nb_list = list_off;
    P.nb_foot = foot_off;
    AFB_TRY(P.tile_cells.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(cell_off, 1)));
    AFB_TRY(P.lconn.reserve(sizeof(ushort4) * (size_t)std::max<int64_t>(cell_off, 1)));
    AFB_TRY(P.foot.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(foot_off, 1)));
    AFB_TRY(P.unit_base.reserve(sizeof(uint32_t) * (size_t)std::max<int64_t>(unit_off, 1)));
    AFB_TRY(P.unit_len.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(unit_off, 1)));
    AFB_TRY(P.gpos.reserve(sizeof(uint32_t) * 32 * (size_t)std::max<int64_t>(unit_off, 1)));
    AFB_TRY(P.gpos2.reserve(sizeof(uint32_t) * 32 * (size_t)std::max<int64_t>(unit_off, 1)));
    AFB_TRY(P.lists.reserve(sizeof(uint16_t) * (size_t)std::max<int64_t>(list_off, 8)));
    AFB_CUDA(cudaMemcpyAsync(P.tile_desc.p, hdesc.data(), sizeof(TileDesc) * (size_t)nb_tile, cudaMemcpyHostToDevice, st));
    AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
    AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
    const size_t smem = sizeof(BuilderSmem);
    const int grid = std::min<int>(nb_tile, ctx->sm_count);
    if (npc == 4) {
      AFB_CUDA(cudaFuncSetAttribute(k_tile_build<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tile_build<4><<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                                      ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(),
                                                      P.node_lrow.as<int32_t>(), P.tile_cells.as<int32_t>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(),
                                                      P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(), P.gpos2.as<uint32_t>(),
                                                      P.lists.as<uint16_t>(), list_off, ctx->tmp_flag.as<int>());
    }
    else {
      AFB_CUDA(cudaFuncSetAttribute(k_tile_build<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_tile_build<3><<<grid, TB_THREADS, smem, st>>>(P.tile_desc.as<TileDesc>(), nb_tile, P.tile_nodes.as<int32_t>(), ctx->conn.as<int32_t>(), ctx->nc_ptr.as<int32_t>(),
                                                      ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(), P.node_tile.as<int32_t>(),
                                                      P.node_lrow.as<int32_t>(), P.tile_cells.as<int32_t>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(),
                                                      P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(), P.gpos2.as<uint32_t>(),
                                                      P.lists.as<uint16_t>(), list_off, ctx->tmp_flag.as<int>());
    }
    AFB_LAUNCH_CHECK(ctx);
    int err = 0;
    AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    if (err == 3) continue; // a tile's lists do not fit the staging buffer: smaller tiles
    AFB_REQUIRE(err == 0, AFB_ERR_CUDA, "tiled gather: plan builder inconsistency (code %d)", err);
    break;
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
  const size_t smem = sizeof(ExecSmem);
  const int grid = std::min<int>(P.nb_tile, ctx->sm_count);
  // values already holding contributions (a second operator added on top) are accumulated into;
  // a freshly reset matrix is simply overwritten
  const int accumulate = ctx->assembled ? 1 : 0;
  if (ctx->npc == 4) {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled<4><<<grid, TG_THREADS, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.nb_tile, ctx->coords.as<double>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(),
                                                                 P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(), P.gpos2.as<uint32_t>(),
                                                                 P.lists.as<uint16_t>(), ctx->values.as<double>(), accumulate, prm);
  }
  else {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled<3><<<grid, TG_THREADS, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.nb_tile, ctx->coords.as<double>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(),
                                                                 P.unit_base.as<uint32_t>(), P.unit_len.as<uint16_t>(), P.gpos.as<uint32_t>(), P.gpos2.as<uint32_t>(),
                                                                 P.lists.as<uint16_t>(), ctx->values.as<double>(), accumulate, prm);
  }
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

} // namespace afb
