// Inspector of the chained-slice executor (chain.cuh, chain_exec.cu).  Runs once per mesh and
// ownership mode, at the first scalar tiled assembly.
//
//   columns   nodes are binned across the two cross axes (one in 2-D), and ordered along the sweep
//             axis inside each column (64-bit radix sort of (column, coordinate) keys)
//   slices    every column is cut greedily into slices of about one layer of nodes: a cut is placed
//             at the first large coordinate gap once the slice holds 3/4 of the target rows (layered
//             meshes: exactly one layer), else at the target; columns whose slices exceed an
//             executor limit are re-cut with a smaller cap
//   cells     per slice the cells touching its (owned) rows.  A cell that also touched the previous
//             slice of the same segment was computed there and is inherited (cache slot of the
//             previous slice, other region); the rest is computed by this slice ("new")
//   lists     per computed matrix entry the cache indices of its contributions, 4 per list row,
//             entries sorted by list rows and cut in units of 32; entries towards the next slice of
//             the segment are computed here and mirrored into the next slice's staging buffer
//   schedule  segments (<= CHAIN_SEG_MAX slices of a column) are dealt to the executor's CTAs by
//             decreasing cost (longest processing time first)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "chain.cuh"
#include "element.cuh"
#include "tiles.cuh" // off_pair, pack_rowinfo

namespace afb {

constexpr int CB_THREADS = 256;
constexpr int CB_CELLS = 2048;   // cells touching a slice the builders can hold
constexpr int CB_HASH = 2048;    // footprint hash slots
constexpr unsigned CB_EMPTY = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// columns
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ch_order_f64(double x)
{
  unsigned long long u = (unsigned long long)__double_as_longlong(x);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
static double ch_unorder_f64(unsigned long long u)
{
  u = (u >> 63) ? (u & 0x7FFFFFFFFFFFFFFFull) : ~u;
  double x;
  memcpy(&x, &u, sizeof(x));
  return x;
}

__global__ void __launch_bounds__(256) k_chain_bbox(const double* __restrict__ coords, int32_t nb_node, unsigned long long* __restrict__ box)
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
      atomicMin(box + a, ch_order_f64(mn[a]));
      atomicMax(box + 3 + a, ch_order_f64(mx[a]));
    }
  }
}

struct ColumnGrid {
  double x0[3], inv_h[2], inv_s; // cross-axis bins, sweep normalisation
  int ax[2], s;                  // cross axes (ax[1] = -1 in 2-D), sweep axis
  int g[2];
};

constexpr int CH_ZBITS = 40;

__global__ void __launch_bounds__(256) k_chain_keys(const double* __restrict__ coords, int32_t nb_node, ColumnGrid cg, unsigned long long* __restrict__ keys, int32_t* __restrict__ ids,
                                                     int32_t* __restrict__ col_count)
{
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb_node) return;
  int c[2] = { 0, 0 };
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (cg.ax[k] < 0) continue;
    const int v = (int)((coords[3 * (int64_t)i + cg.ax[k]] - cg.x0[cg.ax[k]]) * cg.inv_h[k]);
    c[k] = min(max(v, 0), cg.g[k] - 1);
  }
  const int32_t col = c[0] + cg.g[0] * c[1];
  double u = (coords[3 * (int64_t)i + cg.s] - cg.x0[cg.s]) * cg.inv_s;
  u = fmin(fmax(u, 0.0), 1.0);
  const unsigned long long z = (unsigned long long)(u * (double)((1ull << CH_ZBITS) - 1));
  keys[i] = ((unsigned long long)col << CH_ZBITS) | z;
  ids[i] = i;
  atomicAdd(col_count + col, 1);
}

// one warp per column: greedy cuts (see the file header); first[p] = 1 where a slice starts
__global__ void __launch_bounds__(256) k_chain_cut(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ col_ptr, int32_t nb_col, const int32_t* __restrict__ rcap,
                                                    int rtarget, int32_t* __restrict__ first)
{
  const int lane = threadIdx.x & 31;
  const int32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nb_col) return;
  const int32_t beg = col_ptr[c], end = col_ptr[c + 1];
  const int n = end - beg;
  if (n <= 0) return;
  const unsigned long long zmask = (1ull << CH_ZBITS) - 1;
  const unsigned long long span = (keys[end - 1] & zmask) - (keys[beg] & zmask);
  const int rmax = max(1, rcap[c]);
  const int rt = min(rtarget, rmax);
  const int rmin = max(1, (3 * rt) / 4);
  // a quarter of the expected thickness of a slice of rt rows (consecutive nodes of one layer are much closer, layers further apart)
  const unsigned long long thr = max(1ull, (unsigned long long)((double)span * (double)rt / (double)n * 0.25));
  int32_t pos = beg;
  if (lane == 0) first[beg] = 1;
  while (true) {
    const int32_t w0 = pos + rmin, w1 = min(pos + rmax, end - 1); // candidate cut positions [w0, w1]
    int32_t cut = -1;
    for (int32_t q0 = w0; q0 <= w1 && cut < 0; q0 += 32) {
      const int32_t q = q0 + lane;
      bool big = false;
      if (q <= w1) big = ((keys[q] & zmask) - (keys[q - 1] & zmask)) >= thr;
      const unsigned m = __ballot_sync(0xffffffffu, big);
      if (m) cut = q0 + __ffs(m) - 1;
    }
    if (cut < 0) {
      if (end - pos <= rmax) break; // the rest is the last slice
      cut = pos + rt;
    }
    if (cut >= end) break;
    if (lane == 0) first[cut] = 1;
    pos = cut;
  }
}

__global__ void __launch_bounds__(256) k_chain_assign(const unsigned long long* __restrict__ keys, const int32_t* __restrict__ ids, const int32_t* __restrict__ first,
                                                       const int32_t* __restrict__ excl, int32_t nb_node, int32_t* __restrict__ node_slice, int32_t* __restrict__ slice_start,
                                                       int32_t* __restrict__ slice_col)
{
  const int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nb_node) return;
  const int32_t t = excl[p] + first[p] - 1;
  node_slice[ids[p]] = t;
  if (first[p]) {
    slice_start[t] = p;
    slice_col[t] = (int32_t)(keys[p] >> CH_ZBITS);
  }
}

// per slice: rows in ascending node id, row index of each node, first entry of each row, flags
__global__ void __launch_bounds__(128) k_chain_rows(const int32_t* __restrict__ ids, const int32_t* __restrict__ slice_start, const int32_t* __restrict__ slice_col,
                                                     const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ excl, int32_t nb_slice, int32_t nb_node,
                                                     const int32_t* __restrict__ rows, int seg_len, int rmax, int32_t* __restrict__ slice_nodes, int32_t* __restrict__ node_lrow,
                                                     int32_t* __restrict__ node_e0, SliceDesc* __restrict__ desc, int32_t* __restrict__ seg_idx, int* __restrict__ error)
{
  __shared__ int32_t s_id[1024];
  __shared__ int32_t s_sorted[1024];
  __shared__ int s_deg[1024];
  const int32_t t = blockIdx.x;
  if (t >= nb_slice) return;
  const int32_t beg = slice_start[t], end = (t + 1 < nb_slice) ? slice_start[t + 1] : nb_node;
  const int n = end - beg;
  const int32_t col = slice_col[t];
  const int32_t col_first = excl[col_ptr[col]];            // slice of the column's first node
  const int32_t col_end_pos = col_ptr[col + 1];
  const int idx = t - col_first;
  const int sidx = idx % seg_len;
  if (threadIdx.x == 0) {
    SliceDesc d;
    memset(&d, 0, sizeof(d));
    d.node_off = beg;
    d.nb_row = n;
    d.flags = (sidx & 1 ? CH_FLAG_PARITY : 0) | (sidx == 0 ? CH_FLAG_FIRST : 0) | ((sidx == seg_len - 1 || end >= col_end_pos) ? CH_FLAG_LAST : 0) | (sidx << CH_FLAG_SIDX_SHIFT);
    desc[t] = d;
    seg_idx[t] = sidx;
  }
  if (n > 1024 || n > rmax) { // over the limit: reported through the statistics (nb_row), refined by the host
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      slice_nodes[beg + i] = ids[beg + i];
      node_lrow[ids[beg + i]] = i;
      node_e0[ids[beg + i]] = 0;
    }
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_id[i] = ids[beg + i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t x = s_id[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += s_id[j] < x ? 1 : 0;
    s_sorted[rank] = x;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int32_t x = s_sorted[i];
    slice_nodes[beg + i] = x;
    node_lrow[x] = i;
    s_deg[i] = rows[x + 1] - rows[x];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < n; ++i) {
      const int dg = s_deg[i];
      s_deg[i] = run;
      run += dg;
    }
    desc[t].nb_entry = run;
    if (run > 32767) atomicExch(error, 10);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) node_e0[s_sorted[i]] = s_deg[i];
}

// ---------------------------------------------------------------------------------------------
// cells of a slice
// ---------------------------------------------------------------------------------------------
struct CellCtx {
  const int32_t* conn;
  const int32_t* node_slice;
  const int32_t* node_lrow;
  const int32_t* seg_idx;
  const uint8_t* own;     // nullable: all nodes are rows
  int64_t nb_own_cell;    // cells >= this contribute nothing
};

template <int NPC>
__device__ __forceinline__ bool ch_touches(const CellCtx& C, const int32_t (&nd)[NPC], int32_t t)
{
  bool hit = false;
#pragma unroll
  for (int a = 0; a < NPC; ++a) hit |= (__ldg(C.node_slice + nd[a]) == t) && (!C.own || C.own[nd[a]]);
  return hit;
}

// class of the incidence (row i of slice t, cell): 0 = not the leader (another owned row of the slice with a smaller
// index holds the cell) or a ghost cell; 1 = inherited from the previous slice; 2 = computed here, needed by the
// next slice too (group A); 3 = computed here, dead after this slice (group B; only a segment's first slice
// distinguishes: its two groups go to the two cache regions)
template <int NPC>
__device__ __forceinline__ int ch_cell_class(const CellCtx& C, int32_t t, int sidx, bool last, int i, int32_t r, int32_t cell)
{
  if ((int64_t)cell >= C.nb_own_cell) return 0;
  int32_t nd[NPC];
#pragma unroll
  for (int a = 0; a < NPC; ++a) nd[a] = __ldg(C.conn + (int64_t)cell * NPC + a);
#pragma unroll
  for (int a = 0; a < NPC; ++a)
    if (nd[a] != r && __ldg(C.node_slice + nd[a]) == t && (!C.own || C.own[nd[a]]) && __ldg(C.node_lrow + nd[a]) < i) return 0;
  int run = 0;
  while (run < sidx && ch_touches<NPC>(C, nd, t - 1 - run)) ++run;
  if (run & 1) return 1;
  if (sidx != 0) return 2;
  return (!last && ch_touches<NPC>(C, nd, t + 1)) ? 2 : 3;
}

// per slice: cells (all, computed, group A), footprint of the computed cells, largest valence
template <int NPC>
__global__ void __launch_bounds__(128) k_chain_stats(const SliceDesc* __restrict__ desc, int32_t nb_slice, const int32_t* __restrict__ slice_nodes, CellCtx C,
                                                      const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int fmax, int32_t* __restrict__ stats /* [nb_slice][5] */)
{
  __shared__ int s_all, s_new, s_a, s_v, s_f;
  __shared__ unsigned s_tab[CB_HASH];
  const int32_t t = blockIdx.x;
  if (t >= nb_slice) return;
  if (threadIdx.x == 0) s_all = s_new = s_a = s_v = s_f = 0;
  for (int i = threadIdx.x; i < CB_HASH; i += blockDim.x) s_tab[i] = CB_EMPTY;
  __syncthreads();
  const SliceDesc d = desc[t];
  const int sidx = C.seg_idx[t];
  const bool last = (d.flags & CH_FLAG_LAST) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int call = 0, cnew = 0, ca = 0, v = 0;
  for (int i = warp; i < d.nb_row; i += 4) {
    const int32_t r = slice_nodes[d.node_off + i];
    if (C.own && !C.own[r]) continue;
    const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
    if (lane == 0) v = max(v, qe - qb);
    for (int q = qb + lane; q < qe; q += 32) {
      const int32_t cell = nc_list[q];
      const int cls = ch_cell_class<NPC>(C, t, sidx, last, i, r, cell);
      if (cls == 0) continue;
      ++call;
      if (cls >= 2) {
        ++cnew;
        if (cls == 2) ++ca;
        if (s_f <= fmax) {
#pragma unroll
          for (int a = 0; a < NPC; ++a) {
            const unsigned n = (unsigned)__ldg(C.conn + (int64_t)cell * NPC + a);
            unsigned h = (n * 0x9E3779B1u) >> 21;
            while (true) {
              const unsigned old = atomicCAS(s_tab + h, CB_EMPTY, n);
              if (old == CB_EMPTY) { atomicAdd(&s_f, 1); break; }
              if (old == n) break;
              h = (h + 1) & (CB_HASH - 1);
              if (s_f > fmax) break;
            }
          }
        }
      }
    }
  }
  atomicAdd(&s_all, call);
  atomicAdd(&s_new, cnew);
  atomicAdd(&s_a, ca);
  atomicMax(&s_v, v);
  __syncthreads();
  if (threadIdx.x == 0) {
    stats[5 * t + 0] = s_all;
    stats[5 * t + 1] = s_new;
    stats[5 * t + 2] = s_a;
    stats[5 * t + 3] = s_f;
    stats[5 * t + 4] = s_v;
  }
}

// ascending bitonic sort of n2 (power of two) 32-bit keys in shared memory
__device__ void ch_bitonic_sort(unsigned* s, int n2)
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

__device__ __forceinline__ int ch_find_u32(const unsigned* __restrict__ s, int n, unsigned id)
{
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (s[mid] <= id) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int ch_find_i32_global(const int32_t* __restrict__ a, int n, int32_t id) // index of id in ascending a[0,n), -1 if absent
{
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= id) lo = mid; else hi = mid;
  }
  return (n > 0 && __ldg(a + lo) == id) ? lo : -1;
}

// in-place exclusive scan of n ints in shared memory; returns the total
__device__ int ch_exclusive_scan(int* s, int n, int* s_tmp /* >= 33 ints */)
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

// per slice: computed cells in slot order (group A ascending, then group B ascending), their footprint
// (ascending node ids) and footprint-local connectivity
template <int NPC>
__global__ void __launch_bounds__(CB_THREADS) k_chain_mesh(const SliceDesc* __restrict__ desc, int32_t nb_slice, const int32_t* __restrict__ slice_nodes, CellCtx C,
                                                            const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int32_t* __restrict__ new_cells,
                                                            int32_t* __restrict__ foot, ushort4* __restrict__ lconn, int* __restrict__ error)
{
  __shared__ unsigned s_cells[CB_CELLS];
  __shared__ unsigned s_tab[CB_HASH];
  __shared__ unsigned s_foot[1024];
  __shared__ int s_nc, s_nf;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int32_t t = blockIdx.x; t < nb_slice; t += gridDim.x) {
    const SliceDesc d = desc[t];
    const int sidx = C.seg_idx[t];
    const bool last = (d.flags & CH_FLAG_LAST) != 0;
    if (threadIdx.x == 0) s_nc = s_nf = 0;
    for (int i = threadIdx.x; i < CB_CELLS; i += blockDim.x) s_cells[i] = CB_EMPTY;
    for (int i = threadIdx.x; i < CB_HASH; i += blockDim.x) s_tab[i] = CB_EMPTY;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_foot[i] = CB_EMPTY;
    __syncthreads();
    for (int i = warp; i < d.nb_row; i += nwarp) {
      const int32_t r = slice_nodes[d.node_off + i];
      if (C.own && !C.own[r]) continue;
      const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
      for (int q = qb + lane; q < qe; q += 32) {
        const int32_t cell = nc_list[q];
        const int cls = ch_cell_class<NPC>(C, t, sidx, last, i, r, cell);
        if (cls >= 2) {
          const int pos = atomicAdd(&s_nc, 1);
          if (pos < CB_CELLS) s_cells[pos] = (cls == 3 ? 0x80000000u : 0u) | (unsigned)cell; // group B after group A
        }
      }
    }
    __syncthreads();
    const int CN = s_nc;
    if (CN != d.nb_new || CN > CB_CELLS) {
      if (threadIdx.x == 0) atomicExch(error, 21);
      __syncthreads();
      continue;
    }
    int c2 = 32;
    while (c2 < CN) c2 <<= 1;
    ch_bitonic_sort(s_cells, c2);
    for (int lc = threadIdx.x; lc < CN; lc += blockDim.x) {
      const int32_t cell = (int32_t)(s_cells[lc] & 0x7FFFFFFFu);
      new_cells[d.cell_off + lc] = cell;
#pragma unroll
      for (int a = 0; a < NPC; ++a) {
        const unsigned n = (unsigned)__ldg(C.conn + (int64_t)cell * NPC + a);
        unsigned h = (n * 0x9E3779B1u) >> 21;
        while (true) {
          const unsigned old = atomicCAS(s_tab + h, CB_EMPTY, n);
          if (old == CB_EMPTY) {
            const int pos = atomicAdd(&s_nf, 1);
            if (pos < 1024) s_foot[pos] = n;
            break;
          }
          if (old == n) break;
          h = (h + 1) & (CB_HASH - 1);
        }
      }
    }
    __syncthreads();
    const int F = s_nf;
    if (F != d.nb_foot || F > 1024) {
      if (threadIdx.x == 0) atomicExch(error, 22);
      __syncthreads();
      continue;
    }
    int f2 = 32;
    while (f2 < F) f2 <<= 1;
    ch_bitonic_sort(s_foot, f2);
    for (int f = threadIdx.x; f < F; f += blockDim.x) foot[d.foot_off + f] = (int32_t)s_foot[f];
    for (int lc = threadIdx.x; lc < CN; lc += blockDim.x) {
      const int32_t cell = (int32_t)(s_cells[lc] & 0x7FFFFFFFu);
      unsigned short loc[4] = { 0, 0, 0, 0 };
#pragma unroll
      for (int a = 0; a < NPC; ++a) loc[a] = (unsigned short)ch_find_u32(s_foot, F, (unsigned)__ldg(C.conn + (int64_t)cell * NPC + a));
      lconn[d.cell_off + lc] = make_ushort4(loc[0], loc[1], loc[2], loc[3]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// lists
// ---------------------------------------------------------------------------------------------
constexpr int CLS_SKIP = -1;   // mirror of an entry computed in this slice, entry deposited by the previous slice, derived diagonal, zero row
constexpr int CL_KEYS = 2048;  // power of two >= EMAX of every geometry

struct ChainListSmem {
  int erow[1024 + 1];
  int cnt[CL_KEYS];
  int eoff[CL_KEYS + 1];
  unsigned keys[CL_KEYS];
  uint16_t e2[CL_KEYS];
  uint16_t clist[6 * CB_CELLS];
  unsigned cells[CB_CELLS];     // cell id
  uint16_t cslot[CB_CELLS];     // cache position (region * CS + pos)
  int ubase[CL_KEYS / 32 + 2];
  int unch[CL_KEYS / 32 + 2];
  int tmp[40];
  int ncell;
};

template <int NPC>
__global__ void __launch_bounds__(CB_THREADS, 1)
k_chain_lists(SliceDesc* __restrict__ desc, int32_t nb_slice, const int32_t* __restrict__ slice_nodes, CellCtx C, const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list,
              const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ node_e0, const int32_t* __restrict__ new_cells, int cs /* region stride */,
              int nreg, int emax, int prefill, unsigned char* __restrict__ blob, int* __restrict__ error)
{
  extern __shared__ unsigned char cl_raw[];
  ChainListSmem& S = *reinterpret_cast<ChainListSmem*>(cl_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int plane_stride = nreg * cs;
  constexpr int NPAIR = NPC * (NPC - 1) / 2;
  const unsigned ZERO = (unsigned)(NPAIR * plane_stride);
  for (int32_t t = blockIdx.x; t < nb_slice; t += gridDim.x) {
    const SliceDesc d = desc[t];
    const int R = d.nb_row, E = d.nb_entry;
    const int sidx = C.seg_idx[t];
    const bool last = (d.flags & CH_FLAG_LAST) != 0;
    const int reg_new = ch_reg_new(nreg, sidx), reg_prev = ch_reg_prev(nreg, sidx), reg_b = ch_reg_b(nreg);
    if (threadIdx.x == 0) S.ncell = 0;
    for (int i = threadIdx.x; i <= R; i += blockDim.x) S.erow[i] = i < R ? node_e0[slice_nodes[d.node_off + i]] : E;
    __syncthreads();
    if (E > emax || E > CL_KEYS || R > 1024) {
      if (threadIdx.x == 0) atomicExch(error, 31);
      __syncthreads();
      continue;
    }
    unsigned char* rec = blob + (size_t)d.blob_off * 16;
    // ---- entries: class, mirror; rowinfo ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = slice_nodes[d.node_off + i];
      const bool own_i = !C.own || C.own[r];
      const int rb = rows[r], deg = rows[r + 1] - rb, e0 = S.erow[i];
      int pdiag = 0;
      for (int p0 = 0; p0 < deg; p0 += 32) {
        const int p = p0 + lane;
        int32_t c = -1;
        if (p < deg) {
          c = cols[rb + p];
          int cls = 0;
          unsigned m = CH_NONE16;
          if (!own_i || c == r) cls = CLS_SKIP;
          else if (!C.own || C.own[c]) {
            const int32_t tc = __ldg(C.node_slice + c);
            if (tc == t) {
              const int j = __ldg(C.node_lrow + c);
              if (j < i) cls = CLS_SKIP;
              else {
                const int cb = rows[c], ce = rows[c + 1];
                m = (unsigned)(S.erow[j] + (find_col(cols, cb, ce, r) - cb));
              }
            }
            else if (prefill && tc == t + 1 && !last) {
              const int cb = rows[c], ce = rows[c + 1];
              m = CH_NEXT | (unsigned)(__ldg(node_e0 + c) + (find_col(cols, cb, ce, r) - cb));
            }
            else if (prefill && tc == t - 1 && sidx > 0) cls = CLS_SKIP;
          }
          S.cnt[e0 + p] = cls;
          S.e2[e0 + p] = (uint16_t)m;
        }
        const unsigned hit = __ballot_sync(0xffffffffu, c == r);
        if (hit) pdiag = p0 + __ffs(hit) - 1;
      }
      if (lane == 0) S.keys[i] = pack_rowinfo(e0, pdiag, own_i); // rowinfo parked in keys[] until the record layout is known
    }
    __syncthreads();
    // rowinfo -> registers of the first R threads (keys[] is reused by the sort)
    uint32_t my_rowinfo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * CB_THREADS;
      my_rowinfo[q] = i < R ? S.keys[i] : 0u;
    }
    __syncthreads();
    // ---- cells touching the slice, with their cache positions ----
    for (int i = warp; i < R; i += nwarp) {
      const int32_t r = slice_nodes[d.node_off + i];
      if (C.own && !C.own[r]) continue;
      const int qb = nc_ptr[r], qe = nc_ptr[r + 1];
      for (int q = qb + lane; q < qe; q += 32) {
        const int32_t cell = nc_list[q];
        const int cls = ch_cell_class<NPC>(C, t, sidx, last, i, r, cell);
        if (cls == 0) continue;
        int slot = -1;
        if (cls == 1) { // inherited: position in the previous slice's group A, other region
          const SliceDesc dp = desc[t - 1];
          const int k = ch_find_i32_global(new_cells + dp.cell_off, dp.nb_a, cell);
          if (k >= 0) slot = reg_prev * cs + k;
        }
        else {
          if (cls == 2) {
            const int k = ch_find_i32_global(new_cells + d.cell_off, d.nb_a, cell);
            if (k >= 0) slot = reg_new * cs + k;
          }
          else {
            const int k = ch_find_i32_global(new_cells + d.cell_off + d.nb_a, d.nb_new - d.nb_a, cell);
            if (k >= 0) slot = reg_b * cs + k;
          }
        }
        if (slot < 0) {
          atomicExch(error, 32);
          continue;
        }
        const int pos = atomicAdd(&S.ncell, 1);
        if (pos < CB_CELLS) {
          S.cells[pos] = (unsigned)cell;
          S.cslot[pos] = (uint16_t)slot;
        }
      }
    }
    __syncthreads();
    const int NC = min(S.ncell, CB_CELLS);
    if (S.ncell != d.nb_cell) {
      if (threadIdx.x == 0) atomicExch(error, 33);
      __syncthreads();
      continue;
    }
    // ---- contribution lists of the computed entries (count, scan, fill) ----
    for (int pass = 0; pass < 2; ++pass) {
      for (int lc = threadIdx.x; lc < NC; lc += blockDim.x) {
        const int32_t cell = (int32_t)S.cells[lc];
        const int slot = S.cslot[lc];
        int32_t nd[NPC];
        int li[NPC];
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          nd[a] = __ldg(C.conn + (int64_t)cell * NPC + a);
          li[a] = (__ldg(C.node_slice + nd[a]) == t && (!C.own || C.own[nd[a]])) ? __ldg(C.node_lrow + nd[a]) : -1;
        }
#pragma unroll
        for (int a = 0; a < NPC; ++a) {
          if (li[a] < 0) continue;
          const int rb = __ldg(rows + nd[a]), re = __ldg(rows + nd[a] + 1);
#pragma unroll
          for (int bq = 0; bq < NPC; ++bq) {
            if (bq == a) continue;
            const int e = S.erow[li[a]] + (find_col(cols, rb, re, nd[bq]) - rb);
            if (S.cnt[e] < 0) continue; // skipped class (a computed entry's counter never drops below zero)
            if (pass == 0) atomicAdd(&S.cnt[e], 1);
            else {
              const int k = atomicSub(&S.cnt[e], 1) - 1;
              S.clist[S.eoff[e] + k] = (uint16_t)(off_pair(NPC, a, bq) * plane_stride + slot);
            }
          }
        }
      }
      __syncthreads();
      if (pass == 0) {
        for (int e = threadIdx.x; e <= E; e += blockDim.x) S.eoff[e] = e < E ? max(S.cnt[e], 0) : 0;
        __syncthreads();
        const int total = ch_exclusive_scan(S.eoff, E + 1, S.tmp);
        if (total > 6 * CB_CELLS) {
          if (threadIdx.x == 0) atomicExch(error, 34);
        }
        // entries that are computed but received nothing (isolated pattern entries cannot occur: every off-diagonal
        // entry of the pattern comes from a cell) keep cnt = 0; skipped classes keep cnt < 0
      }
      else {
        for (int e = threadIdx.x; e < E; e += blockDim.x)
          if (S.cnt[e] >= 0) S.cnt[e] = S.eoff[e + 1] - S.eoff[e];
        __syncthreads();
      }
    }
    // ---- fixed summation order: ascending cache index ----
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const int n = S.cnt[e];
      if (n <= 1) continue;
      uint16_t* l = S.clist + S.eoff[e];
      for (int i = 1; i < n; ++i) {
        const uint16_t x = l[i];
        int j = i - 1;
        while (j >= 0 && l[j] > x) {
          l[j + 1] = l[j];
          --j;
        }
        l[j + 1] = x;
      }
    }
    // ---- computed entries by descending list rows, cut into units of 32 ----
    int k2 = 32;
    while (k2 < E) k2 <<= 1;
    for (int e = threadIdx.x; e < k2; e += blockDim.x) {
      unsigned key = 0xFFFFFFFFu;
      if (e < E && S.cnt[e] >= 0) {
        const int nch = (S.cnt[e] + 3) >> 2;
        key = ((unsigned)(0xFFFF - nch) << 16) | (unsigned)e;
      }
      S.keys[e] = key;
    }
    if (threadIdx.x == 0) S.tmp[34] = 0;
    __syncthreads();
    {
      int mine = 0;
      for (int e = threadIdx.x; e < E; e += blockDim.x) mine += S.cnt[e] >= 0 ? 1 : 0;
      atomicAdd(&S.tmp[34], mine);
    }
    ch_bitonic_sort(S.keys, k2);
    const int EC = S.tmp[34];
    const int nunit = (EC + 31) / 32;
    for (int u = threadIdx.x; u <= nunit; u += blockDim.x) {
      int nch = 0;
      if (u < nunit) nch = max(1, (S.cnt[S.keys[u * 32] & 0xFFFFu] + 3) >> 2);
      S.unch[u] = nch;
      S.ubase[u] = nch;
    }
    __syncthreads();
    const int nchunk = ch_exclusive_scan(S.ubase, nunit + 1, S.tmp);
    const int bytes = ch_blob_bytes(nchunk, nunit, R, E);
    if (bytes > d.blob_cap) {
      if (threadIdx.x == 0) atomicExch(error, 35);
      __syncthreads();
      continue;
    }
    if (threadIdx.x == 0) {
      desc[t].nb_unit = nunit;
      desc[t].nb_chunk = nchunk;
      desc[t].blob_bytes = bytes;
    }
    uint2* out_l = reinterpret_cast<uint2*>(rec);
    uint32_t* out_em = reinterpret_cast<uint32_t*>(rec + ch_off_emap(nchunk));
    uint32_t* out_un = reinterpret_cast<uint32_t*>(rec + ch_off_units(nchunk, nunit));
    uint32_t* out_ri = reinterpret_cast<uint32_t*>(rec + ch_off_rowinfo(nchunk, nunit));
    for (int u = threadIdx.x; u < ((nunit + 3) & ~3); u += blockDim.x) out_un[u] = u < nunit ? (((uint32_t)S.ubase[u] << 8) | (uint32_t)S.unch[u]) : 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = threadIdx.x + q * CB_THREADS;
      if (i < R) out_ri[i] = my_rowinfo[q];
    }
    for (int i = R + threadIdx.x; i < ((R + 1 + 3) & ~3); i += blockDim.x) out_ri[i] = pack_rowinfo(E, 0, false);
    {
      uint8_t* out_er = rec + ch_off_erow(nchunk, nunit, R);
      for (int i = warp; i < R; i += nwarp)
        for (int e = S.erow[i] + lane; e < S.erow[i + 1]; e += 32) out_er[e] = (uint8_t)i;
      for (int e = E + threadIdx.x; e < ch_align16(E); e += blockDim.x) out_er[e] = 0;
    }
    for (int x = threadIdx.x; x < nunit * 32; x += blockDim.x) {
      const bool valid = x < EC;
      const int e = valid ? (int)(S.keys[x] & 0xFFFFu) : 0;
      out_em[x] = valid ? ((uint32_t)e | ((uint32_t)S.e2[e] << 16)) : 0xFFFFFFFFu;
      const int u = x >> 5, ln = x & 31;
      const int nch = S.unch[u], n = valid ? S.cnt[e] : 0;
      const uint16_t* src = S.clist + (valid ? S.eoff[e] : 0);
      for (int k = 0; k < nch; ++k) {
        unsigned idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (4 * k + j < n) ? (unsigned)src[4 * k + j] : ZERO;
        out_l[(size_t)(S.ubase[u] + k) * 32 + ln] = make_uint2(idx[0] | (idx[1] << 16), idx[2] | (idx[3] << 16));
      }
    }
    __syncthreads();
  }
}

// index blocks in execution order (what the pipelined executor's loaders stream through the TMA engine)
__global__ void __launch_bounds__(128) k_chain_iblock(const SliceDesc* __restrict__ desc_exec, const int32_t* __restrict__ ib_off, int32_t nb_slice, const int32_t* __restrict__ foot,
                                                       const int32_t* __restrict__ slice_nodes, unsigned char* __restrict__ iblock)
{
  const int32_t e = blockIdx.x;
  if (e >= nb_slice) return;
  const SliceDesc d = desc_exec[e];
  unsigned char* out = iblock + (size_t)ib_off[e] * 16;
  if (threadIdx.x < 16) reinterpret_cast<int32_t*>(out)[threadIdx.x] = reinterpret_cast<const int32_t*>(&d)[threadIdx.x];
  int32_t* of = reinterpret_cast<int32_t*>(out + ch_ib_foot());
  for (int i = threadIdx.x; i < ((d.nb_foot + 3) & ~3); i += blockDim.x) of[i] = i < d.nb_foot ? foot[d.foot_off + i] : -1;
  int32_t* on = reinterpret_cast<int32_t*>(out + ch_ib_nodes(d.nb_foot));
  for (int i = threadIdx.x; i < ((d.nb_row + 3) & ~3); i += blockDim.x) on[i] = i < d.nb_row ? slice_nodes[d.node_off + i] : -1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static ChainPlan* chain_of(afb_ctx* ctx)
{
  if (!ctx->chain) ctx->chain = new ChainPlan();
  return static_cast<ChainPlan*>(ctx->chain);
}

void chain_destroy(afb_ctx* ctx)
{
  if (!ctx->chain) return;
  ChainPlan* P = static_cast<ChainPlan*>(ctx->chain);
  DevBuf* bufs[] = { &P->desc, &P->desc_exec, &P->iblock, &P->ib_off, &P->slice_nodes, &P->node_slice, &P->node_lrow, &P->node_e0, &P->new_cells, &P->lconn, &P->foot, &P->blob, &P->order, &P->cta_ptr, &P->errflag,
                     &P->scratch_a, &P->scratch_b, &P->scratch_c, &P->scratch_d, &P->sort_tmp };
  for (DevBuf* b : bufs) b->release();
  delete P;
  ctx->chain = nullptr;
}

bool chain_plan_valid(const afb_ctx* ctx, int mode, int geom)
{
  const ChainPlan* P = static_cast<const ChainPlan*>(ctx->chain);
  return P && P->valid && P->mesh_gen == ctx->mesh_gen && P->mode == mode && P->geom == geom;
}

template <class F> static int ch_by_npc(int npc, F f)
{
  if (npc == 4) return f(std::integral_constant<int, 4>());
  return f(std::integral_constant<int, 3>());
}

int chain_build(afb_ctx* ctx, int mode_flags, int geom, const ChainLimits& L, int grid)
{
  AFB_REQUIRE(ctx->npc == ctx->dim + 1, AFB_ERR_UNSUPPORTED, "the tiled path is not available for %d-node cells in dimension %d (P1 simplices only); use AFB_VARIANT_NODEWISE", ctx->npc, ctx->dim);
  AFB_REQUIRE(ctx->has_pattern && ctx->b == 1, AFB_ERR_INVALID, "chain inspector: build a scalar pattern first");
  ChainPlan& P = *chain_of(ctx);
  P.valid = false;
  cudaStream_t st = ctx->stream;
  const int32_t nb_node = ctx->nb_node;
  const int dim = ctx->dim, npc = ctx->npc;
  cudaEvent_t e0, e1;
  AFB_CUDA(cudaEventCreate(&e0));
  AFB_CUDA(cudaEventCreate(&e1));
  AFB_CUDA(cudaEventRecord(e0, st));
  P.nb_slice = 0;
  if (nb_node == 0) {
    P.mesh_gen = ctx->mesh_gen;
    P.mode = mode_flags;
    P.geom = geom;
    P.grid = grid;
    P.valid = true;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return AFB_OK;
  }
  // ---- bounding box, sweep axis, column grid ----
  AFB_TRY(P.scratch_d.reserve(sizeof(unsigned long long) * 8));
  unsigned long long init[6] = { ~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull }, got[6];
  AFB_CUDA(cudaMemcpyAsync(P.scratch_d.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  k_chain_bbox<<<std::min(grid_for(nb_node, 256), 4 * ctx->sm_count), 256, 0, st>>>(ctx->coords.as<double>(), nb_node, P.scratch_d.as<unsigned long long>());
  AFB_LAUNCH_CHECK(ctx);
  AFB_CUDA(cudaMemcpyAsync(got, P.scratch_d.p, sizeof(got), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  double lo[3], ext[3];
  int nd_ext = 0;
  double vol = 1.0, ext_max = 0.0;
  for (int a = 0; a < 3; ++a) {
    lo[a] = ch_unorder_f64(got[a]);
    ext[a] = ch_unorder_f64(got[3 + a]) - lo[a];
    if (!(ext[a] > 0.0) || a >= dim) ext[a] = 0.0;
    if (ext[a] > 0.0) {
      vol *= ext[a];
      ++nd_ext;
      ext_max = std::max(ext_max, ext[a]);
    }
  }
  ColumnGrid cg;
  memset(&cg, 0, sizeof(cg));
  // sweep along the last axis with a decent extent (node ids of generated and most imported meshes vary slowest
  // along it: the rows of a slice are then runs of consecutive ids, i.e. contiguous in the value array)
  cg.s = 0;
  for (int a = 0; a < 3; ++a)
    if (ext[a] > 0.0 && ext[a] >= 0.25 * ext_max) cg.s = a;
  const int rtarget = nd_ext >= 3 ? L.rt3 : L.rt2;
  const double spacing = nd_ext ? pow(vol / (double)nb_node, 1.0 / nd_ext) : 1.0; // mean node spacing
  int nax = 0;
  cg.ax[0] = cg.ax[1] = -1;
  cg.g[0] = cg.g[1] = 1;
  for (int a = 0; a < 3; ++a)
    if (a != cg.s && ext[a] > 0.0 && nax < 2) cg.ax[nax++] = a;
  const double width = nax == 2 ? floor(sqrt((double)rtarget)) * spacing : (nax == 1 ? (double)rtarget * spacing : 1.0);
  int64_t nb_col64 = 1;
  for (int k = 0; k < nax; ++k) {
    const int a = cg.ax[k];
    // (nodes per unit length along a) * ext / nodes per column side, rounded up: a column never holds more than the target
    cg.g[k] = std::max(1, (int)ceil((ext[a] + spacing) / width - 1e-9));
    cg.inv_h[k] = (double)cg.g[k] / ext[a];
    nb_col64 *= cg.g[k];
  }
  for (int a = 0; a < 3; ++a) cg.x0[a] = lo[a];
  cg.inv_s = ext[cg.s] > 0.0 ? 1.0 / ext[cg.s] : 0.0;
  AFB_REQUIRE(nb_col64 < (1ll << 23), AFB_ERR_UNSUPPORTED, "chain inspector: column grid too large");
  const int32_t nb_col = (int32_t)nb_col64;
  // ---- keys, sort ----
  AFB_TRY(P.scratch_a.reserve(sizeof(unsigned long long) * 2 * (size_t)nb_node));        // keys in / out
  AFB_TRY(P.scratch_b.reserve(sizeof(int32_t) * (2 * (size_t)nb_node + 2 * ((size_t)nb_col + 2)))); // ids in / out, col_count, col_ptr
  unsigned long long* keys_in = P.scratch_a.as<unsigned long long>();
  unsigned long long* keys = keys_in + nb_node;
  int32_t* ids_in = P.scratch_b.as<int32_t>();
  int32_t* ids = ids_in + nb_node;
  int32_t* col_count = ids + nb_node;
  int32_t* col_ptr = col_count + (nb_col + 2);
  AFB_CUDA(cudaMemsetAsync(col_count, 0, sizeof(int32_t) * (size_t)(nb_col + 2), st));
  k_chain_keys<<<grid_for(nb_node, 256), 256, 0, st>>>(ctx->coords.as<double>(), nb_node, cg, keys_in, ids_in, col_count);
  AFB_LAUNCH_CHECK(ctx);
  int col_bits = 1;
  while ((1ll << col_bits) < nb_col64) ++col_bits;
  size_t tmp_bytes = 0;
  AFB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys, ids_in, ids, nb_node, 0, CH_ZBITS + col_bits, st));
  AFB_TRY(P.sort_tmp.reserve(tmp_bytes));
  AFB_CUDA(cub::DeviceRadixSort::SortPairs(P.sort_tmp.p, tmp_bytes, keys_in, keys, ids_in, ids, nb_node, 0, CH_ZBITS + col_bits, st));
  ctx->launches++;
  AFB_TRY(exclusive_scan_i32(ctx, col_count, col_ptr, nb_col));
  // ---- slices (greedy cuts, refined per column until every slice fits the executor) ----
  AFB_TRY(P.node_slice.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.node_lrow.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.node_e0.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.slice_nodes.reserve(sizeof(int32_t) * (size_t)nb_node));
  AFB_TRY(P.scratch_c.reserve(sizeof(int32_t) * (2 * ((size_t)nb_node + 2) + (size_t)nb_col + 2))); // first, excl, rcap
  int32_t* first = P.scratch_c.as<int32_t>();
  int32_t* excl = first + (nb_node + 2);
  int32_t* rcap = excl + (nb_node + 2);
  std::vector<int32_t> hrcap((size_t)nb_col, std::min(L.rmax, rtarget + rtarget / 4));
  const uint8_t* own = (ctx->all_own || (mode_flags & AFB_FLAG_ALL_ROWS)) ? nullptr : ctx->is_own.as<uint8_t>();
  const int64_t nb_own_cell = (mode_flags & AFB_FLAG_OWN_CELLS_ONLY) ? ctx->nb_own_cell : ctx->nb_cell;
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), st));
  int32_t nb_slice = 0;
  std::vector<SliceDesc> hdesc;
  std::vector<int32_t> hstats, hcol;
  int seg_len = CHAIN_SEG_MAX;
  for (int attempt = 0;; ++attempt) {
    AFB_REQUIRE(attempt < 40, AFB_ERR_UNSUPPORTED, "chain inspector: refinement did not converge");
    AFB_CUDA(cudaMemcpyAsync(rcap, hrcap.data(), sizeof(int32_t) * (size_t)nb_col, cudaMemcpyHostToDevice, st));
    AFB_CUDA(cudaMemsetAsync(first, 0, sizeof(int32_t) * ((size_t)nb_node + 2), st));
    k_chain_cut<<<grid_for(nb_col, 8), 256, 0, st>>>(keys, col_ptr, nb_col, rcap, rtarget, first);
    AFB_LAUNCH_CHECK(ctx);
    AFB_TRY(exclusive_scan_i32(ctx, first, excl, nb_node));
    AFB_CUDA(cudaMemcpyAsync(&nb_slice, excl + nb_node, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    AFB_REQUIRE(nb_slice > 0 && nb_slice < (1 << 30), AFB_ERR_OVERFLOW, "chain inspector: bad slice count %d", nb_slice);
    // segment length: at least ~4 segments per CTA when the mesh allows it
    seg_len = (int)std::min<int64_t>(CHAIN_SEG_MAX, std::max<int64_t>(4, (int64_t)nb_slice / (4 * (int64_t)std::max(grid, 1))));
    AFB_TRY(P.scratch_d.reserve(sizeof(int32_t) * (8 * (size_t)nb_slice + 8))); // slice_start, slice_col, seg_idx, stats[5]
    int32_t* slice_start = P.scratch_d.as<int32_t>();
    int32_t* slice_col = slice_start + nb_slice + 2;
    int32_t* seg_idx = slice_col + nb_slice + 2;
    int32_t* stats = seg_idx + nb_slice + 2;
    k_chain_assign<<<grid_for(nb_node, 256), 256, 0, st>>>(keys, ids, first, excl, nb_node, P.node_slice.as<int32_t>(), slice_start, slice_col);
    AFB_LAUNCH_CHECK(ctx);
    AFB_TRY(P.desc.reserve(sizeof(SliceDesc) * (size_t)nb_slice));
    k_chain_rows<<<nb_slice, 128, 0, st>>>(ids, slice_start, slice_col, col_ptr, excl, nb_slice, nb_node, ctx->rows.as<int32_t>(), seg_len, L.rmax, P.slice_nodes.as<int32_t>(),
                                           P.node_lrow.as<int32_t>(), P.node_e0.as<int32_t>(), P.desc.as<SliceDesc>(), seg_idx, ctx->tmp_flag.as<int>());
    AFB_LAUNCH_CHECK(ctx);
    CellCtx C{ ctx->conn.as<int32_t>(), P.node_slice.as<int32_t>(), P.node_lrow.as<int32_t>(), seg_idx, own, nb_own_cell };
    AFB_TRY(ch_by_npc(npc, [&](auto N) {
      k_chain_stats<decltype(N)::value><<<nb_slice, 128, 0, st>>>(P.desc.as<SliceDesc>(), nb_slice, P.slice_nodes.as<int32_t>(), C, ctx->nc_ptr.as<int32_t>(),
                                                                    ctx->nc_list.as<int32_t>(), L.fmax, stats);
      return AFB_OK;
    }));
    AFB_LAUNCH_CHECK(ctx);
    hdesc.resize(nb_slice);
    hstats.resize(5 * (size_t)nb_slice);
    hcol.resize(nb_slice);
    AFB_CUDA(cudaMemcpyAsync(hdesc.data(), P.desc.p, sizeof(SliceDesc) * (size_t)nb_slice, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaMemcpyAsync(hstats.data(), stats, sizeof(int32_t) * 5 * (size_t)nb_slice, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaMemcpyAsync(hcol.data(), slice_col, sizeof(int32_t) * (size_t)nb_slice, cudaMemcpyDeviceToHost, st));
    AFB_CUDA(cudaStreamSynchronize(st));
    bool ok = true;
    std::vector<uint8_t> bad((size_t)nb_col, 0);
    for (int32_t t = 0; t < nb_slice; ++t) {
      const int call = hstats[5 * t], cnew = hstats[5 * t + 1], ca = hstats[5 * t + 2], F = hstats[5 * t + 3];
      const bool firstslice = (hdesc[t].flags & CH_FLAG_FIRST) != 0;
      bool b = hdesc[t].nb_row > L.rmax || hdesc[t].nb_entry > L.emax || F > L.fmax || call > CB_CELLS;
      if (firstslice) b = b || ca > L.cn || (cnew - ca) > L.cn;
      else b = b || cnew > L.cn;
      if (b) {
        if (hdesc[t].nb_row <= 1) {
          set_error("tiled path: a single row exceeds the slice limits (%d cells / %d entries / %d footprint nodes); use AFB_VARIANT_NODEWISE", L.cn, L.emax, L.fmax);
          cudaEventDestroy(e0);
          cudaEventDestroy(e1);
          return AFB_ERR_UNSUPPORTED;
        }
        bad[hcol[t]] = 1;
        hrcap[hcol[t]] = std::min(hrcap[hcol[t]], std::max(1, (2 * hdesc[t].nb_row) / 3));
        ok = false;
      }
    }
    if (ok) break;
  }
  // ---- offsets ----
  int64_t cell_off = 0, foot_off = 0, blob_off = 0, all_cells = 0, all_new = 0;
  for (int32_t t = 0; t < nb_slice; ++t) {
    SliceDesc& d = hdesc[t];
    const int call = hstats[5 * t], cnew = hstats[5 * t + 1], ca = hstats[5 * t + 2], F = hstats[5 * t + 3], V = hstats[5 * t + 4];
    d.cell_off = (int32_t)cell_off;
    d.nb_new = cnew;
    d.nb_a = (d.flags & CH_FLAG_FIRST) ? ca : cnew;
    d.foot_off = (int32_t)foot_off;
    d.nb_foot = F;
    d.nb_cell = call;
    d.max_val = V;
    d.blob_off = (uint32_t)blob_off;
    // list rows: sum over units of the longest list, <= (contributions/4 + entries)/32 + longest list
    const int64_t contrib = (int64_t)(npc * (npc - 1)) * call;
    const int64_t rows_cap = (contrib / 4 + d.nb_entry) / 32 + (V + 3) / 4 + 2;
    const int units_cap = (d.nb_entry + 31) / 32;
    d.blob_cap = ch_blob_bytes((int)rows_cap, units_cap, d.nb_row, d.nb_entry);
    blob_off += d.blob_cap / 16;
    cell_off += (cnew + 1) & ~1; // even offsets: the connectivity of a slice is moved in 16-byte granules
    foot_off += F;
    all_cells += call;
    all_new += cnew;
    AFB_REQUIRE(blob_off < (1ll << 32) && cell_off < (1ll << 31) && foot_off < (1ll << 31), AFB_ERR_OVERFLOW, "chain inspector: plan exceeds 32-bit offsets");
  }
  P.nb_new_total = cell_off;
  P.nb_foot_total = foot_off;
  P.blob_units = blob_off;
  P.halo = ctx->nb_cell > 0 ? (double)all_new / (double)std::min<int64_t>(nb_own_cell, ctx->nb_cell) : 0.0;
  AFB_TRY(P.new_cells.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(cell_off, 1)));
  AFB_TRY(P.lconn.reserve(sizeof(ushort4) * (size_t)std::max<int64_t>(cell_off + 2, 2)));
  AFB_TRY(P.errflag.reserve(17 * sizeof(unsigned long long)));
  AFB_CUDA(cudaMemsetAsync(P.errflag.p, 0, 17 * sizeof(unsigned long long), st));
  AFB_TRY(P.foot.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(foot_off, 1)));
  AFB_TRY(P.blob.reserve(16 * (size_t)std::max<int64_t>(blob_off, 1)));
  AFB_CUDA(cudaMemcpyAsync(P.desc.p, hdesc.data(), sizeof(SliceDesc) * (size_t)nb_slice, cudaMemcpyHostToDevice, st));
  int32_t* seg_idx = P.scratch_d.as<int32_t>() + 2 * ((size_t)nb_slice + 2);
  CellCtx C{ ctx->conn.as<int32_t>(), P.node_slice.as<int32_t>(), P.node_lrow.as<int32_t>(), seg_idx, own, nb_own_cell };
  {
    const int g = std::min<int>(nb_slice, 8 * ctx->sm_count);
    AFB_TRY(ch_by_npc(npc, [&](auto N) {
      k_chain_mesh<decltype(N)::value><<<g, CB_THREADS, 0, st>>>(P.desc.as<SliceDesc>(), nb_slice, P.slice_nodes.as<int32_t>(), C, ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(),
                                                                  P.new_cells.as<int32_t>(), P.foot.as<int32_t>(), P.lconn.as<ushort4>(), ctx->tmp_flag.as<int>());
      return AFB_OK;
    }));
    AFB_LAUNCH_CHECK(ctx);
  }
  {
    static const int prefill = [] {
      const char* e = getenv("AFB_CHAIN_PREFILL");
      return (e && e[0] == '0') ? 0 : 1;
    }();
    const size_t smem = sizeof(ChainListSmem);
    const int g = std::min<int>(nb_slice, 2 * ctx->sm_count);
    auto go = [&](auto kernel) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kernel<<<g, CB_THREADS, smem, st>>>(P.desc.as<SliceDesc>(), nb_slice, P.slice_nodes.as<int32_t>(), C, ctx->nc_ptr.as<int32_t>(), ctx->nc_list.as<int32_t>(), ctx->rows.as<int32_t>(),
                                          ctx->cols.as<int32_t>(), P.node_e0.as<int32_t>(), P.new_cells.as<int32_t>(), L.cn + 1, L.nreg, L.emax, prefill, P.blob.as<unsigned char>(),
                                          ctx->tmp_flag.as<int>());
      return cudaGetLastError();
    };
    AFB_CUDA(npc == 4 ? go(k_chain_lists<4>) : go(k_chain_lists<3>));
    ctx->launches++;
  }
  // ---- schedule: segments to CTAs, longest first ----
  hdesc.resize(nb_slice);
  AFB_CUDA(cudaMemcpyAsync(hdesc.data(), P.desc.p, sizeof(SliceDesc) * (size_t)nb_slice, cudaMemcpyDeviceToHost, st));
  int err = 0;
  AFB_CUDA(cudaMemcpyAsync(&err, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  AFB_CUDA(cudaStreamSynchronize(st));
  if (err != 0) {
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    AFB_REQUIRE(false, AFB_ERR_CUDA, "chain inspector: plan inconsistency (code %d)", err);
  }
  struct Seg { int32_t first, count; int64_t cost; };
  std::vector<Seg> segs;
  for (int32_t t = 0; t < nb_slice; ++t) {
    if (hdesc[t].flags & CH_FLAG_FIRST) segs.push_back({ t, 0, 0 });
    AFB_REQUIRE(!segs.empty(), AFB_ERR_CUDA, "chain inspector: a chain does not start with a segment");
    Seg& s = segs.back();
    s.count++;
    s.cost += 3 * (int64_t)hdesc[t].nb_new + (int64_t)hdesc[t].nb_entry + 64;
    AFB_REQUIRE(hdesc[t].blob_bytes <= hdesc[t].blob_cap, AFB_ERR_CUDA, "chain inspector: record overflow");
  }
  P.nb_seg = (int32_t)segs.size();
  std::vector<int32_t> sorted(segs.size());
  for (size_t i = 0; i < segs.size(); ++i) sorted[i] = (int32_t)i;
  std::stable_sort(sorted.begin(), sorted.end(), [&](int32_t a, int32_t b) { return segs[a].cost > segs[b].cost; });
  const int G = std::max(1, std::min<int>(grid, (int)segs.size()));
  std::vector<int64_t> load((size_t)G, 0);
  std::vector<std::vector<int32_t>> mine((size_t)G);
  {
    // least-loaded CTA first (binary heap over (load, cta))
    std::vector<std::pair<int64_t, int>> heap;
    for (int c = 0; c < G; ++c) heap.push_back({ 0, c });
    auto cmp = [](const std::pair<int64_t, int>& a, const std::pair<int64_t, int>& b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), cmp);
    for (int32_t si : sorted) {
      std::pop_heap(heap.begin(), heap.end(), cmp);
      auto& top = heap.back();
      mine[top.second].push_back(si);
      top.first += segs[si].cost;
      std::push_heap(heap.begin(), heap.end(), cmp);
    }
  }
  std::vector<int32_t> order;
  order.reserve(nb_slice);
  std::vector<int32_t> cta_ptr((size_t)G + 1, 0);
  for (int c = 0; c < G; ++c) {
    cta_ptr[c] = (int32_t)order.size();
    for (int32_t si : mine[c])
      for (int32_t k = 0; k < segs[si].count; ++k) order.push_back(segs[si].first + k);
  }
  cta_ptr[G] = (int32_t)order.size();
  AFB_REQUIRE((int32_t)order.size() == nb_slice, AFB_ERR_CUDA, "chain inspector: schedule lost slices");
  AFB_TRY(P.order.reserve(sizeof(int32_t) * (size_t)nb_slice));
  AFB_TRY(P.cta_ptr.reserve(sizeof(int32_t) * ((size_t)G + 1)));
  AFB_CUDA(cudaMemcpyAsync(P.order.p, order.data(), sizeof(int32_t) * (size_t)nb_slice, cudaMemcpyHostToDevice, st));
  std::vector<SliceDesc> hexec((size_t)nb_slice);
  for (int32_t i = 0; i < nb_slice; ++i) hexec[i] = hdesc[order[i]];
  AFB_TRY(P.desc_exec.reserve(sizeof(SliceDesc) * (size_t)nb_slice));
  AFB_CUDA(cudaMemcpyAsync(P.desc_exec.p, hexec.data(), sizeof(SliceDesc) * (size_t)nb_slice, cudaMemcpyHostToDevice, st));
  {
    std::vector<int32_t> hoff((size_t)nb_slice + 1);
    int64_t run = 0;
    for (int32_t i = 0; i < nb_slice; ++i) {
      hoff[i] = (int32_t)run;
      run += ch_ib_bytes(hexec[i].nb_foot, hexec[i].nb_row) / 16;
      AFB_REQUIRE(run < (1ll << 31), AFB_ERR_OVERFLOW, "chain inspector: index blocks exceed 32-bit offsets");
    }
    hoff[nb_slice] = (int32_t)run;
    AFB_TRY(P.ib_off.reserve(sizeof(int32_t) * ((size_t)nb_slice + 1)));
    AFB_TRY(P.iblock.reserve(16 * (size_t)std::max<int64_t>(run, 1)));
    AFB_CUDA(cudaMemcpyAsync(P.ib_off.p, hoff.data(), sizeof(int32_t) * ((size_t)nb_slice + 1), cudaMemcpyHostToDevice, st));
    k_chain_iblock<<<nb_slice, 128, 0, st>>>(P.desc_exec.as<SliceDesc>(), P.ib_off.as<int32_t>(), nb_slice, P.foot.as<int32_t>(), P.slice_nodes.as<int32_t>(), P.iblock.as<unsigned char>());
    AFB_LAUNCH_CHECK(ctx);
  }
  AFB_CUDA(cudaMemcpyAsync(P.cta_ptr.p, cta_ptr.data(), sizeof(int32_t) * ((size_t)G + 1), cudaMemcpyHostToDevice, st));
  AFB_CUDA(cudaEventRecord(e1, st));
  AFB_CUDA(cudaEventSynchronize(e1));
  AFB_CUDA(cudaEventElapsedTime(&P.plan_ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  P.nb_slice = nb_slice;
  P.grid = G;
  P.geom = geom;
  P.mode = mode_flags;
  P.mesh_gen = ctx->mesh_gen;
  P.valid = true;
  return AFB_OK;
}

} // namespace afb
