// Sparsity pattern from the mesh tiling (steady-state BuildMatrix).
//
// Replaces the reference's BuildMatrix kernels: the sort-based sparsity
// of modules/testlab/CsrGpuBiliAssembly.cc:42-207 / femutils/BSRFormat.cc:799-1006 (pack edges,
// cub radix sort of 6*nbCell u64 keys, atomic degree count, scan, atomic column slot claim) and the
// connectivity-based one (femutils/BSRFormat.cc:445-790), which walks a node-node connectivity that
// Arcane builds on the host at init (MeshUtils::computeNodeNodeViaEdgeConnectivity).
//
// Here the init-time structure is the mesh tiling of tiles_plan.cu: per tile the footprint nodes in
// ascending id order and, per row, the incident cells' other nodes as packed 10-bit footprint
// indices.  One CTA per tile, one thread per row: the thread ORs its neighbours into a private
// bitmap over the footprint (shared memory, layout [word][row]: conflict-free), so duplicates vanish
// without hashing or sorting, the degree is a popcount, and walking the set bits emits the columns
// already in ascending order.  Pass 1 writes the degrees and parks the tile's columns in a tile-ordered
// scratch (coalesced); after the scan of the degrees pass 2 moves them to their rows (contiguous runs).
// 4 bytes per (row, incident cell) are streamed once; no atomics, no global sort, deterministic.
//
// Connectivity-based variant (AFB_SPARSITY_FROM_CONNECTIVITY, the twin of the reference's
// BSRFormat::computeSparsityAtomicFree, femutils/BSRFormat.cc:445-790, and of _buildMatrixNodeWiseCsr,
// modules/testlab/NodeWiseCsrBiliAssembly.cc:90-152): those walk a node-node connectivity that exists
// since init (MeshUtils::computeNodeNodeViaEdgeConnectivity, modules/testlab/FemModule.cc:124,136) --
// degree = neighbours + 1, scan, columns = the neighbours.  The init-time structure here is the
// tile-local node-node connectivity written once by the inspector (the bitmap kernel above in LOCAL
// mode): per node its degree, per tile row its neighbours (self included) as ascending 16-bit
// footprint indices.  A re-build is then: scan of the degrees -> row_index; k_pattern_nn_place
// translates the footprint indices to node ids and stores them to their rows (coalesced runs).
#include <chrono>
#include <algorithm>

#include "tiles.cuh"

namespace afb {

constexpr int PT_UNROLL = 8;

// block exclusive scan of one int per thread (blockDim <= TG_RMAX); s_wsum: one int per warp
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_wsum)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_wsum[warp] = inc;
  __syncthreads();
  int woff = 0;
  for (int w = 0; w < warp; ++w) woff += s_wsum[w];
  return woff + inc - v;
}

// pass 1: bitmaps -> degrees (to `deg_out`, node order) and the tile's columns, row by row, into the
// tile-ordered scratch (coalesced).  `stale` is raised when a tile holds more entries than its slot.
// LOCAL (inspector, once per tiling): the footprint indices themselves go to `scratch16` = the tile-local
// node-node connectivity of the connectivity-based re-build.
template <bool LOCAL>
__global__ void __launch_bounds__(TG_RMAX)
k_pattern_tiled_extract(const TileDesc* __restrict__ desc, const int32_t* __restrict__ tile_nodes, const uint16_t* __restrict__ rowf, const uint32_t* __restrict__ inc,
                        const uint2* __restrict__ inc_grp, const int32_t* __restrict__ foot, int32_t* __restrict__ deg_out, int32_t* __restrict__ scratch,
                        uint16_t* __restrict__ scratch16, uint16_t* __restrict__ row_e0, int* __restrict__ stale)
{
  extern __shared__ uint32_t pt_smem[];
  __shared__ int s_wsum[TG_RMAX / 32];
  const int32_t t = blockIdx.x;
  const TileDesc d = desc[t];
  const int R = d.nb_row;
  if (R == 0) return;
  const int RP = (R + 31) & ~31;             // bitmap row stride
  const int W = (d.nb_foot + 31) >> 5;       // bitmap words per row
  uint32_t* bm = pt_smem;                    // [W][RP]
  uint32_t* s_foot = bm + W * RP;            // [nb_foot]
  uint32_t* s_cols = s_foot + d.nb_foot;     // [nb_entry]
  int32_t* s_erow = reinterpret_cast<int32_t*>(s_cols + d.nb_entry);  // [RP + 1]
  uint16_t* s_pre = reinterpret_cast<uint16_t*>(s_erow + RP + 1);     // [W][RP]: set bits of the row in the words before w
  const int i = threadIdx.x;
  for (int f = threadIdx.x; f < d.nb_foot; f += blockDim.x) s_foot[f] = (uint32_t)__ldg(foot + d.foot_off + f);
  if (i < RP) {
    for (int w = 0; w < W; ++w) bm[w * RP + i] = 0u;
  }
  int32_t node = -1;
  if (i < R) {
    node = __ldg(tile_nodes + d.node_off + i);
    const unsigned self = __ldg(rowf + d.node_off + i);
    uint32_t* my = bm + i;
    my[(self >> 5) * RP] = 1u << (self & 31);
    const uint2 g = __ldg(inc_grp + (size_t)t * TG_GMAX + (i >> 5));
    const uint32_t* src = inc + d.inc_off + g.x + (i & 31);
    const int len = (int)g.y;
    int k = 0;
    for (; k + PT_UNROLL <= len; k += PT_UNROLL) {
      uint32_t w[PT_UNROLL];
#pragma unroll
      for (int q = 0; q < PT_UNROLL; ++q) w[q] = __ldg(src + (k + q) * 32);
#pragma unroll
      for (int q = 0; q < PT_UNROLL; ++q) {
        const unsigned f0 = w[q] & 1023u, f1 = (w[q] >> 10) & 1023u, f2 = (w[q] >> 20) & 1023u;
        my[(f0 >> 5) * RP] |= 1u << (f0 & 31);
        my[(f1 >> 5) * RP] |= 1u << (f1 & 31);
        my[(f2 >> 5) * RP] |= 1u << (f2 & 31);
      }
    }
    for (; k < len; ++k) {
      const uint32_t w = __ldg(src + k * 32);
      const unsigned f0 = w & 1023u, f1 = (w >> 10) & 1023u, f2 = (w >> 20) & 1023u;
      my[(f0 >> 5) * RP] |= 1u << (f0 & 31);
      my[(f1 >> 5) * RP] |= 1u << (f1 & 31);
      my[(f2 >> 5) * RP] |= 1u << (f2 & 31);
    }
  }
  int deg = 0;
  if (i < RP) {
    const uint32_t* my = bm + i;
    for (int w = 0; w < W; ++w) {
      s_pre[w * RP + i] = (uint16_t)deg;
      deg += __popc(my[w * RP]);
    }
  }
  const int e0 = block_exclusive_scan(deg, s_wsum); // barrier inside: s_foot complete
  if (i < R) {
    deg_out[node] = deg;
    s_erow[i] = e0;
    if (i == R - 1) s_erow[R] = e0 + deg;
    if constexpr (LOCAL) row_e0[d.node_off + i] = (uint16_t)e0;
  }
  __syncthreads();
  const int E = s_erow[R];
  if (E > d.nb_entry) { // the tiling was built for another pattern: never write past the tile's slot
    if (threadIdx.x == 0) atomicExch(stale, 1);
    return;
  }
  // one thread per bitmap word: balanced extraction, columns land in row order
  for (int idx = threadIdx.x; idx < W * RP; idx += blockDim.x) {
    uint32_t bits = bm[idx];
    if (!bits) continue;
    const int w = idx / RP, r = idx - w * RP;
    int k = s_erow[r] + s_pre[idx];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      s_cols[k++] = LOCAL ? (uint32_t)(w * 32 + b) : s_foot[w * 32 + b];
    }
  }
  __syncthreads();
  if constexpr (LOCAL) {
    uint16_t* out = scratch16 + d.ent_off;
    for (int e = threadIdx.x; e < E; e += blockDim.x) out[e] = (uint16_t)s_cols[e];
  }
  else {
    int32_t* out = scratch + d.ent_off;
    for (int e = threadIdx.x; e < E; e += blockDim.x) out[e] = (int32_t)s_cols[e];
  }
}

// connectivity-based re-build, after the scan of the init-time degrees: one CTA per tile, one thread per row
// notes its row's value offset for each of its entries, then one thread per entry translates the 16-bit
// footprint index to the node id and stores it (rows with consecutive node ids are adjacent: contiguous runs)
__global__ void __launch_bounds__(TG_RMAX)
k_pattern_nn_place(const TileDesc* __restrict__ desc, const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ rows, const uint16_t* __restrict__ nn_local,
                   const uint16_t* __restrict__ nn_e0, const int32_t* __restrict__ foot, int32_t* __restrict__ cols, int32_t* __restrict__ nz_per_row,
                   int32_t nb_node, int32_t* __restrict__ check /* device-mapped host words: rows[nb_node], stale flag */)
{
  __shared__ int32_t s_dbase[TG_EMAX];
  pdl_enter();
  const int32_t t = blockIdx.x;
  const TileDesc d = desc[t];
  const int R = d.nb_row, E = d.nb_entry;
  if (t == 0 && threadIdx.x == 0) check[0] = __ldg(rows + nb_node);
  if (R == 0) return;
  const int i = threadIdx.x;
  // everything addressed by the descriptor alone is requested at once; rows[node] and foot[local index] (the tile's
  // footprint: 1.4 KB, L1-resident after the first touch) are the only dependent loads; one barrier per tile
  constexpr int PER = 12;
  const uint16_t* src = nn_local + d.ent_off;
  const int32_t* ft = foot + d.foot_off;
  uint16_t loc[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int e = min(i + q * (int)blockDim.x, E - 1);
    loc[q] = __ldg(src + e);
  }
  int32_t node = 0;
  int e0 = 0, e1 = 0;
  if (i < R) {
    node = __ldg(tile_nodes + d.node_off + i);
    e0 = __ldg(nn_e0 + d.node_off + i);
    e1 = i + 1 < R ? (int)__ldg(nn_e0 + d.node_off + i + 1) : E;
  }
  int32_t col[PER];
#pragma unroll
  for (int q = 0; q < PER; ++q) col[q] = __ldg(ft + loc[q]);
  int bad = 0;
  if (i < R) {
    const int rb = __ldg(rows + node), deg = __ldg(rows + node + 1) - rb;
    nz_per_row[node] = deg;
    if (deg != e1 - e0) bad = 1; // the connectivity was built for another pattern
    else
      for (int e = e0; e < e1; ++e) s_dbase[e] = rb - e0;
  }
  if (__syncthreads_or(bad)) { // the one barrier of the tile
    if (i == 0) check[1] = 1;
    return;
  }
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    const int e = i + q * blockDim.x;
    if (e < E) cols[s_dbase[e] + e] = col[q];
  }
  for (int e = i + PER * blockDim.x; e < E; e += blockDim.x) cols[s_dbase[e] + e] = __ldg(ft + __ldg(src + e));
}

// pass 2 (after the scan of the degrees): the tile's columns move from the scratch to their rows;
// rows with consecutive node ids are adjacent in `cols`, so the stores are contiguous runs
__global__ void __launch_bounds__(TG_RMAX)
k_pattern_tiled_place(const TileDesc* __restrict__ desc, const int32_t* __restrict__ tile_nodes, const int32_t* __restrict__ rows, const int32_t* __restrict__ scratch,
                      int32_t* __restrict__ cols, int32_t* __restrict__ nz_per_row)
{
  __shared__ int s_wsum[TG_RMAX / 32];
  __shared__ int32_t s_erow[TG_RMAX + 1];
  __shared__ int32_t s_shift[TG_RMAX];
  __shared__ uint16_t s_etab[TG_EMAX / 8 + 1];
  const int32_t t = blockIdx.x;
  const TileDesc d = desc[t];
  const int R = d.nb_row;
  if (R == 0) return;
  const int i = threadIdx.x;
  int deg = 0, rb = 0;
  int32_t node = -1;
  if (i < R) {
    node = __ldg(tile_nodes + d.node_off + i);
    rb = __ldg(rows + node);
    deg = __ldg(rows + node + 1) - rb;
  }
  const int e0 = block_exclusive_scan(deg, s_wsum);
  if (i < R) {
    s_erow[i] = e0;
    s_shift[i] = rb - e0;
    nz_per_row[node] = deg;
    for (int q = (e0 + 7) >> 3; (q << 3) < e0 + deg; ++q) s_etab[q] = (uint16_t)i;
    if (i == R - 1) s_erow[R] = e0 + deg;
  }
  __syncthreads();
  const int E = min(s_erow[R], d.nb_entry);
  const int32_t* src = scratch + d.ent_off;
  // four independent loads in flight per thread
  for (int e = threadIdx.x; e < E; e += 4 * blockDim.x) {
    int32_t c[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int eq = e + q * blockDim.x;
      c[q] = eq < E ? __ldg(src + eq) : 0;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int eq = e + q * blockDim.x;
      if (eq < E) {
        int r = s_etab[eq >> 3];
        while (eq >= s_erow[r + 1]) ++r;
        cols[s_shift[r] + eq] = c[q];
      }
    }
  }
}

bool pattern_nn_ready(const afb_ctx* ctx)
{
  const TilePlan& P = ctx->plan;
  return pattern_tiled_ready(ctx) && P.nn_valid && P.nn_mesh_gen == ctx->mesh_gen;
}

bool pattern_tiled_ready(const afb_ctx* ctx)
{
  const TilePlan& P = ctx->plan;
  return P.mesh_valid && P.mesh_gen == ctx->mesh_gen && ctx->npc == ctx->dim + 1;
}

static int pattern_threads(const TilePlan& P) { return std::max(32, (P.max_rows + 31) & ~31); }

int pattern_tiled_extract(afb_ctx* ctx, int32_t* deg, int* stale)
{
  (void)stale;
  TilePlan& P = ctx->plan;
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), ctx->stream));
  if (P.nb_tile == 0) return AFB_OK;
  AFB_TRY(P.col_scratch.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(P.nb_entry, 1)));
  const int threads = pattern_threads(P);
  constexpr size_t WMAX = (TG_FMAX + 31) / 32;
  // bitmap [W][RP] + footprint ids + staged columns + row offsets + per-word prefix counts: bounded by the tile limits
  const size_t smem = sizeof(uint32_t) * (WMAX * (size_t)threads + TG_FMAX + TG_EMAX + (size_t)threads + 1) + sizeof(uint16_t) * WMAX * (size_t)threads;
  AFB_CUDA(cudaFuncSetAttribute(k_pattern_tiled_extract<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_pattern_tiled_extract<false><<<P.nb_tile, threads, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.tile_nodes.as<int32_t>(), P.rowf.as<uint16_t>(),
                                                                             P.inc.as<uint32_t>(), P.inc_grp.as<uint2>(), P.foot.as<int32_t>(), deg,
                                                                             P.col_scratch.as<int32_t>(), nullptr, nullptr, ctx->tmp_flag.as<int>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

// the three arrays of the tile-local node-node connectivity, one allocation (sizes: nb_node and the tiling's nb_entry)
int pattern_nn_reserve(afb_ctx* ctx)
{
  TilePlan& P = ctx->plan;
  return reserve_group(P.arena_nn, { { &P.nn_deg, sizeof(int32_t) * ((size_t)ctx->nb_node + 1) },
                                     { &P.nn_local, sizeof(uint16_t) * (size_t)std::max<int64_t>(P.nb_entry, 1) },
                                     { &P.nn_e0, sizeof(uint16_t) * ((size_t)ctx->nb_node + 1) } });
}

// inspector: the tile-local node-node connectivity (once per mesh tiling)
int pattern_nn_build(afb_ctx* ctx)
{
  TilePlan& P = ctx->plan;
  P.nn_valid = false;
  const bool trace = getenv("AFB_INSPECTOR_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* w) { if (trace) fprintf(stderr, "[inspector]     nn %8.1f us  %s\n", std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(), w); };
  AFB_TRY(pattern_nn_reserve(ctx));
  mark("reserved");
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), ctx->stream));
  AFB_CUDA(cudaMemsetAsync(P.nn_deg.p, 0, sizeof(int32_t) * ((size_t)ctx->nb_node + 1), ctx->stream));
  if (P.nb_tile > 0) {
    const int threads = pattern_threads(P);
    constexpr size_t WMAX = (TG_FMAX + 31) / 32;
    const size_t smem = sizeof(uint32_t) * (WMAX * (size_t)threads + TG_FMAX + TG_EMAX + (size_t)threads + 1) + sizeof(uint16_t) * WMAX * (size_t)threads;
    AFB_CUDA(cudaFuncSetAttribute(k_pattern_tiled_extract<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_pattern_tiled_extract<true><<<P.nb_tile, threads, smem, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.tile_nodes.as<int32_t>(), P.rowf.as<uint16_t>(),
                                                                              P.inc.as<uint32_t>(), P.inc_grp.as<uint2>(), P.foot.as<int32_t>(), P.nn_deg.as<int32_t>(),
                                                                              nullptr, P.nn_local.as<uint16_t>(), P.nn_e0.as<uint16_t>(), ctx->tmp_flag.as<int>());
    AFB_LAUNCH_CHECK(ctx);
  }
  mark("launched");
  int stale = 0;
  AFB_CUDA(cudaMemcpyAsync(&stale, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  mark("synchronised");
  AFB_REQUIRE(stale == 0, AFB_ERR_CUDA, "tile inspector: node-node connectivity does not fit the tiling");
  P.nn_mesh_gen = ctx->mesh_gen;
  P.nn_valid = true;
  return AFB_OK;
}

int pattern_nn_place(afb_ctx* ctx, int32_t* check)
{
  const TilePlan& P = ctx->plan;
  if (P.nb_tile == 0) return AFB_OK;
  AFB_CUDA(launch_pdl(k_pattern_nn_place, P.nb_tile, pattern_threads(P), 0, ctx->stream, P.tile_desc.as<TileDesc>(), P.tile_nodes.as<int32_t>(), ctx->rows.as<int32_t>(),
                      P.nn_local.as<uint16_t>(), P.nn_e0.as<uint16_t>(), P.foot.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->nz_per_row.as<int32_t>(), ctx->nb_node, check));
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

int pattern_tiled_place(afb_ctx* ctx)
{
  const TilePlan& P = ctx->plan;
  if (P.nb_tile == 0) return AFB_OK;
  k_pattern_tiled_place<<<P.nb_tile, pattern_threads(P), 0, ctx->stream>>>(P.tile_desc.as<TileDesc>(), P.tile_nodes.as<int32_t>(), ctx->rows.as<int32_t>(),
                                                                           P.col_scratch.as<int32_t>(), ctx->cols.as<int32_t>(), ctx->nz_per_row.as<int32_t>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

} // namespace afb
