// Mesh connectivity (node -> cells) and the sparsity pattern (row_index / columns).
//
// Replaces, with one warp per node and no global sort:
//   - Arcane's nodeCell connectivity view used by the node-wise back-ends
//     (modules/testlab/NodeWiseCsrBiliAssembly.cc:179,188; femutils/BSRFormat.h:411,428);
//   - the sort-based sparsity of the reference: pack edges -> cub radix sort of
//     6*nbCell u64 keys -> unique-edge degree count (atomics) -> exclusive scan -> atomic
//     column slot claim (modules/testlab/CsrGpuBiliAssembly.cc:42-207,
//     femutils/BSRFormat.cc:799-1006) and the connectivity-based one
//     (femutils/BSRFormat.cc:445-790, NodeWiseCsrBiliAssembly.cc:90-152).
// Here each warp loads the cells incident to its node (<= 4 per lane in registers), and
// extracts the distinct neighbour ids in ascending order with a warp-wide REDUX.MIN per
// entry: degree pass -> prefix scan -> column pass.  Columns come out sorted, the
// diagonal sits at its sorted position, the result is deterministic.
#include "afb_internal.h"

namespace afb {

constexpr unsigned UMAX = 0xFFFFFFFFu;
constexpr int WARPS_PER_BLOCK = 8;

// ---------------------------------------------------------------------------------------------
// node -> cell lists
// ---------------------------------------------------------------------------------------------
template <int NPC>
__global__ void __launch_bounds__(256) k_count_node_cells(const int32_t* __restrict__ conn, int64_t nb_cell, int32_t* __restrict__ deg)
{
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb_cell) return;
  const int32_t* cn = conn + c * NPC;
  if constexpr (NPC == 4) {
    int4 v = *reinterpret_cast<const int4*>(cn);
    atomicAdd(deg + v.x, 1); atomicAdd(deg + v.y, 1); atomicAdd(deg + v.z, 1); atomicAdd(deg + v.w, 1);
  }
  else {
#pragma unroll
    for (int i = 0; i < NPC; ++i) atomicAdd(deg + cn[i], 1);
  }
}

template <int NPC>
__global__ void __launch_bounds__(256) k_fill_node_cells(const int32_t* __restrict__ conn, int64_t nb_cell, const int32_t* __restrict__ ptr,
                                                          int32_t* __restrict__ fill, int32_t* __restrict__ list)
{
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nb_cell) return;
  const int32_t* cn = conn + c * NPC;
#pragma unroll
  for (int i = 0; i < NPC; ++i) {
    int32_t n = cn[i];
    int pos = atomicAdd(fill + n, 1);
    list[ptr[n] + pos] = (int32_t)c;
  }
}

// ascending extraction over K registers per lane; calls emit(count, value) for each distinct value
template <int R, class Emit>
__device__ __forceinline__ int warp_extract_sorted_unique(const unsigned (&cand)[R], Emit emit)
{
  unsigned lo = 0;
  int count = 0;
  while (true) {
    unsigned m = UMAX;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      unsigned c = cand[r];
      m = min(m, c >= lo ? c : UMAX);
    }
    m = __reduce_min_sync(0xffffffffu, m);
    if (m == UMAX) break;
    emit(count, m);
    ++count;
    lo = m + 1u;
  }
  return count;
}

template <int K>
__device__ __forceinline__ void sort_list_regs(int32_t* list, int beg, int val, int lane)
{
  unsigned cand[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int idx = lane + 32 * k;
    cand[k] = idx < val ? (unsigned)list[beg + idx] : UMAX;
  }
  unsigned mine = 0;
  int count = warp_extract_sorted_unique<K>(cand, [&](int cnt, unsigned m) {
    if ((cnt & 31) == lane) mine = m;
    if ((cnt & 31) == 31) list[beg + (cnt & ~31) + lane] = (int32_t)mine;
  });
  int rem = count & 31;
  if (lane < rem) list[beg + (count & ~31) + lane] = (int32_t)mine;
}

// Sort each node's cell list ascending (the atomic fill order is arbitrary).  Cell ids of a
// node are distinct, so "sorted unique" == sorted.
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK) k_sort_node_cells(const int32_t* __restrict__ ptr, int32_t* __restrict__ list, int32_t nb_node, int* __restrict__ max_valence)
{
  const int lane = threadIdx.x & 31;
  int64_t node = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (node >= nb_node) return;
  int beg = ptr[node], val = ptr[node + 1] - beg;
  if (lane == 0 && val > *max_valence) atomicMax(max_valence, val);
  if (val <= 1) return;
  if (val <= 32) {
    // rank sort: one item per lane, rank = number of smaller items (cell ids are distinct)
    const int32_t mine = lane < val ? list[beg + lane] : 0x7fffffff;
    int rank = 0;
    for (int i = 0; i < val; ++i) rank += (__shfl_sync(0xffffffffu, mine, i) < mine) ? 1 : 0;
    if (lane < val) list[beg + rank] = mine;
  }
  else if (val <= 64) sort_list_regs<2>(list, beg, val, lane);
  else if (val <= 128) sort_list_regs<4>(list, beg, val, lane);
  else if (val <= 256) sort_list_regs<8>(list, beg, val, lane);
  else if (lane == 0) {
    // pathological valence: serial insertion sort
    for (int i = 1; i < val; ++i) {
      int32_t x = list[beg + i];
      int j = i - 1;
      while (j >= 0 && list[beg + j] > x) { list[beg + j + 1] = list[beg + j]; --j; }
      list[beg + j + 1] = x;
    }
  }
}

int build_node_cells(afb_ctx* ctx)
{
  const int64_t nb_cell = ctx->nb_cell;
  const int32_t nb_node = ctx->nb_node;
  const int npc = ctx->npc;
  AFB_REQUIRE((int64_t)npc * nb_cell < 2147483647LL, AFB_ERR_OVERFLOW, "node-cell connectivity exceeds Int32 (%lld incidences)", (long long)(npc * nb_cell));
  AFB_TRY(ctx->nc_ptr.reserve(sizeof(int32_t) * ((size_t)nb_node + 1)));
  AFB_TRY(ctx->nc_list.reserve(sizeof(int32_t) * (size_t)(npc * nb_cell)));
  AFB_TRY(ctx->tmp_i32a.reserve(sizeof(int32_t) * ((size_t)nb_node + 2)));
  int32_t* deg = ctx->tmp_i32a.as<int32_t>();
  const int32_t* conn = ctx->conn.as<int32_t>();
  AFB_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * ((size_t)nb_node + 2), ctx->stream));
  int grid = grid_for(nb_cell, 256);
  if (nb_cell > 0) {
    switch (npc) {
    case 3: k_count_node_cells<3><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, deg); break;
    case 4: k_count_node_cells<4><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, deg); break;
    case 6: k_count_node_cells<6><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, deg); break;
    case 8: k_count_node_cells<8><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, deg); break;
    case 10: k_count_node_cells<10><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, deg); break;
    }
    AFB_LAUNCH_CHECK(ctx);
  }
  AFB_TRY(exclusive_scan_i32(ctx, deg, ctx->nc_ptr.as<int32_t>(), nb_node));
  AFB_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * ((size_t)nb_node + 2), ctx->stream));
  if (nb_cell > 0) {
    int32_t* ptr = ctx->nc_ptr.as<int32_t>();
    int32_t* list = ctx->nc_list.as<int32_t>();
    switch (npc) {
    case 3: k_fill_node_cells<3><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, ptr, deg, list); break;
    case 4: k_fill_node_cells<4><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, ptr, deg, list); break;
    case 6: k_fill_node_cells<6><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, ptr, deg, list); break;
    case 8: k_fill_node_cells<8><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, ptr, deg, list); break;
    case 10: k_fill_node_cells<10><<<grid, 256, 0, ctx->stream>>>(conn, nb_cell, ptr, deg, list); break;
    }
    AFB_LAUNCH_CHECK(ctx);
    int* maxv = reinterpret_cast<int*>(deg + nb_node + 1); // zeroed above... but fill[] used deg[0..nb_node)
    k_sort_node_cells<<<grid_for(nb_node, WARPS_PER_BLOCK), 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(ptr, list, nb_node, maxv);
    AFB_LAUNCH_CHECK(ctx);
    AFB_CUDA(cudaMemcpyAsync(&ctx->max_valence, maxv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  }
  return AFB_OK;
}

// ---------------------------------------------------------------------------------------------
// row degree / columns
// ---------------------------------------------------------------------------------------------
template <int NPC, int K>
__device__ __forceinline__ void load_candidates(unsigned (&cand)[K * NPC], const int32_t* __restrict__ conn, const int32_t* __restrict__ list, int beg, int val, int lane)
{
#pragma unroll
  for (int k = 0; k < K; ++k) {
    int idx = lane + 32 * k;
    if (idx < val) {
      const int32_t* cn = conn + (int64_t)list[beg + idx] * NPC;
      if constexpr (NPC == 4) {
        int4 v = __ldg(reinterpret_cast<const int4*>(cn));
        cand[k * 4 + 0] = (unsigned)v.x; cand[k * 4 + 1] = (unsigned)v.y; cand[k * 4 + 2] = (unsigned)v.z; cand[k * 4 + 3] = (unsigned)v.w;
      }
      else {
#pragma unroll
        for (int i = 0; i < NPC; ++i) cand[k * NPC + i] = (unsigned)__ldg(cn + i);
      }
    }
    else {
#pragma unroll
      for (int i = 0; i < NPC; ++i) cand[k * NPC + i] = UMAX;
    }
  }
}

template <int NPC, int K, bool WRITE>
__device__ __forceinline__ int row_unique_regs(const int32_t* __restrict__ conn, const int32_t* __restrict__ list, int beg, int val, int lane, int32_t* __restrict__ cols, int rowbeg)
{
  unsigned cand[K * NPC];
  load_candidates<NPC, K>(cand, conn, list, beg, val, lane);
  unsigned mine = 0;
  int count = warp_extract_sorted_unique<K * NPC>(cand, [&](int cnt, unsigned m) {
    if constexpr (WRITE) {
      if ((cnt & 31) == lane) mine = m;
      if ((cnt & 31) == 31) cols[rowbeg + (cnt & ~31) + lane] = (int32_t)mine;
    }
  });
  if constexpr (WRITE) {
    int rem = count & 31;
    if (lane < rem) cols[rowbeg + (count & ~31) + lane] = (int32_t)mine;
  }
  return count;
}

// valence > 128: candidates are re-read (L1/L2) for every extracted entry
template <int NPC, bool WRITE>
__device__ __noinline__ int row_unique_slow(const int32_t* __restrict__ conn, const int32_t* __restrict__ list, int beg, int val, int lane, int32_t* __restrict__ cols, int rowbeg)
{
  unsigned lo = 0;
  int count = 0;
  while (true) {
    unsigned m = UMAX;
    for (int idx = lane; idx < val; idx += 32) {
      const int32_t* cn = conn + (int64_t)list[beg + idx] * NPC;
#pragma unroll
      for (int i = 0; i < NPC; ++i) {
        unsigned c = (unsigned)__ldg(cn + i);
        m = min(m, c >= lo ? c : UMAX);
      }
    }
    m = __reduce_min_sync(0xffffffffu, m);
    if (m == UMAX) break;
    if (WRITE && lane == 0) cols[rowbeg + count] = (int32_t)m;
    ++count;
    lo = m + 1u;
  }
  return count;
}

// Fast path (valence <= 32, the common case): the warp's candidates (one incident cell per
// lane) are de-duplicated in a per-warp shared-memory hash table with ATOMS.CAS, counted with
// ballots, and -- in the column pass -- compacted and rank-sorted (rank = number of smaller
// ids, read back as shared-memory broadcasts).  ~3x fewer instructions than the REDUX.MIN
// extraction, which stays as the fallback for high valence or a full table.
constexpr int HASH_SLOTS = 128;
constexpr unsigned HASH_EMPTY = 0xFFFFFFFFu;

template <int NPC, bool WRITE>
__device__ __forceinline__ int row_unique_hash(const int32_t* __restrict__ conn, const int32_t* __restrict__ list, int beg, int val, int lane,
                                               unsigned* __restrict__ tab, unsigned* __restrict__ compact, int32_t* __restrict__ cols, int rowbeg)
{
#pragma unroll
  for (int k = 0; k < HASH_SLOTS / 32; ++k) tab[lane + 32 * k] = HASH_EMPTY;
  __syncwarp();
  bool overflow = false;
  if (lane < val) {
    unsigned cand[NPC];
    const int32_t* cn = conn + (int64_t)list[beg + lane] * NPC;
    if constexpr (NPC == 4) {
      int4 v = __ldg(reinterpret_cast<const int4*>(cn));
      cand[0] = (unsigned)v.x; cand[1] = (unsigned)v.y; cand[2] = (unsigned)v.z; cand[3] = (unsigned)v.w;
    }
    else {
#pragma unroll
      for (int i = 0; i < NPC; ++i) cand[i] = (unsigned)__ldg(cn + i);
    }
#pragma unroll
    for (int i = 0; i < NPC; ++i) {
      unsigned h = (cand[i] * 2654435761u) >> 25; // 7 bits
      int probes = 0;
      while (true) {
        unsigned old = atomicCAS(tab + h, HASH_EMPTY, cand[i]);
        if (old == HASH_EMPTY || old == cand[i]) break;
        h = (h + 1) & (HASH_SLOTS - 1);
        if (++probes >= HASH_SLOTS) { overflow = true; break; }
      }
    }
  }
  __syncwarp();
  if (__any_sync(0xffffffffu, overflow)) return -1;
  unsigned item[HASH_SLOTS / 32];
  unsigned bal[HASH_SLOTS / 32];
  int count = 0;
#pragma unroll
  for (int k = 0; k < HASH_SLOTS / 32; ++k) {
    item[k] = tab[lane + 32 * k];
    bal[k] = __ballot_sync(0xffffffffu, item[k] != HASH_EMPTY);
    count += __popc(bal[k]);
  }
  if constexpr (WRITE) {
    const unsigned lt = (1u << lane) - 1u;
    int base = 0;
#pragma unroll
    for (int k = 0; k < HASH_SLOTS / 32; ++k) {
      if (item[k] != HASH_EMPTY) compact[base + __popc(bal[k] & lt)] = item[k];
      base += __popc(bal[k]);
    }
    __syncwarp();
    for (int j = lane; j < count; j += 32) {
      const unsigned mine = compact[j];
      int rank = 0;
      for (int i = 0; i < count; ++i) rank += (compact[i] < mine) ? 1 : 0;
      cols[rowbeg + rank] = (int32_t)mine;
    }
    __syncwarp();
  }
  return count;
}

template <int NPC, bool WRITE>
__global__ void __launch_bounds__(32 * WARPS_PER_BLOCK)
k_row_unique(const int32_t* __restrict__ conn, const int32_t* __restrict__ ptr, const int32_t* __restrict__ list, int32_t nb_node,
             int32_t* __restrict__ deg_out, const int32_t* __restrict__ rows, int32_t* __restrict__ cols, int32_t* __restrict__ nz_per_row)
{
  const int lane = threadIdx.x & 31;
  int64_t node = (int64_t)blockIdx.x * WARPS_PER_BLOCK + (threadIdx.x >> 5);
  if (node >= nb_node) return;
  int beg = ptr[node], val = ptr[node + 1] - beg;
  int rowbeg = 0;
  if constexpr (WRITE) rowbeg = rows[node];
  __shared__ unsigned s_tab[WARPS_PER_BLOCK][HASH_SLOTS];
  __shared__ unsigned s_compact[WARPS_PER_BLOCK][HASH_SLOTS];
  const int w = threadIdx.x >> 5;
  int count = -1;
  if (val == 0) { // isolated node: diagonal only
    count = 1;
    if (WRITE && lane == 0) cols[rowbeg] = (int32_t)node;
  }
  else if (val <= 32) {
    count = row_unique_hash<NPC, WRITE>(conn, list, beg, val, lane, s_tab[w], s_compact[w], cols, rowbeg);
    if (count < 0) count = row_unique_regs<NPC, 1, WRITE>(conn, list, beg, val, lane, cols, rowbeg);
  }
  else if (val <= 64) count = row_unique_regs<NPC, 2, WRITE>(conn, list, beg, val, lane, cols, rowbeg);
  else if (val <= 128 && NPC <= 6) count = row_unique_regs<NPC, 4, WRITE>(conn, list, beg, val, lane, cols, rowbeg);
  else count = row_unique_slow<NPC, WRITE>(conn, list, beg, val, lane, cols, rowbeg);
  if (lane == 0) {
    if constexpr (WRITE) nz_per_row[node] = count;
    else deg_out[node] = count;
  }
}

template <int NPC>
static int launch_row_unique(afb_ctx* ctx, bool write, int32_t* deg)
{
  const int32_t nb_node = ctx->nb_node;
  int grid = grid_for(nb_node, WARPS_PER_BLOCK);
  const int32_t* conn = ctx->conn.as<int32_t>();
  const int32_t* ptr = ctx->nc_ptr.as<int32_t>();
  const int32_t* list = ctx->nc_list.as<int32_t>();
  if (!write)
    k_row_unique<NPC, false><<<grid, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(conn, ptr, list, nb_node, deg, nullptr, nullptr, nullptr);
  else
    k_row_unique<NPC, true><<<grid, 32 * WARPS_PER_BLOCK, 0, ctx->stream>>>(conn, ptr, list, nb_node, nullptr, ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(),
                                                                            ctx->nz_per_row.as<int32_t>());
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

static int dispatch_row_unique(afb_ctx* ctx, bool write, int32_t* deg)
{
  switch (ctx->npc) {
  case 3: return launch_row_unique<3>(ctx, write, deg);
  case 4: return launch_row_unique<4>(ctx, write, deg);
  case 6: return launch_row_unique<6>(ctx, write, deg);
  case 8: return launch_row_unique<8>(ctx, write, deg);
  case 10: return launch_row_unique<10>(ctx, write, deg);
  }
  set_error("unsupported nodes_per_cell %d", ctx->npc);
  return AFB_ERR_UNSUPPORTED;
}

// completes the deferred check of the last steady-state re-build (see build_pattern)
int verify_pending(afb_ctx* ctx)
{
  if (!ctx->check_pending) return AFB_OK;
  ctx->check_pending = false;
  AFB_CUDA(cudaEventSynchronize(ctx->check_event));
  AFB_REQUIRE(ctx->pin_check[1] == 0, AFB_ERR_CUDA, "tiled BuildMatrix: a tile produced more entries than its scratch capacity (stale tiling)");
  AFB_REQUIRE((int64_t)ctx->pin_check[0] == ctx->nnz, AFB_ERR_CUDA, "BuildMatrix re-build: the device computed nnz=%d, the host assumed %lld (same mesh)", ctx->pin_check[0],
              (long long)ctx->nnz);
  return AFB_OK;
}

int build_pattern(afb_ctx* ctx)
{
  AFB_TRY(verify_pending(ctx));
  const int32_t nb_node = ctx->nb_node;
  const int b = ctx->b;
  AFB_TRY(ctx->rows.reserve(sizeof(int32_t) * ((size_t)nb_node + 1)));
  AFB_TRY(ctx->nz_per_row.reserve(sizeof(int32_t) * ((size_t)nb_node + 1)));
  const bool fast = pattern_rows_supported(ctx);
  bool done = false;
  const bool steady = ctx->pattern_mesh_gen == ctx->mesh_gen && ctx->cols.p;
  if (steady && ctx->sparsity_algo != AFB_SPARSITY_FROM_CELLS && pattern_nn_ready(ctx) && ctx->nnz_mesh_gen == ctx->mesh_gen && ctx->pin_check) {
    // steady state, connectivity-based (the reference's computeSparsityAtomicFree walks Arcane's init-time
    // node-node connectivity: femutils/BSRFormat.cc:445-790): scan of the init-time degrees -> row_index,
    // columns from the tile-local node-node connectivity.  nnz is a function of the mesh alone and already
    // known on the host: no host synchronisation; rows[nb_node] and the stale flag are compared later
    // (verify_pending).
    // the two check words live in device-mapped pinned memory and are written by the place kernel itself:
    // no memset and no copy between the BuildMatrix kernels and the assembly that follows
    ctx->pin_check[0] = -1; // (the previous check was completed by verify_pending above: nothing is in flight)
    ctx->pin_check[1] = 0;
    AFB_TRY(exclusive_scan_i32(ctx, ctx->plan.nn_deg.as<int32_t>(), ctx->rows.as<int32_t>(), nb_node));
    AFB_TRY(pattern_nn_place(ctx, ctx->pin_check_dev));
    if (ctx->plan.nb_tile == 0) ctx->pin_check[0] = (int32_t)ctx->nnz; // no tile, no kernel: nothing to check
    AFB_CUDA(cudaEventRecord(ctx->check_event, ctx->stream));
    ctx->check_pending = true;
    done = true;
  }
  if (!done && steady && pattern_tiled_ready(ctx)) {
    // steady state from the cells, with the mesh tiling available (the reference re-builds the sparsity on every
    // AssembleBilinearOperator): per-tile bitmap kernel, degree -> scan -> columns
    AFB_TRY(ctx->tmp_i32b.reserve(sizeof(int32_t) * ((size_t)nb_node + 1)));
    int32_t* deg = ctx->tmp_i32b.as<int32_t>();
    int stale = 0;
    AFB_TRY(pattern_tiled_extract(ctx, deg, &stale));
    AFB_TRY(exclusive_scan_i32(ctx, deg, ctx->rows.as<int32_t>(), nb_node));
    if (ctx->nnz_mesh_gen == ctx->mesh_gen && ctx->pin_check) {
      // Same mesh as the previous complete build: nnz (a function of the mesh alone) is already known on
      // the host, so the host does not wait for the device here.  The device still recomputes everything;
      // its rows[nb_node] and stale flag are copied back asynchronously and compared at the next host
      // synchronisation point or re-build (verify_pending).
      AFB_CUDA(cudaMemcpyAsync(ctx->pin_check, ctx->rows.as<int32_t>() + nb_node, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      AFB_CUDA(cudaMemcpyAsync(ctx->pin_check + 1, ctx->tmp_flag.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      AFB_CUDA(cudaEventRecord(ctx->check_event, ctx->stream));
      ctx->check_pending = true;
    }
    else {
      int32_t nnz32 = 0;
      AFB_CUDA(cudaMemcpyAsync(&nnz32, ctx->rows.as<int32_t>() + nb_node, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
      AFB_CUDA(cudaMemcpyAsync(&stale, ctx->tmp_flag.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      AFB_CUDA(cudaStreamSynchronize(ctx->stream));
      AFB_REQUIRE(stale == 0, AFB_ERR_CUDA, "tiled BuildMatrix: a tile produced more entries than its scratch capacity (stale tiling)");
      AFB_REQUIRE(nnz32 >= 0, AFB_ERR_OVERFLOW, "block nnz exceeds Int32");
      ctx->nnz = nnz32;
      ctx->nnz_mesh_gen = ctx->mesh_gen;
    }
    AFB_TRY(ctx->cols.reserve(sizeof(int32_t) * (size_t)ctx->nnz));
    AFB_TRY(pattern_tiled_place(ctx));
    done = true;
  }
  if (!done && fast && ctx->pattern_mesh_gen == ctx->mesh_gen && ctx->cols.p) {
    // steady state without a tiling: one fused pass into the buffers sized by the first build
    int exceeded = 0;
    int32_t nnz32 = 0;
    AFB_TRY(pattern_rows_fused(ctx, &exceeded, &nnz32));
    if (!exceeded) {
      AFB_REQUIRE(nnz32 >= 0, AFB_ERR_OVERFLOW, "block nnz exceeds Int32");
      ctx->nnz = nnz32;
      done = true;
    }
  }
  if (!done) {
    AFB_TRY(ctx->tmp_i32b.reserve(sizeof(int32_t) * ((size_t)nb_node + 1)));
    int32_t* deg = ctx->tmp_i32b.as<int32_t>();
    if (fast) AFB_TRY(pattern_rows_count(ctx, deg));
    else AFB_TRY(dispatch_row_unique(ctx, false, deg));
    AFB_TRY(exclusive_scan_i32(ctx, deg, ctx->rows.as<int32_t>(), nb_node));
    int32_t nnz32 = 0;
    AFB_CUDA(cudaMemcpyAsync(&nnz32, ctx->rows.as<int32_t>() + nb_node, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(cudaStreamSynchronize(ctx->stream));
    AFB_REQUIRE(nnz32 >= 0, AFB_ERR_OVERFLOW, "block nnz exceeds Int32");
    ctx->nnz = nnz32;
    AFB_TRY(ctx->cols.reserve(sizeof(int32_t) * (size_t)ctx->nnz));
    if (fast) AFB_TRY(pattern_rows_write(ctx));
    else AFB_TRY(dispatch_row_unique(ctx, true, nullptr));
    ctx->pattern_mesh_gen = ctx->mesh_gen;
  }
  if (!ctx->check_pending) ctx->nnz_mesh_gen = ctx->mesh_gen; // nnz was read back synchronously above
  AFB_REQUIRE((int64_t)ctx->nnz * b * b < 2147483647LL, AFB_ERR_OVERFLOW,
              "scalar nnz %lld exceeds the Int32 index space of the reference containers (femutils/BSRFormat.cc:362-364)", (long long)(ctx->nnz * b * b));
  AFB_TRY(ctx->values.reserve(sizeof(double) * (size_t)ctx->nnz * b * b));
  AFB_TRY(ctx->rhs.reserve(sizeof(double) * (size_t)nb_node * b));
  // CsrFormat::initialize fills the values with 0 (femutils/CsrFormatMatrix.cc:35-58).  Here the fill is
  // deferred to the first reader/accumulator (ensure_values_zeroed): a fresh tiled assembly writes
  // every entry exactly once and never needs it.
  ctx->values_dirty = true;
  AFB_CUDA(cudaMemsetAsync(ctx->rhs.p, 0, sizeof(double) * (size_t)nb_node * b, ctx->stream));
  return AFB_OK;
}

} // namespace afb
