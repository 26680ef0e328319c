// Sparsity pattern, thread-per-row kernel (P1 and Tri6 cells).
//
// Replaces the reference's BuildMatrix kernels K1-K10: pack edges -> cub radix
// sort of 6*nbCell u64 keys -> unique-edge degree count (atomics) -> exclusive scan -> atomic
// column slot claim (modules/testlab/CsrGpuBiliAssembly.cc:42-207, femutils/BSRFormat.cc:799-1006)
// and the connectivity-based variant (femutils/BSRFormat.cc:445-790).
//
// One thread owns one row.  It walks the node's incident cells (Arcane's nodeCell view, built by
// connectivity.cu), de-duplicates the candidate neighbours in a private open-addressing hash
// table living in shared memory (layout tab[slot][lane]: every access of a warp is bank-conflict
// free whatever the slot), keeps the distinct ids in a private list, insertion-sorts that list
// (<= ~30 ids) and emits it.  No atomics, no global sort; ~25 warp-instructions per row against
// ~300 for the warp-per-row REDUX.MIN extraction (kept as the fallback for very high valence and
// for Tet10).
//
// Three modes:
//   COUNT : degree only (first build of a mesh: sizes the column/value arrays)
//   WRITE : columns at the offsets of a scanned row array
//   FUSED : single pass -- block-level decoupled look-back over the per-block degree sums gives
//           the row offsets while the sorted lists are still in shared memory (steady state:
//           re-assembly on the same mesh re-builds the pattern, as the reference does on every
//           AssembleBilinearOperator, with no second walk and no separate scan kernel)
#include "afb_internal.h"

namespace afb {

constexpr int PR_THREADS = 128;
constexpr int PR_WARPS = PR_THREADS / 32;
constexpr unsigned PR_EMPTY = 0xFFFFFFFFu;
constexpr unsigned long long LB_FLAG_A = 1ull << 62; // aggregate of this block available
constexpr unsigned long long LB_FLAG_P = 2ull << 62; // inclusive prefix available
constexpr unsigned long long LB_VALUE = (1ull << 62) - 1;

enum { PR_COUNT = 0, PR_WRITE = 1, PR_FUSED = 2 };

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// insert id into the private table column (tab[s*32]) / list (items[i*32]); returns false when full
template <int SLOTS, int ITEMS>
__device__ __forceinline__ bool pr_insert(unsigned* __restrict__ tab, unsigned* __restrict__ items, int& count, unsigned id)
{
  unsigned h = (id * 0x9E3779B1u) >> (32 - (SLOTS == 32 ? 5 : (SLOTS == 64 ? 6 : 7)));
  while (true) {
    const unsigned v = tab[h * 32];
    if (v == id) return true;
    if (v == PR_EMPTY) {
      if (count >= ITEMS) return false;
      tab[h * 32] = id;
      items[count * 32] = id;
      ++count;
      return true;
    }
    h = (h + 1) & (SLOTS - 1);
  }
}

template <int NPC, int SLOTS, int ITEMS, int MODE>
__global__ void __launch_bounds__(PR_THREADS)
k_pattern_rows(const int32_t* __restrict__ conn, const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int32_t nb_node,
               int32_t* __restrict__ deg_out,       // COUNT
               int32_t* __restrict__ rows,          // WRITE: in, FUSED: out (nb_node+1)
               int32_t* __restrict__ cols, int32_t* __restrict__ nz_per_row, int64_t capacity,
               unsigned long long* __restrict__ lb_state, int* __restrict__ lb_ticket, int* __restrict__ status /* [0]=overflow rows, [1]=capacity exceeded */)
{
  extern __shared__ unsigned pr_smem[];
  __shared__ int s_vb, s_base, s_warp_tot[PR_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* tab = pr_smem + warp * (SLOTS + ITEMS) * 32 + lane;
  unsigned* items = tab + SLOTS * 32;

  int vb = blockIdx.x;
  if constexpr (MODE == PR_FUSED) {
    if (threadIdx.x == 0) s_vb = atomicAdd(lb_ticket, 1);
    __syncthreads();
    vb = s_vb;
  }
  const int32_t r = vb * PR_THREADS + threadIdx.x;
  int count = 0;
  bool overflow = false;
  if (r < nb_node) {
#pragma unroll 8
    for (int s = 0; s < SLOTS; ++s) tab[s * 32] = PR_EMPTY;
    pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)r); // the diagonal (isolated nodes keep it)
    const int qb = __ldg(nc_ptr + r), qe = __ldg(nc_ptr + r + 1);
    int q = qb;
    if constexpr (NPC == 4) {
      for (; q + 2 <= qe && !overflow; q += 2) {
        const int32_t c0 = __ldg(nc_list + q), c1 = __ldg(nc_list + q + 1);
        const int4 a = __ldg(reinterpret_cast<const int4*>(conn) + c0);
        const int4 b = __ldg(reinterpret_cast<const int4*>(conn) + c1);
        bool ok = pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)a.x);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)a.y);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)a.z);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)a.w);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)b.x);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)b.y);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)b.z);
        ok = ok && pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)b.w);
        overflow = !ok;
      }
    }
    for (; q < qe && !overflow; ++q) {
      const int32_t* cn = conn + (int64_t)__ldg(nc_list + q) * NPC;
#pragma unroll
      for (int i = 0; i < NPC; ++i)
        if (!pr_insert<SLOTS, ITEMS>(tab, items, count, (unsigned)__ldg(cn + i))) overflow = true;
    }
    if (!overflow && MODE != PR_COUNT) {
      // insertion sort of the private list (ascending)
      for (int i = 1; i < count; ++i) {
        const unsigned x = items[i * 32];
        int j = i - 1;
        while (j >= 0) {
          const unsigned y = items[j * 32];
          if (y <= x) break;
          items[(j + 1) * 32] = y;
          --j;
        }
        items[(j + 1) * 32] = x;
      }
    }
  }
  // rows whose neighbourhood does not fit the private table: counted warp-cooperatively below
  const unsigned ovf_mask = __ballot_sync(0xffffffffu, overflow);
  if (ovf_mask) {
    unsigned m = ovf_mask;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int32_t rr = vb * PR_THREADS + warp * 32 + src;
      const int qb = __ldg(nc_ptr + rr), qe = __ldg(nc_ptr + rr + 1);
      // distinct count by repeated warp-wide minimum (same idea as k_row_unique, any valence)
      unsigned lo = 0;
      int cnt = 0;
      while (true) {
        unsigned mn = PR_EMPTY;
        for (int idx = qb + lane; idx < qe; idx += 32) {
          const int32_t* cn = conn + (int64_t)__ldg(nc_list + idx) * NPC;
#pragma unroll
          for (int i = 0; i < NPC; ++i) {
            const unsigned c = (unsigned)__ldg(cn + i);
            mn = min(mn, c >= lo ? c : PR_EMPTY);
          }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (mn == PR_EMPTY) break;
        ++cnt;
        lo = mn + 1u;
      }
      if (lane == src) count = cnt;
    }
    if (lane == 0 && status) atomicAdd(status, __popc(ovf_mask));
  }

  if constexpr (MODE == PR_COUNT) {
    if (r < nb_node) deg_out[r] = count;
    return;
  }

  // ---- offsets ---------------------------------------------------------------------------------
  int rowbeg = 0;
  if constexpr (MODE == PR_WRITE) {
    if (r < nb_node) rowbeg = rows[r];
  }
  else {
    // block exclusive scan of count
    int inc = count;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp_tot[warp] = inc;
    __syncthreads();
    int woff = 0, btot = 0;
#pragma unroll
    for (int w = 0; w < PR_WARPS; ++w) {
      const int t = s_warp_tot[w];
      if (w < warp) woff += t;
      btot += t;
    }
    const int excl = woff + inc - count;
    if (warp == 0) {
      if (lane == 0) {
        st_relaxed_u64(lb_state + vb, (vb == 0 ? LB_FLAG_P : LB_FLAG_A) | (unsigned long long)btot);
      }
      long long run = 0;
      if (vb > 0) {
        int look = vb - 1;
        while (true) {
          const int idx = look - lane;
          unsigned long long s = idx >= 0 ? ld_relaxed_u64(lb_state + idx) : LB_FLAG_P;
          while (__any_sync(0xffffffffu, (s >> 62) == 0)) {
            if ((s >> 62) == 0) s = ld_relaxed_u64(lb_state + idx);
          }
          const unsigned pmask = __ballot_sync(0xffffffffu, (s >> 62) == 2);
          long long v = (long long)(s & LB_VALUE);
          if (pmask) {
            const int first = __ffs(pmask) - 1;
            if (lane > first) v = 0;
          }
#pragma unroll
          for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
          run += v;
          if (pmask) break;
          look -= 32;
        }
        if (lane == 0) st_relaxed_u64(lb_state + vb, LB_FLAG_P | (unsigned long long)(run + btot));
      }
      if (lane == 0) {
        s_base = (int)run;
        if (run + btot > capacity) status[1] = 1;
        if ((int64_t)(vb + 1) * PR_THREADS >= nb_node) rows[nb_node] = (int32_t)(run + btot);
      }
    }
    __syncthreads();
    rowbeg = s_base + excl;
    if (r < nb_node) rows[r] = rowbeg;
    if ((int64_t)s_base + btot > capacity) return; // reported through status[1]; host re-runs two-pass
  }

  // ---- columns ---------------------------------------------------------------------------------
  if (r < nb_node) {
    nz_per_row[r] = count;
    if (!overflow)
      for (int i = 0; i < count; ++i) cols[rowbeg + i] = (int32_t)items[i * 32];
  }
  if (ovf_mask) {
    unsigned m = ovf_mask;
    while (m) {
      const int src = __ffs(m) - 1;
      m &= m - 1;
      const int32_t rr = vb * PR_THREADS + warp * 32 + src;
      const int rb = __shfl_sync(0xffffffffu, rowbeg, src);
      const int qb = __ldg(nc_ptr + rr), qe = __ldg(nc_ptr + rr + 1);
      unsigned lo = 0;
      int cnt = 0;
      while (true) {
        unsigned mn = PR_EMPTY;
        for (int idx = qb + lane; idx < qe; idx += 32) {
          const int32_t* cn = conn + (int64_t)__ldg(nc_list + idx) * NPC;
#pragma unroll
          for (int i = 0; i < NPC; ++i) {
            const unsigned c = (unsigned)__ldg(cn + i);
            mn = min(mn, c >= lo ? c : PR_EMPTY);
          }
        }
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (mn == PR_EMPTY) break;
        if (lane == 0) cols[rb + cnt] = (int32_t)mn;
        ++cnt;
        lo = mn + 1u;
      }
    }
  }
}

template <int NPC, int SLOTS, int ITEMS>
static int launch_pattern_rows(afb_ctx* ctx, int mode, int32_t* deg)
{
  const int32_t nb_node = ctx->nb_node;
  const int grid = grid_for(nb_node, PR_THREADS);
  const size_t smem = sizeof(unsigned) * PR_WARPS * (SLOTS + ITEMS) * 32;
  const int32_t* conn = ctx->conn.as<int32_t>();
  const int32_t* ptr = ctx->nc_ptr.as<int32_t>();
  const int32_t* list = ctx->nc_list.as<int32_t>();
  int* status = ctx->tmp_flag.as<int>();
  auto set_smem = [&](auto kernel) { return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); };
  if (mode == PR_COUNT) {
    AFB_CUDA(set_smem(k_pattern_rows<NPC, SLOTS, ITEMS, PR_COUNT>));
    k_pattern_rows<NPC, SLOTS, ITEMS, PR_COUNT><<<grid, PR_THREADS, smem, ctx->stream>>>(conn, ptr, list, nb_node, deg, nullptr, nullptr, nullptr, 0, nullptr, nullptr, status);
  }
  else if (mode == PR_WRITE) {
    AFB_CUDA(set_smem(k_pattern_rows<NPC, SLOTS, ITEMS, PR_WRITE>));
    k_pattern_rows<NPC, SLOTS, ITEMS, PR_WRITE><<<grid, PR_THREADS, smem, ctx->stream>>>(conn, ptr, list, nb_node, nullptr, ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(),
                                                                                          ctx->nz_per_row.as<int32_t>(), 0, nullptr, nullptr, status);
  }
  else {
    AFB_CUDA(set_smem(k_pattern_rows<NPC, SLOTS, ITEMS, PR_FUSED>));
    unsigned long long* state = ctx->tmp_lookback.as<unsigned long long>();
    k_pattern_rows<NPC, SLOTS, ITEMS, PR_FUSED><<<grid, PR_THREADS, smem, ctx->stream>>>(conn, ptr, list, nb_node, nullptr, ctx->rows.as<int32_t>(), ctx->cols.as<int32_t>(),
                                                                                          ctx->nz_per_row.as<int32_t>(), (int64_t)(ctx->cols.cap / sizeof(int32_t)), state,
                                                                                          reinterpret_cast<int*>(state + grid), status);
  }
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

static int dispatch_pattern_rows(afb_ctx* ctx, int mode, int32_t* deg)
{
  switch (ctx->npc) {
  case 3: return launch_pattern_rows<3, 32, 24>(ctx, mode, deg);
  case 4: return launch_pattern_rows<4, 64, 48>(ctx, mode, deg);
  case 6: return launch_pattern_rows<6, 64, 48>(ctx, mode, deg);
  }
  set_error("thread-per-row pattern kernel: unsupported nodes_per_cell %d", ctx->npc);
  return AFB_ERR_UNSUPPORTED;
}

bool pattern_rows_supported(const afb_ctx* ctx) { return ctx->npc == 3 || ctx->npc == 4 || ctx->npc == 6; }

// degree pass (first build)
int pattern_rows_count(afb_ctx* ctx, int32_t* deg)
{
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), ctx->stream));
  return dispatch_pattern_rows(ctx, PR_COUNT, deg);
}

// column pass at scanned offsets
int pattern_rows_write(afb_ctx* ctx)
{
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), ctx->stream));
  return dispatch_pattern_rows(ctx, PR_WRITE, nullptr);
}

// single pass into the existing column buffer; *exceeded = 1 when the buffer was too small
// (nothing usable was written: the caller falls back to count + scan + write)
int pattern_rows_fused(afb_ctx* ctx, int* exceeded, int32_t* nnz_out)
{
  const int grid = grid_for(ctx->nb_node, PR_THREADS);
  AFB_TRY(ctx->tmp_flag.reserve(2 * sizeof(int)));
  AFB_TRY(ctx->tmp_lookback.reserve(sizeof(unsigned long long) * ((size_t)grid + 1)));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_flag.p, 0, 2 * sizeof(int), ctx->stream));
  AFB_CUDA(cudaMemsetAsync(ctx->tmp_lookback.p, 0, sizeof(unsigned long long) * ((size_t)grid + 1), ctx->stream));
  AFB_TRY(dispatch_pattern_rows(ctx, PR_FUSED, nullptr));
  int st[2] = { 0, 0 };
  int32_t nnz = 0;
  AFB_CUDA(cudaMemcpyAsync(st, ctx->tmp_flag.p, sizeof(st), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaMemcpyAsync(&nnz, ctx->rows.as<int32_t>() + ctx->nb_node, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  *exceeded = st[1];
  *nnz_out = nnz;
  return AFB_OK;
}

} // namespace afb
