// Tiled gather, software-pipelined executor: one persistent CTA per SM, two element-matrix caches.
//
// Same plan and arithmetic as k_assemble_tiled (tiles_exec.cu); what changes is the schedule.  The
// per-tile phases there (A: geometry on the fp64 pipe, B: gather on the shared-memory pipe, C: write-out)
// run one after the other, separated by four block barriers, and two resident CTAs are left to overlap
// them by chance.  Here tile j's gather and tile j+1's geometry are ONE phase of the same warps (cache
// j&1 is read while cache (j+1)&1 is written), so both pipes have work at all times and a tile costs
// two barriers:
//
//   X(j)  stage rows(j)            B(j): lists/emap(j) [TMA] + Kc[j&1] -> vout        A(j+1): cx[(j+1)&1] -> Kc[(j+1)&1]
//         loads -> registers: lconn(j+2), coords(j+2), unit table(j+1), rows(j+1), desc(j+4)
//   ---- barrier ----
//   Y(j)  TMA lists/emap(j+2) -> buffers j&1        C(j): diagonal + coalesced rows out of vout
//         registers -> shared: cx(j+2) -> cx[j&1], unit table(j+1), desc(j+4); loads: foot idx(j+3), row node(j+2)
//   ---- barrier ----
//
// Every global load is consumed at least one phase after it was issued (indices two phases before the
// data they address), so no warp waits on a dependent load; the bulk inputs (contribution lists, entry
// maps) arrive through the TMA engine two tiles ahead.
#include <algorithm>
#include <cstdlib>

#include "element.cuh"
#include "tiles.cuh"

namespace afb {

constexpr int PG_THREADS = 768;
constexpr int PG_ROUNDS = (TG_CMAX + PG_THREADS - 1) / PG_THREADS;
constexpr int PG_NW = PG_THREADS / 32;
constexpr int PG_RING = 8; // descriptor ring (tiles j .. j+4 live)
static_assert(PG_THREADS >= TG_FMAX && PG_THREADS >= TG_RMAX && PG_THREADS >= TG_UMAX, "one thread per footprint node / row / unit");

template <int NPC> struct PipeK;
template <> struct PipeK<4> {
  static constexpr int N = 6;
  __device__ static __forceinline__ void compute(const double* __restrict__ cx, uint2 ln, const ElemParams&, double (&K)[6])
  {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    const double* p3 = cx + 3 * (ln.y >> 16);
    Tet4Geom g;
    g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(0, 3) * g.s;
    K[3] = g.dot(1, 2) * g.s; K[4] = g.dot(1, 3) * g.s; K[5] = g.dot(2, 3) * g.s;
  }
};
template <> struct PipeK<3> {
  static constexpr int N = 3;
  __device__ static __forceinline__ void compute(const double* __restrict__ cx, uint2 ln, const ElemParams& prm, double (&K)[6])
  {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    Tri3Geom g;
    g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], (prm.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(1, 2) * g.s;
    K[3] = K[4] = K[5] = 0.0;
  }
};

struct PipeSmem {
  double Kc[2][TG_ZERO + 2];
  double cx[2][3 * TG_FMAX];
  double vout[TG_EMAX];
  __align__(16) uint16_t lists[2][TG_LMAX];
  __align__(16) uint32_t emap[2][TG_UMAX * 32];
  int32_t rowbeg[TG_RMAX];        // first value of the row minus its first tile-local entry
  uint32_t rowinfo[TG_RMAX + 1];  // + sentinel
  uint32_t ubase[TG_UMAX];
  uint16_t ulen[TG_UMAX];
  uint16_t etab[TG_EMAX / 8];
  __align__(16) TileDesc desc[PG_RING];
  __align__(8) unsigned long long mbar[2];
};
static_assert(sizeof(PipeSmem) <= 232448, "the pipelined executor must fit one SM (227 KB)");

struct PipeArgs {
  const TileDesc* desc;
  int32_t nb_tile;
  const double* coords;
  const int32_t* foot;
  const ushort4* lconn;
  const int32_t* tile_nodes;
  const uint32_t* rowinfo;
  const int32_t* rows;
  const uint32_t* unit_base;
  const uint16_t* unit_len;
  const uint32_t* emap;
  const uint16_t* lists;
  double* values;
  int accumulate;
};

__device__ __forceinline__ void pipe_mbar_wait(uint32_t mbar, unsigned parity)
{
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}

// bulk copy of a tile's contribution lists (when they fit the staging buffer) and entry map
__device__ __forceinline__ void pipe_tma_tile(const TileDesc& d, const PipeArgs& A, PipeSmem& S, int buf)
{
  const uint32_t lbytes = d.list_len <= TG_LMAX ? (uint32_t)d.list_len * 2u : 0u;
  const uint32_t ebytes = (uint32_t)d.nb_unit * 128u;
  if (lbytes + ebytes == 0) return;
  const uint32_t mbar = smem_u32(&S.mbar[buf]);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(lbytes + ebytes) : "memory");
  if (lbytes)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.lists[buf])), "l"(A.lists + d.list_off),
                 "r"(lbytes), "r"(mbar)
                 : "memory");
  if (ebytes)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.emap[buf])),
                 "l"(A.emap + (size_t)d.unit_off * 32), "r"(ebytes), "r"(mbar)
                 : "memory");
}

template <int NPC>
__device__ __forceinline__ void pipe_phase_a(const TileDesc& d, const double* __restrict__ cx, double* __restrict__ Kc, const uint2 (&ln)[PG_ROUNDS], const ElemParams& prm)
{
#pragma unroll
  for (int r = 0; r < PG_ROUNDS; ++r) {
    const int lc = r * PG_THREADS + threadIdx.x;
    if (lc < d.nb_cell) {
      double K[6];
      PipeK<NPC>::compute(cx, ln[r], prm, K);
#pragma unroll
      for (int p = 0; p < PipeK<NPC>::N; ++p) Kc[p * TG_CS + lc] = K[p];
    }
  }
}

template <int NPC>
__global__ void __launch_bounds__(PG_THREADS, 1) k_assemble_tiled_pipe(PipeArgs A, ElemParams prm)
{
  extern __shared__ __align__(16) unsigned char pp_raw[];
  PipeSmem& S = *reinterpret_cast<PipeSmem*>(pp_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int DW = sizeof(TileDesc) / 4;
  const int tid = threadIdx.x;
  auto tile_of = [&](int j) { return (int64_t)blockIdx.x + (int64_t)j * gridDim.x; };
  // descriptor of tile j into its ring slot (zeros past the end: every loop over it is then empty)
  auto fetch_desc_word = [&](int j) -> int32_t {
    const int64_t t = tile_of(j);
    return (tid < DW && t < A.nb_tile) ? __ldg(reinterpret_cast<const int32_t*>(A.desc + t) + tid) : 0;
  };
  if (tid == 0) {
    S.Kc[0][TG_ZERO] = 0.0;
    S.Kc[1][TG_ZERO] = 0.0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 4 * DW) { // descriptors of tiles 0..3
    const int k = tid / DW, w = tid % DW;
    const int64_t t = tile_of(k);
    reinterpret_cast<int32_t*>(&S.desc[k])[w] = t < A.nb_tile ? __ldg(reinterpret_cast<const int32_t*>(A.desc + t) + w) : 0;
  }
  __syncthreads();
  if (tile_of(0) >= A.nb_tile) return;

  // ---- registers carried across phases ----
  uint2 ln[PG_ROUNDS];              // lconn of the tile whose phase A comes next
  double c0 = 0, c1 = 0, c2 = 0;    // coordinates on their way to cx
  int32_t fidx = 0, node = 0;       // level-1 indices
  int32_t r_rowbeg = 0;             // rows of the tile whose phase C comes next
  uint32_t r_ri = 0, r_ri1 = 0;
  uint32_t r_ubase = 0;
  uint16_t r_ulen = 0;

  auto load_lconn = [&](const TileDesc& d) {
#pragma unroll
    for (int r = 0; r < PG_ROUNDS; ++r) {
      const int lc = min(r * PG_THREADS + tid, d.nb_cell - 1); // never predicated per lane: see TilePrefetch::ln
      if (d.nb_cell > 0) ln[r] = __ldg(reinterpret_cast<const uint2*>(A.lconn) + d.cell_off + lc);
    }
  };
  auto load_fidx = [&](const TileDesc& d) { if (tid < d.nb_foot) fidx = __ldg(A.foot + d.foot_off + tid); };
  auto load_node = [&](const TileDesc& d) { if (tid < d.nb_row) node = __ldg(A.tile_nodes + d.node_off + tid); };
  auto load_coords = [&](const TileDesc& d) {
    if (tid < d.nb_foot) {
      const double* p = A.coords + 3 * (int64_t)fidx;
      c0 = __ldg(p); c1 = __ldg(p + 1); c2 = __ldg(p + 2);
    }
  };
  auto stage_coords = [&](const TileDesc& d, int buf) {
    if (tid < d.nb_foot) {
      S.cx[buf][3 * tid] = c0; S.cx[buf][3 * tid + 1] = c1; S.cx[buf][3 * tid + 2] = c2;
    }
  };
  auto load_rows = [&](const TileDesc& d) {
    if (tid < d.nb_row) {
      r_rowbeg = __ldg(A.rows + node);
      r_ri = __ldg(A.rowinfo + d.node_off + tid);
      r_ri1 = tid + 1 < d.nb_row ? __ldg(A.rowinfo + d.node_off + tid + 1) : pack_rowinfo(d.nb_entry, 0, false);
    }
  };
  auto stage_rows = [&](const TileDesc& d) {
    if (tid < d.nb_row) {
      const int e0 = rowinfo_erow(r_ri), e1 = rowinfo_erow(r_ri1);
      S.rowbeg[tid] = r_rowbeg - e0;
      S.rowinfo[tid] = r_ri;
      if (tid + 1 == d.nb_row) S.rowinfo[tid + 1] = pack_rowinfo(d.nb_entry, 0, false);
      for (int q = (e0 + 7) >> 3; (q << 3) < e1; ++q) S.etab[q] = (uint16_t)tid;
    }
  };
  auto load_units = [&](const TileDesc& d) {
    if (tid < d.nb_unit) {
      r_ubase = __ldg(A.unit_base + d.unit_off + tid);
      r_ulen = __ldg(A.unit_len + d.unit_off + tid);
    }
  };
  auto stage_units = [&](const TileDesc& d) {
    if (tid < d.nb_unit) {
      S.ubase[tid] = r_ubase;
      S.ulen[tid] = r_ulen;
    }
  };

  // ---- prologue: tile 0 is made ready the slow way (dependent loads exposed once) ----
  if (tid == 0) {
    pipe_tma_tile(S.desc[0], A, S, 0);
    pipe_tma_tile(S.desc[1], A, S, 1);
  }
  load_fidx(S.desc[0]);
  load_coords(S.desc[0]);
  stage_coords(S.desc[0], 0);
  load_fidx(S.desc[1]);
  load_coords(S.desc[1]);
  stage_coords(S.desc[1], 1);
  load_lconn(S.desc[0]);
  load_units(S.desc[0]);
  stage_units(S.desc[0]);
  load_node(S.desc[0]);
  __syncthreads();
  pipe_phase_a<NPC>(S.desc[0], S.cx[0], S.Kc[0], ln, prm);
  load_lconn(S.desc[1]);
  load_rows(S.desc[0]);
  load_fidx(S.desc[2]);
  load_node(S.desc[1]);
  __syncthreads();

  for (int j = 0; tile_of(j) < A.nb_tile; ++j) {
    const int buf = j & 1, nbuf = buf ^ 1;
    const TileDesc d = S.desc[j & (PG_RING - 1)];
    const TileDesc d1 = S.desc[(j + 1) & (PG_RING - 1)];
    // ================= X(j): B(j) and A(j+1) =================
    stage_rows(d);
    const int32_t desc_word = fetch_desc_word(j + 4);
    const bool staged = d.list_len <= TG_LMAX;
    if ((staged ? d.list_len : 0) + d.nb_unit > 0) pipe_mbar_wait(smem_u32(&S.mbar[buf]), (unsigned)((j >> 1) & 1));
    {
      const double* __restrict__ Kc = S.Kc[buf];
      const uint32_t* l32 = staged ? reinterpret_cast<const uint32_t*>(S.lists[buf]) : reinterpret_cast<const uint32_t*>(A.lists + d.list_off);
      constexpr uint32_t ZPAIR = (uint32_t)TG_ZERO | ((uint32_t)TG_ZERO << 16);
      for (int u = warp; u < d.nb_unit; u += PG_NW) {
        const uint32_t* l = l32 + (S.ubase[u] >> 1) + lane;
        const int len2 = S.ulen[u] >> 1;
        const uint32_t em = S.emap[buf][u * 32 + lane];
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 1
        for (int k = 0; k < len2; k += 2) {
          const uint32_t i0 = l[k * 32];
          const uint32_t i1 = (k + 1 < len2) ? l[(k + 1) * 32] : ZPAIR;
          acc0 += Kc[i0 & 0xFFFFu]; acc1 += Kc[i0 >> 16];
          acc0 += Kc[i1 & 0xFFFFu]; acc1 += Kc[i1 >> 16];
        }
        if (em != 0xFFFFFFFFu) {
          const double v = acc0 + acc1;
          S.vout[em & 0xFFFFu] = v;
          if ((em >> 16) != TG_NONE16) S.vout[em >> 16] = v;
        }
      }
    }
    pipe_phase_a<NPC>(d1, S.cx[nbuf], S.Kc[nbuf], ln, prm);
    {
      const TileDesc& d2 = S.desc[(j + 2) & (PG_RING - 1)];
      load_lconn(d2);
      load_coords(d2);   // addresses: fidx(j+2), loaded in Y(j-1)
      load_units(d1);
      load_rows(d1);     // addresses: node(j+1), loaded in Y(j-1)
    }
    __syncthreads();
    // ================= Y(j): C(j), staging for the tiles ahead =================
    if (tid == 0) pipe_tma_tile(S.desc[(j + 2) & (PG_RING - 1)], A, S, buf);
    // diagonal = -(sum of the row's off-diagonals), four lanes per row, stored straight to global
    for (int rbase = 0; rbase < d.nb_row; rbase += PG_THREADS / 4) {
      const int i = rbase + (tid >> 2), q = tid & 3;
      double sum = 0.0;
      int ed = -1;
      if (i < d.nb_row) {
        const uint32_t ri = S.rowinfo[i];
        if (rowinfo_own(ri)) {
          const int e0 = rowinfo_erow(ri), e1 = rowinfo_erow(S.rowinfo[i + 1]);
          ed = e0 + rowinfo_pdiag(ri);
          for (int e = e0 + q; e < e1; e += 4)
            if (e != ed) sum += S.vout[e];
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      if (q == 0 && ed >= 0) {
        double* dst = A.values + ((int64_t)S.rowbeg[i] + ed);
        if (A.accumulate) *dst -= sum; else *dst = -sum;
      }
    }
    // the other entries leave shared memory in row order: contiguous, coalesced stores
    for (int e = tid; e < d.nb_entry; e += PG_THREADS) {
      int r = S.etab[e >> 3];
      uint32_t ri = S.rowinfo[r + 1];
      while (e >= rowinfo_erow(ri)) {
        ++r;
        ri = S.rowinfo[r + 1];
      }
      ri = S.rowinfo[r];
      const bool own = rowinfo_own(ri);
      if (own && e == rowinfo_erow(ri) + rowinfo_pdiag(ri)) continue;
      const double v = own ? S.vout[e] : 0.0;
      double* dst = A.values + ((int64_t)S.rowbeg[r] + e);
      if (A.accumulate) *dst += v; else *dst = v;
    }
    {
      const TileDesc& d2 = S.desc[(j + 2) & (PG_RING - 1)];
      stage_coords(d2, buf);   // cx[buf] was last read by A(j) in X(j-1)
      stage_units(d1);         // unit table of B(j+1); B(j) is done
      if (tid < DW) reinterpret_cast<int32_t*>(&S.desc[(j + 4) & (PG_RING - 1)])[tid] = desc_word;
      load_fidx(S.desc[(j + 3) & (PG_RING - 1)]);
      load_node(d2);
    }
    __syncthreads();
  }
}

int assemble_tiled_pipe(afb_ctx* ctx, const ElemParams& prm, int accumulate)
{
  const TilePlan& P = ctx->plan;
  PipeArgs A;
  A.desc = P.tile_desc.as<TileDesc>();
  A.nb_tile = P.nb_tile;
  A.coords = ctx->coords.as<double>();
  A.foot = P.foot.as<int32_t>();
  A.lconn = P.lconn.as<ushort4>();
  A.tile_nodes = P.tile_nodes.as<int32_t>();
  A.rowinfo = P.rowinfo.as<uint32_t>();
  A.rows = ctx->rows.as<int32_t>();
  A.unit_base = P.unit_base.as<uint32_t>();
  A.unit_len = P.unit_len.as<uint16_t>();
  A.emap = P.emap.as<uint32_t>();
  A.lists = P.lists.as<uint16_t>();
  A.values = ctx->values.as<double>();
  A.accumulate = accumulate;
  const int grid = std::min<int>(P.nb_tile, ctx->sm_count);
  const size_t smem = sizeof(PipeSmem);
  if (ctx->npc == 4) {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled_pipe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled_pipe<4><<<grid, PG_THREADS, smem, ctx->stream>>>(A, prm);
  }
  else {
    AFB_CUDA(cudaFuncSetAttribute(k_assemble_tiled_pipe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_assemble_tiled_pipe<3><<<grid, PG_THREADS, smem, ctx->stream>>>(A, prm);
  }
  AFB_LAUNCH_CHECK(ctx);
  return AFB_OK;
}

} // namespace afb
