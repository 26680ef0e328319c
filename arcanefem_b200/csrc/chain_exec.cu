// Scalar tiled-gather executor on chained slices (plan: chain_plan.cu, layout: chain.cuh).
//
// Reference behaviour replaced: _assembleNodeWiseCsrBilinearOperator{Tria3,Tetra4}
// (modules/testlab/NodeWiseCsrBiliAssembly.cc:157-297) and BSRFormat::assembleBilinearAtomicFree
// (femutils/BSRFormat.h:406-577) for b = 1: every matrix row is written exactly once, by one owner,
// without atomics and without a zero fill.
//
// One persistent CTA walks its segments slice by slice:
//   stage    the slice's plan record (contribution lists, entry map, unit and row tables) arrives through the
//            TMA engine (cp.async.bulk + mbarrier) while phase A computes; the next record is pulled into L2;
//            footprint coordinates, local connectivity and row offsets of the next slice travel in registers
//   phase A  one thread per cell the slice computes: geometry once (one determinant, one reciprocal), the 6 (Tet4)
//            / 3 (Tri3) off-diagonal K_e values go to the slice's cache region; the cells shared with the
//            previous slice are already in the other region
//   phase B  one lane per computed entry: 4 cache indices per 64-bit list word, summed in a fixed order; the value
//            goes to the staging row, to its symmetric twin, or to the next slice's staging buffer
//   phase C  4 lanes per row: the row leaves shared memory for HBM as it is summed, the diagonal is minus the sum
//            of the off-diagonals (zero row sums of the stiffness matrix)
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "chain.cuh"
#include "element.cuh"
#include "tiles.cuh"

namespace afb {

struct ChainArgs {
  const SliceDesc* desc;
  const int32_t* order;
  const int32_t* cta_ptr;
  const double* coords;
  const int32_t* foot;
  const uint2* lconn;
  const int32_t* slice_nodes;
  const int32_t* rows;
  const unsigned char* blob;
  double* values;
  int accumulate;
  int stage_max;   // records larger than this are read from global memory (test knob; <= G::BLOB)
};

template <class G, int NPC>
struct ChainSmem {
  static constexpr int NPAIR = NPC * (NPC - 1) / 2;
  static constexpr int ZERO = NPAIR * G::PLANE;
  double Kc[ZERO + 1];
  double cx[3 * G::FMAX];
  double vout[2][G::EMAX];
  __align__(16) unsigned char blob[G::BLOB];
  int32_t rowbeg[G::RMAX];
  __align__(16) SliceDesc desc[4];
  __align__(8) unsigned long long mbar;
};

template <int ROUNDS>
struct ChainPrefetch {
  double c0, c1, c2;
  uint2 ln[ROUNDS];
  int32_t rowbeg;
  int32_t fidx, node;
};

template <int NPC>
__device__ __forceinline__ void chain_cell(const double* __restrict__ cx, uint2 ln, const ElemParams& prm, double (&K)[6])
{
  if constexpr (NPC == 4) {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    const double* p3 = cx + 3 * (ln.y >> 16);
    Tet4Geom g;
    g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(0, 3) * g.s;
    K[3] = g.dot(1, 2) * g.s; K[4] = g.dot(1, 3) * g.s; K[5] = g.dot(2, 3) * g.s;
  }
  else {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    Tri3Geom g;
    g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], (prm.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(1, 2) * g.s;
    K[3] = K[4] = K[5] = 0.0;
  }
}

__device__ __forceinline__ void ch_mbar_wait(uint32_t mbar, unsigned parity)
{
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}

template <class G, int NPC>
__global__ void __launch_bounds__(G::THREADS, G::MINB) k_assemble_chain(ChainArgs A, ElemParams prm)
{
  static_assert(G::NREG == 2 && G::FMAX <= G::THREADS && G::RMAX <= G::THREADS, "phase-separated executor: two regions, one thread per footprint node / row");
  using SM = ChainSmem<G, NPC>;
  constexpr int T = G::THREADS, NW = T / 32, ROUNDS = G::ROUNDS, NPAIR = SM::NPAIR;
  constexpr int DW = sizeof(SliceDesc) / 4;
  extern __shared__ __align__(16) unsigned char ch_raw[];
  SM& S = *reinterpret_cast<SM*>(ch_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t it0 = __ldg(A.cta_ptr + blockIdx.x), it1 = __ldg(A.cta_ptr + blockIdx.x + 1);
  if (it0 >= it1) return;
  const uint32_t mbar = smem_u32(&S.mbar);
  if (tid == 0) {
    S.Kc[SM::ZERO] = 0.0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 3 * DW) { // descriptors of the first three slices
    const int k = tid / DW, w = tid % DW;
    if (it0 + k < it1) {
      const int32_t s = __ldg(A.order + it0 + k);
      reinterpret_cast<int32_t*>(&S.desc[k])[w] = __ldg(reinterpret_cast<const int32_t*>(A.desc + s) + w);
    }
  }
  __syncthreads();
  ChainPrefetch<ROUNDS> pf;
  auto level1 = [&](const SliceDesc& d) {
    if (tid < d.nb_foot) pf.fidx = __ldg(A.foot + d.foot_off + tid);
    if (tid < d.nb_row) pf.node = __ldg(A.slice_nodes + d.node_off + tid);
  };
  auto level2 = [&](const SliceDesc& d) {
    if (tid < d.nb_foot) {
      const double* p = A.coords + 3 * (int64_t)pf.fidx;
      pf.c0 = __ldg(p);
      pf.c1 = __ldg(p + 1);
      pf.c2 = __ldg(p + 2);
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const int lc = min(r * T + tid, d.nb_new - 1);
      if (d.nb_new > 0) pf.ln[r] = __ldg(A.lconn + d.cell_off + lc);
    }
    if (tid < d.nb_row) pf.rowbeg = __ldg(A.rows + pf.node);
  };
  auto stage_coords = [&](const SliceDesc& d) {
    if (tid < d.nb_foot) {
      S.cx[3 * tid] = pf.c0;
      S.cx[3 * tid + 1] = pf.c1;
      S.cx[3 * tid + 2] = pf.c2;
    }
  };
  level1(S.desc[0]);
  level2(S.desc[0]); // the only exposed dependent load of the kernel
  if (it0 + 1 < it1) level1(S.desc[1]);
  stage_coords(S.desc[0]);
  unsigned parity = 0;
  int slot = 0;
  for (int32_t it = it0; it < it1; ++it) {
    // ---- barrier 0: coordinates are staged (tail of the previous iteration); every warp is done with the previous
    //      slice's phase C, so the record buffer, the row offsets and the staging row of two slices ago are free ----
    __syncthreads();
    const SliceDesc d = S.desc[slot];
    const bool staged = d.blob_bytes <= A.stage_max;
    if (tid < d.nb_row) S.rowbeg[tid] = pf.rowbeg;
    if (tid == 0) {
      if (staged && d.blob_bytes > 0) {
        const uint32_t bytes = (uint32_t)d.blob_bytes;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.blob)), "l"(A.blob + (size_t)d.blob_off * 16),
                     "r"(bytes), "r"(mbar)
                     : "memory");
      }
      if (it + 1 < it1) { // the next slice's record: into L2 now, into shared memory when its turn comes
        const SliceDesc& dn = S.desc[(slot + 1) & 3];
        if (dn.blob_bytes > 0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.blob + (size_t)dn.blob_off * 16), "r"((uint32_t)dn.blob_bytes) : "memory");
      }
    }
    // ---- phase A ----
    const int par = (d.flags & CH_FLAG_PARITY) ? 1 : 0;
    {
      const int base_own = par * G::CS, base_other = (1 - par) * G::CS - d.nb_a;
#pragma unroll
      for (int r = 0; r < ROUNDS; ++r) {
        const int lc = r * T + tid;
        if (lc < d.nb_new) {
          double K[6];
          chain_cell<NPC>(S.cx, pf.ln[r], prm, K);
          const int pos = lc < d.nb_a ? base_own + lc : base_other + lc;
#pragma unroll
          for (int p = 0; p < NPAIR; ++p) S.Kc[p * G::PLANE + pos] = K[p];
        }
      }
      // a segment's first slice computes up to two regions' worth of cells: the rounds beyond the prefetched ones
      for (int lc = ROUNDS * T + tid; lc < d.nb_new; lc += T) {
        const uint2 ln = __ldg(A.lconn + d.cell_off + lc);
        double K[6];
        chain_cell<NPC>(S.cx, ln, prm, K);
        const int pos = lc < d.nb_a ? base_own + lc : base_other + lc;
#pragma unroll
        for (int p = 0; p < NPAIR; ++p) S.Kc[p * G::PLANE + pos] = K[p];
      }
    }
    // software pipeline: data of the next slice (addresses already in registers), indices of the one after it,
    // descriptor of the one after that -- all in flight during phases B and C
    const int nslot = (slot + 1) & 3, nnslot = (slot + 2) & 3, nnnslot = (slot + 3) & 3;
    if (it + 1 < it1) level2(S.desc[nslot]);
    if (it + 2 < it1) level1(S.desc[nnslot]);
    int32_t desc_word = 0;
    if (tid < DW && it + 3 < it1) {
      const int32_t s = __ldg(A.order + it + 3);
      desc_word = __ldg(reinterpret_cast<const int32_t*>(A.desc + s) + tid);
    }
    __syncthreads(); // ---- barrier 1: the element cache is complete ----
    if (staged && d.blob_bytes > 0) {
      ch_mbar_wait(mbar, parity);
      parity ^= 1u;
    }
    double* vcur = S.vout[par];
    double* vnext = S.vout[1 - par];
    auto phase_b = [&](const unsigned char* rec) {
      const uint2* lists = reinterpret_cast<const uint2*>(rec);
      const uint32_t* emap = reinterpret_cast<const uint32_t*>(rec + ch_off_emap(d.nb_chunk));
      const uint32_t* units = reinterpret_cast<const uint32_t*>(rec + ch_off_units(d.nb_chunk, d.nb_unit));
#pragma unroll 1
      for (int u = warp; u < d.nb_unit; u += NW) {
        const uint32_t uw = units[u];
        const int nch = (int)(uw & 0xFFu);
        const uint2* l = lists + (size_t)(uw >> 8) * 32 + lane;
        const uint32_t em = emap[u * 32 + lane];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = 0;
#pragma unroll 1
        for (; k + 1 < nch; k += 2) { // two list words = 8 cache gathers in flight
          const uint2 w0 = l[k * 32], w1 = l[(k + 1) * 32];
          const double x0 = S.Kc[w0.x & 0xFFFFu], x1 = S.Kc[w0.x >> 16], x2 = S.Kc[w0.y & 0xFFFFu], x3 = S.Kc[w0.y >> 16];
          const double y0 = S.Kc[w1.x & 0xFFFFu], y1 = S.Kc[w1.x >> 16], y2 = S.Kc[w1.y & 0xFFFFu], y3 = S.Kc[w1.y >> 16];
          a0 += x0; a1 += x1; a2 += x2; a3 += x3;
          a0 += y0; a1 += y1; a2 += y2; a3 += y3;
        }
        if (k < nch) {
          const uint2 w0 = l[k * 32];
          a0 += S.Kc[w0.x & 0xFFFFu]; a1 += S.Kc[w0.x >> 16]; a2 += S.Kc[w0.y & 0xFFFFu]; a3 += S.Kc[w0.y >> 16];
        }
        if (em != 0xFFFFFFFFu) {
          const double v = (a0 + a1) + (a2 + a3);
          vcur[em & 0xFFFFu] = v;
          const uint32_t hi = em >> 16;
          if (hi != CH_NONE16) {
            if (hi & CH_NEXT) vnext[hi & 0x7FFFu] = v;
            else vcur[hi] = v;
          }
        }
      }
    };
    // ---- phase B ----
    if (staged) phase_b(S.blob);
    else phase_b(A.blob + (size_t)d.blob_off * 16);
    __syncthreads(); // ---- barrier 2: the staging row is complete ----
    // ---- phase C: rows leave for HBM, 4 lanes per row ----
    auto phase_c = [&](const unsigned char* rec) {
      const uint32_t* rowinfo = reinterpret_cast<const uint32_t*>(rec + ch_off_rowinfo(d.nb_chunk, d.nb_unit));
      const int q = lane & 3;
#pragma unroll 1
      for (int ib = warp * 8; ib < d.nb_row; ib += NW * 8) {
        const int i = ib + (lane >> 2);
        double sum = 0.0;
        double* dst = nullptr;
        int ed = -1;
        bool own = false;
        if (i < d.nb_row) {
          const uint32_t ri = rowinfo[i];
          const int e0 = rowinfo_erow(ri), e1 = rowinfo_erow(rowinfo[i + 1]);
          own = rowinfo_own(ri);
          ed = e0 + rowinfo_pdiag(ri);
          dst = A.values + ((int64_t)S.rowbeg[i] - e0);
          if (own) {
            for (int e = e0 + q; e < e1; e += 4) {
              if (e != ed) {
                const double v = vcur[e];
                sum += v;
                if (A.accumulate) dst[e] += v; else dst[e] = v;
              }
            }
          }
          else if (!A.accumulate) { // rows of non-owned nodes stay zero (the isOwn gate of the reference)
            for (int e = e0 + q; e < e1; e += 4) dst[e] = 0.0;
          }
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        if (own && q == 0) {
          if (A.accumulate) dst[ed] -= sum; else dst[ed] = -sum;
        }
      }
    };
    if (staged) phase_c(S.blob);
    else phase_c(A.blob + (size_t)d.blob_off * 16);
    // ---- tail: what the next slice's phase A reads (other warps may still be in phase C: neither touches cx) ----
    if (tid < DW && it + 3 < it1) reinterpret_cast<int32_t*>(&S.desc[nnnslot])[tid] = desc_word;
    if (it + 1 < it1) stage_coords(S.desc[nslot]);
    slot = nslot;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// geometry of the phase-separated executor: B (2 CTAs/SM x 384 threads) unless AFB_CHAIN_GEOM=A (3 CTAs/SM x 256 threads)
static int chain_geometry_choice()
{
  static const int g = [] {
    const char* e = getenv("AFB_CHAIN_GEOM");
    return (e && (e[0] == 'A' || e[0] == 'a')) ? 0 : 1;
  }();
  return g;
}


float chain_plan_ms(const afb_ctx* ctx)
{
  const ChainPlan* P = static_cast<const ChainPlan*>(ctx->chain);
  return (P && P->valid && P->mesh_gen == ctx->mesh_gen) ? P->plan_ms : -1.0f;
}

template <class G>
static int chain_run(afb_ctx* ctx, const ElemParams& prm, int mode, int geom, int accumulate)
{
  const int grid_full = G::MINB * ctx->sm_count;
  if (!chain_plan_valid(ctx, mode, geom)) AFB_TRY(chain_build(ctx, mode, geom, chain_limits<G>(), grid_full));
  ChainPlan& P = *static_cast<ChainPlan*>(ctx->chain);
  // values already holding contributions (a second operator added on top) are accumulated into; a fresh matrix is
  // simply overwritten: every entry of every row is written exactly once, no zero fill needed
  if (accumulate) AFB_TRY(ensure_values_zeroed(ctx));
  else ctx->values_dirty = false;
  if (P.nb_slice == 0) return AFB_OK;
  ChainArgs A;
  A.desc = P.desc.as<SliceDesc>();
  A.order = P.order.as<int32_t>();
  A.cta_ptr = P.cta_ptr.as<int32_t>();
  A.coords = ctx->coords.as<double>();
  A.foot = P.foot.as<int32_t>();
  A.lconn = P.lconn.as<uint2>();
  A.slice_nodes = P.slice_nodes.as<int32_t>();
  A.rows = ctx->rows.as<int32_t>();
  A.blob = P.blob.as<unsigned char>();
  A.values = ctx->values.as<double>();
  A.accumulate = accumulate;
  A.stage_max = (int)std::min<int64_t>(G::BLOB, ctx->tiled_stage_limit);
  auto go = [&](auto kernel, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<P.grid, G::THREADS, smem, ctx->stream>>>(A, prm);
    return cudaGetLastError();
  };
  cudaError_t e = ctx->npc == 4 ? go(k_assemble_chain<G, 4>, sizeof(ChainSmem<G, 4>)) : go(k_assemble_chain<G, 3>, sizeof(ChainSmem<G, 3>));
  AFB_CUDA(e);
  ctx->launches++;
  return AFB_OK;
}

int chain_assemble(afb_ctx* ctx, const ElemParams& prm, int flags, int accumulate)
{
  const int mode = flags & (AFB_FLAG_ALL_ROWS | AFB_FLAG_OWN_CELLS_ONLY);
  if (ctx->tiled_exec == AFB_TILED_EXEC_CHAIN_FLOW) return flow_assemble(ctx, prm, flags, accumulate);
  const int geom = chain_geometry_choice();
  static_assert(sizeof(ChainSmem<ChainGeomA, 4>) <= 233472 / ChainGeomA::MINB - 1024, "geometry A: MINB CTAs must fit one SM");
  static_assert(sizeof(ChainSmem<ChainGeomB, 4>) <= 233472 / ChainGeomB::MINB - 1024, "geometry B: MINB CTAs must fit one SM");
  return geom == 0 ? chain_run<ChainGeomA>(ctx, prm, mode, geom, accumulate) : chain_run<ChainGeomB>(ctx, prm, mode, geom, accumulate);
}

} // namespace afb
