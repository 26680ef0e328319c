// Tiled gather assembly: the B200 path of the atomic-free ("node-wise") back-ends.
//
// Reference behaviour replaced: _assembleNodeWiseCsrBilinearOperator{Tria3,Tetra4}
// (modules/testlab/NodeWiseCsrBiliAssembly.cc:157-297) and BSRFormat::assembleBilinearAtomicFree
// (femutils/BSRFormat.h:406-577): every matrix row is written by exactly one owner, without
// atomics.  The reference does it with one thread per node that recomputes the geometry of
// every incident cell (4x redundant fp64 work on tetrahedra, valence-divergent).  Here one CTA
// finishes one tile of rows (inspector: tiles_plan.cu), two CTAs resident per SM:
//
//   stage    footprint coordinates, row offsets and unit tables go to shared memory; the tile's
//            contribution lists arrive asynchronously through the TMA engine (cp.async.bulk)
//   phase A  one thread per tile cell: geometry once (one determinant, one reciprocal); scalar
//            operators cache the 6 (Tet4) / 3 (Tri3) off-diagonal K_e values, vector operators
//            cache sqrt(s)*grad(phi_a) (the blocks are rank-one sums of those)
//   phase B  one lane per matrix entry: sum the cached contributions in a fixed order (ascending
//            cell id => bit-reproducible); symmetric twins inside the tile are summed once
//   phase C  scalar: the diagonal is minus the sum of the row's off-diagonals (zero row sums of
//            the stiffness matrix), rows leave shared memory as contiguous, coalesced stores;
//            vector: blocks are written from registers in either BSR value layout
//   Rows are written exactly once, so a fresh assembly needs no zero fill of `values`.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "chain.cuh"
#include "element.cuh"
#include "tiles.cuh"

namespace afb {

// ---------------------------------------------------------------------------------------------
// element caches
// ---------------------------------------------------------------------------------------------
template <int NPC> struct OffDiagK;
template <> struct OffDiagK<4> {
  static constexpr int N = 6;
  template <bool COEF = false> __device__ static __forceinline__ void compute(const double* __restrict__ cx, uint2 ln, const ElemParams&, double (&K)[6], double coef = 1.0)
  {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    const double* p3 = cx + 3 * (ln.y >> 16);
    Tet4Geom g;
    g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    if (COEF) g.s *= coef; // as the cell-wise kernel scales it (element.cuh: g.s *= p.scale)
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(0, 3) * g.s;
    K[3] = g.dot(1, 2) * g.s; K[4] = g.dot(1, 3) * g.s; K[5] = g.dot(2, 3) * g.s;
  }
};
template <> struct OffDiagK<3> {
  static constexpr int N = 3;
  template <bool COEF = false> __device__ static __forceinline__ void compute(const double* __restrict__ cx, uint2 ln, const ElemParams& prm, double (&K)[6], double coef = 1.0)
  {
    const double* p0 = cx + 3 * (ln.x & 0xFFFFu);
    const double* p1 = cx + 3 * (ln.x >> 16);
    const double* p2 = cx + 3 * (ln.y & 0xFFFFu);
    Tri3Geom g;
    g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], (prm.flags & AFB_FLAG_SIGNED_TRI_AREA) != 0);
    if (COEF) g.s *= coef;
    K[0] = g.dot(0, 1) * g.s; K[1] = g.dot(0, 2) * g.s; K[2] = g.dot(1, 2) * g.s;
    K[3] = K[4] = K[5] = 0.0;
  }
};

struct ExecSmem {
  double Kc[TG_ZERO + 1];
  double cx[3 * TG_FMAX];
  double vout[TG_EMAX];
  __align__(16) uint16_t lists[TG_LMAX]; // phase B input; afterwards (phase C, write-out) reused as OutTables
  int32_t rowbeg[TG_RMAX];        // first value of the row minus its first tile-local entry: dest(e) = e + rowbeg[row(e)]
  uint32_t rowinfo[TG_RMAX + 1];  // + sentinel (first entry = nb_entry)
  uint32_t ubase[TG_UMAX];
  uint16_t ulen[TG_UMAX];
  __align__(16) TileDesc desc[4]; // ring: current tile of this CTA and the three after it
  __align__(8) unsigned long long mbar;
};
// write-out table, built by phase C over the (by then consumed) contribution lists; the tables staged per
// tile (rowbeg, rowinfo, ...) are not read by the write-out, so staging the next tile does not wait for it
struct OutTables {
  int32_t dbase[TG_EMAX]; // value offset of the entry's row minus the row's first tile-local entry: dest(e) = e + dbase[e]
};
static_assert(sizeof(OutTables) <= sizeof(uint16_t) * TG_LMAX, "write-out tables alias the list region");
static_assert(sizeof(ExecSmem) <= TG_SMEM_LIMIT, "TG_MINB executor CTAs must fit one SM (228 KB, 1 KB reserved per CTA)");

constexpr int TV_ROUNDS = (TV_CMAX + TG_THREADS - 1) / TG_THREADS;
constexpr int TG_PF_ROUNDS = TG_ROUNDS > TV_ROUNDS ? TG_ROUNDS : TV_ROUNDS;
constexpr int TG_UPW = (TG_UMAX + TG_THREADS / 32 - 1) / (TG_THREADS / 32); // units per warp

// inputs of the next tile a thread carries in registers across phases B and C
struct TilePrefetch {
  double c0, c1, c2;      // coordinates of footprint node `threadIdx.x`
  uint2 ln[TG_PF_ROUNDS];   // local connectivity of this thread's cells (4 x 16 bit, unpacked at use: a
                            // predicated ushort4 load makes ptxas merge halves right after the load = a stall)
  uint32_t ubase;         // unit table entry `threadIdx.x`
  uint16_t ulen;
  int32_t rowbeg;         // row `threadIdx.x`: first value of the row, plan word
  uint32_t rowinfo;
  int32_t fidx, node;     // level-1 indices (footprint node, row node) of the tile after that
  uint32_t em[TG_UPW];    // entry map words of this warp's units (scalar executor)
  uint2 unit;             // unit record `threadIdx.x` (row-ordered vector executor)
  int32_t cid[TG_PF_ROUNDS]; // global ids of this thread's cells (scalar executor with a per-cell coefficient)
};

struct ExecArgs {
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
  union {                      // (one slot: a larger argument block costs the plain scalar executor two registers)
    const uint32_t* emap_rows; // vector plans only
    const int32_t* tile_cells; // scalar executor with a per-cell coefficient (ElemParams::cell_coef): global ids of the tiles' cells
  };
  const uint16_t* lists;
  const uint2* vr_units;     // row-ordered vector plans only
  double* values;
  int accumulate;
  int list_stage_max; // tiles with more 16-bit list slots read their lists from global memory (<= the staging buffer)
};

// The next tiles' inputs travel in two waves so that no warp ever waits on a dependent load:
// level 1 (indices: footprint node ids, row node ids) is requested TWO tiles ahead, level 2 (the data
// those indices address, plus everything addressed directly) one tile ahead, by which time its
// addresses sit in registers.
__device__ __forceinline__ void prefetch_level1(const TileDesc& d, const ExecArgs& A, TilePrefetch& pf)
{
  if ((int)threadIdx.x < d.nb_foot) pf.fidx = __ldg(A.foot + d.foot_off + threadIdx.x);
  if ((int)threadIdx.x < d.nb_row) pf.node = __ldg(A.tile_nodes + d.node_off + threadIdx.x);
}

template <int ROUNDS, int THREADS, bool EMAP = false, bool COEF = false>
__device__ __forceinline__ void prefetch_level2(const TileDesc& d, const ExecArgs& A, TilePrefetch& pf)
{
  if constexpr (EMAP) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < TG_UPW; ++q) {
      const int u = warp + q * (THREADS / 32);
      if (u < d.nb_unit) pf.em[q] = __ldg(A.emap + (size_t)(d.unit_off + u) * 32 + lane);
    }
  }
  if ((int)threadIdx.x < d.nb_foot) {
    const double* p = A.coords + 3 * (int64_t)pf.fidx;
    pf.c0 = __ldg(p);
    pf.c1 = __ldg(p + 1);
    pf.c2 = __ldg(p + 2);
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int lc = min(r * THREADS + (int)threadIdx.x, d.nb_cell - 1);
    if (d.nb_cell > 0) pf.ln[r] = __ldg(reinterpret_cast<const uint2*>(A.lconn) + d.cell_off + lc);
    if constexpr (COEF)
      if (d.nb_cell > 0) pf.cid[r] = __ldg(A.tile_cells + d.cell_off + lc);
  }
  if ((int)threadIdx.x < d.nb_unit) {
    pf.ubase = __ldg(A.unit_base + d.unit_off + threadIdx.x);
    pf.ulen = __ldg(A.unit_len + d.unit_off + threadIdx.x);
  }
  if ((int)threadIdx.x < d.nb_row) {
    pf.rowbeg = __ldg(A.rows + pf.node);
    pf.rowinfo = __ldg(A.rowinfo + d.node_off + threadIdx.x);
  }
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, unsigned parity)
{
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// scalar executor (b = 1, zero-row-sum operators: Poisson)
// ---------------------------------------------------------------------------------------------
// registers -> shared memory.  stage_early: what phases C/write-out of the previous tile do not read (coordinates:
// phase A only; unit tables: phase B only) -- done in the tail of the previous iteration, while other warps may
// still be writing out.  stage_rows: the row tables phase C reads -- done after the top barrier.
template <class SM>
__device__ __forceinline__ void stage_early(SM& S, const TileDesc& d, const TilePrefetch& pf)
{
  if ((int)threadIdx.x < d.nb_foot) {
    S.cx[3 * threadIdx.x] = pf.c0;
    S.cx[3 * threadIdx.x + 1] = pf.c1;
    S.cx[3 * threadIdx.x + 2] = pf.c2;
  }
  if ((int)threadIdx.x < d.nb_unit) {
    S.ubase[threadIdx.x] = pf.ubase;
    S.ulen[threadIdx.x] = pf.ulen;
  }
}
template <class SM>
__device__ __forceinline__ void stage_rows(SM& S, const TileDesc& d, const TilePrefetch& pf)
{
  if ((int)threadIdx.x < d.nb_row) {
    S.rowbeg[threadIdx.x] = pf.rowbeg - rowinfo_erow(pf.rowinfo);
    S.rowinfo[threadIdx.x] = pf.rowinfo;
  }
  else if ((int)threadIdx.x == d.nb_row) S.rowinfo[threadIdx.x] = pack_rowinfo(d.nb_entry, 0, false);
}

template <int NPC, bool COEF>
__device__ __forceinline__ void assemble_tiled_body(const ExecArgs& A, const ElemParams& prm)
{
  extern __shared__ __align__(16) unsigned char ex_raw[];
  ExecSmem& S = *reinterpret_cast<ExecSmem*>(ex_raw);
  pdl_enter();
  OutTables& O = *reinterpret_cast<OutTables*>(S.lists);
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
  if (threadIdx.x < 3 * DW) { // descriptors of the first three tiles
    const int k = threadIdx.x / DW, w = threadIdx.x % DW;
    const int64_t tt = (int64_t)t + (int64_t)k * gridDim.x;
    if (tt < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[k])[w] = __ldg(reinterpret_cast<const int32_t*>(A.desc + tt) + w);
  }
  __syncthreads();
  unsigned parity = 0;
  int slot = 0;
  TilePrefetch pf;
#ifdef AFB_EXP_STAGGER
  if (blockIdx.x >= gridDim.x / 2) { // experiment: start the second CTA of every SM half a tile later
    const long long t0 = clock64();
    while (clock64() - t0 < AFB_EXP_STAGGER) {}
  }
#endif
  if (t < A.nb_tile) {
    prefetch_level1(S.desc[0], A, pf);
    prefetch_level2<TG_ROUNDS, TG_THREADS, true, COEF>(S.desc[0], A, pf); // the only exposed dependent load of the kernel
    if ((int64_t)t + gridDim.x < A.nb_tile) prefetch_level1(S.desc[1], A, pf);
    stage_early(S, S.desc[0], pf);
  }
  while (t < A.nb_tile) {
    // ---- barrier 1: coordinates and unit tables are staged (tail of the previous iteration); every warp is done
    //      with the previous tile's phase C and write-out, so the row tables and the list region may be overwritten ----
    __syncthreads();
    const TileDesc d = S.desc[slot];
    const bool staged = d.list_len <= A.list_stage_max; // lists of an oversized tile are read from global memory
    stage_rows(S, d, pf);
    if (threadIdx.x == 0 && staged && d.list_len > 0) {
      const uint32_t bytes = (uint32_t)d.list_len * 2u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.lists)), "l"(A.lists + d.list_off),
                   "r"(bytes), "r"(mbar)
                   : "memory");
    }
    // ---- phase A: off-diagonal element-matrix values of the tile's cells, once each ----
    double cf[COEF ? TG_ROUNDS : 1]; // (the cells' ids came with the level-2 wave: one exposed load per tile, all rounds in flight together)
    if constexpr (COEF) {
#pragma unroll
      for (int r = 0; r < TG_ROUNDS; ++r) cf[r] = r * TG_THREADS + (int)threadIdx.x < d.nb_cell ? __ldg(prm.cell_coef + pf.cid[r]) : 1.0;
    }
#pragma unroll
    for (int r = 0; r < TG_ROUNDS; ++r) {
      const int lc = r * TG_THREADS + threadIdx.x;
#ifdef AFB_EXP_A_PCT
      if (lc < d.nb_cell * AFB_EXP_A_PCT / 100) {
#else
      if (lc < d.nb_cell) {
#endif
        double K[6];
        if (COEF) OffDiagK<NPC>::template compute<true>(S.cx, pf.ln[r], prm, K, cf[r]);
        else OffDiagK<NPC>::compute(S.cx, pf.ln[r], prm, K);
#pragma unroll
        for (int p = 0; p < OffDiagK<NPC>::N; ++p) S.Kc[p * TG_CS + lc] = K[p];
      }
    }
    // software pipeline: level-2 inputs of the next tile (addresses already in registers), level-1
    // indices of the tile after it, descriptor of the tile after that -- all in flight during B and C
    const int64_t tn = (int64_t)t + gridDim.x, tnn = tn + gridDim.x, tnnn = tnn + gridDim.x;
    const int nslot = (slot + 1) & 3, nnslot = (slot + 2) & 3, nnnslot = (slot + 3) & 3;
    // (the entry-map words of the current tile move to their own registers first)
    uint32_t em[TG_UPW];
#pragma unroll
    for (int q = 0; q < TG_UPW; ++q) em[q] = pf.em[q];
    if (tn < A.nb_tile) prefetch_level2<TG_ROUNDS, TG_THREADS, true, COEF>(S.desc[nslot], A, pf);
    if (tnn < A.nb_tile) prefetch_level1(S.desc[nnslot], A, pf);
    int32_t desc_word = 0;
    if (threadIdx.x < DW && tnnn < A.nb_tile) desc_word = __ldg(reinterpret_cast<const int32_t*>(A.desc + tnnn) + threadIdx.x);
    __syncthreads(); // ---- barrier 2: the element cache is complete ----
    if (staged && d.list_len > 0) {
      mbar_wait(mbar, parity);
      parity ^= 1u;
    }
    // ---- phase B: one warp per unit of 32 entries with equally long contribution lists ----
    {
      const uint32_t* l32 = staged ? reinterpret_cast<const uint32_t*>(S.lists) : reinterpret_cast<const uint32_t*>(A.lists + d.list_off);
      constexpr uint32_t ZPAIR = (uint32_t)TG_ZERO | ((uint32_t)TG_ZERO << 16);
#pragma unroll
      for (int q = 0; q < TG_UPW; ++q) {
        const int u = warp + q * NW;
#ifdef AFB_EXP_B_PCT
        if (u < d.nb_unit * AFB_EXP_B_PCT / 100) {
#else
        if (u < d.nb_unit) {
#endif
          const uint32_t* l = l32 + (S.ubase[u] >> 1) + lane;
          const int len2 = S.ulen[u] >> 1;
          // four list words (8 contributions) per step: the list loads, then the cache gathers, are all in
          // flight together (the lane's sum is latency-bound otherwise); fixed association => reproducible
          double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
#pragma unroll 1
          for (int k = 0; k < len2; k += 4) {
            const uint32_t i0 = l[k * 32];
            const uint32_t i1 = (k + 1 < len2) ? l[(k + 1) * 32] : ZPAIR;
            const uint32_t i2 = (k + 2 < len2) ? l[(k + 2) * 32] : ZPAIR;
            const uint32_t i3 = (k + 3 < len2) ? l[(k + 3) * 32] : ZPAIR;
            const double a0 = S.Kc[i0 & 0xFFFFu], a1 = S.Kc[i0 >> 16], a2 = S.Kc[i1 & 0xFFFFu], a3 = S.Kc[i1 >> 16];
            const double a4 = S.Kc[i2 & 0xFFFFu], a5 = S.Kc[i2 >> 16], a6 = S.Kc[i3 & 0xFFFFu], a7 = S.Kc[i3 >> 16];
            acc0 += a0; acc1 += a1; acc2 += a2; acc3 += a3;
            acc0 += a4; acc1 += a5; acc2 += a6; acc3 += a7;
          }
          acc0 += acc2;
          acc1 += acc3;
          const uint32_t w = em[q];
          if (w != 0xFFFFFFFFu) {
            const double v = acc0 + acc1;
            S.vout[w & 0xFFFFu] = v;
            if ((w >> 16) != TG_NONE16) S.vout[w >> 16] = v;
          }
        }
      }
    }
    __syncthreads(); // ---- barrier 3: off-diagonals are in vout; the list region is free ----
    // ---- phase C + write-out, one contiguous range of rows per warp (no block barrier in between):
    //      G lanes per row derive the diagonal = -(sum of the row's off-diagonals) and note the row's value
    //      offset for each of its entries; then the warp's entries leave shared memory in row order
    //      (contiguous, coalesced stores) ----
    {
      const int rw = (d.nb_row + NW - 1) / NW;             // rows per warp (<= 32)
      const int r0 = min(warp * rw, d.nb_row), r1 = min(r0 + rw, d.nb_row);
      const int gs = rw <= 8 ? 2 : (rw <= 16 ? 1 : 0);     // log2(lanes per row)
      const int i = r0 + (lane >> gs), q = lane & ((1 << gs) - 1), G = 1 << gs;
      double sum = 0.0;
      int ed = -1;
      if (i < r1) {
        const uint32_t ri = S.rowinfo[i];
        const int e0 = rowinfo_erow(ri), e1 = rowinfo_erow(S.rowinfo[i + 1]);
        const int32_t rb = S.rowbeg[i];
        if (rowinfo_own(ri)) {
          ed = e0 + rowinfo_pdiag(ri);
          double s1 = 0.0;
          int e = e0 + q;
          for (; e + G < e1; e += 2 * G) { // two independent loads and sums per step (the diagonal slot holds stale data: select, never add)
            const double v0 = S.vout[e], v1 = S.vout[e + G];
            O.dbase[e] = rb;
            O.dbase[e + G] = rb;
            sum += e != ed ? v0 : 0.0;
            s1 += e + G != ed ? v1 : 0.0;
          }
          if (e < e1) {
            const double v0 = S.vout[e];
            O.dbase[e] = rb;
            sum += e != ed ? v0 : 0.0;
          }
          sum += s1;
        }
        else { // rows of non-owned nodes stay zero (the isOwn gate of the reference)
          for (int e = e0 + q; e < e1; e += G) {
            O.dbase[e] = rb;
            S.vout[e] = 0.0;
          }
        }
      }
      if (gs >= 1) sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      if (gs >= 2) sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      __syncwarp(); // the row's other lanes have passed their (discarded) read of the diagonal slot
      if (q == 0 && ed >= 0) S.vout[ed] = -sum;
      __syncwarp();
      const int eb = rowinfo_erow(S.rowinfo[r0]), ee = rowinfo_erow(S.rowinfo[r1]);
#ifdef AFB_EXP_OUT_PCT
      for (int e = eb + lane; e < eb + (ee - eb) * AFB_EXP_OUT_PCT / 100; e += 32) {
#else
      for (int e = eb + lane; e < ee; e += 32) {
#endif
        double* dst = A.values + ((int64_t)O.dbase[e] + e);
        const double v = S.vout[e];
        if (A.accumulate) *dst += v; else *dst = v;
      }
    }
    // ---- tail: stage what the next tile's phases A and B read (other warps may still be in phase C / write-out) ----
    if (threadIdx.x < DW && tnnn < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[nnnslot])[threadIdx.x] = desc_word;
    if (tn >= A.nb_tile) break;
    stage_early(S, S.desc[nslot], pf);
    t = (int32_t)tn;
    slot = nslot;
  }
}

template <int NPC>
__global__ void __launch_bounds__(TG_THREADS, TG_MINB) k_assemble_tiled(ExecArgs A, ElemParams prm)
{
  assemble_tiled_body<NPC, false>(A, prm);
}

// the same executor with a per-cell multiplier of the element matrix (afb_set_cell_coefficient: conductivity of the fourier /
// electrostatics / FourierNL modules); its own kernel so that the plain one keeps its registers and its instruction stream
template <int NPC>
__global__ void __launch_bounds__(TG_THREADS, TG_MINB) k_assemble_tiled_coef(ExecArgs A, ElemParams prm)
{
  assemble_tiled_body<NPC, true>(A, prm);
}

// ---------------------------------------------------------------------------------------------
// vector executor (b = DIM dofs per node: isotropic elasticity)
//   K_ab = s [ lambda c_a c_b^T + mu c_b c_a^T + mu (c_a.c_b) I ]  with g_a = sqrt(s) c_a:
//   sum over cells of K_ab = lambda M + mu M^T + mu tr(M) I,  M = sum g_a (x) g_b
// ---------------------------------------------------------------------------------------------
template <int DIM>
struct VecSmem {
  double G[TV_PLANES * TV_CS];
  double cx[3 * TG_FMAX];
  int32_t rowbeg[TG_RMAX];
  uint32_t rowinfo[TG_RMAX];
  uint32_t ubase[TG_UMAX];
  uint16_t ulen[TG_UMAX];
  __align__(16) uint16_t lists[TV_LMAX > 0 ? TV_LMAX : 8]; // the tile's contribution lists (TMA bulk copy, in flight during phase A)
  __align__(16) TileDesc desc[4];
  __align__(8) unsigned long long mbar;
};
static_assert(sizeof(VecSmem<3>) <= TG_SMEM_LIMIT, "TG_MINB vector-executor CTAs must fit one SM");



// OP 0: isotropic elasticity; OP 1 (Tri3 only): the bilaplacian's mixed form, dofs (u1,u2) per node
//   K_ab = [[0, s d_a.d_b], [s d_a.d_b, area (1 + delta_ab)]]   (modules/bilaplacian/ElementMatrix.h:37-45): the off-diagonal
//   term is tr(M), the mass-like term sums a seventh cached plane (the cells' areas)
template <int NPC, int LAYOUT, int OP = 0>
__global__ void __launch_bounds__(TG_THREADS, TG_MINB) k_assemble_tiled_vec(ExecArgs A, ElemParams prm)
{
  constexpr int DIM = NPC - 1, B = DIM;
  static_assert(OP == 0 || NPC == 3, "the bilaplacian executor is for Tri3");
  extern __shared__ __align__(16) unsigned char ex_raw[];
  VecSmem<DIM>& S = *reinterpret_cast<VecSmem<DIM>*>(ex_raw);
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TG_THREADS / 32;
  constexpr int DW = sizeof(TileDesc) / 4;
  int32_t t = blockIdx.x;
  const uint32_t mbar = smem_u32(&S.mbar);
  unsigned parity = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero slot of every plane (list padding)
  if (threadIdx.x < TV_PLANES) S.G[threadIdx.x * TV_CS + TV_CS - 1] = 0.0;
  if (threadIdx.x < 3 * DW) {
    const int k = threadIdx.x / DW, w = threadIdx.x % DW;
    const int64_t tt = (int64_t)t + (int64_t)k * gridDim.x;
    if (tt < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[k])[w] = __ldg(reinterpret_cast<const int32_t*>(A.desc + tt) + w);
  }
  __syncthreads();
  int slot = 0;
  TilePrefetch pf;
  if (t < A.nb_tile) {
    prefetch_level1(S.desc[0], A, pf);
    prefetch_level2<TV_ROUNDS, TG_THREADS>(S.desc[0], A, pf);
    if ((int64_t)t + gridDim.x < A.nb_tile) prefetch_level1(S.desc[1], A, pf);
  }
  const double lam = prm.p0, mu = prm.p1;
  while (t < A.nb_tile) {
    const TileDesc d = S.desc[slot];
    // every warp is past the previous tile's phase B (end-of-loop barrier): the list buffer may be overwritten
    const bool staged = TV_LMAX > 0 && d.list_len > 0 && d.list_len <= A.list_stage_max;
    if (threadIdx.x == 0 && staged) {
      const uint32_t bytes = (uint32_t)d.list_len * 2u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.lists)), "l"(A.lists + d.list_off),
                   "r"(bytes), "r"(mbar)
                   : "memory");
    }
    if ((int)threadIdx.x < d.nb_foot) {
      S.cx[3 * threadIdx.x] = pf.c0;
      S.cx[3 * threadIdx.x + 1] = pf.c1;
      S.cx[3 * threadIdx.x + 2] = pf.c2;
    }
    if ((int)threadIdx.x < d.nb_unit) {
      S.ubase[threadIdx.x] = pf.ubase;
      S.ulen[threadIdx.x] = pf.ulen;
    }
    if ((int)threadIdx.x < d.nb_row) {
      S.rowbeg[threadIdx.x] = pf.rowbeg;
      S.rowinfo[threadIdx.x] = pf.rowinfo;
    }
    __syncthreads();
    // ---- phase A: g_a = sqrt(s) * cofactor gradients ----
#pragma unroll
    for (int r = 0; r < TV_ROUNDS; ++r) {
      const int lc = r * TG_THREADS + threadIdx.x;
      if (lc < d.nb_cell) {
        const uint2 ln = pf.ln[r];
        if constexpr (NPC == 4) {
          const double* p0 = S.cx + 3 * (ln.x & 0xFFFFu);
          const double* p1 = S.cx + 3 * (ln.x >> 16);
          const double* p2 = S.cx + 3 * (ln.y & 0xFFFFu);
          const double* p3 = S.cx + 3 * (ln.y >> 16);
          Tet4Geom g;
          g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
          const double q = sqrt(g.s);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) S.G[(a * 3 + k) * TV_CS + lc] = g.c[a][k] * q;
        }
        else {
          const double* p0 = S.cx + 3 * (ln.x & 0xFFFFu);
          const double* p1 = S.cx + 3 * (ln.x >> 16);
          const double* p2 = S.cx + 3 * (ln.y & 0xFFFFu);
          Tri3Geom g;
          g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], false);
          const double q = sqrt(g.s);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int k = 0; k < 2; ++k) S.G[(a * 2 + k) * TV_CS + lc] = g.c[a][k] * q;
          if constexpr (OP == 1) S.G[6 * TV_CS + lc] = g.area;
        }
      }
    }
    const int64_t tn = (int64_t)t + gridDim.x, tnn = tn + gridDim.x, tnnn = tnn + gridDim.x;
    const int nslot = (slot + 1) & 3, nnslot = (slot + 2) & 3, nnnslot = (slot + 3) & 3;
    if (tn < A.nb_tile) prefetch_level2<TV_ROUNDS, TG_THREADS>(S.desc[nslot], A, pf);
    if (tnn < A.nb_tile) prefetch_level1(S.desc[nnslot], A, pf);
    int32_t desc_word = 0;
    if (threadIdx.x < DW && tnnn < A.nb_tile) desc_word = __ldg(reinterpret_cast<const int32_t*>(A.desc + tnnn) + threadIdx.x);
    __syncthreads();
    // ---- phase B: one lane per block entry; M accumulated in registers, lists streamed from global
    //      (dynamic unit hand-out and deeper list prefetch were measured slower) ----
    if (staged) {
      mbar_wait(mbar, parity);
      parity ^= 1u;
    }
    // (generic loads below: the lists live in shared memory when staged, in global memory for an oversized tile)
    const uint32_t* l32 = staged ? reinterpret_cast<const uint32_t*>(S.lists) : reinterpret_cast<const uint32_t*>(A.lists + d.list_off);
    for (int u = warp; u < d.nb_unit; u += NW) {
      const uint32_t em = __ldg(A.emap + (size_t)(d.unit_off + u) * 32 + lane);
      const uint32_t er = __ldg(A.emap_rows + (size_t)(d.unit_off + u) * 32 + lane);
      const uint32_t* l = l32 + (S.ubase[u] >> 1) + lane;
      const int len2 = S.ulen[u] >> 1;
      double M[DIM][DIM];
      double asum = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) M[i][j] = 0.0;
      uint32_t w = len2 > 0 ? l[0] : 0u;
      for (int k = 0; k < len2; ++k) {
        const uint32_t wn = (k + 1 < len2) ? l[(k + 1) * 32] : 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t code = h ? (w >> 16) : (w & 0xFFFFu);
          const uint32_t plane = code / TV_CS, lc = code - plane * TV_CS;
          const uint32_t a = plane / NPC, b = plane - a * NPC;
          double ga[DIM], gb[DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            ga[i] = S.G[(a * DIM + i) * TV_CS + lc];
            gb[i] = S.G[(b * DIM + i) * TV_CS + lc];
          }
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j) M[i][j] = fma(ga[i], gb[j], M[i][j]);
          if constexpr (OP == 1) asum = fma(S.G[6 * TV_CS + lc], a == b ? 2.0 : 1.0, asum);
        }
        w = wn;
      }
      if (em != 0xFFFFFFFFu) {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) tr += M[i][i];
        double blk[B * B];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int j = 0; j < DIM; ++j) blk[i * B + j] = lam * M[i][j] + mu * M[j][i] + (i == j ? mu * tr : 0.0);
        if constexpr (OP == 1) {
          blk[0] = 0.0;
          blk[1] = tr;
          blk[2] = tr;
          blk[3] = asum;
        }
        // (row, position) of the entry and of its mirror: rows from the plan, positions from the tile-local entry index
        auto emit = [&](int e, int lo, bool transpose) {
          const int e0 = rowinfo_erow(S.rowinfo[lo]);
          const int nz = (lo + 1 < d.nb_row ? rowinfo_erow(S.rowinfo[lo + 1]) : d.nb_entry) - e0;
          const int rb = S.rowbeg[lo], p = rb + (e - e0);
#pragma unroll
          for (int i = 0; i < B; ++i)
#pragma unroll
            for (int j = 0; j < B; ++j) {
              double* dst = A.values + value_index<B, LAYOUT>(rb, nz, p, i, j);
              const double v = transpose ? blk[j * B + i] : blk[i * B + j];
              if (A.accumulate) *dst += v; else *dst = v;
            }
        };
        emit((int)(em & 0xFFFFu), (int)(er & 0xFFFFu), false);
        if ((em >> 16) != TG_NONE16) emit((int)(em >> 16), (int)(er >> 16), true);
      }
    }
    if (threadIdx.x < DW && tnnn < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[nnnslot])[threadIdx.x] = desc_word;
    __syncthreads();
    t = (int32_t)tn;
    slot = nslot;
    if (tn >= A.nb_tile) break;
  }
}

// ---------------------------------------------------------------------------------------------
// row-ordered vector executor (isotropic elasticity, b = DIM)
//   Same cache as above (g_a = sqrt(s) grad phi_a per tile cell), different division of phase B: a unit is up to 32
//   CONSECUTIVE entries of whole rows (tiles_plan.cu: k_tile_rowlists), one lane per entry, M = sum g_a (x) g_b in
//   registers.  The blocks of a unit go to a per-warp staging buffer; the diagonal block is minus the sum of the row's
//   other blocks (block row sums of the stiffness matrix are zero: sum_b g_b = 0 in every cell), so it needs neither a list
//   nor the longest loop of the tile; then the warp writes the unit as a few contiguous runs (a row's blocks are
//   contiguous in both value layouts).  No mirror writes, no per-entry maps.
// ---------------------------------------------------------------------------------------------
template <int DIM>
struct RowsSmem {
  static constexpr int BS = DIM == 3 ? VR_BSTRIDE3 : VR_BSTRIDE2;
  double G[TV_PLANES * VR_CS];
  double cx[3 * TG_FMAX];
  double stage[(TG_THREADS / 32) * VR_STAGE];
  int2 rowtab[TG_RMAX];          // value offset of (row, entry e, block row i, column j) = x + EB * e + i * y + j
  uint32_t rowinfo[TG_RMAX + 1]; // + sentinel (first entry = nb_entry)
  uint2 units[VR_UMAX];
  uint16_t erow_of[TG_EMAX];     // tile row of every entry
  __align__(16) TileDesc desc[4];
};
static_assert(sizeof(RowsSmem<3>) <= TG_SMEM_LIMIT, "TG_MINB row-ordered executor CTAs must fit one SM");
constexpr int VR_ROUNDS = (VR_CMAX + TG_THREADS - 1) / TG_THREADS;
static_assert(VR_ROUNDS <= TG_PF_ROUNDS, "prefetch registers");

template <int ROUNDS>
__device__ __forceinline__ void prefetch_rows_level2(const TileDesc& d, const ExecArgs& A, TilePrefetch& pf)
{
  if ((int)threadIdx.x < d.nb_foot) {
    const double* p = A.coords + 3 * (int64_t)pf.fidx;
    pf.c0 = __ldg(p);
    pf.c1 = __ldg(p + 1);
    pf.c2 = __ldg(p + 2);
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int lc = min(r * TG_THREADS + (int)threadIdx.x, d.nb_cell - 1);
    if (d.nb_cell > 0) pf.ln[r] = __ldg(reinterpret_cast<const uint2*>(A.lconn) + d.cell_off + lc);
  }
  if ((int)threadIdx.x < d.nb_unit) pf.unit = __ldg(A.vr_units + d.unit_off + threadIdx.x);
  if ((int)threadIdx.x < d.nb_row) {
    pf.rowbeg = __ldg(A.rows + pf.node);
    pf.rowinfo = __ldg(A.rowinfo + d.node_off + threadIdx.x);
  }
}

template <int NPC, int LAYOUT>
__global__ void __launch_bounds__(TG_THREADS, TG_MINB) k_assemble_rows_vec(ExecArgs A, ElemParams prm)
{
  constexpr int DIM = NPC - 1, B = DIM, BB = B * B, BS = RowsSmem<DIM>::BS;
  extern __shared__ __align__(16) unsigned char ex_raw[];
  RowsSmem<DIM>& S = *reinterpret_cast<RowsSmem<DIM>*>(ex_raw);
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TG_THREADS / 32;
  constexpr int DW = sizeof(TileDesc) / 4;
  int32_t t = blockIdx.x;
  if (threadIdx.x < TV_PLANES) S.G[threadIdx.x * VR_CS + VR_CS - 1] = 0.0; // zero slot of every plane (list padding)
  if (threadIdx.x < 3 * DW) {
    const int k = threadIdx.x / DW, w = threadIdx.x % DW;
    const int64_t tt = (int64_t)t + (int64_t)k * gridDim.x;
    if (tt < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[k])[w] = __ldg(reinterpret_cast<const int32_t*>(A.desc + tt) + w);
  }
  __syncthreads();
  int slot = 0;
  TilePrefetch pf;
  if (t < A.nb_tile) {
    prefetch_level1(S.desc[0], A, pf);
    prefetch_rows_level2<VR_ROUNDS>(S.desc[0], A, pf);
    if ((int64_t)t + gridDim.x < A.nb_tile) prefetch_level1(S.desc[1], A, pf);
  }
  const double lam = prm.p0, mu = prm.p1;
  double* const stg = S.stage + warp * VR_STAGE;
  while (t < A.nb_tile) {
    const TileDesc d = S.desc[slot];
    if ((int)threadIdx.x < d.nb_foot) {
      S.cx[3 * threadIdx.x] = pf.c0;
      S.cx[3 * threadIdx.x + 1] = pf.c1;
      S.cx[3 * threadIdx.x + 2] = pf.c2;
    }
    if ((int)threadIdx.x < d.nb_unit) S.units[threadIdx.x] = pf.unit;
    if ((int)threadIdx.x < d.nb_row) S.rowinfo[threadIdx.x] = pf.rowinfo;
    else if ((int)threadIdx.x == d.nb_row) S.rowinfo[threadIdx.x] = pack_rowinfo(d.nb_entry, 0, false);
    __syncthreads();
    // the next tile's contribution lists start their way from DRAM to L2 now (one bulk prefetch): phase B reads them from L2
    if (threadIdx.x == 0 && (int64_t)t + gridDim.x < A.nb_tile) {
      const TileDesc& dn = S.desc[(slot + 1) & 3];
      if (dn.list_len > 0)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(A.lists + dn.list_off), "r"((uint32_t)dn.list_len * 2u) : "memory");
    }
    // ---- phase A: g_a = sqrt(s) * cofactor gradients ----
#pragma unroll
    for (int r = 0; r < VR_ROUNDS; ++r) {
      const int lc = r * TG_THREADS + threadIdx.x;
      if (lc < d.nb_cell) {
        const uint2 ln = pf.ln[r];
        if constexpr (NPC == 4) {
          const double* p0 = S.cx + 3 * (ln.x & 0xFFFFu);
          const double* p1 = S.cx + 3 * (ln.x >> 16);
          const double* p2 = S.cx + 3 * (ln.y & 0xFFFFu);
          const double* p3 = S.cx + 3 * (ln.y >> 16);
          Tet4Geom g;
          g.init_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
          const double q = sqrt(g.s);
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int k = 0; k < 3; ++k) S.G[(a * 3 + k) * VR_CS + lc] = g.c[a][k] * q;
        }
        else {
          const double* p0 = S.cx + 3 * (ln.x & 0xFFFFu);
          const double* p1 = S.cx + 3 * (ln.x >> 16);
          const double* p2 = S.cx + 3 * (ln.y & 0xFFFFu);
          Tri3Geom g;
          g.init_xy(p0[0], p0[1], p1[0], p1[1], p2[0], p2[1], false);
          const double q = sqrt(g.s);
#pragma unroll
          for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int k = 0; k < 2; ++k) S.G[(a * 2 + k) * VR_CS + lc] = g.c[a][k] * q;
        }
      }
    }
    // row tables (thread per row): where the row's values live, and the row of each of its entries
    if ((int)threadIdx.x < d.nb_row) {
      const int e0 = rowinfo_erow(pf.rowinfo), nz = rowinfo_erow(S.rowinfo[threadIdx.x + 1]) - e0;
      if constexpr (LAYOUT == AFB_LAYOUT_PER_BLOCK) S.rowtab[threadIdx.x] = make_int2(BB * (pf.rowbeg - e0), B);
      else S.rowtab[threadIdx.x] = make_int2(BB * pf.rowbeg - B * e0, B * nz);
      for (int x = 0; x < nz; ++x) S.erow_of[e0 + x] = (uint16_t)threadIdx.x;
    }
    const int64_t tn = (int64_t)t + gridDim.x, tnn = tn + gridDim.x, tnnn = tnn + gridDim.x;
    const int nslot = (slot + 1) & 3, nnslot = (slot + 2) & 3, nnnslot = (slot + 3) & 3;
    if (tn < A.nb_tile) prefetch_rows_level2<VR_ROUNDS>(S.desc[nslot], A, pf);
    if (tnn < A.nb_tile) prefetch_level1(S.desc[nnslot], A, pf);
    int32_t desc_word = 0;
    if (threadIdx.x < DW && tnnn < A.nb_tile) desc_word = __ldg(reinterpret_cast<const int32_t*>(A.desc + tnnn) + threadIdx.x);
    __syncthreads();
    // ---- phase B: one lane per entry of the unit ----
    // staged block of lane el, element (i, j): per-block layout [entry][i][j] (stride BS, odd); per-row layout [i][entry][j]
    // (plane stride PS): either way the write-out below reads consecutive words
    constexpr int PS = DIM == 3 ? VR_PSTRIDE3 : VR_PSTRIDE2;
    auto sidx = [](int el, int i, int j) { return LAYOUT == AFB_LAYOUT_PER_BLOCK ? el * BS + i * B + j : i * PS + el * B + j; };
    const uint32_t* l32 = reinterpret_cast<const uint32_t*>(A.lists + d.list_off);
    constexpr int LW = 4; // list words (2 contributions each) held in registers per unit; longer lists continue from memory
    uint32_t wq[LW];
    uint2 U = make_uint2(0u, 0u);
    auto fetch = [&](int u) {
      if (u < d.nb_unit) {
        U = S.units[u];
        const uint32_t* l = l32 + (U.x >> 1) + lane;
        const int len2 = (int)(U.y >> 18);
#pragma unroll
        for (int k = 0; k < LW; ++k) wq[k] = k < len2 ? __ldg(l + k * 32) : (uint32_t)(VR_CS - 1) * 0x10001u;
      }
    };
    fetch(warp);
    for (int u = warp; u < d.nb_unit; u += NW) {
      const int first = (int)(U.y & 0xFFFu), cnt = (int)((U.y >> 12) & 31u) + 1, len2 = (int)(U.y >> 18);
      const uint32_t* l = l32 + (U.x >> 1) + lane;
      uint32_t wc[LW];
#pragma unroll
      for (int k = 0; k < LW; ++k) wc[k] = wq[k];
      fetch(u + NW); // the next unit's words travel while this one is summed
      // where this lane's entry goes (read back through shuffles by the write-out)
      int my_base = 0, my_y = 0;
      if (lane < cnt) {
        const int2 T = S.rowtab[S.erow_of[first + lane]];
        my_base = T.x + (LAYOUT == AFB_LAYOUT_PER_BLOCK ? BB : B) * (first + lane);
        my_y = T.y;
      }
      double M[DIM][DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) M[i][j] = 0.0;
      auto add2 = [&](uint32_t w) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t code = h ? (w >> 16) : (w & 0xFFFFu);
          const uint32_t lc = code & VR_LC_MASK, a = code >> (VR_LC_BITS + 2), b = (code >> VR_LC_BITS) & 3u;
          const double* pa = S.G + (a * (DIM * VR_CS) + lc);
          const double* pb = S.G + (b * (DIM * VR_CS) + lc);
          double ga[DIM], gb[DIM];
#pragma unroll
          for (int i = 0; i < DIM; ++i) {
            ga[i] = pa[i * VR_CS];
            gb[i] = pb[i * VR_CS];
          }
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j) M[i][j] = fma(ga[i], gb[j], M[i][j]);
        }
      };
      if (len2 <= LW) {
#pragma unroll
        for (int k = 0; k < LW; ++k)
          if (k < len2) add2(wc[k]);
      }
      else {
#pragma unroll
        for (int k = 0; k < LW; ++k) add2(wc[k]);
        for (int k = LW; k < len2; ++k) add2(__ldg(l + k * 32));
      }
      {
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) tr += M[i][i];
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int j = 0; j < DIM; ++j) stg[sidx(lane, i, j)] = lam * M[i][j] + mu * M[j][i] + (i == j ? mu * tr : 0.0);
      }
      __syncwarp();
      // diagonal blocks of the rows that lie wholly inside the unit: minus the sum of the row's other blocks (zero block row
      // sums; the diagonal's own slot was staged as zero, so the whole row is summed).  The block is symmetric: its upper
      // triangle is summed and mirrored.  Lanes: NS components x 2 halves of the row x RPP rows per pass.
      {
        constexpr int NS = B * (B + 1) / 2, RPP = (32 / NS) / 2;
        const int r0 = S.erow_of[first], r1 = S.erow_of[first + cnt - 1];
        const int g = lane / NS, c = lane - g * NS, q = g >> 1, h = g & 1;
        const int ci = B == 3 ? (c < 3 ? 0 : (c < 5 ? 1 : 2)) : (c < 2 ? 0 : 1), cj = B == 3 ? (c < 3 ? c : (c < 5 ? c - 2 : 2)) : (c < 2 ? c : 1);
        constexpr int XS = LAYOUT == AFB_LAYOUT_PER_BLOCK ? BS : B; // stride between the entries of a row
        for (int rb = r0; rb <= r1; rb += RPP) {
          const int rr = rb + q;
          double sum = 0.0;
          int dslot = -1;
          if (g < 2 * RPP && rr <= r1) {
            const uint32_t ri = S.rowinfo[rr];
            const int e0 = rowinfo_erow(ri), nz = rowinfo_erow(S.rowinfo[rr + 1]) - e0;
            if (nz <= 32) {
              const int half = (nz + 1) >> 1, xe = h ? nz : half;
              const double* src = stg + sidx(e0 - first, ci, cj);
              double s0 = 0.0, s1 = 0.0;
              int x = h ? half : 0;
              for (; x + 1 < xe; x += 2) {
                s0 += src[x * XS];
                s1 += src[(x + 1) * XS];
              }
              if (x < xe) s0 += src[x * XS];
              sum = s0 + s1;
              dslot = e0 - first + rowinfo_pdiag(ri);
            }
          }
          const double other = __shfl_down_sync(0xffffffffu, sum, NS); // the second half of the row sits NS lanes above
          __syncwarp();
          if (h == 0 && dslot >= 0) {
            const double v = -(sum + other);
            stg[sidx(dslot, ci, cj)] = v;
            if (ci != cj) stg[sidx(dslot, cj, ci)] = v;
          }
        }
      }
      __syncwarp();
      // write-out: contiguous runs
      auto put = [&](auto acc) {
        constexpr bool ACC = decltype(acc)::value;
        if constexpr (LAYOUT == AFB_LAYOUT_PER_BLOCK && BS == BB) {
          // a row's blocks are one contiguous run, in the staging buffer and in `values`: copied row by row
          const int r0 = S.erow_of[first], r1 = S.erow_of[first + cnt - 1];
          for (int rr = r0; rr <= r1; ++rr) {
            const int e0 = rowinfo_erow(S.rowinfo[rr]), e1 = rowinfo_erow(S.rowinfo[rr + 1]);
            const int lo = max(e0, first), hi = min(e1, first + cnt);
            double* dst = A.values + ((int64_t)S.rowtab[rr].x + (int64_t)BB * lo);
            const double* src = stg + (lo - first) * BS;
            for (int x = lane; x < (hi - lo) * BB; x += 32) {
              if (ACC) dst[x] += src[x]; else dst[x] = src[x];
            }
          }
        }
        else if constexpr (LAYOUT == AFB_LAYOUT_PER_BLOCK) {
          for (int x0 = 0; x0 < cnt * BB; x0 += 32) {
            const int x = x0 + lane, el = x / BB, c = x - el * BB;
            const int base = __shfl_sync(0xffffffffu, my_base, el & 31);
            if (x < cnt * BB) {
              double* dst = A.values + ((int64_t)base + c);
              const double v = stg[el * BS + c];
              if (ACC) *dst += v; else *dst = v;
            }
          }
        }
        else {
          for (int x0 = 0; x0 < cnt * B; x0 += 32) {
            const int x = x0 + lane, el = x / B, j = x - el * B;
            const int base = __shfl_sync(0xffffffffu, my_base, el & 31), ys = __shfl_sync(0xffffffffu, my_y, el & 31);
            if (x < cnt * B) {
              double* dst = A.values + ((int64_t)base + j);
#pragma unroll
              for (int i = 0; i < B; ++i) {
                const double v = stg[i * PS + x];
                if (ACC) dst[i * ys] += v; else dst[i * ys] = v;
              }
            }
          }
        }
      };
      if (A.accumulate) put(std::true_type()); else put(std::false_type());
      __syncwarp();
    }
    if (threadIdx.x < DW && tnnn < A.nb_tile) reinterpret_cast<int32_t*>(&S.desc[nnnslot])[threadIdx.x] = desc_word;
    __syncthreads();
    t = (int32_t)tn;
    slot = nslot;
    if (tn >= A.nb_tile) break;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int ensure_values_zeroed(afb_ctx* ctx)
{
  if (!ctx->values_dirty) return AFB_OK;
  AFB_CUDA(cudaMemsetAsync(ctx->values.p, 0, sizeof(double) * (size_t)ctx->nnz * ctx->b * ctx->b, ctx->stream));
  ctx->values_dirty = false;
  return AFB_OK;
}

int assemble_tiled(afb_ctx* ctx, int op, const double* params, int layout, int flags)
{
  const bool vec = ctx->b > 1;
  AFB_REQUIRE(ctx->npc == ctx->dim + 1 && ((op == AFB_OP_POISSON && !vec) || (op == AFB_OP_ELASTICITY && vec) || (op == AFB_OP_BILAPLACIAN && vec && ctx->npc == 3)),
              AFB_ERR_UNSUPPORTED,
              "AFB_VARIANT_TILED_GATHER is not available for operator %d on %d-node cells (P1 Poisson, P1 elasticity, Tri3 bilaplacian only); use AFB_VARIANT_NODEWISE", op, ctx->npc);
  AFB_REQUIRE(!ctx->has_cell_coef || (!vec && ctx->tiled_exec == AFB_TILED_EXEC_BRICKS), AFB_ERR_UNSUPPORTED,
              "AFB_VARIANT_TILED_GATHER takes a per-cell coefficient (afb_set_cell_coefficient) for the Poisson operator with the brick executor only; use AFB_VARIANT_NODEWISE");
  TilePlan& P = ctx->plan;
  const int mode = flags & (AFB_FLAG_ALL_ROWS | AFB_FLAG_OWN_CELLS_ONLY);
  // scalar operators: the chained-slice executors when selected (afb_set_tiled_executor; chain_exec.cu, chain_flow.cu)
  if (!vec && ctx->tiled_exec != AFB_TILED_EXEC_BRICKS) {
    ElemParams prm;
    prm.p0 = params ? params[0] : 0.0;
    prm.p1 = params ? params[1] : 0.0;
    prm.flags = flags;
    return chain_assemble(ctx, prm, flags, (ctx->assembled || ctx->values_touched) ? 1 : 0);
  }
  const bool rows_exec = vec && op == AFB_OP_ELASTICITY && ctx->vec_rows(); // row-ordered units (k_assemble_rows_vec)
  const int cls = !vec ? 0 : (rows_exec ? 2 : 1);
  if (!P.mesh_valid || P.mesh_gen != ctx->mesh_gen || P.mesh_b_class != cls) AFB_TRY(build_tile_mesh(ctx, cls));
  if (!P.lists_valid || P.lists_mesh_gen != ctx->mesh_gen || P.lists_b != ctx->b || P.lists_mode != mode || P.lists_kind != (rows_exec ? 1 : 0))
    AFB_TRY(rows_exec ? build_tile_rowlists(ctx, mode) : build_tile_lists(ctx, mode));
  ElemParams prm;
  prm.p0 = params ? params[0] : 0.0;
  prm.p1 = params ? params[1] : 0.0;
  prm.flags = flags;
  if (P.nb_tile == 0) return AFB_OK;
  // values already holding contributions (a second operator added on top) are accumulated into;
  // a fresh matrix is simply overwritten (every entry of every row is written exactly once; the
  // vector executor does not write the rows of non-owned nodes: those need the zero fill)
  const int accumulate = (ctx->assembled || ctx->values_touched) ? 1 : 0;
  const bool all_rows = ctx->all_own || (flags & AFB_FLAG_ALL_ROWS);
  if (accumulate || (vec && !all_rows)) AFB_TRY(ensure_values_zeroed(ctx));
  else ctx->values_dirty = false;
  ExecArgs A;
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
  if (vec) A.emap_rows = P.emap_rows.as<uint32_t>();
  else A.tile_cells = P.tile_cells.as<int32_t>();
  A.lists = P.lists.as<uint16_t>();
  A.vr_units = rows_exec ? P.vr_units.as<uint2>() : nullptr;
  A.values = ctx->values.as<double>();
  A.accumulate = accumulate;
  A.list_stage_max = (int)std::min<int64_t>(vec ? TV_LMAX : TG_LMAX, ctx->tiled_stage_limit / 2);
  if (ctx->has_cell_coef) prm.cell_coef = ctx->cell_coef.as<double>();
  const int grid = std::min<int>(P.nb_tile, TG_MINB * ctx->sm_count);
  auto go = [&](auto kernel, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = launch_pdl(kernel, grid, TG_THREADS, smem, ctx->stream, A, prm);
    return e != cudaSuccess ? e : cudaGetLastError();
  };
  cudaError_t e;
  if (!vec && ctx->has_cell_coef) e = ctx->npc == 4 ? go(k_assemble_tiled_coef<4>, sizeof(ExecSmem)) : go(k_assemble_tiled_coef<3>, sizeof(ExecSmem));
  else if (!vec) e = ctx->npc == 4 ? go(k_assemble_tiled<4>, sizeof(ExecSmem)) : go(k_assemble_tiled<3>, sizeof(ExecSmem));
  else if (rows_exec && ctx->npc == 4)
    e = layout == AFB_LAYOUT_PER_BLOCK ? go(k_assemble_rows_vec<4, AFB_LAYOUT_PER_BLOCK>, sizeof(RowsSmem<3>)) : go(k_assemble_rows_vec<4, AFB_LAYOUT_PER_ROW>, sizeof(RowsSmem<3>));
  else if (rows_exec)
    e = layout == AFB_LAYOUT_PER_BLOCK ? go(k_assemble_rows_vec<3, AFB_LAYOUT_PER_BLOCK>, sizeof(RowsSmem<2>)) : go(k_assemble_rows_vec<3, AFB_LAYOUT_PER_ROW>, sizeof(RowsSmem<2>));
  else if (ctx->npc == 4)
    e = layout == AFB_LAYOUT_PER_BLOCK ? go(k_assemble_tiled_vec<4, AFB_LAYOUT_PER_BLOCK>, sizeof(VecSmem<3>)) : go(k_assemble_tiled_vec<4, AFB_LAYOUT_PER_ROW>, sizeof(VecSmem<3>));
  else if (op == AFB_OP_BILAPLACIAN)
    e = layout == AFB_LAYOUT_PER_BLOCK ? go(k_assemble_tiled_vec<3, AFB_LAYOUT_PER_BLOCK, 1>, sizeof(VecSmem<2>)) : go(k_assemble_tiled_vec<3, AFB_LAYOUT_PER_ROW, 1>, sizeof(VecSmem<2>));
  else
    e = layout == AFB_LAYOUT_PER_BLOCK ? go(k_assemble_tiled_vec<3, AFB_LAYOUT_PER_BLOCK>, sizeof(VecSmem<2>)) : go(k_assemble_tiled_vec<3, AFB_LAYOUT_PER_ROW>, sizeof(VecSmem<2>));
  AFB_CUDA(e);
  ctx->launches++;
  return AFB_OK;
}

} // namespace afb
