// Pipelined (warp-specialised) executor of the scalar tiled gather on chained slices -- the default executor of
// AFB_VARIANT_TILED_GATHER for b = 1 (plan: chain_plan.cu, layout: chain.cuh; the phase-separated twin is chain_exec.cu).
//
// Reference behaviour replaced: _assembleNodeWiseCsrBilinearOperator{Tria3,Tetra4}
// (modules/testlab/NodeWiseCsrBiliAssembly.cc:157-297) and BSRFormat::assembleBilinearAtomicFree
// (femutils/BSRFormat.h:406-577) for b = 1: every matrix row is written exactly once, by one owner, without atomics
// and without a zero fill.
//
// The phases of a slice use different pipes of the SM (element matrices: fp64; gather: shared memory; write-out: HBM).
// Separated by block barriers they run one after the other and every barrier waits for the slowest warp
// (profiles/r02c: 30 % of the stall samples).  Here ONE persistent CTA per SM runs the phases as a pipeline of warp
// groups that hand slices to each other through mbarriers (no block barrier after the prologue):
//
//   loaders (4 warps, slices in turn)     descriptor, TMA bulk copies of the plan record and the local connectivity,
//                                         gathers of footprint coordinates and row offsets          -> full[j % 2]
//   element warps (phase A)               a stream of 32-cell batches, dealt round-robin over the warps across slice
//                                         boundaries; K_e off-diagonals to cache region sidx % 3      -> a_done[j % 3]
//   gather warps (phase B)                units of 32 entries: 4 cache indices per 64-bit list word; values to the
//                                         slice's staging row, the twin entry, or the next slice's row -> b_done[j % 3]
//   row warps (phase C)                   diagonal = -(sum of the row's off-diagonals), then the warp's rows leave shared
//                                         memory as contiguous, coalesced stores                      -> c_done[j % 3]
//
// Ring depths: inputs 2 slices, cache regions 3 (the element warps run up to one slice ahead of the gather), staging
// rows 3.  A region / buffer is reused only after the mbarrier of its last reader has completed (see the waits).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "chain.cuh"
#include "element.cuh"
#include "tiles.cuh"

namespace afb {

struct FlowArgs {
  const SliceDesc* desc;   // in execution order: the slices of CTA c are desc[cta_ptr[c] .. cta_ptr[c + 1])
  const int32_t* order;
  const int32_t* cta_ptr;
  const double* coords;
  const int32_t* foot;
  const uint2* lconn;
  const int32_t* slice_nodes;
  const int32_t* rows;
  const unsigned char* blob;
  const unsigned char* iblock; // index blocks in execution order (descriptor | footprint node ids | row node ids)
  const int32_t* ib_off;       // their starts, 16-byte units
  double* values;
  int* error;      // device flag: set (and the kernel trapped) when a pipeline wait times out
  int accumulate;
  int stage_max;   // records larger than this are read from global memory (test knob; <= G::BLOB)
  int prof;        // accumulate the cycles spent in every pipeline wait (error[1..16] as 64-bit counters)
};

#ifndef AFB_FL_NA
#define AFB_FL_NL 4
#define AFB_FL_NA 16
#define AFB_FL_NB 4
#define AFB_FL_NC 4
#endif
constexpr int FL_NL = AFB_FL_NL, FL_NA = AFB_FL_NA, FL_NB = AFB_FL_NB, FL_NC = AFB_FL_NC; // loader / element / gather / row warps
constexpr int FL_IBD = 6;                                                   // index blocks in flight
constexpr int FL_IB_MAX = 64 + 4 * ChainGeomF::FMAX + 4 * ChainGeomF::RMAX; // descriptor + footprint + rows
static_assert((FL_NL + FL_NA + FL_NB + FL_NC) * 32 == ChainGeomF::THREADS, "warp roles cover the CTA");

template <class G, int NPC>
struct FlowSmem {
  static constexpr int NPAIR = NPC * (NPC - 1) / 2;
  static constexpr int ZERO = NPAIR * G::PLANE;
  double Kc[ZERO + 1];
  double vout[3][G::EMAX];
  // inputs of the element warps (ring of 2): footprint coordinates, local connectivity of the first CN cells, descriptor
  double cx[2][3 * G::FMAX];
  __align__(16) uint2 lconn[2][G::CN];
  __align__(16) SliceDesc desc_a[2];
  // inputs of the gather and row warps (ring of 3): plan record, first value of every row, descriptor
  __align__(16) unsigned char blob[3][G::BLOB];
  int32_t rowbeg[3][G::RMAX];
  __align__(16) SliceDesc desc_b[3];
  // index blocks of the slices ahead (ring of FL_IBD, TMA)
  __align__(16) unsigned char ib[FL_IBD][FL_IB_MAX];
  __align__(8) unsigned long long bar_fa[2], bar_fb[3], bar_a[3], bar_b[3], bar_c[3], bar_ib[FL_IBD];
};

__device__ __forceinline__ void fl_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// error[0]: time-out code; error[1..16]: cycles spent waiting, per wait site (only when profiling: AFB_FLOW_PROF=1)
__device__ __forceinline__ void fl_wait(uint32_t bar, unsigned parity, int* error, int code, bool prof = false)
{
  unsigned done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) { // ~2 s: a broken hand-off must not hang the device
      atomicExch(error, code);
      __threadfence_system();
      __trap();
    }
  }
  if (prof && (threadIdx.x & 31) == 0) atomicAdd(reinterpret_cast<unsigned long long*>(error) + code, (unsigned long long)(clock64() - t0));
}

template <int NPC>
__device__ __forceinline__ void flow_cell(const double* __restrict__ cx, uint2 ln, const ElemParams& prm, double (&K)[6])
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

template <class G, int NPC>
__global__ void __launch_bounds__(G::THREADS, 1) k_assemble_flow(FlowArgs A, ElemParams prm)
{
  using SM = FlowSmem<G, NPC>;
  constexpr int NPAIR = SM::NPAIR;
  constexpr int DW = sizeof(SliceDesc) / 4;
  extern __shared__ __align__(16) unsigned char fl_raw[];
  SM& S = *reinterpret_cast<SM*>(fl_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t it0 = __ldg(A.cta_ptr + blockIdx.x);
  const int n = __ldg(A.cta_ptr + blockIdx.x + 1) - it0; // slices of this CTA
  if (n <= 0) return;
  if (tid == 0) {
    S.Kc[SM::ZERO] = 0.0;
    for (int k = 0; k < 2; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.bar_fa[k])), "r"(FL_NL * 32 + 2)); // every loader thread's async copies + TMA issue (tx bytes) + descriptor
    for (int k = 0; k < 3; ++k) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.bar_fb[k])), "r"(FL_NL * 32 + 2));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.bar_a[k])), "r"(FL_NA));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.bar_b[k])), "r"(FL_NB));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&S.bar_c[k])), "r"(FL_NC));
    }
    for (int k = 0; k < FL_IBD; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar_ib[k])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto bar_fa = [&](int j) { return smem_u32(&S.bar_fa[j & 1]); };
  auto bar_fb = [&](int j) { return smem_u32(&S.bar_fb[j % 3]); };
  auto bar_a = [&](int j) { return smem_u32(&S.bar_a[j % 3]); };
  auto bar_b = [&](int j) { return smem_u32(&S.bar_b[j % 3]); };
  auto bar_c = [&](int j) { return smem_u32(&S.bar_c[j % 3]); };
  auto par2 = [](int j) { return (unsigned)((j >> 1) & 1); };
  auto par3 = [](int j) { return (unsigned)((j / 3) & 1); };

  if (warp < FL_NL) {
    // =============================== loaders ===============================
    // The FL_NL loader warps stage every slice together (thread t: footprint nodes t, t + 128, ...; row t) and never wait
    // for memory themselves:
    //   index blocks   (descriptor, footprint node ids, row node ids; contiguous in execution order) stream through the
    //                  TMA engine into a ring FL_IBD slices deep
    //   iteration j    pulls the coordinates / row offsets of slice j + 3 and the record / connectivity of slice j + 2 into
    //                  L2; once the slots of slice j are free it issues the gathers of its coordinates and row offsets as
    //                  asynchronous copies (cp.async, global -> shared) and the TMA copies of its connectivity and plan
    //                  record -- all of them complete on the slot's mbarrier
    constexpr int LT = FL_NL * 32;
    constexpr int FQ = (G::FMAX + LT - 1) / LT, RQ = (G::RMAX + LT - 1) / LT;
    const int t = tid; // loader warps are the first warps of the CTA
    auto ib_bar = [&](int j) { return smem_u32(&S.bar_ib[j % FL_IBD]); };
    auto ib_par = [](int j) { return (unsigned)((j / FL_IBD) & 1); };
    auto ib_desc = [&](int j) { return reinterpret_cast<const SliceDesc*>(S.ib[j % FL_IBD]); };
    auto ib_issue = [&](int j, int32_t o0, int32_t o1) { // thread 0: index block of slice j -> ring
      const uint32_t bytes = (uint32_t)(o1 - o0) * 16u;
      const uint32_t bar = ib_bar(j);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.ib[j % FL_IBD])), "l"(A.iblock + (size_t)o0 * 16), "r"(bytes),
                   "r"(bar)
                   : "memory");
    };
    auto pull_l2 = [&](int j) { // the gathers slice j will do: lines into L2 now
      const SliceDesc* d = ib_desc(j);
      const int nb_foot = d->nb_foot, nb_row = d->nb_row;
      const int32_t* fidx = reinterpret_cast<const int32_t*>(S.ib[j % FL_IBD] + ch_ib_foot());
      const int32_t* nidx = reinterpret_cast<const int32_t*>(S.ib[j % FL_IBD] + ch_ib_nodes(nb_foot));
#pragma unroll
      for (int k = 0; k < FQ; ++k) {
        const int i = k * LT + t;
        if (i < nb_foot) {
          const char* q = reinterpret_cast<const char*>(A.coords + 3 * (int64_t)fidx[i]);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 16)); // a 24-byte record may straddle two sectors
        }
      }
#pragma unroll
      for (int k = 0; k < RQ; ++k) {
        const int i = k * LT + t;
        if (i < nb_row) asm volatile("prefetch.global.L2 [%0];" ::"l"(A.rows + nidx[i]));
      }
    };
    auto loaders_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(FL_NL * 32) : "memory"); };
    int32_t off_next = 0, off_next1 = 0; // thread 0: bounds of the index block it issues next
    if (t == 0) {
      for (int j = 0; j < min(n, FL_IBD); ++j) ib_issue(j, __ldg(A.ib_off + it0 + j), __ldg(A.ib_off + it0 + j + 1));
      if (FL_IBD < n) {
        off_next = __ldg(A.ib_off + it0 + FL_IBD);
        off_next1 = __ldg(A.ib_off + it0 + FL_IBD + 1);
      }
    }
    int ibw = 0; // first index block this warp has not waited for yet (every phase of every ring slot is observed in order)
    auto ib_need = [&](int q) {
      for (; ibw <= q && ibw < n; ++ibw) fl_wait(ib_bar(ibw), ib_par(ibw), A.error, 11, A.prof != 0);
    };
    for (int j = 0; j < n; ++j) {
      // ---- requests for the slices ahead ----
      ib_need(j + 3);
      if (j + 3 < n) pull_l2(j + 3);
      if (j + 2 < n) { // record and connectivity of slice j + 2: read exactly once, straight from HBM otherwise
        const SliceDesc* d2 = ib_desc(j + 2);
        const char* b1 = reinterpret_cast<const char*>(A.lconn + d2->cell_off);
        for (size_t o = (size_t)t * 128; o < (size_t)d2->nb_new * 8; o += (size_t)LT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b1 + o));
        const char* b2 = reinterpret_cast<const char*>(A.blob + (size_t)d2->blob_off * 16);
        for (size_t o = (size_t)t * 128; o < (size_t)d2->blob_bytes; o += (size_t)LT * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(b2 + o));
      }
      // ---- slots of slice j ----
      const int pa = j & 1, pb = j % 3;
      const SliceDesc* dj = ib_desc(j);
      const int32_t cell_off = dj->cell_off, nb_new = dj->nb_new, blob_bytes = dj->blob_bytes, nb_foot = dj->nb_foot, nb_row = dj->nb_row;
      const uint32_t blob_off = dj->blob_off;
      const bool staged = blob_bytes <= A.stage_max;
      const int32_t dword = lane < DW ? reinterpret_cast<const int32_t*>(dj)[lane] : 0;
      const int32_t* fidx = reinterpret_cast<const int32_t*>(S.ib[j % FL_IBD] + ch_ib_foot());
      const int32_t* nidx = reinterpret_cast<const int32_t*>(S.ib[j % FL_IBD] + ch_ib_nodes(nb_foot));
      // element-warp inputs: last read by the element warps of slice j - 2
      if (j >= 2) fl_wait(bar_a(j - 2), par3(j - 2), A.error, 1, A.prof != 0);
      {
        const uint32_t bar = bar_fa(j);
        if (t == 0) {
          const uint32_t b_conn = (uint32_t)((min(nb_new, G::CN) + 1) & ~1) * 8u; // cell offsets are even: 16-byte granules
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_conn) : "memory");
          if (b_conn)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.lconn[pa])), "l"(A.lconn + cell_off), "r"(b_conn),
                         "r"(bar)
                         : "memory");
        }
        const uint32_t cx = smem_u32(S.cx[pa]);
#pragma unroll
        for (int k = 0; k < FQ; ++k) {
          const int i = k * LT + t;
          if (i < nb_foot) {
            const double* q = A.coords + 3 * (int64_t)fidx[i];
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(cx + 24u * (uint32_t)i), "l"(q) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(cx + 24u * (uint32_t)i + 8u), "l"(q + 1) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(cx + 24u * (uint32_t)i + 16u), "l"(q + 2) : "memory");
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); // this thread's copies -> one arrival
        if (warp == 0) {
          if (lane < DW) reinterpret_cast<int32_t*>(&S.desc_a[pa])[lane] = dword;
          __syncwarp();
          if (lane == 0) fl_arrive(bar);
        }
      }
      // gather / row-warp inputs: last read by the row warps of slice j - 3
      if (j >= 3) fl_wait(bar_c(j - 3), par3(j - 3), A.error, 2, A.prof != 0);
      {
        const uint32_t bar = bar_fb(j);
        if (t == 0) {
          const uint32_t b_blob = staged ? (uint32_t)blob_bytes : 0u;
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b_blob) : "memory");
          if (b_blob)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(S.blob[pb])), "l"(A.blob + (size_t)blob_off * 16),
                         "r"(b_blob), "r"(bar)
                         : "memory");
        }
        const uint32_t rbs = smem_u32(S.rowbeg[pb]);
#pragma unroll
        for (int k = 0; k < RQ; ++k) {
          const int i = k * LT + t;
          if (i < nb_row) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rbs + 4u * (uint32_t)i), "l"(A.rows + nidx[i]) : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        if (warp == 0) {
          if (lane < DW) reinterpret_cast<int32_t*>(&S.desc_b[pb])[lane] = dword;
          __syncwarp();
          if (lane == 0) fl_arrive(bar);
        }
      }
      // ---- the index block of slice j is consumed by all loader warps: its slot takes the block of slice j + FL_IBD ----
      loaders_sync();
      if (t == 0 && j + FL_IBD < n) {
        ib_issue(j + FL_IBD, off_next, off_next1);
        off_next = off_next1;
        if (j + FL_IBD + 1 < n) off_next1 = __ldg(A.ib_off + it0 + j + FL_IBD + 2); // consumed one iteration from now
      }
    }
  }
  else if (warp < FL_NL + FL_NA) {
    // =============================== element warps (phase A) ===============================
    const int w = warp - FL_NL;
    int64_t g = w;      // next global batch of this warp
    int64_t gbase = 0;  // global number of the current slice's first batch
    for (int j = 0; j < n; ++j) {
      fl_wait(bar_fa(j), par2(j), A.error, 3, A.prof != 0);
      const SliceDesc& d = S.desc_a[j & 1];
      const int nb_new = d.nb_new, nb_a = d.nb_a, flags = d.flags;
      const uint2* lc_g = A.lconn + d.cell_off;
      const int nbatch = (nb_new + 31) >> 5;
      if (g < gbase + nbatch) {
        // the regions written here were last read by the gather of two slices ago (a segment's first slice also
        // overwrites the region the previous slice inherited from: wait for the previous gather)
        if (flags & CH_FLAG_FIRST) {
          if (j >= 1) fl_wait(bar_b(j - 1), par3(j - 1), A.error, 4, A.prof != 0);
        }
        else if (j >= 2) fl_wait(bar_b(j - 2), par3(j - 2), A.error, 5, A.prof != 0);
        const int sidx = flags >> CH_FLAG_SIDX_SHIFT;
        const int base_own = ch_reg_new(3, sidx) * G::CS, base_b = ch_reg_b(3) * G::CS - nb_a;
        const double* cx = S.cx[j & 1];
        const uint2* lc_s = S.lconn[j & 1];
        for (; g < gbase + nbatch; g += FL_NA) {
          const int lc = (int)(g - gbase) * 32 + lane;
          if (lc < nb_new) {
            double K[6];
            // (a segment's first slice may compute more cells than the staged connectivity holds: the rest comes from global memory)
            flow_cell<NPC>(cx, lc < G::CN ? lc_s[lc] : __ldg(lc_g + lc), prm, K);
            const int pos = lc < nb_a ? base_own + lc : base_b + lc;
#pragma unroll
            for (int q = 0; q < NPAIR; ++q) S.Kc[q * G::PLANE + pos] = K[q];
          }
        }
      }
      gbase += nbatch;
      __syncwarp();
      if (lane == 0) fl_arrive(bar_a(j));
    }
  }
  else if (warp < FL_NL + FL_NA + FL_NB) {
    // =============================== gather warps (phase B) ===============================
    const int w = warp - FL_NL - FL_NA;
    for (int j = 0; j < n; ++j) {
      fl_wait(bar_fb(j), par3(j), A.error, 6, A.prof != 0);
      const SliceDesc& d = S.desc_b[j % 3];
      const int nb_unit = d.nb_unit, nb_chunk = d.nb_chunk;
      const bool staged = d.blob_bytes <= A.stage_max;
      const unsigned char* rec_g = A.blob + (size_t)d.blob_off * 16;
      fl_wait(bar_a(j), par3(j), A.error, 7, A.prof != 0);
      if (j >= 2) fl_wait(bar_c(j - 2), par3(j - 2), A.error, 8, A.prof != 0); // the next slice's staging row was last read by the row warps two slices ago
      double* vcur = S.vout[j % 3];
      double* vnext = S.vout[(j + 1) % 3];
      auto phase_b = [&](const unsigned char* rec) {
        const uint2* lists = reinterpret_cast<const uint2*>(rec);
        const uint32_t* emap = reinterpret_cast<const uint32_t*>(rec + ch_off_emap(nb_chunk));
        const uint32_t* units = reinterpret_cast<const uint32_t*>(rec + ch_off_units(nb_chunk, nb_unit));
#pragma unroll 1
        for (int u = (w + j) % FL_NB; u < nb_unit; u += FL_NB) {
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
      if (staged) phase_b(S.blob[j % 3]);
      else phase_b(rec_g);
      __syncwarp();
      if (lane == 0) fl_arrive(bar_b(j));
    }
  }
  else {
    // =============================== row warps (phase C) ===============================
    const int w = warp - FL_NL - FL_NA - FL_NB;
    for (int j = 0; j < n; ++j) {
      fl_wait(bar_fb(j), par3(j), A.error, 9, A.prof != 0);
      const SliceDesc& d = S.desc_b[j % 3];
      const int nb_unit = d.nb_unit, nb_chunk = d.nb_chunk, R = d.nb_row;
      const bool staged = d.blob_bytes <= A.stage_max;
      const unsigned char* rec_g = A.blob + (size_t)d.blob_off * 16;
      fl_wait(bar_b(j), par3(j), A.error, 10, A.prof != 0);
      double* vcur = S.vout[j % 3];
      int32_t* rowbeg = S.rowbeg[j % 3];
      auto phase_c = [&](const unsigned char* rec) {
        const uint32_t* rowinfo = reinterpret_cast<const uint32_t*>(rec + ch_off_rowinfo(nb_chunk, nb_unit));
        const uint8_t* erow = rec + ch_off_erow(nb_chunk, nb_unit, R);
        const int rw = (R + FL_NC - 1) / FL_NC;
        const int r0 = min(w * rw, R), r1 = min(r0 + rw, R);
        // step 1, one lane per row: the diagonal; the row's destination offset
#pragma unroll 1
        for (int i = r0 + lane; i < r1; i += 32) {
          const uint32_t ri = rowinfo[i];
          const int e0 = rowinfo_erow(ri), e1 = rowinfo_erow(rowinfo[i + 1]);
          const int ed = e0 + rowinfo_pdiag(ri);
          rowbeg[i] -= e0;
          if (rowinfo_own(ri)) {
            double s0 = 0.0, s1 = 0.0;
            int e = e0;
            for (; e + 1 < e1; e += 2) {
              const double v0 = vcur[e], v1 = vcur[e + 1];
              s0 += e != ed ? v0 : 0.0;
              s1 += e + 1 != ed ? v1 : 0.0;
            }
            if (e < e1) s0 += e != ed ? vcur[e] : 0.0;
            vcur[ed] = -(s0 + s1);
          }
          else { // rows of non-owned nodes stay zero (the isOwn gate of the reference)
            for (int e = e0; e < e1; ++e) vcur[e] = 0.0;
          }
        }
        __syncwarp();
        // step 2, one lane per entry: the warp's rows leave shared memory as contiguous stores
        const int eb = rowinfo_erow(rowinfo[r0]), ee = rowinfo_erow(rowinfo[r1]);
#pragma unroll 1
        for (int e = eb + lane; e < ee; e += 128) { // four entries per lane in flight (row table -> row offset -> store is a dependent chain)
          int64_t off[4];
          double v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int eq = min(e + 32 * q, ee - 1);
            off[q] = (int64_t)rowbeg[erow[eq]] + eq;
            v[q] = vcur[eq];
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (e + 32 * q < ee) {
              double* dst = A.values + off[q];
              if (A.accumulate) *dst += v[q]; else *dst = v[q];
            }
          }
        }
      };
      if (staged) phase_c(S.blob[j % 3]);
      else phase_c(rec_g);
      __syncwarp();
      if (lane == 0) fl_arrive(bar_c(j));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

int flow_assemble(afb_ctx* ctx, const ElemParams& prm, int flags, int accumulate)
{
  using G = ChainGeomF;
  const int mode = flags & (AFB_FLAG_ALL_ROWS | AFB_FLAG_OWN_CELLS_ONLY);
  static_assert(sizeof(FlowSmem<G, 4>) <= 232448 - 1024, "the pipelined executor must fit one SM (227 KB per CTA)");
  if (!chain_plan_valid(ctx, mode, 2)) AFB_TRY(chain_build(ctx, mode, 2, chain_limits<G>(), ctx->sm_count));
  ChainPlan& P = *static_cast<ChainPlan*>(ctx->chain);
  if (accumulate) AFB_TRY(ensure_values_zeroed(ctx));
  else ctx->values_dirty = false;
  if (P.nb_slice == 0) return AFB_OK;
  FlowArgs A;
  A.desc = P.desc_exec.as<SliceDesc>();
  A.order = P.order.as<int32_t>();
  A.cta_ptr = P.cta_ptr.as<int32_t>();
  A.coords = ctx->coords.as<double>();
  A.foot = P.foot.as<int32_t>();
  A.lconn = P.lconn.as<uint2>();
  A.slice_nodes = P.slice_nodes.as<int32_t>();
  A.rows = ctx->rows.as<int32_t>();
  A.blob = P.blob.as<unsigned char>();
  A.iblock = P.iblock.as<unsigned char>();
  A.ib_off = P.ib_off.as<int32_t>();
  A.values = ctx->values.as<double>();
  A.error = P.errflag.as<int>();
  static const int prof = [] {
    const char* e = getenv("AFB_FLOW_PROF");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  A.prof = prof;
  A.accumulate = accumulate;
  A.stage_max = (int)std::min<int64_t>(G::BLOB, ctx->tiled_stage_limit);
  auto go = [&](auto kernel, size_t smem) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<P.grid, G::THREADS, smem, ctx->stream>>>(A, prm);
    return cudaGetLastError();
  };
  cudaError_t e = ctx->npc == 4 ? go(k_assemble_flow<G, 4>, sizeof(FlowSmem<G, 4>)) : go(k_assemble_flow<G, 3>, sizeof(FlowSmem<G, 3>));
  AFB_CUDA(e);
  ctx->launches++;
  if (prof) { // development aid: where the pipeline waits (cycles summed over warps and CTAs of this launch)
    unsigned long long h[17];
    AFB_CUDA(cudaStreamSynchronize(ctx->stream));
    AFB_CUDA(cudaMemcpy(h, P.errflag.p, sizeof(h), cudaMemcpyDeviceToHost));
    static const char* site[] = { "", "loader<-A(j-2)", "loader<-C(j-3)", "A<-full", "A<-B(j-1)", "A<-B(j-2)", "B<-full", "B<-A", "B<-C(j-2)", "C<-full", "C<-B", "loader<-iblock" };
    fprintf(stderr, "[flow] waits (Mcycles over all warps):");
    for (int k = 1; k <= 11; ++k) fprintf(stderr, " %s=%.1f", site[k], 1e-6 * (double)h[k]);
    fprintf(stderr, "\n");
    AFB_CUDA(cudaMemset(reinterpret_cast<char*>(P.errflag.p) + 8, 0, sizeof(h) - 8));
  }
  return AFB_OK;
}

} // namespace afb
