// Chained slices: the tiling of the scalar tiled-gather executor (chain_plan.cu / chain_exec.cu).
//
// The brick tiling of tiles_plan.cu recomputes every cell that touches a tile but is owned by a
// neighbouring tile (halo): factor 1.77 on 125-row bricks, and the element phase is fp64- and
// shared-memory-bound.  Here the rows are cut into *columns* (bins across two axes) that are swept
// along the third: a column is a chain of thin slices (about one layer of nodes each), consecutive
// slices are assembled back to back by the same CTA, and the element matrices of the cells two
// consecutive slices share stay in shared memory (two cache regions used alternately: the cells a
// slice computes go to region `parity`, the cells it inherits sit in the other one).  Only the
// lateral halo of the column is recomputed: ((a+1)/a)^2 for an a x a column, 1.23 at a = 9.
// Matrix entries between consecutive slices are summed once, by the earlier slice, which also
// deposits the value in the next slice's staging buffer (the element matrices are symmetric).
//
// Chains are cut in segments (a segment starts with a slice that computes all its cells); segments
// are the unit of load balancing over the persistent CTAs.
#pragma once

#include "afb_internal.h"

namespace afb {

// executor geometry ------------------------------------------------------------------------------
// CN      cells a slice may compute (= size of one cache region); a segment's first slice may compute 2*CN
// RMAX    rows per slice;  EMAX matrix entries per slice;  FMAX footprint nodes of the computed cells
// BLOB    bytes of the slice's plan record staged through the TMA engine (lists, entry map, unit and row tables)
// NREG    cache regions: 2 for the phase-separated executor (a slice's cells replace those of two slices ago), 3 for the
//         pipelined executor (the element phase runs one slice ahead of the gather phase)
template <int THREADS_, int MINB_, int CN_, int RMAX_, int EMAX_, int FMAX_, int BLOB_, int RTARGET3_, int RTARGET2_, int NREG_ = 2>
struct ChainGeom {
  static constexpr int THREADS = THREADS_, MINB = MINB_, CN = CN_, RMAX = RMAX_, EMAX = EMAX_, FMAX = FMAX_, BLOB = BLOB_;
  static constexpr int RT3 = RTARGET3_, RT2 = RTARGET2_, NREG = NREG_;
  static constexpr int CS = CN + 1;              // region stride inside a plane (odd: conflict-free plane offsets)
  static constexpr int PLANE = NREG * CS;        // one plane holds the regions back to back
  static constexpr int ROUNDS = (CN + THREADS - 1) / THREADS;
  static_assert(BLOB % 16 == 0, "TMA bulk copies move multiples of 16 bytes");
};
// A: 3 CTAs/SM x 256 threads, 6x6 columns;  B: 2 CTAs/SM x 384 threads, 8x8 columns (one more node per side fits: bins are ragged on jittered meshes)
// F: the pipelined executor, 1 CTA/SM, warp-specialised (chain_flow.cu)
using ChainGeomA = ChainGeom<256, 3, 384, 96, 1024, 256, 10240, 36, 96>;
using ChainGeomB = ChainGeom<384, 2, 608, 128, 1536, 384, 16384, 64, 128>;
#ifndef AFB_FL_THREADS
#define AFB_FL_THREADS 896
#endif
using ChainGeomF = ChainGeom<AFB_FL_THREADS, 1, 608, 128, 1536, 448, 16384, 64, 128, 3>;

struct ChainLimits { // what the plan builder needs to know about the executor it builds for
  int threads, minb, cn, rmax, emax, fmax, blob, rt3, rt2, nreg;
};
template <class G> inline ChainLimits chain_limits() { return { G::THREADS, G::MINB, G::CN, G::RMAX, G::EMAX, G::FMAX, G::BLOB, G::RT3, G::RT2, G::NREG }; }

// cache regions of a slice with index `sidx` inside its segment: where its computed cells go, where the cells inherited from
// the previous slice sit, and where a segment's first slice puts the cells that die with it (group B)
__host__ __device__ __forceinline__ int ch_reg_new(int nreg, int sidx) { return nreg == 2 ? (sidx & 1) : sidx % 3; }
__host__ __device__ __forceinline__ int ch_reg_prev(int nreg, int sidx) { return nreg == 2 ? 1 - (sidx & 1) : (sidx + 2) % 3; }
__host__ __device__ __forceinline__ int ch_reg_b(int nreg) { return nreg == 2 ? 1 : 2; }

constexpr int CHAIN_SEG_MAX = 32;          // slices per segment (upper bound; fewer on small meshes for load balance)
constexpr unsigned CH_NONE16 = 0xFFFFu;
constexpr unsigned CH_NEXT = 0x8000u;      // entry-map flag: the mirror entry lives in the NEXT slice's staging buffer

enum : int32_t { CH_FLAG_PARITY = 1, CH_FLAG_FIRST = 2, CH_FLAG_LAST = 4, CH_FLAG_SIDX_SHIFT = 8 }; // flags >> 8 = index of the slice inside its segment

struct SliceDesc {
  int32_t node_off, nb_row;    // rows: slice_nodes[node_off, node_off + nb_row) (ascending node ids)
  int32_t cell_off, nb_new;    // cells the slice computes: lconn[cell_off, cell_off + nb_new)
  int32_t foot_off, nb_foot;   // footprint of those cells: foot[foot_off, ...) (ascending node ids)
  uint32_t blob_off;           // plan record, in 16-byte units
  int32_t blob_bytes;          // bytes to stage (multiple of 16)
  int32_t nb_entry;            // matrix entries of the slice's rows
  int32_t nb_unit;             // units of 32 computed entries
  int32_t nb_chunk;            // list rows (32 lanes x 4 contributions each)
  int32_t flags;               // CH_FLAG_*
  int32_t nb_cell;             // cells touching the slice (computed + inherited)
  int32_t nb_a;                // computed cells stored in the slice's own region (the rest of a segment's first slice goes to the other one)
  int32_t blob_cap;            // capacity reserved for the record (bytes)
  int32_t max_val;             // largest node valence in the slice
};
static_assert(sizeof(SliceDesc) == 64, "SliceDesc is copied as 16 words");

// record layout (offsets in bytes from the record's start; every part is 16-byte aligned)
//   lists   uint2[nb_chunk][32]       4 cache indices (16 bit) per lane and list row
//   emap    uint32[nb_unit][32]       own entry | mirror << 16 (CH_NEXT: in the next slice's buffer); 0xFFFFFFFF = padding lane
//   units   uint32[nb_unit]           first list row << 8 | list rows
//   rowinfo uint32[nb_row + 1]        first entry | diagonal position << 16 | owned << 31 (sentinel: nb_entry)
//   erow    uint8[nb_entry]           row (inside the slice) of every entry
__host__ __device__ __forceinline__ int ch_align16(int x) { return (x + 15) & ~15; }
// index block of a slice: descriptor | footprint node ids | row node ids
__host__ __device__ __forceinline__ int ch_ib_foot() { return 64; }
__host__ __device__ __forceinline__ int ch_ib_nodes(int nb_foot) { return 64 + ch_align16(4 * nb_foot); }
__host__ __device__ __forceinline__ int ch_ib_bytes(int nb_foot, int nb_row) { return ch_ib_nodes(nb_foot) + ch_align16(4 * nb_row); }
__host__ __device__ __forceinline__ int ch_off_emap(int nb_chunk) { return nb_chunk * 256; }
__host__ __device__ __forceinline__ int ch_off_units(int nb_chunk, int nb_unit) { return nb_chunk * 256 + nb_unit * 128; }
__host__ __device__ __forceinline__ int ch_off_rowinfo(int nb_chunk, int nb_unit) { return ch_off_units(nb_chunk, nb_unit) + ch_align16(4 * nb_unit); }
__host__ __device__ __forceinline__ int ch_off_erow(int nb_chunk, int nb_unit, int nb_row) { return ch_off_rowinfo(nb_chunk, nb_unit) + ch_align16(4 * (nb_row + 1)); }
__host__ __device__ __forceinline__ int ch_blob_bytes(int nb_chunk, int nb_unit, int nb_row, int nb_entry) { return ch_off_erow(nb_chunk, nb_unit, nb_row) + ch_align16(nb_entry); }

struct ChainPlan {
  bool valid = false;
  uint64_t mesh_gen = ~0ull;
  int mode = 0;                 // ownership flags the lists were built for
  int geom = 0;                 // 0 = A, 1 = B, 2 = F
  int grid = 0;                 // CTAs the schedule was built for
  int32_t nb_slice = 0, nb_seg = 0;
  int64_t nb_new_total = 0, nb_foot_total = 0, blob_units = 0;
  float plan_ms = 0.f;
  double halo = 0.0;            // computed cells / cells
  DevBuf desc;          // SliceDesc[nb_slice]
  DevBuf desc_exec;     // the same descriptors in execution order (order[])
  DevBuf iblock;        // per slice, in execution order: descriptor (64 B) | footprint node ids | row node ids (16-byte padded parts)
  DevBuf ib_off;        // int32[nb_slice + 1]: start of every index block, in 16-byte units
  DevBuf slice_nodes;   // int32[nb_node]: rows of each slice, concatenated
  DevBuf node_slice;    // int32[nb_node]
  DevBuf node_lrow;     // int32[nb_node]: row index inside its slice
  DevBuf node_e0;       // int32[nb_node]: first entry of the node's row inside its slice
  DevBuf new_cells;     // int32: global ids of the computed cells of each slice (slot order)
  DevBuf lconn;         // ushort4 per computed cell
  DevBuf foot;          // int32: footprints
  DevBuf blob;          // plan records
  DevBuf order;         // int32: slices in execution order, per CTA [cta_ptr[c], cta_ptr[c+1])
  DevBuf cta_ptr;       // int32[grid + 1]
  DevBuf errflag;       // int: pipeline time-out code of the executor (0 = fine)
  DevBuf scratch_a, scratch_b, scratch_c, scratch_d, sort_tmp;
};

struct ElemParams;
int chain_assemble(afb_ctx* ctx, const ElemParams& prm, int flags, int accumulate);
int flow_assemble(afb_ctx* ctx, const ElemParams& prm, int flags, int accumulate);
int chain_build(afb_ctx* ctx, int mode_flags, int geom, const ChainLimits& L, int grid);
bool chain_plan_valid(const afb_ctx* ctx, int mode, int geom);
void chain_destroy(afb_ctx* ctx);

} // namespace afb
