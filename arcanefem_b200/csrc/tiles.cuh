// Shared declarations of the tiled path (inspector: tiles_plan.cu; executors: tiles_exec.cu;
// sparsity pattern from tiles: pattern_tiled.cu).
//
// A *tile* is a spatial brick of matrix rows (nodes) small enough that everything one CTA needs
// to finish those rows lives in shared memory: the coordinates of the tile's footprint (rows +
// halo nodes), the element matrices of every cell touching the tile, and the tile's matrix
// entries.  Two CTAs are resident per SM (113 KB each), so one tile's element phase (fp64 pipe)
// overlaps the other's gather phase (shared-memory pipe) and its HBM traffic.
#pragma once

#include "afb_internal.h"

namespace afb {

// ---- executor geometry (TG_MINB CTAs per SM; overridable for tuning builds) --------------------
#ifndef AFB_TG_THREADS
#define AFB_TG_THREADS 384
#define AFB_TG_MINB 2
#define AFB_TG_CMAX 1408
#define AFB_TG_EMAX 2304
#define AFB_TG_FMAX 384
#define AFB_TG_LMAX 7680
#define AFB_TG_RT3 125
#define AFB_TG_RT2 288
#endif
constexpr int TG_THREADS = AFB_TG_THREADS; // executor CTA
constexpr int TG_MINB = AFB_TG_MINB;       // CTAs per SM
constexpr int TG_CMAX = AFB_TG_CMAX;       // cells per tile
constexpr int TG_CS = TG_CMAX + 1;         // cache plane stride (odd)
constexpr int TG_KP = 6;                   // cached values per cell: the off-diagonal pairs of a 4-node cell
constexpr int TG_ZERO = TG_KP * TG_CS;     // cache slot that holds 0.0 (list padding)
constexpr int TG_EMAX = AFB_TG_EMAX;       // matrix entries per tile
constexpr int TG_FMAX = AFB_TG_FMAX;       // footprint nodes per tile
constexpr int TG_RMAX = TG_FMAX;           // rows per tile
constexpr int TG_LMAX = AFB_TG_LMAX;       // 16-bit list slots per tile (TMA-staged)
constexpr int TG_RT3 = AFB_TG_RT3;         // target rows per tile, 3-D / 2-D
constexpr int TG_RT2 = AFB_TG_RT2;
constexpr int TG_SMEM_LIMIT = (233472 / TG_MINB) - 1024; // 228 KB per SM, 1 KB reserved per CTA
static_assert(TG_THREADS >= TG_FMAX && TG_THREADS >= TG_EMAX / 32, "one thread per footprint node / row / unit");
constexpr int TG_UMAX = TG_EMAX / 32;      // units (32 entries with equally long lists) per tile
constexpr int TG_ROUNDS = (TG_CMAX + TG_THREADS - 1) / TG_THREADS;
constexpr int TG_GMAX = TG_RMAX / 32;      // row groups per tile (incidence lists)
constexpr unsigned TG_NONE16 = 0xFFFFu;

// vector (b = 2, 3) executor: cache of sqrt(s)*grad(phi_a) per cell, blocks accumulated in registers
constexpr int TV_PLANES = 12;              // 4 nodes x 3 components
#ifndef AFB_TV_LMAX
#define AFB_TV_LMAX 10240
#endif
constexpr int TV_LMAX = AFB_TV_LMAX;       // 16-bit list slots of a tile staged through the TMA engine (0: lists read from global memory)
constexpr int TV_CMAX_RAW = (TG_SMEM_LIMIT - 3 * 8 * TG_FMAX - 8 * TG_RMAX - 1024 - 2 * TV_LMAX) / (8 * TV_PLANES) - 1;
constexpr int TV_CMAX = TV_CMAX_RAW < 1024 ? TV_CMAX_RAW : 1024;
constexpr int TV_CS = TV_CMAX + 1;
#ifndef AFB_TV_RT3
#define AFB_TV_RT3 100
#endif
constexpr int TV_RT3 = AFB_TV_RT3 > 8 ? AFB_TV_RT3 : 8; // target rows per tile of the vector executor in 3-D (measured at n=140, b=3:
                                                        // 60 / 75 / 88 / 100 / 112 / 125 / 150 rows -> 2.44 / 2.45 / 2.41 / 2.32 / 2.35 / 2.35 / 2.40 ms;
                                                        // bricks over the cell limit are cut by the per-brick refinement)
#ifndef AFB_TV_RT2
#define AFB_TV_RT2 352
#endif
constexpr int TV_RT2 = AFB_TV_RT2;         // target rows per tile of the vector executor in 2-D (measured: 224 / 288 / 352 -> 2.23 / 2.39 / 2.20 ms at C5)

// row-ordered vector executor (elasticity): lanes = consecutive entries of whole rows, blocks staged per warp and written as
// contiguous runs, the diagonal block derived from the row's off-diagonal blocks (zero block row sums)
constexpr int VR_BSTRIDE3 = 9, VR_BSTRIDE2 = 5;      // doubles per staged block (odd: conflict-free half-warps)
constexpr int VR_UMAX = TG_THREADS;                  // units per tile (one thread stages one unit record)
constexpr int VR_FIXED = 3 * 8 * TG_FMAX + (TG_THREADS / 32) * 304 * 8 + 8 * TG_RMAX + 4 * (TG_RMAX + 1) + 8 * VR_UMAX + 2 * TG_EMAX + 4 * 64 + 64;
constexpr int VR_CMAX_RAW = (TG_SMEM_LIMIT - VR_FIXED) / (8 * TV_PLANES) - 1;
#ifndef AFB_VR_CS_MOD
#define AFB_VR_CS_MOD 1
#endif
// cache plane stride: the largest value that fits with VR_CS % 16 == AFB_VR_CS_MOD (1: consecutive planes shift by one 8-byte
// bank, so the gathers of different nodes of one cell never collide)
constexpr int VR_CS_FIT = (VR_CMAX_RAW < 1023 ? VR_CMAX_RAW : 1023) + 1;
constexpr int VR_CS = VR_CS_FIT - ((VR_CS_FIT - AFB_VR_CS_MOD) & 15);
constexpr int VR_CMAX = VR_CS - 1;
constexpr int VR_LC_BITS = 10;                       // list code = (a * 4 + b) << 10 | cache slot of the cell (a, b: nodes of the cell)
constexpr unsigned VR_LC_MASK = (1u << VR_LC_BITS) - 1u;
static_assert(VR_CS <= (1 << VR_LC_BITS), "cache slots must fit the list code");
constexpr int VR_PSTRIDE3 = 99, VR_PSTRIDE2 = 66;    // per-row value layout: staged as [block row][entry][column], plane stride
constexpr int VR_STAGE = 3 * VR_PSTRIDE3 + 7;        // doubles of staging per warp (>= 32 * VR_BSTRIDE3)
#ifndef AFB_VR_RT3
#define AFB_VR_RT3 80
#endif
#ifndef AFB_VR_RT2
#define AFB_VR_RT2 320
#endif
constexpr int VR_RT3 = AFB_VR_RT3, VR_RT2 = AFB_VR_RT2;
#ifndef AFB_VR_SLACK
#define AFB_VR_SLACK 0
#endif
#ifndef AFB_VR_SHARE
#define AFB_VR_SHARE 1
#endif
constexpr bool VR_SHARE = AFB_VR_SHARE != 0; // list ordering prefers cells another lane of the half-warp reads in the same step
constexpr int VR_SLACK = AFB_VR_SLACK; // extra (padding) steps per unit: room for the plan to dodge shared-memory bank conflicts
// unit record (uint2): x = first 16-bit slot of the unit's lists inside the tile's list region;
// y = first tile-local entry | (entries - 1) << 12 | list length (contributions per lane, even) << 17
__host__ __device__ __forceinline__ uint32_t vr_pack_unit(int first, int cnt, int len) { return (uint32_t)first | ((uint32_t)(cnt - 1) << 12) | ((uint32_t)len << 17); }

struct TileDesc {
  int32_t node_off, nb_row;    // rows (node ids ascending) in tile_nodes / rowinfo
  int32_t cell_off, nb_cell;   // tile_cells / lconn
  int32_t foot_off, nb_foot;   // foot (ascending node ids)
  uint32_t inc_off;            // first word of the tile in `inc`
  int32_t nb_group;            // row groups of 32
  int32_t unit_off, nb_unit;   // unit tables / emap (value plan)
  uint32_t list_off;           // first 16-bit slot of the tile in `lists` (multiple of 8)
  int32_t list_len;            // used slots (multiple of 8)
  int32_t nb_entry;            // matrix entries of the tile's rows
  int32_t max_val;             // largest node valence in the tile
  uint32_t ent_off;            // first entry of the tile in the tile-ordered column scratch (capacity nb_entry)
  int32_t pad1;
};
static_assert(sizeof(TileDesc) == 64, "TileDesc is copied as 16 words");

// off-diagonal pair (a<b) of an NPC-node cell -> cache plane
__host__ __device__ __forceinline__ constexpr int off_pair(int npc, int a, int b)
{
  const int lo = a < b ? a : b, hi = a < b ? b : a;
  return lo * (2 * npc - lo - 1) / 2 + (hi - lo - 1);
}

// rowinfo word of a tile row: first entry (tile-local), position of the diagonal, ownership
__host__ __device__ __forceinline__ uint32_t pack_rowinfo(int erow, int pdiag, bool own) { return (uint32_t)erow | ((uint32_t)pdiag << 16) | (own ? 0x80000000u : 0u); }
__host__ __device__ __forceinline__ int rowinfo_erow(uint32_t w) { return (int)(w & 0xFFFFu); }
__host__ __device__ __forceinline__ int rowinfo_pdiag(uint32_t w) { return (int)((w >> 16) & 0x7FFFu); }
__host__ __device__ __forceinline__ bool rowinfo_own(uint32_t w) { return (w >> 31) != 0; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

} // namespace afb
