// Internal declarations shared by the CUDA translation units of libafb200.so.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <algorithm>
#include <initializer_list>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "afb200.h"

namespace afb {

void set_error(const char* fmt, ...);

#define AFB_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      afb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return AFB_ERR_CUDA;                                                                      \
    }                                                                                           \
  } while (0)

#define AFB_TRY(expr)          \
  do {                         \
    int _rc = (expr);          \
    if (_rc != AFB_OK) return _rc; \
  } while (0)

#define AFB_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      afb::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

// Grow-only device buffer (no cudaMalloc/cudaFree in steady state: pattern rebuilds and
// re-assemblies of the same mesh reuse their memory).
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool owned = true;
  int reserve(size_t bytes)
  {
    if (bytes <= cap && p) return AFB_OK;
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
      p = nullptr;
      return AFB_ERR_CUDA;
    }
    cap = bytes;
    return AFB_OK;
  }
  void alias(const void* ext, size_t bytes)
  {
    if (p && owned) cudaFree(p);
    p = const_cast<void*>(ext);
    cap = bytes;
    owned = false;
  }
  void release()
  {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

// Several buffers that are (re)built together share one allocation: a cudaMalloc costs ~0.3 ms whatever its size, and the
// inspector needs about twenty buffers per mesh.  The members become non-owning views into `arena` (256-byte aligned).
struct GroupItem {
  DevBuf* buf;
  size_t bytes;
};
inline int reserve_group(DevBuf& arena, std::initializer_list<GroupItem> items)
{
  size_t total = 0;
  for (const GroupItem& it : items) total += (std::max<size_t>(it.bytes, 16) + 255) & ~(size_t)255;
  int rc = arena.reserve(total);
  if (rc != AFB_OK) return rc;
  size_t off = 0;
  for (const GroupItem& it : items) {
    const size_t sz = (std::max<size_t>(it.bytes, 16) + 255) & ~(size_t)255;
    it.buf->alias(static_cast<char*>(arena.p) + off, sz);
    off += sz;
  }
  return AFB_OK;
}

// Plan of the tiled path (inspector output, tiles_plan.cu).  The mesh tiling is a pure function
// of the mesh topology and coordinates; the value plan additionally depends on the pattern, the
// block size and the ownership mode.  Both survive pattern re-builds on the same mesh.
struct TilePlan {
  // ---- mesh tiling ----
  bool mesh_valid = false;
  uint64_t mesh_gen = ~0ull;
  int mesh_b_class = 0;        // 0: scalar limits (TG_CMAX), 1: vector limits (TV_CMAX), 2: row-ordered vector executor (VR_CMAX)
  int32_t nb_tile = 0;
  int max_rows = 0;
  int64_t nb_tile_cell = 0, nb_foot = 0, nb_inc = 0, nb_entry = 0;
  float mesh_ms = 0.f, lists_ms = 0.f;
  DevBuf tile_desc;   // TileDesc[nb_tile]
  DevBuf tile_nodes;  // int32: rows (node ids, ascending) of each tile, concatenated
  DevBuf tile_cells;  // int32: global ids of the cells touching each tile (ascending), concatenated
  DevBuf foot;        // int32: footprint nodes of each tile (ascending ids), concatenated
  DevBuf lconn;       // ushort4 per tile cell: footprint-local node indices
  DevBuf rowf;        // uint16 per tile row: footprint index of the row's node
  DevBuf inc;         // uint32 per (row, incident cell): the cell's other nodes, 3 x 10-bit footprint indices
  DevBuf inc_grp;     // uint2 per (tile, group of 32 rows): offset (words, inside the tile) and list length
  DevBuf node_tile;   // int32[nb_node]: tile of each node
  DevBuf node_lrow;   // int32[nb_node]: row index inside its tile
  // ---- tile-local node-node connectivity (connectivity-based BuildMatrix, pattern_tiled.cu) ----
  bool nn_valid = false;
  uint64_t nn_mesh_gen = ~0ull;
  DevBuf nn_deg;      // int32[nb_node]: neighbours + 1 of each node
  DevBuf nn_e0;       // uint16[nb_node] (tile-row order): first tile-local entry of each tile row
  DevBuf nn_local;    // uint16[nb_entry]: per tile row its neighbours (self included) as ascending footprint indices
  std::vector<int32_t> hdesc_host; // host copy of tile_desc (16 words per tile)
  // ---- value plan ----
  bool lists_valid = false;
  uint64_t lists_mesh_gen = ~0ull;
  int lists_b = 0, lists_mode = 0, lists_kind = 0; // kind 0: units of equally long lists (+ mirrors); 1: row-ordered units (vr_units)
  int64_t nb_unit = 0, nb_list = 0;
  DevBuf rowinfo;     // uint32 per tile row: first entry | diagonal position << 16 | owned << 31
  DevBuf unit_base;   // uint32 per unit: offset (16-bit slots) of the unit's index slab inside the tile's list region
  DevBuf unit_len;    // uint16 per unit: contributions per entry of the unit (even)
  DevBuf emap;        // uint32 per (unit, lane): tile-local entry | mirror entry << 16; 0xFFFFFFFF = padding lane
  DevBuf emap_rows;   // uint32 per (unit, lane), vector plans only: tile row of the entry | tile row of the mirror << 16
  DevBuf vr_units;    // uint2 per unit of a row-ordered plan (tiles.cuh: vr_pack_unit)
  DevBuf lists;       // uint16: cache indices, per unit [len/2][32 lanes][2]; one contiguous region per tile (TMA bulk copy)
  DevBuf col_scratch; // int32[nb_entry]: columns in tile order between the two BuildMatrix passes (pattern_tiled.cu)
  DevBuf scratch_a, scratch_b, scratch_c, stats; // builder scratch
  DevBuf arena_nodes, arena_desc, arena_tiles, arena_nn, arena_lists; // shared allocations of the buffers above (reserve_group)
};

} // namespace afb

struct afb_xplan;
void xplan_detach(afb_xplan* x); // mgpu.cu: the plan's context is going away

struct afb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;

  // mesh
  int dim = 0, npc = 0;
  int32_t nb_node = 0, nb_own_node = 0;
  int64_t nb_cell = 0;
  int64_t nb_own_cell = 0;      // cells [0, nb_own_cell) belong to this sub-domain, the rest are ghost cells
  bool has_mesh = false;
  afb::DevBuf coords, conn, is_own; // double[nb_node*3], int32[nb_cell*npc], uint8[nb_node] (may be null => all owned)
  afb::DevBuf cell_coef;            // double[nb_cell]: per-cell multiplier of the Poisson element matrix (afb_set_cell_coefficient)
  bool has_cell_coef = false;
  bool all_own = true;
  afb::DevBuf nc_ptr, nc_list;      // node -> cells (int32[nb_node+1], int32[nb_cell*npc]), ascending cell ids
  int max_valence = 0;

  // pattern
  bool has_pattern = false;
  int b = 1;
  int64_t nnz = 0; // block nnz
  afb::DevBuf rows, cols, nz_per_row, coo_rows; // int32
  bool coo_rows_valid = false;
  afb::DevBuf values; // double[nnz*b*b]
  bool values_dirty = false; // allocated but not yet zeroed (a fresh tiled assembly overwrites every entry)
  afb::DevBuf rhs;    // double[nb_node*b]
  int layout = AFB_LAYOUT_PER_BLOCK;
  bool assembled = false;
  bool values_touched = false; // something other than an assembly wrote into `values` since the last reset / pattern build
  // expanded scalar CSR of a b>1 matrix (BSRMatrix::toCsr)
  afb::DevBuf csr_rows, csr_cols, csr_nbcol;
  afb::DevBuf ij_rows, ij_cols;  // afb_get_ij_arrays: global row numbers, columns in the solver's numbering
  bool csr_valid = false;

  // Dirichlet state
  afb::DevBuf dir_node;                     // uint8[nb_node]
  bool has_dir_nodes = false;
  afb::DevBuf elim_info, elim_value;        // uint8[nb_dof], double[nb_dof]
  afb::DevBuf forced_info, forced_value;    // uint8[nb_dof], double[nb_dof]
  bool has_elim = false, has_forced = false, has_rc = false;
  afb::DevBuf saved_values;                 // pre-elimination copy for the RC RHS correction
  bool saved_valid = false;

  // scratch
  afb::DevBuf tmp_i32a, tmp_i32b, tmp_scan, tmp_ids, tmp_vals, tmp_flag, tmp_lookback, scan_state, solver_work;
  uint64_t mesh_gen = 0;        // bumped by afb_set_mesh / afb_mesh_generate_box
  uint64_t pattern_mesh_gen = ~0ull; // mesh generation the column buffer was last sized for

  afb::TilePlan plan;

  // deferred host-side check of a steady-state pattern re-build (connectivity.cu: verify_pending)
  int32_t* pin_check = nullptr;     // pinned, device-mapped int32[2]: rows[nb_node] and the stale flag of the last re-build
  int32_t* pin_check_dev = nullptr; // its device address (the connectivity-based re-build writes the two words from its kernel: no copy in the stream)
  cudaEvent_t check_event = nullptr;
  bool check_pending = false;
  int sparsity_algo = 0;            // AFB_SPARSITY_*
  int tiled_exec = 0;               // AFB_TILED_EXEC_*
  int vec_exec = 0;                 // AFB_VEC_EXEC_* (afb_set_vector_executor)
  bool vec_rows() const { return vec_exec == 1 /*ROWS*/ || (vec_exec == 0 /*AUTO*/ && npc == 4); }
  int64_t tiled_stage_limit = 1ll << 40; // afb_set_tiled_stage_limit
  void* p2p = nullptr;              // afb::P2PState (p2p.cu): ghost-row exchange over NVLink peer memory
  std::vector<afb_xplan*> xplans;   // exchange plans living on this context (mgpu.cu): detached by afb_destroy
  void* chain = nullptr;            // afb::ChainPlan (chain_plan.cu): plan of the scalar tiled-gather executor
  unsigned scan_tickets = 0, scan_epoch = 0; // chained scan (scan.cu): tiles handed out so far, epoch of the last call
  uint64_t nnz_mesh_gen = ~0ull;    // mesh generation ctx->nnz was last read back for

  // timings
  cudaEvent_t ev[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
  bool timed[3] = { false, false, false };
};

namespace afb {

// ---- scan.cu -------------------------------------------------------------------------------
// out[i] = sum_{j<i} in[j] for i in [0,n]; out has n+1 entries (out[n] = total). in/out may alias
// only if identical pointers are NOT used (separate buffers required).
int exclusive_scan_i32(afb_ctx* ctx, const int32_t* in, int32_t* out, int64_t n);

// ---- connectivity.cu -------------------------------------------------------------------------
int build_node_cells(afb_ctx* ctx);
int build_pattern(afb_ctx* ctx);
int verify_pending(afb_ctx* ctx);

// ---- pattern_rows.cu -------------------------------------------------------------------------
bool pattern_rows_supported(const afb_ctx* ctx);
int pattern_rows_count(afb_ctx* ctx, int32_t* deg);
int pattern_rows_write(afb_ctx* ctx);
int pattern_rows_fused(afb_ctx* ctx, int* exceeded, int32_t* nnz_out);

// ---- assemble.cu -----------------------------------------------------------------------------
int assemble_bilinear(afb_ctx* ctx, int op, const double* params, int format, int variant, int layout, int flags);

// ---- tiles_plan.cu / tiles_exec.cu / pattern_tiled.cu ----------------------------------------
int build_tile_mesh(afb_ctx* ctx, int cls = -1); // cls: TilePlan::mesh_b_class wanted (-1: from the context)
int build_tile_lists(afb_ctx* ctx, int mode_flags);
int build_tile_rowlists(afb_ctx* ctx, int mode_flags);
int assemble_tiled(afb_ctx* ctx, int op, const double* params, int layout, int flags);
bool pattern_tiled_ready(const afb_ctx* ctx);
int rhs_neumann(afb_ctx* ctx, int64_t nb_face, const int32_t* faces_dev, int kind, int nb_value, const double* values, int skip_dirichlet);
int solve_pcg(afb_ctx* ctx, double rtol, double atol, int max_iter, double* x_out, int mem_space, int* iterations, double* residual);
void p2p_destroy(afb_ctx* ctx);
void chain_destroy(afb_ctx* ctx);
float chain_plan_ms(const afb_ctx* ctx);
int p2p_export(afb_ctx* ctx, void* values_handle, void* flags_handle);
int p2p_connect(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles, const int64_t* pull_first,
                const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count);
struct P2PEndpoint { // what a rank tells its neighbours about its arrays (p2p_export_ex)
  unsigned char values_handle[64], flags_handle[64]; // CUDA IPC handles (peers in other processes)
  uint64_t pid, values_ptr, flags_ptr;               // raw pointers (peers in the same process)
  int32_t device, pad;
};
int p2p_export_ex(afb_ctx* ctx, P2PEndpoint* ep);
int p2p_connect_ex(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles, const P2PEndpoint* endpoints,
                   const int64_t* pull_first, const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count);
int p2p_exchange(afb_ctx* ctx, int async);
int p2p_wait(afb_ctx* ctx);
int p2p_status(afb_ctx* ctx, int* status);
int p2p_wait_stats(afb_ctx* ctx, double* ready_wait_us, double* pulled_wait_us, int64_t* nb_exchange);
int p2p_disconnect(afb_ctx* ctx);
bool pattern_nn_ready(const afb_ctx* ctx);
int pattern_nn_build(afb_ctx* ctx);
int pattern_nn_reserve(afb_ctx* ctx);
int pattern_nn_place(afb_ctx* ctx, int32_t* check);
int pattern_tiled_extract(afb_ctx* ctx, int32_t* deg, int* stale);
int pattern_tiled_place(afb_ctx* ctx);
int ensure_values_zeroed(afb_ctx* ctx);

// ---- linear.cu (rhs, dirichlet, views) -------------------------------------------------------
int rhs_source(afb_ctx* ctx, const double* f, int nb_f, int nodewise, int signed_area);
int dirichlet_penalty(afb_ctx* ctx, int weak, double penalty, int32_t n, const int32_t* dof_ids, const double* g);
int apply_matrix_transformation(afb_ctx* ctx, int quirk);
int apply_rhs_transformation(afb_ctx* ctx);
int ensure_coo_rows(afb_ctx* ctx);
int ensure_scalar_csr(afb_ctx* ctx);
int lookup_value_slots(afb_ctx* ctx, int64_t n, const int32_t* dof_rows, const int32_t* dof_cols, int64_t* slots);
int add_values_at(afb_ctx* ctx, int64_t n, const int64_t* slots, const double* contrib);
int renumber_columns(afb_ctx* ctx, const int32_t* dof_local_to_global, int32_t* out);
int fill_iota(afb_ctx* ctx, int32_t first, int32_t n, int32_t* out);
int scatter_flags(afb_ctx* ctx, uint8_t* flags, double* vals, uint8_t flag, int32_t n, const int32_t* ids, const double* v);

// ---- mesh_gen.cu -----------------------------------------------------------------------------
int generate_box(afb_ctx* ctx, int dim, int n, double jitter, uint32_t seed, int k_lo, int k_hi, int ghost_cell_layer);

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// NVTX range over an API phase, named after the reference's time-stats scopes / ProfileRegion labels (BuildMatrix, AddAndCompute:
// modules/testlab/CsrGpuBiliAssembly.cc:313-338).  Header-only NVTX: a no-op unless a profiler is attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Programmatic dependent launch for the kernels of the steady-state step (scan -> column placement -> assembly):
// the grid may be scheduled while the kernel before it on the stream drains, which hides the launch latency between
// dependent kernels (a few microseconds per boundary; 3 boundaries per step).  A kernel launched this way calls
// pdl_enter() before its first global memory access: griddepcontrol.wait returns once the previous grid has completed
// and its writes are visible, so the ordering seen by the kernel body is the stream order.  AFB_NO_PDL=1 launches normally.
bool pdl_enabled();
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

#define AFB_LAUNCH_CHECK(ctx)                                                                 \
  do {                                                                                        \
    (ctx)->launches++;                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess) {                                                                  \
      afb::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return AFB_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

} // namespace afb
