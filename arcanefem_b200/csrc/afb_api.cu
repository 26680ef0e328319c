// C ABI entry points of libafb200.so (include/afb200.h).
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "afb_internal.h"

namespace afb {
bool pdl_enabled()
{
  static const bool on = [] { const char* e = getenv("AFB_NO_PDL"); return !(e && *e && *e != '0'); }();
  return on;
}

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static int check_ctx(afb_ctx* ctx)
{
  AFB_REQUIRE(ctx != nullptr, AFB_ERR_INVALID, "null context");
  AFB_CUDA(cudaSetDevice(ctx->device));
  return AFB_OK;
}

static int time_begin(afb_ctx* ctx, int phase)
{
  AFB_CUDA(cudaEventRecord(ctx->ev[2 * phase], ctx->stream));
  return AFB_OK;
}
static int time_end(afb_ctx* ctx, int phase)
{
  AFB_CUDA(cudaEventRecord(ctx->ev[2 * phase + 1], ctx->stream));
  ctx->timed[phase] = true;
  return AFB_OK;
}

static void invalidate_pattern(afb_ctx* ctx)
{
  ctx->has_pattern = false;
  ctx->assembled = false;
  ctx->values_touched = false;
  ctx->coo_rows_valid = false;
  ctx->csr_valid = false;
  ctx->saved_valid = false;
  ctx->has_elim = ctx->has_forced = ctx->has_rc = false;
}

// copy a host or device array into an owned device buffer
static int upload(afb_ctx* ctx, DevBuf& buf, const void* src, size_t bytes, int mem_space)
{
  if (mem_space == AFB_MEM_DEVICE) {
    buf.alias(src, bytes);
    return AFB_OK;
  }
  if (!buf.owned) buf.release();
  AFB_TRY(buf.reserve(bytes));
  if (bytes) AFB_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return AFB_OK;
}

// stage a small host/device id or value list on the device
static int stage(afb_ctx* ctx, DevBuf& buf, const void* src, size_t bytes, int mem_space, const void** dev)
{
  if (mem_space == AFB_MEM_DEVICE) {
    *dev = src;
    return AFB_OK;
  }
  AFB_TRY(buf.reserve(bytes));
  if (bytes) AFB_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  *dev = buf.p;
  return AFB_OK;
}

} // namespace afb

using namespace afb;

extern "C" {

const char* afb_last_error(void) { return g_err; }
const char* afb_version(void) { return "arcanefem_b200 0.1 (sm_100a)"; }

int afb_create(int device, afb_ctx** out)
{
  AFB_REQUIRE(out != nullptr, AFB_ERR_INVALID, "null output pointer");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s): libafb200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return AFB_ERR_CUDA;
  }
  AFB_REQUIRE(device >= 0 && device < count, AFB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  AFB_CUDA(cudaSetDevice(device));
  afb_ctx* ctx = new afb_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  AFB_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  AFB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  for (int i = 0; i < 6; ++i) AFB_CUDA(cudaEventCreate(&ctx->ev[i]));
  AFB_CUDA(cudaEventCreateWithFlags(&ctx->check_event, cudaEventDisableTiming));
  AFB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->pin_check), 2 * sizeof(int32_t), cudaHostAllocMapped));
  AFB_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->pin_check_dev), ctx->pin_check, 0));
  if (const char* ex = getenv("AFB_VEC_EXEC")) ctx->vec_exec = !strcmp(ex, "units") ? AFB_VEC_EXEC_UNITS : !strcmp(ex, "rows") ? AFB_VEC_EXEC_ROWS : AFB_VEC_EXEC_AUTO; // afb_set_vector_executor
  if (const char* ex = getenv("AFB_TILED_EXEC")) { // default executor of the scalar tiled gather (afb_set_tiled_executor)
    if (!strcmp(ex, "chain")) ctx->tiled_exec = AFB_TILED_EXEC_CHAIN;
    else if (!strcmp(ex, "flow")) ctx->tiled_exec = AFB_TILED_EXEC_CHAIN_FLOW;
  }
  *out = ctx;
  return AFB_OK;
}

int afb_destroy(afb_ctx* ctx)
{
  if (!ctx) return AFB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (afb_xplan* x : std::vector<afb_xplan*>(ctx->xplans)) xplan_detach(x); // plans outliving their context must not touch it
  ctx->xplans.clear();
  p2p_destroy(ctx);
  chain_destroy(ctx);
  DevBuf* bufs[] = { &ctx->coords, &ctx->conn, &ctx->is_own, &ctx->nc_ptr, &ctx->nc_list, &ctx->rows, &ctx->cols, &ctx->nz_per_row, &ctx->coo_rows, &ctx->values,
                     &ctx->rhs, &ctx->csr_rows, &ctx->csr_cols, &ctx->csr_nbcol, &ctx->ij_rows, &ctx->ij_cols, &ctx->dir_node, &ctx->elim_info, &ctx->elim_value, &ctx->forced_info,
                     &ctx->forced_value, &ctx->saved_values, &ctx->tmp_i32a, &ctx->tmp_i32b, &ctx->tmp_scan, &ctx->tmp_ids, &ctx->tmp_vals, &ctx->tmp_flag, &ctx->tmp_lookback, &ctx->scan_state, &ctx->solver_work,
                     &ctx->plan.tile_desc, &ctx->plan.tile_nodes, &ctx->plan.tile_cells, &ctx->plan.unit_base, &ctx->plan.unit_len, &ctx->plan.emap, &ctx->plan.rowinfo, &ctx->plan.foot, &ctx->plan.lconn, &ctx->plan.lists,
                     &ctx->plan.rowf, &ctx->plan.inc, &ctx->plan.inc_grp, &ctx->plan.col_scratch, &ctx->plan.nn_deg, &ctx->plan.nn_local, &ctx->plan.nn_e0, &ctx->plan.emap_rows, &ctx->plan.vr_units,
                     &ctx->plan.node_tile, &ctx->plan.node_lrow, &ctx->plan.scratch_a, &ctx->plan.scratch_b, &ctx->plan.scratch_c, &ctx->plan.stats,
                     &ctx->plan.arena_nodes, &ctx->plan.arena_desc, &ctx->plan.arena_tiles, &ctx->plan.arena_nn, &ctx->plan.arena_lists };
  for (DevBuf* b : bufs) b->release();
  for (int i = 0; i < 6; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->check_event) cudaEventDestroy(ctx->check_event);
  if (ctx->pin_check) cudaFreeHost(ctx->pin_check);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return AFB_OK;
}

int afb_set_stream(afb_ctx* ctx, void* cuda_stream)
{
  AFB_TRY(check_ctx(ctx));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  AFB_TRY(verify_pending(ctx));
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return AFB_OK;
}

int afb_options_from_name(const char* name, int* format, int* variant, int* sparsity)
{
  AFB_REQUIRE(name && format && variant && sparsity, AFB_ERR_INVALID, "afb_options_from_name: null argument");
  char low[32];
  size_t n = 0;
  for (; name[n] && n + 1 < sizeof(low); ++n) low[n] = (char)((name[n] >= 'A' && name[n] <= 'Z') ? name[n] - 'A' + 'a' : name[n]);
  low[n] = 0;
  struct Row { const char* name; int format, variant, sparsity; };
  static const Row table[] = {
    // host back-ends of testlab (same matrices as their device twins)
    { "legacy", AFB_FORMAT_CSR, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "dok", AFB_FORMAT_CSR, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "coo", AFB_FORMAT_COO, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "coo-sorting", AFB_FORMAT_COO, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "csr", AFB_FORMAT_CSR, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    // device back-ends
    { "coo-gpu", AFB_FORMAT_COO, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "coo-sorting-gpu", AFB_FORMAT_COO, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "csr-gpu", AFB_FORMAT_CSR, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "nwcsr", AFB_FORMAT_CSR, AFB_VARIANT_TILED_GATHER, AFB_SPARSITY_FROM_CONNECTIVITY },
    { "blcsr", AFB_FORMAT_CSR, AFB_VARIANT_NODEWISE, AFB_SPARSITY_FROM_CONNECTIVITY },
    { "bsr", AFB_FORMAT_BSR, AFB_VARIANT_CELLWISE_ATOMIC, AFB_SPARSITY_FROM_CELLS },
    { "bsr-atomic-free", AFB_FORMAT_BSR, AFB_VARIANT_TILED_GATHER, AFB_SPARSITY_FROM_CONNECTIVITY },
    { "af-bsr", AFB_FORMAT_BSR, AFB_VARIANT_TILED_GATHER, AFB_SPARSITY_FROM_CONNECTIVITY },
  };
  for (const Row& r : table)
    if (!strcmp(low, r.name)) {
      *format = r.format;
      *variant = r.variant;
      *sparsity = r.sparsity;
      return AFB_OK;
    }
  set_error("afb_options_from_name: unknown matrix format option '%s'", name);
  return AFB_ERR_INVALID;
}

int afb_set_tiled_executor(afb_ctx* ctx, int executor)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(executor >= AFB_TILED_EXEC_BRICKS && executor <= AFB_TILED_EXEC_CHAIN_FLOW, AFB_ERR_INVALID, "afb_set_tiled_executor: unknown executor %d", executor);
  ctx->tiled_exec = executor;
  return AFB_OK;
}

int afb_set_vector_executor(afb_ctx* ctx, int executor)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(executor >= AFB_VEC_EXEC_AUTO && executor <= AFB_VEC_EXEC_UNITS, AFB_ERR_INVALID, "afb_set_vector_executor: unknown executor %d", executor);
  ctx->vec_exec = executor;
  return AFB_OK;
}

int afb_set_tiled_stage_limit(afb_ctx* ctx, int64_t bytes)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(bytes >= 0, AFB_ERR_INVALID, "afb_set_tiled_stage_limit: negative limit");
  ctx->tiled_stage_limit = bytes;
  return AFB_OK;
}

int afb_set_sparsity_algorithm(afb_ctx* ctx, int algorithm)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(algorithm >= AFB_SPARSITY_AUTO && algorithm <= AFB_SPARSITY_FROM_CONNECTIVITY, AFB_ERR_INVALID, "afb_set_sparsity_algorithm: unknown algorithm %d", algorithm);
  ctx->sparsity_algo = algorithm;
  return AFB_OK;
}

// ---- ghost-row exchange over NVLink peer memory (p2p.cu) ----
int afb_p2p_export(afb_ctx* ctx, void* values_handle, void* flags_handle)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && values_handle && flags_handle, AFB_ERR_INVALID, "afb_p2p_export: no pattern / null handle buffer");
  return p2p_export(ctx, values_handle, flags_handle);
}

int afb_p2p_connect(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles, const int64_t* pull_first,
                    const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_p2p_connect: no pattern");
  return p2p_connect(ctx, my_rank, nb_peer, peer_rank, values_handles, flags_handles, pull_first, pull_count, slots, send_first, send_count);
}

int afb_p2p_exchange(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  return p2p_exchange(ctx, 0);
}

int afb_p2p_exchange_async(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  return p2p_exchange(ctx, 1);
}

int afb_p2p_wait(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  return p2p_wait(ctx);
}

int afb_p2p_wait_stats(afb_ctx* ctx, double* ready_wait_us, double* pulled_wait_us, int64_t* nb_exchange)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ready_wait_us && pulled_wait_us && nb_exchange, AFB_ERR_INVALID, "afb_p2p_wait_stats: null");
  return p2p_wait_stats(ctx, ready_wait_us, pulled_wait_us, nb_exchange);
}

int afb_p2p_status(afb_ctx* ctx, int* status)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(status, AFB_ERR_INVALID, "afb_p2p_status: null");
  return p2p_status(ctx, status);
}

int afb_p2p_disconnect(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  return p2p_disconnect(ctx);
}

int afb_synchronize(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  AFB_TRY(verify_pending(ctx));
  return AFB_OK;
}

int afb_set_mesh(afb_ctx* ctx, int dim, int npc, int32_t nb_node, int64_t nb_cell, const double* xyz, const int32_t* cell_nodes, const uint8_t* node_is_own, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(dim == 2 || dim == 3, AFB_ERR_UNSUPPORTED, "BSRFormat(initialize): Only supports 2D and 3D (dim=%d)", dim);
  AFB_REQUIRE((dim == 2 && (npc == 3 || npc == 6 || npc == 4)) || (dim == 3 && (npc == 4 || npc == 10 || npc == 8)), AFB_ERR_UNSUPPORTED,
              "unsupported cell type: %d nodes per cell in dimension %d (Tri3 / Tri6 / Quad4, Tet4 / Tet10 / Hexa8)", npc, dim);
  AFB_REQUIRE(nb_node > 0 && nb_cell >= 0, AFB_ERR_INVALID, "bad mesh sizes nb_node=%d nb_cell=%lld", nb_node, (long long)nb_cell);
  AFB_REQUIRE(xyz && (cell_nodes || nb_cell == 0), AFB_ERR_INVALID, "null mesh arrays");
  // device arrays are used in place (zero-copy) and read with vector loads: 16-byte aligned connectivity rows, 8-byte coordinates
  AFB_REQUIRE(mem_space != AFB_MEM_DEVICE || ((reinterpret_cast<uintptr_t>(cell_nodes) & 15u) == 0 && (reinterpret_cast<uintptr_t>(xyz) & 7u) == 0), AFB_ERR_INVALID,
              "afb_set_mesh(AFB_MEM_DEVICE): cell_nodes must be 16-byte aligned and xyz 8-byte aligned (pass a copy, or use AFB_MEM_HOST)");
  invalidate_pattern(ctx);
  ctx->has_mesh = false;
  ctx->has_cell_coef = false; // per-cell data of the previous mesh
  ctx->dim = dim;
  ctx->npc = npc;
  ctx->nb_node = nb_node;
  ctx->nb_cell = nb_cell;
  AFB_TRY(upload(ctx, ctx->coords, xyz, sizeof(double) * 3 * (size_t)nb_node, mem_space));
  AFB_TRY(upload(ctx, ctx->conn, cell_nodes, sizeof(int32_t) * (size_t)npc * (size_t)nb_cell, mem_space));
  ctx->all_own = (node_is_own == nullptr);
  ctx->nb_own_node = nb_node;
  ctx->nb_own_cell = nb_cell;
  if (node_is_own) AFB_TRY(upload(ctx, ctx->is_own, node_is_own, (size_t)nb_node, mem_space));
  if (node_is_own && mem_space == AFB_MEM_HOST) {
    int32_t own = 0;
    for (int32_t i = 0; i < nb_node; ++i) own += node_is_own[i] ? 1 : 0;
    ctx->nb_own_node = own;
  }
  ctx->has_dir_nodes = false;
  ctx->has_mesh = true;
  ctx->mesh_gen++;
  AFB_TRY(time_begin(ctx, 0));
  AFB_TRY(build_node_cells(ctx));
  AFB_TRY(time_end(ctx, 0));
  return AFB_OK;
}

int afb_update_coordinates(afb_ctx* ctx, const double* xyz, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh && xyz, AFB_ERR_INVALID, "afb_update_coordinates: no mesh / null coordinates");
  const size_t bytes = sizeof(double) * 3 * (size_t)ctx->nb_node;
  if (mem_space == AFB_MEM_DEVICE) {
    if (xyz != ctx->coords.p) {
      if (ctx->coords.owned) AFB_CUDA(cudaMemcpyAsync(ctx->coords.p, xyz, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
      else ctx->coords.alias(xyz, bytes);
    }
  }
  else {
    if (!ctx->coords.owned) { // the context aliased the caller's device array so far: own a copy from now on
      ctx->coords.release();
      AFB_TRY(ctx->coords.reserve(bytes));
    }
    AFB_CUDA(cudaMemcpyAsync(ctx->coords.p, xyz, bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  return AFB_OK;
}

int afb_set_cell_coefficient(afb_ctx* ctx, const double* coefficient, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh, AFB_ERR_INVALID, "afb_set_cell_coefficient: no mesh set");
  if (!coefficient) {
    ctx->has_cell_coef = false;
    return AFB_OK;
  }
  AFB_REQUIRE(mem_space == AFB_MEM_HOST || mem_space == AFB_MEM_DEVICE, AFB_ERR_INVALID, "afb_set_cell_coefficient: unknown memory space %d", mem_space);
  AFB_TRY(upload(ctx, ctx->cell_coef, coefficient, sizeof(double) * (size_t)ctx->nb_cell, mem_space));
  ctx->has_cell_coef = true;
  return AFB_OK;
}

int afb_set_own_cell_count(afb_ctx* ctx, int64_t nb_own_cell)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh, AFB_ERR_INVALID, "afb_set_own_cell_count: no mesh");
  AFB_REQUIRE(nb_own_cell >= 0 && nb_own_cell <= ctx->nb_cell, AFB_ERR_INVALID, "nb_own_cell %lld out of range [0,%lld]", (long long)nb_own_cell, (long long)ctx->nb_cell);
  ctx->nb_own_cell = nb_own_cell;
  ctx->plan.lists_valid = false;
  return AFB_OK;
}

int afb_get_own_cell_count(afb_ctx* ctx, int64_t* nb_own_cell)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh && nb_own_cell, AFB_ERR_INVALID, "afb_get_own_cell_count: no mesh / null output");
  *nb_own_cell = ctx->nb_own_cell;
  return AFB_OK;
}

int afb_mesh_generate_box(afb_ctx* ctx, int dim, int n, double jitter, uint32_t seed, int k_lo, int k_hi, int ghost_cell_layer)
{
  AFB_TRY(check_ctx(ctx));
  invalidate_pattern(ctx);
  ctx->has_mesh = false;
  if (!ctx->coords.owned) ctx->coords.release();
  if (!ctx->conn.owned) ctx->conn.release();
  if (!ctx->is_own.owned) ctx->is_own.release();
  AFB_TRY(generate_box(ctx, dim, n, jitter, seed, k_lo, k_hi, ghost_cell_layer));
  ctx->mesh_gen++;
  ctx->has_dir_nodes = false;
  AFB_TRY(time_begin(ctx, 0));
  AFB_TRY(build_node_cells(ctx));
  AFB_TRY(time_end(ctx, 0));
  return AFB_OK;
}

int afb_build_pattern(afb_ctx* ctx, int nb_dof_per_node, int32_t* nb_block_row, int64_t* nb_block_nnz)
{
  afb::NvtxRange nvtx_range("BuildMatrix");
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh, AFB_ERR_INVALID, "afb_build_pattern: no mesh set");
  // BSRMatrix::initialize argument checks (femutils/BSRFormat.cc:51-55)
  AFB_REQUIRE(nb_dof_per_node >= 1 && nb_dof_per_node <= 3, AFB_ERR_INVALID, "BSRMatrix(initialize): block_size must be 1, 2 or 3 (got %d)", nb_dof_per_node);
  // a mesh that was assembled through the chained-slice executor and is built again: create the init-time node-node
  // connectivity of the connectivity-based BuildMatrix now (once per mesh; the first assembly did not need it)
  if (ctx->has_pattern && ctx->b == 1 && nb_dof_per_node == 1 && ctx->pattern_mesh_gen == ctx->mesh_gen && ctx->sparsity_algo != AFB_SPARSITY_FROM_CELLS &&
      chain_plan_ms(ctx) >= 0.0f && !pattern_nn_ready(ctx) && ctx->npc == ctx->dim + 1)
    AFB_TRY(build_tile_mesh(ctx));
  invalidate_pattern(ctx);
  ctx->b = nb_dof_per_node;
  AFB_TRY(time_begin(ctx, 1));
  AFB_TRY(build_pattern(ctx));
  AFB_TRY(time_end(ctx, 1));
  ctx->has_pattern = true;
  // explicit connectivity-based sparsity: the init-time connectivity is created right after the first build
  // of a mesh (AUTO creates it with the first tiled assembly)
  if (ctx->sparsity_algo == AFB_SPARSITY_FROM_CONNECTIVITY && !pattern_nn_ready(ctx) && ctx->npc == ctx->dim + 1) AFB_TRY(build_tile_mesh(ctx));
  if (nb_block_row) *nb_block_row = ctx->nb_node;
  if (nb_block_nnz) *nb_block_nnz = ctx->nnz;
  return AFB_OK;
}

int afb_reset_values(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_reset_values: no pattern");
  AFB_CUDA(cudaMemsetAsync(ctx->values.p, 0, sizeof(double) * (size_t)ctx->nnz * ctx->b * ctx->b, ctx->stream));
  ctx->values_dirty = false;
  ctx->assembled = false;
  ctx->values_touched = false;
  ctx->saved_valid = false;
  return AFB_OK;
}

int afb_assemble_bilinear(afb_ctx* ctx, int op, const double* params, int nb_params, int format, int variant, int value_layout, int flags)
{
  afb::NvtxRange nvtx_range("AddAndCompute");
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_assemble_bilinear: build the pattern first");
  AFB_REQUIRE(format == AFB_FORMAT_CSR || format == AFB_FORMAT_COO || format == AFB_FORMAT_BSR, AFB_ERR_INVALID, "unknown matrix format %d", format);
  AFB_REQUIRE(value_layout == AFB_LAYOUT_PER_BLOCK || value_layout == AFB_LAYOUT_PER_ROW, AFB_ERR_INVALID, "unknown value layout %d", value_layout);
  const int need_b = (op == AFB_OP_POISSON || op == AFB_OP_DIFFUSION_REACTION) ? 1 : ((op == AFB_OP_ELASTICITY || op == AFB_OP_ELASTODYNAMICS) ? ctx->dim : (op == AFB_OP_BILAPLACIAN ? 2 : -1));
  AFB_REQUIRE(need_b > 0, AFB_ERR_INVALID, "unknown operator %d", op);
  AFB_REQUIRE(need_b == ctx->b, AFB_ERR_INVALID, "operator %d needs %d dof per node, pattern was built with %d", op, need_b, ctx->b);
  AFB_REQUIRE(op != AFB_OP_ELASTICITY || (params && nb_params >= 2), AFB_ERR_INVALID, "elasticity needs params = {lambda, mu}");
  AFB_REQUIRE(op != AFB_OP_DIFFUSION_REACTION || (params && nb_params >= 2), AFB_ERR_INVALID, "diffusion-reaction needs params = {alpha, beta}");
  AFB_REQUIRE(op != AFB_OP_ELASTODYNAMICS || (params && nb_params >= 3), AFB_ERR_INVALID, "elastodynamics needs params = {c0, c1, c2}");
  AFB_REQUIRE(!(ctx->b > 1 && format != AFB_FORMAT_BSR), AFB_ERR_INVALID, "CSR/COO back-ends hold one dof per node; use AFB_FORMAT_BSR for b=%d", ctx->b);
  AFB_REQUIRE(!(ctx->assembled && ctx->layout != value_layout), AFB_ERR_INVALID, "values already hold the other layout; afb_reset_values first");
  ctx->layout = value_layout;
  AFB_TRY(time_begin(ctx, 2));
  if (variant == AFB_VARIANT_TILED_GATHER) AFB_TRY(assemble_tiled(ctx, op, params, value_layout, flags));
  else {
    AFB_TRY(ensure_values_zeroed(ctx));
    AFB_TRY(assemble_bilinear(ctx, op, params, format, variant, value_layout, flags));
  }
  AFB_TRY(time_end(ctx, 2));
  ctx->assembled = true;
  return AFB_OK;
}

int afb_rhs_reset(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_rhs_reset: no pattern");
  AFB_CUDA(cudaMemsetAsync(ctx->rhs.p, 0, sizeof(double) * (size_t)ctx->nb_node * ctx->b, ctx->stream));
  return AFB_OK;
}

int afb_assemble_rhs_source(afb_ctx* ctx, const double* f, int nb_f, int nodewise, int signed_tri_area)
{
  afb::NvtxRange nvtx_range("AssembleLinearOperator(source)");
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && f, AFB_ERR_INVALID, "afb_assemble_rhs_source: no pattern / null source");
  return rhs_source(ctx, f, nb_f, nodewise, signed_tri_area);
}

int afb_assemble_rhs_neumann(afb_ctx* ctx, int64_t nb_face, const int32_t* face_nodes, int kind, int nb_value, const double* values, int skip_dirichlet, int mem_space)
{
  afb::NvtxRange nvtx_range("AssembleLinearOperator(boundary)");
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && (nb_face == 0 || face_nodes) && values, AFB_ERR_INVALID, "afb_assemble_rhs_neumann: no pattern / null argument");
  // faces: 2-node edges (any 2-D mesh with straight edges: Tri3, Quad4), 3-node triangles (Tet4), 4-node quadrilaterals (Hexa8)
  const bool hexa = ctx->dim == 3 && ctx->npc == 8;
  AFB_REQUIRE(ctx->npc == ctx->dim + 1 || (ctx->dim == 2 && ctx->npc == 4) || hexa, AFB_ERR_UNSUPPORTED,
              "afb_assemble_rhs_neumann: edges of Tri3 / Quad4 meshes, triangles of Tet4 meshes and quadrilaterals of Hexa8 meshes only");
  AFB_REQUIRE(kind == AFB_NEUMANN_FLUX || kind == AFB_NEUMANN_TRACTION, AFB_ERR_INVALID, "afb_assemble_rhs_neumann: unknown kind %d", kind);
  if (kind == AFB_NEUMANN_FLUX)
    AFB_REQUIRE(nb_value == 1 || nb_value == ctx->dim, AFB_ERR_INVALID, "afb_assemble_rhs_neumann: a flux takes 1 value or one per space dimension (got %d)", nb_value);
  else
    AFB_REQUIRE(nb_value == ctx->b, AFB_ERR_INVALID, "afb_assemble_rhs_neumann: a traction takes one value per DoF of a node (%d, got %d)", ctx->b, nb_value);
  if (nb_face <= 0) return AFB_OK;
  const void* faces = nullptr;
  AFB_TRY(stage(ctx, ctx->tmp_ids, face_nodes, sizeof(int32_t) * (size_t)nb_face * (size_t)(hexa ? 4 : ctx->dim), mem_space, &faces));
  return rhs_neumann(ctx, nb_face, static_cast<const int32_t*>(faces), kind, nb_value, values, skip_dirichlet);
}

int afb_set_dirichlet_nodes(afb_ctx* ctx, int32_t n, const int32_t* node_ids, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh, AFB_ERR_INVALID, "afb_set_dirichlet_nodes: no mesh");
  AFB_TRY(ctx->dir_node.reserve((size_t)ctx->nb_node));
  AFB_CUDA(cudaMemsetAsync(ctx->dir_node.p, 0, (size_t)ctx->nb_node, ctx->stream));
  ctx->has_dir_nodes = false;
  if (n <= 0 || !node_ids) return AFB_OK;
  const void* ids = nullptr;
  AFB_TRY(stage(ctx, ctx->tmp_ids, node_ids, sizeof(int32_t) * (size_t)n, mem_space, &ids));
  AFB_TRY(scatter_flags(ctx, ctx->dir_node.as<uint8_t>(), nullptr, 1, n, (const int32_t*)ids, nullptr));
  ctx->has_dir_nodes = true;
  return AFB_OK;
}

int afb_dirichlet_penalty(afb_ctx* ctx, int weak, double penalty, int32_t n, const int32_t* dof_ids, const double* g, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_dirichlet_penalty: no pattern");
  if (n <= 0) return AFB_OK;
  AFB_TRY(ensure_values_zeroed(ctx));
  ctx->values_touched = true;
  const void *ids = nullptr, *vals = nullptr;
  AFB_TRY(stage(ctx, ctx->tmp_ids, dof_ids, sizeof(int32_t) * (size_t)n, mem_space, &ids));
  AFB_TRY(stage(ctx, ctx->tmp_vals, g, sizeof(double) * (size_t)n, mem_space, &vals));
  return dirichlet_penalty(ctx, weak, penalty, n, (const int32_t*)ids, (const double*)vals);
}

static int ensure_dof_flags(afb_ctx* ctx, DevBuf& info, DevBuf& value, bool& has)
{
  if (has) return AFB_OK;
  const size_t nb_dof = (size_t)ctx->nb_node * ctx->b;
  AFB_TRY(info.reserve(nb_dof));
  AFB_TRY(value.reserve(sizeof(double) * nb_dof));
  AFB_CUDA(cudaMemsetAsync(info.p, 0, nb_dof, ctx->stream));
  AFB_CUDA(cudaMemsetAsync(value.p, 0, sizeof(double) * nb_dof, ctx->stream));
  has = true;
  return AFB_OK;
}

int afb_set_elimination(afb_ctx* ctx, int type, int32_t n, const int32_t* dof_ids, const double* g, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_set_elimination: no pattern");
  AFB_REQUIRE(type == AFB_ELIMINATE_ROW || type == AFB_ELIMINATE_ROW_COLUMN, AFB_ERR_INVALID, "bad elimination type %d", type);
  if (n <= 0) return AFB_OK;
  AFB_TRY(ensure_dof_flags(ctx, ctx->elim_info, ctx->elim_value, ctx->has_elim));
  const void *ids = nullptr, *vals = nullptr;
  AFB_TRY(stage(ctx, ctx->tmp_ids, dof_ids, sizeof(int32_t) * (size_t)n, mem_space, &ids));
  AFB_TRY(stage(ctx, ctx->tmp_vals, g, sizeof(double) * (size_t)n, mem_space, &vals));
  AFB_TRY(scatter_flags(ctx, ctx->elim_info.as<uint8_t>(), ctx->elim_value.as<double>(), (uint8_t)type, n, (const int32_t*)ids, (const double*)vals));
  if (type == AFB_ELIMINATE_ROW_COLUMN) ctx->has_rc = true;
  return AFB_OK;
}

int afb_set_forced_values(afb_ctx* ctx, int32_t n, const int32_t* dof_ids, const double* v, int mem_space)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_set_forced_values: no pattern");
  if (n <= 0) return AFB_OK;
  AFB_TRY(ensure_dof_flags(ctx, ctx->forced_info, ctx->forced_value, ctx->has_forced));
  const void *ids = nullptr, *vals = nullptr;
  AFB_TRY(stage(ctx, ctx->tmp_ids, dof_ids, sizeof(int32_t) * (size_t)n, mem_space, &ids));
  AFB_TRY(stage(ctx, ctx->tmp_vals, v, sizeof(double) * (size_t)n, mem_space, &vals));
  return scatter_flags(ctx, ctx->forced_info.as<uint8_t>(), ctx->forced_value.as<double>(), 1, n, (const int32_t*)ids, (const double*)vals);
}

int afb_clear_dirichlet(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  ctx->has_elim = ctx->has_forced = ctx->has_rc = false;
  ctx->saved_valid = false;
  ctx->has_dir_nodes = false;
  return AFB_OK;
}

int afb_apply_matrix_transformation(afb_ctx* ctx, int replicate_column0_quirk)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_apply_matrix_transformation: no pattern");
  AFB_TRY(ensure_values_zeroed(ctx));
  return apply_matrix_transformation(ctx, replicate_column0_quirk);
}

int afb_apply_rhs_transformation(afb_ctx* ctx)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_apply_rhs_transformation: no pattern");
  AFB_TRY(ensure_values_zeroed(ctx));
  return apply_rhs_transformation(ctx);
}

static int single_slot(afb_ctx* ctx, int32_t dof_row, int32_t dof_col, int64_t* slot)
{
  const int32_t nb_dof = ctx->nb_node * ctx->b;
  AFB_REQUIRE(dof_row >= 0 && dof_row < nb_dof && dof_col >= 0 && dof_col < nb_dof, AFB_ERR_INVALID, "DoF (%d,%d) out of range [0,%d)", dof_row, dof_col, nb_dof);
  AFB_TRY(ensure_values_zeroed(ctx));
  AFB_TRY(ctx->tmp_ids.reserve(2 * sizeof(int32_t) + sizeof(int64_t) + 8));
  int32_t rc[2] = { dof_row, dof_col };
  char* base = ctx->tmp_ids.as<char>();
  int64_t* dslot = reinterpret_cast<int64_t*>(base + 8);
  AFB_CUDA(cudaMemcpyAsync(base, rc, sizeof(rc), cudaMemcpyHostToDevice, ctx->stream));
  AFB_TRY(lookup_value_slots(ctx, 1, reinterpret_cast<int32_t*>(base), reinterpret_cast<int32_t*>(base) + 1, dslot));
  AFB_CUDA(cudaMemcpyAsync(slot, dslot, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  AFB_REQUIRE(*slot >= 0, AFB_ERR_INVALID, "entry (%d,%d) is not in the sparsity pattern", dof_row, dof_col);
  return AFB_OK;
}

int afb_matrix_get_value(afb_ctx* ctx, int32_t dof_row, int32_t dof_col, double* value)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && value, AFB_ERR_INVALID, "afb_matrix_get_value: no pattern / null output");
  int64_t slot = -1;
  AFB_TRY(single_slot(ctx, dof_row, dof_col, &slot));
  AFB_CUDA(cudaMemcpyAsync(value, ctx->values.as<double>() + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  return AFB_OK;
}

int afb_matrix_set_value(afb_ctx* ctx, int32_t dof_row, int32_t dof_col, double value, int mode)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_matrix_set_value: no pattern");
  int64_t slot = -1;
  AFB_TRY(single_slot(ctx, dof_row, dof_col, &slot));
  AFB_TRY(ensure_values_zeroed(ctx));
  ctx->values_touched = true; // a later tiled assembly adds on top instead of overwriting (the reference's += semantics)
  if (mode == 1) {
    double old = 0.0;
    AFB_CUDA(cudaMemcpyAsync(&old, ctx->values.as<double>() + slot, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(cudaStreamSynchronize(ctx->stream));
    value += old;
  }
  AFB_CUDA(cudaMemcpyAsync(ctx->values.as<double>() + slot, &value, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  return AFB_OK;
}

int afb_get_csr_view(afb_ctx* ctx, const int32_t** rows, const int32_t** rows_nb_column, const int32_t** columns, double** values, int32_t* nb_row, int64_t* nnz)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_get_csr_view: no pattern");
  AFB_TRY(ensure_values_zeroed(ctx));
  if (ctx->b == 1) {
    if (rows) *rows = ctx->rows.as<int32_t>();
    if (rows_nb_column) *rows_nb_column = ctx->nz_per_row.as<int32_t>();
    if (columns) *columns = ctx->cols.as<int32_t>();
    if (nb_row) *nb_row = ctx->nb_node;
    if (nnz) *nnz = ctx->nnz;
  }
  else {
    // "BSRFormat(toLinearSystem): Linear system was set to use CSR but is incompatible" (femutils/BSRFormat.cc:382-383)
    AFB_REQUIRE(!ctx->assembled || ctx->layout == AFB_LAYOUT_PER_ROW, AFB_ERR_INVALID,
                "CSR view of a b=%d matrix needs AFB_LAYOUT_PER_ROW values (does_linear_system_use_csr)", ctx->b);
    AFB_TRY(ensure_scalar_csr(ctx));
    if (rows) *rows = ctx->csr_rows.as<int32_t>();
    if (rows_nb_column) *rows_nb_column = ctx->csr_nbcol.as<int32_t>();
    if (columns) *columns = ctx->csr_cols.as<int32_t>();
    if (nb_row) *nb_row = ctx->nb_node * ctx->b;
    if (nnz) *nnz = ctx->nnz * ctx->b * ctx->b;
  }
  if (values) *values = ctx->values.as<double>();
  return AFB_OK;
}

int afb_get_bsr(afb_ctx* ctx, const int32_t** rows_index, const int32_t** columns, double** values, const int32_t** nb_nz_per_row, int32_t* nb_block_row,
                int64_t* nb_col, int* block_size, int* value_layout)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_get_bsr: no pattern");
  AFB_TRY(ensure_values_zeroed(ctx));
  if (rows_index) *rows_index = ctx->rows.as<int32_t>();
  if (columns) *columns = ctx->cols.as<int32_t>();
  if (values) *values = ctx->values.as<double>();
  if (nb_nz_per_row) *nb_nz_per_row = ctx->nz_per_row.as<int32_t>();
  if (nb_block_row) *nb_block_row = ctx->nb_node;
  if (nb_col) *nb_col = ctx->nnz;
  if (block_size) *block_size = ctx->b;
  if (value_layout) *value_layout = ctx->layout;
  return AFB_OK;
}

int afb_get_coo(afb_ctx* ctx, const int32_t** coo_rows, const int32_t** coo_cols, double** values, int64_t* nnz)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_get_coo: no pattern");
  AFB_REQUIRE(ctx->b == 1, AFB_ERR_UNSUPPORTED, "COO view is defined for one dof per node");
  AFB_TRY(ensure_values_zeroed(ctx));
  AFB_TRY(ensure_coo_rows(ctx));
  if (coo_rows) *coo_rows = ctx->coo_rows.as<int32_t>();
  if (coo_cols) *coo_cols = ctx->cols.as<int32_t>();
  if (values) *values = ctx->values.as<double>();
  if (nnz) *nnz = ctx->nnz;
  return AFB_OK;
}

int afb_get_rhs(afb_ctx* ctx, double** rhs, int32_t* nb_dof)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_get_rhs: no pattern");
  if (rhs) *rhs = ctx->rhs.as<double>();
  if (nb_dof) *nb_dof = ctx->nb_node * ctx->b;
  return AFB_OK;
}

int afb_get_mesh(afb_ctx* ctx, int* dim, int* npc, int32_t* nb_node, int64_t* nb_cell, int32_t* nb_own_node, const double** xyz, const int32_t** cell_nodes,
                 const uint8_t** node_is_own)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_mesh, AFB_ERR_INVALID, "afb_get_mesh: no mesh");
  if (dim) *dim = ctx->dim;
  if (npc) *npc = ctx->npc;
  if (nb_node) *nb_node = ctx->nb_node;
  if (nb_cell) *nb_cell = ctx->nb_cell;
  if (nb_own_node) *nb_own_node = ctx->nb_own_node;
  if (xyz) *xyz = ctx->coords.as<double>();
  if (cell_nodes) *cell_nodes = ctx->conn.as<int32_t>();
  if (node_is_own) *node_is_own = ctx->all_own ? nullptr : ctx->is_own.as<uint8_t>();
  return AFB_OK;
}

int afb_copy_to_host(afb_ctx* ctx, int which, void* dst, size_t* bytes)
{
  AFB_TRY(check_ctx(ctx));
  AFB_TRY(verify_pending(ctx));
  const void* src = nullptr;
  size_t n = 0;
  const int b = ctx->b;
  const bool need_pattern = which <= AFB_ARRAY_CSR_NB_COLUMN;
  AFB_REQUIRE(!need_pattern || ctx->has_pattern, AFB_ERR_INVALID, "afb_copy_to_host: no pattern");
  AFB_REQUIRE(need_pattern || ctx->has_mesh, AFB_ERR_INVALID, "afb_copy_to_host: no mesh");
  switch (which) {
  case AFB_ARRAY_ROWS: src = ctx->rows.p; n = sizeof(int32_t) * ((size_t)ctx->nb_node + 1); break;
  case AFB_ARRAY_COLUMNS: src = ctx->cols.p; n = sizeof(int32_t) * (size_t)ctx->nnz; break;
  case AFB_ARRAY_VALUES:
    AFB_TRY(ensure_values_zeroed(ctx));
    src = ctx->values.p; n = sizeof(double) * (size_t)ctx->nnz * b * b; break;
  case AFB_ARRAY_NZ_PER_ROW: src = ctx->nz_per_row.p; n = sizeof(int32_t) * (size_t)ctx->nb_node; break;
  case AFB_ARRAY_RHS: src = ctx->rhs.p; n = sizeof(double) * (size_t)ctx->nb_node * b; break;
  case AFB_ARRAY_COO_ROWS:
    AFB_TRY(ensure_coo_rows(ctx));
    src = ctx->coo_rows.p; n = sizeof(int32_t) * (size_t)ctx->nnz; break;
  case AFB_ARRAY_CSR_ROWS:
    AFB_TRY(ensure_scalar_csr(ctx));
    src = ctx->csr_rows.p; n = sizeof(int32_t) * ((size_t)ctx->nb_node * b + 1); break;
  case AFB_ARRAY_CSR_COLUMNS:
    AFB_TRY(ensure_scalar_csr(ctx));
    src = ctx->csr_cols.p; n = sizeof(int32_t) * (size_t)ctx->nnz * b * b; break;
  case AFB_ARRAY_CSR_NB_COLUMN:
    AFB_TRY(ensure_scalar_csr(ctx));
    src = ctx->csr_nbcol.p; n = sizeof(int32_t) * (size_t)ctx->nb_node * b; break;
  case AFB_ARRAY_COORDS: src = ctx->coords.p; n = sizeof(double) * 3 * (size_t)ctx->nb_node; break;
  case AFB_ARRAY_CELL_NODES: src = ctx->conn.p; n = sizeof(int32_t) * (size_t)ctx->npc * (size_t)ctx->nb_cell; break;
  case AFB_ARRAY_NODE_CELL_PTR: src = ctx->nc_ptr.p; n = sizeof(int32_t) * ((size_t)ctx->nb_node + 1); break;
  case AFB_ARRAY_NODE_CELL_LIST: src = ctx->nc_list.p; n = sizeof(int32_t) * (size_t)ctx->npc * (size_t)ctx->nb_cell; break;
  default:
    set_error("afb_copy_to_host: unknown array %d", which);
    return AFB_ERR_INVALID;
  }
  if (bytes) *bytes = n;
  if (dst && n) {
    AFB_CUDA(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream));
    AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  return AFB_OK;
}

int afb_lookup_value_slots(afb_ctx* ctx, int64_t n, const int32_t* dof_rows, const int32_t* dof_cols, int64_t* slots)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_lookup_value_slots: no pattern");
  return lookup_value_slots(ctx, n, dof_rows, dof_cols, slots);
}

int afb_add_values_at(afb_ctx* ctx, int64_t n, const int64_t* slots, const double* contrib)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_add_values_at: no pattern");
  AFB_TRY(ensure_values_zeroed(ctx));
  ctx->values_touched = true;
  return add_values_at(ctx, n, slots, contrib);
}

int afb_values_tail(afb_ctx* ctx, int32_t first_block_row, int64_t* first_value, int64_t* nb_values)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern, AFB_ERR_INVALID, "afb_values_tail: no pattern");
  AFB_REQUIRE(first_block_row >= 0 && first_block_row <= ctx->nb_node, AFB_ERR_INVALID, "row %d out of range", first_block_row);
  int32_t rb = 0;
  AFB_CUDA(cudaMemcpyAsync(&rb, ctx->rows.as<int32_t>() + first_block_row, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t bb = (int64_t)ctx->b * ctx->b;
  if (first_value) *first_value = (int64_t)rb * bb;
  if (nb_values) *nb_values = (ctx->nnz - rb) * bb;
  return AFB_OK;
}

int afb_renumber_columns(afb_ctx* ctx, const int32_t* dof_local_to_global, int32_t* out)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && dof_local_to_global && out, AFB_ERR_INVALID, "afb_renumber_columns: no pattern / null arrays");
  return renumber_columns(ctx, dof_local_to_global, out);
}

int afb_last_timings(afb_ctx* ctx, float* connectivity_ms, float* pattern_ms, float* assemble_ms)
{
  AFB_TRY(check_ctx(ctx));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  AFB_TRY(verify_pending(ctx));
  float* out[3] = { connectivity_ms, pattern_ms, assemble_ms };
  for (int p = 0; p < 3; ++p) {
    if (!out[p]) continue;
    *out[p] = -1.0f;
    if (ctx->timed[p]) AFB_CUDA(cudaEventElapsedTime(out[p], ctx->ev[2 * p], ctx->ev[2 * p + 1]));
  }
  return AFB_OK;
}

int afb_get_ij_arrays(afb_ctx* ctx, int32_t first_own_row, int32_t nb_own_row, const int32_t* dof_local_to_global, const int32_t** ncols, const int32_t** rows,
                      const int32_t** cols, const double** values, int64_t* nb_values)
{
  AFB_TRY(check_ctx(ctx));
  const int32_t *v_rows = nullptr, *v_nbc = nullptr, *v_cols = nullptr;
  double* v_vals = nullptr;
  int32_t nb_row = 0;
  int64_t nnz = 0;
  AFB_TRY(afb_get_csr_view(ctx, &v_rows, &v_nbc, &v_cols, &v_vals, &nb_row, &nnz));
  AFB_REQUIRE(nb_own_row >= 0 && nb_own_row <= nb_row && first_own_row >= 0, AFB_ERR_INVALID, "afb_get_ij_arrays: nb_own_row %d outside [0,%d]", nb_own_row, nb_row);
  AFB_TRY(ctx->ij_rows.reserve(sizeof(int32_t) * (size_t)std::max(nb_own_row, 1)));
  AFB_TRY(fill_iota(ctx, first_own_row, nb_own_row, ctx->ij_rows.as<int32_t>()));
  const int32_t* c = v_cols;
  if (dof_local_to_global) {
    AFB_TRY(ctx->ij_cols.reserve(sizeof(int32_t) * (size_t)std::max<int64_t>(nnz, 1)));
    AFB_TRY(renumber_columns(ctx, dof_local_to_global, ctx->ij_cols.as<int32_t>()));
    c = ctx->ij_cols.as<int32_t>();
  }
  int32_t own_end = 0; // values of the owned rows: rows[nb_own_row]
  AFB_CUDA(cudaMemcpyAsync(&own_end, v_rows + nb_own_row, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ncols) *ncols = v_nbc;
  if (rows) *rows = ctx->ij_rows.as<int32_t>();
  if (cols) *cols = c;
  if (values) *values = v_vals;
  if (nb_values) *nb_values = own_end;
  return AFB_OK;
}

int afb_memcpy_to_host(afb_ctx* ctx, void* dst_host, const void* src_device, size_t bytes)
{
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(bytes == 0 || (dst_host && src_device), AFB_ERR_INVALID, "afb_memcpy_to_host: null pointer");
  AFB_TRY(verify_pending(ctx));
  if (bytes) AFB_CUDA(cudaMemcpyAsync(dst_host, src_device, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  return AFB_OK;
}

int afb_solve_pcg(afb_ctx* ctx, double rtol, double atol, int max_iter, double* x, int mem_space, int* iterations, double* residual)
{
  afb::NvtxRange nvtx_range("StationarySolve");
  AFB_TRY(check_ctx(ctx));
  AFB_REQUIRE(ctx->has_pattern && ctx->assembled, AFB_ERR_INVALID, "afb_solve_pcg: assemble the matrix first");
  AFB_REQUIRE(rtol >= 0.0 && atol >= 0.0 && max_iter >= 0, AFB_ERR_INVALID, "afb_solve_pcg: negative tolerance / iteration count");
  AFB_TRY(verify_pending(ctx));
  return solve_pcg(ctx, rtol, atol, max_iter, x, mem_space, iterations, residual);
}

int afb_inspector_timings(afb_ctx* ctx, float* mesh_tiling_ms, float* value_plan_ms)
{
  AFB_TRY(check_ctx(ctx));
  const TilePlan& P = ctx->plan;
  if (mesh_tiling_ms) *mesh_tiling_ms = (P.mesh_valid && P.mesh_gen == ctx->mesh_gen) ? P.mesh_ms : -1.0f;
  if (value_plan_ms) *value_plan_ms = (P.lists_valid && P.lists_mesh_gen == ctx->mesh_gen) ? P.lists_ms : -1.0f;
  // scalar assemblies run on the chained-slice plan (chain_plan.cu): its inspector is the value plan of this context
  if (value_plan_ms && chain_plan_ms(ctx) >= 0.0f) *value_plan_ms = chain_plan_ms(ctx);
  return AFB_OK;
}

int64_t afb_launch_count(afb_ctx* ctx) { return ctx ? ctx->launches : 0; }

} // extern "C"
