// Bilinear-form assembly, straightforward variants:
//   * cell-wise + fp64 atomics  (reference K11/K12/K15: modules/testlab/CsrGpuBiliAssembly.cc:339-372,
//     CooGpuBiliAssembly.cc:267-351, femutils/BSRFormat.h:257-370)
//   * node-wise, atomic-free    (reference K13/K16: modules/testlab/NodeWiseCsrBiliAssembly.cc:259-296,
//     femutils/BSRFormat.h:406-537)
// Differences to the reference kernels: cell connectivity is read as one 128-bit load,
// geometry is computed once per element visit (one determinant, one reciprocal), the slot
// of (row, col) is found by binary search in the ascending row instead of a linear scan,
// value offsets are 64-bit.  The B200-tuned path is the tiled gather in tiles.cu.
#include "element.cuh"

namespace afb {

template <int NPC>
__device__ __forceinline__ void load_cell_nodes(const int32_t* __restrict__ conn, int64_t cell, int32_t (&nd)[NPC])
{
  const int32_t* cn = conn + cell * NPC;
  if constexpr (NPC == 4) {
    int4 v = __ldg(reinterpret_cast<const int4*>(cn));
    nd[0] = v.x; nd[1] = v.y; nd[2] = v.z; nd[3] = v.w;
  }
  else if constexpr (NPC % 2 == 0) {
#pragma unroll
    for (int i = 0; i < NPC / 2; ++i) {
      int2 v = __ldg(reinterpret_cast<const int2*>(cn) + i);
      nd[2 * i] = v.x; nd[2 * i + 1] = v.y;
    }
  }
  else {
#pragma unroll
    for (int i = 0; i < NPC; ++i) nd[i] = __ldg(cn + i);
  }
}

// ---------------------------------------------------------------------------------------------
// cell-wise atomic
// ---------------------------------------------------------------------------------------------
// (vector reductions -- red.global.add.v2 -- exist for f32 / f16x2 / bf16x2 only: ptxas rejects .v2.f64, so b > 1 blocks stay scalar REDs)
template <class E, int LAYOUT, bool COO>
__global__ void __launch_bounds__(128)
k_assemble_cellwise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own, int64_t nb_cell,
                    const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, const int32_t* __restrict__ coo_rows, int64_t nnz,
                    double* __restrict__ values, ElemParams prm)
{
  constexpr int NPC = E::NPC, B = E::B;
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= nb_cell) return;
  int32_t nd[NPC];
  load_cell_nodes<NPC>(conn, cell, nd);
  E e;
  if (prm.cell_coef) prm.scale = __ldg(prm.cell_coef + cell);
  e.init(coords, nd, prm);
#pragma unroll(NPC <= 4 ? NPC : 1)
  for (int a = 0; a < NPC; ++a) {
    const int32_t r = nd[a];
    if (is_own && !is_own[r]) continue;
    // COO back-end: the reference locates the row segment by a binary search over the COO row array for every entry
    // (femutils/CooFormatMatrix.h:308-353: ~28 dependent loads per entry at 100 M cells).  The COO row array here is the
    // expansion of row_index, which the pattern build keeps: the segment is read directly, as in the CSR back-end.
    (void)coo_rows;
    (void)nnz;
    const int rb = __ldg(rows + r), re = __ldg(rows + r + 1);
    const int nz = re - rb;
#pragma unroll
    for (int bc = 0; bc < NPC; ++bc) {
      const int p = find_col(cols, rb, re, nd[bc]);
      double blk[B * B];
      e.block(a, bc, blk);
#pragma unroll
      for (int i = 0; i < B; ++i)
#pragma unroll
        for (int j = 0; j < B; ++j) atomicAdd(values + value_index<B, LAYOUT>(rb, nz, p, i, j), blk[i * B + j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// node-wise (thread per block row, plain += : each row has exactly one writer)
// ---------------------------------------------------------------------------------------------
template <class E, int LAYOUT, int A>
__device__ __forceinline__ void add_element_row(const E& e, const int32_t (&nd)[E::NPC], const int32_t* __restrict__ cols, int rb, int re, int nz, double* __restrict__ values)
{
  constexpr int NPC = E::NPC, B = E::B;
#pragma unroll
  for (int bc = 0; bc < NPC; ++bc) {
    const int p = find_col(cols, rb, re, nd[bc]);
    double blk[B * B];
    e.block(A, bc, blk);
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
      for (int j = 0; j < B; ++j) {
        double* v = values + value_index<B, LAYOUT>(rb, nz, p, i, j);
        *v += blk[i * B + j];
      }
  }
}

template <class E, int LAYOUT>
__device__ __forceinline__ void add_element_row_dyn(const E& e, int a, const int32_t (&nd)[E::NPC], const int32_t* __restrict__ cols, int rb, int re, int nz, double* __restrict__ values)
{
  constexpr int NPC = E::NPC, B = E::B;
#pragma unroll 1
  for (int bc = 0; bc < NPC; ++bc) {
    const int p = find_col(cols, rb, re, nd[bc]);
    double blk[B * B];
    e.block(a, bc, blk);
#pragma unroll
    for (int i = 0; i < B; ++i)
#pragma unroll
      for (int j = 0; j < B; ++j) {
        double* v = values + value_index<B, LAYOUT>(rb, nz, p, i, j);
        *v += blk[i * B + j];
      }
  }
}

template <class E, int LAYOUT>
__global__ void __launch_bounds__(128)
k_assemble_nodewise(const double* __restrict__ coords, const int32_t* __restrict__ conn, const uint8_t* __restrict__ is_own, int32_t nb_node,
                    const int32_t* __restrict__ nc_ptr, const int32_t* __restrict__ nc_list, int64_t nb_own_cell,
                    const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, double* __restrict__ values, ElemParams prm)
{
  constexpr int NPC = E::NPC, B = E::B;
  const int32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nb_node) return;
  if (is_own && !is_own[r]) return;
  const int rb = __ldg(rows + r), re = __ldg(rows + r + 1);
  const int nz = re - rb;
  const int qb = __ldg(nc_ptr + r), qe = __ldg(nc_ptr + r + 1);
  for (int q = qb; q < qe; ++q) {
    const int64_t cell = __ldg(nc_list + q);
    if (cell >= nb_own_cell) continue; // ghost cells contribute nothing in AFB_FLAG_OWN_CELLS_ONLY mode
    int32_t nd[NPC];
    load_cell_nodes<NPC>(conn, cell, nd);
    int a = 0;
#pragma unroll
    for (int i = 1; i < NPC; ++i)
      if (nd[i] == r) a = i;
    E e;
    if (prm.cell_coef) prm.scale = __ldg(prm.cell_coef + cell);
    e.init(coords, nd, prm);
    if constexpr (NPC <= 4) {
      // compile-time row index: keeps the cofactors in registers (no dynamic indexing)
      switch (a) {
      case 0: add_element_row<E, LAYOUT, 0>(e, nd, cols, rb, re, nz, values); break;
      case 1: add_element_row<E, LAYOUT, 1>(e, nd, cols, rb, re, nz, values); break;
      case 2: add_element_row<E, LAYOUT, 2>(e, nd, cols, rb, re, nz, values); break;
      default: add_element_row<E, LAYOUT, (NPC > 3 ? 3 : 0)>(e, nd, cols, rb, re, nz, values); break;
      }
    }
    else
      add_element_row_dyn<E, LAYOUT>(e, a, nd, cols, rb, re, nz, values);
  }
}

// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
template <class E>
static int launch(afb_ctx* ctx, int format, int variant, int layout, const ElemParams& prm)
{
  const double* coords = ctx->coords.as<double>();
  const int32_t* conn = ctx->conn.as<int32_t>();
  // domain-decomposition modes (afb200.h): ALL_ROWS drops the isOwn gate, OWN_CELLS_ONLY the ghost cells
  const uint8_t* own = (ctx->all_own || (prm.flags & AFB_FLAG_ALL_ROWS)) ? nullptr : ctx->is_own.as<uint8_t>();
  const int64_t nb_cell = (prm.flags & AFB_FLAG_OWN_CELLS_ONLY) ? ctx->nb_own_cell : ctx->nb_cell;
  const int32_t* rows = ctx->rows.as<int32_t>();
  const int32_t* cols = ctx->cols.as<int32_t>();
  double* values = ctx->values.as<double>();
  if (variant == AFB_VARIANT_CELLWISE_ATOMIC) {
    if (nb_cell == 0) return AFB_OK;
    int grid = grid_for(nb_cell, 128);
    if (format == AFB_FORMAT_COO) {
      AFB_TRY(ensure_coo_rows(ctx));
      const int32_t* coo = ctx->coo_rows.as<int32_t>();
      if (layout == AFB_LAYOUT_PER_BLOCK)
        k_assemble_cellwise<E, AFB_LAYOUT_PER_BLOCK, true><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, nb_cell, rows, cols, coo, ctx->nnz, values, prm);
      else
        k_assemble_cellwise<E, AFB_LAYOUT_PER_ROW, true><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, nb_cell, rows, cols, coo, ctx->nnz, values, prm);
    }
    else {
      if (layout == AFB_LAYOUT_PER_BLOCK)
        k_assemble_cellwise<E, AFB_LAYOUT_PER_BLOCK, false><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, nb_cell, rows, cols, nullptr, ctx->nnz, values, prm);
      else
        k_assemble_cellwise<E, AFB_LAYOUT_PER_ROW, false><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, nb_cell, rows, cols, nullptr, ctx->nnz, values, prm);
    }
    AFB_LAUNCH_CHECK(ctx);
    return AFB_OK;
  }
  if (variant == AFB_VARIANT_NODEWISE) {
    if (ctx->nb_node == 0) return AFB_OK;
    int grid = grid_for(ctx->nb_node, 128);
    const int32_t* ptr = ctx->nc_ptr.as<int32_t>();
    const int32_t* list = ctx->nc_list.as<int32_t>();
    if (layout == AFB_LAYOUT_PER_BLOCK)
      k_assemble_nodewise<E, AFB_LAYOUT_PER_BLOCK><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, ctx->nb_node, ptr, list, nb_cell, rows, cols, values, prm);
    else
      k_assemble_nodewise<E, AFB_LAYOUT_PER_ROW><<<grid, 128, 0, ctx->stream>>>(coords, conn, own, ctx->nb_node, ptr, list, nb_cell, rows, cols, values, prm);
    AFB_LAUNCH_CHECK(ctx);
    return AFB_OK;
  }
  set_error("unknown assembly variant %d", variant);
  return AFB_ERR_INVALID;
}

int assemble_bilinear(afb_ctx* ctx, int op, const double* params, int format, int variant, int layout, int flags)
{
  ElemParams prm;
  prm.p0 = params ? params[0] : 0.0;
  prm.p1 = params ? params[1] : 0.0;
  prm.p2 = (params && op == AFB_OP_ELASTODYNAMICS) ? params[2] : 0.0;
  prm.flags = flags;
  const int npc = ctx->npc, dim = ctx->dim;
  if (ctx->has_cell_coef) {
    AFB_REQUIRE((op == AFB_OP_POISSON || op == AFB_OP_DIFFUSION_REACTION) && (npc == dim + 1 || npc == (1 << dim)), AFB_ERR_UNSUPPORTED,
                "a per-cell coefficient (afb_set_cell_coefficient) applies to the Poisson / diffusion-reaction operators on Tri3 / Tet4 / Quad4 / Hexa8 cells");
    prm.cell_coef = ctx->cell_coef.as<double>();
  }
  if (op == AFB_OP_POISSON) {
    if (npc == 4 && dim == 3) return launch<Tet4Poisson>(ctx, format, variant, layout, prm);
    if (npc == 3 && dim == 2) return launch<Tri3Poisson>(ctx, format, variant, layout, prm);
    if (npc == 6 && dim == 2) return launch<Tri6Poisson>(ctx, format, variant, layout, prm);
    if (npc == 10 && dim == 3) return launch<Tet10Poisson>(ctx, format, variant, layout, prm);
    if (npc == 4 && dim == 2) return launch<Quad4Poisson>(ctx, format, variant, layout, prm);
    if (npc == 8 && dim == 3) return launch<Hexa8Poisson>(ctx, format, variant, layout, prm);
  }
  else if (op == AFB_OP_ELASTICITY) {
    if (npc == 4 && dim == 3) return launch<Tet4Elasticity>(ctx, format, variant, layout, prm);
    if (npc == 3 && dim == 2) return launch<Tri3Elasticity>(ctx, format, variant, layout, prm);
    if (npc == 4 && dim == 2) return launch<Quad4Elasticity>(ctx, format, variant, layout, prm);
    if (npc == 8 && dim == 3) return launch<Hexa8Elasticity>(ctx, format, variant, layout, prm);
  }
  else if (op == AFB_OP_BILAPLACIAN) {
    if (npc == 3 && dim == 2) return launch<Tri3Bilaplacian>(ctx, format, variant, layout, prm);
  }
  else if (op == AFB_OP_DIFFUSION_REACTION) {
    if (npc == 4 && dim == 3) return launch<Tet4DiffReact>(ctx, format, variant, layout, prm);
    if (npc == 3 && dim == 2) return launch<Tri3DiffReact>(ctx, format, variant, layout, prm);
    if (npc == 4 && dim == 2) return launch<Quad4DiffReact>(ctx, format, variant, layout, prm);
    if (npc == 8 && dim == 3) return launch<Hexa8DiffReact>(ctx, format, variant, layout, prm);
  }
  else if (op == AFB_OP_ELASTODYNAMICS) {
    if (npc == 4 && dim == 3) return launch<Tet4Elastodynamics>(ctx, format, variant, layout, prm);
    if (npc == 3 && dim == 2) return launch<Tri3Elastodynamics>(ctx, format, variant, layout, prm);
    if (npc == 4 && dim == 2) return launch<Quad4Elastodynamics>(ctx, format, variant, layout, prm);
    if (npc == 8 && dim == 3) return launch<Hexa8Elastodynamics>(ctx, format, variant, layout, prm);
  }
  // mirrors BSRFormat::computeNbColumns returning 0 / testlab _checkCellType FATAL for
  // unsupported cell types (femutils/BSRFormat.cc:339-341, modules/testlab/FemModule.cc:688-699)
  set_error("operator %d is not implemented for %d-node cells in dimension %d", op, npc, dim);
  return AFB_ERR_UNSUPPORTED;
}

} // namespace afb
