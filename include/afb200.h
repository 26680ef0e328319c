/*
 * afb200.h — C ABI of the B200-native bilinear-form assembly path (libafb200.so).
 *
 * This is the drop-in boundary for ArcaneFEM's assembly back-ends: everything between
 * "mesh connectivity + node coordinates in" and "row/col/val (+rhs) arrays out".
 * Plain pointers and sizes only; no C++/torch types.  Each entry point names the
 * reference interface it replaces (paths relative to the ArcaneFEM source root).
 * The thin C++ façade that keeps the reference's class names (CsrFormat, BSRFormat,
 * BSRMatrix, CooFormat, CsrFormatMatrixView, DoFLinearSystem) lives in
 * include/arcanefem_b200/ and calls only these functions.  INTEGRATION.md shows the
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns AFB_OK (0) or a negative error code; afb_last_error()
 *     gives the thread-local message (the reference throws ARCANE_FATAL / ARCANE_THROW:
 *     the façade converts non-zero codes into exceptions).
 *   - one afb_ctx per GPU, used from one host thread at a time (the reference is
 *     single-threaded per MPI rank with one RunQueue: modules/testlab/FemModule.cc:114).
 *   - all work is stream-ordered on the context's CUDA stream (afb_set_stream lets the
 *     host use its own stream, e.g. torch's current stream); nothing here falls back to
 *     the CPU: a missing device or kernel failure is an error.
 *   - the caller owns host arrays; device arrays returned by afb_get_* stay owned by the
 *     context and remain valid until the next afb_build_pattern / afb_destroy
 *     (same contract as DoFLinearSystem::setCSRValues: femutils/DoFLinearSystem.h:318-325).
 *   - indices are Int32, reals are IEEE fp64, DoF id = node_lid*b + component
 *     (femutils/FemDoFsOnNodes.cc:79-111).  Row arrays are written with nb_row+1 entries;
 *     the first nb_row are exactly the reference's sentinel-less arrays
 *     (femutils/CsrFormatMatrix.h:71-80, femutils/BSRFormat.h:131-134).
 *   - columns are emitted ascending inside each row (the reference's intra-row order is
 *     non-deterministic / connectivity-order dependent; HYPRE and PETSc accept any order).
 */
#ifndef AFB200_H
#define AFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFB_API __attribute__((visibility("default")))

typedef struct afb_ctx afb_ctx;

enum {
  AFB_OK = 0,
  AFB_ERR_INVALID = -1,     /* bad argument / call order (ArgumentException in the reference)   */
  AFB_ERR_CUDA = -2,        /* CUDA runtime or kernel failure                                    */
  AFB_ERR_UNSUPPORTED = -3, /* NotImplementedException / NotSupportedException in the reference  */
  AFB_ERR_OVERFLOW = -4     /* Int32 index space exceeded (femutils/BSRFormat.cc:362-364)        */
};

/* operator tag: replaces the device lambda template parameter of
 * BSRFormat::assembleBilinear* (femutils/BSRFormat.h:218-236) */
enum {
  AFB_OP_POISSON = 0,    /* modules/testlab/FemModule.h:342-538, FemModule.cc:267-315, modules/poisson/ElementMatrix.h:28-118 */
  AFB_OP_ELASTICITY = 1, /* modules/elasticity/ElementMatrix.h:41-301; params = {lambda, mu}                               */
  AFB_OP_BILAPLACIAN = 2, /* modules/bilaplacian/ElementMatrix.h:30-47 (Tri3, 2 DoF/node)                                   */
  /* params = { alpha, beta }: alpha * stiffness + beta * consistent mass, 1 DoF/node -- the acoustics module (modules/acoustics/
   * ElementMatrix.h:14,29: alpha = -1, beta = kc2; ElementMatrixHexQuad.h: alpha = +1) and the heat module's matrix (modules/heat/
   * ElementMatrix.h: alpha = lambda, beta = 1/dt).  Tri3 / Tet4 / Quad4 / Hexa8, cell-wise and node-wise variants; a per-cell
   * coefficient (afb_set_cell_coefficient) multiplies the stiffness part. */
  AFB_OP_DIFFUSION_REACTION = 3,
  /* params = { c0, c1, c2 }: the Newmark-beta / generalised-alpha matrix of the elastodynamics module (modules/elastodynamics/
   * ElementMatrix.h:41-60 Tria3, :150-196 Tetra4; coefficients FemModule.cc:197-225): the elasticity matrix with lambda = c1,
   * mu = c2 plus c0 times the consistent mass on every component (ElementMatrixHexQuad.h for Quad4 / Hexa8: the same, per Gauss point).
   * dim DoF/node, BSR; Tri3 / Tet4 / Quad4 / Hexa8, cell-wise and node-wise variants. */
  AFB_OP_ELASTODYNAMICS = 4
};

/* matrix format = how entries are located during the scatter and which view is native
 * (.arc options csr-gpu / coo-gpu / bsr; modules/testlab/Fem.axl) */
enum {
  AFB_FORMAT_CSR = 0, /* per-row search in [row[r], row[r+1])      : CsrGpuBiliAssembly.cc:356-366 */
  AFB_FORMAT_COO = 1, /* binary search over the COO row array      : CooFormatMatrix.h:308-353     */
  AFB_FORMAT_BSR = 2  /* block rows, b = dof per node              : femutils/BSRFormat.h:257-577   */
};

/* assembly variant */
enum {
  AFB_VARIANT_CELLWISE_ATOMIC = 0, /* thread per cell + fp64 atomics: csr-gpu / coo-gpu / bsr                       */
  AFB_VARIANT_NODEWISE = 1,        /* thread per row, no atomics: nwcsr / bsr-atomic-free (AF-CSR_GPU / AF-BSR_GPU)  */
  AFB_VARIANT_TILED_GATHER = 2     /* B200 path: row tiles staged in shared memory, every row written exactly once
                                      (P1 Poisson, P1 elasticity, Tri3 bilaplacian; any format)                       */
};

/* sparsity algorithm of a re-build on an unchanged mesh (the first build of a mesh always starts from the cells) */
enum {
  AFB_SPARSITY_AUTO = 0,             /* FROM_CONNECTIVITY once the init-time connectivity exists (first tiled assembly), else FROM_CELLS */
  AFB_SPARSITY_FROM_CELLS = 1,       /* from the cells' nodes, as computeSparsityAtomic / _computeSparsity
                                        (femutils/BSRFormat.cc:799-1006, modules/testlab/CsrGpuBiliAssembly.cc:42-207)            */
  AFB_SPARSITY_FROM_CONNECTIVITY = 2 /* from the init-time node-node connectivity, as computeSparsityAtomicFree /
                                        _buildMatrixNodeWiseCsr (femutils/BSRFormat.cc:445-790,
                                        modules/testlab/NodeWiseCsrBiliAssembly.cc:90-152); P1 cells, else FROM_CELLS              */
};

/* BSR value layout (femutils/BSRFormat.cc:367: per-block unless the solver consumes CSR) */
enum {
  AFB_LAYOUT_PER_BLOCK = 0, /* begin*b*b + i*b + j                   (femutils/BSRFormat.h:292-296) */
  AFB_LAYOUT_PER_ROW = 1    /* row[r]*b*b + b*(x + i*nz_r) + j       (femutils/BSRFormat.h:356)      */
};

enum { AFB_MEM_HOST = 0, AFB_MEM_DEVICE = 1 };

/* flags for afb_assemble_bilinear */
enum {
  /* Tri3: use the SIGNED area like the testlab CSR/COO back-ends (modules/testlab/FemModule.h:359);
   * default is the unsigned area of the BSR lambdas (femutils/ArcaneFemFunctionsGpu.h:76-86). */
  AFB_FLAG_SIGNED_TRI_AREA = 1,
  /* Domain decomposition with ghost-row exchange (design B): only the sub-domain's
   * own cells [0, nb_own_cell) contribute (afb_set_own_cell_count) ... */
  AFB_FLAG_OWN_CELLS_ONLY = 2,
  /* ... into the rows of ALL local nodes, ghost nodes included (no isOwn gate): the ghost rows then
   * hold the partial sums that the owner adds (afb_add_values_at). */
  AFB_FLAG_ALL_ROWS = 4
};

/* matrix elimination type: femutils/FemUtilsGlobal.h:51-62 */
enum { AFB_ELIMINATE_NONE = 0, AFB_ELIMINATE_ROW = 1, AFB_ELIMINATE_ROW_COLUMN = 2 };

/* arrays addressable through afb_copy_to_host */
enum {
  AFB_ARRAY_ROWS = 0,         /* int32[nb_block_row+1]                */
  AFB_ARRAY_COLUMNS = 1,      /* int32[nb_block_nnz]                  */
  AFB_ARRAY_VALUES = 2,       /* double[nb_block_nnz*b*b]             */
  AFB_ARRAY_NZ_PER_ROW = 3,   /* int32[nb_block_row]                  */
  AFB_ARRAY_RHS = 4,          /* double[nb_block_row*b]               */
  AFB_ARRAY_COO_ROWS = 5,     /* int32[nb_block_nnz]                  */
  AFB_ARRAY_CSR_ROWS = 6,     /* expanded scalar CSR (b>1): int32[nb_row*b+1] */
  AFB_ARRAY_CSR_COLUMNS = 7,  /* int32[nb_block_nnz*b*b]              */
  AFB_ARRAY_CSR_NB_COLUMN = 8,/* int32[nb_row*b]                      */
  AFB_ARRAY_COORDS = 9,       /* double[nb_node*3]                    */
  AFB_ARRAY_CELL_NODES = 10,  /* int32[nb_cell*npc]                   */
  AFB_ARRAY_NODE_CELL_PTR = 11, /* int32[nb_node+1]                   */
  AFB_ARRAY_NODE_CELL_LIST = 12 /* int32[nb_cell*npc]                 */
};

/* ---- lifetime ------------------------------------------------------------------------ */

/* ctor of CsrFormat/BSRFormat + the RunQueue they hold (femutils/BSRFormat.cc:233-240). */
AFB_API int afb_create(int device, afb_ctx** out);
AFB_API int afb_destroy(afb_ctx* ctx);
AFB_API const char* afb_last_error(void);
AFB_API const char* afb_version(void);
/* Use the caller's CUDA stream (cudaStream_t) for all subsequent work; NULL = own stream. */
AFB_API int afb_set_stream(afb_ctx* ctx, void* cuda_stream);
/* RunQueue::barrier() */
AFB_API int afb_synchronize(afb_ctx* ctx);

/* ---- mesh in -------------------------------------------------------------------------- */

/*
 * Replaces the IMesh* / UnstructuredMeshConnectivityView / VariableNodeReal3 /
 * ItemGenericInfoListView::isOwn inputs of every back-end
 * (modules/testlab/CsrGpuBiliAssembly.cc:330-335, femutils/BSRFormat.h:260-263).
 * dim in {2,3}; nodes_per_cell in {3,4} (P1) or {6,10} (P2, reference node order
 * femutils/ArcaneFemFunctions.h:3245-3262,3893-3911); xyz = AoS double[nb_node][3];
 * cell_nodes = int32[nb_cell][npc]; node_is_own = uint8[nb_node] or NULL (all owned).
 * Also builds the node->cell connectivity (Arcane's nodeCell view), ascending cell ids.
 * mem_space HOST copies to the device; DEVICE keeps the caller's pointers (zero-copy): cell_nodes must then be 16-byte
 * aligned (rows are read with 128-bit loads) and xyz 8-byte aligned, else AFB_ERR_INVALID.
 */
AFB_API int afb_set_mesh(afb_ctx* ctx, int dim, int nodes_per_cell, int32_t nb_node, int64_t nb_cell,
                         const double* xyz, const int32_t* cell_nodes, const uint8_t* node_is_own, int mem_space);

/*
 * New node coordinates on the same topology (a time loop / moving mesh: what changes between two
 * AssembleBilinearOperator calls of a reference module is the field data, not Arcane's connectivity).  Everything derived
 * from the connectivity alone -- node->cell lists, the sparsity structures, the plans of the tiled executors -- stays valid.
 * xyz: [nb_node][3] AoS like afb_set_mesh.
 */
AFB_API int afb_update_coordinates(afb_ctx* ctx, const double* xyz, int mem_space);

/*
 * Per-cell coefficient of the Poisson operator: the element matrix of cell c is multiplied by coefficient[c] -- the
 * conductivity m_cell_lambda of the reference's fourier / heat modules (modules/fourier/ElementMatrix.h:11-57,
 * `area * lambda * (dxU ^ dxU) + ...`; multi-material cases set it per cell group), the permittivity of electrostatics.
 * The FourierNL module's Picard loop re-assembles with a new one every iteration (modules/fouriernl/ElementMatrix.h:29-41: lambda at
 * the mean of the previous iterate over the cell).
 * [nb_cell] doubles, copied; NULL switches it off; a new mesh drops it.  Cell-wise and node-wise variants on Tri3 / Tet4 /
 * Quad4 / Hexa8; the tiled gather takes it for AFB_OP_POISSON on Tri3 / Tet4 with the brick executor (its plan does not depend on
 * the coefficient: a new coefficient costs no inspector run), AFB_ERR_UNSUPPORTED for the chained-slice executors.
 */
AFB_API int afb_set_cell_coefficient(afb_ctx* ctx, const double* coefficient, int mem_space);

/* Cells [0, nb_own_cell) belong to this sub-domain, cells [nb_own_cell, nb_cell) are ghost cells
 * (Arcane's one-layer ghost cells; Cell::isOwn()).  Default after afb_set_mesh: all cells own. */
AFB_API int afb_set_own_cell_count(afb_ctx* ctx, int64_t nb_own_cell);
AFB_API int afb_get_own_cell_count(afb_ctx* ctx, int64_t* nb_own_cell);

/*
 * Synthetic structured box of SURVEY.md §8(d), generated on the device (bench + tests):
 * [0,1]^dim, n^dim cubes, Kuhn 6-tet / 2-triangle split, node id = i+(n+1)(j+(n+1)k),
 * deterministic interior jitter (bit-identical to arcanefem_b200/mesh.py::box_mesh).
 * Domain-decomposition slab = cube layers [k_lo,k_hi) of the last axis (k_lo=0,k_hi=n = whole box).
 * Node ownership follows the reference's rule "one owner per node": the planes k_lo+1..k_hi are
 * owned (plane 0 too when k_lo = 0); plane k_lo > 0 belongs to the lower neighbour.  Local node
 * numbering is owned-first: owned planes ascending, then the bottom ghost plane, then -- with
 * ghost_cell_layer != 0 and k_hi < n -- the top ghost plane k_hi+1, whose cube layer k_hi is appended
 * after the own cells as Arcane's one-layer ghost cells (afb_get_mesh reports nb_own_node; the own
 * cell count is set as by afb_set_own_cell_count).  With the ghost layer every owned row sees all its
 * cells (reference scheme, no exchange); without it -- or with AFB_FLAG_OWN_CELLS_ONLY -- the rows of
 * the interface planes hold partial sums to be exchanged.
 */
AFB_API int afb_mesh_generate_box(afb_ctx* ctx, int dim, int n, double jitter, uint32_t seed, int k_lo, int k_hi, int ghost_cell_layer);

/* ---- sparsity --------------------------------------------------------------------------- */

/*
 * Replaces BSRFormat::initialize + computeSparsity{,Atomic,AtomicFree}
 * (femutils/BSRFormat.cc:350-372,777-790,799-1006), FemModuleTestlab::_computeSparsity /
 * _buildMatrix* (modules/testlab/CsrGpuBiliAssembly.cc:187-207, CsrBiliAssembly.cc:23-92,
 * NodeWiseCsrBiliAssembly.cc:115-152, CooGpuBiliAssembly.cc:76-232) and
 * CsrFormat/CooFormat/BSRMatrix::initialize (allocation + zero fill).
 * Block pattern over nodes (P1: node + edge neighbours; P2: all nodes sharing a cell),
 * nb_block_nnz = nbNode + 2*nbEdge.  Also fills nb_nz_per_row (computeNzPerRowArray,
 * femutils/BSRFormat.cc:400-440) and zeroes values/rhs.
 */
AFB_API int afb_build_pattern(afb_ctx* ctx, int nb_dof_per_node, int32_t* nb_block_row, int64_t* nb_block_nnz);

/*
 * Which of the reference's two sparsity algorithms a re-build on an unchanged mesh follows (AFB_SPARSITY_*):
 * BSRFormat::computeSparsityAtomic (from the cells) or BSRFormat::computeSparsityAtomicFree (from the node-node
 * connectivity Arcane holds since init; here: the tile-local node-node connectivity of the mesh tiling).
 * Both produce the same row_index / columns (ascending); the setting persists until changed.
 */
AFB_API int afb_set_sparsity_algorithm(afb_ctx* ctx, int algorithm);

/*
 * Executor behind AFB_VARIANT_TILED_GATHER for b = 1 (the node-wise, atomic-free back-ends of the reference:
 * modules/testlab/NodeWiseCsrBiliAssembly.cc:157-297, femutils/BSRFormat.h:406-577).  All three write every row exactly
 * once, bit-reproducibly, and agree to 1e-12; they differ in how the work is cut and scheduled on the SMs:
 *   AFB_TILED_EXEC_BRICKS       spatial bricks of rows, one CTA per brick, halo cells recomputed (default: measured fastest)
 *   AFB_TILED_EXEC_CHAIN        columns swept slice by slice, cells shared by consecutive slices stay in shared memory
 *   AFB_TILED_EXEC_CHAIN_FLOW   the same slices through a warp-specialised mbarrier pipeline, one CTA per SM
 * b > 1 always runs on bricks.  The setting persists until changed; AFB_TILED_EXEC in the environment sets the default
 * ("bricks", "chain", "flow").
 */
enum { AFB_TILED_EXEC_BRICKS = 0, AFB_TILED_EXEC_CHAIN = 1, AFB_TILED_EXEC_CHAIN_FLOW = 2 };
AFB_API int afb_set_tiled_executor(afb_ctx* ctx, int executor);
/*
 * Executor behind AFB_VARIANT_TILED_GATHER for elasticity (b = 2, 3; BSRFormat::assembleBilinearAtomicFree with a b x b
 * block per node pair, femutils/BSRFormat.h:406-577).  Both write every block exactly once and agree to 1e-12:
 *   AFB_VEC_EXEC_ROWS    lanes = consecutive entries of whole rows; blocks staged per warp and written as contiguous runs;
 *                        the diagonal block is minus the sum of the row's other blocks (measured fastest on Tet4, b = 3)
 *   AFB_VEC_EXEC_UNITS   entries grouped by list length, symmetric twins summed once, blocks written from registers
 *                        (measured fastest on Tri3, b = 2, where a row has 7 short lists)
 *   AFB_VEC_EXEC_AUTO    rows for Tet4, units for Tri3 (default)
 * The bilaplacian always runs on AFB_VEC_EXEC_UNITS (its blocks have no zero row sums).  The setting persists until changed;
 * AFB_VEC_EXEC in the environment sets the default ("auto", "rows", "units").
 */
enum { AFB_VEC_EXEC_AUTO = 0, AFB_VEC_EXEC_ROWS = 1, AFB_VEC_EXEC_UNITS = 2 };
AFB_API int afb_set_vector_executor(afb_ctx* ctx, int executor);
/*
 * Tuning / test knob of the tiled executors: plan records (contribution lists) larger than `bytes` are not staged in shared
 * memory through the TMA engine but read from global memory (the branch oversized tiles take).  Default: the executor's
 * staging capacity.  bytes = 0 sends every tile through the global-memory branch.
 */
AFB_API int afb_set_tiled_stage_limit(afb_ctx* ctx, int64_t bytes);

/*
 * The reference's matrix-format options, kept as names: testlab's boolean options of modules/testlab/Fem.axl:42-95
 * ("legacy", "coo", "coo-sorting", "coo-gpu", "coo-sorting-gpu", "csr", "csr-gpu", "nwcsr", "blcsr", "bsr", "bsr-atomic-free")
 * and the production modules' <matrix-format> strings ("DOK", "BSR", "AF-BSR": modules/poisson/Fem.axl:31,
 * modules/elasticity/Fem.axl:37).  Writes the AFB_FORMAT_*, AFB_VARIANT_* and AFB_SPARSITY_* values that back-end maps
 * to here (host-only back-ends map to the device variant producing the same matrix).  Case-insensitive;
 * AFB_ERR_INVALID for an unknown name.
 */
AFB_API int afb_options_from_name(const char* matrix_format_option, int* format, int* variant, int* sparsity);

/* ---- bilinear form ------------------------------------------------------------------------ */

/* BSRFormat::resetMatrixValues / DoFLinearSystem::clearValues (values only) */
AFB_API int afb_reset_values(afb_ctx* ctx);

/*
 * Replaces BSRFormat::assembleBilinear{Atomic,AtomicFree} (femutils/BSRFormat.h:257-577) and
 * FemModuleTestlab::_assemble{Csr,Coo,NodeWiseCsr}...BilinearOperator{TRIA3,TETRA4}
 * (modules/testlab/CsrGpuBiliAssembly.cc:222-374, CooGpuBiliAssembly.cc:237-352,
 * NodeWiseCsrBiliAssembly.cc:157-297).  Adds into `values` (call afb_reset_values first for
 * a fresh matrix).  Rows of non-owned nodes are left untouched (isOwn gate).
 * op/params: see AFB_OP_*; format/variant/layout: see enums; flags: AFB_FLAG_*.
 */
AFB_API int afb_assemble_bilinear(afb_ctx* ctx, int op, const double* params, int nb_params,
                                  int format, int variant, int value_layout, int flags);

/* ---- linear form / Dirichlet ------------------------------------------------------------- */

/* rhs_values.fill(0) (modules/testlab/FemModule.cc:724-725) */
AFB_API int afb_rhs_reset(afb_ctx* ctx);

/*
 * Constant source term, b components f[0..b): rhs[dof(n,k)] += f[k]*meas/npc on P1 simplices; on Quad4 / Hexa8
 * rhs[dof(n,k)] += f[k] * integral of N_n by the 2x2 / 2x2x2 Gauss rule (femutils/ArcaneFemFunctions.cc:222-290,437-483).
 * nodewise=0: cell-wise with atomics, skips nodes flagged by afb_set_dirichlet_nodes
 *   (modules/testlab/FemModule.cc:836-868,1358-1532; modules/elasticity/BodyForce.h:93-104);
 * nodewise=1: per-node sum over incident cells, rhs = sum
 *   (femutils/ArcaneFemFunctionsGpu.h:675-708).
 * signed_tri_area: testlab uses the signed triangle area (FemModule.cc:1762).
 */
AFB_API int afb_assemble_rhs_source(afb_ctx* ctx, const double* f, int nb_f, int nodewise, int signed_tri_area);

/*
 * Boundary integrals of the RHS over P1 faces (edges of a Tri3 mesh, triangles of a Tet4 mesh), added to rhs:
 *   AFB_NEUMANN_FLUX, 1 value       rhs[dof(n,0)] += value * measure / nn                    constant flux
 *   AFB_NEUMANN_FLUX, dim values    rhs[dof(n,0)] += (N . q) * measure / nn                  flux vector q, unit normal N
 *   AFB_NEUMANN_TRACTION, b values  rhs[dof(n,k)] += t[k] * measure / nn                     traction (NULL components = 0)
 * for the owned nodes n of every face (nn = nodes per face).  Replaces the flux part of _assembleCsrGpuLinearOperator
 * (modules/testlab/FemModule.cc:1534-1706), BoundaryConditions{2D,3D}::applyNeumannToRhs{Tria3,Tetra4}
 * (femutils/ArcaneFemFunctionsGpu.cc:679-738,1082-1141) and applyTractionToRhs{Tria3,Tetra4}
 * (femutils/ArcaneFemFunctions.h:2854-2885,2188-2220; a host loop upstream).
 * face_nodes[nb_face][dim]: Arcane's faceNode order of the group's faces, with the first two nodes swapped for faces that
 * are not "subdomain boundary outside" -- the swap computeNormalFace / computeNormalTriangle apply
 * (femutils/ArcaneFemFunctionsGpu.h:159-214), so that N = (y1-y0, x0-x1)/|.| resp. (n1-n0)x(n2-n0)/|.| points outward.
 * Edges of a Quad4 mesh are taken like those of a Tri3 mesh (the 2-point rule of applyNeumannToRhsQuad4,
 * femutils/ArcaneFemFunctions.h:2762-2825, integrates the same linear functions).  Hexa8 meshes: face_nodes[nb_face][4],
 * flux only; 2x2 Gauss rule on the bilinear patch, detJ = |dr/dxi x dr/deta| and the unit normal at every Gauss point from the
 * face's node order AS ARCANE STORES IT (smallest unique id first, then towards its smaller neighbour) -- upstream makes no
 * outside-of-the-domain test there (femutils/ArcaneFemFunctions.h:1843-1953), so its q.n term follows the numbering.
 * skip_dirichlet != 0: nodes marked by afb_set_dirichlet_nodes receive nothing (testlab: FemModule.cc:1577).
 */
enum { AFB_NEUMANN_FLUX = 0, AFB_NEUMANN_TRACTION = 1 };
AFB_API int afb_assemble_rhs_neumann(afb_ctx* ctx, int64_t nb_face, const int32_t* face_nodes, int kind, int nb_value, const double* values, int skip_dirichlet,
                                     int mem_space);

/* Node flags m_u_dirichlet (modules/testlab/FemModule.cc:647-677); NULL / n=0 clears. */
AFB_API int afb_set_dirichlet_nodes(afb_ctx* ctx, int32_t n, const int32_t* node_ids, int mem_space);

/*
 * Penalty / weak penalty on a list of DoFs: A[i,i] = P (weak: += P), rhs[i] = P*g
 * (modules/testlab/FemModule.cc:728-790,1201-1313).  Must be called after the assembly
 * (stream order guarantees the atomics have completed).
 */
AFB_API int afb_dirichlet_penalty(afb_ctx* ctx, int weak, double penalty, int32_t n, const int32_t* dof_ids, const double* g, int mem_space);

/* DoFLinearSystem::eliminateRow / eliminateRowColumn (femutils/DoFLinearSystem.h) on a DoF list;
 * type = AFB_ELIMINATE_*.  DoFLinearSystem::clearValues resets them: afb_clear_dirichlet. */
AFB_API int afb_set_elimination(afb_ctx* ctx, int type, int32_t n, const int32_t* dof_ids, const double* g, int mem_space);
/* forced_info / forced_value (femutils/ArcaneFemFunctionsGpu.cc:54-105) */
AFB_API int afb_set_forced_values(afb_ctx* ctx, int32_t n, const int32_t* dof_ids, const double* v, int mem_space);
AFB_API int afb_clear_dirichlet(afb_ctx* ctx);

/*
 * CsrDoFLinearSystemImpl::applyMatrixTransformation (femutils/CsrDoFLinearSystemImpl.cc:235-242):
 * row elimination, row+column elimination (pre-elimination values are kept for the RHS step,
 * replacing the host OrderedRowColumnMap), forced diagonal values.  Works on the scalar CSR
 * view (b>1 needs AFB_LAYOUT_PER_ROW).  replicate_column0_quirk != 0 reproduces the reference's
 * `if (column_index > 0)` (CsrDoFLinearSystemImpl.cc:111), i.e. column 0 is never touched by the
 * row+column pass.
 */
AFB_API int afb_apply_matrix_transformation(afb_ctx* ctx, int replicate_column0_quirk);
/* CsrDoFLinearSystemImpl::applyRHSTransformation (:247-253): rhs[col] -= A[row,col]*g_row for
 * RC-eliminated rows (DoFLinearSystemImplBase.cc:55-88), then rhs[row] = g_row. */
AFB_API int afb_apply_rhs_transformation(afb_ctx* ctx);

/*
 * Single-entry host access of the reference containers (UVM dereference in the reference):
 * BSRMatrix::getValue / setValue / addValue (femutils/BSRFormat.h:89-104, index by
 * BSRMatrix::findValueIndex femutils/BSRFormat.cc:79-106) and CsrFormat::matrixSetValue /
 * matrixAddValue (femutils/CsrFormatMatrix.h:58-70,107-110).  Scalar DoF ids; synchronous.
 * mode: 0 = set, 1 = add.  AFB_ERR_INVALID when (dof_row, dof_col) is not in the pattern.
 */
AFB_API int afb_matrix_get_value(afb_ctx* ctx, int32_t dof_row, int32_t dof_col, double* value);
AFB_API int afb_matrix_set_value(afb_ctx* ctx, int32_t dof_row, int32_t dof_col, double value, int mode);

/* ---- views out (device pointers) ---------------------------------------------------------- */

/*
 * CsrFormat::view() / CSRFormatView (femutils/CsrFormatMatrixView.h:135-210), laid out as
 * HYPRE_IJMatrixSetValues(nrows, ncols=rows_nb_column, rows, cols, values) expects
 * (femutils/HypreDoFLinearSystem.cc:501-514).  For b>1 this is BSRMatrix::toCsr
 * (femutils/BSRFormat.cc:110-172): rows[nb_row*b+1], col = block_col*b+k, values shared
 * (requires AFB_LAYOUT_PER_ROW), built on the device instead of the reference's host loops.
 */
AFB_API int afb_get_csr_view(afb_ctx* ctx, const int32_t** rows, const int32_t** rows_nb_column, const int32_t** columns,
                             double** values, int32_t* nb_row, int64_t* nnz);
/* BSRMatrix arrays (femutils/BSRFormat.h:131-134) */
AFB_API int afb_get_bsr(afb_ctx* ctx, const int32_t** rows_index, const int32_t** columns, double** values,
                        const int32_t** nb_nz_per_row, int32_t* nb_block_row, int64_t* nb_col, int* block_size, int* value_layout);
/* CooFormat arrays / _translateCSRToCOO (femutils/CsrFormatMatrix.cc:161-184) = the
 * MatSetPreallocationCOOLocal(nnz, coo_rows, coo_cols) + MatSetValuesCOO(values) layout
 * (femutils/PetscDoFLinearSystem.cc:329-345,398). */
AFB_API int afb_get_coo(afb_ctx* ctx, const int32_t** coo_rows, const int32_t** coo_cols, double** values, int64_t* nnz);
AFB_API int afb_get_rhs(afb_ctx* ctx, double** rhs, int32_t* nb_dof);
/* raw mesh pointers on the device (coords, cell_nodes, node_is_own) */
AFB_API int afb_get_mesh(afb_ctx* ctx, int* dim, int* nodes_per_cell, int32_t* nb_node, int64_t* nb_cell, int32_t* nb_own_node,
                         const double** xyz, const int32_t** cell_nodes, const uint8_t** node_is_own);

/* UVM host access of the reference (getValue, dumps, tests): synchronous copy; returns the
 * number of bytes of the array in *bytes when dst == NULL. */
AFB_API int afb_copy_to_host(afb_ctx* ctx, int which, void* dst, size_t* bytes);

/* ---- multi-GPU support (domain decomposition) ------------------------------ */

/*
 * Slot lookup on the owner side of a ghost-row exchange: for n scalar (dof_row, dof_col)
 * pairs (device arrays, local numbering) writes the index into `values` or -1.
 * Replaces the role of the ghost DoF synchronisation in
 * HypreDoFLinearSystemImpl::_computeMatrixNumeration (femutils/HypreDoFLinearSystem.cc:209-249).
 */
AFB_API int afb_lookup_value_slots(afb_ctx* ctx, int64_t n, const int32_t* dof_rows, const int32_t* dof_cols, int64_t* slots);
/* values[slots[i]] += contrib[i] (unique slots per call; device arrays) */
AFB_API int afb_add_values_at(afb_ctx* ctx, int64_t n, const int64_t* slots, const double* contrib);
/* Index of the first value of block row `first_block_row` and the number of doubles from
 * there to the end: the ghost rows' partial sums are one contiguous tail of `values`
 * when ghosts are numbered last (zero-copy NCCL send buffer). */
AFB_API int afb_values_tail(afb_ctx* ctx, int32_t first_block_row, int64_t* first_value, int64_t* nb_values);

/*
 * Ghost-row exchange over NVLink peer memory (one process per GPU, CUDA IPC): the owner of a node pulls the
 * partial rows its neighbours computed straight out of their `values` arrays and adds them, in ONE kernel per
 * assembly (signal -> pull + add -> acknowledge -> the sender zeroes its ghost rows), instead of an NCCL
 * send/recv group plus zero-fill and accumulate launches.  The reference leaves this step to the solver's
 * parallel matrix (HYPRE IJ off-processor entries, femutils/HypreDoFLinearSystem.cc:461-520).
 *   afb_p2p_export     IPC handles (AFB_P2P_HANDLE_BYTES each) of this context's values array and flag block;
 *                      the host exchanges them with the neighbours (any transport: torch.distributed, MPI)
 *   afb_p2p_connect    per neighbour k: its rank, its two handles, the slice [pull_first, +pull_count) of ITS
 *                      values holding partial rows of nodes this rank owns, the device array slots[k] (this
 *                      rank's value slot of every double of that slice, afb_lookup_value_slots), and the slice
 *                      [send_first, +send_count) of THIS rank's values the neighbour pulls (zeroed afterwards).
 *                      Neighbour relations must be symmetric (counts may be 0).  Ranks < 64.
 *   afb_p2p_exchange   stream-ordered after afb_assemble_bilinear(AFB_FLAG_OWN_CELLS_ONLY | AFB_FLAG_ALL_ROWS);
 *                      every rank of the decomposition must call it once per assembly
 *   afb_p2p_exchange_async / afb_p2p_wait   the same kernel on an internal high-priority side stream, ordered after the
 *                      work already queued on the context stream; afb_p2p_wait makes the context stream wait for it.
 *                      Between the two the caller may queue work that does not touch `values` (in a time loop: the next
 *                      afb_build_pattern), which then overlaps the ranks' synchronisation.  Everything reading or writing
 *                      `values` must come after afb_p2p_wait (afb_p2p_status and a second exchange wait by themselves).
 *   afb_p2p_status     synchronises; 0 = fine, 1/2 = a neighbour did not show up within the kernel's time-out
 *   afb_p2p_disconnect closes the mappings (also done by afb_destroy); collective by convention
 * Re-export and re-connect after afb_build_pattern moved the values array (afb_p2p_exchange reports it).
 */
#define AFB_P2P_HANDLE_BYTES 64
AFB_API int afb_p2p_export(afb_ctx* ctx, void* values_handle, void* flags_handle);
AFB_API int afb_p2p_connect(afb_ctx* ctx, int my_rank, int nb_peer, const int32_t* peer_rank, const void* values_handles, const void* flags_handles,
                            const int64_t* pull_first, const int64_t* pull_count, const int64_t* const* slots, const int64_t* send_first, const int64_t* send_count);
AFB_API int afb_p2p_exchange(afb_ctx* ctx);
AFB_API int afb_p2p_exchange_async(afb_ctx* ctx);
AFB_API int afb_p2p_wait(afb_ctx* ctx);
AFB_API int afb_p2p_status(afb_ctx* ctx, int* status);
/* Rank skew as a number: microseconds the exchange kernels since the last call spent waiting for the neighbours' rows to
 * become ready, and for the neighbours to acknowledge their pulls (first block of each neighbour, summed over neighbours
 * and exchanges), and the number of exchanges.  Synchronises; read and clear. */
AFB_API int afb_p2p_wait_stats(afb_ctx* ctx, double* ready_wait_us, double* pulled_wait_us, int64_t* nb_exchange);
AFB_API int afb_p2p_disconnect(afb_ctx* ctx);

/*
 * Columns of the scalar CSR view in the solver's global numbering: out[j] = dof_local_to_global[col[j]]
 * (device arrays; dof_local_to_global has nb_row entries, out has nnz entries of the CSR view).
 * Replaces the host loop of HypreDoFLinearSystemImpl::solve that renumbers the columns through
 * m_dof_matrix_numbering when running in parallel (femutils/HypreDoFLinearSystem.cc:390-406).
 */
AFB_API int afb_renumber_columns(afb_ctx* ctx, const int32_t* dof_local_to_global, int32_t* out);

/*
 * The arrays of HYPRE_IJMatrixSetValues(ij_A, nrows, ncols, rows, cols, values) exactly as HypreDoFLinearSystemImpl::solve
 * passes them (femutils/HypreDoFLinearSystem.cc:501-514), as device pointers owned by the context: `ncols` = number of
 * columns of every row, `rows` = global row numbers first_own_row + i, `cols` = columns in the solver's global numbering
 * (dof_local_to_global: device array of nb_row entries, afb_xplan_numbering; NULL = sequential, columns unchanged),
 * `values` = the assembled values in place.  nb_own_row rows are handed over (the owned rows come first).  Scalar CSR view:
 * b > 1 needs the per-row value layout, like afb_get_csr_view.
 */
AFB_API int afb_get_ij_arrays(afb_ctx* ctx, int32_t first_own_row, int32_t nb_own_row, const int32_t* dof_local_to_global, const int32_t** ncols, const int32_t** rows,
                              const int32_t** cols, const double** values, int64_t* nb_values);
/* device -> host copy on the context's stream, synchronised (what UVM host access gives the reference) */
AFB_API int afb_memcpy_to_host(afb_ctx* ctx, void* dst_host, const void* src_device, size_t bytes);

/* ---- solve ---------------------------------------------------------------- */

/*
 * Jacobi-preconditioned conjugate gradient on the matrix and RHS of the context, as assembled (CSR, or BSR in either
 * value layout): stands in for the HYPRE / PETSc solve the reference hands the same arrays to
 * (femutils/HypreDoFLinearSystem.cc:461-520, PetscDoFLinearSystem.cc:329-398) so that golden solution files can be
 * checked end to end on the GPU.  SPD systems only (Poisson, elasticity; Dirichlet by penalty or elimination).
 * Stops when sqrt(r . D^-1 r) <= max(rtol * its initial value, atol) or after max_iter iterations (then an error is
 * returned, x still holds the last iterate).  x: nb_row*b doubles, host or device (mem_space); may be NULL.
 */
AFB_API int afb_solve_pcg(afb_ctx* ctx, double rtol, double atol, int max_iter, double* x, int mem_space, int* iterations, double* residual);

/* ---- instrumentation ---------------------------------------------------------------------- */

/* Milliseconds spent by the last call of each phase, measured with CUDA events on the
 * context stream (Timer::Action scopes "BuildMatrix" / "AddAndCompute" of the reference:
 * modules/testlab/CsrGpuBiliAssembly.cc:313-337).  Synchronises the stream. */
AFB_API int afb_last_timings(afb_ctx* ctx, float* connectivity_ms, float* pattern_ms, float* assemble_ms);
/* One-time cost of the tiled path's inspector for the current mesh (milliseconds, CUDA events): mesh tiling incl. the
 * tile-local node-node connectivity, and the value plan (contribution lists); -1 when not built.  The analogue of
 * the init-time connectivity the reference builds on the host (MeshUtils::computeNodeNodeViaEdgeConnectivity). */
AFB_API int afb_inspector_timings(afb_ctx* ctx, float* mesh_tiling_ms, float* value_plan_ms);
/* number of kernels launched by this context so far (bench.py "gpu_launches") */
AFB_API int64_t afb_launch_count(afb_ctx* ctx);


/* =============================================================================================
 * Domain decomposition behind the C ABI.  What the reference gets from Arcane's partitioner, from
 * Arcane's variable synchronisation and from the solver's parallel matrix (femutils/HypreDoFLinearSystem.cc:209-249:
 * allGather(nb_own_row) + ghost synchronize; :461-520: off-processor entries) is done here by the library; the host
 * program only provides two communication primitives (MPI in the reference's world, torch.distributed in bench.py,
 * shared memory in tests/cpp/mgpu_driver.cpp).
 * ============================================================================================= */
typedef struct afb_transport {
  void* user;
  int32_t rank, world;
  /* every rank contributes `bytes` bytes; recv receives world * bytes, in rank order.  Collective. */
  int (*allgather)(void* user, const void* send, int64_t bytes, void* recv);
  /* for k < nb_peer: send send_bytes[k] bytes to rank peer[k] and receive recv_bytes[k] bytes from it (sizes are known on
   * both sides; zero sizes allowed; peers are symmetric).  device_memory != 0: the buffers are device pointers. */
  int (*exchange)(void* user, int32_t nb_peer, const int32_t* peer, const void* const* send, const int64_t* send_bytes, void* const* recv, const int64_t* recv_bytes,
                  int device_memory);
  int32_t exchange_takes_device_memory; /* the exchange callback can move device buffers (e.g. NCCL): the fall-back transport then sends the matrix rows in place */
  int32_t pad;
} afb_transport;

/*
 * Mesh ingestion: Gmsh 4.1 files, binary or ASCII (host only, no context needed).  Stands in for Arcane's MshMeshReader, which every
 * .arc file of the reference selects through `<filename>meshes/....msh</filename>` (e.g. modules/testlab/inputs/Test.L-shape.2D.arc:17-21),
 * with the conventions the assembly path and the golden solution files rely on: node uniqueId = gmsh node tag, local id = rank of
 * the tag; cells = the elements of the highest dimension (one type per mesh), ordered by element tag; groups by physical name:
 * surfaces (`<surface>`: (dim-1) elements, plus their node group as modules/testlab/FemModule.cc:657-663 takes it), volumes
 * (`<material-property><volume>`) and points (`<dirichlet-point><node>`).  AFB_ERR_INVALID with the byte offset in afb_last_error
 * for anything malformed or truncated.
 */
enum { AFB_MSH_GROUP_POINTS = 0, AFB_MSH_GROUP_FACES = 1, AFB_MSH_GROUP_CELLS = 2 };
typedef struct afb_msh afb_msh;
AFB_API int afb_msh_read(const char* path, afb_msh** out);
AFB_API int afb_msh_destroy(afb_msh* m);
AFB_API int afb_msh_sizes(const afb_msh* m, int* dim, int* nodes_per_cell, int32_t* nb_node, int64_t* nb_cell, int32_t* nb_group);
/* arrays in afb_set_mesh's layout (any pointer may be NULL): xyz [nb_node*3], cell_nodes [nb_cell*nodes_per_cell] local ids, node_uid [nb_node] */
AFB_API int afb_msh_get(const afb_msh* m, double* xyz, int32_t* cell_nodes, int64_t* node_uid);
/* group g: name (owned by the mesh), kind (AFB_MSH_GROUP_*), items (faces / cells / points), nodes per item (faces), size of its node group (0 for cells) */
AFB_API int afb_msh_group(const afb_msh* m, int32_t g, const char** name, int* kind, int64_t* nb_item, int* nodes_per_item, int64_t* nb_group_node);
/* items: faces [nb_item*nodes_per_item] node ids in file order (afb_assemble_rhs_neumann's input), cells / points [nb_item] ascending ids;
 * nodes: the group's nodes, ascending (afb_set_dirichlet_nodes' input) */
AFB_API int afb_msh_group_get(const afb_msh* m, int32_t g, int32_t* items, int32_t* nodes);

/*
 * Partition of a global mesh into `world` sub-domains by recursive coordinate bisection of the cell centroids (what the
 * reference asks of Arcane's partitioner; structured boxes come out as slabs / bricks).  Semantics of an Arcane sub-domain
 * as ArcaneFEM sees it: own cells + one layer of ghost cells, every node has exactly one owner (the lowest rank among the
 * cells touching it), local numbering owned-first, ghosts grouped by ascending owner then global id.  Host arrays.
 */
typedef struct afb_partition afb_partition;
AFB_API int afb_partition_create(int dim, int nodes_per_cell, int32_t nb_node, int64_t nb_cell, const double* xyz, const int32_t* cell_nodes, int world, afb_partition** out);
AFB_API int afb_partition_destroy(afb_partition* p);
/* sizes of rank's sub-domain */
AFB_API int afb_partition_sizes(const afb_partition* p, int rank, int32_t* nb_node, int32_t* nb_own_node, int64_t* nb_cell, int64_t* nb_own_cell);
/* arrays of rank's sub-domain (any pointer may be NULL): xyz [nb_node*3], cell_nodes [nb_cell*npc] local ids, is_own [nb_node], node_gid [nb_node],
 * node_owner [nb_node], cell_gid [nb_cell] */
AFB_API int afb_partition_get(const afb_partition* p, int rank, double* xyz, int32_t* cell_nodes, uint8_t* is_own, int64_t* node_gid, int32_t* node_owner, int64_t* cell_gid);

/*
 * Host-only index logic of the ghost-row exchange of one rank (no CUDA call: also used by the CPU tests).
 * rows_tail [nb_ghost + 1]: block-row offsets (absolute, i.e. row_index[nb_own_node .. nb_node]) of the ghost rows;
 * cols_tail: local column ids of those rows (columns[rows_tail[0] .. rows_tail[nb_ghost])).  Collective over the transport.
 */
typedef struct afb_xplan_host afb_xplan_host;
AFB_API int afb_xplan_host_create(const afb_transport* t, int nb_dof_per_node, int value_layout, int32_t nb_node, int32_t nb_own_node, const int64_t* node_gid,
                                  const int32_t* node_owner, const int32_t* rows_tail, const int32_t* cols_tail, afb_xplan_host** out);
AFB_API int afb_xplan_host_destroy(afb_xplan_host* h);
/* neighbours (ascending rank): the slice [send_first, +send_count) of this rank's values array each one receives, and the
 * number of doubles it sends here */
AFB_API int afb_xplan_host_peers(const afb_xplan_host* h, int32_t* nb_peer, const int32_t** peer, const int64_t** send_first, const int64_t** send_count, const int64_t** recv_count);
/* (dof_row, dof_col) of every double neighbour k sends, in ITS memory order, as local DoF ids of this rank */
AFB_API int afb_xplan_host_pairs(const afb_xplan_host* h, int32_t k, const int32_t** dof_rows, const int32_t** dof_cols);
/* global numbering of the rows (HypreDoFLinearSystemImpl::_computeMatrixNumeration): first_dof [world + 1], dof_l2g [nb_node * b].  Collective. */
AFB_API int afb_xplan_host_numbering(afb_xplan_host* h, int64_t* first_dof, int32_t* dof_l2g);

/*
 * The exchange itself on a context (after afb_build_pattern and a first afb_assemble_bilinear with
 * AFB_FLAG_OWN_CELLS_ONLY | AFB_FLAG_ALL_ROWS, which fixes the value layout).  Collective.  Transport of the rows:
 * one pull kernel over peer memory (NVLink; CUDA IPC between processes, plain pointers inside one) when every rank can map
 * its neighbours -- agreed on collectively -- else the transport's exchange callback + accumulate kernels.
 */
typedef struct afb_xplan afb_xplan;
AFB_API int afb_xplan_create(afb_ctx* ctx, const afb_transport* t, const int64_t* node_gid, const int32_t* node_owner, int32_t nb_own_node, int allow_peer_memory, afb_xplan** out);
AFB_API int afb_xplan_destroy(afb_xplan* x);
AFB_API int afb_xplan_exchange(afb_xplan* x);   /* after every such assembly; the peer-memory kernel runs on a side stream */
AFB_API int afb_xplan_wait(afb_xplan* x);       /* orders the context stream after the exchange; reports a timed-out exchange */
AFB_API int afb_xplan_numbering(afb_xplan* x, int64_t* first_dof, int32_t* dof_l2g);
/* transport_kind: 0 none (single rank), 1 peer memory, 2 callback; bytes per exchange */
AFB_API int afb_xplan_info(const afb_xplan* x, int32_t* nb_peer, int64_t* bytes_sent, int64_t* bytes_received, int32_t* transport_kind, const char** why_not_peer_memory);

#ifdef __cplusplus
}
#endif
#endif /* AFB200_H */
