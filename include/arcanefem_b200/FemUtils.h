// arcanefem_b200/FemUtils.h -- C++ host side of the B200 assembly path, above the C ABI (afb200.h).
//
// The classes keep the names, method names, argument meaning and error behaviour of the ArcaneFEM
// containers they stand in for (paths relative to the ArcaneFEM source root):
//
//   CsrFormatMatrixView  femutils/CsrFormatMatrixView.h:135-210 (+ CsrRowColumnIndex / CsrRowColumnIterator / CsrRow :39-126)
//   Real4, RealMatrix<N,M>, RealVector<N>   femutils/FemUtils.h:35-598  (in FemTypes.h)
//   CsrFormat            femutils/CsrFormatMatrix.h:37-142, CsrFormatMatrix.cc:35-121
//   CooFormat            femutils/CooFormatMatrix.h:38-306
//   BSRMatrix, BSRFormat femutils/BSRFormat.h:77-249, BSRFormat.cc:48-106,350-395
//   DoFLinearSystem      femutils/DoFLinearSystem.h:154-409 (the part the assembly path touches:
//                        setCSRValues / hasSetCSRValues / clearValues / eliminateRow / eliminateRowColumn /
//                        applyMatrixTransformation / applyRHSTransformation; solve() stays with HYPRE/PETSc)
//
// What differs, and why:
//   * Arcane types are replaced by plain arrays: `MeshArrays` instead of IMesh* + connectivity views,
//     Int32 DoF ids instead of DoFLocalId (dof = node_lid * nb_dof + component,
//     femutils/FemDoFsOnNodes.cc:79-111).
//   * the element physics is an operator tag (`Operator`) instead of a device lambda template parameter
//     (femutils/BSRFormat.h:218-236): a lambda cannot cross a C ABI.
//   * arrays live in HBM; the views hand out DEVICE pointers laid out exactly as the reference hands them
//     to HYPRE_IJMatrixSetValues / MatSetValuesCOO (femutils/HypreDoFLinearSystem.cc:501-514,
//     femutils/PetscDoFLinearSystem.cc:329-345,398).  Host access goes through copyToHost()/getValue().
//   * errors: the reference throws (ARCANE_FATAL / ARCANE_THROW); here every non-zero C status becomes a
//     FatalError carrying afb_last_error().
//   * no CPU fallback: without a CUDA device the Context constructor throws.
// Header-only; link with libafb200.so.
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "afb200.h"
#include "arcanefem_b200/FemTypes.h"

namespace arcanefem_b200 {

using Int8 = std::int8_t;
using Int32 = std::int32_t;
using Int64 = std::int64_t;
// Real is declared in FemTypes.h

//! ARCANE_FATAL / ARCANE_THROW equivalent
class FatalError : public std::runtime_error {
 public:
  FatalError(int code, const std::string& msg) : std::runtime_error(msg), m_code(code) {}
  int code() const { return m_code; }

 private:
  int m_code;
};

inline void check(int rc)
{
  if (rc != AFB_OK) throw FatalError(rc, afb_last_error());
}

//! replaces the device lambda handed to BSRFormat::assembleBilinear*
enum class Operator { Poisson = AFB_OP_POISSON, Elasticity = AFB_OP_ELASTICITY, Bilaplacian = AFB_OP_BILAPLACIAN, DiffusionReaction = AFB_OP_DIFFUSION_REACTION,
                      Elastodynamics = AFB_OP_ELASTODYNAMICS };

//! femutils/FemUtilsGlobal.h:51-62
enum class eMatrixEliminationType { None = AFB_ELIMINATE_NONE, Row = AFB_ELIMINATE_ROW, RowColumn = AFB_ELIMINATE_ROW_COLUMN };

//! what the back-ends read from IMesh*: Real3 coordinates (AoS), cell -> node local ids, node ownership
struct MeshArrays {
  int dim = 3;
  int nodes_per_cell = 4;
  Int32 nb_node = 0;
  Int64 nb_cell = 0;
  const Real* coords = nullptr;          // [nb_node][3]
  const Int32* cell_nodes = nullptr;     // [nb_cell][nodes_per_cell]
  const std::uint8_t* node_is_own = nullptr; // [nb_node] or null (all owned)
  Int64 nb_own_cell = -1;                // cells [0,nb_own_cell) are own, the rest ghost cells; -1: all own
  int mem_space = AFB_MEM_HOST;          // AFB_MEM_DEVICE: the pointers are device pointers (zero copy)
};

//! RunQueue + memory resource of the reference back-ends: one per GPU
class Context {
 public:
  explicit Context(int device = 0) { check(afb_create(device, &m_ctx)); }
  ~Context() { afb_destroy(m_ctx); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  afb_ctx* handle() const { return m_ctx; }
  void setStream(void* cuda_stream) { check(afb_set_stream(m_ctx, cuda_stream)); }
  void barrier() { check(afb_synchronize(m_ctx)); } // RunQueue::barrier()
  void setMesh(const MeshArrays& m)
  {
    check(afb_set_mesh(m_ctx, m.dim, m.nodes_per_cell, m.nb_node, m.nb_cell, m.coords, m.cell_nodes, m.node_is_own, m.mem_space));
    if (m.nb_own_cell >= 0) check(afb_set_own_cell_count(m_ctx, m.nb_own_cell));
  }
  template <class T> std::vector<T> copyToHost(int which)
  {
    size_t bytes = 0;
    check(afb_copy_to_host(m_ctx, which, nullptr, &bytes));
    std::vector<T> out(bytes / sizeof(T));
    check(afb_copy_to_host(m_ctx, which, out.data(), &bytes));
    return out;
  }

 private:
  afb_ctx* m_ctx = nullptr;
};

//! SmallSpan of a device array
template <class T> struct DeviceSpan {
  T* ptr = nullptr;
  Int64 n = 0;
  AFB_HOST_DEVICE T* data() const { return ptr; }
  AFB_HOST_DEVICE Int64 size() const { return n; }
  AFB_HOST_DEVICE T& operator[](Int64 i) const { return ptr[i]; } // dereferences DEVICE memory: for kernels over the view
};

/*---------------------------------------------------------------------------*/
//! Position of one stored entry in the columns / values arrays (femutils/CsrFormatMatrixView.h:39-62); -1 = none
class CsrRowColumnIndex {
 public:
  using IndexType = Int32;
  CsrRowColumnIndex() = default;
  AFB_HOST_DEVICE explicit constexpr CsrRowColumnIndex(IndexType index) : m_index(index) {}
  AFB_HOST_DEVICE constexpr IndexType value() const { return m_index; }
  AFB_HOST_DEVICE constexpr operator IndexType() const { return m_index; }
  AFB_HOST_DEVICE constexpr bool isNull() const { return m_index < 0; }

 private:
  IndexType m_index = -1;
};
//! Walks the entries of one row (femutils/CsrFormatMatrixView.h:67-96)
class CsrRowColumnIterator {
 public:
  CsrRowColumnIterator() = default;
  AFB_HOST_DEVICE explicit constexpr CsrRowColumnIterator(Int32 index) : m_index(index) {}
  AFB_HOST_DEVICE constexpr CsrRowColumnIndex operator*() const { return CsrRowColumnIndex(m_index); }
  AFB_HOST_DEVICE constexpr CsrRowColumnIterator& operator++()
  {
    ++m_index;
    return *this;
  }
  friend AFB_HOST_DEVICE constexpr bool operator!=(const CsrRowColumnIterator& a, const CsrRowColumnIterator& b) { return a.m_index != b.m_index; }
  AFB_HOST_DEVICE constexpr bool isValid() const { return m_index != -1; }

 private:
  Int32 m_index = -1;
};
//! The entries [begin, end) of one row, for range-based for (femutils/CsrFormatMatrixView.h:101-126)
class CsrRow {
 public:
  CsrRow() = default;
  AFB_HOST_DEVICE constexpr CsrRow(Int32 begin, Int32 end) : m_begin(begin), m_end(end) {}
  AFB_HOST_DEVICE constexpr CsrRowColumnIterator begin() const { return CsrRowColumnIterator(m_begin); }
  AFB_HOST_DEVICE constexpr CsrRowColumnIterator end() const { return CsrRowColumnIterator(m_end); }
  AFB_HOST_DEVICE constexpr Int32 size() const { return m_end - m_begin; }

 private:
  Int32 m_begin = -1, m_end = -1;
};

/*---------------------------------------------------------------------------*/
//! femutils/CsrFormatMatrixView.h:135-210 (device pointers; rows has nbRow entries, no sentinel)
class CsrFormatMatrixView {
 public:
  CsrFormatMatrixView() = default;
  CsrFormatMatrixView(DeviceSpan<const Int32> rows, DeviceSpan<const Int32> rows_nb_column, DeviceSpan<const Int32> columns, DeviceSpan<Real> values)
  : m_matrix_rows(rows), m_matrix_rows_nb_column(rows_nb_column), m_matrix_columns(columns), m_values(values) {}
  DeviceSpan<const Int32> rows() const { return m_matrix_rows; }
  DeviceSpan<const Int32> rowsNbColumn() const { return m_matrix_rows_nb_column; }
  DeviceSpan<const Int32> columns() const { return m_matrix_columns; }
  DeviceSpan<Real> values() const { return m_values; }
  AFB_HOST_DEVICE Int32 nbRow() const { return (Int32)m_matrix_rows.size(); }
  AFB_HOST_DEVICE Int32 nbColumn() const { return (Int32)m_matrix_columns.size(); }
  AFB_HOST_DEVICE Int32 nbValue() const { return (Int32)m_values.size(); }
  // Element access as in the reference (femutils/CsrFormatMatrixView.h:171-215).  The spans point into HBM: these are for
  // kernels that receive the view by value (or for a view built over host copies, as tests/cpp/types_driver.cpp does).
  AFB_HOST_DEVICE Int32 row(Int32 index) const { return m_matrix_rows[index]; }
  AFB_HOST_DEVICE Int32 nbColumnForRow(Int32 row) const { return m_matrix_rows_nb_column[row]; }
  AFB_HOST_DEVICE Int32 column(CsrRowColumnIndex rc) const { return m_matrix_columns[rc.value()]; }
  AFB_HOST_DEVICE Real& value(CsrRowColumnIndex rc) const { return m_values[rc.value()]; }
  //! entries of `row`: rows() has no sentinel, the last row ends at nbColumn()
  AFB_HOST_DEVICE CsrRow rowRange(Int32 row) const { return CsrRow(m_matrix_rows[row], row + 1 == nbRow() ? nbColumn() : m_matrix_rows[row + 1]); }
  //! linear search of `column_id` in `row`; null index when absent
  AFB_HOST_DEVICE CsrRowColumnIndex tryFindColumnInRow(Int32 row, Int32 column_id) const
  {
    for (CsrRowColumnIndex rc : rowRange(row))
      if (column(rc) == column_id) return rc;
    return CsrRowColumnIndex();
  }

 private:
  DeviceSpan<const Int32> m_matrix_rows, m_matrix_rows_nb_column, m_matrix_columns;
  DeviceSpan<Real> m_values;
};
using CSRFormatView = CsrFormatMatrixView;

/*---------------------------------------------------------------------------*/
//! The part of DoFLinearSystem the assembly path talks to (femutils/DoFLinearSystem.h:154-409).
class DoFLinearSystem {
 public:
  explicit DoFLinearSystem(Context& ctx) : m_ctx(ctx) {}
  //! The view must remain valid until solve() (femutils/DoFLinearSystem.h:318-325)
  void setCSRValues(const CSRFormatView& view) { m_view = view; m_has_view = true; }
  bool hasSetCSRValues() const { return true; } // a CSR-consuming implementation (Hypre/PETSc/Alien)
  const CSRFormatView& getCSRValues() const { return m_view; }
  void clearValues()
  {
    m_has_view = false;
    m_view = CSRFormatView();
    check(afb_clear_dirichlet(m_ctx.handle()));
  }
  //! DoFLinearSystem::matrixAddValue / matrixSetValue (femutils/DoFLinearSystem.h:186-194): one entry of the assembled matrix,
  //! DoF ids; an entry outside the pattern is an error (the reference's CSR implementation has no slot for it either)
  void matrixAddValue(Int32 row, Int32 column, Real value) { check(afb_matrix_set_value(m_ctx.handle(), row, column, value, 1)); }
  void matrixSetValue(Int32 row, Int32 column, Real value) { check(afb_matrix_set_value(m_ctx.handle(), row, column, value, 0)); }
  Real matrixGetValue(Int32 row, Int32 column)
  {
    Real v = 0.0;
    check(afb_matrix_get_value(m_ctx.handle(), row, column, &v));
    return v;
  }
  void eliminateRow(Int32 dof, Real value) { check(afb_set_elimination(m_ctx.handle(), AFB_ELIMINATE_ROW, 1, &dof, &value, AFB_MEM_HOST)); }
  void eliminateRowColumn(Int32 dof, Real value) { check(afb_set_elimination(m_ctx.handle(), AFB_ELIMINATE_ROW_COLUMN, 1, &dof, &value, AFB_MEM_HOST)); }
  void setForcedValues(Int32 n, const Int32* dofs, const Real* values) { check(afb_set_forced_values(m_ctx.handle(), n, dofs, values, AFB_MEM_HOST)); }
  //! CsrDoFLinearSystemImpl::applyMatrixTransformation / applyRHSTransformation (femutils/CsrDoFLinearSystemImpl.cc:235-253)
  void applyMatrixTransformation(bool replicate_column0_quirk = true) { check(afb_apply_matrix_transformation(m_ctx.handle(), replicate_column0_quirk ? 1 : 0)); }
  void applyRHSTransformation() { check(afb_apply_rhs_transformation(m_ctx.handle())); }
  DeviceSpan<Real> rhs()
  {
    Real* p = nullptr;
    Int32 n = 0;
    check(afb_get_rhs(m_ctx.handle(), &p, &n));
    return { p, n };
  }
  bool hasView() const { return m_has_view; }
  //! DoFLinearSystem::solve (femutils/DoFLinearSystem.cc; HypreDoFLinearSystem.cc:461-520 behind it): here a Jacobi-PCG on the
  //! arrays as assembled -- a stand-in for tests; a production build hands getCSRValues() to HYPRE / PETSc.
  //! solution_host: nb_row*b doubles.  Returns the number of iterations.
  Int32 solve(Real* solution_host, Real rtol = 1.0e-10, Real atol = 0.0, Int32 max_iter = 10000)
  {
    int it = 0;
    double res = 0.0;
    check(afb_solve_pcg(m_ctx.handle(), rtol, atol, max_iter, solution_host, AFB_MEM_HOST, &it, &res));
    return it;
  }

 private:
  Context& m_ctx;
  CSRFormatView m_view;
  bool m_has_view = false;
};

/*---------------------------------------------------------------------------*/
//! RHS terms of femutils/ArcaneFemFunctionsGpu.h (BoundaryConditions::applyConstantSourceToRhs :675-708, applyNeumannToRhs,
//! ArcaneFemFunctions.h applyTractionToRhs*): groups become plain face lists (faceNode order, outward-normal swap applied)
namespace BoundaryConditions {
inline void applyConstantSourceToRhs(Context& ctx, const Real* f, int nb_component, bool nodewise = true)
{
  check(afb_assemble_rhs_source(ctx.handle(), f, nb_component, nodewise ? 1 : 0, 0));
}
inline void applyNeumannToRhs(Context& ctx, Int64 nb_face, const Int32* face_nodes, int nb_value, const Real* values, bool skip_dirichlet_nodes = false)
{
  check(afb_assemble_rhs_neumann(ctx.handle(), nb_face, face_nodes, AFB_NEUMANN_FLUX, nb_value, values, skip_dirichlet_nodes ? 1 : 0, AFB_MEM_HOST));
}
inline void applyTractionToRhs(Context& ctx, Int64 nb_face, const Int32* face_nodes, int nb_dof_per_node, const Real* traction)
{
  check(afb_assemble_rhs_neumann(ctx.handle(), nb_face, face_nodes, AFB_NEUMANN_TRACTION, nb_dof_per_node, traction, 0, AFB_MEM_HOST));
}
} // namespace BoundaryConditions

/*---------------------------------------------------------------------------*/
//! femutils/CsrFormatMatrix.h:37-142 (one DoF per node: the testlab csr / csr-gpu / nwcsr back-ends)
class CsrFormat {
 public:
  explicit CsrFormat(Context& ctx) : m_ctx(ctx) {}
  //! CsrFormat::initialize + FemModuleTestlab::_computeSparsity (modules/testlab/CsrGpuBiliAssembly.cc:187-207):
  //! allocation AND sparsity in one step (the reference re-does both on every assembly)
  void initialize(const MeshArrays& mesh)
  {
    m_ctx.setMesh(mesh);
    computeSparsity();
  }
  void computeSparsity()
  {
    Int32 nb_row = 0;
    Int64 nnz = 0;
    check(afb_build_pattern(m_ctx.handle(), 1, &nb_row, &nnz));
    m_nb_row = nb_row;
    m_nnz = (Int32)nnz;
  }
  //! _assembleCsrGPUBilinearOperator{TRIA3,TETRA4} (csr-gpu), _assembleNodeWiseCsrBilinearOperator* (nwcsr)
  void assembleBilinear(Operator op, bool atomic_free, bool tiled = true, int flags = AFB_FLAG_SIGNED_TRI_AREA)
  {
    const int variant = !atomic_free ? AFB_VARIANT_CELLWISE_ATOMIC : (tiled ? AFB_VARIANT_TILED_GATHER : AFB_VARIANT_NODEWISE);
    check(afb_assemble_bilinear(m_ctx.handle(), (int)op, nullptr, 0, AFB_FORMAT_CSR, variant, AFB_LAYOUT_PER_BLOCK, flags));
  }
  void matrixAddValue(Int32 row, Int32 column, Real value)
  {
    if (value == 0.0) return; // femutils/CsrFormatMatrix.h:64
    check(afb_matrix_set_value(m_ctx.handle(), row, column, value, 1));
  }
  void matrixSetValue(Int32 row, Int32 column, Real value) { check(afb_matrix_set_value(m_ctx.handle(), row, column, value, 0)); }
  Real getValue(Int32 row, Int32 column)
  {
    Real v = 0;
    check(afb_matrix_get_value(m_ctx.handle(), row, column, &v));
    return v;
  }
  CsrFormatMatrixView view()
  {
    const Int32 *rows = nullptr, *nbc = nullptr, *cols = nullptr;
    Real* vals = nullptr;
    Int32 nb_row = 0;
    Int64 nnz = 0;
    check(afb_get_csr_view(m_ctx.handle(), &rows, &nbc, &cols, &vals, &nb_row, &nnz));
    return CsrFormatMatrixView({ rows, nb_row }, { nbc, nb_row }, { cols, nnz }, { vals, nnz });
  }
  //! CsrFormat::translateToLinearSystem (femutils/CsrFormatMatrix.cc:63-111): zero-copy hand-off
  void translateToLinearSystem(DoFLinearSystem& linear_system) { linear_system.setCSRValues(view()); }
  Int32 nbRow() const { return m_nb_row; }
  Int32 m_nnz = 0;

 private:
  Context& m_ctx;
  Int32 m_nb_row = 0;
};

/*---------------------------------------------------------------------------*/
//! femutils/CooFormatMatrix.h:38-306 (coo-gpu / coo-sorting-gpu back-ends): rows expanded on the device
class CooFormat {
 public:
  explicit CooFormat(Context& ctx) : m_ctx(ctx) {}
  void initialize(const MeshArrays& mesh)
  {
    m_ctx.setMesh(mesh);
    Int32 nb_row = 0;
    Int64 nnz = 0;
    check(afb_build_pattern(m_ctx.handle(), 1, &nb_row, &nnz));
    m_nnz = (Int32)nnz;
  }
  void assembleBilinear(Operator op, int flags = AFB_FLAG_SIGNED_TRI_AREA)
  {
    check(afb_assemble_bilinear(m_ctx.handle(), (int)op, nullptr, 0, AFB_FORMAT_COO, AFB_VARIANT_CELLWISE_ATOMIC, AFB_LAYOUT_PER_BLOCK, flags));
  }
  //! row / column / value arrays = MatSetPreallocationCOOLocal + MatSetValuesCOO layout
  void arrays(DeviceSpan<const Int32>& row, DeviceSpan<const Int32>& col, DeviceSpan<Real>& val)
  {
    const Int32 *r = nullptr, *c = nullptr;
    Real* v = nullptr;
    Int64 nnz = 0;
    check(afb_get_coo(m_ctx.handle(), &r, &c, &v, &nnz));
    row = { r, nnz };
    col = { c, nnz };
    val = { v, nnz };
  }
  Int32 m_nnz = 0;

 private:
  Context& m_ctx;
};

/*---------------------------------------------------------------------------*/
//! femutils/BSRFormat.h:77-137
class BSRMatrix {
  friend class BSRFormat;

 public:
  explicit BSRMatrix(Context& ctx) : m_ctx(ctx) {}
  Real getValue(Int32 row, Int32 col) const
  {
    Real v = 0;
    check(afb_matrix_get_value(m_ctx.handle(), row, col, &v));
    return v;
  }
  void setValue(Int32 row, Int32 col, Real value) { check(afb_matrix_set_value(m_ctx.handle(), row, col, value, 0)); }
  void addValue(Int32 row, Int32 col, Real value) { check(afb_matrix_set_value(m_ctx.handle(), row, col, value, 1)); }
  bool orderValuePerBlock() const { return m_order_values_per_block; }
  Int32 nbNonZero() const { return m_nb_non_zero_value; }
  Int32 nbColumn() const { return m_nb_col; }
  Int32 nbRow() const { return m_nb_row; }
  Int8 nbBlock() const { return m_nb_block; }
  DeviceSpan<Real> values() const { return m_values; }
  DeviceSpan<const Int32> columns() const { return m_columns; }
  DeviceSpan<const Int32> rowsIndex() const { return m_rows_index; }
  DeviceSpan<const Int32> nbNonZeroPerRows() const { return m_nb_non_zero_per_rows; }
  //! BSRMatrix::toCsr (femutils/BSRFormat.cc:110-172): expanded on the device, values shared
  CsrFormatMatrixView toCsr()
  {
    const Int32 *rows = nullptr, *nbc = nullptr, *cols = nullptr;
    Real* vals = nullptr;
    Int32 nb_row = 0;
    Int64 nnz = 0;
    check(afb_get_csr_view(m_ctx.handle(), &rows, &nbc, &cols, &vals, &nb_row, &nnz));
    return CsrFormatMatrixView({ rows, nb_row }, { nbc, nb_row }, { cols, nnz }, { vals, nnz });
  }
  //! BSRMatrix::dump (femutils/BSRFormat.cc:202-225): same text layout
  void dump(const std::string& filename, Context& ctx)
  {
    auto v = ctx.copyToHost<Real>(AFB_ARRAY_VALUES);
    auto c = ctx.copyToHost<Int32>(AFB_ARRAY_COLUMNS);
    auto r = ctx.copyToHost<Int32>(AFB_ARRAY_ROWS);
    FILE* f = std::fopen(filename.c_str(), "w");
    if (!f) throw FatalError(AFB_ERR_INVALID, "cannot open " + filename);
    std::fprintf(f, "size :%d\n", m_nb_row);
    for (size_t i = 0; i + 1 < r.size(); ++i) std::fprintf(f, "%d ", r[i]);
    std::fprintf(f, "\n");
    for (Int32 x : c) std::fprintf(f, "%d ", x);
    std::fprintf(f, "\n");
    for (Real x : v) std::fprintf(f, "%.17g ", x);
    std::fprintf(f, "\n");
    std::fclose(f);
  }

 private:
  void _refresh()
  {
    const Int32 *rows = nullptr, *cols = nullptr, *nz = nullptr;
    Real* vals = nullptr;
    Int32 nbr = 0;
    Int64 nb_col = 0;
    int b = 1, layout = 0;
    check(afb_get_bsr(m_ctx.handle(), &rows, &cols, &vals, &nz, &nbr, &nb_col, &b, &layout));
    m_nb_row = nbr;
    m_nb_col = (Int32)nb_col;
    m_nb_block = (Int8)b;
    m_nb_non_zero_value = (Int32)(nb_col * b * b);
    m_values = { vals, nb_col * b * b };
    m_columns = { cols, nb_col };
    m_rows_index = { rows, nbr };
    m_nb_non_zero_per_rows = { nz, nbr };
  }
  Context& m_ctx;
  bool m_order_values_per_block = true;
  Int32 m_nb_non_zero_value = 0, m_nb_col = 0, m_nb_row = 0;
  Int8 m_nb_block = 1;
  DeviceSpan<Real> m_values;
  DeviceSpan<const Int32> m_columns, m_rows_index, m_nb_non_zero_per_rows;
};

/*---------------------------------------------------------------------------*/
//! femutils/BSRFormat.h:157-249
class BSRFormat {
 public:
  explicit BSRFormat(Context& ctx) : m_ctx(ctx), m_bsr_matrix(ctx) {}
  //! BSRFormat::initialize (femutils/BSRFormat.cc:350-372)
  void initialize(const MeshArrays& mesh, Int8 nb_dof, bool does_linear_system_use_csr, bool use_atomic_free = false)
  {
    if (mesh.dim != 2 && mesh.dim != 3) throw FatalError(AFB_ERR_UNSUPPORTED, "BSRFormat(initialize): Only supports 2D and 3D");
    m_nb_dof = nb_dof;
    m_use_csr_in_linear_system = does_linear_system_use_csr;
    m_use_atomic_free = use_atomic_free;
    m_bsr_matrix.m_order_values_per_block = !does_linear_system_use_csr; // femutils/BSRFormat.cc:367
    m_ctx.setMesh(mesh);
    m_initialized = true;
  }
  void computeSparsity()
  {
    if (!m_initialized) throw FatalError(AFB_ERR_INVALID, "BSRFormat(computeSparsity): initialize() first");
    Int32 nb_row = 0;
    Int64 nnz = 0;
    check(afb_build_pattern(m_ctx.handle(), m_nb_dof, &nb_row, &nnz));
    m_bsr_matrix._refresh();
  }
  //! from the cells (femutils/BSRFormat.cc:799-1006) / from the init-time node-node connectivity (:445-790);
  //! both emit the same deterministic pattern (columns ascending)
  void computeSparsityAtomic()
  {
    check(afb_set_sparsity_algorithm(m_ctx.handle(), AFB_SPARSITY_FROM_CELLS));
    computeSparsity();
  }
  void computeSparsityAtomicFree()
  {
    check(afb_set_sparsity_algorithm(m_ctx.handle(), AFB_SPARSITY_FROM_CONNECTIVITY));
    computeSparsity();
  }
  //! assembleBilinearAtomic(lambda cell -> RealMatrix): cell-wise, fp64 atomics
  void assembleBilinearAtomic(Operator op, const Real* params = nullptr, int nb_params = 0) { _assemble(op, params, nb_params, AFB_VARIANT_CELLWISE_ATOMIC); }
  //! assembleBilinearAtomicFree(lambda (cell, i) -> RealMatrix<1,n>): every row written once; B200: tiled gather
  void assembleBilinearAtomicFree(Operator op, const Real* params = nullptr, int nb_params = 0)
  {
    try {
      _assemble(op, params, nb_params, AFB_VARIANT_TILED_GATHER);
    }
    catch (const FatalError& e) {
      if (e.code() != AFB_ERR_UNSUPPORTED) throw;
      _assemble(op, params, nb_params, AFB_VARIANT_NODEWISE); // P2 cells, bilaplacian: thread-per-row kernel
    }
  }
  void assembleBilinear(Operator op, const Real* params = nullptr, int nb_params = 0)
  {
    if (m_use_atomic_free) assembleBilinearAtomicFree(op, params, nb_params);
    else assembleBilinearAtomic(op, params, nb_params);
  }
  BSRMatrix& matrix() { return m_bsr_matrix; }
  void resetMatrixValues() { check(afb_reset_values(m_ctx.handle())); }
  //! BSRFormat::toLinearSystem (femutils/BSRFormat.cc:377-395)
  void toLinearSystem(DoFLinearSystem& linear_system)
  {
    if (!m_use_csr_in_linear_system)
      throw FatalError(AFB_ERR_INVALID, "BSRFormat(toLinearSystem): Linear system was set to use CSR but is incompatible");
    linear_system.setCSRValues(m_bsr_matrix.toCsr());
  }
  void dumpMatrix(const std::string& filename) { m_bsr_matrix.dump(filename, m_ctx); }

 private:
  void _assemble(Operator op, const Real* params, int nb_params, int variant)
  {
    const int layout = m_bsr_matrix.m_order_values_per_block ? AFB_LAYOUT_PER_BLOCK : AFB_LAYOUT_PER_ROW;
    check(afb_assemble_bilinear(m_ctx.handle(), (int)op, params, nb_params, AFB_FORMAT_BSR, variant, layout, 0));
    m_bsr_matrix._refresh();
  }
  Context& m_ctx;
  Int8 m_nb_dof = 1;
  bool m_use_csr_in_linear_system = false, m_use_atomic_free = false, m_initialized = false;
  BSRMatrix m_bsr_matrix;
};

} // namespace arcanefem_b200
