// Hand-off of the assembled arrays to the reference's solver back-ends, compile-time optional:
//   AFB_HAVE_HYPRE   HypreDoFLinearSystemImpl::solve's matrix set-up (femutils/HypreDoFLinearSystem.cc:366-514):
//                    IJMatrixCreate / SetObjectType(PARCSR) / Initialize_v2(device memory) / SetValues(nrows, ncols, rows, cols,
//                    values) with DEVICE pointers / Assemble / GetObject -- the arrays come straight from the context
//                    (afb_get_ij_arrays): no host copy, no host renumbering loop
//   AFB_HAVE_PETSC   PetscDoFLinearSystemImpl's COO hand-off (femutils/PetscDoFLinearSystem.cc:329-345,398):
//                    MatSetPreallocationCOOLocal(nnz, coo_rows, coo_cols) + MatSetValuesCOO(values, INSERT_VALUES)
// Include HYPRE.h + HYPRE_IJ_mv.h / petscmat.h BEFORE this header (tests/cpp/mock_solvers.h provides the same names for the
// image without the libraries and records the arguments, so the argument layout is checked on the GPU by the test-suite).
// HYPRE_Int / HYPRE_BigInt / PetscInt must be 32-bit, as the reference assumes when it passes its Int32 arrays (:501-514).
#pragma once

#include <stdexcept>
#include <string>

#include "afb200.h"

namespace arcanefem_b200 {

inline void handoffCheck(int rc, const char* what)
{
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + afb_last_error());
}

#ifdef AFB_HAVE_HYPRE
// Rows [first_own_row, first_own_row + nb_own_row) of the global matrix belong to this rank (afb_xplan_numbering: first_dof);
// dof_local_to_global (device, nb_row entries) is NULL in sequential runs.  Returns the assembled IJ matrix.
inline HYPRE_IJMatrix hypreSetCSRValues(afb_ctx* ctx, MPI_Comm comm, int first_own_row, int nb_own_row, const int32_t* dof_local_to_global, HYPRE_ParCSRMatrix* parcsr_out)
{
  static_assert(sizeof(HYPRE_Int) == 4 && sizeof(HYPRE_BigInt) == 4, "the reference passes Int32 arrays: HYPRE without --enable-bigint / mixedint");
  const int32_t *ncols = nullptr, *rows = nullptr, *cols = nullptr;
  const double* values = nullptr;
  int64_t nb_values = 0;
  handoffCheck(afb_get_ij_arrays(ctx, first_own_row, nb_own_row, dof_local_to_global, &ncols, &rows, &cols, &values, &nb_values), "afb_get_ij_arrays");
  handoffCheck(afb_synchronize(ctx), "afb_synchronize"); // HYPRE works on its own stream
  HYPRE_IJMatrix ij_A = nullptr;
  const int first_row = first_own_row, last_row = first_own_row + nb_own_row - 1;
  if (HYPRE_IJMatrixCreate(comm, first_row, last_row, first_row, last_row, &ij_A)) throw std::runtime_error("HYPRE_IJMatrixCreate");
  if (HYPRE_IJMatrixSetObjectType(ij_A, HYPRE_PARCSR)) throw std::runtime_error("HYPRE_IJMatrixSetObjectType");
  if (HYPRE_IJMatrixInitialize_v2(ij_A, HYPRE_MEMORY_DEVICE)) throw std::runtime_error("HYPRE_IJMatrixInitialize_v2");
  // GPU pointers; efficient in large chunks (one call for all owned rows, as the reference does)
  if (HYPRE_IJMatrixSetValues(ij_A, nb_own_row, const_cast<HYPRE_Int*>(reinterpret_cast<const HYPRE_Int*>(ncols)), reinterpret_cast<const HYPRE_BigInt*>(rows),
                              reinterpret_cast<const HYPRE_BigInt*>(cols), values))
    throw std::runtime_error("HYPRE_IJMatrixSetValues");
  if (HYPRE_IJMatrixAssemble(ij_A)) throw std::runtime_error("HYPRE_IJMatrixAssemble");
  if (parcsr_out && HYPRE_IJMatrixGetObject(ij_A, reinterpret_cast<void**>(parcsr_out))) throw std::runtime_error("HYPRE_IJMatrixGetObject");
  return ij_A;
}
#endif

#ifdef AFB_HAVE_PETSC
// `mat` has its sizes and local-to-global mapping set (MatSetSizes / MatSetLocalToGlobalMapping, :322-325); local (row, col)
// indices are handed over, exactly the reference's _translateCSRToCOO + column copy, without either kernel: the COO row
// array is kept by the context.
inline void petscSetCOOValues(afb_ctx* ctx, Mat mat, bool preallocate)
{
  static_assert(sizeof(PetscInt) == 4, "the reference copies Int32 columns into PetscInt arrays: 32-bit PetscInt");
  const int32_t *coo_rows = nullptr, *coo_cols = nullptr;
  double* values = nullptr;
  int64_t nnz = 0;
  handoffCheck(afb_get_coo(ctx, &coo_rows, &coo_cols, &values, &nnz), "afb_get_coo");
  handoffCheck(afb_synchronize(ctx), "afb_synchronize");
  if (preallocate) {
    if (MatSetPreallocationCOOLocal(mat, (PetscCount)nnz, const_cast<PetscInt*>(reinterpret_cast<const PetscInt*>(coo_rows)), const_cast<PetscInt*>(reinterpret_cast<const PetscInt*>(coo_cols))))
      throw std::runtime_error("MatSetPreallocationCOOLocal");
  }
  if (MatSetValuesCOO(mat, values, INSERT_VALUES)) throw std::runtime_error("MatSetValuesCOO");
  if (MatAssemblyBegin(mat, MAT_FINAL_ASSEMBLY) || MatAssemblyEnd(mat, MAT_FINAL_ASSEMBLY)) throw std::runtime_error("MatAssembly");
}
#endif

} // namespace arcanefem_b200
