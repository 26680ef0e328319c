// Stand-ins for HYPRE.h / HYPRE_IJ_mv.h / petscmat.h in an image without the libraries: the names and signatures
// include/arcanefem_b200/SolverHandoff.h compiles against (HYPRE 2.27+, PETSc 3.18+), recording every argument so that the
// test driver can check the hand-off (device pointers, array layout) against the oracle.  Test infrastructure only.
#pragma once
#include <cstdint>
#include <vector>

typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
typedef int HYPRE_Int;
typedef int HYPRE_BigInt;
typedef double HYPRE_Complex;
typedef int HYPRE_MemoryLocation;
#define HYPRE_MEMORY_HOST 0
#define HYPRE_MEMORY_DEVICE 1
#define HYPRE_PARCSR 5555
struct hypre_IJMatrix_struct {
  MPI_Comm comm;
  HYPRE_BigInt ilower, iupper, jlower, jupper;
  HYPRE_Int object_type = -1;
  HYPRE_MemoryLocation memory = -1;
  HYPRE_Int nrows = -1;
  const HYPRE_Int* ncols = nullptr;
  const HYPRE_BigInt *rows = nullptr, *cols = nullptr;
  const HYPRE_Complex* values = nullptr;
  bool assembled = false;
  std::vector<const char*> calls;
};
typedef hypre_IJMatrix_struct* HYPRE_IJMatrix;
typedef hypre_IJMatrix_struct* HYPRE_ParCSRMatrix;
inline HYPRE_Int HYPRE_IJMatrixCreate(MPI_Comm comm, HYPRE_BigInt ilower, HYPRE_BigInt iupper, HYPRE_BigInt jlower, HYPRE_BigInt jupper, HYPRE_IJMatrix* matrix)
{
  *matrix = new hypre_IJMatrix_struct{ comm, ilower, iupper, jlower, jupper };
  (*matrix)->calls.push_back("Create");
  return 0;
}
inline HYPRE_Int HYPRE_IJMatrixSetObjectType(HYPRE_IJMatrix m, HYPRE_Int type) { m->object_type = type; m->calls.push_back("SetObjectType"); return 0; }
inline HYPRE_Int HYPRE_IJMatrixInitialize_v2(HYPRE_IJMatrix m, HYPRE_MemoryLocation loc) { m->memory = loc; m->calls.push_back("Initialize_v2"); return 0; }
inline HYPRE_Int HYPRE_IJMatrixSetValues(HYPRE_IJMatrix m, HYPRE_Int nrows, HYPRE_Int* ncols, const HYPRE_BigInt* rows, const HYPRE_BigInt* cols, const HYPRE_Complex* values)
{
  m->nrows = nrows; m->ncols = ncols; m->rows = rows; m->cols = cols; m->values = values;
  m->calls.push_back("SetValues");
  return 0;
}
inline HYPRE_Int HYPRE_IJMatrixAssemble(HYPRE_IJMatrix m) { m->assembled = true; m->calls.push_back("Assemble"); return 0; }
inline HYPRE_Int HYPRE_IJMatrixGetObject(HYPRE_IJMatrix m, void** object) { *object = m; m->calls.push_back("GetObject"); return 0; }

typedef int PetscInt;
typedef int64_t PetscCount;
typedef double PetscScalar;
typedef int PetscErrorCode;
enum InsertMode { NOT_SET_VALUES, INSERT_VALUES, ADD_VALUES };
enum MatAssemblyType { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 };
struct _p_Mat {
  PetscCount ncoo = -1;
  const PetscInt *coo_i = nullptr, *coo_j = nullptr;
  const PetscScalar* v = nullptr;
  InsertMode mode = NOT_SET_VALUES;
  int assembled = 0;
};
typedef _p_Mat* Mat;
inline PetscErrorCode MatSetPreallocationCOOLocal(Mat A, PetscCount ncoo, PetscInt coo_i[], PetscInt coo_j[]) { A->ncoo = ncoo; A->coo_i = coo_i; A->coo_j = coo_j; return 0; }
inline PetscErrorCode MatSetValuesCOO(Mat A, const PetscScalar v[], InsertMode mode) { A->v = v; A->mode = mode; return 0; }
inline PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType) { A->assembled++; return 0; }
inline PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType) { A->assembled++; return 0; }
