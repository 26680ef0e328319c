// Test driver of include/arcanefem_b200/SolverHandoff.h against recording stand-ins of HYPRE / PETSc (mock_solvers.h):
// assembles the mesh file, hands the matrix over the way HypreDoFLinearSystemImpl::solve and PetscDoFLinearSystemImpl do,
// and writes what the "solver" received (copied back from the device pointers it was given) for the pytest to compare.
//   handoff_driver <mesh.bin> <first_own_row> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "mock_solvers.h"
#define AFB_HAVE_HYPRE 1
#define AFB_HAVE_PETSC 1
#include "arcanefem_b200/SolverHandoff.h"

template <class T> static void put(FILE* f, const std::vector<T>& v)
{
  const long long n = (long long)v.size();
  std::fwrite(&n, sizeof(n), 1, f);
  std::fwrite(v.data(), sizeof(T), v.size(), f);
}
template <class T> static std::vector<T> fetch(afb_ctx* ctx, const T* dev, size_t n)
{
  std::vector<T> h(n);
  arcanefem_b200::handoffCheck(afb_memcpy_to_host(ctx, h.data(), dev, sizeof(T) * n), "afb_memcpy_to_host");
  return h;
}

int main(int argc, char** argv)
{
  if (argc < 4) return 2;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4];
  if (std::fread(hdr, sizeof(int), 4, f) != 4) return 2;
  std::vector<double> coords((size_t)hdr[2] * 3);
  std::vector<int32_t> cells((size_t)hdr[3] * hdr[1]);
  if (std::fread(coords.data(), sizeof(double), coords.size(), f) != coords.size()) return 2;
  if (std::fread(cells.data(), sizeof(int32_t), cells.size(), f) != cells.size()) return 2;
  std::fclose(f);
  const int first_row = std::atoi(argv[2]);
  afb_ctx* ctx = nullptr;
  if (afb_create(0, &ctx) != 0) {
    std::fprintf(stderr, "%s\n", afb_last_error());
    return 3;
  }
  try {
    using arcanefem_b200::handoffCheck;
    handoffCheck(afb_set_mesh(ctx, hdr[0], hdr[1], hdr[2], hdr[3], coords.data(), cells.data(), nullptr, AFB_MEM_HOST), "afb_set_mesh");
    int32_t nbr = 0;
    int64_t nnz = 0;
    handoffCheck(afb_build_pattern(ctx, 1, &nbr, &nnz), "afb_build_pattern");
    handoffCheck(afb_assemble_bilinear(ctx, AFB_OP_POISSON, nullptr, 0, AFB_FORMAT_CSR, AFB_VARIANT_TILED_GATHER, AFB_LAYOUT_PER_BLOCK, 0), "afb_assemble_bilinear");
    // sequential hand-off (columns unchanged); the parallel column renumbering is covered by tests/test_distributed.py
    HYPRE_ParCSRMatrix par = nullptr;
    HYPRE_IJMatrix A = arcanefem_b200::hypreSetCSRValues(ctx, MPI_COMM_WORLD, first_row, nbr, nullptr, &par);
    FILE* o = std::fopen(argv[3], "wb");
    if (!o) return 2;
    put(o, std::vector<int32_t>{ A->ilower, A->iupper, A->jlower, A->jupper, A->object_type, A->memory, A->nrows, A->assembled ? 1 : 0, (int32_t)A->calls.size(), par == A ? 1 : 0 });
    std::vector<int32_t> ncols = fetch(ctx, A->ncols, (size_t)nbr);
    put(o, ncols);
    put(o, fetch(ctx, A->rows, (size_t)nbr));
    put(o, fetch(ctx, A->cols, (size_t)nnz));
    put(o, fetch(ctx, A->values, (size_t)nnz));
    _p_Mat mat;
    arcanefem_b200::petscSetCOOValues(ctx, &mat, true);
    put(o, std::vector<int32_t>{ (int32_t)mat.ncoo, (int32_t)mat.mode, mat.assembled });
    put(o, fetch(ctx, mat.coo_i, (size_t)nnz));
    put(o, fetch(ctx, mat.coo_j, (size_t)nnz));
    put(o, fetch(ctx, mat.v, (size_t)nnz));
    std::fclose(o);
    delete A;
  }
  catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    afb_destroy(ctx);
    return 4;
  }
  afb_destroy(ctx);
  return 0;
}
