// A user kernel over the device view, written against the reference's CsrFormatMatrixView interface
// (femutils/CsrFormatMatrixView.h:135-215: rowRange / column / value / tryFindColumnInRow) and its value types
// (femutils/FemUtils.h RealVector / RealMatrix): what an ArcaneFEM module's own RUNCOMMAND loop over the matrix looks
// like once the containers are the B200 ones.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -Iinclude tests/cpp/view_kernel.cu -Larcanefem_b200 -lafb200
//   view_kernel <n>     box of n^3 cubes; prints "view ok" when the checks hold
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "arcanefem_b200/FemUtils.h"

using namespace arcanefem_b200;

// per row: sum of the row, the diagonal found by search, and the number of entries walked
__global__ void k_row_checks(CsrFormatMatrixView view, Real* row_sum, Real* diagonal, Int32* walked)
{
  const Int32 r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= view.nbRow()) return;
  RealVector<2> acc; // [sum, sum of |a|]
  Int32 n = 0;
  for (CsrRowColumnIndex rc : view.rowRange(r)) {
    acc(0) += view.value(rc);
    acc(1) += fabs(view.value(rc));
    ++n;
  }
  const CsrRowColumnIndex d = view.tryFindColumnInRow(r, r);
  row_sum[r] = acc(0) / (acc(1) > 0.0 ? acc(1) : 1.0);
  diagonal[r] = d.isNull() ? -1.0 : view.value(d);
  walked[r] = n;
}

int main(int argc, char** argv)
{
  const int n = argc > 1 ? std::atoi(argv[1]) : 6;
  try {
    Context ctx(0);
    check(afb_mesh_generate_box(ctx.handle(), 3, n, 0.2, 1234u, 0, n, 0)); // the whole box, generated in HBM
    CsrFormat csr(ctx);
    csr.computeSparsity();
    csr.assembleBilinear(Operator::Poisson, true);
    DoFLinearSystem ls(ctx);
    csr.translateToLinearSystem(ls);
    const CsrFormatMatrixView view = ls.getCSRValues();
    const Int32 nr = view.nbRow();
    Real *d_sum = nullptr, *d_diag = nullptr;
    Int32* d_walked = nullptr;
    if (cudaMalloc(&d_sum, sizeof(Real) * nr) != cudaSuccess || cudaMalloc(&d_diag, sizeof(Real) * nr) != cudaSuccess || cudaMalloc(&d_walked, sizeof(Int32) * nr) != cudaSuccess) return 3;
    ctx.barrier();
    k_row_checks<<<(nr + 127) / 128, 128>>>(view, d_sum, d_diag, d_walked);
    if (cudaDeviceSynchronize() != cudaSuccess) return 3;
    std::vector<Real> sum(nr), diag(nr);
    std::vector<Int32> walked(nr);
    cudaMemcpy(sum.data(), d_sum, sizeof(Real) * nr, cudaMemcpyDeviceToHost);
    cudaMemcpy(diag.data(), d_diag, sizeof(Real) * nr, cudaMemcpyDeviceToHost);
    cudaMemcpy(walked.data(), d_walked, sizeof(Int32) * nr, cudaMemcpyDeviceToHost);
    const std::vector<int> nbcol = ctx.copyToHost<int>(AFB_ARRAY_NZ_PER_ROW);
    long long total = 0;
    for (Int32 r = 0; r < nr; ++r) {
      if (std::fabs(sum[r]) > 1e-12) { std::printf("row %d: relative row sum %g\n", r, sum[r]); return 4; } // stiffness rows sum to zero
      if (!(diag[r] > 0.0)) { std::printf("row %d: diagonal %g\n", r, diag[r]); return 4; }
      if (walked[r] != nbcol[r]) { std::printf("row %d: walked %d entries of %d\n", r, walked[r], nbcol[r]); return 4; }
      total += walked[r];
    }
    if (total != view.nbValue()) return 4;
    // host-side single-entry access of DoFLinearSystem agrees with what the kernel saw
    if (ls.matrixGetValue(0, 0) != diag[0]) return 5;
    ls.matrixAddValue(0, 0, 1.0);
    if (ls.matrixGetValue(0, 0) != diag[0] + 1.0) return 5;
    ls.matrixSetValue(0, 0, diag[0]);
    if (ls.matrixGetValue(0, 0) != diag[0]) return 5;
    std::printf("view ok: %d rows, %lld entries\n", nr, total);
    cudaFree(d_sum);
    cudaFree(d_diag);
    cudaFree(d_walked);
  }
  catch (const std::exception& e) {
    std::printf("error: %s\n", e.what());
    return 1;
  }
  return 0;
}
