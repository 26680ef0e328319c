// Test driver of the C++ facade (include/arcanefem_b200/FemUtils.h): the call sequence an ArcaneFEM module
// makes (modules/testlab/FemModule.cc:349-399, modules/elasticity/FemModule.cc:236-271), on a mesh file
// written by the pytest; results go back as a binary file and are compared with the oracle there.
//   facade_driver <mesh.bin> <mode> <out.bin>
//   modes: csr-gpu | nwcsr | coo-gpu | bsr | af-bsr | elasticity-bsr | elasticity-af-bsr-csr | solve
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "arcanefem_b200/FemUtils.h"

using namespace arcanefem_b200;

template <class T> static void put(FILE* f, const std::vector<T>& v)
{
  const long long n = (long long)v.size();
  std::fwrite(&n, sizeof(n), 1, f);
  std::fwrite(v.data(), sizeof(T), v.size(), f);
}

int main(int argc, char** argv)
{
  if (argc < 4) return 2;
  const std::string mode = argv[2];
  try {
    Context ctx(0);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[4];
    if (std::fread(hdr, sizeof(int), 4, f) != 4) return 2;
    std::vector<double> coords((size_t)hdr[2] * 3);
    std::vector<int> cells((size_t)hdr[3] * hdr[1]);
    if (std::fread(coords.data(), sizeof(double), coords.size(), f) != coords.size()) return 2;
    if (std::fread(cells.data(), sizeof(int), cells.size(), f) != cells.size()) return 2;
    std::fclose(f);
    MeshArrays mesh;
    mesh.dim = hdr[0];
    mesh.nodes_per_cell = hdr[1];
    mesh.nb_node = hdr[2];
    mesh.nb_cell = hdr[3];
    mesh.coords = coords.data();
    mesh.cell_nodes = cells.data();

    DoFLinearSystem linear_system(ctx);
    std::vector<int> rows, cols, extra;
    std::vector<double> vals;
    if (mode == "csr-gpu" || mode == "nwcsr") {
      CsrFormat csr(ctx);
      csr.initialize(mesh);
      csr.assembleBilinear(Operator::Poisson, mode == "nwcsr");
      // penalty on DoF 0 as FemModuleTestlab does through matrixSetValue
      csr.matrixSetValue(0, 0, 1.0e30);
      if (csr.getValue(0, 0) != 1.0e30) return 4;
      csr.translateToLinearSystem(linear_system);
      if (!linear_system.hasView() || linear_system.getCSRValues().nbRow() != mesh.nb_node) return 4;
      extra.push_back(csr.m_nnz);
    }
    else if (mode == "solve") {
      // assembly -> source term + flux on the faces of the first cell -> penalty on DoF 0 -> DoFLinearSystem::solve
      CsrFormat csr(ctx);
      csr.initialize(mesh);
      csr.assembleBilinear(Operator::Poisson, true);
      const double f = 1.0, g = 0.25, penalty = 1.0e30, q = 2.0;
      BoundaryConditions::applyConstantSourceToRhs(ctx, &f, 1, false);
      std::vector<int> face(cells.begin(), cells.begin() + mesh.dim); // one face: the first `dim` nodes of cell 0
      BoundaryConditions::applyNeumannToRhs(ctx, 1, face.data(), 1, &q);
      const int dof0 = 0;
      check(afb_dirichlet_penalty(ctx.handle(), 0, penalty, 1, &dof0, &g, AFB_MEM_HOST));
      csr.translateToLinearSystem(linear_system);
      std::vector<double> sol((size_t)mesh.nb_node);
      const int it = linear_system.solve(sol.data(), 1.0e-13, 0.0, 20000);
      extra.push_back(it);
      rows = ctx.copyToHost<int>(AFB_ARRAY_ROWS);
      cols = ctx.copyToHost<int>(AFB_ARRAY_COLUMNS);
      FILE* o = std::fopen(argv[3], "wb");
      if (!o) return 2;
      put(o, rows);
      put(o, cols);
      put(o, sol);
      put(o, extra);
      std::fclose(o);
      return 0;
    }
    else if (mode == "coo-gpu") {
      CooFormat coo(ctx);
      coo.initialize(mesh);
      coo.assembleBilinear(Operator::Poisson);
      DeviceSpan<const Int32> r, c;
      DeviceSpan<Real> v;
      coo.arrays(r, c, v);
      extra = ctx.copyToHost<int>(AFB_ARRAY_COO_ROWS);
    }
    else {
      const bool elast = mode.rfind("elasticity", 0) == 0;
      const bool af = mode.find("af-bsr") != std::string::npos;
      const bool use_csr = mode.find("-csr") != std::string::npos;
      BSRFormat bsr(ctx);
      bsr.initialize(mesh, (Int8)(elast ? mesh.dim : 1), use_csr, af);
      bsr.computeSparsity();
      const double E = 21.0e5, nu = 0.28;
      const double prm[2] = { E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu)) }; // modules/elasticity/FemModule.cc:171-172
      bsr.assembleBilinear(elast ? Operator::Elasticity : Operator::Poisson, elast ? prm : nullptr, elast ? 2 : 0);
      bsr.resetMatrixValues();
      bsr.assembleBilinear(elast ? Operator::Elasticity : Operator::Poisson, elast ? prm : nullptr, elast ? 2 : 0);
      BSRMatrix& m = bsr.matrix();
      extra = { m.nbRow(), m.nbColumn(), m.nbNonZero(), (int)m.nbBlock(), (int)m.orderValuePerBlock() };
      bool threw = false;
      try {
        bsr.toLinearSystem(linear_system);
      }
      catch (const FatalError&) {
        threw = true;
      }
      if (threw == use_csr) return 4; // per-block values cannot be handed to a CSR solver (femutils/BSRFormat.cc:382-383)
      if (use_csr) {
        extra.push_back(linear_system.getCSRValues().nbRow());
        extra.push_back(linear_system.getCSRValues().nbValue());
      }
    }
    rows = ctx.copyToHost<int>(AFB_ARRAY_ROWS);
    cols = ctx.copyToHost<int>(AFB_ARRAY_COLUMNS);
    vals = ctx.copyToHost<double>(AFB_ARRAY_VALUES);
    FILE* o = std::fopen(argv[3], "wb");
    if (!o) return 2;
    put(o, rows);
    put(o, cols);
    put(o, vals);
    put(o, extra);
    std::fclose(o);
  }
  catch (const FatalError& e) {
    std::fprintf(stderr, "FatalError(%d): %s\n", e.code(), e.what());
    return 3;
  }
  return 0;
}
