// A host without Arcane and without Python: reads a Gmsh file with the C ABI's reader (afb_msh_*), and -- in "solve" mode, on a GPU --
// runs the testlab Poisson sequence from the file alone (mesh, named surface -> Dirichlet nodes, matrix, source term, penalty, PCG)
// and compares the nodal solution with one of the reference's golden files (uid value per line, femutils/FemUtils.cc:108-172).
//   msh_driver info  <mesh.msh>
//   msh_driver solve <mesh.msh> <f> <penalty> <golden.txt> <surface> <value> [<surface> <value> ...]
#include <afb200.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#define CHECK(call)                                                         \
  do {                                                                      \
    if ((call) != AFB_OK) {                                                 \
      std::fprintf(stderr, "%s failed: %s\n", #call, afb_last_error());     \
      return 2;                                                             \
    }                                                                       \
  } while (0)

int main(int argc, char** argv)
{
  if (argc < 3) return 64;
  const std::string mode = argv[1];
  afb_msh* m = nullptr;
  CHECK(afb_msh_read(argv[2], &m));
  int dim = 0, npc = 0;
  int32_t nb_node = 0, nb_group = 0;
  int64_t nb_cell = 0;
  CHECK(afb_msh_sizes(m, &dim, &npc, &nb_node, &nb_cell, &nb_group));
  std::printf("mesh dim=%d npc=%d nodes=%d cells=%lld groups=%d\n", dim, npc, nb_node, (long long)nb_cell, nb_group);
  for (int32_t g = 0; g < nb_group; ++g) {
    const char* name = nullptr;
    int kind = 0, npi = 0;
    int64_t nb_item = 0, nb_gnode = 0;
    CHECK(afb_msh_group(m, g, &name, &kind, &nb_item, &npi, &nb_gnode));
    std::printf("group %s kind=%d items=%lld npi=%d nodes=%lld\n", name, kind, (long long)nb_item, npi, (long long)nb_gnode);
  }
  if (mode == "info") {
    CHECK(afb_msh_destroy(m));
    return 0;
  }
  if (argc < 8 || (argc - 6) % 2 != 0) return 64;
  const double f = std::atof(argv[3]), penalty = std::atof(argv[4]);
  std::vector<double> xyz(3 * (size_t)nb_node);
  std::vector<int32_t> cells((size_t)nb_cell * npc);
  std::vector<int64_t> uid((size_t)nb_node);
  CHECK(afb_msh_get(m, xyz.data(), cells.data(), uid.data()));
  // Dirichlet values in .arc order, later surfaces overwrite earlier ones (modules/testlab/FemModule.cc:647-677)
  std::vector<double> value((size_t)nb_node, 0.0);
  std::vector<char> fixed((size_t)nb_node, 0);
  for (int a = 6; a + 1 < argc; a += 2) {
    bool found = false;
    for (int32_t g = 0; g < nb_group && !found; ++g) {
      const char* name = nullptr;
      int kind = 0;
      int64_t nb_gnode = 0;
      CHECK(afb_msh_group(m, g, &name, &kind, nullptr, nullptr, &nb_gnode));
      if (kind == AFB_MSH_GROUP_CELLS || std::strcmp(name, argv[a]) != 0) continue;
      std::vector<int32_t> nodes((size_t)nb_gnode);
      CHECK(afb_msh_group_get(m, g, nullptr, nodes.data()));
      for (int32_t n : nodes) {
        fixed[n] = 1;
        value[n] = std::atof(argv[a + 1]);
      }
      found = true;
    }
    if (!found) {
      std::fprintf(stderr, "no surface or point named %s\n", argv[a]);
      return 3;
    }
  }
  CHECK(afb_msh_destroy(m));
  std::vector<int32_t> ids;
  std::vector<double> g_values;
  for (int32_t n = 0; n < nb_node; ++n)
    if (fixed[n]) {
      ids.push_back(n);
      g_values.push_back(value[n]);
    }

  afb_ctx* ctx = nullptr;
  CHECK(afb_create(0, &ctx));
  CHECK(afb_set_mesh(ctx, dim, npc, nb_node, nb_cell, xyz.data(), cells.data(), nullptr, AFB_MEM_HOST));
  int32_t nb_row = 0;
  int64_t nnz = 0;
  CHECK(afb_build_pattern(ctx, 1, &nb_row, &nnz));
  CHECK(afb_assemble_bilinear(ctx, AFB_OP_POISSON, nullptr, 0, AFB_FORMAT_CSR, AFB_VARIANT_TILED_GATHER, AFB_LAYOUT_PER_BLOCK, AFB_FLAG_SIGNED_TRI_AREA));
  CHECK(afb_set_dirichlet_nodes(ctx, (int32_t)ids.size(), ids.data(), AFB_MEM_HOST));
  CHECK(afb_rhs_reset(ctx));
  CHECK(afb_assemble_rhs_source(ctx, &f, 1, 0, 1));
  CHECK(afb_dirichlet_penalty(ctx, 0, penalty, (int32_t)ids.size(), ids.data(), g_values.data(), AFB_MEM_HOST));
  std::vector<double> u((size_t)nb_node);
  int iterations = 0;
  double residual = 0;
  CHECK(afb_solve_pcg(ctx, 1e-13, 0.0, 5000, u.data(), AFB_MEM_HOST, &iterations, &residual));
  CHECK(afb_destroy(ctx));

  std::unordered_map<long long, double> golden;
  if (FILE* fp = std::fopen(argv[5], "r")) {
    long long id;
    double v;
    while (std::fscanf(fp, "%lld %lf", &id, &v) == 2) golden[id] = v;
    std::fclose(fp);
  }
  if (golden.size() != (size_t)nb_node) {
    std::fprintf(stderr, "golden file lists %zu nodes, the mesh has %d\n", golden.size(), nb_node);
    return 4;
  }
  double worst = 0;
  for (int32_t n = 0; n < nb_node; ++n) {
    auto it = golden.find((long long)uid[n]);
    if (it == golden.end()) return 5;
    const double r = it->second, v = u[n];
    if (std::fabs(r) < 1e-16 && std::fabs(v) < 1e-16) continue;
    worst = std::fmax(worst, std::fabs(r - v) / std::fmax(std::fabs(r), std::fabs(v)));
  }
  std::printf("rows=%d nnz=%lld pcg_iterations=%d worst=%.3e\n", nb_row, (long long)nnz, iterations, worst);
  if (!(worst < 1e-6)) return 6;
  std::printf("golden ok\n");
  return 0;
}
