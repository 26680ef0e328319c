// Test driver of the domain-decomposition layer of the C ABI (include/afb200.h: afb_partition_*, afb_xplan_*), no Python in
// the loop: one host thread per rank (what one MPI rank is to the reference), every rank with its own context on the box's
// GPU(s), an in-process transport (allgather / neighbour exchange through shared memory) standing in for MPI.
//   mgpu_driver <mesh.bin> <world> <b> <layout> <allow_peer_memory> <out.bin>
// Writes per rank: rows, columns, values of its sub-domain after assembly + exchange (3 repetitions of BuildMatrix + assembly +
// exchange: steady state), node_gid, nb_own_node, and the global numbering; the pytest compares with the global oracle matrix.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "afb200.h"

struct Shared { // in-process "MPI": a barrier and mailboxes
  int world = 0;
  std::mutex m;
  std::condition_variable cv;
  int count = 0, gen = 0;
  std::vector<const void*> ag_send;
  std::vector<std::vector<const void*>> box; // box[src][dst]: buffer src sends to dst in the current exchange
  void barrier()
  {
    std::unique_lock<std::mutex> lk(m);
    const int g = gen;
    if (++count == world) {
      count = 0;
      ++gen;
      cv.notify_all();
    }
    else if (!cv.wait_for(lk, std::chrono::seconds(120), [&] { return gen != g; })) {
      std::fprintf(stderr, "mgpu_driver: a rank did not reach the barrier (another rank failed)\n");
      std::_Exit(6);
    }
  }
};

struct Rank {
  Shared* sh;
  int rank;
};

static int t_allgather(void* user, const void* send, int64_t bytes, void* recv)
{
  Rank* r = static_cast<Rank*>(user);
  Shared& S = *r->sh;
  S.ag_send[r->rank] = send;
  S.barrier();
  for (int q = 0; q < S.world; ++q) std::memcpy(static_cast<char*>(recv) + (size_t)q * bytes, S.ag_send[q], (size_t)bytes);
  S.barrier();
  return 0;
}

static int t_exchange(void* user, int32_t nb_peer, const int32_t* peer, const void* const* send, const int64_t* send_bytes, void* const* recv, const int64_t* recv_bytes,
                      int device_memory)
{
  if (device_memory) return 1; // this transport moves host memory only (exchange_takes_device_memory = 0)
  Rank* r = static_cast<Rank*>(user);
  Shared& S = *r->sh;
  for (int k = 0; k < nb_peer; ++k) S.box[r->rank][peer[k]] = send[k];
  S.barrier();
  for (int k = 0; k < nb_peer; ++k)
    if (recv_bytes[k]) std::memcpy(recv[k], S.box[peer[k]][r->rank], (size_t)recv_bytes[k]);
  (void)send_bytes;
  S.barrier();
  return 0;
}

template <class T> static void put(FILE* f, const std::vector<T>& v)
{
  const long long n = (long long)v.size();
  std::fwrite(&n, sizeof(n), 1, f);
  std::fwrite(v.data(), sizeof(T), v.size(), f);
}

#define CHECK(call)                                                                         \
  do {                                                                                      \
    const int rc_ = (call);                                                                 \
    if (rc_ != 0) {                                                                         \
      std::fprintf(stderr, "rank %d: %s -> %d: %s\n", rank, #call, rc_, afb_last_error()); \
      failed = true;                                                                        \
      return;                                                                               \
    }                                                                                       \
  } while (0)

struct Result {
  std::vector<int32_t> rows, cols, l2g;
  std::vector<double> vals;
  std::vector<int64_t> gid, first;
  int32_t nb_own = 0, kind = 0;
};

int main(int argc, char** argv)
{
  if (argc < 7) return 2;
  const int world = std::atoi(argv[2]), b = std::atoi(argv[3]), layout = std::atoi(argv[4]), allow_p2p = std::atoi(argv[5]);
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int hdr[4];
  if (std::fread(hdr, sizeof(int), 4, f) != 4) return 2;
  const int dim = hdr[0], npc = hdr[1], nb_node = hdr[2], nb_cell = hdr[3];
  std::vector<double> coords((size_t)nb_node * 3);
  std::vector<int32_t> cells((size_t)nb_cell * npc);
  if (std::fread(coords.data(), sizeof(double), coords.size(), f) != coords.size()) return 2;
  if (std::fread(cells.data(), sizeof(int32_t), cells.size(), f) != cells.size()) return 2;
  std::fclose(f);

  afb_partition* part = nullptr;
  if (afb_partition_create(dim, npc, nb_node, nb_cell, coords.data(), cells.data(), world, &part) != 0) {
    std::fprintf(stderr, "partition: %s\n", afb_last_error());
    return 4;
  }
  Shared sh;
  sh.world = world;
  sh.ag_send.resize(world);
  sh.box.assign(world, std::vector<const void*>(world, nullptr));
  std::vector<Result> res(world);
  bool failed = false;
  auto body = [&](int rank) {
    Rank me{ &sh, rank };
    afb_transport t;
    std::memset(&t, 0, sizeof(t));
    t.user = &me;
    t.rank = rank;
    t.world = world;
    t.allgather = t_allgather;
    t.exchange = t_exchange;
    int32_t nn = 0, no = 0;
    int64_t nc = 0, noc = 0;
    CHECK(afb_partition_sizes(part, rank, &nn, &no, &nc, &noc));
    std::vector<double> xyz((size_t)nn * 3);
    std::vector<int32_t> cn((size_t)nc * npc), owner(nn);
    std::vector<uint8_t> own(nn);
    Result& R = res[rank];
    R.gid.resize(nn);
    R.nb_own = no;
    CHECK(afb_partition_get(part, rank, xyz.data(), cn.data(), own.data(), R.gid.data(), owner.data(), nullptr));
    afb_ctx* ctx = nullptr;
    int rc = afb_create(0, &ctx);
    if (rc != 0) {
      std::fprintf(stderr, "rank %d: afb_create: %s\n", rank, afb_last_error());
      failed = true;
      // keep the collective calls of the other ranks from dead-locking: nothing collective has happened yet, all ranks fail alike
      return;
    }
    CHECK(afb_set_mesh(ctx, dim, npc, nn, nc, xyz.data(), cn.data(), own.data(), AFB_MEM_HOST));
    CHECK(afb_set_own_cell_count(ctx, noc));
    const double E = 21.0e5, nu = 0.28;
    const double prm[2] = { E * nu / ((1 + nu) * (1 - 2 * nu)), E / (2 * (1 + nu)) };
    afb_xplan* x = nullptr;
    for (int rep = 0; rep < 3; ++rep) {
      int32_t nbr = 0;
      int64_t nnz = 0;
      CHECK(afb_build_pattern(ctx, b, &nbr, &nnz));
      if (x) CHECK(afb_xplan_wait(x));
      CHECK(afb_assemble_bilinear(ctx, b == 1 ? AFB_OP_POISSON : AFB_OP_ELASTICITY, prm, 2, b == 1 ? AFB_FORMAT_CSR : AFB_FORMAT_BSR, AFB_VARIANT_TILED_GATHER, layout,
                                  AFB_FLAG_OWN_CELLS_ONLY | AFB_FLAG_ALL_ROWS));
      if (!x) CHECK(afb_xplan_create(ctx, &t, R.gid.data(), owner.data(), no, allow_p2p, &x));
      // Ranks of ONE process share the device's allocator: a cudaFree / cudaMalloc of one rank (grow-only scratch buffers of
      // the first builds) waits for every kernel on the device, including a neighbour's exchange kernel that is itself
      // waiting for this rank.  With one process per GPU (the deployment) that coupling does not exist; here the ranks meet
      // before the exchange and after it.
      sh.barrier();
      CHECK(afb_xplan_exchange(x));
      CHECK(afb_xplan_wait(x));
      CHECK(afb_synchronize(ctx));
      sh.barrier();
    }
    int32_t np = 0;
    CHECK(afb_xplan_info(x, &np, nullptr, nullptr, &R.kind, nullptr));
    size_t bytes = 0;
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_ROWS, nullptr, &bytes));
    R.rows.resize(bytes / 4);
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_ROWS, R.rows.data(), &bytes));
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_COLUMNS, nullptr, &bytes));
    R.cols.resize(bytes / 4);
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_COLUMNS, R.cols.data(), &bytes));
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_VALUES, nullptr, &bytes));
    R.vals.resize(bytes / 8);
    CHECK(afb_copy_to_host(ctx, AFB_ARRAY_VALUES, R.vals.data(), &bytes));
    R.first.resize(world + 1);
    R.l2g.resize((size_t)nn * b);
    CHECK(afb_xplan_numbering(x, R.first.data(), R.l2g.data()));
    // odd ranks destroy the context first: the plan must survive that (afb_destroy detaches it) and still be destroyable
    sh.barrier();
    if (rank & 1) CHECK(afb_destroy(ctx));
    else CHECK(afb_xplan_destroy(x));
    sh.barrier();
    if (rank & 1) {
      if (afb_xplan_exchange(x) == 0) {
        std::fprintf(stderr, "rank %d: exchange on a plan whose context is gone did not fail\n", rank);
        failed = true;
      }
      CHECK(afb_xplan_destroy(x));
    }
    else CHECK(afb_destroy(ctx));
  };
  {
    // no GPU: afb_create fails on every rank before anything collective -- report like the facade driver does
    afb_ctx* probe = nullptr;
    if (afb_create(0, &probe) != 0) {
      std::fprintf(stderr, "%s\n", afb_last_error());
      afb_partition_destroy(part);
      return 3;
    }
    afb_destroy(probe);
  }
  std::vector<std::thread> th;
  for (int r = 0; r < world; ++r) th.emplace_back(body, r);
  for (auto& t : th) t.join();
  afb_partition_destroy(part);
  if (failed) return 5;
  FILE* o = std::fopen(argv[6], "wb");
  if (!o) return 2;
  for (int r = 0; r < world; ++r) {
    put(o, res[r].rows);
    put(o, res[r].cols);
    put(o, res[r].vals);
    put(o, res[r].gid);
    put(o, std::vector<int32_t>{ res[r].nb_own, res[r].kind });
    put(o, res[r].first);
    put(o, res[r].l2g);
  }
  std::fclose(o);
  return 0;
}
