// Domain decomposition behind the C ABI (afb200.h: afb_partition_*, afb_xplan_host_*, afb_xplan_*).
//
// Reference behaviour replaced:
//   * Arcane's mesh partitioner + ghost layer (every rank holds its own cells and one layer of ghost cells, every node has
//     one owner; ArcaneFEM only sees the result: isOwn gates, modules/testlab/CsrGpuBiliAssembly.cc:273,351);
//   * HypreDoFLinearSystemImpl::_computeMatrixNumeration (femutils/HypreDoFLinearSystem.cc:209-249): allGather of the owned
//     row counts, exclusive scan, ghost rows learn their global number from the owner (variable synchronize);
//   * the off-processor rows of the solver's parallel matrix (femutils/HypreDoFLinearSystem.cc:461-520): here every rank
//     assembles its OWN cells into the rows of ALL its local nodes and the partial ghost rows travel to their owners.
// Communication of the set-up goes through two host callbacks (afb_transport); the rows themselves move in one kernel over
// NVLink peer memory (p2p.cu) or, where peer memory cannot be mapped, through the transport.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "afb_internal.h"

using namespace afb;

// ---------------------------------------------------------------------------------------------
// partition (host)
// ---------------------------------------------------------------------------------------------
struct afb_partition {
  int dim = 0, npc = 0, world = 0;
  struct Sub {
    std::vector<double> xyz;
    std::vector<int32_t> cells;
    std::vector<uint8_t> is_own;
    std::vector<int64_t> node_gid, cell_gid;
    std::vector<int32_t> node_owner;
    int32_t nb_own_node = 0;
    int64_t nb_own_cell = 0;
  };
  std::vector<Sub> sub;
};

// recursive coordinate bisection: cells [begin, end) of `order` go to ranks [r0, r1)
static void rcb(std::vector<int64_t>& order, int64_t begin, int64_t end, int r0, int r1, const std::vector<double>& cent, std::vector<int32_t>& cell_rank)
{
  if (r1 - r0 <= 1) {
    for (int64_t i = begin; i < end; ++i) cell_rank[order[i]] = r0;
    return;
  }
  double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
  for (int64_t i = begin; i < end; ++i)
    for (int a = 0; a < 3; ++a) {
      const double v = cent[3 * order[i] + a];
      lo[a] = std::min(lo[a], v);
      hi[a] = std::max(hi[a], v);
    }
  // longest axis; among (nearly) equal extents the LAST one: node ids of generated and most imported meshes vary slowest along it
  double longest = 0.0;
  for (int a = 0; a < 3; ++a) longest = std::max(longest, hi[a] - lo[a]);
  int axis = 0;
  for (int a = 0; a < 3; ++a)
    if (hi[a] - lo[a] >= longest * (1.0 - 1e-9)) axis = a;
  const int rm = r0 + (r1 - r0) / 2;
  const int64_t mid = begin + (end - begin) * (rm - r0) / (r1 - r0);
  std::nth_element(order.begin() + begin, order.begin() + mid, order.begin() + end, [&](int64_t a, int64_t b) {
    const double va = cent[3 * a + axis], vb = cent[3 * b + axis];
    return va < vb || (va == vb && a < b);
  });
  rcb(order, begin, mid, r0, rm, cent, cell_rank);
  rcb(order, mid, end, rm, r1, cent, cell_rank);
}

extern "C" {

int afb_partition_create(int dim, int npc, int32_t nb_node, int64_t nb_cell, const double* xyz, const int32_t* cell_nodes, int world, afb_partition** out)
{
  AFB_REQUIRE(out && xyz && cell_nodes && world >= 1 && nb_node >= 0 && nb_cell >= 0 && npc >= 2 && (dim == 2 || dim == 3), AFB_ERR_INVALID, "afb_partition_create: bad arguments");
  *out = nullptr;
  afb_partition* P = new afb_partition();
  P->dim = dim;
  P->npc = npc;
  P->world = world;
  std::vector<double> cent(3 * (size_t)nb_cell, 0.0);
  for (int64_t c = 0; c < nb_cell; ++c)
    for (int a = 0; a < npc; ++a) {
      const int32_t n = cell_nodes[c * npc + a];
      if (n < 0 || n >= nb_node) {
        delete P;
        AFB_REQUIRE(false, AFB_ERR_INVALID, "afb_partition_create: cell %lld references node %d outside [0,%d)", (long long)c, n, nb_node);
      }
      for (int k = 0; k < dim; ++k) cent[3 * c + k] += xyz[3 * (size_t)n + k] / npc;
    }
  std::vector<int64_t> order((size_t)nb_cell);
  std::iota(order.begin(), order.end(), 0);
  std::vector<int32_t> cell_rank((size_t)nb_cell, 0);
  rcb(order, 0, nb_cell, 0, world, cent, cell_rank);
  // owner of a node: the lowest rank among its cells (isolated nodes: rank 0)
  std::vector<int32_t> node_owner((size_t)nb_node, world);
  for (int64_t c = 0; c < nb_cell; ++c)
    for (int a = 0; a < npc; ++a) {
      int32_t& o = node_owner[cell_nodes[c * npc + a]];
      o = std::min(o, cell_rank[c]);
    }
  for (auto& o : node_owner)
    if (o == world) o = 0;
  P->sub.resize(world);
  std::vector<int32_t> g2l((size_t)nb_node);
  for (int r = 0; r < world; ++r) {
    afb_partition::Sub& S = P->sub[r];
    std::vector<int64_t> own_cells, ghost_cells;
    for (int64_t c = 0; c < nb_cell; ++c) {
      if (cell_rank[c] == r) own_cells.push_back(c);
      else {
        bool touches = false;
        for (int a = 0; a < npc; ++a) touches |= node_owner[cell_nodes[c * npc + a]] == r;
        if (touches) ghost_cells.push_back(c);
      }
    }
    S.nb_own_cell = (int64_t)own_cells.size();
    S.cell_gid = own_cells;
    S.cell_gid.insert(S.cell_gid.end(), ghost_cells.begin(), ghost_cells.end());
    std::vector<uint8_t> used((size_t)nb_node, 0);
    for (int64_t c : S.cell_gid)
      for (int a = 0; a < npc; ++a) used[cell_nodes[c * npc + a]] = 1;
    std::vector<int64_t> owned, ghosts;
    for (int32_t n = 0; n < nb_node; ++n) {
      if (node_owner[n] == r) owned.push_back(n); // (owned nodes without a cell here cannot exist: the owner holds a cell)
      else if (used[n]) ghosts.push_back(n);
    }
    std::stable_sort(ghosts.begin(), ghosts.end(), [&](int64_t a, int64_t b) { return node_owner[a] != node_owner[b] ? node_owner[a] < node_owner[b] : a < b; });
    S.nb_own_node = (int32_t)owned.size();
    S.node_gid = owned;
    S.node_gid.insert(S.node_gid.end(), ghosts.begin(), ghosts.end());
    const size_t nl = S.node_gid.size();
    S.xyz.resize(3 * nl);
    S.is_own.resize(nl);
    S.node_owner.resize(nl);
    for (size_t i = 0; i < nl; ++i) {
      const int64_t g = S.node_gid[i];
      g2l[g] = (int32_t)i;
      for (int k = 0; k < 3; ++k) S.xyz[3 * i + k] = xyz[3 * (size_t)g + k];
      S.node_owner[i] = node_owner[g];
      S.is_own[i] = node_owner[g] == r ? 1 : 0;
    }
    S.cells.resize(S.cell_gid.size() * (size_t)npc);
    for (size_t c = 0; c < S.cell_gid.size(); ++c)
      for (int a = 0; a < npc; ++a) S.cells[c * npc + a] = g2l[cell_nodes[S.cell_gid[c] * npc + a]];
  }
  *out = P;
  return AFB_OK;
}

int afb_partition_destroy(afb_partition* p)
{
  delete p;
  return AFB_OK;
}

int afb_partition_sizes(const afb_partition* p, int rank, int32_t* nb_node, int32_t* nb_own_node, int64_t* nb_cell, int64_t* nb_own_cell)
{
  AFB_REQUIRE(p && rank >= 0 && rank < p->world, AFB_ERR_INVALID, "afb_partition_sizes: bad rank");
  const auto& S = p->sub[rank];
  if (nb_node) *nb_node = (int32_t)S.node_gid.size();
  if (nb_own_node) *nb_own_node = S.nb_own_node;
  if (nb_cell) *nb_cell = (int64_t)S.cell_gid.size();
  if (nb_own_cell) *nb_own_cell = S.nb_own_cell;
  return AFB_OK;
}

int afb_partition_get(const afb_partition* p, int rank, double* xyz, int32_t* cell_nodes, uint8_t* is_own, int64_t* node_gid, int32_t* node_owner, int64_t* cell_gid)
{
  AFB_REQUIRE(p && rank >= 0 && rank < p->world, AFB_ERR_INVALID, "afb_partition_get: bad rank");
  const auto& S = p->sub[rank];
  if (xyz) memcpy(xyz, S.xyz.data(), sizeof(double) * S.xyz.size());
  if (cell_nodes) memcpy(cell_nodes, S.cells.data(), sizeof(int32_t) * S.cells.size());
  if (is_own) memcpy(is_own, S.is_own.data(), S.is_own.size());
  if (node_gid) memcpy(node_gid, S.node_gid.data(), sizeof(int64_t) * S.node_gid.size());
  if (node_owner) memcpy(node_owner, S.node_owner.data(), sizeof(int32_t) * S.node_owner.size());
  if (cell_gid) memcpy(cell_gid, S.cell_gid.data(), sizeof(int64_t) * S.cell_gid.size());
  return AFB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// exchange plan: host index logic
// ---------------------------------------------------------------------------------------------
struct afb_xplan_host {
  afb_transport t;
  int b = 1, layout = 0;
  int32_t nb_node = 0, nb_own = 0;
  std::vector<int64_t> node_gid;
  std::vector<int32_t> node_owner;
  std::vector<int64_t> gid_sorted;   // for gid -> local id
  std::vector<int32_t> gid_order;
  std::vector<int32_t> peer;                       // ascending ranks
  std::vector<int64_t> send_first, send_count, recv_count;
  std::vector<std::vector<int32_t>> dof_rows, dof_cols; // per peer
  struct Range { int32_t owner, g0, g1; };
  std::vector<Range> ghost_ranges;                 // ghost nodes [g0, g1) owned by `owner`
  int32_t local_of(int64_t gid) const
  {
    auto it = std::lower_bound(gid_sorted.begin(), gid_sorted.end(), gid);
    if (it == gid_sorted.end() || *it != gid) return -1;
    return gid_order[it - gid_sorted.begin()];
  }
};

// variable-size exchange with every rank that has something for us or we for it; returns per-rank received arrays
static int exchange_i64(const afb_transport& t, const std::vector<std::vector<int64_t>>& out, std::vector<std::vector<int64_t>>& in)
{
  const int W = t.world;
  std::vector<int64_t> counts((size_t)W, 0), all((size_t)W * W, 0);
  for (int q = 0; q < W; ++q) counts[q] = (int64_t)out[q].size();
  AFB_REQUIRE(t.allgather(t.user, counts.data(), (int64_t)sizeof(int64_t) * W, all.data()) == 0, AFB_ERR_INVALID, "transport: allgather failed");
  in.assign(W, {});
  std::vector<int32_t> peers;
  std::vector<const void*> sp;
  std::vector<void*> rp;
  std::vector<int64_t> sb, rb;
  for (int q = 0; q < W; ++q) {
    if (q == t.rank) continue;
    const int64_t ns = counts[q], nr = all[(size_t)q * W + t.rank];
    if (ns == 0 && nr == 0) continue;
    in[q].resize((size_t)nr);
    peers.push_back(q);
    sp.push_back(out[q].data());
    sb.push_back(ns * (int64_t)sizeof(int64_t));
    rp.push_back(in[q].data());
    rb.push_back(nr * (int64_t)sizeof(int64_t));
  }
  // every rank calls the exchange, also with no peer: the transport may be collective
  AFB_REQUIRE(t.exchange(t.user, (int32_t)peers.size(), peers.data(), sp.data(), sb.data(), rp.data(), rb.data(), 0) == 0, AFB_ERR_INVALID, "transport: exchange failed");
  return AFB_OK;
}

extern "C" {

int afb_xplan_host_create(const afb_transport* t, int b, int layout, int32_t nb_node, int32_t nb_own, const int64_t* node_gid, const int32_t* node_owner,
                          const int32_t* rows_tail, const int32_t* cols_tail, afb_xplan_host** out)
{
  AFB_REQUIRE(out && t && t->allgather && t->exchange && t->world >= 1 && t->rank >= 0 && t->rank < t->world, AFB_ERR_INVALID, "afb_xplan_host_create: bad transport");
  AFB_REQUIRE(b >= 1 && b <= 3 && nb_node >= 0 && nb_own >= 0 && nb_own <= nb_node && node_gid && node_owner && rows_tail, AFB_ERR_INVALID, "afb_xplan_host_create: bad arguments");
  *out = nullptr;
  auto* H = new afb_xplan_host();
  H->t = *t;
  H->b = b;
  H->layout = layout;
  H->nb_node = nb_node;
  H->nb_own = nb_own;
  H->node_gid.assign(node_gid, node_gid + nb_node);
  H->node_owner.assign(node_owner, node_owner + nb_node);
  H->gid_order.resize(nb_node);
  std::iota(H->gid_order.begin(), H->gid_order.end(), 0);
  std::stable_sort(H->gid_order.begin(), H->gid_order.end(), [&](int32_t a, int32_t c) { return node_gid[a] < node_gid[c]; });
  H->gid_sorted.resize(nb_node);
  for (int32_t i = 0; i < nb_node; ++i) H->gid_sorted[i] = node_gid[H->gid_order[i]];
  const int W = t->world, me = t->rank;
  auto fail = [&](const char* msg) {
    delete H;
    set_error("afb_xplan_host_create: %s", msg);
    return AFB_ERR_INVALID;
  };
  // ghost nodes are numbered last, grouped by ascending owner: contiguous ranges per owner
  for (int32_t g = nb_own; g < nb_node;) {
    const int32_t q = node_owner[g];
    if (q == me || q < 0 || q >= W) return fail("ghost nodes must follow the owned ones and carry a valid owner");
    int32_t e = g;
    while (e < nb_node && node_owner[e] == q) ++e;
    if (!H->ghost_ranges.empty() && H->ghost_ranges.back().owner >= q) return fail("ghost nodes must be grouped by ascending owner rank");
    H->ghost_ranges.push_back({ q, g, e });
    g = e;
  }
  for (int32_t g = 0; g < nb_own; ++g)
    if (node_owner[g] != me) return fail("the first nb_own_node nodes must be owned by this rank");
  const int64_t bb = (int64_t)b * b;
  // what this rank sends: per owner the (row gid, col gid) of every block entry of its ghost rows, in memory order
  std::vector<std::vector<int64_t>> outv((size_t)W), inv;
  std::vector<int64_t> my_first((size_t)W, 0), my_count((size_t)W, 0);
  const int64_t base = rows_tail[0];
  for (const auto& R : H->ghost_ranges) {
    const int64_t r0 = rows_tail[R.g0 - nb_own], r1 = rows_tail[R.g1 - nb_own];
    auto& v = outv[R.owner];
    v.reserve((size_t)(2 * (r1 - r0)));
    for (int32_t g = R.g0; g < R.g1; ++g)
      for (int64_t p = rows_tail[g - nb_own]; p < rows_tail[g - nb_own + 1]; ++p) {
        const int32_t c = cols_tail[p - base];
        if (c < 0 || c >= nb_node) return fail("column id outside the local nodes");
        v.push_back(node_gid[g]);
        v.push_back(node_gid[c]);
      }
    my_first[R.owner] = r0 * bb;
    my_count[R.owner] = (r1 - r0) * bb;
  }
  int rc = exchange_i64(*t, outv, inv);
  if (rc != AFB_OK) {
    delete H;
    return rc;
  }
  for (int q = 0; q < W; ++q) {
    if (q == me) continue;
    const bool sends = my_count[q] > 0, recvs = !inv[q].empty();
    if (!sends && !recvs) continue;
    H->peer.push_back(q);
    H->send_first.push_back(my_first[q]);
    H->send_count.push_back(my_count[q]);
    const size_t n = inv[q].size() / 2;
    std::vector<int32_t> lr(n), lc(n);
    for (size_t i = 0; i < n; ++i) {
      lr[i] = H->local_of(inv[q][2 * i]);
      lc[i] = H->local_of(inv[q][2 * i + 1]);
      if (lr[i] < 0 || lr[i] >= nb_own) return fail("received a partial row of a node this rank does not own");
      if (lc[i] < 0) return fail("a neighbour's partial row references a node unknown here (ghost layer missing)");
    }
    // scalar (dof_row, dof_col) pairs in the SENDER's memory order (BSRMatrix::findValueIndex layouts, femutils/BSRFormat.cc:79-106)
    std::vector<int32_t> dr, dc;
    dr.reserve(n * bb);
    dc.reserve(n * bb);
    if (b == 1) {
      dr = lr;
      dc = lc;
    }
    else if (layout == AFB_LAYOUT_PER_BLOCK) {
      for (size_t i = 0; i < n; ++i)
        for (int ii = 0; ii < b; ++ii)
          for (int jj = 0; jj < b; ++jj) {
            dr.push_back(lr[i] * b + ii);
            dc.push_back(lc[i] * b + jj);
          }
    }
    else { // per row: index = rb*b*b + b*(x + i*nz) + j  ->  order i, x, j inside a block row
      for (size_t s = 0; s < n;) {
        size_t e = s;
        while (e < n && inv[q][2 * e] == inv[q][2 * s]) ++e;
        for (int ii = 0; ii < b; ++ii)
          for (size_t x = s; x < e; ++x)
            for (int jj = 0; jj < b; ++jj) {
              dr.push_back(lr[x] * b + ii);
              dc.push_back(lc[x] * b + jj);
            }
        s = e;
      }
    }
    H->recv_count.push_back((int64_t)dr.size());
    H->dof_rows.push_back(std::move(dr));
    H->dof_cols.push_back(std::move(dc));
  }
  *out = H;
  return AFB_OK;
}

int afb_xplan_host_destroy(afb_xplan_host* h)
{
  delete h;
  return AFB_OK;
}

int afb_xplan_host_peers(const afb_xplan_host* h, int32_t* nb_peer, const int32_t** peer, const int64_t** send_first, const int64_t** send_count, const int64_t** recv_count)
{
  AFB_REQUIRE(h, AFB_ERR_INVALID, "afb_xplan_host_peers: null plan");
  if (nb_peer) *nb_peer = (int32_t)h->peer.size();
  if (peer) *peer = h->peer.data();
  if (send_first) *send_first = h->send_first.data();
  if (send_count) *send_count = h->send_count.data();
  if (recv_count) *recv_count = h->recv_count.data();
  return AFB_OK;
}

int afb_xplan_host_pairs(const afb_xplan_host* h, int32_t k, const int32_t** dof_rows, const int32_t** dof_cols)
{
  AFB_REQUIRE(h && k >= 0 && k < (int32_t)h->peer.size(), AFB_ERR_INVALID, "afb_xplan_host_pairs: bad neighbour index");
  if (dof_rows) *dof_rows = h->dof_rows[k].data();
  if (dof_cols) *dof_cols = h->dof_cols[k].data();
  return AFB_OK;
}

int afb_xplan_host_numbering(afb_xplan_host* h, int64_t* first_dof, int32_t* dof_l2g)
{
  AFB_REQUIRE(h && first_dof && dof_l2g, AFB_ERR_INVALID, "afb_xplan_host_numbering: null argument");
  const afb_transport& t = h->t;
  const int W = t.world, me = t.rank, b = h->b;
  // allGather(nb_own_row) + exclusive scan (femutils/HypreDoFLinearSystem.cc:224-233)
  const int64_t mine = (int64_t)h->nb_own * b;
  std::vector<int64_t> all((size_t)W, 0);
  AFB_REQUIRE(t.allgather(t.user, &mine, (int64_t)sizeof(int64_t), all.data()) == 0, AFB_ERR_INVALID, "transport: allgather failed");
  first_dof[0] = 0;
  for (int q = 0; q < W; ++q) first_dof[q + 1] = first_dof[q] + all[q];
  AFB_REQUIRE(first_dof[W] < (1ll << 31), AFB_ERR_OVERFLOW, "global row index exceeds Int32 (HYPRE_Int)");
  std::vector<int64_t> l2g_node((size_t)h->nb_node, -1);
  for (int32_t i = 0; i < h->nb_own; ++i) l2g_node[i] = first_dof[me] / b + i;
  // ghost nodes ask their owner for its local id (the reference's variable synchronize(), :236-245)
  std::vector<std::vector<int64_t>> ask((size_t)W), asked, answer((size_t)W), back;
  for (const auto& R : h->ghost_ranges) ask[R.owner].assign(h->node_gid.begin() + R.g0, h->node_gid.begin() + R.g1);
  AFB_TRY(exchange_i64(t, ask, asked));
  for (int q = 0; q < W; ++q) {
    answer[q].resize(asked[q].size());
    for (size_t i = 0; i < asked[q].size(); ++i) {
      const int32_t lid = h->local_of(asked[q][i]);
      AFB_REQUIRE(lid >= 0 && lid < h->nb_own, AFB_ERR_INVALID, "numbering: rank %d asks for node %lld, which rank %d does not own", q, (long long)asked[q][i], me);
      answer[q][i] = lid;
    }
  }
  AFB_TRY(exchange_i64(t, answer, back));
  for (const auto& R : h->ghost_ranges) {
    AFB_REQUIRE((int32_t)back[R.owner].size() == R.g1 - R.g0, AFB_ERR_INVALID, "numbering: owner %d answered %zu of %d ghost nodes", R.owner, back[R.owner].size(), R.g1 - R.g0);
    for (int32_t g = R.g0; g < R.g1; ++g) l2g_node[g] = first_dof[R.owner] / b + back[R.owner][g - R.g0];
  }
  for (int32_t i = 0; i < h->nb_node; ++i)
    for (int k = 0; k < b; ++k) dof_l2g[(size_t)i * b + k] = (int32_t)(l2g_node[i] * b + k);
  return AFB_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------
// exchange on a context
// ---------------------------------------------------------------------------------------------
struct afb_xplan {
  afb_ctx* ctx = nullptr;
  afb_xplan_host* host = nullptr;
  afb_transport t;
  int kind = 0; // 0 none, 1 peer memory, 2 callback
  std::string why_not_p2p;
  std::vector<DevBuf> slots;   // per peer: value slot of every double it sends
  std::vector<DevBuf> recvbuf; // callback transport: device landing buffers
  std::vector<std::vector<double>> hsend, hrecv; // callback transport without device buffers: host staging
  int64_t bytes_sent = 0, bytes_recv = 0;
  const void* values_base = nullptr;
};

// afb_destroy of the plan's context: release what lives on that context now, keep the host part for afb_xplan_destroy
void xplan_detach(afb_xplan* x)
{
  if (!x || !x->ctx) return;
  if (x->kind == 1) p2p_disconnect(x->ctx);
  for (auto& b : x->slots) b.release();
  for (auto& b : x->recvbuf) b.release();
  x->ctx = nullptr;
  x->kind = 0;
}

extern "C" {

int afb_xplan_destroy(afb_xplan* x)
{
  if (!x) return AFB_OK;
  if (x->ctx) {
    auto& v = x->ctx->xplans;
    v.erase(std::remove(v.begin(), v.end(), x), v.end());
    cudaSetDevice(x->ctx->device);
    if (x->kind == 1) p2p_disconnect(x->ctx);
  }
  for (auto& b : x->slots) b.release();
  for (auto& b : x->recvbuf) b.release();
  afb_xplan_host_destroy(x->host);
  delete x;
  return AFB_OK;
}

int afb_xplan_create(afb_ctx* ctx, const afb_transport* t, const int64_t* node_gid, const int32_t* node_owner, int32_t nb_own, int allow_peer_memory, afb_xplan** out)
{
  AFB_REQUIRE(ctx && out && t, AFB_ERR_INVALID, "afb_xplan_create: null argument");
  AFB_REQUIRE(ctx->has_pattern && ctx->values.p, AFB_ERR_INVALID, "afb_xplan_create: build the pattern and assemble once first (the assembly fixes the value layout)");
  *out = nullptr;
  AFB_CUDA(cudaSetDevice(ctx->device));
  const int32_t nb_node = ctx->nb_node;
  const int b = ctx->b;
  AFB_REQUIRE(nb_own >= 0 && nb_own <= nb_node, AFB_ERR_INVALID, "afb_xplan_create: nb_own_node out of range");
  AFB_TRY(verify_pending(ctx));
  // block pattern of the ghost rows
  std::vector<int32_t> rows_tail((size_t)(nb_node - nb_own) + 1);
  AFB_CUDA(cudaMemcpyAsync(rows_tail.data(), ctx->rows.as<int32_t>() + nb_own, sizeof(int32_t) * rows_tail.size(), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  const int64_t c0 = rows_tail.front(), c1 = rows_tail.back();
  std::vector<int32_t> cols_tail((size_t)std::max<int64_t>(c1 - c0, 1));
  if (c1 > c0) AFB_CUDA(cudaMemcpyAsync(cols_tail.data(), ctx->cols.as<int32_t>() + c0, sizeof(int32_t) * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, ctx->stream));
  AFB_CUDA(cudaStreamSynchronize(ctx->stream));
  afb_xplan* X = new afb_xplan();
  X->ctx = ctx;
  X->t = *t;
  int rc = afb_xplan_host_create(t, b, ctx->layout, nb_node, nb_own, node_gid, node_owner, rows_tail.data(), cols_tail.data(), &X->host);
  if (rc != AFB_OK) {
    delete X;
    return rc;
  }
  const afb_xplan_host& H = *X->host;
  const int np = (int)H.peer.size();
  X->slots.resize(np);
  auto bail = [&](int code) {
    afb_xplan_destroy(X);
    return code;
  };
  // value slot of every double a neighbour sends (its memory order, this rank's layout)
  int bad = 0;
  for (int k = 0; k < np; ++k) {
    const int64_t n = H.recv_count[k];
    X->bytes_recv += 8 * n;
    X->bytes_sent += 8 * H.send_count[k];
    if (n == 0) continue;
    DevBuf dr, dc;
    if (dr.reserve(sizeof(int32_t) * (size_t)n) || dc.reserve(sizeof(int32_t) * (size_t)n) || X->slots[k].reserve(sizeof(int64_t) * (size_t)n)) return bail(AFB_ERR_CUDA);
    cudaMemcpyAsync(dr.p, H.dof_rows[k].data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(dc.p, H.dof_cols[k].data(), sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream);
    rc = lookup_value_slots(ctx, n, dr.as<int32_t>(), dc.as<int32_t>(), X->slots[k].as<int64_t>());
    std::vector<int64_t> hs((size_t)n);
    cudaMemcpyAsync(hs.data(), X->slots[k].p, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    dr.release();
    dc.release();
    if (rc != AFB_OK) return bail(rc);
    for (int64_t s : hs) bad |= s < 0;
  }
  if (bad) {
    set_error("afb_xplan_create: a neighbour's partial row has an entry outside this rank's pattern");
    return bail(AFB_ERR_INVALID);
  }
  X->values_base = ctx->values.p;
  if (t->world == 1 || np == 0) X->kind = np == 0 ? 0 : 2;
  // ---- peer memory: every step is attempted on every rank and the outcome agreed on collectively ----
  int ok = allow_peer_memory ? 1 : 0;
  P2PEndpoint mine;
  memset(&mine, 0, sizeof(mine));
  if (ok && p2p_export_ex(ctx, &mine) != AFB_OK) {
    ok = 0;
    X->why_not_p2p = afb_last_error();
  }
  std::vector<P2PEndpoint> all((size_t)t->world);
  if (t->allgather(t->user, &mine, (int64_t)sizeof(P2PEndpoint), all.data()) != 0) {
    set_error("transport: allgather failed");
    return bail(AFB_ERR_INVALID);
  }
  // the slice every neighbour pulls from me is what I told it; what I pull from it is what it tells me
  std::vector<std::vector<int64_t>> tell((size_t)t->world), told;
  for (int k = 0; k < np; ++k) tell[H.peer[k]] = { H.send_first[k], H.send_count[k] };
  rc = exchange_i64(*t, tell, told);
  if (rc != AFB_OK) return bail(rc);
  if (ok && np > 0) {
    std::vector<P2PEndpoint> eps((size_t)np);
    std::vector<int64_t> pull_first((size_t)np, 0), pull_count((size_t)np, 0);
    std::vector<const int64_t*> sl((size_t)np, nullptr);
    for (int k = 0; k < np; ++k) {
      eps[k] = all[H.peer[k]];
      if (told[H.peer[k]].size() == 2) {
        pull_first[k] = told[H.peer[k]][0];
        pull_count[k] = told[H.peer[k]][1];
      }
      if (pull_count[k] != H.recv_count[k]) ok = 0;
      sl[k] = X->slots[k].as<int64_t>();
    }
    if (!ok) X->why_not_p2p = "a neighbour's slice and the local slot list differ in length";
    else if (p2p_connect_ex(ctx, t->rank, np, H.peer.data(), nullptr, nullptr, eps.data(), pull_first.data(), pull_count.data(), sl.data(), H.send_first.data(), H.send_count.data()) != AFB_OK) {
      ok = 0;
      X->why_not_p2p = afb_last_error();
    }
  }
  int32_t flag = ok ? 0 : 1;
  std::vector<int32_t> flags((size_t)t->world, 0);
  if (t->allgather(t->user, &flag, (int64_t)sizeof(int32_t), flags.data()) != 0) {
    set_error("transport: allgather failed");
    return bail(AFB_ERR_INVALID);
  }
  int any_bad = 0;
  for (int32_t f : flags) any_bad |= f;
  if (!any_bad && np > 0) X->kind = 1;
  else if (np > 0) {
    if (ok) {
      p2p_disconnect(ctx);
      if (X->why_not_p2p.empty()) X->why_not_p2p = allow_peer_memory ? "peer-memory mapping failed on another rank" : "not requested";
    }
    X->kind = 2;
    X->recvbuf.resize(np);
    X->hsend.resize(np);
    X->hrecv.resize(np);
    for (int k = 0; k < np; ++k) {
      if (H.recv_count[k] && X->recvbuf[k].reserve(sizeof(double) * (size_t)H.recv_count[k])) return bail(AFB_ERR_CUDA);
      if (!t->exchange_takes_device_memory) {
        X->hsend[k].resize((size_t)H.send_count[k]);
        X->hrecv[k].resize((size_t)H.recv_count[k]);
      }
    }
  }
  ctx->xplans.push_back(X);
  *out = X;
  return AFB_OK;
}

int afb_xplan_exchange(afb_xplan* x)
{
  afb::NvtxRange nvtx_range("GhostRowExchange");
  AFB_REQUIRE(x, AFB_ERR_INVALID, "afb_xplan_exchange: null plan");
  AFB_REQUIRE(x->ctx, AFB_ERR_INVALID, "afb_xplan_exchange: the plan's context was destroyed");
  afb_ctx* ctx = x->ctx;
  AFB_REQUIRE(ctx->values.p == x->values_base, AFB_ERR_INVALID, "afb_xplan_exchange: the values array moved since afb_xplan_create (create the plan again)");
  if (x->kind == 0) return AFB_OK;
  if (x->kind == 1) return p2p_exchange(ctx, 1);
  // callback transport: rows leave in place (device-capable transport) or through host staging, then accumulate + zero
  const afb_xplan_host& H = *x->host;
  const int np = (int)H.peer.size();
  double* values = ctx->values.as<double>();
  std::vector<const void*> sp((size_t)np);
  std::vector<void*> rp((size_t)np);
  std::vector<int64_t> sb((size_t)np), rb((size_t)np);
  const bool dev = x->t.exchange_takes_device_memory != 0;
  for (int k = 0; k < np; ++k) {
    sb[k] = 8 * H.send_count[k];
    rb[k] = 8 * H.recv_count[k];
    if (dev) {
      sp[k] = values + H.send_first[k];
      rp[k] = x->recvbuf[k].p;
    }
    else {
      if (sb[k]) AFB_CUDA(cudaMemcpyAsync(x->hsend[k].data(), values + H.send_first[k], (size_t)sb[k], cudaMemcpyDeviceToHost, ctx->stream));
      sp[k] = x->hsend[k].data();
      rp[k] = x->hrecv[k].data();
    }
  }
  AFB_CUDA(cudaStreamSynchronize(ctx->stream)); // the assembly (and the staging copies) are complete before the transport reads
  AFB_REQUIRE(x->t.exchange(x->t.user, np, H.peer.data(), sp.data(), sb.data(), rp.data(), rb.data(), dev ? 1 : 0) == 0, AFB_ERR_INVALID, "transport: exchange failed");
  for (int k = 0; k < np; ++k) {
    if (sb[k]) AFB_CUDA(cudaMemsetAsync(values + H.send_first[k], 0, (size_t)sb[k], ctx->stream)); // ghost rows are zero in the reference (isOwn gates)
    if (!rb[k]) continue;
    if (!dev) AFB_CUDA(cudaMemcpyAsync(x->recvbuf[k].p, x->hrecv[k].data(), (size_t)rb[k], cudaMemcpyHostToDevice, ctx->stream));
    AFB_TRY(add_values_at(ctx, H.recv_count[k], x->slots[k].as<int64_t>(), x->recvbuf[k].as<double>()));
  }
  return AFB_OK;
}

int afb_xplan_wait(afb_xplan* x)
{
  AFB_REQUIRE(x, AFB_ERR_INVALID, "afb_xplan_wait: null plan");
  if (x->kind == 1) return p2p_wait(x->ctx);
  return AFB_OK;
}

int afb_xplan_numbering(afb_xplan* x, int64_t* first_dof, int32_t* dof_l2g)
{
  AFB_REQUIRE(x, AFB_ERR_INVALID, "afb_xplan_numbering: null plan");
  return afb_xplan_host_numbering(x->host, first_dof, dof_l2g);
}

int afb_xplan_info(const afb_xplan* x, int32_t* nb_peer, int64_t* bytes_sent, int64_t* bytes_received, int32_t* transport_kind, const char** why_not_peer_memory)
{
  AFB_REQUIRE(x, AFB_ERR_INVALID, "afb_xplan_info: null plan");
  if (nb_peer) *nb_peer = (int32_t)x->host->peer.size();
  if (bytes_sent) *bytes_sent = x->bytes_sent;
  if (bytes_received) *bytes_received = x->bytes_recv;
  if (transport_kind) *transport_kind = x->kind;
  if (why_not_peer_memory) *why_not_peer_memory = x->why_not_p2p.c_str();
  return AFB_OK;
}

} // extern "C"
