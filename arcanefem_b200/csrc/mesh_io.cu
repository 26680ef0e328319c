// Gmsh 4.1 reader (binary and ASCII) behind the C ABI: the mesh-ingestion step of SURVEY.md §8(f).3.
//
// The reference leaves this to Arcane's MshMeshReader (selected by `<filename>meshes/*.msh</filename>` in every .arc file, e.g.
// modules/testlab/inputs/Test.L-shape.2D.arc:17-21).  What the assembly path sees of it, and what this file reproduces
// (verified against the reference's golden solution files, which are keyed by node uniqueId):
//  * node uniqueId = gmsh node tag; node local id = rank of the tag in ascending order;
//  * cells = the elements of the highest dimension present, one cell type per mesh, ordered by element tag;
//  * a named surface (`<surface>left</surface>`) = the (dim-1) elements of every entity carrying that physical name; its node
//    group = the union of their nodes (modules/testlab/FemModule.cc:657-663 `face_group.nodeGroup()`);
//  * a named volume (`<material-property><volume>`) = the cells of the entities carrying the name;
//  * a named point (`<dirichlet-point><node>`) = the nodes of the 0-dimensional elements carrying the name.
// Host-only code: no CUDA call, usable before a context exists.
#include "afb_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// gmsh element type -> {dimension, nodes}
bool gmsh_type(int t, int& dim, int& npn)
{
  switch (t) {
  case 15: dim = 0; npn = 1; return true;
  case 1: dim = 1; npn = 2; return true;
  case 8: dim = 1; npn = 3; return true;
  case 2: dim = 2; npn = 3; return true;
  case 3: dim = 2; npn = 4; return true;
  case 9: dim = 2; npn = 6; return true;
  case 16: dim = 2; npn = 8; return true;
  case 10: dim = 2; npn = 9; return true;
  case 4: dim = 3; npn = 4; return true;
  case 5: dim = 3; npn = 8; return true;
  case 6: dim = 3; npn = 6; return true;
  case 7: dim = 3; npn = 5; return true;
  case 11: dim = 3; npn = 10; return true;
  case 17: dim = 3; npn = 20; return true;
  case 12: dim = 3; npn = 27; return true;
  }
  return false;
}

struct Failure {
  std::string what;
};

// cursor over the file image; every read is bounds-checked and throws Failure
struct Cursor {
  const char* p;
  const char* end;
  bool binary = false;

  [[noreturn]] void fail(const std::string& w) const { throw Failure{ w }; }

  std::string line()
  {
    if (p >= end) fail("unexpected end of file");
    const char* e = (const char*)memchr(p, '\n', (size_t)(end - p));
    const char* stop = e ? e : end;
    const char* a = p;
    const char* b = stop;
    while (a < b && (*a == ' ' || *a == '\t' || *a == '\r')) ++a;
    while (b > a && (b[-1] == ' ' || b[-1] == '\t' || b[-1] == '\r')) --b;
    p = e ? e + 1 : end;
    return std::string(a, b);
  }
  void expect(const char* tag)
  {
    std::string s = line();
    while (s.empty() && p < end) s = line();
    if (s != tag) fail(std::string("expected ") + tag + ", found '" + s.substr(0, 40) + "'");
  }
  template <typename T> T raw()
  {
    if ((size_t)(end - p) < sizeof(T)) fail("truncated binary section");
    T v;
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  void skip(size_t bytes)
  {
    if ((size_t)(end - p) < bytes) fail("truncated binary section");
    p += bytes;
  }
  long long ascii_int()
  {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    if (p >= end) fail("unexpected end of file");
    char* e = nullptr;
    const long long v = strtoll(p, &e, 10);
    if (e == p) fail("integer expected");
    p = e;
    return v;
  }
  double ascii_real()
  {
    while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    if (p >= end) fail("unexpected end of file");
    char* e = nullptr;
    const double v = strtod(p, &e);
    if (e == p) fail("number expected");
    p = e;
    return v;
  }
  // the two encodings of the format's scalar kinds
  long long size_value() { return binary ? (long long)raw<uint64_t>() : ascii_int(); }
  long long int_value() { return binary ? (long long)raw<int32_t>() : ascii_int(); }
  double real_value() { return binary ? raw<double>() : ascii_real(); }
};

struct Block {
  int dim, entity, type, npn;
  std::vector<int64_t> rows; // [n][1 + npn]: element tag, node tags
  int64_t count() const { return (int64_t)rows.size() / (1 + npn); }
};

struct Group {
  std::string name;
  int kind;                   // AFB_MSH_GROUP_*
  int npi = 1;                // nodes per face
  std::vector<int32_t> items; // faces: [n][npi] node ids; cells: cell ids; points: node ids
  std::vector<int32_t> nodes; // sorted unique node ids (faces, points)
};

void sort_unique(std::vector<int32_t>& v)
{
  std::sort(v.begin(), v.end());
  v.erase(std::unique(v.begin(), v.end()), v.end());
}

} // namespace

struct afb_msh {
  int dim = 0, npc = 0;
  std::vector<double> xyz;
  std::vector<int64_t> uid;
  std::vector<int32_t> cells;
  std::vector<Group> groups;
};

namespace {

void parse(Cursor& c, afb_msh& M)
{
  std::map<std::pair<int, int>, std::string> phys_name;      // (dim, physical tag) -> name
  std::map<std::pair<int, int>, std::vector<int>> ent_phys;  // (dim, entity tag) -> physical tags
  std::vector<int64_t> node_tag;
  std::vector<double> node_xyz;
  std::vector<Block> blocks;
  bool have_format = false;

  while (c.p < c.end) {
    const std::string tag = c.line();
    if (tag.empty()) continue;
    if (tag == "$MeshFormat") {
      const std::string h = c.line();
      double version = 0;
      int file_type = -1, data_size = 0;
      if (sscanf(h.c_str(), "%lf %d %d", &version, &file_type, &data_size) != 3) c.fail("bad $MeshFormat line");
      if (version < 4.1 || version >= 4.2) c.fail("need the msh 4.1 format, file says '" + h + "'");
      if (data_size != 8) c.fail("need data-size 8");
      c.binary = file_type == 1;
      if (c.binary) {
        if (c.raw<int32_t>() != 1) c.fail("big-endian msh files are not supported");
        c.line();
      }
      c.expect("$EndMeshFormat");
      have_format = true;
    }
    else if (!have_format) {
      c.fail("not a msh file ($MeshFormat expected first)");
    }
    else if (tag == "$PhysicalNames") {
      const int n = atoi(c.line().c_str());
      for (int i = 0; i < n; ++i) {
        const std::string s = c.line();
        int d = 0, t = 0, used = 0;
        if (sscanf(s.c_str(), "%d %d %n", &d, &t, &used) < 2) c.fail("bad $PhysicalNames entry");
        std::string name = s.substr((size_t)used);
        if (name.size() >= 2 && name.front() == '"' && name.back() == '"') name = name.substr(1, name.size() - 2);
        phys_name[{ d, t }] = name;
      }
      c.expect("$EndPhysicalNames");
    }
    else if (tag == "$Entities") {
      long long counts[4];
      for (auto& k : counts) k = c.size_value();
      for (int d = 0; d < 4; ++d)
        for (long long i = 0; i < counts[d]; ++i) {
          const int etag = (int)c.int_value();
          for (int k = 0; k < (d == 0 ? 3 : 6); ++k) c.real_value();
          const long long nphys = c.size_value();
          std::vector<int>& ph = ent_phys[{ d, etag }];
          for (long long k = 0; k < nphys; ++k) ph.push_back((int)c.int_value());
          if (d > 0) {
            const long long nb = c.size_value();
            for (long long k = 0; k < nb; ++k) c.int_value();
          }
        }
      if (c.binary) c.line();
      c.expect("$EndEntities");
    }
    else if (tag == "$Nodes") {
      const long long nblocks = c.size_value();
      const long long nnodes = c.size_value();
      c.size_value();
      c.size_value();
      if (nnodes < 0 || nnodes > 0x7fffffffLL) c.fail("node count out of range");
      node_tag.reserve((size_t)nnodes);
      node_xyz.reserve(3 * (size_t)nnodes);
      for (long long b = 0; b < nblocks; ++b) {
        c.int_value();
        c.int_value();
        const long long parametric = c.int_value();
        const long long n = c.size_value();
        if (parametric != 0) c.fail("parametric node blocks are not supported");
        if (n < 0 || (long long)node_tag.size() + n > nnodes) c.fail("node blocks exceed the announced node count");
        for (long long i = 0; i < n; ++i) node_tag.push_back(c.size_value());
        for (long long i = 0; i < 3 * n; ++i) node_xyz.push_back(c.real_value());
      }
      if (c.binary) c.line();
      c.expect("$EndNodes");
    }
    else if (tag == "$Elements") {
      const long long nblocks = c.size_value();
      c.size_value();
      c.size_value();
      c.size_value();
      for (long long b = 0; b < nblocks; ++b) {
        Block B;
        B.dim = (int)c.int_value();
        B.entity = (int)c.int_value();
        B.type = (int)c.int_value();
        const long long n = c.size_value();
        int tdim = 0;
        if (!gmsh_type(B.type, tdim, B.npn)) c.fail("unknown gmsh element type " + std::to_string(B.type));
        if (n < 0) c.fail("negative element count");
        const size_t words = (size_t)n * (size_t)(1 + B.npn);
        if (c.binary && (size_t)(c.end - c.p) < 8 * words) c.fail("truncated $Elements section");
        B.rows.resize(words);
        for (size_t i = 0; i < words; ++i) B.rows[i] = c.size_value();
        blocks.push_back(std::move(B));
      }
      if (c.binary) c.line();
      c.expect("$EndElements");
    }
    else if (tag[0] == '$') {
      // a section the assembly path does not use (periodic links, ghost elements, parametrisations, post-processing views)
      const std::string endtag = "$End" + tag.substr(1);
      const char* q = c.p;
      const char* found = nullptr;
      while (q < c.end) {
        const char* d = (const char*)memchr(q, '$', (size_t)(c.end - q));
        if (!d) break;
        if ((size_t)(c.end - d) >= endtag.size() && memcmp(d, endtag.data(), endtag.size()) == 0) { found = d; break; }
        q = d + 1;
      }
      if (!found) c.fail("section " + tag + " is not closed");
      c.p = found;
      c.line();
    }
    else {
      c.fail("unexpected text outside a section: '" + tag.substr(0, 40) + "'");
    }
  }
  if (node_tag.empty()) c.fail("no $Nodes section");
  if (blocks.empty()) c.fail("no $Elements section");

  // local ids: rank of the tag
  const size_t nn = node_tag.size();
  std::vector<int32_t> order(nn);
  for (size_t i = 0; i < nn; ++i) order[i] = (int32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return node_tag[a] < node_tag[b]; });
  M.uid.resize(nn);
  M.xyz.resize(3 * nn);
  std::unordered_map<int64_t, int32_t> lid;
  lid.reserve(2 * nn);
  for (size_t i = 0; i < nn; ++i) {
    const int32_t s = order[i];
    M.uid[i] = node_tag[s];
    for (int k = 0; k < 3; ++k) M.xyz[3 * i + k] = node_xyz[3 * (size_t)s + k];
    if (!lid.emplace(node_tag[s], (int32_t)i).second) c.fail("node tag " + std::to_string(node_tag[s]) + " appears twice");
  }
  auto local = [&](int64_t t) -> int32_t {
    auto it = lid.find(t);
    if (it == lid.end()) c.fail("an element references node tag " + std::to_string(t) + ", which $Nodes does not define");
    return it->second;
  };

  M.dim = 0;
  for (const Block& B : blocks) M.dim = std::max(M.dim, B.dim);
  if (M.dim < 2) c.fail("no 2-D or 3-D elements in the file");
  int cell_type = -1;
  size_t nb_cell = 0;
  for (const Block& B : blocks)
    if (B.dim == M.dim) {
      if (cell_type >= 0 && cell_type != B.type) c.fail("mixed cell types (" + std::to_string(cell_type) + " and " + std::to_string(B.type) + ")");
      cell_type = B.type;
      M.npc = B.npn;
      nb_cell += (size_t)B.count();
    }
  // cells ordered by element tag
  struct Ref { int64_t tag; const int64_t* row; };
  std::vector<Ref> refs;
  refs.reserve(nb_cell);
  for (const Block& B : blocks)
    if (B.dim == M.dim)
      for (int64_t i = 0; i < B.count(); ++i) refs.push_back({ B.rows[(size_t)i * (1 + B.npn)], &B.rows[(size_t)i * (1 + B.npn)] });
  std::stable_sort(refs.begin(), refs.end(), [](const Ref& a, const Ref& b) { return a.tag < b.tag; });
  M.cells.resize(nb_cell * (size_t)M.npc);
  std::unordered_map<int64_t, int32_t> cell_of_tag;
  cell_of_tag.reserve(2 * nb_cell);
  for (size_t i = 0; i < nb_cell; ++i) {
    for (int a = 0; a < M.npc; ++a) M.cells[i * M.npc + a] = local(refs[i].row[1 + a]);
    cell_of_tag[refs[i].tag] = (int32_t)i;
  }

  // groups, in the order their first block appears in the file: cell groups, face groups, then point groups
  auto find_group = [&](const std::string& name, int kind) -> Group* {
    for (Group& g : M.groups)
      if (g.name == name && g.kind == kind) return &g;
    return nullptr;
  };
  auto names_of = [&](const Block& B) {
    std::vector<std::string> out;
    auto e = ent_phys.find({ B.dim, B.entity });
    if (e != ent_phys.end())
      for (int pt : e->second) {
        auto n = phys_name.find({ B.dim, pt });
        if (n != phys_name.end()) out.push_back(n->second);
      }
    return out;
  };
  for (const Block& B : blocks)
    if (B.dim == M.dim)
      for (const std::string& name : names_of(B)) {
        Group* g = find_group(name, AFB_MSH_GROUP_CELLS);
        if (!g) { M.groups.push_back(Group{ name, AFB_MSH_GROUP_CELLS }); g = &M.groups.back(); }
        for (int64_t i = 0; i < B.count(); ++i) g->items.push_back(cell_of_tag[B.rows[(size_t)i * (1 + B.npn)]]);
      }
  for (Group& g : M.groups) sort_unique(g.items);
  for (const Block& B : blocks)
    if (B.dim == M.dim - 1)
      for (const std::string& name : names_of(B)) {
        Group* g = find_group(name, AFB_MSH_GROUP_FACES);
        if (!g) { M.groups.push_back(Group{ name, AFB_MSH_GROUP_FACES, B.npn }); g = &M.groups.back(); }
        if (g->npi != B.npn) c.fail("surface '" + name + "' mixes face types");
        for (int64_t i = 0; i < B.count(); ++i)
          for (int a = 0; a < B.npn; ++a) g->items.push_back(local(B.rows[(size_t)i * (1 + B.npn) + 1 + a]));
      }
  for (Group& g : M.groups)
    if (g.kind == AFB_MSH_GROUP_FACES) { g.nodes = g.items; sort_unique(g.nodes); }
  for (const Block& B : blocks)
    if (B.dim == 0)
      for (const std::string& name : names_of(B)) {
        if (find_group(name, AFB_MSH_GROUP_FACES)) continue; // a surface of that name wins
        Group* g = find_group(name, AFB_MSH_GROUP_POINTS);
        if (!g) { M.groups.push_back(Group{ name, AFB_MSH_GROUP_POINTS }); g = &M.groups.back(); }
        for (int64_t i = 0; i < B.count(); ++i) g->items.push_back(local(B.rows[(size_t)i * 2 + 1]));
      }
  for (Group& g : M.groups)
    if (g.kind == AFB_MSH_GROUP_POINTS) { sort_unique(g.items); g.nodes = g.items; }
}

} // namespace

extern "C" {

int afb_msh_read(const char* path, afb_msh** out)
{
  AFB_REQUIRE(path && out, AFB_ERR_INVALID, "afb_msh_read: null argument");
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  AFB_REQUIRE(f != nullptr, AFB_ERR_INVALID, "afb_msh_read: cannot open '%s'", path);
  std::vector<char> image;
  {
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) image.insert(image.end(), buf, buf + n);
    fclose(f);
  }
  image.push_back('\0'); // strtod / strtoll stop here
  afb_msh* M = new afb_msh();
  Cursor c{ image.data(), image.data() + image.size() - 1 };
  try {
    parse(c, *M);
  }
  catch (const Failure& e) {
    delete M;
    AFB_REQUIRE(false, AFB_ERR_INVALID, "afb_msh_read('%s'): %s (byte %lld)", path, e.what.c_str(), (long long)(c.p - image.data()));
  }
  catch (const std::bad_alloc&) {
    delete M;
    AFB_REQUIRE(false, AFB_ERR_INVALID, "afb_msh_read('%s'): out of host memory", path);
  }
  *out = M;
  return AFB_OK;
}

int afb_msh_destroy(afb_msh* m)
{
  delete m;
  return AFB_OK;
}

int afb_msh_sizes(const afb_msh* m, int* dim, int* nodes_per_cell, int32_t* nb_node, int64_t* nb_cell, int32_t* nb_group)
{
  AFB_REQUIRE(m, AFB_ERR_INVALID, "afb_msh_sizes: null mesh");
  if (dim) *dim = m->dim;
  if (nodes_per_cell) *nodes_per_cell = m->npc;
  if (nb_node) *nb_node = (int32_t)m->uid.size();
  if (nb_cell) *nb_cell = (int64_t)(m->cells.size() / (size_t)m->npc);
  if (nb_group) *nb_group = (int32_t)m->groups.size();
  return AFB_OK;
}

int afb_msh_get(const afb_msh* m, double* xyz, int32_t* cell_nodes, int64_t* node_uid)
{
  AFB_REQUIRE(m, AFB_ERR_INVALID, "afb_msh_get: null mesh");
  if (xyz) memcpy(xyz, m->xyz.data(), m->xyz.size() * sizeof(double));
  if (cell_nodes) memcpy(cell_nodes, m->cells.data(), m->cells.size() * sizeof(int32_t));
  if (node_uid) memcpy(node_uid, m->uid.data(), m->uid.size() * sizeof(int64_t));
  return AFB_OK;
}

int afb_msh_group(const afb_msh* m, int32_t g, const char** name, int* kind, int64_t* nb_item, int* nodes_per_item, int64_t* nb_group_node)
{
  AFB_REQUIRE(m && g >= 0 && g < (int32_t)m->groups.size(), AFB_ERR_INVALID, "afb_msh_group: group %d of %d", g, m ? (int)m->groups.size() : 0);
  const Group& G = m->groups[g];
  if (name) *name = G.name.c_str();
  if (kind) *kind = G.kind;
  if (nb_item) *nb_item = (int64_t)(G.items.size() / (size_t)G.npi);
  if (nodes_per_item) *nodes_per_item = G.npi;
  if (nb_group_node) *nb_group_node = (int64_t)G.nodes.size();
  return AFB_OK;
}

int afb_msh_group_get(const afb_msh* m, int32_t g, int32_t* items, int32_t* nodes)
{
  AFB_REQUIRE(m && g >= 0 && g < (int32_t)m->groups.size(), AFB_ERR_INVALID, "afb_msh_group_get: group %d of %d", g, m ? (int)m->groups.size() : 0);
  const Group& G = m->groups[g];
  if (items) memcpy(items, G.items.data(), G.items.size() * sizeof(int32_t));
  if (nodes) memcpy(nodes, G.nodes.data(), G.nodes.size() * sizeof(int32_t));
  return AFB_OK;
}

} // extern "C"
