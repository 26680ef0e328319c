"""Host-side mesh inputs for the assembly path: Gmsh 4.1 binary reader and the
synthetic structured-box generators of SURVEY.md §8(d).

Nothing here is on the hot path; it produces the plain arrays the C ABI takes
(`afb_set_mesh`): AoS coordinates ``float64[nb_node,3]`` (Arcane ``VariableNodeReal3``
layout) and ``int32[nb_cell,npc]`` connectivity (``cnc.nodeId(cell,i)``).

Reference conventions mirrored (verified against the golden files):
node uniqueId = gmsh node tag, node local id = rank of the tag in ascending order;
a `<surface>` Dirichlet set = union of the nodes of all (dim-1) elements whose
entity carries that physical name (modules/testlab/FemModule.cc:657-663).
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field

import numpy as np

# gmsh element type -> (dim, nodes)
_GMSH_TYPES = {15: (0, 1), 1: (1, 2), 2: (2, 3), 4: (3, 4), 8: (1, 3), 9: (2, 6), 11: (3, 10),
               3: (2, 4), 5: (3, 8), 10: (2, 9), 16: (2, 8), 17: (3, 20), 12: (3, 27), 6: (3, 6), 7: (3, 5)}


@dataclass
class Mesh:
    dim: int
    coords: np.ndarray            # float64 [nb_node,3]
    cells: np.ndarray             # int32 [nb_cell,npc]
    node_uid: np.ndarray          # int64 [nb_node] (gmsh tag; box meshes: = local id)
    # physical name -> int32 node ids (sorted unique) of its (dim-1) elements
    groups: dict = field(default_factory=dict)
    # physical name -> int32 cell ids of a named volume (material regions)
    cell_groups: dict = field(default_factory=dict)
    # physical name -> int32 [nb_face, nodes_per_face] boundary elements
    faces: dict = field(default_factory=dict)

    @property
    def nb_node(self):
        return int(self.coords.shape[0])

    @property
    def nb_cell(self):
        return int(self.cells.shape[0])

    @property
    def npc(self):
        return int(self.cells.shape[1])


def read_msh(path: str) -> Mesh:
    """Gmsh 4.1 binary reader (enough for Tri3/Tet4 (+Tri6/Tet10) cells and physical groups)."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0

    def line():
        nonlocal pos
        e = data.index(b"\n", pos)
        s = data[pos:e].decode("ascii", "replace").strip()
        pos = e + 1
        return s

    def unpack(fmt):
        nonlocal pos
        sz = struct.calcsize(fmt)
        v = struct.unpack_from(fmt, data, pos)
        pos += sz
        return v

    phys_names = {}          # (dim, tag) -> name
    ent_phys = {}            # (dim, entity tag) -> [phys tags]
    node_tags = None
    node_xyz = None
    blocks = []              # (entityDim, entityTag, etype, ndarray[n, 1+npn])
    while pos < len(data):
        tag = line()
        if tag == "$MeshFormat":
            hdr = line().split()
            assert hdr[0].startswith("4.1") and hdr[1] == "1", f"need gmsh 4.1 binary, got {hdr}"
            one, = unpack("<i")
            assert one == 1, "big-endian msh unsupported"
            line()  # rest of line
            assert line() == "$EndMeshFormat"
        elif tag == "$PhysicalNames":
            n = int(line())
            for _ in range(n):
                parts = line().split(" ", 2)
                phys_names[(int(parts[0]), int(parts[1]))] = parts[2].strip().strip('"')
            assert line() == "$EndPhysicalNames"
        elif tag == "$Entities":
            counts = unpack("<4Q")
            for _ in range(counts[0]):
                etag, = unpack("<i")
                unpack("<3d")
                nphys, = unpack("<Q")
                ent_phys[(0, etag)] = list(unpack(f"<{nphys}i")) if nphys else []
            for d in (1, 2, 3):
                for _ in range(counts[d]):
                    etag, = unpack("<i")
                    unpack("<6d")
                    nphys, = unpack("<Q")
                    ent_phys[(d, etag)] = list(unpack(f"<{nphys}i")) if nphys else []
                    nb, = unpack("<Q")
                    if nb:
                        unpack(f"<{nb}i")
            line()
            assert line() == "$EndEntities"
        elif tag == "$Nodes":
            nblocks, nnodes, _mn, _mx = unpack("<4Q")
            tags = np.empty(nnodes, dtype=np.int64)
            xyz = np.empty((nnodes, 3), dtype=np.float64)
            k = 0
            for _ in range(nblocks):
                _ed, _et, param = unpack("<3i")
                n, = unpack("<Q")
                assert param == 0
                tags[k:k + n] = np.frombuffer(data, dtype="<u8", count=n, offset=pos)
                pos += 8 * n
                xyz[k:k + n] = np.frombuffer(data, dtype="<f8", count=3 * n, offset=pos).reshape(n, 3)
                pos += 24 * n
                k += n
            node_tags, node_xyz = tags, xyz
            line()
            assert line() == "$EndNodes"
        elif tag == "$Elements":
            nblocks, _ne, _mn, _mx = unpack("<4Q")
            for _ in range(nblocks):
                ed, et, etype = unpack("<3i")
                n, = unpack("<Q")
                npn = _GMSH_TYPES[etype][1]
                arr = np.frombuffer(data, dtype="<u8", count=n * (1 + npn), offset=pos).reshape(n, 1 + npn).astype(np.int64)
                pos += 8 * n * (1 + npn)
                blocks.append((ed, et, etype, arr))
            line()
            assert line() == "$EndElements"
        elif tag.startswith("$"):
            end = "$End" + tag[1:]
            e = data.index(end.encode(), pos)
            pos = e
            line()
        # else: blank line

    order = np.argsort(node_tags, kind="stable")
    uid = node_tags[order]
    coords = np.ascontiguousarray(node_xyz[order])
    lid_of_tag = {int(t): i for i, t in enumerate(uid)}
    tag2lid = np.vectorize(lambda t: lid_of_tag[int(t)], otypes=[np.int32])

    dim = max(ed for ed, _, _, _ in blocks)
    cell_blocks = [(ed, et, ty, a) for ed, et, ty, a in blocks if ed == dim]
    etypes = {ty for _, _, ty, _ in cell_blocks}
    assert len(etypes) == 1, f"mixed cell types {etypes}"
    # cells ordered by element tag (Arcane creates items in file order; fixtures are tag-ordered)
    allc = np.concatenate([a for _, _, _, a in cell_blocks], axis=0)
    allc = allc[np.argsort(allc[:, 0], kind="stable")]
    cells = tag2lid(allc[:, 1:]).astype(np.int32)
    cell_of_tag = {int(t): i for i, t in enumerate(allc[:, 0])}
    cell_groups = {}
    for ed, et, ty, a in cell_blocks:  # named volumes: <material-property><volume> of the reference's .arc files
        for ptag in ent_phys.get((ed, et), []):
            name = phys_names.get((ed, ptag))
            if name is not None:
                ids = np.array([cell_of_tag[int(t)] for t in a[:, 0]], dtype=np.int32)
                cell_groups[name] = np.unique(np.concatenate([cell_groups.get(name, np.empty(0, np.int32)), ids])).astype(np.int32)

    groups, faces = {}, {}
    for ed, et, ty, a in blocks:
        if ed != dim - 1:
            continue
        for ptag in ent_phys.get((ed, et), []):
            name = phys_names.get((ed, ptag))
            if name is None:
                continue
            fl = tag2lid(a[:, 1:]).astype(np.int32)
            faces.setdefault(name, []).append(fl)
    for name, lst in faces.items():
        fl = np.concatenate(lst, axis=0)
        faces[name] = fl
        groups[name] = np.unique(fl.ravel()).astype(np.int32)
    # named points (physical groups of dimension 0: the <dirichlet-point> targets of the reference's .arc files)
    for ed, et, ty, a in blocks:
        if ed != 0:
            continue
        for ptag in ent_phys.get((0, et), []):
            name = phys_names.get((0, ptag))
            if name is not None and name not in faces:
                pts = tag2lid(a[:, 1:]).astype(np.int32).ravel()
                groups[name] = np.unique(np.concatenate([groups.get(name, np.empty(0, np.int32)), pts])).astype(np.int32)
    return Mesh(dim=dim, coords=coords, cells=np.ascontiguousarray(cells), node_uid=uid.astype(np.int64), groups=groups, faces=faces, cell_groups=cell_groups)


def read_msh_native(path: str) -> Mesh:
    """The same mesh through the C ABI's reader (`afb_msh_*`, csrc/mesh_io.cu: Gmsh 4.1 binary and ASCII) -- what a C++ host links."""
    import ctypes as C

    from . import capi as A
    lib = A.lib()
    h = C.c_void_p()
    A._check(lib.afb_msh_read(os.fsencode(path), C.byref(h)))
    try:
        dim, npc, nn, ng = C.c_int(), C.c_int(), C.c_int32(), C.c_int32()
        nc = C.c_int64()
        A._check(lib.afb_msh_sizes(h, C.byref(dim), C.byref(npc), C.byref(nn), C.byref(nc), C.byref(ng)))
        coords = np.empty((nn.value, 3), dtype=np.float64)
        cells = np.empty((nc.value, npc.value), dtype=np.int32)
        uid = np.empty(nn.value, dtype=np.int64)
        A._check(lib.afb_msh_get(h, A._ptr(coords), A._ptr(cells), A._ptr(uid)))
        groups, faces, cell_groups = {}, {}, {}
        for g in range(ng.value):
            name, kind, npi = C.c_char_p(), C.c_int(), C.c_int()
            n_item, n_node = C.c_int64(), C.c_int64()
            A._check(lib.afb_msh_group(h, g, C.byref(name), C.byref(kind), C.byref(n_item), C.byref(npi), C.byref(n_node)))
            items = np.empty((n_item.value, npi.value), dtype=np.int32)
            nodes = np.empty(n_node.value, dtype=np.int32)
            A._check(lib.afb_msh_group_get(h, g, A._ptr(items), A._ptr(nodes)))
            key = name.value.decode()
            if kind.value == A.MSH_GROUP_CELLS:
                cell_groups[key] = items.ravel()
            elif kind.value == A.MSH_GROUP_FACES:
                faces[key] = items
                groups[key] = nodes
            else:
                groups[key] = nodes
    finally:
        lib.afb_msh_destroy(h)
    return Mesh(dim=dim.value, coords=coords, cells=cells, node_uid=uid, groups=groups, faces=faces, cell_groups=cell_groups)


def orient_boundary_faces(mesh: Mesh, faces: np.ndarray) -> np.ndarray:
    """Boundary faces (edges of a Tri3 mesh, triangles of a Tet4 mesh) with their first two nodes ordered so that the
    reference's normal formulas point OUT of the domain: N = (y1-y0, x0-x1) in 2-D, (n1-n0) x (n2-n0) in 3-D
    (femutils/ArcaneFemFunctionsGpu.h:159-214).  This is the swap those helpers apply when Arcane reports the face as
    not "subdomain boundary outside"; here the side is found from the cell the face belongs to."""
    faces = np.array(faces, dtype=np.int32, copy=True)
    nn = faces.shape[1]
    if mesh.cells.shape[1] == 2 ** mesh.dim:  # Quad4 (edges) / Hexa8 (Quad4 faces): side found from the cell's centroid
        return _orient_q1_faces(mesh, faces)
    corner = mesh.cells[:, :mesh.dim + 1]
    key = lambda a: tuple(sorted(int(x) for x in a))
    owner = {}
    for c, cn in enumerate(corner):
        for k in range(mesh.dim + 1):
            owner.setdefault(key(np.delete(cn, k)), (c, int(cn[k])))
    for f in range(faces.shape[0]):
        c, opp = owner[key(faces[f, :mesh.dim])]
        p0, p1 = mesh.coords[faces[f, 0]], mesh.coords[faces[f, 1]]
        if mesh.dim == 2:
            n = np.array([p1[1] - p0[1], p0[0] - p1[0], 0.0])
        else:
            n = np.cross(p1 - p0, mesh.coords[faces[f, 2]] - p0)
        if np.dot(n, mesh.coords[opp] - p0) > 0.0:
            faces[f, [0, 1]] = faces[f, [1, 0]]
    assert nn == mesh.dim
    return faces


def arcane_face_node_order(mesh: Mesh, faces: np.ndarray) -> np.ndarray:
    """Faces with their nodes in the order Arcane stores them: the node with the smallest unique id first, then round the
    face towards its neighbour with the smaller unique id.  The reference's Quad4-face flux integral (Hexa8 meshes,
    femutils/ArcaneFemFunctions.h:1843-1953) takes its normal dr/dxi x dr/deta from that order without the
    outside-of-the-domain test the P1 helpers make: its q.n term follows the numbering, not the geometry."""
    faces = np.asarray(faces, dtype=np.int32)
    out = np.empty_like(faces)
    nn = faces.shape[1]
    for f in range(faces.shape[0]):
        u = mesh.node_uid[faces[f]]
        k = int(np.argmin(u))
        step = 1 if u[(k + 1) % nn] < u[(k - 1) % nn] else -1
        out[f] = [faces[f, (k + step * i) % nn] for i in range(nn)]
    return out


_HEXA_FACES = ((0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7))


def _orient_q1_faces(mesh: Mesh, faces: np.ndarray) -> np.ndarray:
    """Edges of a Quad4 mesh / Quad4 faces of a Hexa8 mesh, node order such that N = (y1-y0, x0-x1) resp. the patch normal
    dr/dxi x dr/deta (femutils/ArcaneFemFunctions.h:1895-1915) points out of the cell the face belongs to."""
    key = lambda a: tuple(sorted(int(x) for x in a))
    owner = {}
    local = ((0, 1), (1, 2), (2, 3), (3, 0)) if mesh.dim == 2 else _HEXA_FACES
    for c, cn in enumerate(mesh.cells):
        for lf in local:
            owner.setdefault(key(cn[list(lf)]), c)
    for f in range(faces.shape[0]):
        c = owner[key(faces[f])]
        centre = mesh.coords[mesh.cells[c]].mean(axis=0)
        p = mesh.coords[faces[f]]
        if mesh.dim == 2:
            n = np.array([p[1][1] - p[0][1], p[0][0] - p[1][0], 0.0])
        else:  # patch normal at the face centre: (dr/dxi) x (dr/deta)
            t1 = 0.25 * (-p[0] + p[1] + p[2] - p[3])
            t2 = 0.25 * (-p[0] - p[1] + p[2] + p[3])
            n = np.cross(t1, t2)
        if np.dot(n, p.mean(axis=0) - centre) < 0.0:
            faces[f] = faces[f, ::-1] if mesh.dim == 2 else faces[f, [0, 3, 2, 1]]
    return faces


# ---------------------------------------------------------------------------
# Synthetic structured boxes.  The device generator
# (`afb_mesh_generate_box`, csrc/mesh_gen.cu) produces bit-identical arrays.
# ---------------------------------------------------------------------------
_KUHN_PERMS = ((0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0))


def jitter_unit(ids: np.ndarray, comp: int, seed: int) -> np.ndarray:
    """hash(id, comp, seed) -> [-0.5, 0.5), 32-bit murmur3 finaliser (same on device)."""
    h = (ids.astype(np.uint64) * np.uint64(3) + np.uint64(comp)) & np.uint64(0xFFFFFFFF)
    h = (h * np.uint64(0x9E3779B1)) & np.uint64(0xFFFFFFFF)
    h ^= np.uint64(seed & 0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return h.astype(np.float64) * (1.0 / 4294967296.0) - 0.5


def box_mesh(dim: int, n: int, jitter: float = 0.2, seed: int = 12345) -> Mesh:
    """[0,1]^dim box, n^dim cubes; 3-D: Kuhn 6-tet split, cell id = 6*cube+perm;
    2-D: two CCW triangles per square along the (0,0)-(1,1) diagonal.
    Node id = i + (n+1)(j + (n+1)k); interior nodes jittered by jitter/n * hash."""
    m = n + 1
    if dim == 3:
        k, j, i = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    else:
        j, i = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")
        k = np.zeros_like(i)
    i, j, k = i.ravel(), j.ravel(), k.ravel()
    ids = (i + m * (j + m * k)).astype(np.int64)
    interior = (i > 0) & (i < n) & (j > 0) & (j < n)
    if dim == 3:
        interior &= (k > 0) & (k < n)
    coords = np.zeros((ids.size, 3), dtype=np.float64)
    for c, idx in enumerate((i, j, k)[:dim]):
        u = jitter_unit(ids, c, seed)
        off = np.where(interior, jitter * u, 0.0)
        coords[:, c] = (idx.astype(np.float64) + off) / float(n)
    if dim == 3:
        ck, cj, ci = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        ci, cj, ck = ci.ravel(), cj.ravel(), ck.ravel()
        v0 = ci + m * (cj + m * ck)
        step = (1, m, m * m)
        cells = np.empty((ci.size, 6, 4), dtype=np.int64)
        for p, perm in enumerate(_KUHN_PERMS):
            a = v0
            cells[:, p, 0] = a
            for s, ax in enumerate(perm):
                a = a + step[ax]
                cells[:, p, s + 1] = a
        cells = cells.reshape(-1, 4)
    else:
        cj, ci = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        ci, cj = ci.ravel(), cj.ravel()
        v00 = ci + m * cj
        cells = np.empty((ci.size, 2, 3), dtype=np.int64)
        cells[:, 0] = np.stack([v00, v00 + 1, v00 + 1 + m], axis=1)
        cells[:, 1] = np.stack([v00, v00 + 1 + m, v00 + m], axis=1)
        cells = cells.reshape(-1, 3)
    groups = {"zmin" if dim == 3 else "ymin": np.nonzero((k if dim == 3 else j) == 0)[0].astype(np.int32)}
    return Mesh(dim=dim, coords=coords, cells=np.ascontiguousarray(cells.astype(np.int32)), node_uid=ids, groups=groups)


def box_mesh_q1(dim: int, n: int, jitter: float = 0.2, seed: int = 12345) -> Mesh:
    """The same jittered node grid as box_mesh, cut in n^dim Q1 cells: Quad4 (counter-clockwise) in 2-D, Hexa8 in 3-D (bottom face
    counter-clockwise, then the top face: the node order of the reference's shape functions, femutils/ShapeFunctions.h:123-129,
    :314-346)."""
    base = box_mesh(dim, n, jitter, seed)
    m = n + 1
    if dim == 2:
        cj, ci = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
        v = (ci + m * cj).ravel()
        cells = np.stack([v, v + 1, v + 1 + m, v + m], axis=1)
    else:
        ck, cj, ci = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
        v = (ci + m * (cj + m * ck)).ravel()
        mm = m * m
        cells = np.stack([v, v + 1, v + 1 + m, v + m, v + mm, v + 1 + mm, v + 1 + m + mm, v + m + mm], axis=1)
    return Mesh(dim=dim, coords=base.coords, cells=np.ascontiguousarray(cells.astype(np.int32)), node_uid=base.node_uid, groups=base.groups)


def to_p2(mesh: Mesh) -> Mesh:
    """P1 simplex mesh -> P2 (Tri6/Tet10): one node per edge, ids appended after the
    vertex ids in ascending (min,max) vertex-pair order; mid-edge coordinates are the
    endpoint averages.  Local node order follows the reference's shape functions
    (femutils/ArcaneFemFunctions.h:3245-3262 Tri6: 3=(0,1) 4=(1,2) 5=(2,0);
    :3893-3911 Tet10: 4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3))."""
    npc = mesh.npc
    pairs = [(0, 1), (1, 2), (2, 0)] if npc == 3 else [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
    c = mesh.cells.astype(np.int64)
    keys = []
    for a, b in pairs:
        lo = np.minimum(c[:, a], c[:, b])
        hi = np.maximum(c[:, a], c[:, b])
        keys.append((lo << 32) | hi)
    keys = np.stack(keys, axis=1)
    uniq, inv = np.unique(keys.ravel(), return_inverse=True)
    mid_ids = (mesh.nb_node + inv.reshape(keys.shape)).astype(np.int32)
    lo = (uniq >> 32).astype(np.int64)
    hi = (uniq & 0xFFFFFFFF).astype(np.int64)
    mid_xyz = 0.5 * (mesh.coords[lo] + mesh.coords[hi])
    coords = np.concatenate([mesh.coords, mid_xyz], axis=0)
    cells = np.concatenate([mesh.cells, mid_ids], axis=1).astype(np.int32)
    uid = np.concatenate([mesh.node_uid, np.arange(mesh.nb_node, coords.shape[0], dtype=np.int64)])
    return Mesh(dim=mesh.dim, coords=np.ascontiguousarray(coords), cells=np.ascontiguousarray(cells), node_uid=uid, groups=dict(mesh.groups))


def subdivide(mesh: Mesh, times: int = 1) -> Mesh:
    """Uniform refinement of a P1 simplex mesh, the stand-in for Arcane's `<subdivider><nb-subdivision>` mesh option
    (e.g. modules/elasticity/inputs/bar.3D.Dirichlet.bodyForce.arc:18-20): every edge gets a midpoint node (ids appended after the
    existing nodes in ascending (min,max) vertex-pair order, as `to_p2`), a triangle becomes 4 triangles, a tetrahedron 8
    (4 corner tetrahedra + the inner octahedron cut along its shortest diagonal), all with the parent's orientation.  Children of cell c are cells [k*c, k*c+k) (k = 4 or 8); boundary faces, node groups and cell groups follow.
    The node and cell numbering is this repo's, not Arcane's (the subdivider is not part of the reference's sources)."""
    for _ in range(times):
        npc = mesh.npc
        if npc not in (3, 4) or npc != mesh.dim + 1:
            raise ValueError("subdivide: P1 simplex cells only")
        p2 = to_p2(mesh)
        c = p2.cells.astype(np.int64)
        if npc == 3:
            # Tri6 local order: 3=(0,1) 4=(1,2) 5=(2,0)
            kids = [(0, 3, 5), (3, 1, 4), (5, 4, 2), (3, 4, 5)]
        else:
            # Tet10 local order: 4=(0,1) 5=(1,2) 6=(0,2) 7=(0,3) 8=(1,3) 9=(2,3); inner octahedron: see below
            kids = [(0, 4, 6, 7), (4, 1, 5, 8), (6, 5, 2, 9), (7, 8, 9, 3),
                    (4, 6, 7, 8), (4, 5, 6, 8), (6, 5, 9, 8), (6, 9, 7, 8)]
        cells = np.stack([c[:, list(k)] for k in kids], axis=1)
        coords = p2.coords
        if npc == 4:
            # the octahedron is cut along its shortest diagonal (ties: 6-8, then 4-9, then 7-5), which keeps the children's shape
            # bounded under repeated refinement (a Kuhn box refined once this way is the Kuhn box of twice the resolution)
            inner = [[(4, 6, 7, 8), (4, 5, 6, 8), (6, 5, 9, 8), (6, 9, 7, 8)],
                     [(4, 9, 5, 6), (4, 9, 6, 7), (4, 9, 7, 8), (4, 9, 8, 5)],
                     [(7, 5, 4, 6), (7, 5, 6, 9), (7, 5, 9, 8), (7, 5, 8, 4)]]
            diag = [(6, 8), (4, 9), (7, 5)]
            length = np.stack([np.linalg.norm(coords[c[:, a]] - coords[c[:, b]], axis=1) for a, b in diag], axis=1)
            best = length.min(axis=1, keepdims=True)
            choice = np.argmax(length <= best * (1.0 + 1e-12), axis=1)  # first diagonal within rounding of the shortest
            for d in (1, 2):
                sel = choice == d
                for q in range(4):
                    cells[sel, 4 + q] = c[sel][:, list(inner[d][q])]
        cells = cells.reshape(-1, npc)
        # keep the parent's orientation (sign of the signed measure) in every child
        def signed(cc):
            x = coords[cc]
            if npc == 3:
                e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
                return e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0]
            return np.einsum("ij,ij->i", np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), x[:, 3] - x[:, 0])
        flip = np.sign(signed(cells)) != np.repeat(np.sign(signed(mesh.cells.astype(np.int64))), len(kids))
        cells[flip, -2], cells[flip, -1] = cells[flip, -1].copy(), cells[flip, -2].copy()
        # mid-edge id of a vertex pair: the sorted unique keys of to_p2
        pairs = [(0, 1), (1, 2), (2, 0)] if npc == 3 else [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
        mc = mesh.cells.astype(np.int64)
        keys = np.unique(np.concatenate([(np.minimum(mc[:, a], mc[:, b]) << 32) | np.maximum(mc[:, a], mc[:, b]) for a, b in pairs]))
        def mid(a, b):
            k = (np.minimum(a, b).astype(np.int64) << 32) | np.maximum(a, b).astype(np.int64)
            pos = np.searchsorted(keys, k)
            if np.any(pos >= keys.size) or np.any(keys[np.minimum(pos, keys.size - 1)] != k):
                raise ValueError("subdivide: a boundary face has an edge that is not an edge of the mesh")
            return mesh.nb_node + pos
        faces = {}
        for name, f in mesh.faces.items():
            f = np.asarray(f, dtype=np.int64)
            if f.ndim != 2 or f.shape[0] == 0:
                faces[name] = f.astype(np.int32)
            elif f.shape[1] == 2:  # edge -> 2 edges, same direction
                m01 = mid(f[:, 0], f[:, 1])
                faces[name] = np.stack([np.stack([f[:, 0], m01], 1), np.stack([m01, f[:, 1]], 1)], 1).reshape(-1, 2).astype(np.int32)
            elif f.shape[1] == 3:  # triangle -> 4 triangles, same normal
                a, b, cc = f[:, 0], f[:, 1], f[:, 2]
                ab, bc, ca = mid(a, b), mid(b, cc), mid(cc, a)
                faces[name] = np.stack([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1), np.stack([ca, bc, cc], 1), np.stack([ab, bc, ca], 1)], 1).reshape(-1, 3).astype(np.int32)
            else:
                faces[name] = f.astype(np.int32)  # points
        groups = {}
        for name, ids in mesh.groups.items():
            if name in faces and faces[name].ndim == 2 and faces[name].shape[1] >= 2:
                groups[name] = np.unique(faces[name]).astype(np.int32)
            else:
                groups[name] = np.asarray(ids, dtype=np.int32)
        k = len(kids)
        cell_groups = {name: (np.asarray(ids, dtype=np.int64)[:, None] * k + np.arange(k)[None, :]).reshape(-1).astype(np.int32) for name, ids in mesh.cell_groups.items()}
        mesh = Mesh(dim=mesh.dim, coords=np.ascontiguousarray(coords), cells=np.ascontiguousarray(cells.astype(np.int32)), node_uid=p2.node_uid,
                    groups=groups, cell_groups=cell_groups, faces=faces)
    return mesh


def box_counts(dim: int, n: int):
    """(nb_cell, nb_node, nb_edge, nnz) of the P1 box."""
    if dim == 3:
        nb_node = (n + 1) ** 3
        nb_edge = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
        return 6 * n ** 3, nb_node, nb_edge, nb_node + 2 * nb_edge
    nb_node = (n + 1) ** 2
    nb_edge = 2 * n * (n + 1) + n * n
    return 2 * n * n, nb_node, nb_edge, nb_node + 2 * nb_edge


# ---------------------------------------------------------------------------
# Domain decomposition.  The reference gets its sub-domains from
# Arcane's partitioner at mesh-read time: every rank holds its own cells plus one
# layer of ghost cells, each node has exactly one owner, and only rows of owned
# nodes are assembled (modules/testlab/CsrGpuBiliAssembly.cc:273,351).  These
# helpers produce sub-domains with the same semantics as plain arrays.
# ---------------------------------------------------------------------------
@dataclass
class Subdomain:
    rank: int
    world: int
    dim: int
    coords: np.ndarray            # float64 [nb_node,3]   local nodes: owned first, then ghosts ordered by (owner, global id)
    cells: np.ndarray             # int32 [nb_cell,npc]   own cells first, then ghost cells (local node ids)
    nb_own_cell: int
    nb_own_node: int
    is_own: np.ndarray            # uint8 [nb_node]
    node_gid: np.ndarray          # int64 [nb_node]  global node id
    node_owner: np.ndarray        # int32 [nb_node]  owning rank
    cell_gid: np.ndarray          # int64 [nb_cell]

    @property
    def nb_node(self):
        return int(self.coords.shape[0])

    @property
    def nb_cell(self):
        return int(self.cells.shape[0])

    @property
    def npc(self):
        return int(self.cells.shape[1])


def slab_layers(n: int, world: int, rank: int):
    """cube layers [k_lo,k_hi) of rank's slab along the last axis (balanced)."""
    base, rem = divmod(n, world)
    k_lo = rank * base + min(rank, rem)
    return k_lo, k_lo + base + (1 if rank < rem else 0)


def box_slab_numbering(dim: int, n: int, k_lo: int, k_hi: int, ghost_cell_layer: bool):
    """Global node ids / owners (relative: -1 lower neighbour, 0 self, +1 upper neighbour) of the local
    nodes of `afb_mesh_generate_box(dim, n, ..., k_lo, k_hi, ghost_cell_layer)`, in local order."""
    m = n + 1
    plane = m * m if dim == 3 else m
    ghost = bool(ghost_cell_layer) and k_hi < n
    own_planes = list(range(k_lo + 1 if k_lo > 0 else 0, k_hi + 1))
    planes = own_planes + ([k_lo] if k_lo > 0 else []) + ([k_hi + 1] if ghost else [])
    rel = [0] * len(own_planes) + ([-1] if k_lo > 0 else []) + ([1] if ghost else [])
    inl = np.arange(plane, dtype=np.int64)
    gid = np.concatenate([k * plane + inl for k in planes])
    owner_rel = np.concatenate([np.full(plane, r, dtype=np.int32) for r in rel])
    nb_own = len(own_planes) * plane
    cubes_per_layer = n * n if dim == 3 else n
    cells_per_cube = 6 if dim == 3 else 2
    nb_own_cell = (k_hi - k_lo) * cubes_per_layer * cells_per_cube
    nb_cell = nb_own_cell + (cubes_per_layer * cells_per_cube if ghost else 0)
    cell_gid = k_lo * cubes_per_layer * cells_per_cube + np.arange(nb_cell, dtype=np.int64)
    return gid, owner_rel, nb_own, nb_own_cell, cell_gid


def box_slab(dim: int, n: int, world: int, rank: int, ghost_cell_layer: bool = True, jitter: float = 0.2, seed: int = 12345) -> Subdomain:
    """Host mirror of the device slab generator (bit-identical arrays): rank's z-slab of the box."""
    full = box_mesh(dim, n, jitter, seed)
    k_lo, k_hi = slab_layers(n, world, rank)
    gid, owner_rel, nb_own, nb_own_cell, cell_gid = box_slab_numbering(dim, n, k_lo, k_hi, ghost_cell_layer)
    g2l = np.full(full.nb_node, -1, dtype=np.int64)
    g2l[gid] = np.arange(gid.size)
    cells = g2l[full.cells[cell_gid]].astype(np.int32)
    assert (cells >= 0).all()
    return Subdomain(rank=rank, world=world, dim=dim, coords=np.ascontiguousarray(full.coords[gid]), cells=np.ascontiguousarray(cells), nb_own_cell=int(nb_own_cell),
                     nb_own_node=int(nb_own), is_own=(owner_rel == 0).astype(np.uint8), node_gid=gid, node_owner=(rank + owner_rel).astype(np.int32), cell_gid=cell_gid)


def partition_mesh(mesh: Mesh, world: int, axis: int | None = None) -> list:
    """Generic slab-like partition of any mesh into `world` sub-domains: cells are sorted by centroid along
    `axis` (default: the last one) and cut in equal parts; a node belongs to the lowest rank among its
    cells; every rank also gets the ghost cells touching its owned nodes (one ghost layer)."""
    axis = mesh.dim - 1 if axis is None else axis
    cent = mesh.coords[mesh.cells][:, :, axis].mean(axis=1)
    order = np.argsort(cent, kind="stable")
    cell_rank = np.empty(mesh.nb_cell, dtype=np.int32)
    bounds = [(mesh.nb_cell * r) // world for r in range(world + 1)]
    for r in range(world):
        cell_rank[order[bounds[r]:bounds[r + 1]]] = r
    node_owner = np.full(mesh.nb_node, world, dtype=np.int32)
    np.minimum.at(node_owner, mesh.cells.ravel(), np.repeat(cell_rank, mesh.npc))
    node_owner[node_owner == world] = 0  # isolated nodes
    subs = []
    for r in range(world):
        own_cells = np.nonzero(cell_rank == r)[0]
        touches_owned = (node_owner[mesh.cells] == r).any(axis=1)
        ghost_cells = np.nonzero(touches_owned & (cell_rank != r))[0]
        cell_gid = np.concatenate([own_cells, ghost_cells]).astype(np.int64)
        used = np.unique(mesh.cells[cell_gid].ravel()) if cell_gid.size else np.empty(0, dtype=np.int64)
        owned = np.nonzero(node_owner == r)[0]
        ghosts = np.setdiff1d(used, owned)
        ghosts = ghosts[np.lexsort((ghosts, node_owner[ghosts]))]
        gid = np.concatenate([owned, ghosts]).astype(np.int64)
        g2l = np.full(mesh.nb_node, -1, dtype=np.int64)
        g2l[gid] = np.arange(gid.size)
        cells = g2l[mesh.cells[cell_gid]].astype(np.int32).reshape(-1, mesh.npc)
        subs.append(Subdomain(rank=r, world=world, dim=mesh.dim, coords=np.ascontiguousarray(mesh.coords[gid]), cells=np.ascontiguousarray(cells),
                              nb_own_cell=int(own_cells.size), nb_own_node=int(owned.size), is_own=(node_owner[gid] == r).astype(np.uint8),
                              node_gid=gid, node_owner=node_owner[gid].astype(np.int32), cell_gid=cell_gid))
    return subs
