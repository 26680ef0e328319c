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
    return Mesh(dim=dim, coords=coords, cells=np.ascontiguousarray(cells), node_uid=uid.astype(np.int64), groups=groups, faces=faces)


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


def box_counts(dim: int, n: int):
    """(nb_cell, nb_node, nb_edge, nnz) of the P1 box."""
    if dim == 3:
        nb_node = (n + 1) ** 3
        nb_edge = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
        return 6 * n ** 3, nb_node, nb_edge, nb_node + 2 * nb_edge
    nb_node = (n + 1) ** 2
    nb_edge = 2 * n * (n + 1) + n * n
    return 2 * n * n, nb_node, nb_edge, nb_node + 2 * nb_edge
