"""Mesh ingestion helpers next to the assembly path: the uniform subdivider that stands in for Arcane's
`<subdivider><nb-subdivision>` option.  CPU: structural properties and the oracle on the refined mesh; GPU: the CUDA path on a refined
fixture against the oracle."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from arcanefem_b200 import mesh as M
from oracle import oracle as O
from tests import cases as CS


def _measure(m):
    x = m.coords[m.cells.astype(np.int64)]
    if m.dim == 2:
        e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
        return 0.5 * (e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0])
    return np.einsum("ij,ij->i", np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), x[:, 3] - x[:, 0]) / 6.0


def _face_keys(cells):
    """sorted node tuples of all (dim-1)-faces of simplex cells, one row per (cell, face)"""
    npc = cells.shape[1]
    out = []
    for skip in range(npc):
        out.append(np.sort(np.delete(cells, skip, axis=1), axis=1))
    return np.concatenate(out, axis=0)


MESHES = ["box2", "box3", "L-shape.msh", "sphere_cut.msh", "bar.msh"]


def _mesh(name):
    if name == "box2":
        return M.box_mesh(2, 5)
    if name == "box3":
        return M.box_mesh(3, 3)
    return M.read_msh(os.path.join(CS.GOLDEN, name))


@pytest.mark.parametrize("name", MESHES)
def test_subdivide_structure(name):
    m = _mesh(name)
    r = M.subdivide(m)
    k = 4 if m.dim == 2 else 8
    assert r.nb_cell == k * m.nb_cell and r.dim == m.dim and r.npc == m.npc
    # one new node per edge
    fk = _face_keys(m.cells) if m.dim == 2 else None
    c = m.cells.astype(np.int64)
    pairs = [(a, b) for a in range(m.npc) for b in range(a + 1, m.npc)]
    edges = np.unique(np.concatenate([np.sort(c[:, list(p)], axis=1) for p in pairs]), axis=0)
    assert r.nb_node == m.nb_node + edges.shape[0]
    assert np.array_equal(r.coords[: m.nb_node], m.coords)
    assert np.allclose(r.coords[m.nb_node:], 0.5 * (m.coords[edges[:, 0]] + m.coords[edges[:, 1]]), rtol=0, atol=0)
    # children tile their parent: measures add up per parent, orientation kept, no degenerate child
    vm, vr = _measure(m), _measure(r).reshape(m.nb_cell, k)
    assert np.allclose(vr.sum(axis=1), vm, rtol=1e-13, atol=0)
    assert np.all(np.sign(vr) == np.sign(vm)[:, None])
    assert np.abs(vr).min() >= 0.99 * np.abs(vm).min() / k * (1.0 if m.dim == 2 else 0.5)
    # conforming: every face belongs to one or two cells, and the boundary has exactly k/2 times as many faces
    _, cnt_m = np.unique(_face_keys(m.cells), axis=0, return_counts=True)
    _, cnt_r = np.unique(_face_keys(r.cells), axis=0, return_counts=True)
    assert cnt_r.max() <= 2 and cnt_m.max() <= 2
    assert (cnt_r == 1).sum() == (k // 2) * (cnt_m == 1).sum()
    del fk


@pytest.mark.parametrize("name", ["L-shape.msh", "sphere_cut.msh", "bar.msh"])
def test_subdivide_boundary_groups(name):
    m = _mesh(name)
    r = M.subdivide(m)
    boundary = {tuple(f) for f in np.unique(_face_keys(r.cells), axis=0, return_counts=True)[0][np.unique(_face_keys(r.cells), axis=0, return_counts=True)[1] == 1]}
    for gname, f in m.faces.items():
        f = np.asarray(f)
        if f.ndim != 2 or f.shape[1] != m.dim:
            continue
        rf = r.faces[gname]
        assert rf.shape == (f.shape[0] * (2 if m.dim == 2 else 4), m.dim)
        assert all(tuple(sorted(x)) in boundary for x in rf.tolist())        # still boundary faces of the refined mesh
        assert np.array_equal(r.groups[gname], np.unique(rf))
        assert set(m.groups[gname].tolist()) <= set(r.groups[gname].tolist())
        # same total measure of the group
        def size(mm, ff):
            x = mm.coords[np.asarray(ff, dtype=np.int64)]
            if mm.dim == 2:
                return np.linalg.norm(x[:, 1] - x[:, 0], axis=1).sum()
            return 0.5 * np.linalg.norm(np.cross(x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]), axis=1).sum()
        assert np.isclose(size(m, f), size(r, rf), rtol=1e-12)


def test_subdivide_twice_and_cell_groups():
    m = _mesh("box2")
    m.cell_groups["left"] = np.arange(0, m.nb_cell, 3, dtype=np.int32)
    r2 = M.subdivide(m, 2)
    assert r2.nb_cell == 16 * m.nb_cell
    area_m = np.abs(_measure(m))[m.cell_groups["left"]].sum()
    area_r = np.abs(_measure(r2))[r2.cell_groups["left"]].sum()
    assert np.isclose(area_m, area_r, rtol=1e-12) and r2.cell_groups["left"].size == 16 * m.cell_groups["left"].size
    with pytest.raises(ValueError):
        M.subdivide(M.box_mesh_q1(2, 2))


@pytest.mark.parametrize("dim", [2, 3])
def test_poisson_converges_under_subdivision(dim):
    """-lap u = f with a smooth manufactured solution, Dirichlet on the whole boundary, through the oracle: the nodal error drops by
    2.4-4x per uniform refinement (P1; 4x where the refined box is again a Kuhn box, less where the octahedron cut picks the other of two
    equally short diagonals) -- the refined mesh is a valid finite-element mesh, not just a valid file."""
    def solve(m):
        x = m.coords
        u_ex = np.sin(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * (np.sin(np.pi * x[:, 2]) if dim == 3 else 1.0)
        rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
        vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_COMPACT)
        K = sp.csr_matrix((vals, cols, rows), shape=(m.nb_node, m.nb_node))
        # consistent load of f = dim*pi^2*u through the P1 mass matrix (lumped by rows is enough for the rate)
        vol = np.abs(_measure(m))
        lump = np.zeros(m.nb_node)
        np.add.at(lump, m.cells.astype(np.int64).ravel(), np.repeat(vol / m.npc, m.npc))
        rhs = lump * (dim * np.pi ** 2) * u_ex
        on_b = np.zeros(m.nb_node, dtype=bool)
        on_b[np.any((np.abs(x[:, :dim]) < 1e-12) | (np.abs(x[:, :dim] - 1.0) < 1e-12), axis=1)] = True
        free = ~on_b
        u = np.zeros(m.nb_node)
        u[free] = spla.spsolve(K[free][:, free].tocsc(), rhs[free])
        return np.abs(u - u_ex).max()
    m = M.box_mesh(dim, 4, jitter=0.0)
    e0 = solve(m)
    e1 = solve(M.subdivide(m))
    e2 = solve(M.subdivide(m, 2))
    assert e1 < e0 / 2.5 and e2 < e1 / 2.0, (e0, e1, e2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sphere_cut.msh", "L-shape.msh"])
def test_refined_fixture_through_the_cuda_path(name):
    """pattern bit-exact and values to 1e-12 on a once-refined reference fixture (unstructured, not a box), all three variants"""
    from arcanefem_b200 import capi as A
    m = M.subdivide(_mesh(name))
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    ref = O.assemble(m.dim, m.coords, m.cells, rows_ref, cols_ref, form=O.FORM_COMPACT)
    rowmax = np.repeat(np.maximum.reduceat(np.abs(ref), rows_ref[:-1]), np.diff(rows_ref))
    with A.Context(0) as ctx:
        ctx.set_mesh(m.dim, m.coords, m.cells)
        ctx.build_pattern(1)
        assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref) and np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
        for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
            ctx.reset_values()
            ctx.assemble(A.OP_POISSON, variant=variant)
            got = ctx.to_host(A.ARRAY_VALUES)
            err = np.abs(got - ref) / np.maximum(np.maximum(np.abs(got), np.abs(ref)), rowmax)
            assert err.max() <= 1e-12, (variant, err.max())


# ---- the C ABI's Gmsh reader (afb_msh_*, csrc/mesh_io.cu): host-only, so checked on the CPU -------------------------------------

ALL_MSH = sorted(f for f in os.listdir(CS.GOLDEN) if f.endswith(".msh"))


def _same_mesh(a, b):
    assert a.dim == b.dim
    assert np.array_equal(a.node_uid, b.node_uid)
    assert np.array_equal(a.coords, b.coords)          # bit-exact: the same doubles, reordered the same way
    assert np.array_equal(a.cells, b.cells) and a.cells.dtype == b.cells.dtype
    for attr in ("groups", "faces", "cell_groups"):
        da, db = getattr(a, attr), getattr(b, attr)
        assert sorted(da) == sorted(db), attr
        for k in da:
            assert np.array_equal(da[k], db[k]), (attr, k)


@pytest.mark.parametrize("name", ALL_MSH)
def test_native_msh_reader_matches_on_every_reference_fixture(name):
    """every mesh file of the reference this repo carries (binary msh 4.1; Tri3, Quad4, Tet4, Hexa8; surfaces, volumes, named points):
    nodes, cells and all groups identical to the Python reader the golden-solution tests are pinned with"""
    path = os.path.join(CS.GOLDEN, name)
    _same_mesh(M.read_msh_native(path), M.read_msh(path))


_ASCII_MSH = """$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
3
0 7 "corner"
1 5 "left"
2 9 "plate"
$EndPhysicalNames
$Entities
1 1 1 0
1 0 0 0 1 7
1 0 0 0 0 1 0 1 5 2 1 -1
1 0 0 0 1 1 0 1 9 1 1
$EndEntities
$Comment
a section the path does not use, with a $ in it
$EndComment
$Nodes
2 4 1 40
0 1 0 1
30
0 0 0
2 1 0 3
40
10
20
1 1 0
1 0 0
0 1 0
$EndNodes
$Elements
3 4 1 4
0 1 15 1
1 30
1 1 1 1
2 30 20
2 1 2 2
4 10 20 40
3 30 10 20
$EndElements
"""


def test_native_msh_reader_ascii(tmp_path):
    """the ASCII flavour of msh 4.1, node tags out of order, cells out of tag order, an unknown section in between"""
    p = tmp_path / "tiny.msh"
    p.write_text(_ASCII_MSH)
    m = M.read_msh_native(str(p))
    assert (m.dim, m.nb_node, m.nb_cell, m.npc) == (2, 4, 2, 3)
    assert m.node_uid.tolist() == [10, 20, 30, 40]
    assert m.coords.tolist() == [[1, 0, 0], [0, 1, 0], [0, 0, 0], [1, 1, 0]]
    assert m.cells.tolist() == [[2, 0, 1], [0, 1, 3]]                   # element tags 3, 4 -> local node ids
    assert m.faces["left"].tolist() == [[2, 1]] and m.groups["left"].tolist() == [1, 2]
    assert m.groups["corner"].tolist() == [2]
    assert m.cell_groups["plate"].tolist() == [0, 1]


def test_native_msh_reader_rejects_bad_files(tmp_path):
    from arcanefem_b200 import capi as A
    with pytest.raises(A.AfbError, match="cannot open"):
        M.read_msh_native(str(tmp_path / "missing.msh"))
    data = open(os.path.join(CS.GOLDEN, "L-shape.msh"), "rb").read()
    for cut in (len(data) // 3, len(data) // 2, len(data) - 20):        # truncated inside a binary section / before the end tag
        p = tmp_path / f"cut{cut}.msh"
        p.write_bytes(data[:cut])
        with pytest.raises(A.AfbError, match="afb_msh_read"):
            M.read_msh_native(str(p))
    p = tmp_path / "v2.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(A.AfbError, match="4.1"):
        M.read_msh_native(str(p))
    p = tmp_path / "dangling.msh"
    p.write_text(_ASCII_MSH.replace("4 10 20 40", "4 10 20 99"))
    with pytest.raises(A.AfbError, match="node tag 99"):
        M.read_msh_native(str(p))
    p = tmp_path / "text.msh"
    p.write_text("hello\n")
    with pytest.raises(A.AfbError, match="not a msh file"):
        M.read_msh_native(str(p))
