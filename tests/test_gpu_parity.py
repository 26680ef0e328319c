"""GPU parity tests (run on the B200 box with `-m gpu`): the CUDA path, called through
the C ABI, against the CPU oracle on the same inputs.

Bars: sparsity pattern, DoF numbering and every integer array
bit-exact; fp64 values within 1e-12 relative, measured as
    |a-b| <= 1e-12 * max(|a|, |b|, max|row|)
because entries are sums of <= ~30 mixed-sign terms (exact cancellations exist on
right-angled boxes) and the atomic variant's summation order is arbitrary.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from arcanefem_b200 import capi as A
from arcanefem_b200 import mesh as M
from oracle import oracle as O
from tests import cases as CS

pytestmark = pytest.mark.gpu

TOL = 1.0e-12


@pytest.fixture(scope="module")
def ctx():
    c = A.Context(0)
    yield c
    c.close()


def _fixture_mesh(name):
    return CS.load_mesh(name)


MESHES = {
    "L-shape_2D": lambda: _fixture_mesh("L-shape.msh"),
    "circle_2D": lambda: _fixture_mesh("circle_cut.msh"),
    "porous_2D": lambda: _fixture_mesh("porous-medium.msh"),
    "L-shape_3D": lambda: _fixture_mesh("L-shape-3D.msh"),
    "sphere_3D": lambda: _fixture_mesh("sphere_cut.msh"),
    "bar_3D": lambda: _fixture_mesh("bar_dynamic_3D.msh"),
    "box2d_n17": lambda: M.box_mesh(2, 17),
    "box3d_n9": lambda: M.box_mesh(3, 9),
    "box3d_n6_nojitter": lambda: M.box_mesh(3, 6, jitter=0.0),
    "L-shape_2D_P2": lambda: M.to_p2(_fixture_mesh("L-shape.msh")),
    "sphere_3D_P2": lambda: M.to_p2(_fixture_mesh("sphere_cut.msh")),
    "box3d_n4_P2": lambda: M.to_p2(M.box_mesh(3, 4)),
}
_cache = {}


def get_mesh(name):
    if name not in _cache:
        _cache[name] = MESHES[name]()
    return _cache[name]


def row_scaled_close(got, ref, rows, b=1, layout=O.LAYOUT_PER_BLOCK, tol=TOL):
    """|a-b| <= tol*max(|a|,|b|,rowmax) with rowmax over the block row of the reference."""
    nbr = rows.shape[0] - 1
    bb = b * b
    # every block row occupies the contiguous range [rows[r]*bb, rows[r+1]*bb) in both layouts
    rowmax = np.zeros(nbr)
    seg = np.repeat(np.arange(nbr), np.diff(rows) * bb)
    np.maximum.at(rowmax, seg, np.abs(ref))
    scale = np.maximum(np.maximum(np.abs(got), np.abs(ref)), rowmax[seg])
    err = np.abs(got - ref)
    bad = err > tol * scale
    worst = float(np.max(err / np.where(scale > 0, scale, 1.0))) if err.size else 0.0
    assert not bad.any(), f"{bad.sum()} entries beyond {tol}: worst scaled error {worst}"
    return worst


# ---------------------------------------------------------------------------------------------
# connectivity + pattern: bit-exact
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(MESHES))
def test_pattern_bit_exact(ctx, name):
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    nbr, nnz = ctx.build_pattern(1)
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    assert nbr == m.nb_node and nnz == cols_ref.size
    assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
    assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
    assert np.array_equal(ctx.to_host(A.ARRAY_NZ_PER_ROW), np.diff(rows_ref))
    # node -> cell lists ascending
    ptr, lst = ctx.to_host(A.ARRAY_NODE_CELL_PTR), ctx.to_host(A.ARRAY_NODE_CELL_LIST)
    order = np.argsort(m.cells.ravel(), kind="stable")
    ref_list = (order // m.npc).astype(np.int32)
    ref_ptr = np.concatenate([[0], np.cumsum(np.bincount(m.cells.ravel(), minlength=m.nb_node))]).astype(np.int32)
    assert np.array_equal(ptr, ref_ptr) and np.array_equal(lst, ref_list)
    # COO rows (_translateCSRToCOO)
    assert np.array_equal(ctx.to_host(A.ARRAY_COO_ROWS), O.csr_to_coo_rows(rows_ref))
    # re-build on the same mesh (steady state: fused single-pass kernel with look-back offsets)
    for b in (1, 2):
        assert ctx.build_pattern(b) == (nbr, nnz)
        assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_NZ_PER_ROW), np.diff(rows_ref))
        assert not ctx.to_host(A.ARRAY_VALUES).any()


def test_pattern_with_isolated_node_and_high_valence(ctx):
    # a fan of 150 triangles around node 0 (valence > 128: slow path) + one isolated node
    k = 150
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    coords = np.zeros((k + 2, 3))
    coords[1:k + 1, 0], coords[1:k + 1, 1] = np.cos(ang), np.sin(ang)
    coords[k + 1] = (5.0, 5.0, 0.0)
    cells = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
    ctx.set_mesh(2, coords, cells)
    ctx.build_pattern(1)
    rows_ref, cols_ref = O.build_pattern(3, k + 2, cells)
    assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
    assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_NODEWISE)
    ref = O.assemble(2, coords, cells, rows_ref, cols_ref, form=O.FORM_BSR)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows_ref)
    ctx.build_pattern(1)  # fused re-build: node 0 overflows the private table (warp-cooperative path)
    assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
    assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)


@pytest.mark.parametrize("algo", [A.SPARSITY_AUTO, A.SPARSITY_FROM_CELLS, A.SPARSITY_FROM_CONNECTIVITY])
@pytest.mark.parametrize("name", ["L-shape_2D", "porous_2D", "sphere_3D", "bar_3D", "box3d_n9", "box2d_n17"])
def test_tiled_pattern_rebuild_bit_exact(ctx, name, algo):
    """Steady-state BuildMatrix: once the tile inspector has run (first tiled assembly, or the first build
    with SPARSITY_FROM_CONNECTIVITY), the pattern is re-built from the cells by the per-tile bitmap kernel
    (computeSparsityAtomic) or from the tile-local node-node connectivity (computeSparsityAtomicFree; AUTO
    picks it); rows/columns stay bit-exact, the values of a fresh pattern read as zero, and a tiled
    assembly right after needs no zero fill."""
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.set_sparsity_algorithm(algo)
    try:
        _tiled_pattern_rebuild(ctx, m)
    finally:
        ctx.set_sparsity_algorithm(A.SPARSITY_AUTO)


def _tiled_pattern_rebuild(ctx, m):
    nbr, nnz = ctx.build_pattern(1)
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    v1 = ctx.to_host(A.ARRAY_VALUES)
    ref = O.assemble(m.dim, m.coords, m.cells, rows_ref, cols_ref, form=O.FORM_BSR)
    row_scaled_close(v1, ref, rows_ref)
    for rep in range(2):
        assert ctx.build_pattern(1) == (nbr, nnz)
        assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_NZ_PER_ROW), np.diff(rows_ref))
        if rep == 0:
            assert not ctx.to_host(A.ARRAY_VALUES).any()
        ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
        assert np.array_equal(ctx.to_host(A.ARRAY_VALUES), v1), "tiled gather must be bit-reproducible"
    # another block size on the same mesh re-uses the tiling for the pattern
    assert ctx.build_pattern(m.dim) == (nbr, nnz)
    assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
    assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
    # ... and a vector operator on it (the tiling is re-cut for b > 1, the connectivity with it)
    if m.npc in (3, 4):
        lam, mu = 1.0e6, 8.0e5
        ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER)
        assert ctx.build_pattern(m.dim) == (nbr, nnz)
        assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
        assert np.array_equal(ctx.to_host(A.ARRAY_NZ_PER_ROW), np.diff(rows_ref))


def test_ownership_modes(ctx):
    """Domain-decomposition flags: OWN_CELLS_ONLY drops the ghost cells, ALL_ROWS the isOwn gate.
    Every variant must agree with the atomic cell-wise one (itself checked against the oracle)."""
    m = get_mesh("box3d_n9")
    nb_own_cell = (m.nb_cell * 2) // 3
    own = np.ones(m.nb_node, dtype=np.uint8)
    own[::5] = 0
    ctx.set_mesh(3, m.coords, m.cells, own)
    ctx.set_own_cell_count(nb_own_cell)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    full = O.assemble(3, m.coords, m.cells[:nb_own_cell], rows, cols, form=O.FORM_BSR)
    seg = np.repeat(np.arange(m.nb_node), np.diff(rows))
    for flags, ref in ((A.FLAG_OWN_CELLS_ONLY | A.FLAG_ALL_ROWS, full), (A.FLAG_OWN_CELLS_ONLY, np.where(own[seg] != 0, full, 0.0))):
        for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
            ctx.reset_values()
            ctx.assemble(A.OP_POISSON, variant=variant, flags=flags)
            row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows)


def test_tiled_gather_limits(ctx):
    """Tiles shrink until they fit the shared-memory cache; a single row that cannot fit is an
    explicit error (no silent fallback), the other variants still work."""
    # fan of 300 triangles: node 0 has valence 300 (fits a tile: 384 footprint nodes); 3000 cannot fit
    for k, ok in ((300, True), (3000, False)):
        ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
        coords = np.zeros((k + 1, 3))
        coords[1:, 0], coords[1:, 1] = np.cos(ang), np.sin(ang)
        cells = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
        ctx.set_mesh(2, coords, cells)
        ctx.build_pattern(1)
        rows_ref, cols_ref = O.build_pattern(3, k + 1, cells)
        ref = O.assemble(2, coords, cells, rows_ref, cols_ref, form=O.FORM_BSR)
        if ok:
            ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
            row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows_ref)
        else:
            with pytest.raises(A.AfbError):
                ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
            ctx.assemble(A.OP_POISSON, variant=A.VARIANT_NODEWISE)
            row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows_ref)


@pytest.mark.parametrize("k", [20, 40, 70])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_row_ordered_executor_with_cut_rows(k, layout):
    """rows longer than a unit (32 entries) are cut into units of their own and their diagonal block is summed from a list
    instead of the row's other blocks: an "orange" of k tetrahedra around an axis (3-D, the two axis nodes have k + 2
    entries) and a fan of k triangles (2-D, the centre has k + 1), both vector executors"""
    ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
    lam, mu = O.lame(21.0e5, 0.28)
    c3 = np.zeros((k + 2, 3))
    c3[0, 2], c3[1, 2] = -0.7, 0.9
    c3[2:, 0], c3[2:, 1], c3[2:, 2] = np.cos(ang), 1.1 * np.sin(ang), 0.05 * np.cos(3 * ang)
    t3 = np.array([[0, 1, 2 + i, 2 + (i + 1) % k] for i in range(k)], dtype=np.int32)
    c2 = np.zeros((k + 1, 3))
    c2[1:, 0], c2[1:, 1] = np.cos(ang), 0.8 * np.sin(ang)
    t2 = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
    with A.Context(0) as c:
        for dim, coords, cells in ((3, c3, t3), (2, c2, t2)):
            c.set_mesh(dim, coords, cells)
            c.build_pattern(dim)
            rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
            assert int(np.diff(rows).max()) == k + (2 if dim == 3 else 1)
            ref = O.assemble(dim, coords, cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout, nodewise=True)
            for ex in (A.VEC_EXEC_ROWS, A.VEC_EXEC_UNITS):
                c.set_vector_executor(ex)
                c.reset_values()
                c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
                row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows, b=dim, layout=layout)


def test_device_box_generator_bit_identical(ctx):
    for dim, n in ((3, 7), (2, 13)):
        ref = M.box_mesh(dim, n)
        info = ctx.generate_box(dim, n)
        assert (info["nb_node"], info["nb_cell"]) == (ref.nb_node, ref.nb_cell)
        assert np.array_equal(ctx.to_host(A.ARRAY_COORDS).reshape(-1, 3), ref.coords)
        assert np.array_equal(ctx.to_host(A.ARRAY_CELL_NODES).reshape(-1, dim + 1), ref.cells)
        nbc, nbn, nbe, nnz = M.box_counts(dim, n)
        assert ctx.build_pattern(1) == (nbn, nnz)


# ---------------------------------------------------------------------------------------------
# values: every format x variant x layout against the oracle
# ---------------------------------------------------------------------------------------------
P1_POISSON = ["L-shape_2D", "circle_2D", "porous_2D", "L-shape_3D", "sphere_3D", "box2d_n17", "box3d_n9", "box3d_n6_nojitter"]


@pytest.mark.parametrize("name", P1_POISSON)
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_COO, A.VARIANT_CELLWISE_ATOMIC),
                                         (A.FORMAT_CSR, A.VARIANT_NODEWISE), (A.FORMAT_BSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_NODEWISE),
                                         (A.FORMAT_CSR, A.VARIANT_TILED_GATHER), (A.FORMAT_BSR, A.VARIANT_TILED_GATHER)],
                         ids=["csr-gpu", "coo-gpu", "nwcsr", "bsr", "af-bsr", "tiled-csr", "tiled-bsr"])
def test_poisson_values(ctx, name, fmt, variant):
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    # the reference formulation each back-end uses
    form = {A.FORMAT_CSR: O.FORM_COMPACT, A.FORMAT_COO: O.FORM_COMPACT, A.FORMAT_BSR: O.FORM_BSR}[fmt]
    nodewise = variant in (A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER)
    if nodewise and fmt == A.FORMAT_CSR:
        form = O.FORM_NODEWISE
    flags = A.FLAG_SIGNED_TRI_AREA if fmt != A.FORMAT_BSR else 0
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=variant, flags=flags)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=form, nodewise=nodewise)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows)
    # re-assembly after reset gives the same matrix; without reset it accumulates (adds)
    v1 = ctx.to_host(A.ARRAY_VALUES)
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=variant, flags=flags)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), 2.0 * v1, rows)
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=variant, flags=flags)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), v1, rows)


@pytest.mark.parametrize("name", P1_POISSON)
@pytest.mark.parametrize("fmt", [A.FORMAT_CSR, A.FORMAT_BSR], ids=["tiled-csr", "tiled-bsr"])
def test_poisson_values_with_cell_coefficient_tiled(ctx, name, fmt):
    """per-cell conductivity (afb_set_cell_coefficient: fourier, electrostatics, FourierNL modules) through the tiled executor
    (k_assemble_tiled_coef) against the oracle; a new coefficient on the same plan, then none again"""
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    form = O.FORM_NODEWISE if fmt == A.FORMAT_CSR else O.FORM_BSR
    flags = A.FLAG_SIGNED_TRI_AREA if fmt != A.FORMAT_BSR else 0
    rng = np.random.default_rng(7)
    for trial in range(2):
        cc = 10.0 ** rng.uniform(-2.0, 2.0, m.nb_cell)
        ctx.set_cell_coefficient(cc)
        ctx.reset_values()
        ctx.assemble(A.OP_POISSON, fmt=fmt, variant=A.VARIANT_TILED_GATHER, flags=flags)
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=form, nodewise=True, cell_coef=cc)
        row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows)
    ctx.set_cell_coefficient(None)
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=A.VARIANT_TILED_GATHER, flags=flags)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=form, nodewise=True), rows)
    # the chained-slice executors and the vector executors do not take it: a clear error, no silent fallback
    ctx.set_cell_coefficient(1.5)
    ctx.set_tiled_executor(A.TILED_EXEC_CHAIN)
    with pytest.raises(A.AfbError, match="per-cell coefficient"):
        ctx.assemble(A.OP_POISSON, fmt=fmt, variant=A.VARIANT_TILED_GATHER, flags=flags)
    ctx.set_tiled_executor(A.TILED_EXEC_BRICKS)
    ctx.set_cell_coefficient(None)


def test_full_size_cell_coefficient_tiled_against_cellwise(ctx):
    """C2-size box: the tiled executor with a per-cell coefficient equals the atomic cell-wise variant entry by entry (1e-12 of the row),
    is bit-reproducible, and a uniform coefficient c equals c times the plain matrix"""
    import torch
    n = 120
    ctx.generate_box(3, n)
    nbc, nbn, nbe, nnz = M.box_counts(3, n)
    ctx.build_pattern(1)
    v = ctx.csr_view()
    rows = A.as_torch(v["rows"], nbn + 1, np.int32, 0).long()
    vals = A.as_torch(v["values"], nnz, np.float64, 0)
    cc = 10.0 ** np.random.default_rng(11).uniform(-1.0, 1.0, nbc)
    ctx.set_cell_coefficient(cc)
    out = {}
    for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_TILED_GATHER, A.VARIANT_TILED_GATHER):
        ctx.reset_values()
        ctx.assemble(A.OP_POISSON, variant=variant)
        ctx.synchronize()
        x = vals.clone()
        torch.cuda.synchronize()
        if variant in out:
            assert bool(torch.equal(x, out[variant]))
        out[variant] = x
    rid = torch.repeat_interleave(torch.arange(nbn, device="cuda"), rows[1:] - rows[:-1])
    scale = torch.zeros(nbn, dtype=torch.float64, device="cuda").scatter_reduce_(0, rid, out[A.VARIANT_CELLWISE_ATOMIC].abs(), "amax")
    assert float(((out[A.VARIANT_TILED_GATHER] - out[A.VARIANT_CELLWISE_ATOMIC]).abs() / scale[rid]).max()) < 1e-12
    ctx.set_cell_coefficient(2.5)
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    ctx.synchronize()
    scaled = vals.clone()
    torch.cuda.synchronize()
    ctx.set_cell_coefficient(None)
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    ctx.synchronize()
    assert float(((scaled - 2.5 * vals).abs() / scale[rid]).max()) < 1e-12


def test_signed_area_of_clockwise_triangles(ctx):
    """the testlab compact path gives a negative-definite K_e for a
    clockwise triangle, the BSR path does not."""
    m = get_mesh("L-shape_2D")
    cells = m.cells[:, [0, 2, 1]].copy()
    ctx.set_mesh(2, m.coords, cells)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    ctx.assemble(A.OP_POISSON, fmt=A.FORMAT_CSR, flags=A.FLAG_SIGNED_TRI_AREA)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), O.assemble(2, m.coords, cells, rows, cols, form=O.FORM_COMPACT), rows)
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR)
    ref = O.assemble(2, m.coords, cells, rows, cols, form=O.FORM_BSR)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows)
    assert ref[rows[0]:rows[1]].max() > 0


@pytest.mark.parametrize("name", ["bar_3D", "sphere_3D", "box3d_n9", "L-shape_2D", "box2d_n17"])
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER], ids=["bsr", "af-bsr", "tiled-bsr"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_elasticity_values(ctx, name, variant, layout):
    m = get_mesh(name)
    b = m.dim
    lam, mu = O.lame(21.0e5, 0.28)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(b)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=layout)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout,
                     nodewise=variant != A.VARIANT_CELLWISE_ATOMIC)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows, b=b, layout=layout)
    if variant == A.VARIANT_TILED_GATHER:
        # accumulate on top, then a fresh matrix again
        v1 = ctx.to_host(A.ARRAY_VALUES)
        ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=layout)
        row_scaled_close(ctx.to_host(A.ARRAY_VALUES), 2.0 * v1, rows, b=b, layout=layout)
        ctx.reset_values()
        ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=layout)
        assert np.array_equal(ctx.to_host(A.ARRAY_VALUES), v1), "tiled gather must be bit-reproducible"
    if layout == A.LAYOUT_PER_ROW:
        # BSRMatrix::toCsr hand-off arrays, bit-exact
        crow, ccol, nbc = O.bsr_to_csr(b, rows, cols)
        assert np.array_equal(ctx.to_host(A.ARRAY_CSR_ROWS), crow)
        assert np.array_equal(ctx.to_host(A.ARRAY_CSR_COLUMNS), ccol)
        assert np.array_equal(ctx.to_host(A.ARRAY_CSR_NB_COLUMN), nbc)
        v = ctx.csr_view()
        assert v["nb_row"] == m.nb_node * b and v["nnz"] == cols.size * b * b
    else:
        with pytest.raises(A.AfbError):
            ctx.csr_view()


@pytest.mark.parametrize("name", ["bar_3D", "sphere_3D", "box3d_n9", "L-shape_2D", "box2d_n17"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_elasticity_vector_executors_agree(name, layout):
    """afb_set_vector_executor: the row-ordered executor (default) and the unit executor against the oracle and each other,
    fresh / accumulated / after a reset, also when the contribution lists are read from global memory"""
    m = get_mesh(name)
    b = m.dim
    lam, mu = O.lame(21.0e5, 0.28)
    out = {}
    with A.Context(0) as c:
        c.set_mesh(m.dim, m.coords, m.cells)
        c.build_pattern(b)
        rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout, nodewise=True)
        for ex in (A.VEC_EXEC_ROWS, A.VEC_EXEC_UNITS, A.VEC_EXEC_ROWS):
            c.set_vector_executor(ex)
            for limit in (None, 0):
                if limit is not None:
                    c.set_tiled_stage_limit(limit)
                c.reset_values()
                c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
                v1 = c.to_host(A.ARRAY_VALUES)
                row_scaled_close(v1, ref, rows, b=b, layout=layout)
                c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
                row_scaled_close(c.to_host(A.ARRAY_VALUES), 2.0 * v1, rows, b=b, layout=layout)
                if ex in out:
                    assert np.array_equal(out[ex], v1), "bit-reproducible per executor"
                out[ex] = v1
            c.set_tiled_stage_limit(1 << 30)
        with pytest.raises(A.AfbError):
            c.set_vector_executor(7)
    row_scaled_close(out[A.VEC_EXEC_ROWS], out[A.VEC_EXEC_UNITS], rows, b=b, layout=layout)


@pytest.mark.parametrize("name", ["L-shape_2D", "box2d_n17", "porous_2D"])
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER], ids=["bsr", "af-bsr", "tiled"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_bilaplacian_values(ctx, name, variant, layout):
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(2)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    ctx.assemble(A.OP_BILAPLACIAN, fmt=A.FORMAT_BSR, variant=variant, layout=layout)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_BILAPLACIAN, form=O.FORM_BSR, layout=layout, nodewise=variant == A.VARIANT_NODEWISE)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows, b=2, layout=layout)


@pytest.mark.parametrize("name", ["L-shape_2D_P2", "sphere_3D_P2", "box3d_n4_P2"])
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_COO, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_NODEWISE)],
                         ids=["csr", "coo", "af-bsr"])
def test_p2_poisson_values(ctx, name, fmt, variant):
    m = get_mesh(name)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=variant)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, nodewise=variant == A.VARIANT_NODEWISE)
    row_scaled_close(ctx.to_host(A.ARRAY_VALUES), ref, rows, tol=1e-11)
    # stiffness annihilates constants
    Acsr = sp.csr_matrix((ctx.to_host(A.ARRAY_VALUES), cols, rows))
    assert np.max(np.abs(Acsr @ np.ones(m.nb_node))) < 1e-10 * np.abs(Acsr).max()


def test_is_own_gate(ctx):
    """Rows of non-owned nodes stay zero (modules/testlab/CsrGpuBiliAssembly.cc:351)."""
    m = get_mesh("sphere_3D")
    own = np.ones(m.nb_node, dtype=np.uint8)
    own[::3] = 0
    ctx.set_mesh(m.dim, m.coords, m.cells, is_own=own)
    ctx.build_pattern(1)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
        ctx.reset_values()
        ctx.assemble(A.OP_POISSON, variant=variant)
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_COMPACT, is_own=own)
        got = ctx.to_host(A.ARRAY_VALUES)
        row_scaled_close(got, ref, rows)
        for r in np.nonzero(own == 0)[0]:
            assert not got[rows[r]:rows[r + 1]].any()


# ---------------------------------------------------------------------------------------------
# end-to-end anchor: reference golden solutions through the GPU-assembled system
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CS.POISSON_CASES))
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_COO, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_CSR, A.VARIANT_NODEWISE),
                                         (A.FORMAT_BSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_NODEWISE), (A.FORMAT_CSR, A.VARIANT_TILED_GATHER)],
                         ids=["csr-gpu", "coo-gpu", "nwcsr", "bsr", "af-bsr", "tiled"])
def test_poisson_golden_solution(ctx, name, fmt, variant):
    case = CS.POISSON_CASES[name]
    m = _fixture_mesh(case["mesh"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    ctx.assemble(A.OP_POISSON, fmt=fmt, variant=variant, flags=A.FLAG_SIGNED_TRI_AREA if fmt != A.FORMAT_BSR else 0)
    ctx.set_dirichlet_nodes(ids)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"], nodewise=False, signed_tri_area=True)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    rows, cols, vals, rhs = (ctx.to_host(w) for w in (A.ARRAY_ROWS, A.ARRAY_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    # rhs parity with the oracle (atomic order differs)
    isd = np.zeros(m.nb_node, dtype=np.uint8)
    isd[ids] = 1
    rhs_ref = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=True, is_dirichlet=isd)
    rhs_ref[ids] = case["penalty"] * g
    free = isd == 0
    assert np.array_equal(rhs[ids], rhs_ref[ids])
    assert np.all(np.abs(rhs[free] - rhs_ref[free]) <= 1e-12 * np.abs(rhs_ref[free]).max())
    u = spla.spsolve(sp.csr_matrix((vals, cols, rows)).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16)
    assert worst < 1.0e-7


@pytest.mark.parametrize("name", list(CS.ELASTICITY_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
def test_elasticity_golden_solution(ctx, name, variant):
    case = CS.ELASTICITY_CASES[name]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(b)
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=A.LAYOUT_PER_ROW)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"], nodewise=False)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    crow, ccol, vals, rhs = (ctx.to_host(w) for w in (A.ARRAY_CSR_ROWS, A.ARRAY_CSR_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    u = spla.spsolve(sp.csr_matrix((vals, ccol, crow)).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4


# ---------------------------------------------------------------------------------------------
# Dirichlet: penalty / weak penalty / forced / row / row-column elimination, bit-exact
# against the oracle applied to the SAME (GPU-assembled) values
# ---------------------------------------------------------------------------------------------
def _assembled_elasticity(ctx, layout):
    case = CS.ELASTICITY_CASES["bar_2D"]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(b)
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_NODEWISE, layout=layout)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"], nodewise=True)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    return case, m, b, ids, g


@pytest.mark.parametrize("weak", [False, True], ids=["penalty", "weak-penalty"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_penalty_bit_exact(ctx, weak, layout):
    case, m, b, ids, g = _assembled_elasticity(ctx, layout)
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    vals, rhs = ctx.to_host(A.ARRAY_VALUES).copy(), ctx.to_host(A.ARRAY_RHS).copy()
    ctx.dirichlet_penalty(ids, g + 0.25, 1.0e30, weak=weak)
    for k, d in enumerate(ids):
        s = O.value_index(rows, cols, b, layout, int(d), int(d))
        vals[s] = vals[s] + 1.0e30 if weak else 1.0e30
        rhs[d] = 1.0e30 * (g[k] + 0.25)
    assert np.array_equal(ctx.to_host(A.ARRAY_VALUES), vals)
    assert np.array_equal(ctx.to_host(A.ARRAY_RHS), rhs)


@pytest.mark.parametrize("kind", [A.ELIMINATE_ROW, A.ELIMINATE_ROW_COLUMN], ids=["row", "row-column"])
@pytest.mark.parametrize("quirk", [True, False], ids=["col0-quirk", "no-quirk"])
def test_elimination_bit_exact_and_golden(ctx, kind, quirk):
    case, m, b, ids, g = _assembled_elasticity(ctx, A.LAYOUT_PER_ROW)
    crow, ccol = ctx.to_host(A.ARRAY_CSR_ROWS), ctx.to_host(A.ARRAY_CSR_COLUMNS)
    vals, rhs = ctx.to_host(A.ARRAY_VALUES).copy(), ctx.to_host(A.ARRAY_RHS).copy()
    gg = g + 0.125  # non-zero values exercise the RHS correction
    # forced value on one free DoF as well
    free = int(np.setdiff1d(np.arange(m.nb_node * b), ids)[5])
    ctx.set_elimination(kind, ids, gg)
    ctx.set_forced_values([free], [7.5])
    ctx.apply_matrix_transformation(replicate_column0_quirk=quirk)
    ctx.apply_rhs_transformation()
    info = np.zeros(m.nb_node * b, dtype=np.uint8)
    val = np.zeros(m.nb_node * b)
    info[ids], val[ids] = kind, gg
    finfo = np.zeros(m.nb_node * b, dtype=np.uint8)
    fval = np.zeros(m.nb_node * b)
    finfo[free], fval[free] = 1, 7.5
    O.apply_elimination(crow, ccol, vals, rhs, info, val, forced_info=finfo, forced_value=fval, quirk_skip_col0=quirk)
    assert np.array_equal(ctx.to_host(A.ARRAY_VALUES), vals)
    assert np.array_equal(ctx.to_host(A.ARRAY_RHS), rhs)
    # and the reference's golden field with the case's own values
    ctx.clear_dirichlet()
    ctx.reset_values()
    lam, mu = O.lame(case["E"], case["nu"])
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_CELLWISE_ATOMIC, layout=A.LAYOUT_PER_ROW)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"])
    ctx.set_elimination(kind, ids, g)
    ctx.apply_matrix_transformation(replicate_column0_quirk=quirk)
    ctx.apply_rhs_transformation()
    u = spla.spsolve(sp.csr_matrix((ctx.to_host(A.ARRAY_VALUES), ccol, crow)).tocsc(), ctx.to_host(A.ARRAY_RHS))
    CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)


def test_rhs_neumann_and_traction(ctx):
    """Boundary integrals of the RHS (afb_assemble_rhs_neumann) against the oracle: constant flux, q.n, traction; 2-D and
    3-D; Dirichlet-node skip (testlab) and the isOwn gate.  Sums of <= ~8 fp64 atomics per node: 1e-12 of the largest entry."""
    for name, group in (("circle_2D", "curved"), ("sphere_3D", "curved"), ("bar_3D", "sidesurfaces")):
        m = get_mesh(name)
        faces = M.orient_boundary_faces(m, m.faces[group])
        own = (np.arange(m.nb_node) % 5 != 0).astype(np.uint8)
        isd = np.zeros(m.nb_node, dtype=np.uint8)
        isd[m.groups[group][::3]] = 1
        for b, kind, vals in ((1, A.NEUMANN_FLUX, [1.0e4]), (1, A.NEUMANN_FLUX, [2.9e4, -1.8e4, 0.7e4][:m.dim]),
                              (m.dim, A.NEUMANN_TRACTION, [1.0, -2.5, 0.25][:m.dim]), (m.dim, A.NEUMANN_FLUX, [3.0])):
            for gate in (None, own):
                ctx.set_mesh(m.dim, m.coords, m.cells, gate)
                ctx.build_pattern(b)
                ctx.set_dirichlet_nodes(np.nonzero(isd)[0].astype(np.int32))
                for skip in (False, True):
                    ctx.rhs_reset()
                    ctx.rhs_neumann(faces, vals, kind=kind, skip_dirichlet=skip)
                    ref = np.zeros(m.nb_node * b)
                    O.rhs_neumann(m.dim, b, m.coords, faces, vals, ref, kind=kind, is_own=gate, is_dirichlet=isd if skip else None)
                    got = ctx.to_host(A.ARRAY_RHS)
                    assert np.abs(ref).max() > 0
                    assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref).max()), (name, b, kind, skip)
                ctx.clear_dirichlet()
    # argument checks
    m = get_mesh("circle_2D")
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    with pytest.raises(A.AfbError):
        ctx.rhs_neumann(m.faces["curved"], [1.0, 2.0, 3.0])          # a 2-D flux has 1 or 2 values
    with pytest.raises(A.AfbError):
        ctx.rhs_neumann(m.faces["curved"], [1.0, 2.0], kind=A.NEUMANN_TRACTION)  # b = 1: one traction component


@pytest.mark.parametrize("name", list(CS.NEUMANN_CASES))
def test_poisson_neumann_golden_solution(ctx, name):
    """testlab's flux cases end to end on the GPU-assembled system (modules/testlab/inputs/Test.circle.2D.trac*.arc)."""
    case = CS.NEUMANN_CASES[name]
    m = M.read_msh(os.path.join(CS.GOLDEN, case["mesh"]))
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, flags=A.FLAG_SIGNED_TRI_AREA)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    ctx.set_dirichlet_nodes(ids)
    ctx.rhs_reset()
    ctx.rhs_source([case["f"]], nodewise=False, signed_tri_area=True)
    for group, q in case["neumann"]:
        ctx.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), q, kind=A.NEUMANN_FLUX, skip_dirichlet=True)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    rows, cols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
    u = spla.spsolve(sp.csr_matrix((ctx.to_host(A.ARRAY_VALUES), cols, rows)).tocsc(), ctx.to_host(A.ARRAY_RHS))
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16)
    assert worst < 1.0e-6
    ctx.clear_dirichlet()


def test_elasticity_traction_golden_solution(ctx):
    case = CS.TRACTION_CASE
    m = M.read_msh(os.path.join(CS.GOLDEN, case["mesh"]))
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(b)
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=A.LAYOUT_PER_ROW)
    ctx.rhs_reset()
    for group, t in case["traction"]:
        ctx.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), t, kind=A.NEUMANN_TRACTION)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    crow, ccol, _ = O.bsr_to_csr(b, ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS))
    u = spla.spsolve(sp.csr_matrix((ctx.to_host(A.ARRAY_VALUES), ccol, crow)).tocsc(), ctx.to_host(A.ARRAY_RHS))
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4


# ---------------------------------------------------------------------------------------------
# in-repo Jacobi-PCG (afb_solve_pcg): the reference's golden solution files end to end on the device
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CS.POISSON_CASES) + [f"neumann:{k}" for k in CS.NEUMANN_CASES])
def test_pcg_poisson_golden(ctx, name):
    case = CS.NEUMANN_CASES[name[8:]] if name.startswith("neumann:") else CS.POISSON_CASES[name]
    m = _fixture_mesh(case["mesh"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, flags=A.FLAG_SIGNED_TRI_AREA)
    ctx.set_dirichlet_nodes(ids)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"], nodewise=False, signed_tri_area=True)
    for group, q in case.get("neumann", []):
        ctx.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), q, kind=A.NEUMANN_FLUX, skip_dirichlet=True)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    u, it, res = ctx.solve_pcg(rtol=1e-13, max_iter=5000)
    assert 0 < it < 5000
    rows, cols, vals, rhs = (ctx.to_host(w) for w in (A.ARRAY_ROWS, A.ARRAY_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    u_ref = spla.spsolve(sp.csr_matrix((vals, cols, rows)).tocsc(), rhs)
    assert np.abs(u - u_ref).max() <= 1e-9 * np.abs(u_ref).max()
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16)
    assert worst < 1.0e-6
    ctx.clear_dirichlet()


@pytest.mark.parametrize("name", list(CS.ELASTICITY_CASES))
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_pcg_elasticity_golden(ctx, name, layout):
    case = CS.ELASTICITY_CASES[name]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(b)
    ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
    ctx.rhs_reset()
    ctx.rhs_source(case["f"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    ctx.dirichlet_penalty(ids, g, case["penalty"])
    u, it, res = ctx.solve_pcg(rtol=1e-12, max_iter=20000)
    assert 0 < it < 20000
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4
    # an indefinite / non-assembled system is refused, not iterated on
    ctx.build_pattern(b)
    with pytest.raises(A.AfbError):
        ctx.solve_pcg()


def test_degenerate_meshes(ctx):
    """Ragged / empty inputs: a mesh without cells (every row is its diagonal), a single cell, and a re-used context
    going back to a regular mesh afterwards; every variant, pattern re-builds included."""
    coords = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.3, 0.2, 0.9], [5.0, 5.0, 5.0]])
    for dim, cells in ((2, np.zeros((0, 3), np.int32)), (3, np.zeros((0, 4), np.int32)), (2, np.array([[0, 1, 2]], np.int32)), (3, np.array([[0, 1, 2, 3]], np.int32))):
        ctx.set_mesh(dim, coords, cells)
        rows_ref, cols_ref = O.build_pattern(cells.shape[1], coords.shape[0], cells)
        for rep in range(2):
            nbr, nnz = ctx.build_pattern(1)
            assert (nbr, nnz) == (coords.shape[0], cols_ref.size)
            assert np.array_equal(ctx.to_host(A.ARRAY_ROWS), rows_ref) and np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)
            ref = O.assemble(dim, coords, cells, rows_ref, cols_ref, form=O.FORM_BSR)
            for v in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
                ctx.reset_values()
                ctx.assemble(A.OP_POISSON, variant=v)
                got = ctx.to_host(A.ARRAY_VALUES)
                assert np.all(np.abs(got - ref) <= 1e-12 * max(np.abs(ref).max(), 1.0)), (dim, cells.shape, v)
        ctx.rhs_reset()
        ctx.rhs_source([2.0])
        assert np.all(np.isfinite(ctx.to_host(A.ARRAY_RHS)))
    m = get_mesh("box3d_n6_nojitter")
    ctx.set_mesh(m.dim, m.coords, m.cells)
    ctx.build_pattern(1)
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    assert np.array_equal(ctx.to_host(A.ARRAY_COLUMNS), cols_ref)


def test_rhs_source_variants(ctx):
    for name in ("sphere_3D", "L-shape_2D"):
        m = get_mesh(name)
        for b, f in ((1, [5.5]), (m.dim, [0.5, -1.0, 2.0][:m.dim])):
            ctx.set_mesh(m.dim, m.coords, m.cells)
            ctx.build_pattern(b)
            ctx.rhs_reset()
            ctx.rhs_source(f, nodewise=False, signed_tri_area=(b == 1))
            ref = O.rhs_source_cellwise(m.dim, m.coords, m.cells, f, signed_area=(b == 1))
            got = ctx.to_host(A.ARRAY_RHS)
            assert np.all(np.abs(got - ref) <= 1e-12 * np.abs(ref).max())
            ctx.rhs_source(f, nodewise=True)
            ref = O.rhs_source_nodewise(m.dim, m.coords, m.cells, f)
            got = ctx.to_host(A.ARRAY_RHS)
            assert np.all(np.abs(got - ref) <= 1e-14 * np.abs(ref).max())


# ---------------------------------------------------------------------------------------------
# error behaviour mirrors the reference's exceptions
# ---------------------------------------------------------------------------------------------
def test_errors(ctx):
    m = get_mesh("sphere_3D")
    c2 = A.Context(0)
    with pytest.raises(A.AfbError):
        c2.build_pattern(1)                      # no mesh
    with pytest.raises(A.AfbError):
        c2.set_mesh(3, m.coords, m.cells[:, :3])  # Tri3 cells in 3-D
    c2.set_mesh(m.dim, m.coords, m.cells)
    with pytest.raises(A.AfbError):
        c2.assemble(A.OP_POISSON)                # no pattern
    with pytest.raises(A.AfbError):
        c2.build_pattern(4)                      # block size
    c2.build_pattern(3)
    with pytest.raises(A.AfbError):
        c2.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR)   # b mismatch
    with pytest.raises(A.AfbError):
        c2.assemble(A.OP_ELASTICITY, fmt=A.FORMAT_BSR)  # missing lambda, mu
    with pytest.raises(A.AfbError):
        c2.assemble(A.OP_BILAPLACIAN, fmt=A.FORMAT_BSR)  # Tri3 only
    c2.close()


# ---------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE configs C2 / C3-like)
# ---------------------------------------------------------------------------------------------
def test_full_size_poisson_properties(ctx):
    import torch
    n = 120  # C2: 10 368 000 tets
    info = ctx.generate_box(3, n)
    nbc, nbn, nbe, nnz = M.box_counts(3, n)
    assert (info["nb_cell"], info["nb_node"]) == (nbc, nbn)
    assert ctx.build_pattern(1) == (nbn, nnz)
    v = ctx.csr_view()
    rows = A.as_torch(v["rows"], nbn + 1, np.int32, 0).long()
    cols = A.as_torch(v["columns"], nnz, np.int32, 0).long()
    vals = A.as_torch(v["values"], nnz, np.float64, 0)
    assert bool((rows[1:] > rows[:-1]).all()) and int(rows[-1]) == nnz
    rid = torch.repeat_interleave(torch.arange(nbn, device="cuda"), rows[1:] - rows[:-1])
    # ascending columns inside rows, diagonal present
    same_row = rid[1:] == rid[:-1]
    assert bool((cols[1:][same_row] > cols[:-1][same_row]).all())
    assert int((cols == rid).sum()) == nbn
    results = {}
    for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
        ctx.reset_values()
        ctx.assemble(A.OP_POISSON, variant=variant)
        ctx.synchronize()
        x = vals.clone()
        torch.cuda.synchronize()  # the context owns a non-blocking stream: finish the copy before it works again
        if variant == A.VARIANT_TILED_GATHER:
            # fixed summation order: a second assembly is bit-identical
            ctx.reset_values()
            ctx.assemble(A.OP_POISSON, variant=variant)
            ctx.synchronize()
            assert bool(torch.equal(x, vals))
        results[variant] = x
        rowmax = torch.zeros(nbn, dtype=torch.float64, device="cuda").scatter_reduce(0, rid, x.abs(), "amax")
        rowsum = torch.zeros(nbn, dtype=torch.float64, device="cuda").scatter_add(0, rid, x)
        assert float((rowsum.abs() / rowmax).max()) < 1e-12      # constants are in the kernel
        assert bool((x[cols == rid] > 0).all())                   # positive diagonal
        # symmetry: entry (r,c) equals entry (c,r); key-sort both orientations
        k1 = rid * nbn + cols
        k2 = cols * nbn + rid
        xt = x[torch.argsort(k2)]
        assert bool(torch.equal(torch.sort(k1).values, torch.sort(k2).values))
        assert float(((x - xt).abs() / rowmax[rid]).max()) < 1e-12
    a = results[A.VARIANT_CELLWISE_ATOMIC]
    rowmax = torch.zeros(nbn, dtype=torch.float64, device="cuda").scatter_reduce(0, rid, a.abs(), "amax")
    for other in (A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
        assert float(((a - results[other]).abs() / rowmax[rid]).max()) < 1e-12
    # idempotence of the steady-state BuildMatrix at full size: both re-build algorithms reproduce the first pattern bit for bit
    rows0, cols0 = rows.clone(), cols.clone()
    torch.cuda.synchronize()
    for algo in (A.SPARSITY_FROM_CELLS, A.SPARSITY_FROM_CONNECTIVITY):
        ctx.set_sparsity_algorithm(algo)
        try:
            assert ctx.build_pattern(1) == (nbn, nnz)
            v2 = ctx.csr_view()
            assert bool(torch.equal(A.as_torch(v2["rows"], nbn + 1, np.int32, 0).long(), rows0))
            assert bool(torch.equal(A.as_torch(v2["columns"], nnz, np.int32, 0).long(), cols0))
            ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
            ctx.synchronize()
            assert bool(torch.equal(A.as_torch(v2["values"], nnz, np.float64, 0), results[A.VARIANT_TILED_GATHER]))
        finally:
            ctx.set_sparsity_algorithm(A.SPARSITY_AUTO)


def test_full_size_elasticity_rigid_body_modes(ctx):
    import torch
    n = 48  # 663 552 tets, b=3
    ctx.generate_box(3, n)
    nbc, nbn, nbe, nnz = M.box_counts(3, n)
    lam, mu = O.lame(21.0e5, 0.28)
    ctx.build_pattern(3)
    for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
        ctx.reset_values()
        ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=A.LAYOUT_PER_ROW)
        v = ctx.csr_view()
        crow = A.as_torch(v["rows"], v["nb_row"] + 1, np.int32, 0)
        ccol = A.as_torch(v["columns"], v["nnz"], np.int32, 0)
        vals = A.as_torch(v["values"], v["nnz"], np.float64, 0)
        Acsr = torch.sparse_csr_tensor(crow, ccol, vals, size=(v["nb_row"], v["nb_row"]))
        xyz = A.as_torch(ctx.mesh_info()["xyz"], (nbn, 3), np.float64, 0)
        scale = float(vals.abs().max())
        # translations and infinitesimal rotations are in the kernel of the stiffness matrix
        modes = []
        for k in range(3):
            t = torch.zeros(nbn, 3, dtype=torch.float64, device="cuda")
            t[:, k] = 1.0
            modes.append(t)
        rx = torch.stack([torch.zeros_like(xyz[:, 0]), -xyz[:, 2], xyz[:, 1]], dim=1)
        rz = torch.stack([-xyz[:, 1], xyz[:, 0], torch.zeros_like(xyz[:, 0])], dim=1)
        modes += [rx, rz]
        for mode in modes:
            r = Acsr @ mode.reshape(-1)
            assert float(r.abs().max()) < 1e-11 * scale


# ---------------------------------------------------------------------------------------------
# full-size comparison with the CPU oracle: digests (sum |a_ij|, trace) of the whole matrix and the values of a
# fixed sample of rows, recorded by tests/golden/make_box_checksums.py (the oracle at the BASELINE sizes)
# ---------------------------------------------------------------------------------------------
def _golden_digest(key):
    import json
    with open(os.path.join(CS.GOLDEN, "box_checksums.json")) as f:
        return json.load(f)[key]


FULL_SIZE = [
    ("poisson3d_n24", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER]),
    ("poisson3d_n120", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER]),      # C2
    ("poisson3d_n256", [A.VARIANT_TILED_GATHER]),                                                      # C4
    ("elasticity3d_n24", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER]),
    ("elasticity3d_n100", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_TILED_GATHER]),                        # 6 M cells, b=3
    ("elasticity3d_n203", [A.VARIANT_TILED_GATHER]),                                                   # C3
]


@pytest.mark.parametrize("key,variants", FULL_SIZE, ids=[k for k, _ in FULL_SIZE])
def test_full_size_against_oracle_digest(key, variants):
    import torch
    g = _golden_digest(key)
    n, b, op = g["n"], g["b"], g["op"]
    with A.Context(0) as c:
        info = c.generate_box(3, n)
        assert (info["nb_cell"], info["nb_node"]) == (g["nb_cell"], g["nb_node"])
        nbr, nnz = c.build_pattern(b)
        assert nnz == g["nnz"]
        rows = A.as_torch((c.csr_view() if b == 1 else c.bsr_view())["rows" if b == 1 else "rows_index"], nbr + 1, np.int32, 0)
        rows_h = rows.cpu().numpy()
        sample = sorted(int(k) for k in g["sample_rows"])
        UNITS = 100 + A.VARIANT_TILED_GATHER  # the tiled gather through the other vector executor (afb_set_vector_executor)
        if b > 1 and n <= 100 and A.VARIANT_TILED_GATHER in variants:
            variants = variants + [UNITS]
        for layout in ([A.LAYOUT_PER_BLOCK] if b == 1 else [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW]):
            for variant in variants:
                c.set_vector_executor(A.VEC_EXEC_UNITS if variant == UNITS else A.VEC_EXEC_AUTO)
                variant = A.VARIANT_TILED_GATHER if variant == UNITS else variant
                c.reset_values()
                c.assemble(op, params=g["params"], fmt=A.FORMAT_CSR if b == 1 else A.FORMAT_BSR, variant=variant, layout=layout)
                c.synchronize()
                v = c.csr_view() if b == 1 else c.bsr_view()
                cols = A.as_torch(v["columns"], nnz, np.int32, 0)
                vals = A.as_torch(v["values"], nnz * b * b, np.float64, 0)
                # whole-matrix digests
                import bench
                abs_sum, trace = bench.values_digest(torch, rows, cols, vals, nbr, b, layout)
                assert abs(abs_sum - g["abs_sum"]) <= 1e-12 * g["abs_sum"], (key, variant, layout, abs_sum, g["abs_sum"])
                assert abs(trace - g["trace"]) <= 1e-12 * abs(g["trace"]), (key, variant, layout, trace, g["trace"])
                # sampled rows: columns bit-exact, values to the parity bar
                for r in sample:
                    ref = g["sample_rows"][str(r)]
                    lo, hi = int(rows_h[r]), int(rows_h[r + 1])
                    assert cols[lo:hi].cpu().tolist() == ref["cols"]
                    got = vals[lo * b * b:hi * b * b].cpu().numpy()
                    want = np.asarray(ref["vals"])
                    if b > 1 and layout == A.LAYOUT_PER_ROW:  # per block (p, i, j) -> per row (i, p, j)
                        want = want.reshape(hi - lo, b, b).transpose(1, 0, 2).reshape(-1)
                    scale = np.abs(want).max()
                    assert np.abs(got - want).max() <= TOL * scale, (key, variant, layout, r)


# ---------------------------------------------------------------------------------------------
# the three executors behind VARIANT_TILED_GATHER (afb_set_tiled_executor) and their read-from-global-memory
# branch (afb_set_tiled_stage_limit = 0: no tile stages its plan record in shared memory)
# ---------------------------------------------------------------------------------------------
EXECUTORS = [(A.TILED_EXEC_BRICKS, "bricks"), (A.TILED_EXEC_CHAIN, "chain"), (A.TILED_EXEC_CHAIN_FLOW, "flow")]


@pytest.fixture
def exec_ctx():
    c = A.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", P1_POISSON)
@pytest.mark.parametrize("executor", [e for e, _ in EXECUTORS], ids=[n for _, n in EXECUTORS])
@pytest.mark.parametrize("stage_limit", [None, 0], ids=["staged", "from-global"])
def test_tiled_executors_poisson(exec_ctx, name, executor, stage_limit):
    c = exec_ctx
    m = get_mesh(name)
    c.set_tiled_executor(executor)
    if stage_limit is not None:
        c.set_tiled_stage_limit(stage_limit)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=O.FORM_BSR, nodewise=True)
    c.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER)
    v1 = c.to_host(A.ARRAY_VALUES)
    row_scaled_close(v1, ref, rows)
    # steady state: pattern re-build (connectivity-based from the second build on) + assembly, bit-reproducible
    for _ in range(2):
        c.build_pattern(1)
        c.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER)
        assert np.array_equal(c.to_host(A.ARRAY_VALUES), v1)
    # a second operator on top accumulates
    c.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER)
    row_scaled_close(c.to_host(A.ARRAY_VALUES), 2.0 * v1, rows)


@pytest.mark.parametrize("executor", [e for e, _ in EXECUTORS], ids=[n for _, n in EXECUTORS])
def test_tiled_executors_ownership_modes(exec_ctx, executor):
    c = exec_ctx
    c.set_tiled_executor(executor)
    m = get_mesh("box3d_n9")
    nb_own_cell = (m.nb_cell * 2) // 3
    own = np.ones(m.nb_node, dtype=np.uint8)
    own[::5] = 0
    c.set_mesh(3, m.coords, m.cells, own)
    c.set_own_cell_count(nb_own_cell)
    c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    full = O.assemble(3, m.coords, m.cells[:nb_own_cell], rows, cols, form=O.FORM_BSR)
    seg = np.repeat(np.arange(m.nb_node), np.diff(rows))
    for flags, ref in ((A.FLAG_OWN_CELLS_ONLY | A.FLAG_ALL_ROWS, full), (A.FLAG_OWN_CELLS_ONLY, np.where(own[seg] != 0, full, 0.0)), (0, None)):
        if ref is None:  # ghost cells recomputed, owned rows only (the reference's scheme)
            ref = np.where(own[seg] != 0, O.assemble(3, m.coords, m.cells, rows, cols, form=O.FORM_BSR), 0.0)
        c.reset_values()
        c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, flags=flags)
        row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows)


@pytest.mark.parametrize("executor", [e for e, _ in EXECUTORS], ids=[n for _, n in EXECUTORS])
def test_tiled_executors_limits_and_degenerate(exec_ctx, executor):
    """High-valence fan: 300 triangles around one node fit a tile / slice; 3000 do not: an explicit error, no silent fallback."""
    c = exec_ctx
    c.set_tiled_executor(executor)
    for k, ok in ((300, True), (3000, False)):
        ang = np.linspace(0, 2 * np.pi, k, endpoint=False)
        coords = np.zeros((k + 1, 3))
        coords[1:, 0], coords[1:, 1] = np.cos(ang), np.sin(ang)
        cells = np.array([[0, 1 + i, 1 + (i + 1) % k] for i in range(k)], dtype=np.int32)
        c.set_mesh(2, coords, cells)
        c.build_pattern(1)
        rows_ref, cols_ref = O.build_pattern(3, k + 1, cells)
        ref = O.assemble(2, coords, cells, rows_ref, cols_ref, form=O.FORM_BSR)
        if ok:
            c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
            row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows_ref)
        else:
            with pytest.raises(A.AfbError):
                c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
            c.assemble(A.OP_POISSON, variant=A.VARIANT_NODEWISE)
            row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows_ref)
    # one cell, and a mesh with an isolated node
    coords = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [5, 5, 5]], dtype=np.float64)
    cells = np.array([[0, 1, 2, 3]], dtype=np.int32)
    c.set_mesh(3, coords, cells)
    c.build_pattern(1)
    rows_ref, cols_ref = O.build_pattern(4, 5, cells)
    ref = O.assemble(3, coords, cells, rows_ref, cols_ref, form=O.FORM_BSR)
    c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows_ref)


@pytest.mark.parametrize("name", ["bar_3D", "box3d_n9", "box2d_n17"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_vector_executor_lists_from_global_memory(exec_ctx, name, layout):
    """k_assemble_tiled_vec with every tile's contribution lists read from global memory (the branch oversized tiles take)."""
    c = exec_ctx
    c.set_tiled_stage_limit(0)
    m = get_mesh(name)
    b = m.dim
    lam, mu = O.lame(21.0e5, 0.28)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(b)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout, nodewise=True)
    c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
    row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows, b=b, layout=layout)


@pytest.mark.parametrize("executor", [A.TILED_EXEC_CHAIN, A.TILED_EXEC_CHAIN_FLOW], ids=["chain", "flow"])
def test_chain_executors_full_size_digest(executor):
    """C2 (10.4 M Tet4) through the chained-slice executors against the oracle's full-size digest."""
    import torch
    import bench
    g = _golden_digest("poisson3d_n120")
    with A.Context(0) as c:
        c.set_tiled_executor(executor)
        c.generate_box(3, 120)
        nbr, nnz = c.build_pattern(1)
        for _ in range(2):
            c.build_pattern(1)
            c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
        abs_sum, trace, _ = bench.matrix_digest(torch, A, c, 0, nbr)
        assert abs(abs_sum - g["abs_sum"]) <= 1e-12 * g["abs_sum"] and abs(trace - g["trace"]) <= 1e-12 * g["trace"]


@pytest.mark.parametrize("executor", [e for e, _ in EXECUTORS], ids=[n for _, n in EXECUTORS])
def test_update_coordinates_keeps_plans(exec_ctx, executor):
    """afb_update_coordinates: new coordinates on the same topology (time loop); sparsity structures and executor plans
    survive, every variant follows the new coordinates."""
    c = exec_ctx
    c.set_tiled_executor(executor)
    m = get_mesh("box3d_n9")
    c.set_mesh(3, m.coords, m.cells)
    c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    c.build_pattern(1)  # (the second build of a mesh creates the init-time connectivity of the connectivity-based BuildMatrix)
    c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    t0 = c.inspector_timings()
    rng = np.random.default_rng(7)
    for step in range(2):
        moved = m.coords * (1.0 + 0.1 * (step + 1)) + 0.02 / 9 * rng.standard_normal(m.coords.shape)
        c.update_coordinates(moved)
        ref = O.assemble(3, moved, m.cells, rows, cols, form=O.FORM_BSR)
        for variant in (A.VARIANT_TILED_GATHER, A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE):
            c.build_pattern(1)
            assert np.array_equal(c.to_host(A.ARRAY_ROWS), rows) and np.array_equal(c.to_host(A.ARRAY_COLUMNS), cols)
            c.assemble(A.OP_POISSON, variant=variant)
            row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows)
    assert c.inspector_timings() == t0, "the inspector must not run again"


@pytest.mark.parametrize("executor", [e for e, _ in EXECUTORS], ids=[n for _, n in EXECUTORS])
def test_values_written_before_the_assembly_are_kept(exec_ctx, executor):
    """matrixAddValue / weak penalty before assembleBilinear: every variant adds on top, as the reference's += does."""
    c = exec_ctx
    c.set_tiled_executor(executor)
    m = get_mesh("box3d_n6_nojitter")
    c.set_mesh(3, m.coords, m.cells)
    c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    ref = O.assemble(3, m.coords, m.cells, rows, cols, form=O.FORM_BSR)
    dofs = np.array([0, 5, 11], dtype=np.int32)
    for variant in (A.VARIANT_TILED_GATHER, A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE):
        c.build_pattern(1)
        c.dirichlet_penalty(dofs, np.zeros(3), 3.5, weak=True)  # A[i,i] += 3.5 before anything is assembled
        c.assemble(A.OP_POISSON, variant=variant)
        want = ref.copy()
        for d in dofs:
            lo, hi = rows[d], rows[d + 1]
            want[lo + int(np.nonzero(cols[lo:hi] == d)[0][0])] += 3.5
        row_scaled_close(c.to_host(A.ARRAY_VALUES), want, rows)


def test_unaligned_device_connectivity_is_rejected(exec_ctx):
    import torch
    m = get_mesh("box3d_n6_nojitter")
    xyz = torch.from_numpy(m.coords).cuda()
    flat = torch.zeros(m.cells.size + 1, dtype=torch.int32, device="cuda")
    flat[1:] = torch.from_numpy(m.cells.reshape(-1)).cuda()
    shifted = flat[1:].view(-1, 4)  # 4 bytes off a 16-byte boundary
    with pytest.raises(A.AfbError, match="aligned"):
        exec_ctx.set_mesh(3, xyz, shifted, mem_space=A.MEM_DEVICE)


# ---------------------------------------------------------------------------------------------
# Q1 cells: Quad4 / Hexa8 Poisson (modules/poisson/ElementMatrixHexQuad.h) with their all-pairs sparsity
# (femutils/BSRFormat.cc:284-338: edges + face / body diagonals of every cell)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dim,n", [(2, 13), (3, 6)], ids=["quad4", "hexa8"])
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_COO, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_CELLWISE_ATOMIC),
                                         (A.FORMAT_BSR, A.VARIANT_NODEWISE)], ids=["csr-gpu", "coo-gpu", "bsr", "af-bsr"])
def test_q1_poisson_values(exec_ctx, dim, n, fmt, variant):
    c = exec_ctx
    m = M.box_mesh_q1(dim, n)
    c.set_mesh(m.dim, m.coords, m.cells)
    nbr, nnz = c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    assert np.array_equal(rows, rows_ref) and np.array_equal(cols, cols_ref)
    # interior rows: 9 (quad) / 27 (hexa) entries = the node, its edges and the face / body diagonals of its cells
    assert np.diff(rows).max() == 3 ** dim
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=O.FORM_HOST, nodewise=variant == A.VARIANT_NODEWISE)
    for rep in range(2):  # second round: steady-state BuildMatrix
        c.build_pattern(1)
        assert np.array_equal(c.to_host(A.ARRAY_ROWS), rows) and np.array_equal(c.to_host(A.ARRAY_COLUMNS), cols)
        c.assemble(A.OP_POISSON, fmt=fmt, variant=variant)
        row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows)
    with pytest.raises(A.AfbError, match="P1"):
        c.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    # source term by the 2x2 / 2x2x2 Gauss rule, cell-wise (atomics) and node-wise
    ref_rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, 5.5)
    for nodewise in (False, True):
        c.rhs_reset()
        c.rhs_source(5.5, nodewise=nodewise)
        assert np.all(np.abs(c.to_host(A.ARRAY_RHS) - ref_rhs) <= 1e-12 * np.abs(ref_rhs).max())


@pytest.mark.parametrize("dim,n", [(2, 9), (3, 5)], ids=["quad4", "hexa8"])
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_q1_elasticity_values(exec_ctx, dim, n, variant, layout):
    c = exec_ctx
    m = M.box_mesh_q1(dim, n)
    lam, mu = O.lame(21.0e5, 0.28)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(dim)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout, nodewise=variant == A.VARIANT_NODEWISE)
    c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=layout)
    row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows, b=dim, layout=layout)
    ref_rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, [1.5, -2.0, 0.5][:dim])
    for nodewise in (False, True):
        c.rhs_reset()
        c.rhs_source([1.5, -2.0, 0.5][:dim], nodewise=nodewise)
        assert np.all(np.abs(c.to_host(A.ARRAY_RHS) - ref_rhs) <= 1e-12 * np.abs(ref_rhs).max())
    with pytest.raises(A.AfbError):
        c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)


@pytest.mark.parametrize("name", list(CS.Q1_ELASTICITY_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
def test_q1_elasticity_golden_solution(exec_ctx, name, variant):
    """Quad4 / Hexa8 elasticity against the elasticity module's own golden solution files (modules/elasticity/check/*quad*, *hexa*)"""
    c = exec_ctx
    case = CS.Q1_ELASTICITY_CASES[name]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(b)
    c.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=variant, layout=A.LAYOUT_PER_ROW)
    c.rhs_reset()
    c.rhs_source(case["f"], nodewise=variant == A.VARIANT_NODEWISE)
    for group, t in case.get("traction", []):
        c.rhs_neumann(m.faces[group], t, kind=A.NEUMANN_TRACTION)
    c.dirichlet_penalty(ids, g, case["penalty"])
    crow, ccol, vals, rhs = (c.to_host(w) for w in (A.ARRAY_CSR_ROWS, A.ARRAY_CSR_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    u = spla.spsolve(sp.csr_matrix((vals, ccol, crow)).tocsc(), rhs)
    golden = CS.load_golden(case["golden"], b)
    assert CS.compare_to_golden(m, u, golden, b, eps=1.0e-3, min_value=max(1.0e-10, CS.golden_floor(case, golden)), subset=True) < 1.0e-4


@pytest.mark.parametrize("mesh", ["L-shape_2D", "sphere_3D", "quad4", "hexa8"])
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_COO, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_NODEWISE)],
                         ids=["csr-gpu", "coo-gpu", "af-bsr"])
def test_diffusion_reaction_values(exec_ctx, mesh, fmt, variant):
    """alpha * stiffness + beta * mass (acoustics / heat matrices), uniform and per-cell alpha, against the oracle"""
    c = exec_ctx
    m = M.box_mesh_q1(2, 11) if mesh == "quad4" else M.box_mesh_q1(3, 5) if mesh == "hexa8" else get_mesh(mesh)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    coef = 0.5 + np.arange(m.nb_cell) % 7
    for params, cc in (([-1.0, 1.1], None), ([2.5, 40.0], coef)):
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=params, nodewise=variant == A.VARIANT_NODEWISE, cell_coef=cc)
        c.set_cell_coefficient(cc)
        c.reset_values()
        c.assemble(A.OP_DIFFUSION_REACTION, params=params, fmt=fmt, variant=variant)
        row_scaled_close(c.to_host(A.ARRAY_VALUES), ref, rows)
    c.set_cell_coefficient(None)
    with pytest.raises(A.AfbError):
        c.assemble(A.OP_DIFFUSION_REACTION, params=[1.0, 1.0], fmt=fmt, variant=A.VARIANT_TILED_GATHER)
    with pytest.raises(A.AfbError, match="alpha"):
        c.assemble(A.OP_DIFFUSION_REACTION, fmt=fmt, variant=variant)


@pytest.mark.parametrize("mesh", ["L-shape_2D", "sphere_3D", "box3", "quad4", "hexa8"])
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
@pytest.mark.parametrize("layout", [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW], ids=["per-block", "per-row"])
def test_elastodynamics_values(exec_ctx, mesh, variant, layout):
    """stiffness + mass matrix of the elastodynamics module (c0, c1, c2 of a Newmark step) against the oracle"""
    c = exec_ctx
    m = M.box_mesh(3, 6) if mesh == "box3" else M.box_mesh_q1(2, 9) if mesh == "quad4" else M.box_mesh_q1(3, 4) if mesh == "hexa8" else get_mesh(mesh)
    b = m.dim
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(b)
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    for params in ([625.0, 576.9230769, 384.6153846], [0.0, 1.0e6, 8.0e5], [3.0, 0.0, 0.0]):
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTODYNAMICS, form=O.FORM_BSR, params=params, layout=layout, nodewise=variant == A.VARIANT_NODEWISE)
        c.reset_values()
        c.assemble(A.OP_ELASTODYNAMICS, params=params, fmt=A.FORMAT_BSR, variant=variant, layout=layout)
        got = c.to_host(A.ARRAY_VALUES)
        scale = np.abs(ref).max()
        assert np.abs(got - ref).max() <= 1e-12 * scale
    with pytest.raises(A.AfbError):
        c.assemble(A.OP_ELASTODYNAMICS, params=[1.0, 1.0, 1.0], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
    with pytest.raises(A.AfbError, match="c0"):
        c.assemble(A.OP_ELASTODYNAMICS, params=[1.0, 1.0], fmt=A.FORMAT_BSR, variant=variant, layout=layout)


@pytest.mark.parametrize("name", list(CS.ELASTODYNAMICS_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
def test_elastodynamics_golden_solution(exec_ctx, name, variant):
    """the elastodynamics module's golden displacement files through its Newmark-beta time loop: operator, mass matrix, body force and
    traction assembled on the GPU, penalty Dirichlet rows on the GPU, the per-step solves on the host"""
    c = exec_ctx
    case = CS.ELASTODYNAMICS_CASES[name]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    _, _, c0, c1, c2, _, _ = CS.newmark_coefficients(case)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    c.set_mesh(m.dim, m.coords, m.cells)
    # consistent mass (one DoF per node), then the block system
    c.build_pattern(1)
    c.assemble(A.OP_DIFFUSION_REACTION, params=[0.0, 1.0], fmt=A.FORMAT_BSR, variant=variant)
    mass = sp.csr_matrix((c.to_host(A.ARRAY_VALUES), c.to_host(A.ARRAY_COLUMNS), c.to_host(A.ARRAY_ROWS)))
    c.build_pattern(b)
    damping = None
    if CS.newmark_damping_terms(case, None) is not None:  # Rayleigh damping, generalized-alpha: the elasticity matrices of the right-hand side terms, from the GPU too
        parts = []
        for prm in ([1.0, 0.0], [0.0, 1.0]):
            c.reset_values()
            c.assemble(A.OP_ELASTICITY, params=prm, fmt=A.FORMAT_BSR, variant=variant, layout=A.LAYOUT_PER_ROW)
            parts.append(sp.csr_matrix((c.to_host(A.ARRAY_VALUES).copy(), c.to_host(A.ARRAY_CSR_COLUMNS).copy(), c.to_host(A.ARRAY_CSR_ROWS).copy())))
        c.reset_values()
        damping = CS.newmark_damping_terms(case, lambda lam, mu, x: lam * (parts[0] @ x) + mu * (parts[1] @ x))
    c.assemble(A.OP_ELASTODYNAMICS, params=[c0, c1, c2], fmt=A.FORMAT_BSR, variant=variant, layout=A.LAYOUT_PER_ROW)
    c.rhs_reset()
    c.rhs_source(case["f"], nodewise=variant == A.VARIANT_NODEWISE)
    for group, t in case["traction"]:
        c.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), t, kind=A.NEUMANN_TRACTION)
    static = c.to_host(A.ARRAY_RHS).copy()

    def unit_rhs(group, comp):  # a unit traction on the surface, assembled on the GPU
        c.rhs_reset()
        c.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), [1.0 if i == comp else 0.0 for i in range(b)], kind=A.NEUMANN_TRACTION)
        return c.to_host(A.ARRAY_RHS).copy()

    table_rhs = CS.transient_traction(case, b, unit_rhs)
    c.dirichlet_penalty(ids, g, case["penalty"])
    crow, ccol, vals = (c.to_host(w) for w in (A.ARRAY_CSR_ROWS, A.ARRAY_CSR_COLUMNS, A.ARRAY_VALUES))
    lu = spla.splu(sp.csr_matrix((vals, ccol, crow)).tocsc())

    def solve_step(dynamic, t):
        rhs = static + dynamic + table_rhs(t)
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    u = CS.newmark_time_loop(case, m.nb_node * b, solve_step, lambda x: (mass @ x.reshape(m.nb_node, b)).reshape(-1), damping)
    golden = CS.load_golden(case["golden"], b)
    assert CS.compare_to_golden(m, u, golden, b, eps=1.0e-4, min_value=CS.golden_floor(case, golden), subset=True) < case.get("tol", 1.0e-5)


@pytest.mark.parametrize("name", list(CS.HEAT_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
def test_heat_golden_solution(exec_ctx, name, variant):
    """the heat module's golden temperature files through its implicit Euler time loop: lambda * stiffness + mass / dt
    (OP_DIFFUSION_REACTION) and the mass matrix of the right-hand side assembled on the GPU, penalty rows on the GPU, the per-step
    solves on the host"""
    c = exec_ctx
    case = CS.HEAT_CASES[name]
    m = _fixture_mesh(case["mesh"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    c.assemble(A.OP_DIFFUSION_REACTION, params=[0.0, 1.0], fmt=A.FORMAT_BSR, variant=variant)
    mass = sp.csr_matrix((c.to_host(A.ARRAY_VALUES).copy(), c.to_host(A.ARRAY_COLUMNS).copy(), c.to_host(A.ARRAY_ROWS).copy()))
    c.reset_values()
    c.assemble(A.OP_DIFFUSION_REACTION, params=[case["lam"], 1.0 / case["dt"]], fmt=A.FORMAT_BSR, variant=variant)
    # convection surfaces: their face mass matrices go into the device matrix as the module's matrixAddValue calls do (modules/heat/FemModule.cc:317-331)
    B, static = CS.convection_boundary_terms(m, case)
    if B.nnz:
        import torch
        Bc = B.tocoo()
        slots = torch.empty(Bc.nnz, dtype=torch.int64, device="cuda:0")
        c.lookup_value_slots(Bc.nnz, torch.from_numpy(Bc.row.astype(np.int32)).cuda(), torch.from_numpy(Bc.col.astype(np.int32)).cuda(), slots)
        assert bool((slots >= 0).all())
        c.add_values_at(Bc.nnz, slots, torch.from_numpy(np.ascontiguousarray(Bc.data)).cuda())
        c.synchronize()
    c.rhs_reset()
    for group, q in case.get("neumann", []):
        c.rhs_neumann(m.faces[group], q, kind=A.NEUMANN_FLUX)
    static = static + c.to_host(A.ARRAY_RHS)
    c.dirichlet_penalty(ids, g, case["penalty"])
    rows, cols, vals = (c.to_host(w) for w in (A.ARRAY_ROWS, A.ARRAY_COLUMNS, A.ARRAY_VALUES))
    lu = spla.splu(sp.csr_matrix((vals, cols, rows)).tocsc())

    def solve_step(rhs):
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    T = CS.heat_time_loop(case, m.nb_node, solve_step, lambda x: mass @ x, static)
    assert CS.compare_to_golden(m, T, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True) < case.get("tol", 1.0e-7)


@pytest.mark.parametrize("name", list(CS.FOURIERNL_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER], ids=["bsr", "af-bsr", "tiled"])
def test_fouriernl_golden_solution(exec_ctx, name, variant):
    """the FourierNL module's Picard loop with everything but the update of the conductivity on the device: per iteration the values are
    reset, the Poisson operator is re-assembled with the new per-cell conductivity on the unchanged pattern, the penalty rows are set and
    the system is solved by the PCG stand-in; the converged temperature against the module's golden files"""
    c = exec_ctx
    case = CS.FOURIERNL_CASES[name]
    m = _fixture_mesh(case["mesh"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    launches = []

    def solve(lam):
        n0 = c.launch_count()
        c.reset_values()
        c.set_cell_coefficient(lam)
        c.assemble(A.OP_POISSON, fmt=A.FORMAT_BSR, variant=variant)
        c.rhs_reset()
        c.dirichlet_penalty(ids, g, case["penalty"])
        u, it, _ = c.solve_pcg(rtol=1e-13, max_iter=20000)
        assert 0 < it < 20000
        launches.append(c.launch_count() - n0)
        return u.copy()

    u, iters = CS.picard_loop(case, m, solve)
    c.set_cell_coefficient(None)
    assert 2 < iters < case["max_iters"] and min(launches) > 0
    assert CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True) < 1.0e-4


def _values_per_row(b, rows, values, layout):
    """block values in the order of the expanded scalar CSR (= the per-row layout): [block row][i][entry][j]"""
    if layout == A.LAYOUT_PER_ROW:
        return np.array(values, copy=True)
    out = np.empty_like(values)
    bb = b * b
    for n in range(rows.size - 1):
        lo, cnt = int(rows[n]), int(rows[n + 1] - rows[n])
        blk = np.asarray(values[lo * bb:(lo + cnt) * bb]).reshape(cnt, b, b)          # [entry][i][j]
        out[lo * bb:(lo + cnt) * bb] = blk.transpose(1, 0, 2).reshape(-1)             # [i][entry][j]
    return out


@pytest.mark.parametrize("name", list(CS.SOILDYNAMICS_CASES))
@pytest.mark.parametrize("variant,layout", [(A.VARIANT_CELLWISE_ATOMIC, A.LAYOUT_PER_BLOCK), (A.VARIANT_NODEWISE, A.LAYOUT_PER_ROW)], ids=["bsr", "af-bsr"])
def test_soildynamics_golden_solution(exec_ctx, name, variant, layout):
    """the soildynamics module's golden displacement files (Tri3): Newmark matrix, mass, body force, traction and penalty rows on the GPU;
    the paraxial boundary entries are added to the assembled device matrix the way the module calls BSRMatrix::addValue
    (modules/soildynamics/Paraxial.h:153-186) -- one afb_lookup_value_slots + one afb_add_values_at over the boundary's (row, column) pairs"""
    import torch
    c = exec_ctx
    case = CS.SOILDYNAMICS_CASES[name]
    m = _fixture_mesh(case["mesh"])
    b = m.dim
    k = CS.soildynamics_coefficients(case)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    c.assemble(A.OP_DIFFUSION_REACTION, params=[0.0, 1.0], fmt=A.FORMAT_BSR, variant=variant)
    mass = sp.csr_matrix((c.to_host(A.ARRAY_VALUES).copy(), c.to_host(A.ARRAY_COLUMNS).copy(), c.to_host(A.ARRAY_ROWS).copy()))
    c.build_pattern(b)
    c.assemble(A.OP_ELASTODYNAMICS, params=[k["c0"], k["lam"], k["mu"]], fmt=A.FORMAT_BSR, variant=variant, layout=layout)
    B = sum(CS.paraxial_boundary_matrix(m, m.faces[grp], k["cp"], k["cs"]) for grp in case["paraxial"]).tocoo()
    n = B.nnz
    slots = torch.empty(n, dtype=torch.int64, device="cuda:0")
    c.lookup_value_slots(n, torch.from_numpy(B.row.astype(np.int32)).cuda(), torch.from_numpy(B.col.astype(np.int32)).cuda(), slots)
    assert bool((slots >= 0).all()) and int(torch.unique(slots).numel()) == n
    c.add_values_at(n, slots, torch.from_numpy(np.ascontiguousarray(k["c7"] * B.data)).cuda())
    c.synchronize()
    # what the device matrix holds now, entry by entry, against the oracle's matrix plus the same boundary terms
    rows, cols = c.to_host(A.ARRAY_ROWS), c.to_host(A.ARRAY_COLUMNS)
    crow, ccol = c.to_host(A.ARRAY_CSR_ROWS).copy(), c.to_host(A.ARRAY_CSR_COLUMNS).copy()
    got = _values_per_row(b, rows, c.to_host(A.ARRAY_VALUES), layout)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTODYNAMICS, form=O.FORM_BSR, params=[k["c0"], k["lam"], k["mu"]], layout=O.LAYOUT_PER_ROW,
                     nodewise=variant == A.VARIANT_NODEWISE)
    want = CS.add_in_pattern(crow, ccol, ref, k["c7"] * B)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    c.rhs_reset()
    c.rhs_source(case["f"], nodewise=variant == A.VARIANT_NODEWISE)
    for group, t in case["traction"]:
        c.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), t, kind=A.NEUMANN_TRACTION)
    static = c.to_host(A.ARRAY_RHS).copy()

    def unit_rhs(group, comp):  # a unit traction on the surface, assembled on the GPU
        c.rhs_reset()
        c.rhs_neumann(M.orient_boundary_faces(m, m.faces[group]), [1.0 if i == comp else 0.0 for i in range(b)], kind=A.NEUMANN_TRACTION)
        return c.to_host(A.ARRAY_RHS).copy()

    table_rhs = CS.transient_traction(case, b, unit_rhs)
    c.dirichlet_penalty(ids, g, case["penalty"])
    vals = _values_per_row(b, rows, c.to_host(A.ARRAY_VALUES), layout)
    lu = spla.splu(sp.csr_matrix((vals, ccol, crow)).tocsc())
    Bc = B.tocsr()

    source = CS.double_couple_rhs(case, m)

    def step(U, V, Acc, t):
        rhs = static + table_rhs(t) + (mass @ (k["c0"] * U + k["c3"] * V + k["c4"] * Acc).reshape(m.nb_node, b)).reshape(-1) + Bc @ (k["c7"] * U - k["c8"] * V + k["c9"] * Acc)
        if source is not None:
            source(rhs, t)
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    u = CS.soildynamics_time_loop(case, m.nb_node * b, step)
    golden = CS.load_golden(case["golden"], b)
    assert CS.compare_to_golden(m, u, golden, b, eps=1.0e-4, min_value=CS.golden_floor(case, golden), subset=True) < case.get("tol", 1.0e-5)


@pytest.mark.parametrize("name", list(CS.ACOUSTICS_CASES))
@pytest.mark.parametrize("variant", [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE], ids=["bsr", "af-bsr"])
def test_acoustics_golden_solution(exec_ctx, name, variant):
    """the acoustics module's golden solution files (Helmholtz matrix = OP_DIFFUSION_REACTION, flux on the inner boundary)"""
    c = exec_ctx
    case = CS.ACOUSTICS_CASES[name]
    m = _fixture_mesh(case["mesh"])
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    c.assemble(A.OP_DIFFUSION_REACTION, params=[case["alpha"], case["kc2"]], fmt=A.FORMAT_BSR, variant=variant)
    c.rhs_reset()
    for group, q in case["neumann"]:
        c.rhs_neumann(m.faces[group], q, kind=A.NEUMANN_FLUX)
    rows, cols, vals, rhs = (c.to_host(w) for w in (A.ARRAY_ROWS, A.ARRAY_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    u = spla.spsolve(sp.csr_matrix((vals, cols, rows)).tocsc(), rhs)
    assert CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True) < case.get("tol", 1.0e-7)


@pytest.mark.parametrize("name", list(CS.Q1_CASES))
@pytest.mark.parametrize("fmt,variant", [(A.FORMAT_CSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_CELLWISE_ATOMIC), (A.FORMAT_BSR, A.VARIANT_NODEWISE)],
                         ids=["csr-gpu", "bsr", "af-bsr"])
def test_q1_poisson_golden_solution(exec_ctx, name, fmt, variant):
    """Quad4 / Hexa8 Poisson of the production module against its own golden solution files (modules/poisson/check/*quad*, *hexa*):
    matrix, source term, scalar Neumann flux and penalty through the C ABI, solved by the PCG stand-in and by a direct solve"""
    c = exec_ctx
    case = CS.Q1_CASES[name]
    m = _fixture_mesh(case["mesh"])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    c.set_mesh(m.dim, m.coords, m.cells)
    c.build_pattern(1)
    coef = CS.cell_coefficient(m, case)
    c.set_cell_coefficient(coef)  # (None: off)
    tiled = None
    if coef is not None and m.npc == m.dim + 1 and fmt != A.FORMAT_COO:  # the B200 executor takes the conductivity too
        c.assemble(A.OP_POISSON, fmt=fmt, variant=A.VARIANT_TILED_GATHER)
        tiled = c.to_host(A.ARRAY_VALUES).copy()
        c.reset_values()
    c.assemble(A.OP_POISSON, fmt=fmt, variant=variant)
    if tiled is not None:
        row_scaled_close(tiled, c.to_host(A.ARRAY_VALUES), c.to_host(A.ARRAY_ROWS))
    c.set_cell_coefficient(None)
    c.rhs_reset()
    c.rhs_source(case["f"], nodewise=variant == A.VARIANT_NODEWISE)
    for group, q in case.get("neumann", []):
        c.rhs_neumann(CS.boundary_faces(m, case, group), q, kind=A.NEUMANN_FLUX)
    c.dirichlet_penalty(ids, g, case["penalty"])
    rows, cols, vals, rhs = (c.to_host(w) for w in (A.ARRAY_ROWS, A.ARRAY_COLUMNS, A.ARRAY_VALUES, A.ARRAY_RHS))
    golden = CS.load_golden(case["golden"], 1)
    u = spla.spsolve(sp.csr_matrix((vals, cols, rows)).tocsc(), rhs)
    assert CS.compare_to_golden(m, u, golden, 1, eps=1.0e-4, min_value=1.0e-16, subset=True) < case.get("tol", 1.0e-7)
    x, it, res = c.solve_pcg(rtol=1e-12, max_iter=20000)
    assert CS.compare_to_golden(m, x, golden, 1, eps=1.0e-4, min_value=1.0e-16, subset=True) < max(case.get("tol", 0.0), 1.0e-6)
