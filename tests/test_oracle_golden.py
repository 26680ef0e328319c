"""Pins the CPU oracle against the reference's own golden solution vectors
(SURVEY.md §8c): every matrix format / formulation of the oracle must reproduce the
golden nodal fields through a host sparse solve at the reference's own tolerance
(testlab 1e-4, modules/testlab/FemModule.cc:2078-2081; elasticity 1e-3,
modules/elasticity/Fem.axl:33).  We observe <= 2e-9 / 2e-5."""
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from arcanefem_b200 import mesh as M
from oracle import oracle as O
from tests import cases as CS


def _csr(rows, cols, vals):
    n = rows.shape[0] - 1
    return sp.csr_matrix((vals, cols, rows), shape=(n, n))


def _load(case):
    return CS.load_mesh(case["mesh"])


@pytest.mark.parametrize("name", list(CS.POISSON_CASES))
def test_mesh_counts_and_nnz(name):
    case = CS.POISSON_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    # nnz = nbNode + 2*nbEdge (modules/testlab/CsrGpuBiliAssembly.cc:193)
    c = m.cells.astype(np.int64)
    pairs = [(i, j) for i in range(m.npc) for j in range(i + 1, m.npc)]
    keys = np.concatenate([(np.minimum(c[:, i], c[:, j]) << 32) | np.maximum(c[:, i], c[:, j]) for i, j in pairs])
    nb_edge = np.unique(keys).size
    assert rows[-1] == m.nb_node + 2 * nb_edge == cols.size
    # ascending, diagonal present
    for r in range(m.nb_node):
        seg = cols[rows[r]:rows[r + 1]]
        assert np.all(np.diff(seg) > 0) and r in seg
    # the host "BuildMatrix" restatement yields the same canonical pattern
    rows2, cols2 = O.build_pattern_host(m.npc, m.nb_node, m.cells, cols.size)
    assert np.array_equal(rows, rows2) and np.array_equal(cols, cols2)


def test_fixture_sizes():
    # SURVEY.md App. D fixture sizes
    expect = {"L-shape.msh": (151, 254), "circle_cut.msh": (101, 166), "porous-medium.msh": (1011, 1753), "bilap.msh": (63, 92),
              "L-shape-3D.msh": (108, 259), "sphere_cut.msh": (194, 527), "bar_dynamic_3D.msh": (64, 120)}
    for f, (nn, nc) in expect.items():
        m = M.read_msh(os.path.join(CS.GOLDEN, f))
        assert (m.nb_node, m.nb_cell) == (nn, nc), f


VARIANTS = [
    ("csr-host", dict(form=O.FORM_HOST, skip_zero=True)),              # csr / coo / legacy back-ends
    ("csr-gpu", dict(form=O.FORM_COMPACT)),                            # csr-gpu / coo-gpu
    ("nwcsr", dict(form=O.FORM_NODEWISE, nodewise=True)),              # nwcsr / blcsr
    ("bsr", dict(form=O.FORM_BSR)),                                    # bsr
    ("af-bsr", dict(form=O.FORM_BSR, nodewise=True)),                  # bsr-atomic-free
]


@pytest.mark.parametrize("name", list(CS.POISSON_CASES))
@pytest.mark.parametrize("variant", VARIANTS, ids=[v[0] for v in VARIANTS])
def test_poisson_golden(name, variant):
    case = CS.POISSON_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, **variant[1])
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    isd = np.zeros(m.nb_node, dtype=np.uint8)
    isd[ids] = 1
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=True, is_dirichlet=isd)
    O.dirichlet_penalty(rows, cols, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(rows, cols, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16)
    assert worst < 1.0e-7


@pytest.mark.parametrize("name", list(CS.Q1_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_q1_poisson_golden(name, nodewise):
    """Quad4 / Hexa8 Poisson of the production module (modules/poisson/ElementMatrixHexQuad.h, source term
    femutils/ArcaneFemFunctions.cc:222-290,437-483) against the module's own golden solution files."""
    case = CS.Q1_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=O.FORM_BSR, nodewise=nodewise, cell_coef=CS.cell_coefficient(m, case))
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"])
    for group, q in case.get("neumann", []):  # edges of the Quad4 mesh / Quad4 faces of the Hexa8 mesh, outward
        O.rhs_neumann(m.dim, 1, m.coords, CS.boundary_faces(m, case, group), q, rhs, kind=O.NEUMANN_FLUX)
    O.dirichlet_penalty(rows, cols, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(rows, cols, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True)
    assert worst < case.get("tol", 1.0e-7)


@pytest.mark.parametrize("name", list(CS.ACOUSTICS_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_acoustics_golden(name, nodewise):
    """alpha * stiffness + beta * mass (OP_DIFFUSION_REACTION) against the acoustics module's golden solution files"""
    case = CS.ACOUSTICS_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=[case["alpha"], case["kc2"]], nodewise=nodewise)
    rhs = np.zeros(m.nb_node)
    for group, q in case["neumann"]:
        O.rhs_neumann(m.dim, 1, m.coords, m.faces[group], q, rhs, kind=O.NEUMANN_FLUX)
    u = spla.spsolve(_csr(rows, cols, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True)
    assert worst < case.get("tol", 1.0e-7)


@pytest.mark.parametrize("name", list(CS.NEUMANN_CASES))
def test_poisson_neumann_golden(name):
    """Constant flux term (modules/testlab/FemModule.cc:1534-1706): scalar value and q.n with the outward normal."""
    case = CS.NEUMANN_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=O.FORM_COMPACT)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    isd = np.zeros(m.nb_node, dtype=np.uint8)
    isd[ids] = 1
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=True, is_dirichlet=isd)
    for group, q in case["neumann"]:
        O.rhs_neumann(m.dim, 1, m.coords, M.orient_boundary_faces(m, m.faces[group]), q, rhs, kind=O.NEUMANN_FLUX, is_dirichlet=isd)
    O.dirichlet_penalty(rows, cols, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(rows, cols, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16)
    assert worst < 1.0e-6


def test_elasticity_traction_golden():
    """Traction term (femutils/ArcaneFemFunctions.h:2854-2885) against modules/elasticity/check/bar.2D.Dirichlet.traction.txt."""
    case = CS.TRACTION_CASE
    m, b, rows, cols, vals, rhs, ids, g = _elasticity_system(case, O.LAYOUT_PER_ROW, False)
    for group, t in case["traction"]:
        O.rhs_neumann(m.dim, b, m.coords, M.orient_boundary_faces(m, m.faces[group]), t, rhs, kind=O.NEUMANN_TRACTION)
    crow, ccol, _ = O.bsr_to_csr(b, rows, cols)
    O.dirichlet_penalty(crow, ccol, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(crow, ccol, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4


@pytest.mark.parametrize("meshfile", ["circle_cut.msh", "sphere_cut.msh", "L-shape-3D.msh", "bar_dynamic_3D.msh"])
def test_neumann_closed_surface_properties(meshfile):
    """Size-independent checks of the flux term in 2-D and 3-D (no reference golden file has a 3-D flux): over the closed
    boundary a constant flux vector integrates to zero (divergence theorem: the normals of orient_boundary_faces point
    outward everywhere), a unit scalar flux integrates to the boundary's measure, and a traction distributes t * measure."""
    m = M.read_msh(os.path.join(CS.GOLDEN, meshfile))
    faces = M.orient_boundary_faces(m, np.concatenate([f for f in m.faces.values()], axis=0))
    # every boundary face exactly once: each appears in exactly one cell
    assert np.unique(np.sort(faces, axis=1), axis=0).shape[0] == faces.shape[0]
    p = m.coords[faces]
    if m.dim == 2:
        meas = np.linalg.norm((p[:, 1] - p[:, 0])[:, :2], axis=1)
    else:
        meas = 0.5 * np.linalg.norm(np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]), axis=1)
    q = [2.9e4, -1.8e4, 0.7e4][:m.dim]
    rhs = np.zeros(m.nb_node)
    O.rhs_neumann(m.dim, 1, m.coords, faces, q, rhs, kind=O.NEUMANN_FLUX)
    assert abs(rhs.sum()) <= 1e-10 * np.abs(q).max() * meas.sum()
    rhs[:] = 0.0
    O.rhs_neumann(m.dim, 1, m.coords, faces, [1.0], rhs, kind=O.NEUMANN_FLUX)
    assert abs(rhs.sum() - meas.sum()) <= 1e-12 * meas.sum()
    t = [1.0, -2.0, 0.5][:m.dim]
    rhs = np.zeros(m.nb_node * m.dim)
    O.rhs_neumann(m.dim, m.dim, m.coords, faces, t, rhs, kind=O.NEUMANN_TRACTION)
    for k in range(m.dim):
        assert abs(rhs[k::m.dim].sum() - t[k] * meas.sum()) <= 1e-12 * abs(t[k]) * meas.sum()


def test_formulations_agree_to_rounding():
    m = _load(CS.POISSON_CASES["sphere_3D"])
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_COMPACT)
    for kw in (dict(form=O.FORM_HOST), dict(form=O.FORM_BSR), dict(form=O.FORM_NODEWISE, nodewise=True), dict(form=O.FORM_BSR, nodewise=True)):
        v = O.assemble(m.dim, m.coords, m.cells, rows, cols, **kw)
        A = _csr(rows, cols, np.abs(ref))
        rowmax = np.repeat(A.max(axis=1).toarray().ravel(), np.diff(rows))
        assert np.all(np.abs(v - ref) <= 1e-12 * np.maximum(np.maximum(np.abs(v), np.abs(ref)), rowmax))


def _elasticity_system(case, layout, nodewise, method="Penalty"):
    m = _load(case)
    b = m.dim
    lam, mu = O.lame(case["E"], case["nu"])
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[lam, mu], layout=layout, nodewise=nodewise)
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=False)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    return m, b, rows, cols, vals, rhs, ids, g


@pytest.mark.parametrize("name", list(CS.ELASTICITY_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_elasticity_golden_per_row_layout(name, nodewise):
    case = CS.ELASTICITY_CASES[name]
    m, b, rows, cols, vals, rhs, ids, g = _elasticity_system(case, O.LAYOUT_PER_ROW, nodewise)
    crow, ccol, nbc = O.bsr_to_csr(b, rows, cols)     # BSRMatrix::toCsr hand-off
    assert np.array_equal(np.diff(crow), nbc)
    O.dirichlet_penalty(crow, ccol, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(crow, ccol, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], b), b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4


@pytest.mark.parametrize("name", list(CS.Q1_ELASTICITY_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_q1_elasticity_golden(name, nodewise):
    """Quad4 / Hexa8 elasticity (modules/elasticity/ElementMatrixHexQuad.h; body force modules/elasticity/BodyForce.h, Gauss rule)
    against the elasticity module's own golden solution files"""
    case = CS.Q1_ELASTICITY_CASES[name]
    m, b, rows, cols, vals, rhs, ids, g = _elasticity_system(case, O.LAYOUT_PER_ROW, nodewise)
    for group, t in case.get("traction", []):  # edges of the Quad4 mesh / Quad4 faces of the Hexa8 mesh (no normal involved)
        O.rhs_neumann(m.dim, b, m.coords, m.faces[group], t, rhs, kind=O.NEUMANN_TRACTION)
    crow, ccol, nbc = O.bsr_to_csr(b, rows, cols)
    O.dirichlet_penalty(crow, ccol, vals, rhs, ids, g, case["penalty"])
    u = spla.spsolve(_csr(crow, ccol, vals).tocsc(), rhs)
    golden = CS.load_golden(case["golden"], b)
    worst = CS.compare_to_golden(m, u, golden, b, eps=1.0e-3, min_value=max(1.0e-10, CS.golden_floor(case, golden)), subset=True)
    assert worst < 1.0e-4


@pytest.mark.parametrize("name", list(CS.ELASTODYNAMICS_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_elastodynamics_golden(name, nodewise):
    """Stiffness + mass operator of the elastodynamics module (modules/elastodynamics/ElementMatrix.h) through the module's Newmark-beta
    time loop, against its own golden displacement files (epsilon 1e-4, FemModule.cc:536-540)."""
    case = CS.ELASTODYNAMICS_CASES[name]
    m = _load(case)
    b = m.dim
    _, _, c0, c1, c2, _, _ = CS.newmark_coefficients(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTODYNAMICS, form=O.FORM_BSR, params=[c0, c1, c2], layout=O.LAYOUT_PER_ROW, nodewise=nodewise)
    mass = _csr(rows, cols, O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=[0.0, 1.0]))
    static = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=False)
    for group, t in case["traction"]:
        O.rhs_neumann(m.dim, b, m.coords, M.orient_boundary_faces(m, m.faces[group]), t, static, kind=O.NEUMANN_TRACTION)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    crow, ccol, _ = O.bsr_to_csr(b, rows, cols)
    lhs = vals.copy()
    O.dirichlet_penalty(crow, ccol, lhs, np.zeros(m.nb_node * b), ids, g, case["penalty"])
    lu = spla.splu(_csr(crow, ccol, lhs).tocsc())

    def mass_times(x):
        return (mass @ x.reshape(m.nb_node, b)).reshape(-1)

    def unit_rhs(group, comp):
        out = np.zeros(m.nb_node * b)
        O.rhs_neumann(m.dim, b, m.coords, M.orient_boundary_faces(m, m.faces[group]), [1.0 if i == comp else 0.0 for i in range(b)], out, kind=O.NEUMANN_TRACTION)
        return out

    table_rhs = CS.transient_traction(case, b, unit_rhs)

    def solve_step(dynamic, t):
        rhs = static + dynamic + table_rhs(t)
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    k_lam = _csr(crow, ccol, O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[1.0, 0.0], layout=O.LAYOUT_PER_ROW, nodewise=nodewise))
    k_mu = _csr(crow, ccol, O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[0.0, 1.0], layout=O.LAYOUT_PER_ROW, nodewise=nodewise))
    damping = CS.newmark_damping_terms(case, lambda lam, mu, x: lam * (k_lam @ x) + mu * (k_mu @ x))
    u = CS.newmark_time_loop(case, m.nb_node * b, solve_step, mass_times, damping)
    golden = CS.load_golden(case["golden"], b)
    worst = CS.compare_to_golden(m, u, golden, b, eps=1.0e-4, min_value=CS.golden_floor(case, golden), subset=True)
    assert worst < case.get("tol", 1.0e-5)
    # composition: the operator is the elasticity matrix with (lambda, mu) = (c1, c2) plus c0 times the mass on every component
    ke = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=[c1, c2], layout=O.LAYOUT_PER_ROW, nodewise=nodewise)
    A = _csr(crow, ccol, ke) + c0 * sp.kron(mass, sp.identity(b), format="csr")
    d = abs(A - _csr(crow, ccol, vals))
    assert d.max() <= 1e-12 * abs(A).max()


@pytest.mark.parametrize("name", list(CS.HEAT_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_heat_golden(name, nodewise):
    """lambda * stiffness + mass / dt (OP_DIFFUSION_REACTION) through the heat module's implicit Euler time loop, against its own golden
    temperature files (modules/heat/check/2d_conduction.txt, 2d_conduction.quad.txt, 3d_conduction.txt)"""
    case = CS.HEAT_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=[case["lam"], 1.0 / case["dt"]], nodewise=nodewise)
    mass = _csr(rows, cols, O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=[0.0, 1.0]))
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)
    B, static = CS.convection_boundary_terms(m, case)   # (empty without convection surfaces)
    for group, q in case.get("neumann", []):
        O.rhs_neumann(m.dim, 1, m.coords, m.faces[group], q, static, kind=O.NEUMANN_FLUX)
    lhs = CS.add_in_pattern(rows, cols, vals, B)
    O.dirichlet_penalty(rows, cols, lhs, np.zeros(m.nb_node), ids, g, case["penalty"])
    lu = spla.splu(_csr(rows, cols, lhs).tocsc())

    def solve_step(rhs):
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    T = CS.heat_time_loop(case, m.nb_node, solve_step, lambda x: mass @ x, static)
    worst = CS.compare_to_golden(m, T, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True)
    assert worst < case.get("tol", 1.0e-7)


@pytest.mark.parametrize("name", list(CS.FOURIERNL_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_fouriernl_golden(name, nodewise):
    """Poisson operator with a per-cell conductivity re-assembled in every Picard iteration of the FourierNL module, against its golden files"""
    case = CS.FOURIERNL_CASES[name]
    m = _load(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], 1)

    def solve(lam):
        vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_POISSON, form=O.FORM_BSR, nodewise=nodewise, cell_coef=lam)
        rhs = np.zeros(m.nb_node)
        O.dirichlet_penalty(rows, cols, vals, rhs, ids, g, case["penalty"])
        return spla.spsolve(_csr(rows, cols, vals).tocsc(), rhs)

    u, iters = CS.picard_loop(case, m, solve)
    assert 2 < iters < case["max_iters"]
    worst = CS.compare_to_golden(m, u, CS.load_golden(case["golden"], 1), 1, eps=1.0e-4, min_value=1.0e-16, subset=True)
    assert worst < 1.0e-4   # (the golden file is a converged Picard iterate at nlin-rtol 1e-5, not an exact solve)


@pytest.mark.parametrize("name", list(CS.SOILDYNAMICS_CASES))
@pytest.mark.parametrize("nodewise", [False, True], ids=["bsr", "af-bsr"])
def test_soildynamics_golden(name, nodewise):
    """The soildynamics module on Tri3: elastodynamics operator + paraxial boundary entries added to the assembled matrix
    (BSRMatrix::addValue in modules/soildynamics/Paraxial.h:153-186), Newmark-beta loop, against the module's golden displacement files."""
    case = CS.SOILDYNAMICS_CASES[name]
    m = _load(case)
    b = m.dim
    k = CS.soildynamics_coefficients(case)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTODYNAMICS, form=O.FORM_BSR, params=[k["c0"], k["lam"], k["mu"]], layout=O.LAYOUT_PER_ROW, nodewise=nodewise)
    mass = _csr(rows, cols, O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_DIFFUSION_REACTION, form=O.FORM_BSR, params=[0.0, 1.0]))
    static = O.rhs_source_cellwise(m.dim, m.coords, m.cells, case["f"], signed_area=False)
    for group, t in case["traction"]:
        O.rhs_neumann(m.dim, b, m.coords, M.orient_boundary_faces(m, m.faces[group]), t, static, kind=O.NEUMANN_TRACTION)
    B = sum(CS.paraxial_boundary_matrix(m, m.faces[g], k["cp"], k["cs"]) for g in case["paraxial"])
    crow, ccol, _ = O.bsr_to_csr(b, rows, cols)
    lv = CS.add_in_pattern(crow, ccol, vals, k["c7"] * B)  # (the boundary entries fall inside the cell pattern)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    O.dirichlet_penalty(crow, ccol, lv, np.zeros(m.nb_node * b), ids, g, case["penalty"])
    lu = spla.splu(_csr(crow, ccol, lv).tocsc())

    source = CS.double_couple_rhs(case, m)

    def unit_rhs(group, comp):
        out = np.zeros(m.nb_node * b)
        O.rhs_neumann(m.dim, b, m.coords, M.orient_boundary_faces(m, m.faces[group]), [1.0 if i == comp else 0.0 for i in range(b)], out, kind=O.NEUMANN_TRACTION)
        return out

    table_rhs = CS.transient_traction(case, b, unit_rhs)

    def step(U, V, A, t):
        rhs = static + table_rhs(t) + (mass @ (k["c0"] * U + k["c3"] * V + k["c4"] * A).reshape(m.nb_node, b)).reshape(-1) + B @ (k["c7"] * U - k["c8"] * V + k["c9"] * A)
        if source is not None:
            source(rhs, t)
        rhs[ids] = case["penalty"] * np.asarray(g)
        return lu.solve(rhs)

    u = CS.soildynamics_time_loop(case, m.nb_node * b, step)
    golden = CS.load_golden(case["golden"], b)
    worst = CS.compare_to_golden(m, u, golden, b, eps=1.0e-4, min_value=CS.golden_floor(case, golden), subset=True)
    assert worst < case.get("tol", 1.0e-5)


def test_elasticity_per_block_layout_equals_per_row():
    case = CS.ELASTICITY_CASES["bar_3D"]
    m, b, rows, cols, v_row, *_ = _elasticity_system(case, O.LAYOUT_PER_ROW, False)
    _, _, _, _, v_blk, *_ = _elasticity_system(case, O.LAYOUT_PER_BLOCK, False)
    for r in range(0, m.nb_node * b, 7):
        for p in range(rows[r // b], rows[r // b + 1]):
            for k in range(b):
                c = cols[p] * b + k
                assert v_row[O.value_index(rows, cols, b, O.LAYOUT_PER_ROW, r, c)] == v_blk[O.value_index(rows, cols, b, O.LAYOUT_PER_BLOCK, r, c)]


@pytest.mark.parametrize("method", ["RowElimination", "RowColumnElimination"])
def test_elasticity_elimination_golden(method):
    """modules/elasticity/CMakeLists.txt:93-106: same golden file through Row / RowColumn elimination."""
    case = CS.ELASTICITY_CASES["bar_2D"]
    m, b, rows, cols, vals, rhs, ids, g = _elasticity_system(case, O.LAYOUT_PER_ROW, False)
    crow, ccol, _ = O.bsr_to_csr(b, rows, cols)
    info = np.zeros(m.nb_node * b, dtype=np.uint8)
    val = np.zeros(m.nb_node * b)
    info[ids] = 1 if method == "RowElimination" else 2
    val[ids] = g
    O.apply_elimination(crow, ccol, vals, rhs, info, val)
    A = _csr(crow, ccol, vals)
    if method == "RowColumnElimination":
        # column 0 quirk aside, the eliminated system is symmetric
        d = (A - A.T).tocoo()
        bad = [(i, j) for i, j, v in zip(d.row, d.col, d.data) if abs(v) > 1e-9 and 0 not in (i, j)]
        assert not bad
    u = spla.spsolve(A.tocsc(), rhs)
    golden = CS.load_golden(case["golden"], b)
    # eliminated DoFs are exact zeros here, golden holds ~1e-29 penalty residue: skipped by min_value
    CS.compare_to_golden(m, u, golden, b, eps=1.0e-3, min_value=1.0e-10)


@pytest.mark.parametrize("method", ["RowElimination", "RowColumnElimination"])
def test_bilaplacian_golden(method):
    """modules/bilaplacian/inputs/direct.arc + internal_hypre_rowColElim.arc -> check/2d_test.txt.
    The saddle-point system [0 S; S M] with penalty 1e30 is numerically singular for
    LAPACK/SuperLU, so the golden field is anchored through the reference's elimination
    methods (exact Dirichlet), which reproduce it to ~1e-6, and through its residual."""
    case = CS.BILAPLACIAN_CASE
    m = _load(case)
    b = 2
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_BILAPLACIAN, form=O.FORM_BSR, layout=O.LAYOUT_PER_ROW)
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, [case["f"], 0.0], signed_area=False)
    crow, ccol, _ = O.bsr_to_csr(b, rows, cols)
    ids, g = CS.dirichlet_dofs(m, case["dirichlet"], b)
    golden = CS.load_golden(case["golden"], b)
    ug = np.array([golden[int(m.node_uid[i // b])][i % b] for i in range(m.nb_node * b)])
    r = O.spmv(crow, ccol, vals, ug) - rhs
    free = np.ones(m.nb_node * b, dtype=bool)
    free[ids] = False
    assert np.max(np.abs(r[free])) < 2e-3 * np.max(np.abs(rhs))
    info = np.zeros(m.nb_node * b, dtype=np.uint8)
    val = np.zeros(m.nb_node * b)
    info[ids] = 1 if method == "RowElimination" else 2
    val[ids] = g
    # DoF 0 (node uid 1, component u1) is a Dirichlet DoF and the (u1,u1) block is zero:
    # with the reference's `column_index > 0` quirk (CsrDoFLinearSystemImpl.cc:111) entry
    # (0,0) would stay 0 and the matrix be singular, so the quirk is switched off here.
    O.apply_elimination(crow, ccol, vals, rhs, info, val, quirk_skip_col0=False)
    u = spla.spsolve(_csr(crow, ccol, vals).tocsc(), rhs)
    worst = CS.compare_to_golden(m, u, golden, b, eps=1.0e-3, min_value=1.0e-10)
    assert worst < 1.0e-4


def test_reference_rank_legs_match_global_oracle():
    """bench.py's CPU arm: the per-rank sequential CSR back-end (init-time node-node connectivity, BuildMatrix =
    allocate + fill + append walk, AddAndCompute = host element matrix + linear-scan add) reproduces the global
    oracle matrix on its owned rows; the first-assembly form (connectivity built inside BuildMatrix) is identical."""
    n = 6
    m = M.box_mesh(3, n)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    ref = O.assemble(3, m.coords, m.cells, rows, cols, form=O.FORM_HOST)
    plane, layer = (n + 1) ** 2, 6 * n * n
    for k0, k1 in ((0, 3), (3, 7)):
        part = (layer * max(k0 - 1, 0), layer * min(k1, n), plane * k0, plane * k1)
        init = O.ReferenceRank(m.cells, *part)
        cap = int(rows[part[3]] - rows[part[2]])
        for ini in (init, None):
            r = O.reference_rank(3, m.coords, m.cells, *part, want_arrays=True, capacity=cap, init=ini)
            assert r["nnz"] == cap
            for lr in range(part[3] - part[2]):
                g = part[2] + lr
                lo, hi = r["rows"][lr], r["rows"][lr + 1]
                c = r["cols"][lo:hi]
                assert c[0] == g, "diagonal first (CsrBiliAssembly.cc:84)"
                order = np.argsort(c)
                assert np.array_equal(c[order], cols[rows[g]:rows[g + 1]])
                want = ref[rows[g]:rows[g + 1]]
                assert np.allclose(r["vals"][lo:hi][order], want, rtol=0, atol=1e-12 * np.abs(want).max())
        init.close()
