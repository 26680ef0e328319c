"""The reference's own end-to-end test cases for the assembly path,
restated as data: mesh fixture, source term, Dirichlet groups in .arc order, penalty,
golden nodal field.  Citations: modules/testlab/inputs/Test.*.arc,
modules/elasticity/inputs/bar.*.arc, modules/bilaplacian/inputs/direct.arc."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

POISSON_CASES = {
    # modules/testlab/inputs/Test.L-shape.2D.csr-gpu.arc
    "L-shape_2D": dict(mesh="L-shape.msh", f=-5.5, dirichlet=[("boundary", 0.5)], penalty=1.0e30,
                       golden="poisson_test_ref_L-shape_2D.txt"),
    # modules/testlab/inputs/Test.L-shape.3D.nwcsr.arc
    "L-shape_3D": dict(mesh="L-shape-3D.msh", f=5.5, dirichlet=[("bot", 50.0), ("bc", 10.0)], penalty=1.0e30,
                       golden="poisson_test_ref_L-shape_3D.txt"),
    # modules/testlab/inputs/Test.circle.2D.csr.arc
    "circle_2D": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e30,
                      golden="poisson_test_ref_circle_2D.txt"),
    # modules/testlab/inputs/Test.sphere.3D.csr-gpu.arc
    "sphere_3D": dict(mesh="sphere_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e31,
                      golden="poisson_test_ref_sphere_3D.txt"),
}

# Quad4 / Hexa8 Poisson: the production poisson module's own tests (modules/poisson/CMakeLists.txt:76-91,183-187;
# inputs/circle.2D.quad.arc, circle.neumann.2D.quad.arc with neumann value=2, sphere.3D.hexa.arc)
Q1_CASES = {
    "circle_2D_quad": dict(mesh="circle_cut.quad.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e30,
                           golden="poisson_test_ref_circle_2D_quad.txt"),
    "circle_scalar_neumann_2D_quad": dict(mesh="circle_cut.quad.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [2.0])], penalty=1.0e30,
                                          golden="poisson_test_ref_circle_scalar_neumann_2D_quad.txt"),
    "sphere_3D_hexa": dict(mesh="sphere_cut.hexa.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e30,
                           golden="poisson_test_ref_sphere_3D_hexa.txt"),
    # inputs/circle.neumann.2D.quad.arc, sphere.neumann.3D.hexa.arc (flux vector q: q.n with the outward normal) and the
    # latter with neumann value=3.1 (CMakeLists.txt:188-198)
    "circle_neumann_2D_quad": dict(mesh="circle_cut.quad.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [-0.35, 1.65])], penalty=1.0e30,
                                   golden="poisson_test_ref_circle_neumann_2D_quad.txt"),
    "sphere_neumann_3D_hexa": dict(mesh="sphere_cut.hexa.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [0.35, 1.65, 3.75])], penalty=1.0e30,
                                   face_order="arcane",  # upstream takes the Quad4-face normal from Arcane's node order (see mesh.arcane_face_node_order)
                                   golden="poisson_test_ref_sphere_neumann_3D_hexa.txt"),
    "sphere_scalar_neumann_3D_hexa": dict(mesh="sphere_cut.hexa.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [3.1])], penalty=1.0e30,
                                          golden="poisson_test_ref_sphere_scalar_neumann_3D_hexa.txt"),
}

# Laplace module (modules/laplace: no source term; inputs/ring.dirichlet-neumann.quad.arc, truncated-cube.neumann.3D.hexa.arc,
# L-shape.3D.arc run as BSR and AF-BSR, CMakeLists.txt:71-89,112-123)
Q1_CASES.update({
    "laplace_ring_quad": dict(mesh="ring.quad.msh", f=0.0, dirichlet=[("inner", 50.0)], neumann=[("outer", [17.8])], penalty=1.0e30,
                              golden="laplace_test_ring_quad.txt"),
    "laplace_truncated_cube_hexa": dict(mesh="truncated_cube.hexa.msh", f=0.0, dirichlet=[("horizontal", 1.8)], neumann=[("bottom", [3.1])], penalty=1.0e30,
                                        golden="laplace_test_trucated-cube_hexa.txt"),
    "laplace_L-shape_3D": dict(mesh="L-shape-3D.msh", f=0.0, dirichlet=[("bot", 50.0), ("bc", 10.0)], penalty=1.0e30,
                               golden="laplace_test_3D_L-shape.txt"),
})

# Production poisson module on P1 cells (BSR-lambda element matrices, BC-service source and flux terms; modules/poisson/inputs/
# circle.2D.arc, circle.neumann.2D.arc, sphere.3D.arc, sphere.neumann.3D.arc; CMakeLists.txt:65-75,173-182)
Q1_CASES.update({
    "poissonmod_circle_2D": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e30, golden="poissonmod_test_ref_circle_2D.txt"),
    "poissonmod_circle_neumann_2D": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [-0.35, 1.65])], penalty=1.0e30,
                                         golden="poissonmod_test_ref_circle_neumann_2D.txt"),
    "poissonmod_sphere_3D": dict(mesh="sphere_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], penalty=1.0e30, golden="poissonmod_test_ref_sphere_3D.txt"),
    "poissonmod_sphere_neumann_3D": dict(mesh="sphere_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [0.35, 1.65, 3.75])], penalty=1.0e30,
                                         golden="poissonmod_test_ref_sphere_neumann_3D.txt"),
})

# <dirichlet-point> on named nodes (modules/poisson/inputs/perforatedSquare.pointDirichlet.2D.arc; the laplace module's
# PointDirichlet.arc is the same problem with its own golden file)
Q1_CASES.update({
    # (these two golden files carry their solver's residual: Dirichlet nodes read 20.0000000000118; compared at 5e-5, upstream's bar is 1e-4)
    "poissonmod_point_dirichlet_2D": dict(mesh="plancher.msh", f=0.0, penalty=1.0e30, golden="poissonmod_test_point_dirichlet_2D.txt", tol=5.0e-5,
                                          dirichlet=[("topLeftCorner", 50.0), ("topRightCorner", 20.0), ("botLeftCorner", 20.0), ("botRightCorner", 50.0)]),
    "laplace_point_dirichlet_2D": dict(mesh="plancher.msh", f=0.0, penalty=1.0e30, golden="laplace_test3_results.txt", tol=5.0e-5,
                                       dirichlet=[("topLeftCorner", 50.0), ("topRightCorner", 20.0), ("botLeftCorner", 20.0), ("botRightCorner", 50.0)]),
})

# Aerodynamics module (potential flow around an airfoil / an airplane: the Poisson matrix, no source; Dirichlet rows by penalty first,
# then the far-field condition psi = y - angle * x (2-D) / z - angle * x (3-D) on the outer boundary, which overwrites them where the
# groups meet: modules/aerodynamics/FemModule.cc:225-252; inputs/Joukowski.arc, Joukowski.quad.arc, Joukowski_3d.arc, Joukowski_3d.hexa.arc)
def _farfield(dim, angle=0.1):
    return lambda x: x[dim - 1] - angle * x[0]


Q1_CASES.update({
    "aerodynamics_2D": dict(mesh="NACA0012.msh", f=0.0, penalty=1.0e30, golden="aerodynamics_test_2d.txt",
                            dirichlet=[("upperAirfoil", 0.0), ("lowerAirfoil", 0.0), ("FarField", _farfield(2))]),
    "aerodynamics_2D_quad": dict(mesh="NACA0012.quad.msh", f=0.0, penalty=1.0e30, golden="aerodynamics_test_2d.quad.txt",
                                 dirichlet=[("upperAirfoil", 0.0), ("lowerAirfoil", 0.0), ("FarField", _farfield(2))]),
    "aerodynamics_3D": dict(mesh="aerodynamics_3d_coarse.msh", f=0.0, penalty=1.0e30, golden="aerodynamics_test_3d.txt",
                            dirichlet=[("airplane", 0.0), ("outer", _farfield(3))]),
    "aerodynamics_3D_hexa": dict(mesh="aerodynamics_3d_coarse.hexa.msh", f=0.0, penalty=1.0e30, golden="aerodynamics_test_3d.hexa.txt",
                                 dirichlet=[("inner", 0.0), ("outer", _farfield(3))]),
})

# more of the same operator on other meshes and boundary data: modules/poisson/inputs/cube.3D.hexa.arc (source + flux over Quad4
# faces), modules/laplace/inputs/truncated-cube.3D.arc (face and point Dirichlet on Tet4), the electrostatics module (rho = 0,
# epsilon = 1: the same stiffness matrix; inputs/box-rods.arc, box-rods.quad.arc, rod-circle.arc, truncated_cube.hexa.arc)
Q1_CASES.update({
    "poissonmod_cube_3D_hexa8": dict(mesh="3x3x3_cube_hexa8.msh", f=9.8, dirichlet=[("left", 0.5)], neumann=[("right", [13.9])], penalty=1.0e30,
                                     golden="poissonmod_test_ref_cube_3D_hexa8.txt"),
    "laplace_truncated_cube_point": dict(mesh="truncated_cube.msh", f=0.0, dirichlet=[("center", 18.8), ("left", 0.8)], penalty=1.0e30,
                                         golden="laplace_test_3D_truncated-cube.txt"),
    "electrostatics_box_rods": dict(mesh="box-rods.msh", f=0.0, dirichlet=[("rod1", -1.0), ("rod2", 1.0), ("external", 0.0)], penalty=1.0e30,
                                    golden="electrostatics_test_1.txt"),
    "electrostatics_box_rods_quad": dict(mesh="box-rods.quad.msh", f=0.0, dirichlet=[("rod1", -1.0), ("rod2", 1.0), ("external", 0.0)], penalty=1.0e30,
                                         golden="electrostatics_box-rods.quad.txt"),
    "electrostatics_rod_circle": dict(mesh="box-rod-circle.msh", f=0.0, dirichlet=[("rod1", -1.0), ("circle", 1.0), ("external", 0.0)], penalty=1.0e30,
                                      golden="electrostatics_test_2.txt"),
    "electrostatics_truncated_cube_hexa": dict(mesh="truncated_cube.hexa.msh", f=0.0, neumann=[("verticalYZ", [-1.0])],
                                               dirichlet=[("verticalYZ", -1.0), ("horizontal", 0.0)], penalty=1.0e30,
                                               golden="electrostatics_truncated-cube.hexa.txt"),
})

# Fourier module (heat conduction: the stiffness operator times a per-cell conductivity, modules/fourier/ElementMatrix.h;
# inputs/conduction.arc, conduction.quad.arc, conduction.3D.arc, conduction.sphere-cut.hexa.arc and the two-material
# conduction.heterogeneous{,.quad}.arc; CMakeLists.txt:91-124).  `conductivity`: a number or {volume name: value}.
Q1_CASES.update({
    # (test1_results / test2_results predate the present inputs: their Dirichlet nodes read 5.0004 where the penalty gives 5 to 1e-12;
    # they hold at upstream's own bar of 1e-4, the four other files to 5e-15)
    "fourier_conduction_2D": dict(mesh="plancher.msh", f=1.0e5, conductivity=1.75, penalty=1.0e12, golden="fourier_test1_results.txt", tol=1.0e-4,
                                  dirichlet=[("Cercle", 50.0), ("Bas", 5.0), ("Haut", 21.0)], neumann=[("Droite", [15.0]), ("Gauche", [0.0])]),
    "fourier_conduction_quad": dict(mesh="plancher.quad4.msh", f=1.0e5, conductivity=1.75, penalty=1.0e12, golden="fourier_conduction_quad.txt",
                                    dirichlet=[("Cercle", 50.0), ("Bas", 5.0), ("Haut", 21.0)], neumann=[("Droite", [15.0]), ("Gauche", [0.0])]),
    "fourier_conduction_3D": dict(mesh="bar_dynamic_3D.msh", f=1.123e-2, conductivity=0.023, penalty=1.0e30, golden="fourier_test_conduction_3D.txt",
                                  dirichlet=[("surfaceleft", 55.0), ("surfaceright", 12.0)]),
    "fourier_conduction_hexa": dict(mesh="sphere_cut.hexa.msh", f=1.123e-2, conductivity=23.5, penalty=1.0e30, golden="fourier_conduction_hexa.txt",
                                    dirichlet=[("horizontal", 55.0)], neumann=[("curved", [1003.67])]),
    "fourier_two_materials": dict(mesh="multi-material.msh", f=15.0, conductivity={"Mat1": 100.0, "Mat2": 1.0}, penalty=1.0e30,
                                  golden="fourier_test2_results.txt", tol=1.0e-4, dirichlet=[("Left", 50.0), ("Right", 5.0)]),
    "fourier_two_materials_quad": dict(mesh="multi-material.quad.msh", f=15.0, conductivity={"Mat1": 100.0, "Mat2": 1.0}, penalty=1.0e30,
                                       golden="fourier_conduction_multi-mat_quad.txt", dirichlet=[("Left", 50.0), ("Right", 5.0)]),
})


# Acoustics module (Helmholtz: alpha * stiffness + kc2 * consistent mass, no Dirichlet condition, scalar flux; modules/acoustics/
# ElementMatrix.h:14,29 uses alpha = -1 on Tri3 / Tet4, ElementMatrixHexQuad.h alpha = +1 on Quad4 / Hexa8; inputs/sub.arc, sub.quad.arc,
# 3d_sub.arc, 3d_sphere_in_sphere.hexa.arc)
ACOUSTICS_CASES = {
    "sub_2D": dict(mesh="sub.msh", alpha=-1.0, kc2=1.1, neumann=[("inner1", [1.0])], golden="acoustics_sub_2D.txt"),
    "sub_2D_quad": dict(mesh="sub.quad.msh", alpha=1.0, kc2=1.1, neumann=[("inner1", [1.0])], golden="acoustics_sub_2D.quad.txt"),
    "sub_3D": dict(mesh="sub_3d.msh", alpha=-1.0, kc2=18.0e5, neumann=[("inner", [11.0e2])], golden="acoustics_sphere_3d.txt"),
    "sphere_in_sphere_hexa": dict(mesh="sphere_in_sphere.hexa.msh", alpha=1.0, kc2=18.0e5, neumann=[("inner", [11.0e2])], golden="acoustics_sphere_3d.hexa.txt",
                                  tol=1.0e-6),  # (indefinite system with kc2 = 1.8e6: 4e-7 between two direct solves)
}


def cell_coefficient(mesh, case):
    """per-cell conductivity of a case ([nb_cell]) or None"""
    k = case.get("conductivity")
    if k is None:
        return None
    out = np.zeros(mesh.cells.shape[0])
    if isinstance(k, dict):
        for name, v in k.items():
            out[mesh.cell_groups[name]] = v
    else:
        out[:] = k
    return out


# Neumann flux cases of testlab (circle_cut.msh; modules/testlab/inputs/Test.circle.2D.trac*.arc): value = scalar flux,
# valueX/valueY = flux vector q (q.n with the outward normal)
NEUMANN_CASES = {
    "trac": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("vertical", [1.0e4])], penalty=1.0e30,
                 golden="poisson_test_ref_circle_trac_2D.txt"),
    "x-trac": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [2.9e4, 0.0])], penalty=1.0e30,
                   golden="poisson_test_ref_circle_x-trac_2D.txt"),
    "y-trac": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [0.0, 2.9e4])], penalty=1.0e30,
                   golden="poisson_test_ref_circle_y-trac_2D.txt"),
    "vect-trac": dict(mesh="circle_cut.msh", f=5.5, dirichlet=[("horizontal", 0.5)], neumann=[("curved", [2.9e4, -1.8e4])], penalty=1.0e30,
                      golden="poisson_test_ref_circle_vect-trac_2D.txt"),
}

# modules/elasticity/inputs/bar.2D.Dirichlet.traction.arc (traction "1.0 NULL" on `right`, no body force)
TRACTION_CASE = dict(mesh="bar.msh", E=21.0e5, nu=0.28, f=[0.0, 0.0], dirichlet=[("left", [0.0, 0.0])], traction=[("right", [1.0, 0.0])],
                     penalty=1.0e30, golden="elasticity_bar.2D.Dirichlet.traction.txt")

ELASTICITY_CASES = {
    # modules/elasticity/inputs/bar.2D.Dirichlet.bodyForce.arc
    "bar_2D": dict(mesh="bar.msh", E=21.0e5, nu=0.28, f=[0.0, -1.0], dirichlet=[("left", [0.0, 0.0])], penalty=1.0e30,
                   golden="elasticity_bar.2D.Dirichlet.bodyForce.txt"),
    # modules/elasticity/inputs/bar.3D.Dirichlet.bodyForce.arc
    "bar_3D": dict(mesh="bar_dynamic_3D.msh", E=21.0e5, nu=0.28, f=[-1.0, 0.0, 0.0],
                   dirichlet=[("surfaceleft", [0.0, 0.0, 0.0]), ("surfaceright", [None, 1.0, None])], penalty=1.0e30,
                   golden="elasticity_bar.3D.Dirichlet.bodyForce.txt"),
}

# Quad4 / Hexa8 elasticity (modules/elasticity/ElementMatrixHexQuad.h): inputs/2D.dirichlet.bodyforce.quad.arc,
# bar.2D.Dirichlet.bodyForce.quad.arc, 3D.dirichlet.bodyforce.hexa.arc
def cartesian_mesh(n, length):
    """Arcane's Cartesian2D / Cartesian3D mesh generator as the reference's .arc files use it (`<generator name="Cartesian2D">` with
    `<x><n>..</n><length>..</length></x>` ..., origin 0, generate-sod-groups): Quad4 / Hexa8 cells, node uniqueId = i + j (nx+1) + k (nx+1)(ny+1),
    face groups XMIN / XMAX / YMIN / ... as node groups (all the Dirichlet conditions need)."""
    from arcanefem_b200 import mesh as M
    dim = len(n)
    npts = [k + 1 for k in n]
    grids = np.meshgrid(*[np.arange(p) for p in npts], indexing="ij")
    stride = [1, npts[0], npts[0] * npts[1]][:dim]
    uid = sum(g * s for g, s in zip(grids, stride)).ravel()
    order = np.argsort(uid)
    ijk = np.stack([g.ravel() for g in grids], axis=1)[order]
    coords = np.zeros((ijk.shape[0], 3))
    for a in range(dim):
        coords[:, a] = ijk[:, a] * (length[a] / n[a])
    node = lambda *idx: sum(i * s for i, s in zip(idx, stride))
    cells = []
    if dim == 2:
        for j in range(n[1]):
            for i in range(n[0]):
                cells.append([node(i, j), node(i + 1, j), node(i + 1, j + 1), node(i, j + 1)])
    else:
        for k in range(n[2]):
            for j in range(n[1]):
                for i in range(n[0]):
                    cells.append([node(i, j, k), node(i + 1, j, k), node(i + 1, j + 1, k), node(i, j + 1, k),
                                  node(i, j, k + 1), node(i + 1, j, k + 1), node(i + 1, j + 1, k + 1), node(i, j + 1, k + 1)])
    groups = {}
    for a, axis in enumerate("XYZ"[:dim]):
        groups[axis + "MIN"] = np.flatnonzero(ijk[:, a] == 0).astype(np.int32)
        groups[axis + "MAX"] = np.flatnonzero(ijk[:, a] == n[a]).astype(np.int32)
    return M.Mesh(dim=dim, coords=coords, cells=np.array(cells, dtype=np.int32), node_uid=np.arange(ijk.shape[0], dtype=np.int64), groups=groups)


def load_mesh(name):
    """a reference mesh file of tests/golden/, or `cartesian:nx,ny[,nz]:lx,ly[,lz]` for Arcane's generated meshes"""
    from arcanefem_b200 import mesh as M
    if name.startswith("cartesian:"):
        _, n, length = name.split(":")
        return cartesian_mesh([int(v) for v in n.split(",")], [float(v) for v in length.split(",")])
    return M.read_msh(os.path.join(GOLDEN, name))


Q1_ELASTICITY_CASES = {
    # Arcane's cartesian generator instead of a mesh file (inputs/bar.2D.cartesian.Dirichlet.bodyForce.arc, bar.3D.cartesian.Dirichlet.bodyForce.arc)
    "cartesian_bar_2D": dict(mesh="cartesian:10,2:1.0,0.1", E=21.0e5, nu=0.28, f=[-1.0, 0.0], dirichlet=[("XMIN", [0.0, 0.0]), ("XMAX", [None, 1.0])],
                             penalty=1.0e30, golden="elasticity_bar.2D.cartesian.Dirichlet.bodyForce.txt"),
    "cartesian_bar_3D": dict(mesh="cartesian:10,2,2:1.0,0.1,0.04", E=21.0e5, nu=0.28, f=[-1.0, 0.0, 0.0],
                             dirichlet=[("XMIN", [0.0, 0.0, 0.0]), ("XMAX", [None, 1.0, None])], penalty=1.0e30,
                             golden="elasticity_bar.3D.cartesian.Dirichlet.bodyForce.txt", min_rel=1.0e-6),  # (mid-plane components are 1e-9 of the deflection: cancellation)
    "five_quads": dict(mesh="five_quads.msh", E=200e9, nu=0.3, f=[-9818949214245.0, -7818949234281.0],
                       dirichlet=[("bot", [0.0, 0.0]), ("top", [1.9, 14.5])], penalty=1.0e30, golden="elasticity_2D.dirichlet.bodyforce.quad.txt"),
    "plate_quad": dict(mesh="plate.quad.msh", E=21.0e5, nu=0.28, f=[0.0, -1.0], dirichlet=[("left", [0.0, 0.0])], penalty=1.0e30,
                       golden="elasticity_bar.2D.Dirichlet.bodyForce.quad.txt"),
    # inputs/2D.dirichlet.traction.bodyforce.quad.arc, bar.2D.traction.bodyforce.quad.arc, 3D.dirichlet.traction.bodyforce.hexa.arc
    "five_quads_traction": dict(mesh="five_quads.msh", E=200e9, nu=0.3, f=[9.8e9, 7.3e9], dirichlet=[("bot", [0.0, 0.0])], traction=[("top", [13.3e9, 14.5e5])],
                                penalty=1.0e30, golden="elasticity_2D.dirichlet.traction.bodyforce.quad.txt"),
    "plate_quad_traction": dict(mesh="plate.quad.msh", E=21.0e5, nu=0.28, f=[3.33, -6.66], dirichlet=[("left", [0.0, 0.0])], traction=[("right", [1.33, 2.13])],
                                penalty=1.0e30, golden="elasticity_bar.2D.traction.bodyforce.quad.txt"),
    "truncated_cube_hexa_traction": dict(mesh="truncated_cube.hexa.msh", E=200e9, nu=0.3, f=[-9.8e1, -7.5e1, 5.9e1], dirichlet=[("top", [0.0, 0.0, 0.0])],
                                         traction=[("bottom", [12.9e11, -14.5e11, -18.8e11])], penalty=1.0e30,
                                         golden="elasticity_3D.dirichlet.traction.bodyforce.hexa.txt"),
    # inputs/bar.2D.traction.bodyforce.arc (Tri3: traction and body force together)
    "bar_2D_traction_bodyforce": dict(mesh="bar.msh", E=21.0e5, nu=0.28, f=[3.33, -6.66], dirichlet=[("left", [0.0, 0.0])], traction=[("right", [1.33, 2.13])],
                                      penalty=1.0e30, golden="elasticity_bar.2D.traction.bodyforce.txt"),
    # inputs/bar.2D.PointDirichlet.Dirichlet.bodyForce.arc (component-wise Dirichlet on faces and on named nodes)
    "bar_2D_point_dirichlet": dict(mesh="bar.msh", E=21.0e5, nu=0.28, f=[0.0, -1.0], penalty=1.0e30, golden="elasticity_bar.2D.PointDirichlet.Dirichlet.bodyForce.txt",
                                   dirichlet=[("left", [0.0, None]), ("right", [1.0, None]), ("botLeft", [0.0, 0.0]), ("botRight", [None, 0.0])]),
    "truncated_cube_hexa": dict(mesh="truncated_cube.hexa.msh", E=200e9, nu=0.3, f=[-9.8e12, -7.5e12, 5.9e12],
                                dirichlet=[("top", [1.0, 2.0, 8.0]), ("bottom", [12.9, -14.5, -18.8])], penalty=1.0e30,
                                golden="elasticity_3D.dirichlet.bodyforce.hexa.txt"),
}

# modules/elastodynamics/inputs/bar.arc, bar.3D.arc: Newmark-beta time loop (gamma = 1/2, beta = 1/4, no damping) on the
# stiffness + mass operator; the golden files hold the displacement of the last step
ELASTODYNAMICS_CASES = {
    "bar_2D": dict(mesh="bar_dynamic.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 0.0],
                   dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[("surfaceright", [0.0, 0.01])], penalty=1.0e30, golden="elastodynamics_2D_bar.txt"),
    "bar_3D": dict(mesh="bar_dynamic_3D.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=0.5, f=[131.0e2, 113.8e6, 567.0e8],
                   dirichlet=[("surfaceleft", [0.0, 0.0, 0.0])], traction=[("surfaceright", [0.0, 1869.1e2, 0.0])], penalty=1.0e30,
                   golden="elastodynamics_bar_3d.txt"),
    # Quad4 / Hexa8 (modules/elastodynamics/ElementMatrixHexQuad.h): inputs/bar.quad.arc, bar.3D.hexa.arc
    "bar_quad": dict(mesh="bar_dynamic_quad.msh", rho=12.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 13.5e2],
                     dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[], penalty=1.0e30, golden="elastodynamics_bar.quad.txt",
                     min_rel=1.0e-8),  # one x-displacement on the symmetry line is 7e-9 against 15 elsewhere: pure cancellation, skipped
    "bar_3D_hexa": dict(mesh="bar_dynamic_3Dhexa.msh", rho=1232434.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=0.5, f=[131.0e2, 113.8e6, 567.0e8],
                        dirichlet=[("left", [0.0, 0.0, 0.0])], traction=[], penalty=1.0e30, golden="elastodynamics_bar_3d.hexa.txt", min_rel=1.0e-8),
    # traction over Quad4 edges / Hexa8 faces, body force + traction, <dirichlet-point> conditions (inputs/bar.dirichlet-traction.quad.arc,
    # bar.3D.dirichlet-traction.hexa.arc, bar.dirichlet.traction.bodyforce.quad.arc, semi-circle.pointBC.arc, truncated-cube.pointBC.arc)
    "bar_quad_traction": dict(mesh="bar_dynamic_quad.msh", rho=12.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 13.5e2],
                              dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[("surfaceright", [1.7e6, 2.7e6])], penalty=1.0e30,
                              golden="elastodynamics_bar_2d_dirichlet-traction.quad.txt", min_rel=1.0e-8),
    "bar_3D_hexa_traction": dict(mesh="bar_dynamic_3Dhexa.msh", rho=1232434.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=0.5, f=[0.0, 0.0, 0.0],
                                 dirichlet=[("left", [0.0, 0.0, 0.0])], traction=[("right", [1.7e6, 2.7e6, 2.4e7])], penalty=1.0e30,
                                 golden="elastodynamics_bar_3d_dirichlet-traction.hexa.txt", min_rel=1.0e-8),
    "bar_quad_traction_bodyforce": dict(mesh="bar_dynamic_quad.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=1.0, f=[0.0, -2000.0],
                                        dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[("surfaceright", [0.0, -1.0])], penalty=1.0e30,
                                        golden="elastodynamics_bar_dirichlet_traction_bodyforce.quad.txt", min_rel=1.0e-8),
    "semi_circle_point": dict(mesh="semi-circle.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=1.0, f=[0.0, 0.0],
                              dirichlet=[("boderCircle", [0.0, 0.0]), ("source", [10.0, 10.0])], traction=[], penalty=1.0e30,
                              golden="elastodynamics_semi-ciricle_point-bc.txt", min_rel=1.0e-8),
    "truncated_cube_point": dict(mesh="truncated_cube.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=1.0, f=[145.5e5, 56456.5e6, 87842.5e5],
                                 dirichlet=[("bottom", [0.0, 0.0, 0.0]), ("center", [18.0, 13.0, 14.0])], traction=[], penalty=1.0e30,
                                 golden="elastodynamics_truncated-cube_point-bc.txt", min_rel=1.0e-8),
    # traction from a table in time (traction_table: surface -> file of `t tx ty tz` rows, linear in between): inputs/bar.transient-traction.arc,
    # bar.transient-traction.quad.arc, bar.3D.transient-traction.arc, bar.3D.transient-traction.hexa.arc
    "bar_2D_transient": dict(mesh="bar_dynamic.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 0.0],
                             dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[], traction_table=[("surfaceright", "elastodynamics_traction_bar_test_1.txt")],
                             penalty=1.0e30, golden="elastodynamics_2D_bar_transient_traction.txt", min_rel=1.0e-8),
    "bar_quad_transient": dict(mesh="bar_dynamic_quad.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 0.0],
                               dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[], traction_table=[("surfaceright", "elastodynamics_traction_bar_test_1.txt")],
                               penalty=1.0e30, golden="elastodynamics_bar_transient-traction.quad.txt", min_rel=1.0e-8),
    # Rayleigh damping (inputs/bar.damping.arc): etam on the mass, etak on the stiffness
    "bar_2D_damping": dict(mesh="bar_dynamic.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 0.0], etam=0.01, etak=0.01,
                           dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[("surfaceright", [0.0, 0.01])], penalty=1.0e30,
                           golden="elastodynamics_2D_bar_constant_traction_damping.txt", min_rel=1.0e-8),
    # generalized-alpha time discretization (inputs/bar.Galpha.arc)
    "bar_2D_galpha": dict(mesh="bar_dynamic.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=2.0, f=[0.0, 0.0], alpm=0.20, alpf=0.40,
                          dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[("surfaceright", [0.0, 0.01])], penalty=1.0e30,
                          golden="elastodynamics_2D_Galpha_time_discretization.txt", min_rel=1.0e-8),
    "bar_quad_three_steps": dict(mesh="bar_dynamic_quad.msh", rho=12.0, lam=576.9230769, mu=384.6153846, dt=0.08, tmax=0.25, f=[0.0, -13.5e5],
                                 dirichlet=[("surfaceleft", [0.0, 0.0])], traction=[], traction_table=[("surfaceright", "elastodynamics_traction_bar_three_steps.txt")],
                                 penalty=1.0e30, golden="elastodynamics_2D_bar_transient_traction_three_steps.txt", min_rel=1.0e-8),  # (inputs/bar.dirichlet.three-step-traction.bodyforce.quad.arc)
    "bar_3D_transient": dict(mesh="bar_dynamic_3D.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.1, tmax=0.5, f=[0.0, 0.0, 0.0],
                             dirichlet=[("surfaceleft", [0.0, 0.0, 0.0])], traction=[], traction_table=[("surfaceright", "elastodynamics_traction_bar_test_1.txt")],
                             penalty=1.0e30, golden="elastodynamics_bar_3d_transient-traction.txt", min_rel=1.0e-8),
    "bar_3D_hexa_transient": dict(mesh="bar_dynamic_3Dhexa.msh", rho=1.0, lam=576.9230769, mu=384.6153846, dt=0.1, tmax=0.5, f=[0.0, 0.0, 0.0],
                                  dirichlet=[("left", [0.0, 0.0, 0.0])], traction=[], traction_table=[("right", "elastodynamics_traction_bar_test_1.txt")],
                                  penalty=1.0e30, golden="elastodynamics_bar_3d_transient-traction.hexa.txt", min_rel=1.0e-8),
}


def transient_traction(case, b, unit_rhs):
    """traction tables of a case: unit_rhs(group, k) -> right-hand side of a unit traction in direction k on the surface (the term is linear in the
    traction vector, femutils/ArcaneFemFunctions.h:2987-2996).  Returns rhs(t), the sum over the surfaces of table(t)[k] * unit_rhs(group, k)."""
    terms = []
    for group, fname in case.get("traction_table", []):
        table = np.loadtxt(os.path.join(GOLDEN, fname))
        terms.append((table, [unit_rhs(group, k) for k in range(b)]))

    def rhs(t):
        out = 0.0
        for table, units in terms:
            for k, u in enumerate(units):
                out = out + float(np.interp(t, table[:, 0], table[:, 1 + k])) * u
        return out
    return rhs


def golden_floor(case, golden):
    """smallest golden value compared: upstream's 1e-14, or `min_rel` times the largest golden value where a case has entries that are
    zero up to cancellation"""
    gmax = max(abs(v) for vals in golden.values() for v in vals)
    return max(1.0e-14, case.get("min_rel", 0.0) * gmax)


def _time_constants(case):
    """modules/elastodynamics/FemModule.cc:197-232: c0 .. c10 of the Newmark-beta scheme (alpm = alpf = 0) and of the generalized-alpha scheme
    (a case with alpm / alpf), Rayleigh damping etam / etak included"""
    rho, dt, lam, mu = case["rho"], case["dt"], case["lam"], case["mu"]
    etam, etak = case.get("etam", 0.0), case.get("etak", 0.0)
    alpm, alpf = case.get("alpm", 0.0), case.get("alpf", 0.0)
    gamma = 0.5 + alpf - alpm
    beta = (1. / 4.) * (gamma + 0.5) * (gamma + 0.5)
    c = [0.0] * 11
    c[0] = rho * (1. - alpm) / (beta * dt * dt) + etam * rho * gamma * (1 - alpf) / beta / dt
    c[1] = lam * (1. - alpf) + lam * etak * gamma * (1. - alpf) / beta / dt
    c[2] = mu * (1. - alpf) + mu * etak * gamma * (1. - alpf) / beta / dt
    c[3] = rho * (1. - alpm) / beta / dt - etam * rho * (1 - gamma * (1 - alpf) / beta)
    c[4] = rho * ((1. - alpm) * (1. - 2. * beta) / 2. / beta - alpm - etam * dt * (1. - alpf) * (1. - gamma / 2 / beta))
    c[5] = lam * alpf - lam * etak * gamma * (1. - alpf) / beta / dt
    c[6] = mu * alpf - mu * etak * gamma * (1. - alpf) / beta / dt
    c[7] = etak * lam * (gamma * (1. - alpf) / beta - 1)
    c[8] = etak * lam * dt * (1. - alpf) * ((1. - 2 * beta) / 2. / beta - (1. - gamma))
    c[9] = etak * mu * (gamma * (1. - alpf) / beta - 1)
    c[10] = etak * mu * dt * (1. - alpf) * ((1. - 2 * beta) / 2. / beta - (1. - gamma))
    return gamma, beta, c


def newmark_coefficients(case):
    """gamma, beta, c0 (mass), c1 (lambda-like), c2 (mu-like), c3, c4 (mass terms of the right-hand side)"""
    gamma, beta, c = _time_constants(case)
    return gamma, beta, c[0], c[1], c[2], c[3], c[4]


def newmark_damping_terms(case, stiff_times):
    """modules/elastodynamics/SourceTerm.h:79-84: the stiffness-type right-hand side terms (Rayleigh damping, generalized-alpha),
    -K(c5, c6) U + K(c7, c9) V + K(c8, c10) A with K(lambda, mu) the elasticity matrix; stiff_times(lambda, mu, x) = K(lambda, mu) x.
    Returns f(U, V, A), or None when c5 .. c10 are all zero."""
    _, _, c = _time_constants(case)
    if not any(c[5:]):
        return None
    return lambda U, V, A: -stiff_times(c[5], c[6], U) + stiff_times(c[7], c[9], V) + stiff_times(c[8], c[10], A)


def newmark_time_loop(case, nb_dof, solve_step, mass_times, damping=None):
    """The module's time loop (FemModule.cc:29-131, 277-330): t starts at dt, the loop ends after the step that starts with
    t >= tmax - dt.  solve_step(rhs_dynamic, t) -> displacement of the step (the caller adds the static loads and the Dirichlet rows);
    mass_times(x) = consistent mass matrix (rho = 1 per component) times x.  Returns the last displacement."""
    gamma, beta, c0, _, _, c3, c4 = newmark_coefficients(case)
    dt = case["dt"]
    t, tmax = dt, case["tmax"] - dt
    U = np.zeros(nb_dof)
    V = np.zeros(nb_dof)
    A = np.zeros(nb_dof)
    dU = U
    while True:
        last = t >= tmax
        dU = solve_step(mass_times(c0 * U + c3 * V + c4 * A) + (0.0 if damping is None else damping(U, V, A)), t)
        a_new = (dU - U - dt * V) / (beta * dt * dt) - (1. - 2. * beta) / (2. * beta) * A
        V = V + dt * ((1. - gamma) * A + gamma * a_new)
        A = a_new
        U = dU
        t += dt
        if last:
            break
    return dU


# Heat module (implicit Euler on lambda * stiffness + mass / dt; modules/heat/ElementMatrix.h:36, :107, ElementMatrixHexQuad.h;
# inputs/conduction.arc, conduction.quad.arc, 3d_conduction.arc).  The golden files hold the temperature of the last step.
HEAT_CASES = {
    "plate_2D": dict(mesh="plate.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("left", 10.0)], penalty=1.0e31,
                     golden="heat_2d_conduction.txt"),
    "plate_2D_quad": dict(mesh="plate.quad.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("left", 10.0)], penalty=1.0e31,
                          golden="heat_2d_conduction.quad.txt"),
    "truncated_cube_3D": dict(mesh="truncated_cube.msh", lam=1.75, dt=0.1, tmax=1.0, Tinit=30.0,
                              dirichlet=[("top", 100.6), ("bottom", 1.6)], penalty=1.0e30, golden="heat_3d_conduction.txt"),
    # convection (Robin) boundaries: (surface, h, T_ext) -- h * face mass matrix into the matrix, h * T_ext * measure / nodes into the right-hand side
    # (modules/heat/FemModule.cc:305-348); flux boundaries (scalar <neumann>) and <dirichlet-point> conditions
    "plate_2D_convection": dict(mesh="plate.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("left", 10.0)], penalty=1.0e31,
                                convection=[("right", 1.0, 20.0), ("top", 1.0, 20.0), ("bottom", 1.0, 20.0)], golden="heat_2d_conduction_convection.txt"),
    "plate_2D_convection_quad": dict(mesh="plate.quad.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("left", 10.0)], penalty=1.0e31,
                                     convection=[("right", 5.0, 15.0)], golden="heat_2d_conduction_convection.quad.txt"),
    "plate_2D_neumann_points": dict(mesh="plate.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("topLeft", 1.8), ("botRight", 31.0)], penalty=1.0e31,
                                    neumann=[("left", [9.6])], golden="heat_2d_conduction_neumann_pointBC.txt"),
    "plate_2D_neumann_points_quad": dict(mesh="plate.quad.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("topLeft", 1.8), ("botRight", 31.0)],
                                         penalty=1.0e31, neumann=[("left", [9.6])], golden="heat_2d_conduction_neumann_pointBC.quad.txt"),
    "truncated_cube_3D_convection_point": dict(mesh="truncated_cube.msh", lam=1.75, dt=0.1, tmax=1.0, Tinit=30.0, dirichlet=[("bottom", 2.6), ("center", 100.6)],
                                               penalty=1.0e30, convection=[("top", 1.9, 30.0)], golden="heat_3d_conduction_convection_pointBC.txt"),
    "truncated_cube_3D_convection_hexa": dict(mesh="truncated_cube.hexa.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("left", 10.0)], penalty=1.0e30,
                                              convection=[("right", 5.0, 15.0)], golden="heat_3d_conduction_convection.hexa.txt"),
    "truncated_cube_3D_neumann_hexa": dict(mesh="truncated_cube.hexa.msh", lam=1.75, dt=0.4, tmax=20.0, Tinit=30.0, dirichlet=[("top", 10.6)], penalty=1.0e30,
                                           neumann=[("bottom", [1.2])], golden="heat_3d_conduction_neumann.hexa.txt"),
}


def add_in_pattern(rows, cols, vals, B):
    """vals + the entries of the scipy matrix B, inside the CSR pattern (rows, cols) with ascending columns per row (what a sequence of
    matrixAddValue calls does); every entry of B has to exist in the pattern"""
    out = np.array(vals, dtype=np.float64, copy=True)
    Bc = B.tocoo()
    for r, c, v in zip(Bc.row, Bc.col, Bc.data):
        lo, hi = int(rows[r]), int(rows[r + 1])
        k = lo + int(np.searchsorted(cols[lo:hi], c))
        assert k < hi and cols[k] == c, (r, c)
        out[k] += v
    return out


def face_measure(mesh, face):
    """femutils/ArcaneFemFunctions.h:172-215: edge length, triangle area, quadrilateral area as two triangles (n1: n2-n1 x n0-n1, n3: n0-n3 x n2-n3)"""
    x = mesh.coords[np.asarray(face, dtype=np.int64)]
    if len(face) == 2:
        return float(np.linalg.norm(x[1] - x[0]))
    if len(face) == 3:
        return float(np.linalg.norm(np.cross(x[1] - x[0], x[2] - x[0]))) / 2.0
    return 0.5 * float(np.linalg.norm(np.cross(x[2] - x[1], x[0] - x[1])) + np.linalg.norm(np.cross(x[0] - x[3], x[2] - x[3])))


def convection_boundary_terms(mesh, case):
    """modules/heat/FemModule.cc:305-348: per face of a convection surface h * factor * massMatrix(1, 1) * measure into the matrix (factor 1/6 on
    edges, 1/12 on triangles, 1/20 on quadrilaterals) and h * T_ext * measure / nodes into the right-hand side.  Returns (scipy CSR, vector)."""
    import scipy.sparse as sp
    n = mesh.nb_node
    rows, cols, vals = [np.empty(0, np.int64)], [np.empty(0, np.int64)], [np.empty(0)]
    rhs = np.zeros(n)
    for group, h, text in case.get("convection", []):
        for face in mesh.faces[group]:
            k = len(face)
            measure = face_measure(mesh, face)
            Ke = h * {2: 1 / 6., 3: 1 / 12., 4: 1 / 20.}[k] * _mass_matrix(np.ones(k), np.ones(k)) * measure
            f = np.asarray(face, dtype=np.int64)
            rows.append(np.repeat(f, k))
            cols.append(np.tile(f, k))
            vals.append(Ke.ravel())
            rhs[f] += (h * text) * measure / k
    B = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
    B.sum_duplicates()
    return B, rhs


def heat_time_loop(case, nb_node, solve_step, mass_times, static=None):
    """The module's time loop (modules/heat/FemModule.cc:74-135, 232-245, 350-363): t starts at 0 and grows by dt after every solve; the
    loop ends with the first step that starts at t >= tmax (that step is still solved, and checked against the golden file).
    solve_step(rhs) -> temperature of the step (the caller sets the Dirichlet rows); mass_times(x) = consistent mass matrix times x.
    The right-hand side of a step is the mass matrix applied to (previous temperature / dt), plus `static` (convection and flux terms)."""
    dt, tmax = case["dt"], case["tmax"]
    t = 0.0
    T = np.full(nb_node, case["Tinit"])
    while True:
        last = t >= tmax
        T = solve_step(mass_times(T * (1.0 / dt)) + (0.0 if static is None else static))
        t += dt
        if last:
            break
    return T


# FourierNL module: -div(lambda(u) grad u) = 0 with lambda(u) = (1 + u)^m, solved by Picard iterations; on Tri3 / Tet4 the conductivity of a
# cell is lambda at the mean of its nodal values of the previous iterate (modules/fouriernl/ElementMatrix.h:29-41, ConductivityCoefficient.h:16),
# i.e. the Poisson operator with a per-cell coefficient re-assembled every iteration (inputs/Test.nonlinear.conduction.arc,
# Test.3d.nonlinear.conduction.arc; defaults of Fem.axl: m = 2, nlin-rtol 1e-5, at most 30 iterations).  The module's Quad4 / Hexa8 matrices take
# lambda at the Gauss points from a nodal field -- a different operator, not covered.
FOURIERNL_CASES = {
    "unit_square_tria": dict(mesh="unit_square.msh", m=2.0, rtol=1.0e-5, max_iters=30, dirichlet=[("left", 0.0), ("right", 1.0)], penalty=1.0e30,
                             golden="fouriernl_conduction_tria.txt"),
    "unit_cube_tetra": dict(mesh="unit_cube.msh", m=2.0, rtol=1.0e-5, max_iters=30, dirichlet=[("left", 0.0), ("right", 1.0)], penalty=1.0e30,
                            golden="fouriernl_conduction_tetra.txt"),
}


def picard_loop(case, mesh, solve_with_conductivity):
    """modules/fouriernl/FemModule.cc:155-210, 466-492: u_k starts at zero; every iteration assembles with lambda((mean of u_k over the cell)),
    solves, and stops when max |u - u_k| < nlin-rtol.  solve_with_conductivity(lambda_per_cell) -> u.  Returns (u, iterations)."""
    uk = np.zeros(mesh.nb_node)
    for it in range(1, case["max_iters"] + 1):
        lam = (1.0 + uk[mesh.cells.astype(np.int64)].sum(axis=1) / mesh.npc) ** case["m"]
        u = solve_with_conductivity(lam)
        if np.abs(u - uk).max() < case["rtol"]:
            return u, it
        uk = u
    raise AssertionError("Picard iterations did not converge")


# Soildynamics module, 2-D Tri3 cases: the elastodynamics matrix (same element matrix, modules/soildynamics/ElementMatrix.h) plus the paraxial
# (absorbing) boundary of modules/soildynamics/Paraxial.h on the surface `lower`: every boundary edge adds c7 * length * P to the matrix and
# length * (c7 U - c8 V + c9 A) . P to the right-hand side of each step (inputs/constant-traction.arc, constant-traction.pointbc.arc)
SOILDYNAMICS_CASES = {
    "semi_circle_constant_traction": dict(mesh="semi-circle-soil.msh", E=6.62e6, nu=0.45, rho=2500.0, dt=0.01, tmax=0.08, f=[0.0, 0.0],
                                          traction=[("input", [0.01, 0.01])], paraxial=["lower"], dirichlet=[], penalty=1.0e30,
                                          golden="soildynamics_test_2D_constant_traction.txt"),
    "semi_circle_constant_traction_pointbc": dict(mesh="semi-circle-soil.msh", E=6.62e6, nu=0.45, rho=2500.0, dt=0.01, tmax=0.08, f=[3359.6, 3452.3],
                                                  traction=[("input", [0.01, 0.01])], paraxial=["lower"], dirichlet=[("source", [0.0, 0.0003])],
                                                  penalty=1.0e30, golden="soildynamics_test_2D_constant_traction_pointbc.txt"),
    # traction table in time + body force + paraxial boundary (inputs/transient-traction.arc)
    "semi_circle_transient_traction": dict(mesh="semi-circle-soil.msh", E=6.62e6, nu=0.45, rho=2500.0, dt=0.01, tmax=0.08, f=[0.0, 315.9], traction=[],
                                           traction_table=[("input", "soildynamics_semi-circle-soil-traction.txt")], paraxial=["lower"], dirichlet=[],
                                           penalty=1.0e30, golden="soildynamics_test_2D_transient_traction.txt"),
    # double-couple source (modules/soildynamics/DoubleCouple.h: the force of a time table overwrites the right-hand side at four named nodes --
    # x-DoF of north (+) / south (-), y-DoF of east (-) / west (+)) in a square with paraxial boundaries all around; material given by wave speeds
    "square_double_couple": dict(mesh="square_double-couple.msh", cs=2.0, cp=4.0, rho=1.0, dt=0.01, tmax=0.2, f=[0.0, 0.0], traction=[],
                                 paraxial=["left", "top", "right", "bottom"], dirichlet=[], penalty=1.0e30,
                                 double_couple=dict(north="sourceT", south="sourceB", east="sourceR", west="sourceL", table="soildynamics_force_loading_dc.txt"),
                                 golden="soildynamics_test_paraxial_results.txt"),
    # 3-D: paraxial triangles of a Tet4 mesh, the double couple acts on the x- and z-DoFs (inputs/3d.double-couple.paraxial.soil.arc; the Hexa8
    # twin's mesh file carries its source points as isolated nodes without physical groups -- not reproduced)
    "cube_double_couple_3D": dict(mesh="cube_double_couple_3d.msh", cs=21.5, cp=40.0, rho=9.0, dt=0.01, tmax=0.1, f=[0.0, 0.0, 0.0], traction=[],
                                  paraxial=["leftsur"], dirichlet=[], penalty=1.0e30,
                                  double_couple=dict(north="dcNorth", south="dcSouth", east="dcEast", west="dcWest", table="soildynamics_force_loading_dc.txt"),
                                  golden="soildynamics_3d_test_paraxial_double_couple.txt"),
    "square_double_couple_bodyforce": dict(mesh="square_double-couple.msh", cs=2.0, cp=4.0, rho=1.0, dt=0.01, tmax=0.2, f=[1255.1, 32289.5], traction=[],
                                           paraxial=["left", "top", "right", "bottom"], dirichlet=[], penalty=1.0e30,
                                           double_couple=dict(north="sourceT", south="sourceB", east="sourceR", west="sourceL", table="soildynamics_force_loading_dc.txt"),
                                           golden="soildynamics_test_paraxial_body-force_results.txt"),
}


def double_couple_rhs(case, mesh):
    """modules/soildynamics/DoubleCouple.h:20-63: returns apply(rhs, t), which overwrites the four source DoFs with the table's force at
    time t (linear interpolation between the rows of the table, femutils/FemUtils.cc:182-212), or None for a case without such a source"""
    dc = case.get("double_couple")
    if dc is None:
        return None
    table = np.loadtxt(os.path.join(GOLDEN, dc["table"]))
    b, ew = mesh.dim, mesh.dim - 1  # north / south act on the x-DoF, east / west on the y-DoF in 2-D and on the z-DoF in 3-D
    dofs = [(b * int(n), +1.0) for n in mesh.groups[dc["north"]]] + [(b * int(n), -1.0) for n in mesh.groups[dc["south"]]] \
        + [(b * int(n) + ew, -1.0) for n in mesh.groups[dc["east"]]] + [(b * int(n) + ew, +1.0) for n in mesh.groups[dc["west"]]]

    def apply(rhs, t):
        force = float(np.interp(t, table[:, 0], table[:, 1]))
        for dof, sign in dofs:
            rhs[dof] = sign * force
        return rhs
    return apply


def soildynamics_coefficients(case):
    """modules/soildynamics/FemModule.cc:156-196: Lame constants and wave speeds from (E, nu, rho), Newmark-beta constants c0..c9"""
    rho, dt = case["rho"], case["dt"]
    if "E" in case:
        E, nu = case["E"], case["nu"]
        mu = E / (2 * (1 + nu))
        lam = E * nu / ((1 + nu) * (1 - 2 * nu))
        cs = np.sqrt(mu / rho)
        cp = np.sqrt((lam + 2. * mu) / rho)
    else:  # wave speeds given
        cs, cp = case["cs"], case["cp"]
        mu = cs * cs * rho
        lam = cp * cp * rho - 2 * mu
    gamma = 0.5
    beta = (1. / 4.) * (gamma + 0.5) * (gamma + 0.5)
    return dict(lam=lam, mu=mu, cs=cs, cp=cp, gamma=gamma, beta=beta, c0=rho / (beta * dt * dt), c3=rho / (beta * dt), c4=rho * (1. / 2. / beta - 1.),
                c7=rho * gamma / beta / dt, c8=rho * (1. - gamma / beta), c9=rho * dt * (1. - gamma / (2. * beta)))


def _mass_matrix(u, v):
    """femutils/FemUtils.h:583-597 massMatrix(lhs, rhs): outer product with a doubled diagonal"""
    m = np.outer(u, v)
    m[np.diag_indices_from(m)] *= 2.
    return m


def paraxial_boundary_matrix(mesh, faces, cp, cs):
    """sum over the boundary faces of measure * P_face as a scipy CSR matrix over the dim * nb_node DoFs: edges in 2-D, triangles / quadrilaterals
    in 3-D (modules/soildynamics/Paraxial.h:27-90 `_computeParaxialElementMatrix{Edge2,Tria3,Quad4}`, scatter as :137-190).  With n the unit
    normal (its sign does not matter): P = sum_a (n_a^2 cp + (1 - n_a^2) cs) M(U_a, U_a) + sum_{a != b} n_a n_b (cp - cs) M(U_a, U_b), divided by
    6 / 12 / 20, U_a = 1 on the a-th DoF of every node of the face."""
    import scipy.sparse as sp
    b = mesh.dim
    rows, cols, vals = [], [], []
    for face in np.asarray(faces, dtype=np.int64):
        k = len(face)
        x = mesh.coords[face]
        if k == 2:
            d = x[1] - x[0]
            nrm = np.array([d[1], -d[0]]) / np.sqrt(d[0] * d[0] + d[1] * d[1])
        elif k == 3:
            nrm = np.cross(x[1] - x[0], x[2] - x[0])
            nrm = nrm / np.linalg.norm(nrm)
        else:  # Newell's formula (femutils/ArcaneFemFunctions.h:493-512)
            nxt = np.roll(x, -1, axis=0)
            nrm = np.array([np.sum((x[:, 1] - nxt[:, 1]) * (x[:, 2] + nxt[:, 2])), np.sum((x[:, 2] - nxt[:, 2]) * (x[:, 0] + nxt[:, 0])),
                            np.sum((x[:, 0] - nxt[:, 0]) * (x[:, 1] + nxt[:, 1]))])
            nrm = nrm / np.linalg.norm(nrm)
        U = [np.tile(np.eye(b)[a], k) for a in range(b)]
        P = np.zeros((b * k, b * k))
        for a in range(b):
            P += (nrm[a] * nrm[a] * cp + (1. - nrm[a] * nrm[a]) * cs) * _mass_matrix(U[a], U[a])
            for c in range(b):
                if c != a:
                    P += (nrm[a] * nrm[c] * (cp - cs)) * _mass_matrix(U[a], U[c])
        P /= {2: 6., 3: 12., 4: 20.}[k]
        dofs = (b * face[:, None] + np.arange(b)[None, :]).ravel()
        rows.append(np.repeat(dofs, b * k))
        cols.append(np.tile(dofs, b * k))
        vals.append((face_measure(mesh, face) * P).ravel())
    B = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(b * mesh.nb_node, b * mesh.nb_node)).tocsr()
    B.sum_duplicates()
    return B


def soildynamics_time_loop(case, nb_dof, step):
    """modules/soildynamics/FemModule.cc:28-62, 75-78, 262-290: t starts at dt, the step that starts with t >= tmax is the last (and the one
    compared with the golden file).  step(U, V, A, t) -> displacement of the step."""
    k = soildynamics_coefficients(case)
    gamma, beta, dt = k["gamma"], k["beta"], case["dt"]
    t = dt
    U, V, A = np.zeros(nb_dof), np.zeros(nb_dof), np.zeros(nb_dof)
    while True:
        last = t >= case["tmax"]
        dU = step(U, V, A, t)
        a_new = (dU - U - dt * V) / (beta * dt * dt) - (1. - 2. * beta) / (2. * beta) * A
        V = V + dt * ((1. - gamma) * A + gamma * a_new)
        A = a_new
        U = dU
        t += dt
        if last:
            return dU


# modules/bilaplacian/inputs/direct.arc
BILAPLACIAN_CASE = dict(mesh="bilap.msh", f=-786.25, dirichlet=[("boundary", [145.5, None])], penalty=1.0e30,
                        golden="bilaplacian_2d_test.txt")


def boundary_faces(mesh, case, group):
    """the group's faces in the node order the case's reference run saw"""
    from arcanefem_b200 import mesh as M
    if case.get("face_order") == "arcane":
        return M.arcane_face_node_order(mesh, mesh.faces[group])
    return M.orient_boundary_faces(mesh, mesh.faces[group])


def load_golden(name, ncomp):
    """'uid v0 [v1 ...]' per line -> dict uid -> values[:ncomp]"""
    out = {}
    with open(os.path.join(GOLDEN, name)) as f:
        for line in f:
            p = line.split()
            if not p:
                continue
            out[int(p[0])] = np.array([float(x) for x in p[1:1 + ncomp]])
    return out


def dirichlet_dofs(mesh, dirichlet, b):
    """.arc order; later groups override earlier ones (m_u is overwritten in order,
    modules/testlab/FemModule.cc:647-677).  Returns (dof_ids, values) sorted by dof."""
    val = {}
    for name, v in dirichlet:
        if callable(v):  # a value per node (far-field condition of the aerodynamics module)
            for node in mesh.groups[name]:
                val[int(node) * b] = float(v(mesh.coords[node]))
            continue
        vs = [v] if b == 1 and not isinstance(v, (list, tuple)) else list(v)
        for node in mesh.groups[name]:
            for k, x in enumerate(vs):
                if x is not None:
                    val[int(node) * b + k] = float(x)
    ids = np.array(sorted(val), dtype=np.int32)
    return ids, np.array([val[i] for i in ids], dtype=np.float64)


def compare_to_golden(mesh, u, golden, b, eps, min_value, subset=False):
    """femutils/FemUtils.cc:108-172 (checkNodeResultFile): relative eps, values below
    min_value skipped.  Returns max relative deviation over compared entries.
    subset: the file lists only some of the nodes (the production modules' check files hold the first few dozen)."""
    worst = 0.0
    assert np.all(np.isfinite(u))
    seen = 0
    for lid in range(mesh.nb_node):
        uid = int(mesh.node_uid[lid])
        if subset and uid not in golden:
            continue
        seen += 1
        ref = golden[uid]
        for k in range(b):
            r, v = ref[k], u[lid * b + k]
            if abs(r) < min_value and abs(v) < min_value:
                continue
            d = abs(r - v) / max(abs(r), abs(v))
            worst = max(worst, d)
    assert worst <= eps, f"max relative deviation {worst} > {eps}"
    assert seen == (len(golden) if subset else mesh.nb_node)
    return worst
