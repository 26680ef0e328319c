"""Copies the reference's own fixtures for the assembly path into tests/golden/
(meshes + golden post-solve nodal fields).  These are DATA files of the reference's
test-suite, not source; they travel with the repo because
/root/reference does not exist on the GPU box.  Run:  python tests/golden/make_golden.py
"""
import os
import shutil

REF = os.environ.get("AFB_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    # meshes (Gmsh 4.1 binary)
    "meshes/msh/L-shape.msh": "L-shape.msh",
    "meshes/msh/L-shape-3D.msh": "L-shape-3D.msh",
    "meshes/msh/circle_cut.msh": "circle_cut.msh",
    "meshes/msh/sphere_cut.msh": "sphere_cut.msh",
    "meshes/msh/porous-medium.msh": "porous-medium.msh",
    "meshes/msh/bar.msh": "bar.msh",
    "meshes/msh/bar_dynamic_3D.msh": "bar_dynamic_3D.msh",
    "meshes/msh/bilap.msh": "bilap.msh",
    # golden nodal solutions ("uid value..." per line)
    "modules/testlab/tests/poisson_test_ref_L-shape_2D.txt": "poisson_test_ref_L-shape_2D.txt",
    "modules/testlab/tests/poisson_test_ref_L-shape_3D.txt": "poisson_test_ref_L-shape_3D.txt",
    "modules/testlab/tests/poisson_test_ref_circle_2D.txt": "poisson_test_ref_circle_2D.txt",
    "modules/testlab/tests/poisson_test_ref_sphere_3D.txt": "poisson_test_ref_sphere_3D.txt",
    "modules/testlab/tests/poisson_test_ref_circle_trac_2D.txt": "poisson_test_ref_circle_trac_2D.txt",
    "modules/testlab/tests/poisson_test_ref_circle_x-trac_2D.txt": "poisson_test_ref_circle_x-trac_2D.txt",
    "modules/testlab/tests/poisson_test_ref_circle_y-trac_2D.txt": "poisson_test_ref_circle_y-trac_2D.txt",
    "modules/testlab/tests/poisson_test_ref_circle_vect-trac_2D.txt": "poisson_test_ref_circle_vect-trac_2D.txt",
    "modules/elasticity/check/bar.2D.Dirichlet.traction.txt": "elasticity_bar.2D.Dirichlet.traction.txt",
    "modules/elasticity/check/bar.2D.Dirichlet.bodyForce.txt": "elasticity_bar.2D.Dirichlet.bodyForce.txt",
    "modules/elasticity/check/bar.3D.Dirichlet.bodyForce.txt": "elasticity_bar.3D.Dirichlet.bodyForce.txt",
    "modules/bilaplacian/check/2d_test.txt": "bilaplacian_2d_test.txt",
    # Quad4 / Hexa8 Poisson (modules/poisson: inputs/circle.2D.quad.arc, circle.neumann.2D.quad.arc with value=2, sphere.3D.hexa.arc)
    "meshes/msh/circle_cut.quad.msh": "circle_cut.quad.msh",
    "meshes/msh/sphere_cut.hexa.msh": "sphere_cut.hexa.msh",
    "modules/poisson/check/poisson_test_ref_circle_2D_quad.txt": "poisson_test_ref_circle_2D_quad.txt",
    "modules/poisson/check/poisson_test_ref_circle_scalar_neumann_2D_quad.txt": "poisson_test_ref_circle_scalar_neumann_2D_quad.txt",
    "modules/poisson/check/poisson_test_ref_sphere_3D_hexa.txt": "poisson_test_ref_sphere_3D_hexa.txt",
    # production poisson module on P1 cells
    "modules/poisson/check/poisson_test_ref_circle_2D.txt": "poissonmod_test_ref_circle_2D.txt",
    "modules/poisson/check/poisson_test_ref_circle_neumann_2D.txt": "poissonmod_test_ref_circle_neumann_2D.txt",
    "modules/poisson/check/poisson_test_ref_sphere_3D.txt": "poissonmod_test_ref_sphere_3D.txt",
    "modules/poisson/check/poisson_test_ref_sphere_neumann_3D.txt": "poissonmod_test_ref_sphere_neumann_3D.txt",
    # point Dirichlet conditions
    "meshes/msh/plancher.msh": "plancher.msh",
    "modules/poisson/check/poisson_test_point_dirichlet_2D.txt": "poissonmod_test_point_dirichlet_2D.txt",
    "modules/laplace/check/test3_results.txt": "laplace_test3_results.txt",
    "modules/elasticity/check/bar.2D.PointDirichlet.Dirichlet.bodyForce.txt": "elasticity_bar.2D.PointDirichlet.Dirichlet.bodyForce.txt",
    # the same operator on other meshes / boundary data (poisson cube, laplace truncated cube, electrostatics module)
    "meshes/msh/3x3x3_cube_hexa8.msh": "3x3x3_cube_hexa8.msh",
    "meshes/msh/truncated_cube.msh": "truncated_cube.msh",
    "meshes/msh/box-rods.msh": "box-rods.msh",
    "meshes/msh/box-rods.quad.msh": "box-rods.quad.msh",
    "meshes/msh/box-rod-circle.msh": "box-rod-circle.msh",
    "modules/poisson/check/poisson_test_ref_cube_3D_hexa8.txt": "poissonmod_test_ref_cube_3D_hexa8.txt",
    "modules/laplace/check/test_3D_truncated-cube.txt": "laplace_test_3D_truncated-cube.txt",
    "modules/electrostatics/check/test_1.txt": "electrostatics_test_1.txt",
    "modules/electrostatics/check/test_2.txt": "electrostatics_test_2.txt",
    "modules/electrostatics/check/box-rods.quad.txt": "electrostatics_box-rods.quad.txt",
    "modules/electrostatics/check/truncated-cube.hexa.txt": "electrostatics_truncated-cube.hexa.txt",
    # acoustics module (stiffness + mass)
    "meshes/msh/sub.msh": "sub.msh",
    "meshes/msh/sub.quad.msh": "sub.quad.msh",
    "meshes/msh/sub_3d.msh": "sub_3d.msh",
    "meshes/msh/sphere_in_sphere.hexa.msh": "sphere_in_sphere.hexa.msh",
    "modules/acoustics/check/sub_2D.txt": "acoustics_sub_2D.txt",
    "modules/acoustics/check/sub_2D.quad.txt": "acoustics_sub_2D.quad.txt",
    "modules/acoustics/check/sphere_3d.txt": "acoustics_sphere_3d.txt",
    "modules/acoustics/check/sphere_3d.hexa.txt": "acoustics_sphere_3d.hexa.txt",
    # fourier module (conductivity per cell)
    "meshes/msh/plancher.quad4.msh": "plancher.quad4.msh",
    "meshes/msh/multi-material.msh": "multi-material.msh",
    "meshes/msh/multi-material.quad.msh": "multi-material.quad.msh",
    "modules/fourier/check/test1_results.txt": "fourier_test1_results.txt",
    "modules/fourier/check/test2_results.txt": "fourier_test2_results.txt",
    "modules/fourier/check/conduction_quad.txt": "fourier_conduction_quad.txt",
    "modules/fourier/check/conduction_multi-mat_quad.txt": "fourier_conduction_multi-mat_quad.txt",
    "modules/fourier/check/test_conduction_3D.txt": "fourier_test_conduction_3D.txt",
    "modules/fourier/check/conduction_hexa.txt": "fourier_conduction_hexa.txt",
    # Laplace module (Quad4, Hexa8, Tet4 through the BSR back-ends)
    "meshes/msh/ring.quad.msh": "ring.quad.msh",
    "modules/laplace/check/test_ring_quad.txt": "laplace_test_ring_quad.txt",
    "modules/laplace/check/test_trucated-cube_hexa.txt": "laplace_test_trucated-cube_hexa.txt",
    "modules/laplace/check/test_3D_L-shape.txt": "laplace_test_3D_L-shape.txt",
    # Quad4 / Hexa8 elasticity (modules/elasticity)
    "meshes/msh/five_quads.msh": "five_quads.msh",
    "meshes/msh/plate.quad.msh": "plate.quad.msh",
    "meshes/msh/truncated_cube.hexa.msh": "truncated_cube.hexa.msh",
    "modules/elasticity/check/2D.dirichlet.bodyforce.quad.txt": "elasticity_2D.dirichlet.bodyforce.quad.txt",
    "modules/elasticity/check/bar.2D.Dirichlet.bodyForce.quad.txt": "elasticity_bar.2D.Dirichlet.bodyForce.quad.txt",
    "modules/elasticity/check/3D.dirichlet.bodyforce.hexa.txt": "elasticity_3D.dirichlet.bodyforce.hexa.txt",
    "modules/elasticity/check/bar.2D.traction.bodyforce.txt": "elasticity_bar.2D.traction.bodyforce.txt",
    "modules/elasticity/check/2D.dirichlet.traction.bodyforce.quad.txt": "elasticity_2D.dirichlet.traction.bodyforce.quad.txt",
    "modules/elasticity/check/bar.2D.traction.bodyforce.quad.txt": "elasticity_bar.2D.traction.bodyforce.quad.txt",
    "modules/elasticity/check/3D.dirichlet.traction.bodyforce.hexa.txt": "elasticity_3D.dirichlet.traction.bodyforce.hexa.txt",
    "modules/poisson/check/poisson_test_ref_circle_neumann_2D_quad.txt": "poisson_test_ref_circle_neumann_2D_quad.txt",
    "modules/poisson/check/poisson_test_ref_sphere_neumann_3D_hexa.txt": "poisson_test_ref_sphere_neumann_3D_hexa.txt",
    "modules/poisson/check/poisson_test_ref_sphere_scalar_neumann_3D_hexa.txt": "poisson_test_ref_sphere_scalar_neumann_3D_hexa.txt",
    # elastodynamics module (Newmark-beta time loop on the stiffness + mass operator): inputs/bar.arc, inputs/bar.3D.arc
    "meshes/msh/bar_dynamic.msh": "bar_dynamic.msh",
    "modules/elastodynamics/check/2D_elastodynamics_bar.txt": "elastodynamics_2D_bar.txt",
    "modules/elastodynamics/check/bar_3d.txt": "elastodynamics_bar_3d.txt",
    # aerodynamics module (potential flow: the Poisson matrix with a far-field penalty condition psi = y - angle * x (z - angle * x in 3-D),
    # modules/aerodynamics/FemModule.cc:236-252; inputs/Joukowski.arc, Joukowski.quad.arc, Joukowski_3d.arc, Joukowski_3d.hexa.arc)
    "meshes/msh/NACA0012.msh": "NACA0012.msh",
    "meshes/msh/NACA0012.quad.msh": "NACA0012.quad.msh",
    "meshes/msh/aerodynamics_3d_coarse.msh": "aerodynamics_3d_coarse.msh",
    "meshes/msh/aerodynamics_3d_coarse.hexa.msh": "aerodynamics_3d_coarse.hexa.msh",
    "modules/aerodynamics/check/test_2d.txt": "aerodynamics_test_2d.txt",
    "modules/aerodynamics/check/test_2d.quad.txt": "aerodynamics_test_2d.quad.txt",
    "modules/aerodynamics/check/test_3d.txt": "aerodynamics_test_3d.txt",
    "modules/aerodynamics/check/test_3d.hexa.txt": "aerodynamics_test_3d.hexa.txt",
    # fouriernl module (Picard iterations on lambda(u) = (1 + u)^m, lambda taken at the cell mean on Tri3 / Tet4: modules/fouriernl/ElementMatrix.h:29-41;
    # inputs/Test.nonlinear.conduction.arc, Test.3d.nonlinear.conduction.arc)
    "meshes/msh/unit_square.msh": "unit_square.msh",
    "meshes/msh/unit_cube.msh": "unit_cube.msh",
    "modules/fouriernl/check/conduction_tria.txt": "fouriernl_conduction_tria.txt",
    "modules/fouriernl/check/conduction_tetra.txt": "fouriernl_conduction_tetra.txt",
    # soildynamics module (the elastodynamics matrix + the paraxial absorbing boundary, modules/soildynamics/Paraxial.h; inputs/constant-traction.arc,
    # constant-traction.pointbc.arc)
    "meshes/msh/semi-circle-soil.msh": "semi-circle-soil.msh",
    "modules/soildynamics/check/test_2D_constant_traction.txt": "soildynamics_test_2D_constant_traction.txt",
    "modules/soildynamics/check/test_2D_constant_traction_pointbc.txt": "soildynamics_test_2D_constant_traction_pointbc.txt",
    # ... and its double-couple source (a force table in time applied at four named nodes, modules/soildynamics/DoubleCouple.h) inside a box of
    # paraxial boundaries: inputs/double-couple.paraxial.arc, double-couple.paraxial.body-force.arc
    "meshes/msh/square_double-couple.msh": "square_double-couple.msh",
    "modules/soildynamics/data/force_loading_dc.txt": "soildynamics_force_loading_dc.txt",
    "modules/soildynamics/check/test_paraxial_results.txt": "soildynamics_test_paraxial_results.txt",
    "modules/soildynamics/check/test_paraxial_body-force_results.txt": "soildynamics_test_paraxial_body-force_results.txt",
    # ... in 3-D (paraxial triangles / quadrilaterals, modules/soildynamics/Paraxial.h:43-90): inputs/3d.double-couple.paraxial.soil.arc
    "meshes/msh/cube_double_couple_3d.msh": "cube_double_couple_3d.msh",
    "modules/soildynamics/check/3d_test_paraxial_double_couple.txt": "soildynamics_3d_test_paraxial_double_couple.txt",
    # ... and a traction table in time: inputs/transient-traction.arc
    "modules/soildynamics/data/semi-circle-soil-traction.txt": "soildynamics_semi-circle-soil-traction.txt",
    "modules/soildynamics/check/test_2D_transient_traction.txt": "soildynamics_test_2D_transient_traction.txt",
    # elastodynamics module, more boundary data on the same operator: inputs/bar.dirichlet-traction.quad.arc, bar.3D.dirichlet-traction.hexa.arc,
    # bar.dirichlet.traction.bodyforce.quad.arc, semi-circle.pointBC.arc, truncated-cube.pointBC.arc
    "meshes/msh/semi-circle.msh": "semi-circle.msh",
    "modules/elastodynamics/check/bar_2d_dirichlet-traction.quad.txt": "elastodynamics_bar_2d_dirichlet-traction.quad.txt",
    "modules/elastodynamics/check/bar_3d_dirichlet-traction.hexa.txt": "elastodynamics_bar_3d_dirichlet-traction.hexa.txt",
    "modules/elastodynamics/check/bar_dirichlet_traction_bodyforce.quad.txt": "elastodynamics_bar_dirichlet_traction_bodyforce.quad.txt",
    "modules/elastodynamics/check/semi-ciricle_point-bc.txt": "elastodynamics_semi-ciricle_point-bc.txt",
    "modules/elastodynamics/check/truncated-cube_point-bc.txt": "elastodynamics_truncated-cube_point-bc.txt",
    # ... and a traction read from a table in time (femutils/ArcaneFemFunctions.h:2959-2997): inputs/bar.transient-traction.arc, bar.transient-traction.quad.arc,
    # bar.3D.transient-traction.arc, bar.3D.transient-traction.hexa.arc
    "modules/elastodynamics/data/traction_bar_test_1.txt": "elastodynamics_traction_bar_test_1.txt",
    "modules/elastodynamics/data/traction_bar_three_steps.txt": "elastodynamics_traction_bar_three_steps.txt",
    # ... Rayleigh damping (etam, etak: stiffness-type terms on the right-hand side, modules/elastodynamics/SourceTerm.h:69-85): inputs/bar.damping.arc
    "modules/elastodynamics/check/2D_elastodynamics_bar_constant_traction_damping.txt": "elastodynamics_2D_bar_constant_traction_damping.txt",
    # ... generalized-alpha time discretization (inputs/bar.Galpha.arc)
    "modules/elastodynamics/check/2D_elastodynamics_Galpha_time_discretization.txt": "elastodynamics_2D_Galpha_time_discretization.txt",
    "modules/elastodynamics/check/2D_elastodynamics_bar_transient_traction_three_steps.txt": "elastodynamics_2D_bar_transient_traction_three_steps.txt",
    "modules/elastodynamics/check/2D_elastodynamics_bar_transient_traction.txt": "elastodynamics_2D_bar_transient_traction.txt",
    "modules/elastodynamics/check/bar_transient-traction.quad.txt": "elastodynamics_bar_transient-traction.quad.txt",
    "modules/elastodynamics/check/bar_3d_transient-traction.txt": "elastodynamics_bar_3d_transient-traction.txt",
    "modules/elastodynamics/check/bar_3d_transient-traction.hexa.txt": "elastodynamics_bar_3d_transient-traction.hexa.txt",
    # elasticity on Arcane's cartesian generator (Quad4 / Hexa8 bars built in tests/cases.py::cartesian_mesh): inputs/bar.2D.cartesian.Dirichlet.bodyForce.arc,
    # bar.3D.cartesian.Dirichlet.bodyForce.arc
    "modules/elasticity/check/bar.2D.cartesian.Dirichlet.bodyForce.txt": "elasticity_bar.2D.cartesian.Dirichlet.bodyForce.txt",
    "modules/elasticity/check/bar.3D.cartesian.Dirichlet.bodyForce.txt": "elasticity_bar.3D.cartesian.Dirichlet.bodyForce.txt",
    # heat module (implicit Euler on lambda * stiffness + mass / dt): inputs/conduction.arc, 3d_conduction.arc, conduction.quad.arc
    "meshes/msh/plate.msh": "plate.msh",
    "modules/heat/check/2d_conduction.txt": "heat_2d_conduction.txt",
    "modules/heat/check/3d_conduction.txt": "heat_3d_conduction.txt",
    "modules/heat/check/2d_conduction.quad.txt": "heat_2d_conduction.quad.txt",
    # ... with convection (Robin) boundaries, flux boundaries and point conditions: inputs/conduction.convection.arc, conduction.convection.quad.arc,
    # conduction.neumann.point-dirichlet.arc, conduction.neumann.point-dirichlet.quad.arc, 3d_conduction.pointBc.convection.arc,
    # 3d_conduction.convection.hexa.arc, 3d_conduction.neumann.hexa.arc
    "modules/heat/check/2d_conduction_convection.txt": "heat_2d_conduction_convection.txt",
    "modules/heat/check/2d_conduction_convection.quad.txt": "heat_2d_conduction_convection.quad.txt",
    "modules/heat/check/2d_conduction_neumann_pointBC.txt": "heat_2d_conduction_neumann_pointBC.txt",
    "modules/heat/check/2d_conduction_neumann_pointBC.quad.txt": "heat_2d_conduction_neumann_pointBC.quad.txt",
    "modules/heat/check/3d_conduction_convection_pointBC.txt": "heat_3d_conduction_convection_pointBC.txt",
    "modules/heat/check/3d_conduction_convection.hexa.txt": "heat_3d_conduction_convection.hexa.txt",
    "modules/heat/check/3d_conduction_neumann.hexa.txt": "heat_3d_conduction_neumann.hexa.txt",
    # the same on Quad4 / Hexa8: inputs/bar.quad.arc, bar.3D.hexa.arc
    "meshes/msh/bar_dynamic_quad.msh": "bar_dynamic_quad.msh",
    "meshes/msh/bar_dynamic_3Dhexa.msh": "bar_dynamic_3Dhexa.msh",
    "modules/elastodynamics/check/bar.quad.txt": "elastodynamics_bar.quad.txt",
    "modules/elastodynamics/check/bar_3d.hexa.txt": "elastodynamics_bar_3d.hexa.txt",
}

if __name__ == "__main__":
    for src, dst in FILES.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, dst))
        os.chmod(os.path.join(HERE, dst), 0o644)
        print("copied", src, "->", dst)
