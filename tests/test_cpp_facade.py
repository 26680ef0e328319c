"""C++ facade (include/arcanefem_b200/FemUtils.h) over the C ABI: compiles with g++ against libafb200.so,
fails loudly without a GPU (no CPU fallback), and -- on the GPU box -- reproduces the oracle through the
reference's call sequence (BSRFormat::initialize -> computeSparsity -> assembleBilinear -> toLinearSystem)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from arcanefem_b200 import mesh as M
from oracle import oracle as O

LIBDIR = os.path.join(ROOT, "arcanefem_b200")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path_factory.mktemp("cpp") / "facade_driver")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_driver.cpp"), "-o", out,
           "-L" + LIBDIR, "-lafb200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return out


def write_mesh(path, m):
    with open(path, "wb") as f:
        np.array([m.dim, m.npc, m.nb_node, m.nb_cell], dtype=np.int32).tofile(f)
        np.ascontiguousarray(m.coords, dtype=np.float64).tofile(f)
        np.ascontiguousarray(m.cells, dtype=np.int32).tofile(f)


def read_out(path):
    arrs = []
    with open(path, "rb") as f:
        for dt in (np.int32, np.int32, np.float64, np.int32):
            n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
            arrs.append(np.fromfile(f, dtype=dt, count=n))
    return arrs


def test_facade_compiles_and_has_no_cpu_fallback(driver, tmp_path):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    m = M.box_mesh(3, 2)
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([driver, str(tmp_path / "m.bin"), "csr-gpu", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["csr-gpu", "nwcsr", "coo-gpu", "bsr", "af-bsr", "af-bsr-csr", "elasticity-bsr", "elasticity-af-bsr-csr"])
@pytest.mark.parametrize("meshname", ["box3d", "L-shape.msh"])
def test_facade_matches_oracle(driver, tmp_path, mode, meshname):
    m = M.box_mesh(3, 6) if meshname == "box3d" else M.read_msh(os.path.join(ROOT, "tests", "golden", meshname))
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([driver, str(tmp_path / "m.bin"), mode, str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows, cols, vals, extra = read_out(tmp_path / "o.bin")
    rows_ref, cols_ref = O.build_pattern(m.npc, m.nb_node, m.cells)
    assert np.array_equal(rows, rows_ref) and np.array_equal(cols, cols_ref)
    elast = mode.startswith("elasticity")
    b = m.dim if elast else 1
    per_row = mode.endswith("-csr")
    layout = O.LAYOUT_PER_ROW if per_row else O.LAYOUT_PER_BLOCK
    if elast:
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=O.OP_ELASTICITY, form=O.FORM_BSR, params=list(O.lame(21.0e5, 0.28)), layout=layout)
    elif mode in ("csr-gpu", "nwcsr", "coo-gpu"):
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_COMPACT)
    else:
        ref = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_BSR)
    if mode in ("csr-gpu", "nwcsr"):
        assert vals[rows[0] + int(np.searchsorted(cols[rows[0]:rows[1]], 0))] == 1.0e30  # matrixSetValue penalty
        ref[rows[0] + int(np.searchsorted(cols[rows[0]:rows[1]], 0))] = 1.0e30
        assert extra[0] == cols.size
    if mode == "coo-gpu":
        assert np.array_equal(extra, O.csr_to_coo_rows(rows))
    bb = b * b
    seg = np.repeat(np.arange(m.nb_node), np.diff(rows) * bb)
    rowmax = np.zeros(m.nb_node)
    np.maximum.at(rowmax, seg, np.abs(ref))
    err = np.abs(vals - ref) / np.maximum(rowmax[seg], 1e-300)
    assert err.max() <= 1e-12
    if mode in ("bsr", "af-bsr", "af-bsr-csr", "elasticity-bsr", "elasticity-af-bsr-csr"):
        assert list(extra[:5]) == [m.nb_node, cols.size, cols.size * bb, b, 0 if per_row else 1]
        if per_row:
            assert list(extra[5:7]) == [m.nb_node * b, cols.size * bb]


@pytest.mark.gpu
@pytest.mark.parametrize("meshname", ["box3d", "L-shape.msh"])
def test_facade_solve(driver, tmp_path, meshname):
    """BoundaryConditions::applyConstantSourceToRhs / applyNeumannToRhs + DoFLinearSystem::solve through the facade."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    m = M.box_mesh(3, 6) if meshname == "box3d" else M.read_msh(os.path.join(ROOT, "tests", "golden", meshname))
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([driver, str(tmp_path / "m.bin"), "solve", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows, cols, sol, extra = read_out(tmp_path / "o.bin")
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_BSR)
    rhs = O.rhs_source_cellwise(m.dim, m.coords, m.cells, [1.0], signed_area=False)
    O.rhs_neumann(m.dim, 1, m.coords, m.cells[:1, :m.dim], [2.0], rhs)
    O.dirichlet_penalty(rows, cols, vals, rhs, np.array([0], dtype=np.int32), np.array([0.25]), 1.0e30)
    ref = spla.spsolve(sp.csr_matrix((vals, cols, rows)).tocsc(), rhs)
    assert 0 < extra[0] < 20000
    assert np.abs(sol - ref).max() <= 1e-8 * np.abs(ref).max()
