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


# ---------------------------------------------------------------------------------------------
# the domain-decomposition layer of the C ABI from C++ only (tests/cpp/mgpu_driver.cpp): afb_partition_* (RCB), per-rank
# contexts, afb_xplan_* with an in-process transport, the peer-memory exchange kernel between contexts of one process
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def mgpu_driver(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path_factory.mktemp("cpp") / "mgpu_driver")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mgpu_driver.cpp"), "-o", out,
           "-L" + LIBDIR, "-lafb200", "-Wl,-rpath," + LIBDIR, "-pthread"]
    subprocess.run(cmd, check=True)
    return out


def test_mgpu_driver_compiles_and_has_no_cpu_fallback(mgpu_driver, tmp_path):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    m = M.box_mesh(3, 2)
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([mgpu_driver, str(tmp_path / "m.bin"), "2", "1", "0", "1", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


def read_ranks(path, world):
    out = []
    with open(path, "rb") as f:
        for _ in range(world):
            arrs = []
            for dt in (np.int32, np.int32, np.float64, np.int64, np.int32, np.int64, np.int32):
                n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
                arrs.append(np.fromfile(f, dtype=dt, count=n))
            out.append(arrs)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [("box3d", 1, O.LAYOUT_PER_BLOCK), ("box3d", 3, O.LAYOUT_PER_ROW), ("sphere", 1, O.LAYOUT_PER_BLOCK), ("box3d", 3, O.LAYOUT_PER_BLOCK)],
                         ids=["box-poisson", "box-elasticity-per-row", "sphere-poisson", "box-elasticity-per-block"])
@pytest.mark.parametrize("p2p", [1, 0], ids=["peer-memory", "callback-transport"])
def test_mgpu_driver_two_rank_exchange_without_python(mgpu_driver, tmp_path, world, case, p2p):
    name, b, layout = case
    m = M.box_mesh(3, 7) if name == "box3d" else M.read_msh(os.path.join(ROOT, "tests", "golden", "sphere_cut.msh"))
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([mgpu_driver, str(tmp_path / "m.bin"), str(world), str(b), str(layout), str(p2p), str(tmp_path / "o.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    op = O.OP_POISSON if b == 1 else O.OP_ELASTICITY
    params = list(O.lame(21.0e5, 0.28)) if b > 1 else None
    grows, gcols = O.build_pattern(m.npc, m.nb_node, m.cells)
    gvals = O.assemble(m.dim, m.coords, m.cells, grows, gcols, op=op, form=O.FORM_BSR, params=params, layout=layout)
    bb = b * b
    seen = np.zeros(m.nb_node, dtype=np.int32)
    firsts = None
    node_number = {}
    for rank, (rows, cols, vals, gid, meta, first, l2g) in enumerate(read_ranks(tmp_path / "o.bin", world)):
        nb_own, kind = int(meta[0]), int(meta[1])
        assert kind == (1 if p2p else 2), "transport kind"
        seen[gid[:nb_own]] += 1
        firsts = first if firsts is None else firsts
        assert np.array_equal(first, firsts) and first[rank + 1] - first[rank] == nb_own * b
        assert np.array_equal(l2g[:nb_own * b], first[rank] + np.arange(nb_own * b))
        for i in range(nb_own):
            node_number[int(gid[i])] = int(l2g[i * b]) // b
        for i in range(nb_own):  # owned rows equal the global matrix rows (columns through the gid map)
            g = int(gid[i])
            lo, hi = int(rows[i]), int(rows[i + 1])
            glo, ghi = int(grows[g]), int(grows[g + 1])
            assert hi - lo == ghi - glo
            order = np.argsort(gid[cols[lo:hi]])
            assert np.array_equal(gid[cols[lo:hi]][order], gcols[glo:ghi])
            got = vals[lo * bb:hi * bb]
            want = gvals[glo * bb:ghi * bb]
            nz = hi - lo
            if b > 1 and layout == O.LAYOUT_PER_ROW:
                got = got.reshape(b, nz, b)[:, order, :].reshape(-1)
            else:
                got = got.reshape(nz, bb)[order].reshape(-1)
            scale = np.abs(want).max()
            assert np.abs(got - want).max() <= 1e-12 * scale, (rank, i)
        assert not vals[int(rows[nb_own]) * bb:].any(), "ghost rows are zero after the exchange"
    assert (seen == 1).all(), "every node has exactly one owner"
    # ghosts learnt their global number from the owner
    for rank, (rows, cols, vals, gid, meta, first, l2g) in enumerate(read_ranks(tmp_path / "o.bin", world)):
        for i in range(int(meta[0]), gid.size):
            assert int(l2g[i * b]) // b == node_number[int(gid[i])]


# ---------------------------------------------------------------------------------------------
# solver hand-off shim (include/arcanefem_b200/SolverHandoff.h) against recording stand-ins of HYPRE / PETSc
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def handoff_driver(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path_factory.mktemp("cpp") / "handoff_driver")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tests", "cpp"),
           os.path.join(ROOT, "tests", "cpp", "handoff_driver.cpp"), "-o", out, "-L" + LIBDIR, "-lafb200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return out


def test_handoff_shim_compiles_against_solver_signatures(handoff_driver, tmp_path):
    """CPU: the shim compiles with -Wall -Werror against the HYPRE IJ / PETSc COO signatures; without a GPU it fails loudly."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present: covered by the gpu tests")
    m = M.box_mesh(3, 2)
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([handoff_driver, str(tmp_path / "m.bin"), "0", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("first_row", [0, 1000])
def test_handoff_arrays_as_hypre_and_petsc_receive_them(handoff_driver, tmp_path, first_row):
    m = M.read_msh(os.path.join(ROOT, "tests", "golden", "sphere_cut.msh"))
    write_mesh(tmp_path / "m.bin", m)
    r = subprocess.run([handoff_driver, str(tmp_path / "m.bin"), str(first_row), str(tmp_path / "o.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, form=O.FORM_NODEWISE, nodewise=True)
    arrs = []
    with open(tmp_path / "o.bin", "rb") as f:
        for dt in (np.int32, np.int32, np.int32, np.int32, np.float64, np.int32, np.int32, np.int32, np.float64):
            n = int(np.fromfile(f, dtype=np.int64, count=1)[0])
            arrs.append(np.fromfile(f, dtype=dt, count=n))
    meta, ncols, hrows, hcols, hvals, pmeta, coo_i, coo_j, coo_v = arrs
    n = m.nb_node
    # IJMatrixCreate(comm, first, last, first, last); PARCSR; device memory; one SetValues for all rows; assembled; GetObject
    assert meta.tolist() == [first_row, first_row + n - 1, first_row, first_row + n - 1, 5555, 1, n, 1, 6, 1]
    assert np.array_equal(ncols, np.diff(rows)) and np.array_equal(hrows, first_row + np.arange(n)) and np.array_equal(hcols, cols)
    scale = np.abs(vals).max()
    assert np.abs(hvals - vals).max() <= 1e-12 * scale
    # MatSetPreallocationCOOLocal(nnz, coo_rows, coo_cols) + MatSetValuesCOO(values, INSERT_VALUES) + MatAssemblyBegin/End
    assert pmeta.tolist() == [cols.size, 1, 2]
    assert np.array_equal(coo_i, np.repeat(np.arange(n), np.diff(rows))) and np.array_equal(coo_j, cols) and np.array_equal(coo_v, hvals)


# ---- value types and CSR row iteration (include/arcanefem_b200/FemTypes.h, CsrRow / CsrRowColumnIndex) ------------------------

def test_value_types_and_row_iteration_on_the_host(tmp_path):
    """CPU: Real4 / RealMatrix / RealVector algebra and the CsrFormatMatrixView accessors (view over host arrays), -Wall -Werror."""
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path / "types_driver")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "types_driver.cpp"), "-o", out,
                    "-L" + LIBDIR, "-lafb200", "-Wl,-rpath," + LIBDIR], check=True)
    r = subprocess.run([out], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "types ok" in r.stdout, r.stdout + r.stderr


@pytest.fixture(scope="module")
def view_kernel(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path_factory.mktemp("cpp") / "view_kernel")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Xcompiler", "-Wall,-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "view_kernel.cu"), "-o", out, "-L" + LIBDIR, "-lafb200", "-Xlinker", "-rpath," + LIBDIR], check=True)
    return out


def test_user_kernel_over_the_device_view_compiles(view_kernel):
    """CPU: a user kernel written against the reference's view interface and value types compiles for sm_100a (__host__ __device__ members)."""
    assert os.path.exists(view_kernel)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [6, 24])
def test_user_kernel_over_the_device_view(view_kernel, n):
    """rowRange / value / tryFindColumnInRow on the device view of an assembled Poisson matrix: zero row sums, positive diagonals,
    every entry walked once; DoFLinearSystem::matrixGetValue / matrixAddValue / matrixSetValue agree with what the kernel read."""
    r = subprocess.run([view_kernel, str(n)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "view ok" in r.stdout, r.stdout + r.stderr


# ---- a host without Arcane and without Python: mesh file -> golden solution (tests/cpp/msh_driver.cpp over afb_msh_*) ------------

@pytest.fixture(scope="module")
def msh_driver(tmp_path_factory):
    if not os.path.exists(os.path.join(LIBDIR, "libafb200.so")):
        pytest.skip("libafb200.so not built")
    out = str(tmp_path_factory.mktemp("cpp") / "msh_driver")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "msh_driver.cpp"), "-o", out,
                    "-L" + LIBDIR, "-lafb200", "-Wl,-rpath," + LIBDIR], check=True)
    return out


def test_msh_driver_reads_the_reference_meshes_on_the_host(msh_driver):
    """CPU: the reader needs neither a GPU nor a context"""
    from tests import cases as CS
    r = subprocess.run([msh_driver, "info", os.path.join(CS.GOLDEN, "L-shape.msh")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mesh dim=2 npc=3 nodes=151 cells=254" in r.stdout and "group boundary kind=1" in r.stdout
    r = subprocess.run([msh_driver, "info", os.path.join(CS.GOLDEN, "missing.msh")], capture_output=True, text=True, timeout=60)
    assert r.returncode == 2 and "cannot open" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["L-shape_2D", "L-shape_3D", "sphere_3D"])
def test_msh_driver_file_to_golden_solution(msh_driver, name):
    """the testlab Poisson cases from the mesh file alone, in C++: afb_msh_* -> afb_set_mesh -> tiled assembly -> penalty -> PCG,
    compared inside the driver with the reference's golden solution file"""
    from tests import cases as CS
    case = CS.POISSON_CASES[name]
    args = [msh_driver, "solve", os.path.join(CS.GOLDEN, case["mesh"]), repr(case["f"]), repr(case["penalty"]), os.path.join(CS.GOLDEN, case["golden"])]
    for group, value in case["dirichlet"]:
        args += [group, repr(value)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "golden ok" in r.stdout, r.stdout + r.stderr
