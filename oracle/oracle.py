"""ctypes loader for the CPU oracle (oracle/afb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Never imported by the arcanefem_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libafb_oracle.so")

OP_POISSON, OP_ELASTICITY, OP_BILAPLACIAN, OP_DIFFUSION_REACTION, OP_ELASTODYNAMICS = 0, 1, 2, 3, 4
FORM_COMPACT, FORM_HOST, FORM_BSR, FORM_NODEWISE = 0, 1, 2, 3
LAYOUT_PER_BLOCK, LAYOUT_PER_ROW = 0, 1

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "afb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_build_pattern.restype = C.c_int64
        _lib.orc_build_pattern_host.restype = C.c_int64
        _lib.orc_value_index.restype = C.c_int64
    return _lib


def _p(a, ty=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint8)


def block_size(op: int, dim: int) -> int:
    return 1 if op in (OP_POISSON, OP_DIFFUSION_REACTION) else (dim if op in (OP_ELASTICITY, OP_ELASTODYNAMICS) else 2)


def lame(E: float, nu: float):
    lam, mu = C.c_double(), C.c_double()
    lib().orc_lame(C.c_double(E), C.c_double(nu), C.byref(lam), C.byref(mu))
    return lam.value, mu.value


def element_matrix(npc, dim, op, form, coords, cell_nodes, params=None):
    b = block_size(op, dim)
    K = np.zeros((npc * b, npc * b))
    coords = _f64(coords)
    cn = _i32(cell_nodes)
    prm = _f64(params if params is not None else [0.0, 0.0, 0.0])
    rc = lib().orc_element_matrix(npc, dim, op, form, _p(prm), _p(coords), _p(cn), _p(K))
    assert rc == 0
    return K


def build_pattern(npc, nb_node, cells):
    cells = _i32(cells)
    nb_cell = cells.shape[0]
    rows = np.empty(nb_node + 1, dtype=np.int32)
    nnz = lib().orc_build_pattern(npc, C.c_int32(nb_node), C.c_int64(nb_cell), _p(cells), _p(rows), None)
    cols = np.empty(nnz, dtype=np.int32)
    lib().orc_build_pattern(npc, C.c_int32(nb_node), C.c_int64(nb_cell), _p(cells), _p(rows), _p(cols))
    return rows, cols


def build_pattern_host(npc, nb_node, cells, capacity):
    cells = _i32(cells)
    rows = np.empty(nb_node + 1, dtype=np.int32)
    cols = np.empty(capacity, dtype=np.int32)
    nnz = lib().orc_build_pattern_host(npc, C.c_int32(nb_node), C.c_int64(cells.shape[0]), _p(cells), _p(rows), _p(cols), C.c_int64(capacity))
    assert nnz >= 0
    return rows, cols[:nnz]


def assemble(mesh_dim, coords, cells, rows, cols, op=OP_POISSON, form=FORM_COMPACT, params=None,
             layout=LAYOUT_PER_BLOCK, nodewise=False, is_own=None, skip_zero=False, cell_coef=None):
    """cell_coef: per-cell multiplier of the element matrix (conductivity of the fourier / heat modules)"""
    coords, cells, rows, cols = _f64(coords), _i32(cells), _i32(rows), _i32(cols)
    if cell_coef is not None:
        cell_coef = _f64(np.broadcast_to(np.asarray(cell_coef, dtype=np.float64), (cells.shape[0],)))
        lib().orc_set_cell_coefficient(_p(cell_coef))
    try:
        return _assemble(mesh_dim, coords, cells, rows, cols, op, form, params, layout, nodewise, is_own, skip_zero)
    finally:
        lib().orc_set_cell_coefficient(None)


def _assemble(mesh_dim, coords, cells, rows, cols, op, form, params, layout, nodewise, is_own, skip_zero):
    npc = cells.shape[1]
    b = block_size(op, mesh_dim)
    vals = np.zeros(int(cols.shape[0]) * b * b)
    prm = _f64(params if params is not None else [0.0, 0.0])
    own = _u8(is_own)
    nb_node = coords.shape[0]
    if nodewise:
        rc = lib().orc_assemble_nodewise(npc, mesh_dim, op, form, _p(prm), C.c_int32(nb_node), C.c_int64(cells.shape[0]), _p(coords), _p(cells), _p(own),
                                         _p(rows), _p(cols), layout, _p(vals))
    else:
        rc = lib().orc_assemble_cellwise(npc, mesh_dim, op, form, _p(prm), C.c_int32(nb_node), C.c_int64(cells.shape[0]), _p(coords), _p(cells), _p(own),
                                         _p(rows), _p(cols), layout, int(skip_zero), _p(vals))
    assert rc == 0, rc
    return vals


def bsr_to_csr(b, rows, cols):
    rows, cols = _i32(rows), _i32(cols)
    nbr = rows.shape[0] - 1
    csr_rows = np.empty(nbr * b + 1, dtype=np.int32)
    csr_cols = np.empty(cols.shape[0] * b * b, dtype=np.int32)
    nbc = np.empty(nbr * b, dtype=np.int32)
    lib().orc_bsr_to_csr(C.c_int32(nbr), b, _p(rows), _p(cols), _p(csr_rows), _p(csr_cols), _p(nbc))
    return csr_rows, csr_cols, nbc


def csr_to_coo_rows(rows):
    rows = _i32(rows)
    out = np.empty(int(rows[-1]), dtype=np.int32)
    lib().orc_csr_to_coo_rows(C.c_int32(rows.shape[0] - 1), _p(rows), _p(out))
    return out


def value_index(rows, cols, b, layout, dof_row, dof_col):
    return lib().orc_value_index(_p(_i32(rows)), _p(_i32(cols)), b, layout, C.c_int32(dof_row), C.c_int32(dof_col))


def rhs_source_cellwise(mesh_dim, coords, cells, f, signed_area=False, is_own=None, is_dirichlet=None):
    coords, cells = _f64(coords), _i32(cells)
    f = _f64(np.atleast_1d(f))
    b = f.shape[0]
    rhs = np.zeros(coords.shape[0] * b)
    lib().orc_rhs_source_cellwise(cells.shape[1], mesh_dim, b, int(signed_area), C.c_int32(coords.shape[0]), C.c_int64(cells.shape[0]), _p(coords), _p(cells),
                                  _p(_u8(is_own)), _p(_u8(is_dirichlet)), _p(f), _p(rhs))
    return rhs


def rhs_source_nodewise(mesh_dim, coords, cells, f, is_own=None):
    coords, cells = _f64(coords), _i32(cells)
    f = _f64(np.atleast_1d(f))
    b = f.shape[0]
    rhs = np.zeros(coords.shape[0] * b)
    lib().orc_rhs_source_nodewise(cells.shape[1], mesh_dim, b, C.c_int32(coords.shape[0]), C.c_int64(cells.shape[0]), _p(coords), _p(cells), _p(_u8(is_own)), _p(f), _p(rhs))
    return rhs


NEUMANN_FLUX, NEUMANN_TRACTION = 0, 1


def rhs_neumann(mesh_dim, b, coords, faces, values, rhs, kind=NEUMANN_FLUX, is_own=None, is_dirichlet=None):
    """rhs += boundary integral over P1 faces (oriented: see orc_rhs_neumann); in place."""
    faces = np.ascontiguousarray(faces, dtype=np.int32)
    values = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64)
    if mesh_dim == 3 and faces.ndim == 2 and faces.shape[1] == 4:  # Quad4 faces of a Hexa8 mesh
        lib().orc_rhs_neumann_quad4(int(b), int(kind), int(values.size), C.c_int64(faces.shape[0]), _p(_f64(coords)), _p(faces), _p(values), _p(_u8(is_own)), _p(_u8(is_dirichlet)), _p(rhs))
        return rhs
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    assert faces.shape[1] == mesh_dim and rhs.dtype == np.float64
    lib().orc_rhs_neumann(int(mesh_dim), int(b), int(kind), int(values.size), C.c_int64(faces.shape[0]), _p(coords), _p(faces), _p(values),
                          _p(_u8(is_own)), _p(_u8(is_dirichlet)), _p(rhs))
    return rhs


def dirichlet_penalty(rows, cols, values, rhs, dof_ids, g, penalty, weak=False):
    dof_ids, g = _i32(dof_ids), _f64(g)
    rc = lib().orc_dirichlet_penalty(int(weak), C.c_double(penalty), C.c_int32(dof_ids.shape[0]), _p(dof_ids), _p(g), _p(_i32(rows)), _p(_i32(cols)), _p(values), _p(rhs))
    assert rc == 0


def apply_elimination(rows, cols, values, rhs, elim_info, elim_value, forced_info=None, forced_value=None, dof_is_own=None, quirk_skip_col0=True):
    rows, cols = _i32(rows), _i32(cols)
    lib().orc_apply_elimination(C.c_int32(rows.shape[0] - 1), _p(rows), _p(cols), _p(values), _p(rhs), _p(_u8(elim_info)), _p(_f64(elim_value)),
                                _p(_u8(forced_info)), _p(_f64(forced_value)), _p(_u8(dof_is_own)), int(quirk_skip_col0))


def spmv(rows, cols, values, x):
    rows, cols = _i32(rows), _i32(cols)
    y = np.empty(rows.shape[0] - 1)
    lib().orc_spmv(C.c_int32(rows.shape[0] - 1), _p(rows), _p(cols), _p(_f64(values)), _p(_f64(x)), _p(y))
    return y


def assemble_csr_host_range(mesh_dim, coords, cells, rows, cols, values, cell_lo, cell_hi, owner_lo, owner_hi):
    """Reference sequential CSR AddAndCompute on a sub-domain (thread-safe for disjoint owner ranges)."""
    rc = lib().orc_assemble_csr_host_range(cells.shape[1], mesh_dim, C.c_int64(cell_lo), C.c_int64(cell_hi), C.c_int32(owner_lo), C.c_int32(owner_hi),
                                           _p(coords), _p(cells), _p(rows), _p(cols), _p(values))
    assert rc == 0, rc


class ReferenceRank:
    """Init-time state of one MPI-rank-equivalent of the reference's sequential CSR back-end: the node-node
    connectivity of its owned nodes (built once, untimed, like FemModule::startInit)."""

    def __init__(self, cells, cell_lo, cell_hi, owner_lo, owner_hi):
        lib().orc_reference_init.restype = C.c_void_p
        self.args = (cell_lo, cell_hi, owner_lo, owner_hi)
        self.h = C.c_void_p(lib().orc_reference_init(cells.shape[1], C.c_int64(cell_lo), C.c_int64(cell_hi), C.c_int32(owner_lo), C.c_int32(owner_hi), _p(cells)))

    def close(self):
        if self.h:
            lib().orc_reference_free(self.h)
            self.h = None

    def __del__(self):
        self.close()


def reference_rank(mesh_dim, coords, cells, cell_lo, cell_hi, owner_lo, owner_hi, want_arrays=False, capacity=0, init=None):
    """One MPI-rank-equivalent of the reference's sequential CSR back-end (BuildMatrix + AddAndCompute)
    on the sub-domain cells [cell_lo,cell_hi) / owned nodes [owner_lo,owner_hi).  ctypes releases the
    GIL, so N python threads run N ranks concurrently.  `init` = ReferenceRank of the same sub-domain: BuildMatrix
    walks the init-time node-node connectivity as the reference does; without it the connectivity is built inside the
    timed BuildMatrix (first assembly).  Returns dict(nnz, checksum, seconds[, rows, cols, vals])."""
    chk = C.c_double()
    sec = (C.c_double * 2)()
    rows = cols = vals = None
    if want_arrays:
        rows = np.empty(owner_hi - owner_lo + 1, dtype=np.int32)
        cols = np.empty(capacity, dtype=np.int32)
        vals = np.empty(capacity, dtype=np.float64)
    tail = (C.c_int64(cell_lo), C.c_int64(cell_hi), C.c_int32(owner_lo), C.c_int32(owner_hi),
            _p(coords), _p(cells), C.byref(chk), sec, _p(rows), _p(cols), _p(vals), C.c_int64(capacity))
    if init is not None:
        assert init.args == (cell_lo, cell_hi, owner_lo, owner_hi)
        lib().orc_reference_rank_nn.restype = C.c_int64
        nnz = lib().orc_reference_rank_nn(init.h, cells.shape[1], mesh_dim, *tail)
    else:
        lib().orc_reference_rank.restype = C.c_int64
        nnz = lib().orc_reference_rank(cells.shape[1], mesh_dim, *tail)
    assert nnz >= 0, nnz
    out = dict(nnz=int(nnz), checksum=chk.value, seconds=(sec[0], sec[1]))
    if want_arrays:
        out.update(rows=rows, cols=cols[:nnz], vals=vals[:nnz])
    return out
