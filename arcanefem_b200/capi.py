"""ctypes binding of the C ABI (include/afb200.h) — what tests/ and bench.py call.

There is no CPU fallback: if libafb200.so is missing or no CUDA device is present the
calls raise.  Device arrays are returned as raw pointers; `Context.to_host` copies them
out and `Context.as_torch` wraps them as zero-copy torch tensors (torch is plumbing:
device memory, streams, torch.distributed).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AFB200_LIB") or os.path.join(_HERE, "libafb200.so")  # AFB200_LIB: tuning builds (scratch/build_variants.sh)

# enums of afb200.h
OP_POISSON, OP_ELASTICITY, OP_BILAPLACIAN, OP_DIFFUSION_REACTION, OP_ELASTODYNAMICS = 0, 1, 2, 3, 4
FORMAT_CSR, FORMAT_COO, FORMAT_BSR = 0, 1, 2
VARIANT_CELLWISE_ATOMIC, VARIANT_NODEWISE, VARIANT_TILED_GATHER = 0, 1, 2
LAYOUT_PER_BLOCK, LAYOUT_PER_ROW = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
FLAG_SIGNED_TRI_AREA, FLAG_OWN_CELLS_ONLY, FLAG_ALL_ROWS = 1, 2, 4
SPARSITY_AUTO, SPARSITY_FROM_CELLS, SPARSITY_FROM_CONNECTIVITY = 0, 1, 2
TILED_EXEC_BRICKS, TILED_EXEC_CHAIN, TILED_EXEC_CHAIN_FLOW = 0, 1, 2
VEC_EXEC_AUTO, VEC_EXEC_ROWS, VEC_EXEC_UNITS = 0, 1, 2
NEUMANN_FLUX, NEUMANN_TRACTION = 0, 1
ELIMINATE_ROW, ELIMINATE_ROW_COLUMN = 1, 2
(ARRAY_ROWS, ARRAY_COLUMNS, ARRAY_VALUES, ARRAY_NZ_PER_ROW, ARRAY_RHS, ARRAY_COO_ROWS, ARRAY_CSR_ROWS, ARRAY_CSR_COLUMNS,
 ARRAY_CSR_NB_COLUMN, ARRAY_COORDS, ARRAY_CELL_NODES, ARRAY_NODE_CELL_PTR, ARRAY_NODE_CELL_LIST) = range(13)

_ARRAY_DTYPE = {ARRAY_ROWS: np.int32, ARRAY_COLUMNS: np.int32, ARRAY_VALUES: np.float64, ARRAY_NZ_PER_ROW: np.int32, ARRAY_RHS: np.float64,
                ARRAY_COO_ROWS: np.int32, ARRAY_CSR_ROWS: np.int32, ARRAY_CSR_COLUMNS: np.int32, ARRAY_CSR_NB_COLUMN: np.int32,
                ARRAY_COORDS: np.float64, ARRAY_CELL_NODES: np.int32, ARRAY_NODE_CELL_PTR: np.int32, ARRAY_NODE_CELL_LIST: np.int32}

MSH_GROUP_POINTS, MSH_GROUP_FACES, MSH_GROUP_CELLS = 0, 1, 2   # afb_msh_group kinds

EXPORTS = [
    "afb_create", "afb_destroy", "afb_last_error", "afb_version", "afb_set_stream", "afb_synchronize", "afb_set_mesh", "afb_update_coordinates", "afb_set_own_cell_count", "afb_set_cell_coefficient", "afb_get_own_cell_count", "afb_renumber_columns", "afb_get_ij_arrays", "afb_memcpy_to_host", "afb_mesh_generate_box",
    "afb_build_pattern", "afb_set_sparsity_algorithm", "afb_set_tiled_executor", "afb_set_vector_executor", "afb_set_tiled_stage_limit", "afb_options_from_name", "afb_reset_values", "afb_assemble_bilinear", "afb_rhs_reset", "afb_assemble_rhs_source", "afb_assemble_rhs_neumann", "afb_set_dirichlet_nodes",
    "afb_dirichlet_penalty", "afb_set_elimination", "afb_set_forced_values", "afb_clear_dirichlet", "afb_apply_matrix_transformation",
    "afb_apply_rhs_transformation", "afb_matrix_get_value", "afb_matrix_set_value", "afb_get_csr_view", "afb_get_bsr", "afb_get_coo", "afb_get_rhs", "afb_get_mesh", "afb_copy_to_host",
    "afb_lookup_value_slots", "afb_add_values_at", "afb_values_tail",
    "afb_p2p_export", "afb_p2p_connect", "afb_p2p_exchange", "afb_p2p_exchange_async", "afb_p2p_wait", "afb_p2p_status", "afb_p2p_wait_stats", "afb_p2p_disconnect", "afb_partition_create", "afb_partition_destroy", "afb_partition_sizes", "afb_partition_get",
    "afb_msh_read", "afb_msh_destroy", "afb_msh_sizes", "afb_msh_get", "afb_msh_group", "afb_msh_group_get",
    "afb_xplan_host_create", "afb_xplan_host_destroy", "afb_xplan_host_peers", "afb_xplan_host_pairs", "afb_xplan_host_numbering",
    "afb_xplan_create", "afb_xplan_destroy", "afb_xplan_exchange", "afb_xplan_wait", "afb_xplan_numbering", "afb_xplan_info", "afb_solve_pcg", "afb_last_timings", "afb_inspector_timings", "afb_launch_count",
]


class AfbError(RuntimeError):
    pass


_lib = None


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into arcanefem_b200/libafb200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfbError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
        _lib = C.CDLL(LIB_PATH)
        _lib.afb_last_error.restype = C.c_char_p
        _lib.afb_version.restype = C.c_char_p
        _lib.afb_launch_count.restype = C.c_int64
        _lib.afb_launch_count.argtypes = [C.c_void_p]
    return _lib


def _check(rc):
    if rc != 0:
        raise AfbError(f"afb error {rc}: {lib().afb_last_error().decode()}")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    # torch tensor
    return C.c_void_p(a.data_ptr())


def options_from_name(name: str):
    """(format, variant, sparsity) of a reference matrix-format option name (Fem.axl): 'csr-gpu', 'nwcsr', 'bsr', 'AF-BSR', ..."""
    f, v, sp = C.c_int(), C.c_int(), C.c_int()
    _check(lib().afb_options_from_name(name.encode(), C.byref(f), C.byref(v), C.byref(sp)))
    return f.value, v.value, sp.value


class Context:
    """One assembly context per GPU (afb_ctx)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        _check(lib().afb_create(int(device), C.byref(self._h)))
        self.device = device
        if stream is not None:
            self.set_stream(stream)
        self.b = 1

    def close(self):
        if self._h:
            lib().afb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing ---------------------------------------------------------------------------
    def set_stream(self, stream_ptr: int | None):
        _check(lib().afb_set_stream(self._h, C.c_void_p(stream_ptr or 0)))

    def synchronize(self):
        _check(lib().afb_synchronize(self._h))

    # -- mesh -----------------------------------------------------------------------------------
    def set_mesh(self, dim, coords, cells, is_own=None, mem_space=MEM_HOST):
        if mem_space == MEM_HOST:
            coords = np.ascontiguousarray(coords, dtype=np.float64)
            cells = np.ascontiguousarray(cells, dtype=np.int32)
            is_own = None if is_own is None else np.ascontiguousarray(is_own, dtype=np.uint8)
        self._keep = (coords, cells, is_own)
        nb_node, npc = int(coords.shape[0]), int(cells.shape[1])
        _check(lib().afb_set_mesh(self._h, int(dim), npc, C.c_int32(nb_node), C.c_int64(int(cells.shape[0])), _ptr(coords), _ptr(cells), _ptr(is_own), mem_space))
        self.dim, self.npc, self.nb_node, self.nb_cell = dim, npc, nb_node, int(cells.shape[0])

    def update_coordinates(self, coords, mem_space=MEM_HOST):
        """New coordinates on the same topology: plans and sparsity structures stay valid (time loop)."""
        if mem_space == MEM_HOST:
            coords = np.ascontiguousarray(coords, dtype=np.float64)
        self._keep_xyz = coords
        _check(lib().afb_update_coordinates(self._h, _ptr(coords), mem_space))

    def set_own_cell_count(self, nb_own_cell):
        _check(lib().afb_set_own_cell_count(self._h, C.c_int64(int(nb_own_cell))))

    def generate_box(self, dim, n, jitter=0.2, seed=12345, k_lo=0, k_hi=None, ghost_cell_layer=False):
        k_hi = n if k_hi is None else k_hi
        _check(lib().afb_mesh_generate_box(self._h, dim, n, C.c_double(jitter), C.c_uint32(seed), k_lo, k_hi, int(ghost_cell_layer)))
        info = self.mesh_info()
        self.dim, self.npc, self.nb_node, self.nb_cell = info["dim"], info["npc"], info["nb_node"], info["nb_cell"]
        return info

    def mesh_info(self):
        dim, npc, nbn, nbo = C.c_int(), C.c_int(), C.c_int32(), C.c_int32()
        nbc = C.c_int64()
        xyz, cn, own = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _check(lib().afb_get_mesh(self._h, C.byref(dim), C.byref(npc), C.byref(nbn), C.byref(nbc), C.byref(nbo), C.byref(xyz), C.byref(cn), C.byref(own)))
        noc = C.c_int64()
        _check(lib().afb_get_own_cell_count(self._h, C.byref(noc)))
        return dict(dim=dim.value, npc=npc.value, nb_node=nbn.value, nb_cell=nbc.value, nb_own_node=nbo.value, nb_own_cell=noc.value,
                    xyz=xyz.value, cell_nodes=cn.value, is_own=own.value)

    # -- pattern / assembly ------------------------------------------------------------------
    def build_pattern(self, nb_dof_per_node=1):
        nbr, nnz = C.c_int32(), C.c_int64()
        _check(lib().afb_build_pattern(self._h, nb_dof_per_node, C.byref(nbr), C.byref(nnz)))
        self.b, self.nb_block_row, self.nnz = nb_dof_per_node, nbr.value, nnz.value
        return nbr.value, nnz.value

    def set_sparsity_algorithm(self, algorithm):
        """SPARSITY_FROM_CELLS = computeSparsityAtomic, SPARSITY_FROM_CONNECTIVITY = computeSparsityAtomicFree (re-builds on an unchanged mesh)."""
        _check(lib().afb_set_sparsity_algorithm(self._h, int(algorithm)))

    def set_cell_coefficient(self, coefficient):
        """per-cell multiplier of the Poisson element matrix ([nb_cell] doubles, or a scalar for all cells); None switches it off"""
        if coefficient is None:
            _check(lib().afb_set_cell_coefficient(self._h, None, MEM_HOST))
            return
        c = np.ascontiguousarray(np.broadcast_to(np.asarray(coefficient, dtype=np.float64), (self.nb_cell,)))
        _check(lib().afb_set_cell_coefficient(self._h, _ptr(c), MEM_HOST))

    def set_vector_executor(self, executor):
        """VEC_EXEC_AUTO (default: rows on Tet4, units on Tri3), VEC_EXEC_ROWS, VEC_EXEC_UNITS: how VARIANT_TILED_GATHER runs for elasticity."""
        _check(lib().afb_set_vector_executor(self._h, int(executor)))

    def set_tiled_executor(self, executor):
        """TILED_EXEC_BRICKS (default), TILED_EXEC_CHAIN, TILED_EXEC_CHAIN_FLOW: how VARIANT_TILED_GATHER runs for b = 1."""
        _check(lib().afb_set_tiled_executor(self._h, int(executor)))

    def set_tiled_stage_limit(self, nbytes):
        """Tiles whose plan record / contribution lists exceed nbytes read them from global memory (test knob)."""
        _check(lib().afb_set_tiled_stage_limit(self._h, C.c_int64(int(nbytes))))

    def reset_values(self):
        _check(lib().afb_reset_values(self._h))

    def assemble(self, op=OP_POISSON, params=None, fmt=FORMAT_CSR, variant=VARIANT_CELLWISE_ATOMIC, layout=LAYOUT_PER_BLOCK, flags=0):
        prm = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
        _check(lib().afb_assemble_bilinear(self._h, op, _ptr(prm), 0 if prm is None else int(prm.size), fmt, variant, layout, flags))
        self.layout = layout

    # -- rhs / dirichlet -----------------------------------------------------------------------
    def rhs_neumann(self, faces, values, kind=NEUMANN_FLUX, skip_dirichlet=False):
        """faces: int32 [nb_face, dim] oriented boundary faces (mesh.orient_boundary_faces), [nb_face, 4] on Hexa8 meshes
        (mesh.arcane_face_node_order for a flux vector); values: flux value, flux vector or traction."""
        faces = np.ascontiguousarray(faces, dtype=np.int32)
        values = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64)
        _check(lib().afb_assemble_rhs_neumann(self._h, C.c_int64(faces.shape[0]), _ptr(faces), int(kind), int(values.size), _ptr(values), int(skip_dirichlet), MEM_HOST))

    def rhs_reset(self):
        _check(lib().afb_rhs_reset(self._h))

    def rhs_source(self, f, nodewise=False, signed_tri_area=False):
        f = np.ascontiguousarray(np.atleast_1d(f), dtype=np.float64)
        _check(lib().afb_assemble_rhs_source(self._h, _ptr(f), int(f.size), int(nodewise), int(signed_tri_area)))

    def set_dirichlet_nodes(self, node_ids):
        ids = np.ascontiguousarray(node_ids, dtype=np.int32)
        _check(lib().afb_set_dirichlet_nodes(self._h, C.c_int32(ids.size), _ptr(ids), MEM_HOST))

    def dirichlet_penalty(self, dof_ids, g, penalty, weak=False):
        ids = np.ascontiguousarray(dof_ids, dtype=np.int32)
        g = np.ascontiguousarray(g, dtype=np.float64)
        _check(lib().afb_dirichlet_penalty(self._h, int(weak), C.c_double(penalty), C.c_int32(ids.size), _ptr(ids), _ptr(g), MEM_HOST))

    def set_elimination(self, kind, dof_ids, g):
        ids = np.ascontiguousarray(dof_ids, dtype=np.int32)
        g = np.ascontiguousarray(g, dtype=np.float64)
        _check(lib().afb_set_elimination(self._h, kind, C.c_int32(ids.size), _ptr(ids), _ptr(g), MEM_HOST))

    def set_forced_values(self, dof_ids, v):
        ids = np.ascontiguousarray(dof_ids, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float64)
        _check(lib().afb_set_forced_values(self._h, C.c_int32(ids.size), _ptr(ids), _ptr(v), MEM_HOST))

    def clear_dirichlet(self):
        _check(lib().afb_clear_dirichlet(self._h))

    def apply_matrix_transformation(self, replicate_column0_quirk=True):
        _check(lib().afb_apply_matrix_transformation(self._h, int(replicate_column0_quirk)))

    def apply_rhs_transformation(self):
        _check(lib().afb_apply_rhs_transformation(self._h))

    # -- views -------------------------------------------------------------------------------------
    def csr_view(self):
        rows, nbc, cols, vals = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nb_row, nnz = C.c_int32(), C.c_int64()
        _check(lib().afb_get_csr_view(self._h, C.byref(rows), C.byref(nbc), C.byref(cols), C.byref(vals), C.byref(nb_row), C.byref(nnz)))
        return dict(rows=rows.value, rows_nb_column=nbc.value, columns=cols.value, values=vals.value, nb_row=nb_row.value, nnz=nnz.value)

    def bsr_view(self):
        rows, cols, vals, nzr = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        nbr, b, lay = C.c_int32(), C.c_int(), C.c_int()
        nbc = C.c_int64()
        _check(lib().afb_get_bsr(self._h, C.byref(rows), C.byref(cols), C.byref(vals), C.byref(nzr), C.byref(nbr), C.byref(nbc), C.byref(b), C.byref(lay)))
        return dict(rows_index=rows.value, columns=cols.value, values=vals.value, nb_nz_per_row=nzr.value, nb_block_row=nbr.value, nb_col=nbc.value,
                    block_size=b.value, layout=lay.value)

    def coo_view(self):
        r, c, v = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nnz = C.c_int64()
        _check(lib().afb_get_coo(self._h, C.byref(r), C.byref(c), C.byref(v), C.byref(nnz)))
        return dict(rows=r.value, cols=c.value, values=v.value, nnz=nnz.value)

    def rhs_view(self):
        p = C.c_void_p()
        n = C.c_int32()
        _check(lib().afb_get_rhs(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def to_host(self, which, out=None):
        nbytes = C.c_size_t()
        _check(lib().afb_copy_to_host(self._h, which, None, C.byref(nbytes)))
        dt = np.dtype(_ARRAY_DTYPE[which])
        if out is None:
            out = np.empty(nbytes.value // dt.itemsize, dtype=dt)
        assert out.nbytes >= nbytes.value
        _check(lib().afb_copy_to_host(self._h, which, _ptr(out), C.byref(nbytes)))
        return out

    # -- multi-GPU helpers -------------------------------------------------------------------------
    def lookup_value_slots(self, n, dof_rows_ptr, dof_cols_ptr, slots_ptr):
        _check(lib().afb_lookup_value_slots(self._h, C.c_int64(n), _ptr(dof_rows_ptr), _ptr(dof_cols_ptr), _ptr(slots_ptr)))

    def add_values_at(self, n, slots_ptr, contrib_ptr):
        _check(lib().afb_add_values_at(self._h, C.c_int64(n), _ptr(slots_ptr), _ptr(contrib_ptr)))

    # ghost-row exchange over NVLink peer memory (p2p.cu)
    def p2p_export(self):
        vh, fh = C.create_string_buffer(64), C.create_string_buffer(64)
        _check(lib().afb_p2p_export(self._h, vh, fh))
        return vh.raw, fh.raw

    def p2p_connect(self, my_rank, peer_rank, values_handles, flags_handles, pull_first, pull_count, slots, send_first, send_count):
        """peer_rank: list of ranks; *_handles: list of 64-byte strings; slots: list of device int64 tensors (or None)."""
        n = len(peer_rank)
        pr = (C.c_int32 * max(n, 1))(*peer_rank)
        vh = C.create_string_buffer(b"".join(values_handles), max(64 * n, 1))
        fh = C.create_string_buffer(b"".join(flags_handles), max(64 * n, 1))
        i64 = lambda a: (C.c_int64 * max(n, 1))(*[int(x) for x in a])
        sl = (C.c_void_p * max(n, 1))(*[None if (t is None or t.numel() == 0) else t.data_ptr() for t in slots])
        self._p2p_keep = list(slots)
        _check(lib().afb_p2p_connect(self._h, int(my_rank), n, pr, vh, fh, i64(pull_first), i64(pull_count), sl, i64(send_first), i64(send_count)))

    def p2p_exchange(self, asynchronous=False):
        _check(lib().afb_p2p_exchange_async(self._h) if asynchronous else lib().afb_p2p_exchange(self._h))

    def p2p_wait(self):
        _check(lib().afb_p2p_wait(self._h))

    def p2p_wait_stats(self):
        """(ready_wait_us, pulled_wait_us, nb_exchange) since the last call: time the exchange kernels waited for the neighbours."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _check(lib().afb_p2p_wait_stats(self._h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def p2p_status(self):
        st = C.c_int()
        _check(lib().afb_p2p_status(self._h, C.byref(st)))
        return st.value

    def p2p_disconnect(self):
        _check(lib().afb_p2p_disconnect(self._h))

    def values_tail(self, first_block_row):
        first, nb = C.c_int64(), C.c_int64()
        _check(lib().afb_values_tail(self._h, C.c_int32(first_block_row), C.byref(first), C.byref(nb)))
        return first.value, nb.value

    # -- instrumentation -----------------------------------------------------------------------------
    def last_timings(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _check(lib().afb_last_timings(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(connectivity_ms=a.value, pattern_ms=b.value, assemble_ms=c.value)

    def solve_pcg(self, rtol=1e-12, atol=0.0, max_iter=10000):
        """Jacobi-PCG on the assembled system; returns (x, iterations, preconditioned residual)."""
        x = np.empty(self.nb_block_row * self.b, dtype=np.float64)
        it, res = C.c_int(), C.c_double()
        _check(lib().afb_solve_pcg(self._h, C.c_double(rtol), C.c_double(atol), int(max_iter), _ptr(x), MEM_HOST, C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def inspector_timings(self):
        a, b = C.c_float(), C.c_float()
        _check(lib().afb_inspector_timings(self._h, C.byref(a), C.byref(b)))
        return dict(mesh_tiling_ms=a.value, value_plan_ms=b.value)

    def launch_count(self):
        return int(lib().afb_launch_count(self._h))


def as_torch(ptr: int, shape, dtype, device: int):
    """Zero-copy torch view of a device array owned by the context."""
    import torch

    class _Holder:
        pass

    tstr = {np.dtype(np.float64): "<f8", np.dtype(np.int32): "<i4", np.dtype(np.int64): "<i8", np.dtype(np.uint8): "|u1"}[np.dtype(dtype)]
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(int(s) for s in np.atleast_1d(shape)), "typestr": tstr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(h, device=f"cuda:{device}")
