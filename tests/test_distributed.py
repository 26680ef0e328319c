"""Domain decomposition: partition helpers, ghost-row exchange plan and global numbering.

CPU part (gloo, world_size 2 and 3): the exchange plan runs on CPU tensors; the per-rank partial
matrices come from the oracle (test infrastructure), the slot lookup is a numpy restatement of
BSRMatrix::findValueIndex.  The distributed result must equal the global oracle matrix row by row.
GPU part (-m gpu): the same check with the CUDA kernels computing the partial rows, the ranks being
emulated one after the other on one device (the exchange itself is the CPU-tested code).
"""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from arcanefem_b200 import mesh as M
from oracle import oracle as O

TOL = 1e-12


def numpy_lookup(rows, cols, b, per_row):
    """(dof_row, dof_col) -> value slot, BSRMatrix::findValueIndex (femutils/BSRFormat.cc:79-106)."""
    def f(dr, dc):
        import torch
        out = np.full(dr.size, -1, dtype=np.int64)
        for k in range(dr.size):
            br, bc = int(dr[k]) // b, int(dc[k]) // b
            rb, re = int(rows[br]), int(rows[br + 1])
            p = rb + int(np.searchsorted(cols[rb:re], bc))
            if p < re and cols[p] == bc:
                i, j = int(dr[k]) - br * b, int(dc[k]) - bc * b
                out[k] = p * b * b + i * b + j if not per_row else rb * b * b + b * ((p - rb) + i * (re - rb)) + j
        return torch.from_numpy(out)
    return f


def expand_block_entries(rows_local, cols_local, run_len, b, layout_per_row):
    """Scalar (dof_row, dof_col) pairs of block entries in the memory order of the sender's values (BSRMatrix::findValueIndex
    layouts, femutils/BSRFormat.cc:79-106) -- numpy restatement of what afb_xplan_host_create does in C++ (checker only)."""
    rows_local = np.asarray(rows_local, dtype=np.int64)
    cols_local = np.asarray(cols_local, dtype=np.int64)
    if b == 1:
        return rows_local.astype(np.int32), cols_local.astype(np.int32)
    ii, jj = np.meshgrid(np.arange(b), np.arange(b), indexing="ij")
    if not layout_per_row:
        dr = (rows_local[:, None, None] * b + ii[None]).reshape(-1)
        dc = (cols_local[:, None, None] * b + jj[None]).reshape(-1)
        return dr.astype(np.int32), dc.astype(np.int32)
    out_r, out_c = [], []
    pos = 0
    for nz in run_len:
        r = rows_local[pos:pos + nz]
        c = cols_local[pos:pos + nz]
        dr = np.repeat(r[None, :, None] * b + np.arange(b)[:, None, None], b, axis=2)
        dc = np.broadcast_to(c[None, :, None] * b + np.arange(b)[None, None, :], (b, nz, b))
        out_r.append(dr.reshape(-1))
        out_c.append(dc.reshape(-1))
        pos += nz
    return np.concatenate(out_r).astype(np.int32), np.concatenate(out_c).astype(np.int32)


def global_reference(mesh, op, params, layout):
    rows, cols = O.build_pattern(mesh.npc, mesh.nb_node, mesh.cells)
    vals = O.assemble(mesh.dim, mesh.coords, mesh.cells, rows, cols, op=op, form=O.FORM_BSR, params=params, layout=layout)
    return rows, cols, vals


def block_row(rows, vals, r, b, per_row):
    """values of block row r as an array [nz, b, b] (both BSR value layouts)."""
    lb, le = int(rows[r]), int(rows[r + 1])
    nz, bb = le - lb, b * b
    seg = vals[lb * bb:le * bb]
    if not per_row:
        return seg.reshape(nz, b, b)
    return seg.reshape(b, nz, b).transpose(1, 0, 2)  # index = rb*bb + b*(x + i*nz) + j


def check_owned_rows(sub, b, layout, lrows, lcols, lvals, grows, gcols, gvals):
    """every owned block row equals the global one (columns matched through the global ids: local columns
    are ascending in LOCAL numbering, ghosts last); ghost rows are zero."""
    per_row = layout == O.LAYOUT_PER_ROW
    bb = b * b
    worst = 0.0
    for i in range(sub.nb_node):
        lb, le = int(lrows[i]), int(lrows[i + 1])
        if i >= sub.nb_own_node:
            assert not lvals[lb * bb:le * bb].any(), f"ghost row {i} is not zero after the exchange"
            continue
        g = int(sub.node_gid[i])
        gb, ge = int(grows[g]), int(grows[g + 1])
        assert le - lb == ge - gb, f"row {i} (global {g}): degree {le - lb} vs {ge - gb}"
        cg = sub.node_gid[lcols[lb:le]]
        order = np.argsort(cg)
        assert np.array_equal(cg[order], gcols[gb:ge]), "column sets differ"
        mine = block_row(lrows, lvals, i, b, per_row)[order]
        ref = block_row(grows, gvals, g, b, per_row)
        scale = max(np.abs(ref).max(), 1e-300)
        worst = max(worst, float(np.abs(mine - ref).max() / scale))
    assert worst <= TOL, f"worst row-scaled error {worst}"
    return worst


def run_rank(rank, world, port, case, errq):
    try:
        import torch
        import torch.distributed as dist
        from arcanefem_b200.distributed import ExchangePlan
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        name, op, layout = case
        mesh = M.box_mesh(3, 5) if name == "box3d" else (M.box_mesh(2, 9) if name == "box2d" else M.read_msh(os.path.join(ROOT, "tests", "golden", name)))
        b = O.block_size(op, mesh.dim)
        params = list(O.lame(21.0e5, 0.28)) if op == O.OP_ELASTICITY else None
        sub = M.partition_mesh(mesh, world)[rank]
        lrows, lcols = O.build_pattern(sub.npc, sub.nb_node, sub.cells)
        # mode "exchange": own cells only, rows of all local nodes
        lvals = O.assemble(sub.dim, sub.coords, sub.cells[:sub.nb_own_cell], lrows, lcols, op=op, form=O.FORM_BSR, params=params, layout=layout)
        vt = torch.from_numpy(lvals)
        own = sub.nb_own_node

        def add_at(slots, buf):
            vt.index_add_(0, slots, buf)

        plan = ExchangePlan(rank, world, sub.node_gid, sub.node_owner, own, b, layout == O.LAYOUT_PER_ROW,
                            tail_pattern=lambda: (lrows[own:], lcols[int(lrows[own]):]),
                            lookup=numpy_lookup(lrows, lcols, b, layout == O.LAYOUT_PER_ROW),
                            values_slice=lambda first, n: vt[first:first + n], add_at=add_at,
                            make_buffer=lambda n: torch.empty(n, dtype=torch.float64))
        plan.exchange()
        grows, gcols, gvals = global_reference(mesh, op, params, layout)
        check_owned_rows(sub, b, layout, lrows, lcols, lvals, grows, gcols, gvals)
        # the replicated scheme (ghost cells recomputed, owned rows only) gives the same owned rows
        rep = O.assemble(sub.dim, sub.coords, sub.cells, lrows, lcols, op=op, form=O.FORM_BSR, params=params, layout=layout, is_own=sub.is_own)
        check_owned_rows(sub, b, layout, lrows, lcols, rep, grows, gcols, gvals)
        # global numbering: owned rows contiguous per rank, ghosts agree with their owner
        num = plan.numbering()
        assert num.first_dof[rank + 1] - num.first_dof[rank] == own * b
        assert np.array_equal(num.dof_l2g[:own * b], num.first_dof[rank] + np.arange(own * b))
        gathered = [None] * world
        dist.all_gather_object(gathered, (sub.node_gid[:own], num.dof_l2g[:own * b:b] // b))
        node_of_gid = {}
        for gids, nums in gathered:
            node_of_gid.update(zip(gids.tolist(), nums.tolist()))
        assert len(node_of_gid) == mesh.nb_node and sorted(node_of_gid.values()) == list(range(mesh.nb_node))
        for i in range(own, sub.nb_node):
            assert num.dof_l2g[i * b] // b == node_of_gid[int(sub.node_gid[i])]
        sent, received = plan.bytes_per_exchange()
        assert sent > 0 or received > 0
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errq.put((rank, traceback.format_exc()))
        raise


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


CASES = [("box3d", O.OP_POISSON, O.LAYOUT_PER_BLOCK), ("box3d", O.OP_ELASTICITY, O.LAYOUT_PER_ROW), ("box3d", O.OP_ELASTICITY, O.LAYOUT_PER_BLOCK),
         ("sphere_cut.msh", O.OP_POISSON, O.LAYOUT_PER_BLOCK), ("L-shape.msh", O.OP_ELASTICITY, O.LAYOUT_PER_ROW)]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}-op{c[1]}-layout{c[2]}" for c in CASES])
def test_ghost_row_exchange_gloo(world, case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=run_rank, args=(r, world, port, case, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errs.append((-1, "rank timed out"))
    assert not errs, "\n".join(f"rank {r}:\n{t}" for r, t in errs)
    assert all(p.exitcode == 0 for p in procs)


def test_partition_covers_mesh_once():
    for mesh in (M.box_mesh(3, 6), M.read_msh(os.path.join(ROOT, "tests", "golden", "sphere_cut.msh"))):
        for world in (2, 3, 4):
            subs = M.partition_mesh(mesh, world)
            own_nodes = np.concatenate([s.node_gid[:s.nb_own_node] for s in subs])
            own_cells = np.concatenate([s.cell_gid[:s.nb_own_cell] for s in subs])
            assert np.array_equal(np.sort(own_nodes), np.arange(mesh.nb_node))
            assert np.array_equal(np.sort(own_cells), np.arange(mesh.nb_cell))
            for s in subs:
                assert (s.is_own[:s.nb_own_node] == 1).all() and (s.is_own[s.nb_own_node:] == 0).all()
                # one ghost layer: every cell touching an owned node is local
                touches = (np.isin(mesh.cells, s.node_gid[:s.nb_own_node])).any(axis=1)
                assert np.array_equal(np.sort(s.cell_gid), np.nonzero(touches | np.isin(np.arange(mesh.nb_cell), s.cell_gid[:s.nb_own_cell]))[0])
                assert np.array_equal(mesh.coords[s.node_gid], s.coords)
                assert np.array_equal(s.node_gid[s.cells], mesh.cells[s.cell_gid])


def test_box_slab_matches_generic_partition():
    for dim, n, world in ((3, 6, 3), (2, 8, 4)):
        for r in range(world):
            a = M.box_slab(dim, n, world, r)
            k_lo, k_hi = M.slab_layers(n, world, r)
            gid, owner_rel, nb_own, nb_own_cell, cell_gid = M.box_slab_numbering(dim, n, k_lo, k_hi, True)
            assert a.nb_own_node == nb_own and a.nb_own_cell == nb_own_cell
            assert (np.diff(a.node_owner[a.nb_own_node:]) >= 0).all()
            full = M.box_mesh(dim, n)
            assert np.array_equal(a.node_gid[a.cells], full.cells[a.cell_gid])


# -------------------------------------------------------------------------------------------------
# GPU: the CUDA kernels compute the partial rows; ranks emulated sequentially on one device
# -------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["atomic", "nodewise", "tiled"])
@pytest.mark.parametrize("case", [("box3d", O.OP_POISSON, O.LAYOUT_PER_BLOCK), ("box3d", O.OP_ELASTICITY, O.LAYOUT_PER_ROW), ("sphere_cut.msh", O.OP_POISSON, O.LAYOUT_PER_BLOCK)],
                         ids=["box-poisson", "box-elasticity-per-row", "sphere-poisson"])
def test_decomposed_assembly_gpu(case, variant):
    from arcanefem_b200 import capi as A
    name, op, layout = case
    mesh = M.box_mesh(3, 7) if name == "box3d" else M.read_msh(os.path.join(ROOT, "tests", "golden", name))
    world = 3
    b = O.block_size(op, mesh.dim)
    params = list(O.lame(21.0e5, 0.28)) if op == O.OP_ELASTICITY else None
    v = {"atomic": A.VARIANT_CELLWISE_ATOMIC, "nodewise": A.VARIANT_NODEWISE, "tiled": A.VARIANT_TILED_GATHER}[variant]
    fmt = A.FORMAT_BSR if b > 1 else A.FORMAT_CSR
    subs = M.partition_mesh(mesh, world)
    grows, gcols, gvals = global_reference(mesh, op, params, layout)
    ctxs, partial = [], []
    for s in subs:
        ctx = A.Context(0)
        ctx.set_mesh(s.dim, s.coords, s.cells, s.is_own)
        ctx.set_own_cell_count(s.nb_own_cell)
        ctx.build_pattern(b)
        # replicated scheme first: owned rows complete, ghost rows zero, no exchange
        ctx.assemble(op, params=params, fmt=fmt, variant=v, layout=layout)
        lrows, lcols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
        check_owned_rows(s, b, layout, lrows, lcols, ctx.to_host(A.ARRAY_VALUES), grows, gcols, gvals)
        # exchange scheme: own cells only, all rows
        ctx.reset_values()
        ctx.assemble(op, params=params, fmt=fmt, variant=v, layout=layout, flags=A.FLAG_OWN_CELLS_ONLY | A.FLAG_ALL_ROWS)
        ctxs.append(ctx)
        partial.append((lrows, lcols, ctx.to_host(A.ARRAY_VALUES)))
    # emulate the exchange through the C ABI (afb_lookup_value_slots / afb_add_values_at)
    import torch
    bb = b * b
    for s, (lrows, lcols, lvals) in zip(subs, partial):
        for i in range(s.nb_own_node, s.nb_node):
            q = int(s.node_owner[i])
            tgt, tsub = ctxs[q], subs[q]
            g2l = {int(g): k for k, g in enumerate(tsub.node_gid)}
            lb, le = int(lrows[i]), int(lrows[i + 1])
            lr = np.full(le - lb, g2l[int(s.node_gid[i])])
            lc = np.array([g2l[int(s.node_gid[c])] for c in lcols[lb:le]])
            dr, dc = expand_block_entries(lr, lc, [le - lb], b, layout == O.LAYOUT_PER_ROW)
            n = dr.size
            slots = torch.empty(n, dtype=torch.int64, device="cuda:0")
            tgt.lookup_value_slots(n, torch.from_numpy(dr).cuda(), torch.from_numpy(dc).cuda(), slots)
            contrib = torch.from_numpy(np.ascontiguousarray(lvals[lb * bb:le * bb])).cuda()
            assert bool((slots >= 0).all())
            tgt.add_values_at(n, slots, contrib)
            tgt.synchronize()
    for s, ctx, (lrows, lcols, _) in zip(subs, ctxs, partial):
        vals = ctx.to_host(A.ARRAY_VALUES)
        bbn = int(lrows[s.nb_own_node]) * bb
        vals[bbn:] = 0.0  # the sender zeroes its ghost rows after the send (ExchangePlan.exchange)
        check_owned_rows(s, b, layout, lrows, lcols, vals, grows, gcols, gvals)
        # columns in the solver's global numbering (owned-first numbering => node_gid-independent check of the kernel)
        l2g = np.arange(s.nb_node * b, dtype=np.int32)[::-1].copy()
        out = torch.empty(lcols.size * bb, dtype=torch.int32, device="cuda:0")
        from ctypes import c_void_p
        A._check(A.lib().afb_renumber_columns(ctx._h, c_void_p(torch.from_numpy(l2g).cuda().data_ptr()), c_void_p(out.data_ptr())))
        ctx.synchronize()
        ref_cols = lcols if b == 1 else O.bsr_to_csr(b, lrows, lcols)[1]
        if b == 1 or layout == O.LAYOUT_PER_ROW:
            assert np.array_equal(out.cpu().numpy(), l2g[ref_cols])
        ctx.close()


@pytest.mark.gpu
def test_device_slab_generator_with_ghost_layer():
    from arcanefem_b200 import capi as A
    with A.Context(0) as ctx:
        for dim, n, world in ((3, 6, 3), (2, 9, 2)):
            for r in range(world):
                k_lo, k_hi = M.slab_layers(n, world, r)
                ref = M.box_slab(dim, n, world, r, ghost_cell_layer=True)
                info = ctx.generate_box(dim, n, k_lo=k_lo, k_hi=k_hi, ghost_cell_layer=True)
                assert (info["nb_node"], info["nb_cell"], info["nb_own_node"], info["nb_own_cell"]) == (ref.nb_node, ref.nb_cell, ref.nb_own_node, ref.nb_own_cell)
                assert np.array_equal(ctx.to_host(A.ARRAY_COORDS).reshape(-1, 3), ref.coords)
                assert np.array_equal(ctx.to_host(A.ARRAY_CELL_NODES).reshape(-1, dim + 1), ref.cells)


# -------------------------------------------------------------------------------------------------
# GPU: ghost rows pulled over peer memory (csrc/p2p.cu).  Two or three processes share the one GPU of
# the test box (CUDA IPC works between processes on the same device; the kernels of the processes are
# time-sliced, the waits inside the kernel see the other process' progress); gloo carries the set-up.
# -------------------------------------------------------------------------------------------------
def run_rank_p2p(rank, world, port, case, transport, errq):
    try:
        import torch
        import torch.distributed as dist
        from arcanefem_b200 import capi as A
        from arcanefem_b200.distributed import DistributedAssembly
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        name, op, layout = case
        mesh = M.box_mesh(3, 7) if name == "box3d" else M.read_msh(os.path.join(ROOT, "tests", "golden", name))
        b = O.block_size(op, mesh.dim)
        params = list(O.lame(21.0e5, 0.28)) if op == O.OP_ELASTICITY else None
        fmt = A.FORMAT_BSR if b > 1 else A.FORMAT_CSR
        s = M.partition_mesh(mesh, world)[rank]
        grows, gcols, gvals = global_reference(mesh, op, params, layout)
        ctx = A.Context(0)
        ctx.set_mesh(s.dim, s.coords, s.cells, s.is_own)
        ctx.set_own_cell_count(s.nb_own_cell)
        da = DistributedAssembly(ctx, rank, world, s.node_gid, s.node_owner, s.nb_own_node, 0, transport=transport, comm_device="cpu")
        for rep in range(3):  # steady state: pattern re-build + assembly + exchange, epochs advance
            ctx.build_pattern(b)
            da.assemble(op, params=params, fmt=fmt, variant=A.VARIANT_TILED_GATHER, layout=layout, mode="exchange")
            if transport == "p2p":
                if rep == 1:
                    da.wait()  # explicit join; p2p_status below joins by itself in the other repetitions
                assert ctx.p2p_status() == 0, "ghost-row exchange timed out"
            else:
                ctx.synchronize()
            lrows, lcols = ctx.to_host(A.ARRAY_ROWS), ctx.to_host(A.ARRAY_COLUMNS)
            check_owned_rows(s, b, layout, lrows, lcols, ctx.to_host(A.ARRAY_VALUES), grows, gcols, gvals)
        if transport == "p2p":  # rank skew as a number (afb_p2p_wait_stats): read and clear
            ready_us, pulled_us, nex = ctx.p2p_wait_stats()
            assert nex == 3 and ready_us >= 0.0 and pulled_us >= 0.0 and ready_us + pulled_us < 3 * 60e6
            assert ctx.p2p_wait_stats()[2] == 0
            dist.barrier()
        if transport == "p2p":
            ctx.p2p_disconnect()
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    except Exception:
        errq.put((rank, traceback.format_exc()))
        raise


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", [("box3d", O.OP_POISSON, O.LAYOUT_PER_BLOCK), ("box3d", O.OP_ELASTICITY, O.LAYOUT_PER_ROW), ("sphere_cut.msh", O.OP_POISSON, O.LAYOUT_PER_BLOCK)],
                         ids=["box-poisson", "box-elasticity-per-row", "sphere-poisson"])
def test_ghost_rows_pulled_over_peer_memory(world, case):
    _run_exchange_processes(world, case, "p2p")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [("box3d", O.OP_POISSON, O.LAYOUT_PER_BLOCK), ("box3d", O.OP_ELASTICITY, O.LAYOUT_PER_ROW)], ids=["box-poisson", "box-elasticity-per-row"])
def test_ghost_rows_through_the_transport_callbacks(case):
    """the fall-back of the exchange (no peer memory): the rows travel through the two transport callbacks (torch.distributed
    here) on contexts with their own streams -- everything is ordered on the context stream by the library"""
    _run_exchange_processes(2, case, "nccl")


def _run_exchange_processes(world, case, transport):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errq = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=run_rank_p2p, args=(r, world, port, case, transport, errq)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    errs = []
    while not errq.empty():
        errs.append(errq.get())
    for p in procs:
        if p.is_alive():
            p.terminate()
            errs.append((-1, "rank timed out"))
    assert not errs, "\n".join(f"rank {r}:\n{t}" for r, t in errs)
    assert all(p.exitcode == 0 for p in procs)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("name", ["box3d", "box2d", "sphere_cut.msh"])
def test_native_rcb_partition_semantics(name, world):
    """afb_partition_* (recursive coordinate bisection, C++): Arcane sub-domain semantics as ArcaneFEM sees them -- every cell
    owned once, every node owned once (lowest rank among its cells), one ghost-cell layer, owned-first numbering with the
    ghosts grouped by ascending owner then global id, balanced cell counts."""
    from arcanefem_b200.distributed import partition_mesh_native
    mesh = M.box_mesh(3, 6) if name == "box3d" else (M.box_mesh(2, 12) if name == "box2d" else M.read_msh(os.path.join(ROOT, "tests", "golden", name)))
    subs = partition_mesh_native(mesh, world)
    cell_seen = np.zeros(mesh.nb_cell, dtype=np.int32)
    node_seen = np.zeros(mesh.nb_node, dtype=np.int32)
    owner_of = np.full(mesh.nb_node, -1)
    for s in subs:
        cell_seen[s.cell_gid[:s.nb_own_cell]] += 1
        node_seen[s.node_gid[:s.nb_own_node]] += 1
        owner_of[s.node_gid[:s.nb_own_node]] = s.rank
    assert (cell_seen == 1).all() and (node_seen == 1).all()
    counts = [s.nb_own_cell for s in subs]
    assert max(counts) - min(counts) <= max(2, world)
    cell_rank = np.empty(mesh.nb_cell, dtype=np.int32)
    for s in subs:
        cell_rank[s.cell_gid[:s.nb_own_cell]] = s.rank
    low = np.full(mesh.nb_node, world)
    np.minimum.at(low, mesh.cells.ravel(), np.repeat(cell_rank, mesh.npc))
    assert np.array_equal(owner_of, np.where(low == world, 0, low))
    for s in subs:
        assert np.array_equal(s.node_owner, owner_of[s.node_gid]) and np.array_equal(s.is_own, (s.node_owner == s.rank).astype(np.uint8))
        assert (s.node_owner[:s.nb_own_node] == s.rank).all() and (s.node_owner[s.nb_own_node:] != s.rank).all()
        gh = s.node_owner[s.nb_own_node:]
        assert (np.diff(gh) >= 0).all()
        for q in np.unique(gh):
            g = s.node_gid[s.nb_own_node:][gh == q]
            assert (np.diff(g) > 0).all()
        assert (np.diff(s.node_gid[:s.nb_own_node]) > 0).all()
        # cells: connectivity maps back, ghost cells = the foreign cells touching an owned node
        assert np.array_equal(s.node_gid[s.cells], mesh.cells[s.cell_gid])
        touches = (owner_of[mesh.cells] == s.rank).any(axis=1)
        expected_ghost = np.nonzero(touches & (cell_rank != s.rank))[0]
        assert np.array_equal(np.sort(s.cell_gid[s.nb_own_cell:]), expected_ghost)
        assert np.allclose(s.coords, mesh.coords[s.node_gid])
