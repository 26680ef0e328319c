"""Domain-decomposed assembly across the GPUs of one box: one process per GPU, NCCL for the
ghost-row exchange.  torch.distributed is plumbing here; the arithmetic (slot lookup,
accumulation of received partial rows, column renumbering) runs in libafb200's kernels.

Reference semantics mirrored (Arcane sub-domains as ArcaneFEM sees them):
  * every node has exactly one owning rank; a rank assembles the rows of its owned nodes only
    (`isOwn` gates: modules/testlab/CsrGpuBiliAssembly.cc:273,351; femutils/BSRFormat.h:287,344,444,512);
  * the solver numbers the rows globally, owned rows of rank r contiguous after those of rank r-1, and
    ghost DoFs learn their global row from their owner
    (HypreDoFLinearSystemImpl::_computeMatrixNumeration, femutils/HypreDoFLinearSystem.cc:209-249:
    `allGather(nb_own_row)` + `m_dof_matrix_numbering.synchronize()`).

Two ways to get the interface rows right:
  mode "replicate" (the reference's own scheme): each rank also holds one layer of ghost cells and
      recomputes them; no communication during assembly;
  mode "exchange" (BASELINE.json north star): each rank computes its own cells only, into the rows of
      ALL its local nodes; the partial sums that land in ghost rows are sent to the owners, which add
      them.  Ghost nodes are numbered last and grouped by owner, so the partial rows bound for one
      neighbour are one contiguous tail slice of `values`: the NCCL send buffer is the matrix itself.

`ExchangePlan` holds only index logic and communication, behind four callables, so that the same code
runs on CPU tensors with the gloo backend in the tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Numbering:
    first_dof: np.ndarray      # int64 [world+1]: first global DoF of every rank (exclusive scan of owned DoF counts)
    dof_l2g: np.ndarray        # int32 [nb_node*b]: local DoF -> global row


def _ranges_by_owner(node_owner, nb_own_node, rank):
    """ghost nodes are ordered by (owner, gid): contiguous local ranges per owner."""
    gh = np.asarray(node_owner[nb_own_node:])
    assert (gh != rank).all(), "ghost nodes must follow the owned ones"
    assert (np.diff(gh) >= 0).all(), "ghost nodes must be grouped by ascending owner rank"
    owners, starts = np.unique(gh, return_index=True)
    ends = list(starts[1:]) + [gh.size]
    return [(int(q), nb_own_node + int(s), nb_own_node + int(e)) for q, s, e in zip(owners, starts, ends)]


def expand_block_entries(rows_local, cols_local, run_len, b, layout_per_row):
    """Scalar (dof_row, dof_col) pairs of block entries in the memory order of the sender's values
    (BSRMatrix::findValueIndex layouts, femutils/BSRFormat.cc:79-106).  rows_local/cols_local: local node
    ids per block entry in CSR order; run_len: entries per block row (needed by the per-row layout)."""
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
        # index = rb*b*b + b*(x + i*nz) + j  ->  order: i, x, j
        dr = np.repeat(r[None, :, None] * b + np.arange(b)[:, None, None], b, axis=2)
        dc = np.broadcast_to(c[None, :, None] * b + np.arange(b)[None, None, :], (b, nz, b))
        out_r.append(dr.reshape(-1))
        out_c.append(dc.reshape(-1))
        pos += nz
    if not out_r:
        return np.empty(0, np.int32), np.empty(0, np.int32)
    return np.concatenate(out_r).astype(np.int32), np.concatenate(out_c).astype(np.int32)


class ExchangePlan:
    """Ghost-row exchange of one rank.

    Callables (device- or host-backed):
      tail_pattern()            -> (rows_tail int32[nb_ghost+1], cols_tail int32[...]) block pattern of the ghost rows
                                   (rows_tail[0] = first block entry of the first ghost row)
      lookup(dof_rows, dof_cols)-> int64 tensor of value slots in THIS rank's layout (-1: entry absent)
      values_slice(first, n)    -> tensor view of `values[first:first+n]` (zero-copy send buffer)
      add_at(slots, contrib)    -> values[slots] += contrib
      make_buffer(n)            -> float64 tensor for receiving
      fence()                   -> orders the context's stream and the transport's stream against each other (both ways)
    comm_device: where the tensors handed to torch.distributed live ("cpu" for gloo, "cuda:i" for nccl).
    After the sends the ghost rows are zeroed (the reference's ghost rows are all-zero: isOwn gate).
    p2p: optional object with export() / connect(...) / exchange() (capi.Context: afb_p2p_*): the per-assembly
    exchange then is ONE kernel pulling the neighbours' partial rows over NVLink peer memory (csrc/p2p.cu);
    torch.distributed only carries the set-up (IPC handles, slice offsets).
    """

    def __init__(self, rank, world, node_gid, node_owner, nb_own_node, b, layout_per_row, tail_pattern, lookup, values_slice, add_at, make_buffer,
                 group=None, comm_device="cpu", p2p=None, fence=None):
        import torch.distributed as dist
        self.fence = fence or (lambda: None)
        self.dist, self.group, self.comm_device = dist, group, comm_device
        self.rank, self.world, self.b = rank, world, b
        self.node_gid = np.asarray(node_gid, dtype=np.int64)
        self.node_owner = np.asarray(node_owner, dtype=np.int32)
        self.nb_own_node = int(nb_own_node)
        self.values_slice, self.add_at, self.make_buffer = values_slice, add_at, make_buffer
        self.send = []   # (peer, first_value, nb_values)
        self.recv = []   # (peer, slots tensor, buffer)
        self.p2p, self.p2p_error = None, None
        self._setup(layout_per_row, tail_pattern, lookup)
        if p2p is not None:
            self._setup_p2p(p2p)

    # -- helpers ---------------------------------------------------------------------------------
    def _exchange_arrays(self, out_by_peer, dtype):
        """Variable-size neighbour exchange of host int64/int32 arrays (setup only)."""
        import torch
        dist = self.dist
        dev = self.comm_device
        counts = torch.zeros(self.world, dtype=torch.int64)
        for q, a in out_by_peer.items():
            counts[q] = a.size
        counts = counts.to(dev)
        allc = [torch.zeros(self.world, dtype=torch.int64, device=dev) for _ in range(self.world)]
        dist.all_gather(allc, counts, group=self.group)
        allc = [c.cpu() for c in allc]
        ops, bufs, keep = [], {}, []
        for q, a in out_by_peer.items():
            if a.size:
                t = torch.from_numpy(np.ascontiguousarray(a.astype(dtype))).to(dev)
                keep.append(t)
                ops.append(dist.P2POp(dist.isend, t, q, group=self.group))
        for q in range(self.world):
            n = int(allc[q][self.rank])
            if q != self.rank and n:
                bufs[q] = torch.empty(n, dtype=torch.from_numpy(np.empty(0, dtype)).dtype, device=dev)
                ops.append(dist.P2POp(dist.irecv, bufs[q], q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return {q: t.cpu().numpy() for q, t in bufs.items()}

    def _local_of_gid(self, gids):
        order = getattr(self, "_gid_order", None)
        if order is None:
            self._gid_order = order = np.argsort(self.node_gid, kind="stable")
            self._gid_sorted = self.node_gid[order]
        pos = np.searchsorted(self._gid_sorted, gids)
        pos = np.minimum(pos, self._gid_sorted.size - 1)
        ok = self._gid_sorted[pos] == gids
        return np.where(ok, order[pos], -1)

    # -- setup --------------------------------------------------------------------------------------
    def _setup(self, layout_per_row, tail_pattern, lookup):
        b = self.b
        ranges = _ranges_by_owner(self.node_owner, self.nb_own_node, self.rank)
        rows_tail, cols_tail = tail_pattern()
        rows_tail = np.asarray(rows_tail, dtype=np.int64)
        base = int(rows_tail[0]) if rows_tail.size else 0
        out = {}
        for q, g0, g1 in ranges:
            r0, r1 = int(rows_tail[g0 - self.nb_own_node]), int(rows_tail[g1 - self.nb_own_node])
            cols = np.asarray(cols_tail[r0 - base:r1 - base], dtype=np.int64)
            deg = np.diff(rows_tail[g0 - self.nb_own_node:g1 - self.nb_own_node + 1])
            row_gid = np.repeat(self.node_gid[g0:g1], deg)
            out[q] = np.stack([row_gid, self.node_gid[cols]], axis=0).reshape(-1)  # [2, n] flattened
            self.send.append((q, r0 * b * b, (r1 - r0) * b * b))
        got = self._exchange_arrays(out, np.int64)
        for q in sorted(got):
            pairs = got[q].reshape(2, -1)
            lr, lc = self._local_of_gid(pairs[0]), self._local_of_gid(pairs[1])
            assert (lr >= 0).all() and (lr < self.nb_own_node).all(), "received a partial row of a node this rank does not own"
            assert (lc >= 0).all(), "a neighbour's partial row references a node unknown here (ghost layer missing)"
            run_len = np.diff(np.concatenate([[0], np.nonzero(np.diff(pairs[0]))[0] + 1, [pairs.shape[1]]])) if pairs.shape[1] else []
            dr, dc = expand_block_entries(lr, lc, run_len, b, layout_per_row)
            slots = lookup(dr, dc)
            assert bool((slots >= 0).all()), "a neighbour's partial row has an entry outside this rank's pattern"
            self.recv.append((q, slots, self.make_buffer(int(dr.size))))

    def _setup_p2p(self, p2p):
        """Collective: IPC handles of every rank's values array, and for every neighbour pair the slice one pulls
        from the other."""
        import torch
        dist = self.dist
        send = {q: (first, n) for q, first, n in self.send}
        recv = {q: slots for q, slots, _ in self.recv}
        peers = sorted(set(send) | set(recv))
        told = self._exchange_arrays({q: np.array(send.get(q, (0, 0)), dtype=np.int64) for q in peers}, np.int64)
        # every step below is attempted on every rank and the outcome agreed on collectively: if peer memory cannot be
        # mapped somewhere (no IPC in the container, no peer access), ALL ranks keep the torch.distributed transport
        err = None
        try:
            vh, fh = p2p.export()
        except Exception as e:  # noqa: BLE001
            err, vh, fh = e, bytes(64), bytes(64)
        mine = torch.tensor(list(vh + fh), dtype=torch.uint8).to(self.comm_device)
        allh = [torch.zeros(128, dtype=torch.uint8, device=self.comm_device) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)
        allh = [bytes(t.cpu().numpy().tobytes()) for t in allh]
        pull_first, pull_count, slots = [], [], []
        for q in peers:
            first, n = (int(x) for x in told[q])
            sl = recv.get(q)
            assert n == (0 if sl is None else int(sl.numel())), "neighbour's slice and the local slot list differ in length"
            pull_first.append(first)
            pull_count.append(n)
            slots.append(sl)
        if err is None:
            try:
                p2p.connect(self.rank, peers, [allh[q][:64] for q in peers], [allh[q][64:] for q in peers], pull_first, pull_count, slots,
                            [send.get(q, (0, 0))[0] for q in peers], [send.get(q, (0, 0))[1] for q in peers])
            except Exception as e:  # noqa: BLE001
                err = e
        ok = torch.tensor([0 if err is None else 1], dtype=torch.int32).to(self.comm_device)
        dist.all_reduce(ok, op=dist.ReduceOp.MAX, group=self.group)
        if int(ok.item()) != 0:
            self.p2p_error = str(err) if err is not None else "peer-memory mapping failed on another rank"
            try:
                p2p.disconnect()
            except Exception:  # noqa: BLE001
                pass
            return
        self.p2p = p2p

    # -- every assembly ------------------------------------------------------------------------------
    def exchange(self):
        """Send the partial ghost rows to their owners and add the received ones (stream-ordered for NCCL)."""
        if self.p2p is not None:
            self.p2p.exchange()
            return
        dist = self.dist
        # the assembly ran on the context's stream, the transport and the torch ops below run on torch's current
        # stream: order them explicitly (host-side fences; this is the portable fall-back, not the fast path)
        self.fence()
        ops = []
        for q, first, n in self.send:
            if n:
                ops.append(dist.P2POp(dist.isend, self.values_slice(first, n), q, group=self.group))
        for q, slots, buf in self.recv:
            ops.append(dist.P2POp(dist.irecv, buf, q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for q, first, n in self.send:
            if n:
                self.values_slice(first, n).zero_()
        self.fence()  # receives and zero fills (torch stream) are complete before the accumulate kernels (context stream)
        for q, slots, buf in self.recv:
            self.add_at(slots, buf)

    def bytes_per_exchange(self):
        return 8 * sum(n for _, _, n in self.send), 8 * sum(int(buf.numel()) for _, _, buf in self.recv)

    # -- global numbering ------------------------------------------------------------------------------
    def numbering(self) -> Numbering:
        """HypreDoFLinearSystemImpl::_computeMatrixNumeration (femutils/HypreDoFLinearSystem.cc:209-249)."""
        import torch
        dist, b = self.dist, self.b
        mine = torch.tensor([self.nb_own_node * b], dtype=torch.int64, device=self.comm_device)
        allc = [torch.zeros(1, dtype=torch.int64, device=self.comm_device) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=self.group)
        first = np.concatenate([[0], np.cumsum([int(c[0]) for c in allc])]).astype(np.int64)
        nb_node = self.node_gid.size
        l2g_node = np.full(nb_node, -1, dtype=np.int64)
        l2g_node[:self.nb_own_node] = first[self.rank] // b + np.arange(self.nb_own_node)
        # ghost nodes ask their owner for its local id (the reference's variable synchronize())
        ranges = _ranges_by_owner(self.node_owner, self.nb_own_node, self.rank)
        asked = self._exchange_arrays({q: self.node_gid[g0:g1] for q, g0, g1 in ranges}, np.int64)
        answers = {}
        for q, gids in asked.items():
            lid = self._local_of_gid(gids)
            assert (lid >= 0).all() and (lid < self.nb_own_node).all()
            answers[q] = lid.astype(np.int64)
        back = self._exchange_arrays(answers, np.int64)
        for q, g0, g1 in ranges:
            l2g_node[g0:g1] = first[q] // b + back[q]
        assert (l2g_node >= 0).all()
        dof = (l2g_node[:, None] * b + np.arange(b)[None, :]).reshape(-1)
        assert dof.max(initial=0) < 2 ** 31, "global row index exceeds Int32 (HYPRE_Int)"
        return Numbering(first_dof=first, dof_l2g=dof.astype(np.int32))


# ---------------------------------------------------------------------------------------------------
# GPU binding
# ---------------------------------------------------------------------------------------------------
class _P2P:
    """afb_p2p_* of one context behind the three calls ExchangePlan needs."""

    def __init__(self, ctx, overlap=True):
        self.ctx, self.overlap = ctx, overlap

    def export(self):
        import os
        if os.environ.get("AFB_P2P_DISABLE"):  # exercises the collective fall-back to the torch.distributed transport
            raise RuntimeError("peer-memory exchange disabled by AFB_P2P_DISABLE")
        return self.ctx.p2p_export()

    def connect(self, *a):
        self.ctx.p2p_connect(*a)

    def exchange(self):
        self.ctx.p2p_exchange(asynchronous=self.overlap)

    def wait(self):
        self.ctx.p2p_wait()

    def disconnect(self):
        self.ctx.p2p_disconnect()


class DistributedAssembly:
    """One rank's share of a domain-decomposed assembly on its GPU (context `ctx`)."""

    def __init__(self, ctx, rank, world, node_gid, node_owner, nb_own_node, device_index, group=None, transport="p2p", comm_device=None, overlap=True):
        """transport "p2p": ghost rows travel in one kernel over NVLink peer memory (CUDA IPC, csrc/p2p.cu);
        "nccl": torch.distributed send/recv of the rows + accumulate kernels.  comm_device: where set-up tensors live
        (default the GPU, for the nccl backend; "cpu" with a gloo group)."""
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.transport, self.comm_device, self.overlap = transport, comm_device, overlap
        self.node_gid, self.node_owner, self.nb_own_node = node_gid, node_owner, int(nb_own_node)
        self.device_index = device_index
        self.plan = None
        self._plan_key = None

    def _build_plan(self, layout):
        import torch
        from . import capi as A
        ctx, dev = self.ctx, self.device_index
        b = ctx.b
        nb_node, nnz = ctx.nb_block_row, ctx.nnz
        v = ctx.bsr_view()
        rows_t = A.as_torch(v["rows_index"], nb_node + 1, np.int32, dev)
        cols_t = A.as_torch(v["columns"], nnz, np.int32, dev)
        vals_t = A.as_torch(v["values"], nnz * b * b, np.float64, dev)
        own = self.nb_own_node

        def tail_pattern():
            rt = rows_t[own:].cpu().numpy()
            ct = cols_t[int(rt[0]):].cpu().numpy() if rt.size else np.empty(0, np.int32)
            return rt, ct

        def lookup(dr, dc):
            n = int(dr.size)
            slots = torch.empty(n, dtype=torch.int64, device=f"cuda:{dev}")
            if n:
                drt = torch.from_numpy(np.ascontiguousarray(dr)).to(f"cuda:{dev}")
                dct = torch.from_numpy(np.ascontiguousarray(dc)).to(f"cuda:{dev}")
                ctx.lookup_value_slots(n, drt, dct, slots)
                ctx.synchronize()
            return slots

        def fence():
            ctx.synchronize()
            torch.cuda.current_stream(dev).synchronize()

        self.plan = ExchangePlan(self.rank, self.world, self.node_gid, self.node_owner, own, b, layout == A.LAYOUT_PER_ROW, tail_pattern, lookup,
                                 values_slice=lambda first, n: vals_t[first:first + n],
                                 add_at=lambda slots, buf: ctx.add_values_at(int(slots.numel()), slots, buf),
                                 make_buffer=lambda n: torch.empty(n, dtype=torch.float64, device=f"cuda:{dev}"), group=self.group,
                                 comm_device=self.comm_device or f"cuda:{dev}", p2p=_P2P(ctx, overlap=self.overlap) if self.transport == "p2p" else None,
                                 fence=fence)

    def assemble(self, op, params=None, fmt=None, variant=None, layout=None, mode="exchange", flags=0):
        """Fresh assembly of this rank's rows (call after ctx.build_pattern).  mode "exchange": own cells
        only + ghost-row exchange; mode "replicate": own + ghost cells, owned rows only, no communication."""
        from . import capi as A
        ctx = self.ctx
        fmt = A.FORMAT_CSR if fmt is None else fmt
        variant = A.VARIANT_TILED_GATHER if variant is None else variant
        layout = A.LAYOUT_PER_BLOCK if layout is None else layout
        self.wait()  # the previous exchange (side stream) is done with `values` before they are overwritten
        if mode == "replicate":
            ctx.assemble(op, params=params, fmt=fmt, variant=variant, layout=layout, flags=flags)
            return
        ctx.assemble(op, params=params, fmt=fmt, variant=variant, layout=layout, flags=flags | A.FLAG_OWN_CELLS_ONLY | A.FLAG_ALL_ROWS)
        key = (ctx.b, layout, ctx.nnz)
        if self.plan is None or self._plan_key != key:
            # the value layout is stored by the assembly: build the plan after the first one
            self._build_plan(layout)
            self._plan_key = key
        try:
            self.plan.exchange()
        except A.AfbError as e:
            if "values array moved" not in str(e):
                raise
            self._build_plan(layout)  # collective: a re-allocation follows the pattern size, identical on the ranks' schedule
            self.plan.exchange()

    def wait(self):
        """Orders the context stream after the exchange of the last assemble() (transport p2p runs it on a side stream so that
        the next BuildMatrix overlaps it).  Call before anything that reads or writes the matrix values."""
        if self.plan is not None and self.plan.p2p is not None:
            self.plan.p2p.wait()

    def numbering(self):
        if self.plan is None:
            raise RuntimeError("numbering() needs the exchange plan: assemble once in mode 'exchange' first")
        return self.plan.numbering()
