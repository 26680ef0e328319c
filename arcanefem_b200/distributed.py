"""Domain-decomposed assembly across the GPUs of one box: one process per GPU.

The logic lives in libafb200 (csrc/mgpu.cu, declared in include/afb200.h): exchange-plan construction, owner lookup,
slot lists, the peer-memory exchange and the solver's global row numbering
(HypreDoFLinearSystemImpl::_computeMatrixNumeration, femutils/HypreDoFLinearSystem.cc:209-249).  This module is the
binding: it hands the library two communication primitives (`afb_transport`: allgather, neighbour exchange) implemented
over torch.distributed -- what MPI is to the reference -- and exposes

  Transport            the two callbacks over a torch.distributed group (gloo on CPU, nccl on GPUs)
  ExchangePlan         the host-only index logic (afb_xplan_host_*), with the data path as callables, so that the CPU tests
                       (gloo, world_size 2 and 3) run the same C++ code as the GPUs
  DistributedAssembly  one rank's share of an assembly on its GPU (afb_xplan_*): assemble own cells into the rows of all
                       local nodes, pull the partial ghost rows over NVLink peer memory (or through the transport)

Reference semantics mirrored (Arcane sub-domains as ArcaneFEM sees them): every node has exactly one owning rank; a rank
assembles the rows of its owned nodes only (`isOwn` gates: modules/testlab/CsrGpuBiliAssembly.cc:273,351); the solver numbers
the rows globally, owned rows of rank r contiguous after those of rank r-1.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np


@dataclass
class Numbering:
    first_dof: np.ndarray      # int64 [world+1]: first global DoF of every rank (exclusive scan of owned DoF counts)
    dof_l2g: np.ndarray        # int32 [nb_node*b]: local DoF -> global row


_ALLGATHER = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p)
_EXCHANGE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int)


class _TransportStruct(C.Structure):
    _fields_ = [("user", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("allgather", _ALLGATHER), ("exchange", _EXCHANGE),
                ("exchange_takes_device_memory", C.c_int32), ("pad", C.c_int32)]


class Transport:
    """afb_transport over a torch.distributed process group.  comm_device: where tensors handed to the backend live
    ("cpu" for gloo, "cuda:i" for nccl).  device_index: the GPU whose memory the library may pass to `exchange`
    (None: the callback only moves host memory)."""

    def __init__(self, rank, world, group=None, comm_device="cpu", device_index=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world, self.comm_device, self.device_index = rank, world, comm_device, device_index
        self.error = None
        self._ag = _ALLGATHER(self._allgather)
        self._ex = _EXCHANGE(self._exchange)
        takes_dev = 1 if (device_index is not None and str(comm_device).startswith("cuda")) else 0
        self.struct = _TransportStruct(None, rank, world, self._ag, self._ex, takes_dev, 0)

    def _host_tensor(self, ptr, nbytes):
        if not nbytes or not ptr:
            return self.torch.empty(0, dtype=self.torch.uint8)
        buf = (C.c_uint8 * int(nbytes)).from_address(ptr)
        return self.torch.frombuffer(buf, dtype=self.torch.uint8, count=int(nbytes))

    def _allgather(self, user, send, nbytes, recv):
        try:
            torch, dist = self.torch, self.dist
            mine = self._host_tensor(send, nbytes).clone().to(self.comm_device)
            out = [torch.zeros(int(nbytes), dtype=torch.uint8, device=self.comm_device) for _ in range(self.world)]
            if self.world > 1:
                dist.all_gather(out, mine, group=self.group)
            else:
                out[0] = mine
            dst = self._host_tensor(recv, nbytes * self.world)
            for q in range(self.world):
                dst[q * nbytes:(q + 1) * nbytes] = out[q].cpu()
            return 0
        except Exception as e:  # noqa: BLE001 -- reported through the C status, re-raised by the caller
            self.error = e
            return 1

    def _exchange(self, user, nb_peer, peer, send, send_bytes, recv, recv_bytes, device_memory):
        try:
            torch, dist = self.torch, self.dist
            ops, landing = [], []
            for k in range(nb_peer):
                q, ns, nr = int(peer[k]), int(send_bytes[k]), int(recv_bytes[k])
                if device_memory:
                    from . import capi as A
                    if ns:
                        ops.append(dist.P2POp(dist.isend, A.as_torch(send[k], (ns,), np.uint8, self.device_index), q, group=self.group))
                    if nr:
                        ops.append(dist.P2POp(dist.irecv, A.as_torch(recv[k], (nr,), np.uint8, self.device_index), q, group=self.group))
                else:
                    if ns:
                        ops.append(dist.P2POp(dist.isend, self._host_tensor(send[k], ns).clone().to(self.comm_device), q, group=self.group))
                    if nr:
                        t = torch.empty(nr, dtype=torch.uint8, device=self.comm_device)
                        landing.append((recv[k], nr, t))
                        ops.append(dist.P2POp(dist.irecv, t, q, group=self.group))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if device_memory:
                torch.cuda.current_stream(self.device_index).synchronize()
            for ptr, nr, t in landing:
                self._host_tensor(ptr, nr)[:] = t.cpu()
            return 0
        except Exception as e:  # noqa: BLE001
            self.error = e
            return 1

    def check(self, rc):
        """raise what a callback caught, else the library's error"""
        from . import capi as A
        if rc != 0:
            if self.error is not None:
                e, self.error = self.error, None
                raise e
            A._check(rc)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class ExchangePlan:
    """Host-only index logic of one rank's ghost-row exchange (afb_xplan_host_*; no GPU needed), with the data path behind
    callables so the CPU tests drive it with numpy / torch CPU tensors:
      tail_pattern()             -> (rows_tail int32[nb_ghost+1], cols_tail int32[...]) block pattern of the ghost rows
                                    (rows_tail[0] = first block entry of the first ghost row)
      lookup(dof_rows, dof_cols) -> int64 tensor of value slots in THIS rank's layout (-1: entry absent)
      values_slice(first, n)     -> tensor view of `values[first:first+n]` (send buffer)
      add_at(slots, contrib)     -> values[slots] += contrib
      make_buffer(n)             -> float64 tensor for receiving
    After the sends the ghost rows are zeroed (the reference's ghost rows are all-zero: isOwn gate)."""

    def __init__(self, rank, world, node_gid, node_owner, nb_own_node, b, layout_per_row, tail_pattern, lookup, values_slice, add_at, make_buffer,
                 group=None, comm_device="cpu"):
        from . import capi as A
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world, self.b = rank, world, b
        self.transport = Transport(rank, world, group=group, comm_device=comm_device)
        self.node_gid, self.node_owner = _i64(node_gid), _i32(node_owner)
        self.values_slice, self.add_at, self.make_buffer = values_slice, add_at, make_buffer
        rows_tail, cols_tail = tail_pattern()
        rows_tail, cols_tail = _i32(rows_tail), _i32(cols_tail)
        if cols_tail.size == 0:
            cols_tail = np.zeros(1, dtype=np.int32)
        self._h = C.c_void_p()
        lib = A.lib()
        self.transport.check(lib.afb_xplan_host_create(C.byref(self.transport.struct), int(b), A.LAYOUT_PER_ROW if layout_per_row else A.LAYOUT_PER_BLOCK,
                                                       C.c_int32(self.node_gid.size), C.c_int32(int(nb_own_node)), A._ptr(self.node_gid), A._ptr(self.node_owner),
                                                       A._ptr(rows_tail), A._ptr(cols_tail), C.byref(self._h)))
        n, peer, sf, sc, rcnt = C.c_int32(), C.POINTER(C.c_int32)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)(), C.POINTER(C.c_int64)()
        A._check(lib.afb_xplan_host_peers(self._h, C.byref(n), C.byref(peer), C.byref(sf), C.byref(sc), C.byref(rcnt)))
        self.send, self.recv = [], []   # (peer, first_value, nb_values) / (peer, slots, buffer)
        for k in range(n.value):
            q = int(peer[k])
            if sc[k]:
                self.send.append((q, int(sf[k]), int(sc[k])))
            if rcnt[k]:
                dr, dc = C.POINTER(C.c_int32)(), C.POINTER(C.c_int32)()
                A._check(lib.afb_xplan_host_pairs(self._h, k, C.byref(dr), C.byref(dc)))
                m = int(rcnt[k])
                slots = lookup(np.ctypeslib.as_array(dr, (m,)).copy(), np.ctypeslib.as_array(dc, (m,)).copy())
                slots = slots if isinstance(slots, torch.Tensor) else torch.from_numpy(np.asarray(slots, dtype=np.int64))
                assert bool((slots >= 0).all()), "a neighbour's partial row has an entry outside this rank's pattern"
                self.recv.append((q, slots, make_buffer(m)))

    def close(self):
        if self._h:
            from . import capi as A
            A.lib().afb_xplan_host_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def exchange(self):
        """Send the partial ghost rows to their owners and add the received ones (CPU data path of the tests)."""
        dist = self.dist
        ops = []
        for q, first, n in self.send:
            ops.append(dist.P2POp(dist.isend, self.values_slice(first, n), q, group=self.group))
        for q, slots, buf in self.recv:
            ops.append(dist.P2POp(dist.irecv, buf, q, group=self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for q, first, n in self.send:
            self.values_slice(first, n).zero_()
        for q, slots, buf in self.recv:
            self.add_at(slots, buf)

    def bytes_per_exchange(self):
        return 8 * sum(n for _, _, n in self.send), 8 * sum(int(buf.numel()) for _, _, buf in self.recv)

    def numbering(self) -> Numbering:
        from . import capi as A
        first = np.zeros(self.world + 1, dtype=np.int64)
        l2g = np.zeros(self.node_gid.size * self.b, dtype=np.int32)
        self.transport.check(A.lib().afb_xplan_host_numbering(self._h, A._ptr(first), A._ptr(l2g)))
        return Numbering(first_dof=first, dof_l2g=l2g)


class _PlanInfo:
    """what bench.py / the tests read off a DistributedAssembly's plan"""

    def __init__(self, nb_peer, sent, received, kind, why):
        self.nb_peer, self.sent, self.received, self.kind, self.why = nb_peer, sent, received, kind, why
        self.p2p = True if kind == 1 else None
        self.p2p_error = why or None

    def bytes_per_exchange(self):
        return self.sent, self.received


class DistributedAssembly:
    """One rank's share of a domain-decomposed assembly on its GPU (context `ctx`): afb_xplan_* of the C ABI."""

    def __init__(self, ctx, rank, world, node_gid, node_owner, nb_own_node, device_index, group=None, transport="p2p", comm_device=None, overlap=True):
        """transport "p2p": ghost rows travel in one kernel over NVLink peer memory (CUDA IPC, csrc/p2p.cu) when every rank can
        map its neighbours, else through the transport; "nccl": always through the transport (torch.distributed send/recv of the
        rows in place + accumulate kernels).  comm_device: where set-up tensors live (default the GPU, for the nccl backend;
        "cpu" with a gloo group)."""
        import os
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.node_gid, self.node_owner, self.nb_own_node = _i64(node_gid), _i32(node_owner), int(nb_own_node)
        self.device_index = device_index
        self.allow_p2p = transport == "p2p" and not os.environ.get("AFB_P2P_DISABLE")  # (the variable forces the transport path)
        self.transport = Transport(rank, world, group=group, comm_device=comm_device or f"cuda:{device_index}", device_index=device_index)
        self._x = C.c_void_p()
        self._key = None
        self.plan = None

    def _destroy_plan(self):
        if self._x:
            from . import capi as A
            A.lib().afb_xplan_destroy(self._x)
            self._x = C.c_void_p()

    def close(self):
        self._destroy_plan()

    def __del__(self):
        try:
            self._destroy_plan()
        except Exception:  # noqa: BLE001
            pass

    def _build_plan(self):
        from . import capi as A
        lib = A.lib()
        self._destroy_plan()
        self.transport.check(lib.afb_xplan_create(self.ctx._h, C.byref(self.transport.struct), A._ptr(self.node_gid), A._ptr(self.node_owner),
                                                  C.c_int32(self.nb_own_node), int(self.allow_p2p), C.byref(self._x)))
        n, s, r, kind, why = C.c_int32(), C.c_int64(), C.c_int64(), C.c_int32(), C.c_char_p()
        A._check(lib.afb_xplan_info(self._x, C.byref(n), C.byref(s), C.byref(r), C.byref(kind), C.byref(why)))
        self.plan = _PlanInfo(n.value, s.value, r.value, kind.value, (why.value or b"").decode())

    def invalidate(self):
        """Call on ALL ranks when the mesh or the pattern size changed: the next assemble() re-creates the plan (collective)."""
        self._key = None

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
        # the plan follows the pattern (block size, value layout, size, storage) and the mesh: every rank sees the same sequence
        # of such changes, so re-creating it is collective by construction; a change seen by one rank alone must be announced
        # with invalidate() on all ranks
        key = (ctx.b, layout, ctx.nnz, ctx.bsr_view()["values"])
        if self._key != key:
            self._build_plan()
            self._key = key
        A._check(A.lib().afb_xplan_exchange(self._x))

    def wait(self):
        """Orders the context stream after the exchange of the last assemble() (the peer-memory kernel runs on a side stream so
        that the next BuildMatrix overlaps it) and raises if that exchange timed out.  Call before anything that reads or
        writes the matrix values."""
        if self._x:
            from . import capi as A
            A._check(A.lib().afb_xplan_wait(self._x))

    def numbering(self) -> Numbering:
        if not self._x:
            raise RuntimeError("numbering() needs the exchange plan: assemble once in mode 'exchange' first")
        from . import capi as A
        b = self.ctx.b
        first = np.zeros(self.world + 1, dtype=np.int64)
        l2g = np.zeros(self.node_gid.size * b, dtype=np.int32)
        self.transport.check(A.lib().afb_xplan_numbering(self._x, A._ptr(first), A._ptr(l2g)))
        return Numbering(first_dof=first, dof_l2g=l2g)


def partition_mesh_native(mesh, world):
    """afb_partition_* (recursive coordinate bisection in C++): list of mesh.Subdomain, same semantics as mesh.partition_mesh."""
    from . import capi as A
    from .mesh import Subdomain
    lib = A.lib()
    h = C.c_void_p()
    coords = np.ascontiguousarray(mesh.coords, dtype=np.float64)
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    A._check(lib.afb_partition_create(int(mesh.dim), int(mesh.npc), C.c_int32(mesh.nb_node), C.c_int64(mesh.nb_cell), A._ptr(coords), A._ptr(cells), int(world), C.byref(h)))
    subs = []
    try:
        for r in range(world):
            nn, no, nc, noc = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int64()
            A._check(lib.afb_partition_sizes(h, r, C.byref(nn), C.byref(no), C.byref(nc), C.byref(noc)))
            xyz = np.empty((nn.value, 3), dtype=np.float64)
            cn = np.empty((nc.value, mesh.npc), dtype=np.int32)
            own = np.empty(nn.value, dtype=np.uint8)
            gid = np.empty(nn.value, dtype=np.int64)
            owner = np.empty(nn.value, dtype=np.int32)
            cgid = np.empty(nc.value, dtype=np.int64)
            A._check(lib.afb_partition_get(h, r, A._ptr(xyz), A._ptr(cn), A._ptr(own), A._ptr(gid), A._ptr(owner), A._ptr(cgid)))
            subs.append(Subdomain(rank=r, world=world, dim=mesh.dim, coords=xyz, cells=cn, nb_own_cell=noc.value, nb_own_node=no.value, is_own=own,
                                  node_gid=gid, node_owner=owner, cell_gid=cgid))
    finally:
        lib.afb_partition_destroy(h)
    return subs
