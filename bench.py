#!/usr/bin/env python
"""bench.py — the assembly hot path on B200, next to the reference's CPU path.

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...   # reference CPU algorithm (oracle port)

Metric (BASELINE.json): assembly elements/s of 3-D Tet4 P1 Poisson into CSR.
One *step* = one `AssembleBilinearOperator` of the reference (the scope its own benchmark
times: modules/testlab/CsrGpuBiliAssembly.cc:313-337) = BuildMatrix (sparsity pattern from
the mesh, allocation/zero fill) + AddAndCompute (element matrices + scatter).  Both
sub-timers are reported in `phases`; `roofline` is the dominant kernel (the value assembly),
`roofline_pattern` the BuildMatrix phase.

Workload at N=1: BASELINE config C2, structured box n=120 (10 368 000 Tet4, 1 771 561 nodes,
nnz 26 223 481), jittered, generated in HBM.  N>1: weak scaling, global box
n = round(120*N^(1/3)) cut in N z-slabs (one slab per GPU, one process per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "assembly elements/s (3-D Tet4 P1 Poisson, CSR; BuildMatrix+AddAndCompute per step)"
UNIT = "elements/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def box_counts(n):
    nb_node = (n + 1) ** 3
    nb_edge = 3 * n * (n + 1) ** 2 + 3 * n * n * (n + 1) + n ** 3
    return 6 * n ** 3, nb_node, nb_edge, nb_node + 2 * nb_edge


def algorithmic_bytes(nb_cell, nb_node, nnz, b=1, npc=4):
    """BASELINE.md §3 / SURVEY.md §8(d)."""
    values = 4 * npc * nb_cell + 24 * nb_node + 8 * b * b * nnz + 4 * nnz + 4 * (nb_node + 1)
    pattern = 4 * npc * nb_cell + 4 * nnz + 4 * (nb_node + 1)
    return values, pattern


def slab_layers(n, world, rank):
    """cube layers [k_lo,k_hi) of rank's z-slab (balanced)."""
    from arcanefem_b200 import mesh as M
    return M.slab_layers(n, world, rank)


def global_n(world, n1):
    return n1 if world == 1 else int(round(n1 * world ** (1.0 / 3.0)))


# -------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe)
# -------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device)],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(names, p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# -------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's sequential CSR back-end, one sub-domain per
# host thread (= what `mpirun -n N Testlab` does; oracle/afb_oracle.c orc_reference_rank)
# -------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_step(mesh, n, nthreads):
    """One full AssembleBilinearOperator (BuildMatrix + AddAndCompute) of the box mesh, split in
    `nthreads` z-slab sub-domains processed concurrently.  Returns seconds (wall)."""
    from oracle import oracle as O
    m = n + 1
    plane_nodes, layer_cells = m * m, 6 * n * n
    parts = []
    base, rem = divmod(m, nthreads)
    k = 0
    for t in range(nthreads):
        k1 = k + base + (1 if t < rem else 0)
        if k1 > k:
            parts.append((k, k1))
        k = k1
    res = [None] * len(parts)

    def work(i):
        k0, k1 = parts[i]
        res[i] = O.reference_rank(3, mesh.coords, mesh.cells, layer_cells * max(k0 - 1, 0), layer_cells * min(k1, n), plane_nodes * k0, plane_nodes * k1)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(parts))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    nnz = sum(r["nnz"] for r in res)
    assert nnz == box_counts(n)[3], (nnz, box_counts(n)[3])
    return dt, max(r["seconds"][0] for r in res), max(r["seconds"][1] for r in res)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from arcanefem_b200 import mesh as M
    world = args.gpus
    n_job = global_n(world, args.n)
    # bounded sample of the same workload: the box of the job when it is CPU-affordable, else a smaller box
    n = min(n_job, args.cpu_n)
    threads = host_threads()
    mesh = M.box_mesh(3, n)
    nb_cell = mesh.nb_cell
    for _ in range(args.warmup):
        reference_step(mesh, n, threads)
    times, tb, ta = [], [], []
    for _ in range(args.steps):
        dt, b_, a_ = reference_step(mesh, n, threads)
        times.append(dt)
        tb.append(b_)
        ta.append(a_)
    total = sum(times)
    value = nb_cell * args.steps / total
    sample = f"box n={n} ({nb_cell} Tet4) of the job's n={n_job}; {threads} host threads = {threads} MPI-rank-like z-slab sub-domains, BuildMatrix+AddAndCompute per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2 3-D Poisson P1 Tet4 CSR, structured box n={n_job}", "sample": sample, "format": "csr (reference CPU back-end, CsrBiliAssembly.cc)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "phases": {"build_matrix_ms": 1e3 * statistics.mean(tb), "add_and_compute_ms": 1e3 * statistics.mean(ta)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(args, n_job):
    """Bounded sample (about 10-30 s of CPU work) of the same workload on the box's host cores."""
    from arcanefem_b200 import mesh as M
    n = min(n_job, args.cpu_n)
    mesh = M.box_mesh(3, n)
    threads = host_threads()
    reference_step(mesh, n, threads)
    reps = 2
    t = sum(reference_step(mesh, n, threads)[0] for _ in range(reps))
    seq_n = min(n, 64)
    mseq = mesh if seq_n == n else M.box_mesh(3, seq_n)
    tseq = reference_step(mseq, seq_n, 1)[0]
    return {"value": mesh.nb_cell * reps / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"box n={n} ({mesh.nb_cell} Tet4), {reps} full assemblies (BuildMatrix+AddAndCompute), {threads} host threads as MPI-rank-like z-slabs",
            "sequential_value": mseq.nb_cell / tseq, "sequential_sample": f"box n={seq_n}, 1 thread"}


# -------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------
VARIANT_NAMES = {0: "cellwise-atomic (csr-gpu)", 1: "nodewise (nwcsr / AF-CSR)", 2: "tiled-gather (atomic-free, B200)"}


VARIANT_KERNEL = {0: "k_assemble_cellwise", 1: "k_assemble_nodewise", 2: "k_assemble_tiled"}


def committed_traffic(kernel, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the value kernel, from the committed
    `ncu --set full` capture of the same kernel on the same box size (profiles/traffic.json); None otherwise
    (a bench run is never taken under the profiler)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if kernel is None or n is None or not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    e = t.get(f"{kernel}:n={n}")
    if not e:
        return None, None
    return float(e["dram_bytes_read"] + e["dram_bytes_write"]), e["source"]


def run_b200(args):
    import torch
    import torch.distributed as dist
    from arcanefem_b200 import capi as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx = A.Context(local_rank, stream=stream.cuda_stream)

    from arcanefem_b200 import mesh as M
    from arcanefem_b200.distributed import DistributedAssembly

    n = global_n(world, args.n)
    k_lo, k_hi = slab_layers(n, world, rank)
    # N>1: one z-slab per GPU with Arcane-like ghosts (one ghost cell layer; each node has one owner)
    info = ctx.generate_box(3, n, k_lo=k_lo, k_hi=k_hi, ghost_cell_layer=world > 1)
    nb_cell_local = info["nb_own_cell"]            # throughput counts every cell once (its owner)
    nbr, nnz = ctx.build_pattern(1)
    bytes_values, bytes_pattern = algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz)
    da = None
    if world > 1:
        gid, owner_rel, nb_own, _, _ = M.box_slab_numbering(3, n, k_lo, k_hi, True)
        da = DistributedAssembly(ctx, rank, world, gid, (rank + owner_rel).astype(np.int32), nb_own, local_rank, transport=args.transport)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    torch.cuda.set_stream(stream)  # NCCL work and torch ops are ordered on the context's stream
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def assemble(c, variant, mode):
        if da is None or c is not ctx:
            c.assemble(A.OP_POISSON, variant=variant)
        else:
            da.assemble(A.OP_POISSON, variant=variant, mode=mode)

    # --- variant choice -----------------------------------------------------------------------
    variants = [A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER]
    if args.variant != "auto":
        variants = [{"atomic": 0, "nodewise": 1, "tiled": 2}[args.variant]]
    per_variant = {}
    for v in variants:
        try:
            ctx.reset_values()
            ctx.assemble(A.OP_POISSON, variant=v)  # inspector / first launch
        except A.AfbError as e:
            if "not available" in str(e):
                continue
            raise
        ts = []
        for _ in range(3):
            ctx.reset_values()
            e0, e1 = ev(), ev()
            e0.record(stream)
            ctx.assemble(A.OP_POISSON, variant=v)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        per_variant[v] = min(ts)
    variant = min(per_variant, key=per_variant.get)
    if world > 1:  # same variant everywhere
        t = torch.tensor([variant], device=dev)
        dist.broadcast(t, 0)
        variant = int(t.item())
    mode = args.mode if world > 1 else "single"
    # BuildMatrix algorithm of the steady-state step: the reference pairs the atomic assembly with the sparsity
    # computed from the cells (computeSparsityAtomic) and the atomic-free assembly with the one walking the
    # init-time node-node connectivity (computeSparsityAtomicFree); --sparsity overrides.
    SPARSITY = {"cells": A.SPARSITY_FROM_CELLS, "connectivity": A.SPARSITY_FROM_CONNECTIVITY}
    sparsity = args.sparsity if args.sparsity != "auto" else ("cells" if variant == A.VARIANT_CELLWISE_ATOMIC else "connectivity")
    per_sparsity = {}
    for name, algo in SPARSITY.items():
        ctx.set_sparsity_algorithm(algo)
        ts = []
        for _ in range(4):
            e0, e1 = ev(), ev()
            e0.record(stream)
            ctx.build_pattern(1)
            e1.record(stream)
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        per_sparsity[name] = min(ts[1:])
    ctx.set_sparsity_algorithm(SPARSITY[sparsity])

    def step(events=None):
        if events is not None:
            events[0].record(stream)
        ctx.build_pattern(1)                      # BuildMatrix: pattern + allocation (+ zero fill when the variant needs it)
        if events is not None:
            events[1].record(stream)
        assemble(ctx, variant, mode)              # AddAndCompute (+ ghost-row exchange over NCCL)
        if events is not None:
            events[2].record(stream)

    other_ms = None
    if world > 1:
        # the other decomposition scheme, for the record (phases.other_scheme_ms)
        other = "replicate" if mode == "exchange" else "exchange"
        for _ in range(3):
            ctx.build_pattern(1)
            assemble(ctx, variant, other)
        barrier()
        e0, e1 = ev(), ev()
        e0.record(stream)
        for _ in range(5):
            ctx.build_pattern(1)
            assemble(ctx, variant, other)
        da.wait()
        e1.record(stream)
        barrier()
        other_ms = (other, e0.elapsed_time(e1) / 5)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    evs = [[ev(), ev(), ev()] for _ in range(args.steps)]
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record(stream)
    for k in range(args.steps):
        step(evs[k])
    if da is not None:
        da.wait()  # the last step's exchange (side stream) belongs to the timed region
    t_end.record(stream)
    barrier()
    if rank == 0:
        time.sleep(0.15)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    inspector = ctx.inspector_timings()
    transport = "none"
    if da is not None and da.plan is not None:
        transport = "p2p" if da.plan.p2p is not None else "nccl"
        if args.transport == "p2p" and transport != "p2p" and rank == 0:
            print(f"bench.py: peer-memory exchange unavailable ({da.plan.p2p_error}); NCCL send/recv used instead", file=sys.stderr)
    if world > 1 and mode == "exchange" and transport == "p2p":
        st = ctx.p2p_status()
        if st != 0:
            raise SystemExit(f"bench.py: ghost-row exchange timed out on rank {rank} (status {st})")
    total_ms = t_start.elapsed_time(t_end)
    pattern_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    values_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    exch_bytes = da.plan.bytes_per_exchange() if (da is not None and da.plan is not None) else (0, 0)

    # --- e2e: host mesh in (pinned) -> C ABI -> host CSR out (pinned), every step ---------------
    e2e_steps = max(1, min(args.steps, 5))
    nb_node_l, nb_cell_all = info["nb_node"], info["nb_cell"]
    coords_h = torch.empty((nb_node_l, 3), dtype=torch.float64, pin_memory=True)
    cells_h = torch.empty((nb_cell_all, 4), dtype=torch.int32, pin_memory=True)
    own_h = torch.empty((nb_node_l,), dtype=torch.uint8, pin_memory=True)
    coords_h.copy_(A.as_torch(info["xyz"], (nb_node_l, 3), np.float64, local_rank))
    cells_h.copy_(A.as_torch(info["cell_nodes"], (nb_cell_all, 4), np.int32, local_rank))
    if info["is_own"]:
        own_h.copy_(A.as_torch(info["is_own"], (nb_node_l,), np.uint8, local_rank))
    rows_h = torch.empty((nbr + 1,), dtype=torch.int32, pin_memory=True)
    cols_h = torch.empty((nnz,), dtype=torch.int32, pin_memory=True)
    vals_h = torch.empty((nnz,), dtype=torch.float64, pin_memory=True)
    ctx2 = A.Context(local_rank, stream=stream.cuda_stream)
    da2 = None
    if world > 1:
        da2 = DistributedAssembly(ctx2, rank, world, da.node_gid, da.node_owner, da.nb_own_node, local_rank, transport=args.transport)

    def e2e_step(v):
        ctx2.set_mesh(3, coords_h.numpy(), cells_h.numpy(), own_h.numpy() if info["is_own"] else None)
        ctx2.set_own_cell_count(info["nb_own_cell"])
        ctx2.build_pattern(1)
        if da2 is None:
            ctx2.assemble(A.OP_POISSON, variant=v)
        else:
            da2.assemble(A.OP_POISSON, variant=v, mode=mode)
            da2.wait()
        ctx2.to_host(A.ARRAY_ROWS, rows_h.numpy())
        ctx2.to_host(A.ARRAY_COLUMNS, cols_h.numpy())
        ctx2.to_host(A.ARRAY_VALUES, vals_h.numpy())

    # a new mesh every step: the tile inspector is not amortised here, so the cheaper of the atomic and
    # the tiled variant is used for the end-to-end number (chosen by one timed trial each, rank 0 decides)
    trial = {}
    for v in sorted({A.VARIANT_CELLWISE_ATOMIC, variant}):
        e2e_step(v)
        barrier()
        t0 = time.perf_counter()
        e2e_step(v)
        barrier()
        trial[v] = time.perf_counter() - t0
    e2e_variant = min(trial, key=trial.get)
    if world > 1:
        t = torch.tensor([e2e_variant], device=dev)
        dist.broadcast(t, 0)
        e2e_variant = int(t.item())
    e2e_step(e2e_variant)
    barrier()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step(e2e_variant)
    e1.record(stream)
    barrier()
    e2e_serial_ms = e0.elapsed_time(e1)
    e2e_ms, e2e_pipelined = e2e_serial_ms, False
    if world == 1 and not args.no_e2e_pipeline:
        try:
            # Streaming form of the same step: several contexts, each on its own stream and driven by its own host thread
            # (ctypes releases the GIL), so one lane's H2D overlaps another's D2H (PCIe: 55 + 51 GB/s one way, 76 GB/s both
            # ways on this box) and the kernels of either.  Every step still copies its own inputs in and its CSR arrays out.
            import concurrent.futures
            nl = max(2, args.e2e_lanes)
            extra = []
            for _ in range(nl - 1):
                st_k = torch.cuda.Stream(device=dev)
                extra.append((st_k, A.Context(local_rank, stream=st_k.cuda_stream),
                              (torch.empty_like(rows_h).pin_memory(), torch.empty_like(cols_h).pin_memory(), torch.empty_like(vals_h).pin_memory())))

            def lane_step(c, outs):
                c.set_mesh(3, coords_h.numpy(), cells_h.numpy(), None)
                c.set_own_cell_count(info["nb_own_cell"])
                c.build_pattern(1)
                c.assemble(A.OP_POISSON, variant=e2e_variant)
                c.to_host(A.ARRAY_ROWS, outs[0].numpy())
                c.to_host(A.ARRAY_COLUMNS, outs[1].numpy())
                c.to_host(A.ARRAY_VALUES, outs[2].numpy())

            lanes = [(ctx2, (rows_h, cols_h, vals_h))] + [(c, o) for _, c, o in extra]
            per_lane = 2 * e2e_steps
            pipe_steps = nl * per_lane
            step_s = 1e-3 * e2e_serial_ms / e2e_steps
            with concurrent.futures.ThreadPoolExecutor(max_workers=nl) as pool:
                def run_lane(k, count):
                    if count > 1:
                        time.sleep(k * step_s / nl)  # lanes start out of phase (inside the timed region): uploads meet downloads
                    for _ in range(count):
                        lane_step(*lanes[k])
                list(pool.map(lambda k: run_lane(k, 1), range(nl)))  # warm-up of every lane
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                list(pool.map(lambda k: run_lane(k, per_lane), range(nl)))
                torch.cuda.synchronize(dev)
                pipe_ms = 1e3 * (time.perf_counter() - t0)
            for _, c, o in extra:
                assert torch.equal(o[2], vals_h) or e2e_variant == A.VARIANT_CELLWISE_ATOMIC, "pipelined lanes disagree"
                assert torch.equal(o[1], cols_h) and torch.equal(o[0], rows_h), "pipelined lanes disagree (pattern)"
                c.close()
            if pipe_ms / pipe_steps < e2e_serial_ms / e2e_steps:
                e2e_ms, e2e_pipelined = pipe_ms * e2e_steps / pipe_steps, True
        except Exception as exc:  # noqa: BLE001 -- the one-step-at-a-time number above stands
            print(f"bench.py: pipelined e2e skipped ({type(exc).__name__}: {exc})", file=sys.stderr)
    h2d = coords_h.numel() * 8 + cells_h.numel() * 4 + (own_h.numel() if info["is_own"] else 0)
    d2h = rows_h.numel() * 4 + cols_h.numel() * 4 + vals_h.numel() * 8
    checksum = float(vals_h.sum())
    ctx2.close()

    # --- reduce over ranks (max time, summed work) ---------------------------------------------
    stats = torch.tensor([total_ms, pattern_ms, values_ms, e2e_ms, float(nb_cell_local), float(launches), float(h2d), float(d2h),
                          float(bytes_values), float(bytes_pattern), e2e_serial_ms], dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx = sm = stats
    total_ms, pattern_ms, values_ms, e2e_ms = (float(mx[i]) for i in range(4))
    cells_all = float(sm[4])
    if rank == 0:
        peak, peak_src = measured_peaks()
        value = cells_all * args.steps / (total_ms * 1e-3)
        ach_values = float(mx[8]) / (values_ms * 1e-3) / 1e9      # slowest rank's kernel on its own slab
        ach_pattern = float(mx[9]) / (pattern_ms * 1e-3) / 1e9
        traffic, traffic_src = committed_traffic(VARIANT_KERNEL.get(variant), n if world == 1 else None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2 3-D Poisson P1 Tet4 CSR, structured box n={n} jitter 0.2 ({int(cells_all)} Tet4), z-slab per GPU",
                       "format": "csr", "variant": VARIANT_NAMES[variant],
                       "sparsity": {"cells": "from the cells (computeSparsityAtomic)",
                                    "connectivity": "from the init-time node-node connectivity (computeSparsityAtomicFree)"}[sparsity], "l2": "inputs larger than L2 (connectivity+values > 126 MB per GPU), no flush",
                       "parallelism": f"slab{world}" + ("" if world == 1 else f" ({mode}: " + (("own cells + ghost rows pulled over NVLink peer memory in one kernel on a side stream, overlapping the next BuildMatrix" if transport == "p2p" else "own cells + NCCL ghost-row exchange") if mode == "exchange" else "ghost cells recomputed, no exchange") + ")")},
            "phases": {"build_matrix_ms": pattern_ms, "add_and_compute_ms": values_ms,
                       "values_only_elements_per_s": cells_all / (values_ms * 1e-3),
                       "variants_ms": {VARIANT_NAMES[k]: v for k, v in per_variant.items()},
                       "build_matrix_ms_by_sparsity": per_sparsity,
                       "inspector_ms_once_per_mesh": inspector,
                       "decomposition": mode, "exchange_bytes_sent_recv_rank0": list(exch_bytes),
                       "other_scheme_ms_per_step": None if other_ms is None else {other_ms[0]: other_ms[1]}},
            "roofline": {"bound": "hbm", "kernel": "value assembly (AddAndCompute)", "achieved": ach_values, "peak": peak, "unit": "GB/s", "frac": ach_values / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": float(mx[8])},
            "roofline_pattern": {"bound": "hbm", "kernel": "BuildMatrix phase (degree, scan, columns)", "achieved": ach_pattern, "peak": peak, "unit": "GB/s",
                                 "frac": ach_pattern / peak, "algorithmic_bytes": float(mx[9])},
            "e2e": {"value": cells_all * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sm[6]), "d2h_bytes_per_step": int(sm[7]),
                    "steps": e2e_steps, "variant": VARIANT_NAMES[e2e_variant],
                    "pipelined": f"{max(2, args.e2e_lanes)} lanes (contexts / streams / host threads) out of phase: one lane's H2D overlaps another's D2H and kernels" if e2e_pipelined else "no (one step at a time)",
                    "one_step_at_a_time_value": cells_all * e2e_steps / (float(mx[10]) * 1e-3),
                    "what": "afb_set_mesh(host) + afb_build_pattern + afb_assemble_bilinear (+ ghost-row exchange) + afb_copy_to_host(rows, columns, values)",
                    "values_checksum": checksum},
            "gpu_launches": int(sm[5]),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(args, n)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=120, help="box size at N=1 (C2: 120)")
    ap.add_argument("--cpu-n", type=int, default=120, help="largest box the CPU legs run (bounded sample)")
    ap.add_argument("--variant", default="auto", choices=["auto", "atomic", "nodewise", "tiled"])
    ap.add_argument("--sparsity", default="auto", choices=["auto", "cells", "connectivity"], help="steady-state BuildMatrix algorithm (auto: by variant, as the reference pairs them)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e: one step at a time only")
    ap.add_argument("--e2e-lanes", type=int, default=4, help="e2e: pipelined lanes (independent problems in flight)")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="N>1, exchange mode: one pull kernel over NVLink peer memory (CUDA IPC) or NCCL send/recv + accumulate kernels")
    ap.add_argument("--mode", default="exchange", choices=["exchange", "replicate"], help="N>1: ghost-row exchange over NCCL (north star) or the reference's ghost-cell replication")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: everything native libraries print there (NCCL's version banner, ...) is sent
    # to stderr for the duration of the run, the line itself goes to the saved descriptor
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    real_stdout = os.fdopen(saved, "w")
    sys.stdout = real_stdout
    try:
        return run_reference(args) if args.impl == "reference" else run_b200(args)
    finally:
        real_stdout.flush()


if __name__ == "__main__":
    sys.exit(main())
