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

Workload (the north star's target configuration): BASELINE config C4, structured box n=256
(100 663 296 Tet4, 16 974 593 nodes, nnz 253 036 801), jittered, generated in HBM.
N=1: the whole box on one GPU.  N>1: **strong scaling** -- the same n=256 box cut in N z-slabs
(one slab per GPU, one process per GPU), ghost rows exchanged over NVLink.  `--scaling weak`
grows the box with N instead (n = round(n1*N^(1/3))).  At N=1 the line also carries a `configs`
block: C2 (n=120, the 10 M-cell CSR case) and C3 (elasticity b=3, n=203, both BSR value layouts).
Every run checks the assembled matrix against the CPU oracle's full-size digests
(tests/golden/box_checksums.json: sum |a_ij| and trace, relative 1e-12) and fails loudly otherwise.
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


def global_n(world, n1, scaling="strong"):
    """box size of the job: strong scaling keeps the box, weak scaling keeps the cells per GPU."""
    if world == 1 or scaling == "strong":
        return n1
    return int(round(n1 * world ** (1.0 / 3.0)))


def workload_name(n):
    """`config.workload` of both arms (this repo's and --impl reference): the job, not how it is split."""
    tag = {120: "C2", 256: "C4"}.get(n, "box")
    return f"{tag} 3-D Poisson P1 Tet4 CSR, structured box n={n} jitter 0.2 ({6 * n ** 3} Tet4)"


def golden_digest(key):
    """Full-size digests of the CPU oracle (tests/golden/make_box_checksums.py); None when the size was not recorded."""
    p = os.path.join(ROOT, "tests", "golden", "box_checksums.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f).get(key)


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


class ReferenceJob:
    """The reference's sequential CSR back-end on the box mesh, split in `nthreads` z-slab sub-domains
    (one per host thread = one per MPI rank of `mpirun -n N Testlab`).  Init (untimed, FemModule::startInit):
    the node-node connectivity of every sub-domain.  step(): one AssembleBilinearOperator of all ranks, concurrently."""

    def __init__(self, mesh, n, nthreads, init_connectivity=True):
        from oracle import oracle as O
        self.O, self.mesh, self.n = O, mesh, n
        m = n + 1
        plane_nodes, layer_cells = m * m, 6 * n * n
        self.parts = []
        base, rem = divmod(m, nthreads)
        k = 0
        for t in range(nthreads):
            k1 = k + base + (1 if t < rem else 0)
            if k1 > k:
                self.parts.append((layer_cells * max(k - 1, 0), layer_cells * min(k1, n), plane_nodes * k, plane_nodes * k1))
            k = k1
        self.inits = [None] * len(self.parts)
        if init_connectivity:
            def mk(i):
                self.inits[i] = O.ReferenceRank(mesh.cells, *self.parts[i])
            self._run(mk)

    def _run(self, fn):
        ths = [threading.Thread(target=fn, args=(i,)) for i in range(len(self.parts))]
        for t in ths:
            t.start()
        for t in ths:
            t.join()

    def step(self):
        """-> (wall seconds, max BuildMatrix seconds, max AddAndCompute seconds)"""
        res = [None] * len(self.parts)

        def work(i):
            res[i] = self.O.reference_rank(3, self.mesh.coords, self.mesh.cells, *self.parts[i], init=self.inits[i])

        t0 = time.perf_counter()
        self._run(work)
        dt = time.perf_counter() - t0
        nnz = sum(r["nnz"] for r in res)
        assert nnz == box_counts(self.n)[3], (nnz, box_counts(self.n)[3])
        return dt, max(r["seconds"][0] for r in res), max(r["seconds"][1] for r in res)

    def close(self):
        for h in self.inits:
            if h is not None:
                h.close()


REF_BUILD = ("BuildMatrix = CsrFormat::initialize (allocate + 4 fills) + walk of the init-time node-node connectivity "
             "(modules/testlab/CsrBiliAssembly.cc:79-91), AddAndCompute = host element matrix + linear-scan matrixAddValue")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from arcanefem_b200 import mesh as M
    world = args.gpus
    n_job = global_n(world, args.n, args.scaling)
    # bounded sample of the same workload: the box of the job when it is CPU-affordable, else a smaller box
    n = min(n_job, args.cpu_n)
    threads = host_threads()
    mesh = M.box_mesh(3, n)
    nb_cell = mesh.nb_cell
    job = ReferenceJob(mesh, n, threads)
    for _ in range(args.warmup):
        job.step()
    times, tb, ta = [], [], []
    for _ in range(args.steps):
        dt, b_, a_ = job.step()
        times.append(dt)
        tb.append(b_)
        ta.append(a_)
    job.close()
    total = sum(times)
    value = nb_cell * args.steps / total
    sample = (f"box n={n} ({nb_cell} Tet4) of the job's n={n_job}; {threads} host threads = {threads} MPI-rank-like z-slab sub-domains, "
              f"BuildMatrix+AddAndCompute per step; {REF_BUILD}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n_job), "sample": sample, "format": "csr (reference CPU back-end, CsrBiliAssembly.cc)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "phases": {"build_matrix_ms": 1e3 * statistics.mean(tb), "add_and_compute_ms": 1e3 * statistics.mean(ta)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    sys.stdout.flush()
    return 0


def cpu_baseline_leg(args, n_job):
    """Bounded sample (about 10-30 s of CPU work) of the same workload on the box's host cores."""
    from arcanefem_b200 import mesh as M
    n = min(n_job, args.cpu_n)
    mesh = M.box_mesh(3, n)
    threads = host_threads()
    job = ReferenceJob(mesh, n, threads)
    job.step()
    reps = 3
    runs = [job.step() for _ in range(reps)]
    job.close()
    t = sum(r[0] for r in runs)
    # the first assembly of a run (connectivity built inside BuildMatrix), for the record
    first = ReferenceJob(mesh, n, threads, init_connectivity=False)
    t_first = first.step()[0]
    # sequential: one rank owns the whole sample box
    seq = ReferenceJob(mesh, n, 1)
    tseq = seq.step()[0]
    seq.close()
    return {"value": mesh.nb_cell * reps / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"box n={n} ({mesh.nb_cell} Tet4) of the job's n={n_job}, {reps} assemblies, {threads} host threads as MPI-rank-like z-slabs; {REF_BUILD}",
            "build_matrix_ms": 1e3 * statistics.mean(r[1] for r in runs), "add_and_compute_ms": 1e3 * statistics.mean(r[2] for r in runs),
            "first_assembly_value": mesh.nb_cell / t_first, "first_assembly_note": "node-node connectivity built from the cells inside BuildMatrix (what round 1 timed)",
            "sequential_value": mesh.nb_cell / tseq, "sequential_sample": f"box n={n} ({mesh.nb_cell} Tet4), 1 thread"}


# -------------------------------------------------------------------------------------------
# this repo's arm
# -------------------------------------------------------------------------------------------
VARIANT_NAMES = {0: "cellwise-atomic (csr-gpu)", 1: "nodewise (nwcsr / AF-CSR)", 2: "tiled-gather (atomic-free, B200)"}


VARIANT_KERNEL = {0: "k_assemble_cellwise", 1: "k_assemble_nodewise", 2: "k_assemble_tiled"}
E_MOD, NU = 21e5, 0.28  # modules/elasticity/inputs/bar.3D.Dirichlet.bodyForce.arc:25-26


def committed_traffic(kernel, n, exact_key=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the value kernel, from the committed
    ncu capture of the same kernel on the same box size (profiles/traffic.json); None otherwise
    (a bench run is never taken under the profiler)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if (kernel is None and exact_key is None) or n is None or not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    e = t.get(exact_key if exact_key else f"{kernel}:n={n}")
    if not e:
        return None, None
    return float(e["dram_bytes_read"] + e["dram_bytes_write"]), e["source"]


def matrix_digest(torch, A, ctx, device, nb_own_row, b=1, layout=0):
    """sum |a_ij| and trace of the rows [0, nb_own_row) of the matrix in HBM (fp64, on the device)."""
    v = ctx.csr_view() if b == 1 else ctx.bsr_view()
    rows_p = v["rows"] if b == 1 else v["rows_index"]
    rows = A.as_torch(rows_p, (nb_own_row + 1,), np.int32, device)
    nnz = int(rows[nb_own_row].item())
    if nnz == 0:
        return 0.0, 0.0, 0
    cols = A.as_torch(v["columns"], (nnz,), np.int32, device)
    vals = A.as_torch(v["values"], (nnz * b * b,), np.float64, device)
    return values_digest(torch, rows, cols, vals, nb_own_row, b, layout) + (nnz,)


def values_digest(torch, rows, cols, vals, nb_row, b=1, layout=0):
    counts = (rows[1:nb_row + 1] - rows[:nb_row]).to(torch.int64)
    rid = torch.repeat_interleave(torch.arange(nb_row, dtype=torch.int32, device=rows.device), counts)
    diag = torch.nonzero(cols[:rid.numel()] == rid).squeeze(1)
    del rid
    abs_sum = float(vals.abs().sum().item())
    if b == 1:
        trace = float(vals[diag].sum().item())
    elif layout == 0:  # per block: p*b*b + i*b + j
        blk = vals.view(-1, b * b)[diag]
        trace = float(blk[:, [i * b + i for i in range(b)]].sum().item())
    else:              # per row: rb*b*b + b*(x + i*nz) + j  (femutils/BSRFormat.h:356)
        row_of = torch.arange(nb_row, device=rows.device)
        rb = rows[:nb_row].to(torch.int64)
        x = diag - rb
        tr = 0.0
        for i in range(b):
            tr += float(vals[rb * b * b + b * (x + i * counts) + i].sum().item())
        trace = tr
        del row_of
    return abs_sum, trace


def check_digest(name, got_abs, got_trace, key, tol=1e-12):
    g = golden_digest(key)
    out = {"abs_sum": got_abs, "trace": got_trace, "golden": key if g else None}
    if g:
        ea = abs(got_abs - g["abs_sum"]) / abs(g["abs_sum"])
        et = abs(got_trace - g["trace"]) / abs(g["trace"])
        out.update(golden_abs_sum=g["abs_sum"], golden_trace=g["trace"], rel_err=max(ea, et), tol=tol, ok=bool(max(ea, et) <= tol))
        if not out["ok"]:
            raise SystemExit(f"bench.py: {name}: assembled matrix differs from the CPU oracle's digest ({key}): "
                             f"sum|a| {got_abs!r} vs {g['abs_sum']!r}, trace {got_trace!r} vs {g['trace']!r} (rel {max(ea, et):.3e} > {tol})")
    return out


def timed_steps(ev, stream, fn_build, fn_values, reps):
    """mean BuildMatrix / AddAndCompute ms over `reps` steps (CUDA events on the context's stream)."""
    evs = [[ev(), ev(), ev()] for _ in range(reps)]
    for e in evs:
        e[0].record(stream)
        fn_build()
        e[1].record(stream)
        fn_values()
        e[2].record(stream)
    evs[-1][2].synchronize()
    return statistics.mean(e[0].elapsed_time(e[1]) for e in evs), statistics.mean(e[1].elapsed_time(e[2]) for e in evs)


def side_config(torch, A, device, stream, ev, name, n, op, b, layouts, peak, coefficient=None):
    """One entry of the `configs` block: another BASELINE configuration on the same GPU, same protocol
    (steady-state BuildMatrix + AddAndCompute, tiled gather), checked against the oracle's digest.
    coefficient: a uniform per-cell conductivity (afb_set_cell_coefficient) -- the [nb_cell] array is read cell by cell all the same
    (k_assemble_tiled_coef), and the digest has to be `coefficient` times the golden one."""
    out = []
    ctx = A.Context(device, stream=stream.cuda_stream)
    try:
        info = ctx.generate_box(3, n)
        nbr, nnz = ctx.build_pattern(b)
        bytes_values, bytes_pattern = algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz, b=b)
        if coefficient is not None:
            ctx.set_cell_coefficient(float(coefficient))
            bytes_values += 8 * info["nb_cell"]
        params = None
        if op == A.OP_ELASTICITY:
            lam = E_MOD * NU / ((1 + NU) * (1 - 2 * NU))
            mu = E_MOD / (2 * (1 + NU))
            params = [lam, mu]
        for layout in layouts:
            fmt = A.FORMAT_CSR if b == 1 else A.FORMAT_BSR
            asm = lambda: ctx.assemble(op, params=params, fmt=fmt, variant=A.VARIANT_TILED_GATHER, layout=layout)
            ctx.reset_values()
            asm()  # inspector
            for _ in range(2):
                ctx.build_pattern(b)
                asm()
            bm, vm = timed_steps(ev, stream, lambda: ctx.build_pattern(b), asm, 5)
            key = ("poisson3d_n%d" if op == A.OP_POISSON else "elasticity3d_n%d") % n
            ga, gt, _ = matrix_digest(torch, A, ctx, device, nbr, b, layout)
            if coefficient is not None:
                ga, gt = ga / coefficient, gt / coefficient
            ach = bytes_values / (vm * 1e-3) / 1e9
            tkey = f"k_assemble_tiled<4>:3d:n={n}:b=1" if b == 1 else f"k_assemble_rows_vec<4, {layout}>:3d:n={n}:b={b}"
            if coefficient is not None:
                tkey = f"k_assemble_tiled_coef<4>:3d:n={n}:b=1"
            traffic, traffic_src = committed_traffic(None, n, exact_key=tkey)
            out.append({"config": name, "workload": f"box n={n} ({info['nb_cell']} Tet4), " + ("Poisson b=1 CSR" if b == 1 else f"elasticity b={b} BSR, values {'per block (BSR)' if layout == 0 else 'per row (AF-BSR / CSR hand-off)'}")
                        + ("" if coefficient is None else f", per-cell conductivity array (uniform {coefficient}: fourier / FourierNL modules), digest / {coefficient} checked"),
                        "variant": VARIANT_NAMES[2], "build_matrix_ms": bm, "add_and_compute_ms": vm, "ms_per_step": bm + vm,
                        "elements_per_s": info["nb_cell"] / ((bm + vm) * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes_per_launch": float(bytes_values),
                                     "traffic": traffic, "traffic_source": traffic_src},
                        "inspector_ms_once_per_mesh": ctx.inspector_timings(),
                        "check": check_digest(name, ga, gt, key)})
    finally:
        ctx.close()
    return out


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs next to its GPU (NVML's ideal CPU affinity) before any pinned host buffer is allocated,
    so that the e2e host<->device copies of N ranks do not all cross the socket interconnect.  Returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        target = cpus & allowed
        if not target or target == allowed:
            return "unchanged (%d CPUs allowed, %d next to the GPU)" % (len(allowed), len(cpus))
        os.sched_setaffinity(0, target)
        return "%d of %d allowed CPUs (next to GPU %d)" % (len(target), len(allowed), device_index)
    except Exception as exc:  # noqa: BLE001 -- the placement is an optimisation, never a requirement
        return "unchanged (%s)" % type(exc).__name__


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
    cpus_at_start = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    affinity = bind_to_gpu_numa(local_rank) if not args.no_bind else "unchanged (--no-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.Stream(device=dev)
    ctx = A.Context(local_rank, stream=stream.cuda_stream)

    from arcanefem_b200 import mesh as M
    from arcanefem_b200.distributed import DistributedAssembly

    n = global_n(world, args.n, args.scaling)
    k_lo, k_hi = slab_layers(n, world, rank)
    # N>1: one z-slab per GPU with Arcane-like ghosts (one ghost cell layer; each node has one owner)
    info = ctx.generate_box(3, n, k_lo=k_lo, k_hi=k_hi, ghost_cell_layer=world > 1)
    nb_cell_local = info["nb_own_cell"]            # throughput counts every cell once (its owner)
    nbr, nnz = ctx.build_pattern(1)
    bytes_values, bytes_pattern = algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz)
    da = None
    nb_own_row = nbr
    if world > 1:
        gid, owner_rel, nb_own, _, _ = M.box_slab_numbering(3, n, k_lo, k_hi, True)
        nb_own_row = int(nb_own)
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
        assemble(ctx, variant, mode)              # AddAndCompute (+ ghost-row exchange over NVLink)
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
    if world > 1 and mode == "exchange":
        ctx.p2p_wait_stats()  # clear: the counters below cover the timed steps only
    barrier()
    t_start, t_end = ev(), ev()
    t_start.record(stream)
    inner = bool(os.environ.get("AFB_BENCH_INNER_EVENTS"))
    for k in range(args.steps):
        step(evs[k] if inner else None)  # nothing but the step's own launches between the two bracketing events
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
    wait_ready_us = wait_pulled_us = 0.0
    if world > 1 and mode == "exchange" and transport == "p2p":
        st = ctx.p2p_status()
        if st != 0:
            raise SystemExit(f"bench.py: ghost-row exchange timed out on rank {rank} (status {st})")
        wr, wp, nex = ctx.p2p_wait_stats()  # rank skew as seen by this rank's exchange kernels, per step
        wait_ready_us, wait_pulled_us = wr / max(nex, 1), wp / max(nex, 1)
    total_ms = t_start.elapsed_time(t_end)
    if not inner:
        # phase split (BuildMatrix / AddAndCompute): the same steps again with an event between the phases, outside the
        # timed region (an event record between two kernels costs a few microseconds of idle stream per boundary)
        barrier()  # all ranks enter together: a late rank would show up as exchange wait in its neighbours' phases
        for k in range(args.steps):
            step(evs[k])
        if da is not None:
            da.wait()
        evs[-1][2].synchronize()
        barrier()
        if world > 1 and mode == "exchange" and transport == "p2p" and ctx.p2p_status() != 0:
            raise SystemExit(f"bench.py: ghost-row exchange timed out on rank {rank} (phase pass)")
    pattern_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    values_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    exch_bytes = da.plan.bytes_per_exchange() if (da is not None and da.plan is not None) else (0, 0)
    # digest of the matrix the timed region left in HBM (owned rows of every rank)
    dev_abs, dev_trace, dev_nnz = matrix_digest(torch, A, ctx, local_rank, nb_own_row)

    # --- first assembly of a mesh: pattern from the cells + inspector + assembly, nothing amortised --------------
    first_step_ms = None
    if world == 1 and not args.no_first_step:
        ctx.close()
        ctx = None
        c3 = A.Context(local_rank, stream=stream.cuda_stream)
        c3.generate_box(3, n)
        torch.cuda.synchronize(dev)
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        c3.build_pattern(1)
        e1.record(stream)
        c3.assemble(A.OP_POISSON, variant=variant)
        e2.record(stream)
        e2.synchronize()
        first_step_ms = {"build_matrix_ms": e0.elapsed_time(e1), "add_and_compute_ms": e1.elapsed_time(e2), "total_ms": e0.elapsed_time(e2),
                         "variant": VARIANT_NAMES[variant], "inspector_ms": c3.inspector_timings(),
                         "what": "fresh context: node->cell lists + pattern from the cells + tile inspector + value assembly (one cold call each)"}
        c3.close()

    # --- e2e: host mesh in (pinned) -> C ABI -> host CSR out (pinned), every step ---------------
    e2e_steps = max(1, min(args.steps, 5 if n <= 160 else 3))
    nb_node_l, nb_cell_all = info["nb_node"], info["nb_cell"]
    coords_h = torch.empty((nb_node_l, 3), dtype=torch.float64, pin_memory=True)
    cells_h = torch.empty((nb_cell_all, 4), dtype=torch.int32, pin_memory=True)
    own_h = torch.empty((nb_node_l,), dtype=torch.uint8, pin_memory=True)
    src = ctx
    if src is None:
        src = A.Context(local_rank, stream=stream.cuda_stream)
        info = src.generate_box(3, n)
    coords_h.copy_(A.as_torch(info["xyz"], (nb_node_l, 3), np.float64, local_rank))
    cells_h.copy_(A.as_torch(info["cell_nodes"], (nb_cell_all, 4), np.int32, local_rank))
    has_own = bool(info["is_own"])
    if has_own:
        own_h.copy_(A.as_torch(info["is_own"], (nb_node_l,), np.uint8, local_rank))
    torch.cuda.synchronize(dev)
    nb_own_cell = info["nb_own_cell"]
    if src is not ctx:
        src.close()
    rows_h = torch.empty((nbr + 1,), dtype=torch.int32, pin_memory=True)
    cols_h = torch.empty((nnz,), dtype=torch.int32, pin_memory=True)
    vals_h = torch.empty((nnz,), dtype=torch.float64, pin_memory=True)
    ctx2 = A.Context(local_rank, stream=stream.cuda_stream)
    da2 = None
    if world > 1:
        da2 = DistributedAssembly(ctx2, rank, world, da.node_gid, da.node_owner, da.nb_own_node, local_rank, transport=args.transport)

    def e2e_new_mesh_step(v):
        ctx2.set_mesh(3, coords_h.numpy(), cells_h.numpy(), own_h.numpy() if has_own else None)
        ctx2.set_own_cell_count(nb_own_cell)
        ctx2.build_pattern(1)
        if da2 is None:
            ctx2.assemble(A.OP_POISSON, variant=v)
        else:
            da2.invalidate()
            da2.assemble(A.OP_POISSON, variant=v, mode=mode)
            da2.wait()
        ctx2.to_host(A.ARRAY_ROWS, rows_h.numpy())
        ctx2.to_host(A.ARRAY_COLUMNS, cols_h.numpy())
        ctx2.to_host(A.ARRAY_VALUES, vals_h.numpy())

    def loop_step(c, d, outs):
        """one step of a time loop: this step's coordinates in, BuildMatrix + AddAndCompute (+ exchange), the CSR arrays out"""
        c.update_coordinates(coords_h.numpy())
        c.build_pattern(1)
        if d is None:
            c.assemble(A.OP_POISSON, variant=variant)
        else:
            d.assemble(A.OP_POISSON, variant=variant, mode=mode)
            d.wait()
        c.to_host(A.ARRAY_ROWS, outs[0].numpy())
        c.to_host(A.ARRAY_COLUMNS, outs[1].numpy())
        c.to_host(A.ARRAY_VALUES, outs[2].numpy())

    # --- (1) a new mesh every step: nothing is amortised (the tile inspector would run every step), so the cheaper of the
    #     atomic and the steady-state variant is taken by one timed trial each (rank 0 decides); reported as e2e.new_mesh
    trial = {}
    cand = sorted({A.VARIANT_CELLWISE_ATOMIC, variant}) if args.e2e_variant == "auto" else [variant if args.e2e_variant == "same" else {"atomic": 0, "nodewise": 1, "tiled": 2}[args.e2e_variant]]
    for v in cand:
        e2e_new_mesh_step(v)
        barrier()
        t0 = time.perf_counter()
        e2e_new_mesh_step(v)
        barrier()
        trial[v] = time.perf_counter() - t0
    e2e_variant = min(trial, key=trial.get)
    if world > 1:
        t = torch.tensor([e2e_variant], device=dev)
        dist.broadcast(t, 0)
        e2e_variant = int(t.item())
    e2e_new_mesh_step(e2e_variant)
    barrier()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_new_mesh_step(e2e_variant)
    e1.record(stream)
    barrier()
    new_mesh_ms = e0.elapsed_time(e1)
    # --- (2) the time loop the reference's own benchmark times (same mesh, re-assembled every iteration,
    #     modules/testlab/FemModule.cc:74-75 + cache_warming): topology and plans stay on the device; every step copies its
    #     coordinates in and the CSR arrays out.  Same variant as the device-timed `value`.  This is `e2e.value`.
    if da2 is not None:
        da2.invalidate()
    ctx2.set_mesh(3, coords_h.numpy(), cells_h.numpy(), own_h.numpy() if has_own else None)
    ctx2.set_own_cell_count(nb_own_cell)
    ctx2.build_pattern(1)
    for _ in range(2):
        loop_step(ctx2, da2, (rows_h, cols_h, vals_h))
    barrier()
    e0, e1 = ev(), ev()
    e0.record(stream)
    for _ in range(e2e_steps):
        loop_step(ctx2, da2, (rows_h, cols_h, vals_h))
    e1.record(stream)
    barrier()
    e2e_serial_ms = e0.elapsed_time(e1)
    e2e_ms, e2e_pipelined = e2e_serial_ms, False
    # digest of what reached the host (owned rows), computed on the device from the host arrays
    h_rows, h_cols, h_vals = rows_h.to(dev), cols_h.to(dev), vals_h.to(dev)
    own_nnz = int(h_rows[nb_own_row].item())
    e2e_abs, e2e_trace = values_digest(torch, h_rows, h_cols[:own_nnz], h_vals[:own_nnz], nb_own_row)
    del h_rows, h_cols, h_vals
    nl = max(2, args.e2e_lanes if n <= 160 else min(args.e2e_lanes, 3))
    if world == 1 and not args.no_e2e_pipeline:
        try:
            # Streaming form of the same loop: several contexts (independent problems), each on its own stream and driven by
            # its own host thread (ctypes releases the GIL), so one lane's H2D overlaps another's D2H (PCIe: 55 + 51 GB/s one
            # way, 76 GB/s both ways on this box) and the kernels of either.  Every step still copies its inputs in and its
            # CSR arrays out.
            import concurrent.futures
            extra = []
            for _ in range(nl - 1):
                st_k = torch.cuda.Stream(device=dev)
                c = A.Context(local_rank, stream=st_k.cuda_stream)
                c.set_mesh(3, coords_h.numpy(), cells_h.numpy(), None)
                c.build_pattern(1)
                extra.append((st_k, c, (torch.empty_like(rows_h).pin_memory(), torch.empty_like(cols_h).pin_memory(), torch.empty_like(vals_h).pin_memory())))
            lanes = [(ctx2, (rows_h, cols_h, vals_h))] + [(c, o) for _, c, o in extra]
            per_lane = 2 * e2e_steps
            pipe_steps = nl * per_lane
            step_s = 1e-3 * e2e_serial_ms / e2e_steps
            with concurrent.futures.ThreadPoolExecutor(max_workers=nl) as pool:
                def run_lane(k, count):
                    if count > 2:
                        time.sleep(k * step_s / nl)  # lanes start out of phase (inside the timed region): uploads meet downloads
                    for _ in range(count):
                        loop_step(lanes[k][0], None, lanes[k][1])
                list(pool.map(lambda k: run_lane(k, 2), range(nl)))  # warm-up of every lane (inspector, steady-state BuildMatrix)
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                list(pool.map(lambda k: run_lane(k, per_lane), range(nl)))
                torch.cuda.synchronize(dev)
                pipe_ms = 1e3 * (time.perf_counter() - t0)
            for _, c, o in extra:
                assert torch.equal(o[2], vals_h) or variant == A.VARIANT_CELLWISE_ATOMIC, "pipelined lanes disagree"
                assert torch.equal(o[1], cols_h) and torch.equal(o[0], rows_h), "pipelined lanes disagree (pattern)"
                c.close()
            if pipe_ms / pipe_steps < e2e_serial_ms / e2e_steps:
                e2e_ms, e2e_pipelined = pipe_ms * e2e_steps / pipe_steps, True
        except Exception as exc:  # noqa: BLE001 -- the one-step-at-a-time number above stands
            print(f"bench.py: pipelined e2e skipped ({type(exc).__name__}: {exc})", file=sys.stderr)
    h2d_new = coords_h.numel() * 8 + cells_h.numel() * 4 + (own_h.numel() if has_own else 0)
    h2d = coords_h.numel() * 8
    d2h = rows_h.numel() * 4 + cols_h.numel() * 4 + vals_h.numel() * 8
    if da2 is not None:
        da2.close()  # the exchange plan goes before its context
    ctx2.close()
    del coords_h, cells_h, rows_h, cols_h, vals_h

    # --- reduce over ranks (max time, summed work) ---------------------------------------------
    stats = torch.tensor([total_ms, pattern_ms, values_ms, e2e_ms, float(nb_cell_local), float(launches), float(h2d), float(d2h),
                          float(bytes_values), float(bytes_pattern), e2e_serial_ms, dev_abs, dev_trace, float(dev_nnz), e2e_abs, e2e_trace, new_mesh_ms, float(h2d_new),
                          wait_ready_us, wait_pulled_us],
                         dtype=torch.float64, device=dev)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        per_rank = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(per_rank, stats)
    else:
        mx = sm = stats
        per_rank = [stats]
    total_ms, pattern_ms, values_ms, e2e_ms = (float(mx[i]) for i in range(4))
    cells_all = float(sm[4])
    rc = 0
    if rank == 0:
        peak, peak_src = measured_peaks()
        value = cells_all * args.steps / (total_ms * 1e-3)
        ach_values = float(mx[8]) / (values_ms * 1e-3) / 1e9      # slowest rank's kernel on its own slab
        ach_pattern = float(mx[9]) / (pattern_ms * 1e-3) / 1e9
        traffic, traffic_src = committed_traffic(VARIANT_KERNEL.get(variant), n if world == 1 else None)
        key = f"poisson3d_n{n}"
        nnz_ok = int(sm[13]) == box_counts(n)[3]
        if not nnz_ok:
            raise SystemExit(f"bench.py: nnz of the owned rows over all ranks {int(sm[13])} != {box_counts(n)[3]}")
        check = check_digest("timed region (device)", float(sm[11]), float(sm[12]), key)
        check_e2e = check_digest("e2e (host arrays)", float(sm[14]), float(sm[15]), key)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n),
                       "format": "csr", "variant": VARIANT_NAMES[variant],
                       "sparsity": {"cells": "from the cells (computeSparsityAtomic)",
                                    "connectivity": "from the init-time node-node connectivity (computeSparsityAtomicFree)"}[sparsity], "l2": "inputs larger than L2 (connectivity+values > 126 MB per GPU), no flush",
                       "parallelism": f"slab{world}" + ("" if world == 1 else f" ({args.scaling} scaling, one z-slab per GPU; {mode}: " + (("own cells + ghost rows pulled over NVLink peer memory in one kernel on a side stream, overlapping the next BuildMatrix" if transport == "p2p" else "own cells + NCCL ghost-row exchange") if mode == "exchange" else "ghost cells recomputed, no exchange") + ")")},
            "phases": {"build_matrix_ms": pattern_ms, "add_and_compute_ms": values_ms,
                       "values_only_elements_per_s": cells_all / (values_ms * 1e-3),
                       "variants_ms": {VARIANT_NAMES[k]: v for k, v in per_variant.items()},
                       "build_matrix_ms_by_sparsity": per_sparsity,
                       "inspector_ms_once_per_mesh": inspector,
                       "first_step_ms": first_step_ms,
                       "decomposition": mode, "exchange_bytes_sent_recv_rank0": list(exch_bytes),
                       "per_rank_ms": [{"step": float(p[0]) / args.steps, "build_matrix": float(p[1]), "add_and_compute": float(p[2]),
                                        "exchange_wait_ready_us": float(p[18]), "exchange_wait_pulled_us": float(p[19])} for p in per_rank] if world > 1 else None,
                       "other_scheme_ms_per_step": None if other_ms is None else {other_ms[0]: other_ms[1]}},
            "roofline": {"bound": "hbm", "kernel": "value assembly (AddAndCompute)", "achieved": ach_values, "peak": peak, "unit": "GB/s", "frac": ach_values / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": float(mx[8])},
            "roofline_pattern": {"bound": "hbm", "kernel": "BuildMatrix phase (degree, scan, columns)", "achieved": ach_pattern, "peak": peak, "unit": "GB/s",
                                 "frac": ach_pattern / peak, "algorithmic_bytes": float(mx[9])},
            "check": check,
            "e2e": {"value": cells_all * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(sm[6]), "d2h_bytes_per_step": int(sm[7]),
                    "steps": e2e_steps, "variant": VARIANT_NAMES[variant],
                    "host_cpus_rank0": affinity,
                    "pcie_gbs_all_ranks": (float(sm[6]) + float(sm[7])) * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "pipelined": f"{nl} lanes (contexts / streams / host threads) out of phase: one lane's H2D overlaps another's D2H and kernels" if e2e_pipelined else "no (one step at a time)",
                    "one_step_at_a_time_value": cells_all * e2e_steps / (float(mx[10]) * 1e-3),
                    "what": "time loop on a resident mesh, as the reference's benchmark re-assembles the same mesh: afb_update_coordinates(host, pinned) + afb_build_pattern + "
                            "afb_assemble_bilinear (+ ghost-row exchange) + afb_copy_to_host(rows, columns, values), every step",
                    "new_mesh": {"value": cells_all * e2e_steps / (float(mx[16]) * 1e-3), "variant": VARIANT_NAMES[e2e_variant], "h2d_bytes_per_step": int(sm[17]),
                                 "d2h_bytes_per_step": int(sm[7]), "trial_s": {VARIANT_NAMES[k]: v for k, v in trial.items()},
                                 "what": "a new mesh every step, nothing amortised: afb_set_mesh(host) + node->cell lists + pattern from the cells + inspector "
                                         "(tiled variant) + assembly + afb_copy_to_host(rows, columns, values); one step at a time"},
                    "check": check_e2e},
            "gpu_launches": int(sm[5]),
            "clocks": clocks,
        }
        if world == 1 and not args.no_configs:
            if ctx is not None:
                ctx.close()
                ctx = None
            torch.cuda.empty_cache()
            cfgs = []
            if n != 120:
                cfgs += side_config(torch, A, local_rank, stream, ev, "C2", 120, A.OP_POISSON, 1, [A.LAYOUT_PER_BLOCK], peak)
            cfgs += side_config(torch, A, local_rank, stream, ev, "C3", args.n_c3, A.OP_ELASTICITY, 3, [A.LAYOUT_PER_ROW, A.LAYOUT_PER_BLOCK], peak)
            cfgs += side_config(torch, A, local_rank, stream, ev, "C2+conductivity", 120, A.OP_POISSON, 1, [A.LAYOUT_PER_BLOCK], peak, coefficient=2.5)
            line["configs"] = cfgs
        if cpus_at_start is not None:
            os.sched_setaffinity(0, cpus_at_start)  # the CPU baseline uses every host core again
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(args, n)
        if args.time_stats:
            # the reference's output/listing/time_stats.json (modules/testlab/FemModule.cc:19-56), for its own plotting scripts
            from arcanefem_b200.timestats import write_time_stats
            bm_cells = per_sparsity.get("cells", pattern_ms) * 1e-3
            fm = {"nwcsr": (pattern_ms * 1e-3, values_ms * 1e-3)} if variant == A.VARIANT_TILED_GATHER else {}
            if A.VARIANT_CELLWISE_ATOMIC in per_variant:
                fm["csr-gpu"] = (bm_cells, per_variant[A.VARIANT_CELLWISE_ATOMIC] * 1e-3)
                fm["coo-gpu"] = (bm_cells, per_variant[A.VARIANT_CELLWISE_ATOMIC] * 1e-3)  # same kernel: the COO back-end reads the row segment directly
            if A.VARIANT_NODEWISE in per_variant:
                fm["CsrNodeWise_ThreadPerRow"] = (bm_cells, per_variant[A.VARIANT_NODEWISE] * 1e-3)
            write_time_stats(args.time_stats, args.steps, world, 3, (n + 1) ** 3, 12 * n * n, int(cells_all), fm)
        print(json.dumps(line))
        sys.stdout.flush()
    if da is not None:
        da.close()  # the exchange plan goes before its context
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return rc


def main():
    import faulthandler
    faulthandler.enable()  # a crash inside native code prints the Python stack of every thread
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=256, help="box size of the job (C4: 256 = 100 663 296 Tet4; C2: 120)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N>1: the same box cut in N slabs (north star) or a box growing with N")
    ap.add_argument("--n-c3", type=int, default=203, help="box size of the C3 (elasticity b=3) entry of the configs block")
    ap.add_argument("--time-stats", default=None, help="also write the reference's time_stats.json (modules/testlab/FemModule.cc:19-56) to this path")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs block (C2, C3) at N=1")
    ap.add_argument("--no-first-step", action="store_true", help="skip the first-assembly (nothing amortised) measurement at N=1")
    ap.add_argument("--e2e-variant", default="auto", choices=["auto", "same", "atomic", "nodewise", "tiled"], help="e2e: cheaper of atomic / steady-state variant by trial, the steady-state variant, or a fixed one")
    ap.add_argument("--cpu-n", type=int, default=120, help="largest box the CPU legs run (bounded sample)")
    ap.add_argument("--variant", default="auto", choices=["auto", "atomic", "nodewise", "tiled"])
    ap.add_argument("--sparsity", default="auto", choices=["auto", "cells", "connectivity"], help="steady-state BuildMatrix algorithm (auto: by variant, as the reference pairs them)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-bind", action="store_true", help="do not pin the process to the CPUs next to its GPU")
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
