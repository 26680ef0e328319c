"""Every assembly / BuildMatrix variant once (after a warm-up), for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`.
NVTX-free: the phases are told apart by the marker kernel launches (afb_rhs_reset = one memset) ... simpler: each phase is run
TWICE in a fixed order that scratch/traffic_from_ncu.py knows."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arcanefem_b200 import capi as A
import bench

def run_tiled_only(dim, n, b):
    """large configurations (C4, C3): the tiled gather and the connectivity-based BuildMatrix only"""
    ctx = A.Context(0)
    ctx.generate_box(dim, n)
    lam = bench.E_MOD * bench.NU / ((1 + bench.NU) * (1 - 2 * bench.NU)); mu = bench.E_MOD / (2 * (1 + bench.NU))
    op, params, fmt = (A.OP_POISSON, None, A.FORMAT_CSR) if b == 1 else (A.OP_ELASTICITY, [lam, mu], A.FORMAT_BSR)
    ctx.build_pattern(b)
    for layout in ([A.LAYOUT_PER_BLOCK] if b == 1 else [A.LAYOUT_PER_ROW, A.LAYOUT_PER_BLOCK]):
        for _ in range(3):
            ctx.build_pattern(b)
            ctx.assemble(op, params=params, fmt=fmt, variant=A.VARIANT_TILED_GATHER, layout=layout)
    ctx.synchronize()
    ctx.close()


def run(dim, n, b, tag):
    ctx = A.Context(0)
    info = ctx.generate_box(dim, n)
    lam = bench.E_MOD * bench.NU / ((1 + bench.NU) * (1 - 2 * bench.NU)); mu = bench.E_MOD / (2 * (1 + bench.NU))
    op, params, fmt = (A.OP_POISSON, None, A.FORMAT_CSR) if b == 1 else (A.OP_ELASTICITY, [lam, mu], A.FORMAT_BSR)
    ctx.build_pattern(b)                                   # first build: from the cells (k_pattern_rows count + write)
    layouts = [A.LAYOUT_PER_BLOCK] if b == 1 else [A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW]
    for layout in layouts:
        for variant in (A.VARIANT_CELLWISE_ATOMIC, A.VARIANT_NODEWISE, A.VARIANT_TILED_GATHER):
            for _ in range(2):
                ctx.reset_values()
                ctx.assemble(op, params=params, fmt=fmt, variant=variant, layout=layout)
        if b > 1:  # the other vector executor
            ctx.set_vector_executor(A.VEC_EXEC_UNITS if dim == 3 else A.VEC_EXEC_ROWS)
            for _ in range(2):
                ctx.reset_values()
                ctx.assemble(op, params=params, fmt=fmt, variant=A.VARIANT_TILED_GATHER, layout=layout)
            ctx.set_vector_executor(A.VEC_EXEC_AUTO)
    if b == 1:
        ctx.reset_values()
        ctx.assemble(op, fmt=A.FORMAT_COO, variant=A.VARIANT_CELLWISE_ATOMIC)
        for ex in (A.TILED_EXEC_CHAIN, A.TILED_EXEC_CHAIN_FLOW):
            ctx.set_tiled_executor(ex)
            for _ in range(2):
                ctx.build_pattern(b)
                ctx.assemble(op, variant=A.VARIANT_TILED_GATHER)
        ctx.set_tiled_executor(A.TILED_EXEC_BRICKS)
    for algo in (A.SPARSITY_FROM_CELLS, A.SPARSITY_FROM_CONNECTIVITY):
        ctx.set_sparsity_algorithm(algo)
        for _ in range(2):
            ctx.build_pattern(b)
    ctx.synchronize()
    ctx.close()

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
{"c4": lambda: run_tiled_only(3, 256, 1), "c3": lambda: run_tiled_only(3, 203, 3), "c2": lambda: run(3, 120, 1, "c2"), "e3": lambda: run(3, 100, 3, "e3"), "e2": lambda: run(2, 2048, 2, "e2"), "p2d": lambda: run(2, 2048, 1, "p2d")}[which]()
