"""CUDA-event timing of BuildMatrix / AddAndCompute per variant (no torch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
ctx = A.Context(0)
ctx.generate_box(3, n)
ctx.build_pattern(1)
for name, v, algo in (("atomic", A.VARIANT_CELLWISE_ATOMIC, A.SPARSITY_FROM_CELLS), ("tiled/cells", A.VARIANT_TILED_GATHER, A.SPARSITY_FROM_CELLS),
                      ("tiled/conn", A.VARIANT_TILED_GATHER, A.SPARSITY_AUTO)):
    ctx.set_sparsity_algorithm(algo)
    tp, ta = [], []
    for _ in range(6):
        ctx.build_pattern(1)
        ctx.assemble(A.OP_POISSON, variant=v)
        t = ctx.last_timings()
        tp.append(t["pattern_ms"]); ta.append(t["assemble_ms"])
    print(f"n={n} {name}: pattern min {min(tp[1:]):.4f} ms, assemble min {min(ta[1:]):.4f} ms   (all pattern {['%.3f' % x for x in tp]}, assemble {['%.3f' % x for x in ta]})")
if len(sys.argv) > 2:
    # elasticity b=3
    import numpy as np
    ctx.generate_box(3, int(sys.argv[2]))
    ctx.build_pattern(3)
    for name, v in (("nodewise", A.VARIANT_NODEWISE), ("tiled", A.VARIANT_TILED_GATHER)):
        tp, ta = [], []
        for _ in range(4):
            ctx.build_pattern(3)
            ctx.assemble(A.OP_ELASTICITY, params=[1.0e6, 8.0e5], fmt=A.FORMAT_BSR, variant=v, layout=A.LAYOUT_PER_ROW)
            t = ctx.last_timings()
            tp.append(t["pattern_ms"]); ta.append(t["assemble_ms"])
        print(f"elasticity n={sys.argv[2]} {name}: pattern min {min(tp[1:]):.4f} ms, assemble min {min(ta[1:]):.4f} ms")
