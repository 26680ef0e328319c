import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
layout = A.LAYOUT_PER_ROW if (len(sys.argv) > 2 and sys.argv[2] == "row") else A.LAYOUT_PER_BLOCK
ctx = A.Context(0)
ctx.generate_box(3, n)
ctx.build_pattern(3)
ta = []
for _ in range(4):
    ctx.build_pattern(3)
    ctx.assemble(A.OP_ELASTICITY, params=[1.0e6, 8.0e5], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
    ta.append(ctx.last_timings()["assemble_ms"])
ctx.synchronize()
print("ok n=%d elasticity tiled assemble ms:" % n, ["%.3f" % x for x in ta], ctx.last_timings())
