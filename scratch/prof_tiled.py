import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
ctx = A.Context(0)
ctx.generate_box(3, int(sys.argv[1]) if len(sys.argv) > 1 else 120)
ctx.build_pattern(1)
for _ in range(3):
    ctx.build_pattern(1)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
ctx.synchronize()
print("ok", ctx.last_timings())
