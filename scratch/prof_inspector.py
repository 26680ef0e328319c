"""One cold tiled assembly (inspector included) for a launch list: python prof_inspector.py [n] [b]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = A.Context(0)
ctx.generate_box(3, n)
ctx.build_pattern(b)
ctx.synchronize()
t0 = time.perf_counter()
if b == 1:
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
else:
    ctx.assemble(A.OP_ELASTICITY, params=[1.0e6, 8.0e5], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=A.LAYOUT_PER_ROW)
ctx.synchronize()
print("first tiled assembly: %.2f ms wall" % (1e3 * (time.perf_counter() - t0)), ctx.inspector_timings())
