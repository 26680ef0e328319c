"""Host-side timeline of the inspector (AFB_INSPECTOR_TRACE=1): python prof_inspector3.py [n ...]"""
import sys, os, time
os.environ["AFB_INSPECTOR_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
def one(n):
    ctx = A.Context(0)
    ctx.generate_box(3, n)
    ctx.build_pattern(1)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    ctx.synchronize()
    print("n=%d first tiled assembly %.2f ms wall" % (n, 1e3 * (time.perf_counter() - t0)), ctx.inspector_timings(), file=sys.stderr, flush=True)
    ctx.close()
for n in [int(a) for a in sys.argv[1:]] or [120, 120, 256, 120]:
    one(n)
