"""Where does the first inspector run of a process spend its time?  python prof_inspector2.py [warm]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
warm = len(sys.argv) > 1 and sys.argv[1] == "warm"
def one(n):
    ctx = A.Context(0)
    ctx.generate_box(3, n)
    ctx.build_pattern(1)
    ctx.synchronize()
    t0 = time.perf_counter()
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    ctx.synchronize()
    dt = 1e3 * (time.perf_counter() - t0)
    print("n=%d first tiled assembly %.1f ms wall" % (n, dt), ctx.inspector_timings(), flush=True)
    ctx.close()
if warm:
    one(8)
for n in (256, 256, 120, 256):
    one(n)
