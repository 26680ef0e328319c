import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arcanefem_b200 import capi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
ctx = A.Context(0)
ctx.generate_box(3, n)
ctx.build_pattern(1)
ctx.assemble(A.OP_POISSON, variant=A.VARIANT_CELLWISE_ATOMIC)
ref = ctx.to_host(A.ARRAY_VALUES)
rows = ctx.to_host(A.ARRAY_ROWS)
outs = []
for it in range(3):
    ctx.reset_values()
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    v = ctx.to_host(A.ARRAY_VALUES)
    outs.append(v)
    bad = np.nonzero(np.abs(v - ref) > 1e-10)[0]
    print("iter", it, "bad entries", bad.size, "first", bad[:5], "last", bad[-5:], "zeros", int((v == 0).sum()))
    if bad.size:
        r = np.searchsorted(rows, bad, side="right") - 1
        print("   bad rows range", r.min(), r.max(), "distinct rows", np.unique(r).size)
print("identical 1,2:", np.array_equal(outs[1], outs[2]), "0,1:", np.array_equal(outs[0], outs[1]))
