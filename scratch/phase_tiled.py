import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
ctx = A.Context(0)
ctx.generate_box(3, int(sys.argv[1]) if len(sys.argv) > 1 else 120)
ctx.build_pattern(1)
for name, fl in (("both", 0), ("B only", 1 << 16), ("A only", 1 << 17), ("none", 3 << 16)):
    ts = []
    for _ in range(4):
        ctx.reset_values()
        ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, flags=fl)
        ts.append(ctx.last_timings()["assemble_ms"])
    print(name, min(ts[1:]))
