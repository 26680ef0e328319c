"""Event-timed AddAndCompute of the vector tiled variant (elasticity b=3)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arcanefem_b200 import capi as A
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 140
layout = int(sys.argv[2]) if len(sys.argv) > 2 else 1
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = A.Context(0, stream=stream.cuda_stream)
info = ctx.generate_box(3, n)
nbr, nnz = ctx.build_pattern(3)
lam = bench.E_MOD * bench.NU / ((1 + bench.NU) * (1 - 2 * bench.NU)); mu = bench.E_MOD / (2 * (1 + bench.NU))
asm = lambda: ctx.assemble(A.OP_ELASTICITY, params=[lam, mu], fmt=A.FORMAT_BSR, variant=A.VARIANT_TILED_GATHER, layout=layout)
asm()
ev = lambda: torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(6):
    ctx.build_pattern(3)
    e0, e1 = ev(), ev(); e0.record(stream); asm(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
bv, _ = bench.algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz, b=3)
print(json.dumps({"n": n, "layout": layout, "lib": os.path.basename(os.environ.get("AFB200_LIB", "default")), "min_ms": min(ts), "frac": bv / (min(ts) * 1e-3) / 1e9 / 6535.7, "inspector": ctx.inspector_timings()}))
