"""Event-timed AddAndCompute of the scalar tiled variant (chained slices vs brick tiles; env AFB_SCALAR_EXEC, AFB_CHAIN_GEOM)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from arcanefem_b200 import capi as A
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = A.Context(0, stream=stream.cuda_stream)
EXEC = {"bricks": 0, "tiles": 0, "chain": 1, "flow": 2}[os.environ.get("AFB_TILED_EXEC", "bricks")]
ctx.set_tiled_executor(EXEC)
info = ctx.generate_box(dim, n)
nbr, nnz = ctx.build_pattern(1)
ev = lambda: torch.cuda.Event(enable_timing=True)
e0, e1 = ev(), ev()
e0.record(stream)
ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
e1.record(stream); e1.synchronize()
first = e0.elapsed_time(e1)
ts = []
for _ in range(8):
    ctx.build_pattern(1)
    e0, e1 = ev(), ev()
    e0.record(stream)
    ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
    e1.record(stream); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
bv, _ = bench.algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz, npc=4 if dim == 3 else 3)
out = {"n": n, "dim": dim, "exec": os.environ.get("AFB_TILED_EXEC", "bricks"), "geom": os.environ.get("AFB_CHAIN_GEOM", "B"), "prefill": os.environ.get("AFB_CHAIN_PREFILL", "1"),
       "first_ms": first, "min_ms": min(ts), "median_ms": float(np.median(ts)), "frac": bv / (min(ts) * 1e-3) / 1e9 / 6535.7, "inspector": ctx.inspector_timings()}
if dim == 3:
    g = bench.golden_digest(f"poisson3d_n{n}")
    if g:
        a, t, _ = bench.matrix_digest(torch, A, ctx, 0, nbr)
        out["rel_err"] = max(abs(a - g["abs_sum"]) / g["abs_sum"], abs(t - g["trace"]) / g["trace"])
print(json.dumps(out))
