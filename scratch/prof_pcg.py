"""Jacobi-PCG on the assembled C2 Poisson system: time per iteration and algorithmic GB/s."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arcanefem_b200 import capi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
ctx = A.Context(0)
info = ctx.generate_box(3, n)
nbr, nnz = ctx.build_pattern(1)
ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER)
ids = np.arange((n + 1) ** 2, dtype=np.int32)          # face z = 0
ctx.set_dirichlet_nodes(ids)
ctx.rhs_reset(); ctx.rhs_source([5.5]); ctx.dirichlet_penalty(ids, np.full(ids.size, 0.5), 1e31)
for iters in (64, 256):
    ctx.synchronize(); t0 = time.perf_counter()
    try:
        x, it, res = ctx.solve_pcg(rtol=0.0, atol=0.0, max_iter=iters)
    except A.AfbError:
        it = iters
    dt = time.perf_counter() - t0
t64 = None
ctx.synchronize(); t0 = time.perf_counter()
try: ctx.solve_pcg(rtol=0.0, atol=0.0, max_iter=64)
except A.AfbError: pass
t64 = time.perf_counter() - t0
ctx.synchronize(); t0 = time.perf_counter()
try: ctx.solve_pcg(rtol=0.0, atol=0.0, max_iter=576)
except A.AfbError: pass
t576 = time.perf_counter() - t0
per_it = (t576 - t64) / 512
bytes_it = nnz * 12 + 4 * (nbr + 1) + 8 * nbr * 2 + 8 * nbr * 6 + 8 * nbr * 3
x, it, res = ctx.solve_pcg(rtol=1e-10, max_iter=20000)
print(json.dumps({"n": n, "rows": nbr, "nnz": nnz, "ms_per_iteration": per_it * 1e3, "algorithmic_GBs": bytes_it / per_it / 1e9, "iterations_to_1e-10": it, "residual": res,
                  "x_min_max": [float(x.min()), float(x.max())]}))
