"""BASELINE configs C2/C3/C4 on one GPU: BuildMatrix / AddAndCompute per variant (CUDA events of the C ABI)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A
PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def algorithmic_bytes(nb_cell, nb_node, nnz, b=1, npc=4):
    return 4 * npc * nb_cell + 24 * nb_node + 8 * b * b * nnz + 4 * nnz + 4 * (nb_node + 1), 4 * npc * nb_cell + 4 * nnz + 4 * (nb_node + 1)

def run(name, n, b, op, variants, params=None, layout=A.LAYOUT_PER_ROW, reps=5, dim=3, p2=False, fmt=None):
    ctx = A.Context(0)
    if p2:
        from arcanefem_b200 import mesh as M
        m = M.to_p2(M.box_mesh(dim, n))
        ctx.set_mesh(dim, m.coords, m.cells)
        info = ctx.mesh_info()
    else:
        info = ctx.generate_box(dim, n)
    nbr, nnz = ctx.build_pattern(b)
    bv, bp = algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz, b, info["npc"])
    out = {"config": name, "n": n, "cells": info["nb_cell"], "nodes": info["nb_node"], "block_nnz": nnz, "b": b, "alg_bytes_values": bv, "alg_bytes_pattern": bp, "variants": {}}
    for vname, v in variants:
        tp, ta = [], []
        for _ in range(reps + 1):
            ctx.build_pattern(b)
            f = fmt if fmt is not None else (A.FORMAT_BSR if b > 1 else A.FORMAT_CSR)
            ctx.assemble(op, params=params, fmt=f, variant=v, layout=layout if b > 1 else A.LAYOUT_PER_BLOCK)
            t = ctx.last_timings()
            tp.append(t["pattern_ms"]); ta.append(t["assemble_ms"])
        p, a = min(tp[1:]), min(ta[1:])
        out["variants"][vname] = {"build_matrix_ms": p, "add_and_compute_ms": a, "elements_per_s": info["nb_cell"] / ((p + a) * 1e-3),
                                  "values_GBs": bv / a / 1e6, "values_frac_of_peak": bv / a / 1e6 / PEAK, "pattern_GBs": bp / p / 1e6}
    ctx.close()
    print(json.dumps(out), flush=True)

which = sys.argv[1:] or ["c2", "c3", "c4"]
if "c2" in which:
    run("C2 Poisson b=1", 120, 1, A.OP_POISSON, [("atomic", A.VARIANT_CELLWISE_ATOMIC), ("nodewise", A.VARIANT_NODEWISE), ("tiled", A.VARIANT_TILED_GATHER)])
if "c3" in which:
    run("C3 elasticity b=3", 203, 3, A.OP_ELASTICITY, [("nodewise", A.VARIANT_NODEWISE), ("tiled", A.VARIANT_TILED_GATHER)], params=[1.0e6, 8.0e5], reps=3)
if "c4" in which:
    run("C4 Poisson b=1", 256, 1, A.OP_POISSON, [("atomic", A.VARIANT_CELLWISE_ATOMIC), ("tiled", A.VARIANT_TILED_GATHER)], reps=3)

if "c5" in which:
    T = [("atomic", A.VARIANT_CELLWISE_ATOMIC), ("nodewise", A.VARIANT_NODEWISE), ("tiled", A.VARIANT_TILED_GATHER)]
    run("C5 2-D Poisson P1 b=1 CSR", 4096, 1, A.OP_POISSON, T, dim=2, reps=3)
    run("C5 2-D Poisson P1 b=1 COO", 4096, 1, A.OP_POISSON, [T[0], T[2]], dim=2, reps=3, fmt=A.FORMAT_COO)
    run("C5 2-D elasticity P1 b=2 BSR", 4096, 2, A.OP_ELASTICITY, T, params=[1.0e6, 8.0e5], dim=2, reps=3)
    run("C5 2-D bilaplacian P1 b=2 BSR", 4096, 2, A.OP_BILAPLACIAN, T, dim=2, reps=3)
    run("C5 2-D Poisson P2 (Tri6) b=1 CSR", 1024, 1, A.OP_POISSON, T[:2], dim=2, reps=3, p2=True)
