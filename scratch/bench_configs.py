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

def run(name, n, b, op, variants, params=None, layout=A.LAYOUT_PER_ROW, reps=5):
    ctx = A.Context(0)
    info = ctx.generate_box(3, n)
    nbr, nnz = ctx.build_pattern(b)
    bv, bp = algorithmic_bytes(info["nb_cell"], info["nb_node"], nnz, b)
    out = {"config": name, "n": n, "cells": info["nb_cell"], "nodes": info["nb_node"], "block_nnz": nnz, "b": b, "alg_bytes_values": bv, "alg_bytes_pattern": bp, "variants": {}}
    for vname, v in variants:
        tp, ta = [], []
        for _ in range(reps + 1):
            ctx.build_pattern(b)
            ctx.assemble(op, params=params, fmt=A.FORMAT_BSR if b > 1 else A.FORMAT_CSR, variant=v, layout=layout if b > 1 else A.LAYOUT_PER_BLOCK)
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
