"""Per-cell coefficient (afb_set_cell_coefficient) through the three variants at C2 / C4: AddAndCompute per variant, plain tiled executor beside it."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arcanefem_b200 import capi as A

def run(n, reps=4):
    ctx = A.Context(0)
    info = ctx.generate_box(3, n)
    ctx.build_pattern(1)
    cc = 10.0 ** np.random.default_rng(3).uniform(-1, 1, info["nb_cell"])
    out = {"n": n, "cells": info["nb_cell"]}
    for label, coef, variants in (("plain", None, [("tiled", A.VARIANT_TILED_GATHER)]),
                                  ("coef", cc, [("atomic", A.VARIANT_CELLWISE_ATOMIC), ("nodewise", A.VARIANT_NODEWISE), ("tiled", A.VARIANT_TILED_GATHER)])):
        ctx.set_cell_coefficient(coef)
        for vname, v in variants:
            ts = []
            for _ in range(reps + 1):
                ctx.build_pattern(1)
                ctx.assemble(A.OP_POISSON, variant=v)
                ts.append(ctx.last_timings()["assemble_ms"])
            out[f"{label}:{vname}_ms"] = min(ts[1:])
    ctx.close()
    print(json.dumps(out), flush=True)

for n in [int(a) for a in sys.argv[1:]] or [120, 256]:
    run(n)
