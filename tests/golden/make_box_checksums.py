#!/usr/bin/env python
"""Full-size known answers for bench.py and the `-m gpu` full-size tests.

Runs the CPU oracle (oracle/afb_oracle.c, the restatement of the reference's assembly) on the
BASELINE boxes and records, per configuration, size-independent digests of the assembled
matrix: sum |a_ij|, sum of the diagonal (trace), nnz, and the values of a fixed sample of rows.
bench.py compares the GPU result of every run (any number of GPUs) against these numbers
(relative 1e-12), so a full-size run carries its own correctness evidence.

    python tests/golden/make_box_checksums.py [c2 c4 c3 ...]      # ~ minutes, tens of GB of host RAM for c4 / c3

Output: tests/golden/box_checksums.json (merged with what is already there).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from arcanefem_b200 import mesh as M  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "box_checksums.json")
E_MOD, NU = 21e5, 0.28  # modules/elasticity/inputs/bar.3D.Dirichlet.bodyForce.arc:25-26

CONFIGS = {
    # name: (dim, n, op, b)
    "poisson3d_n24": (3, 24, O.OP_POISSON, 1),
    "poisson3d_n120": (3, 120, O.OP_POISSON, 1),
    "poisson3d_n256": (3, 256, O.OP_POISSON, 1),
    "elasticity3d_n24": (3, 24, O.OP_ELASTICITY, 3),
    "elasticity3d_n100": (3, 100, O.OP_ELASTICITY, 3),
    "elasticity3d_n203": (3, 203, O.OP_ELASTICITY, 3),
}
ALIAS = {"c2": "poisson3d_n120", "c4": "poisson3d_n256", "c3": "elasticity3d_n203", "small": "poisson3d_n24", "small3": "elasticity3d_n24", "e100": "elasticity3d_n100"}


def sample_rows(nb_row, count=64):
    """A fixed, size-dependent but seed-free sample of rows: evenly spread, plus first and last."""
    idx = np.unique(np.concatenate([[0, nb_row - 1], (np.arange(count, dtype=np.int64) * 2654435761 % nb_row)]))
    return idx.astype(np.int64)


def digest(name):
    dim, n, op, b = CONFIGS[name]
    t0 = time.time()
    m = M.box_mesh(dim, n)
    rows, cols = O.build_pattern(m.npc, m.nb_node, m.cells)
    params = None
    if op == O.OP_ELASTICITY:
        params = O.lame(E_MOD, NU)
    vals = O.assemble(m.dim, m.coords, m.cells, rows, cols, op=op, form=O.FORM_COMPACT, params=params, layout=O.LAYOUT_PER_BLOCK)
    nnz = int(rows[-1])
    counts = np.diff(rows).astype(np.int64)
    rid = np.repeat(np.arange(m.nb_node, dtype=np.int32), counts)
    diag = np.nonzero(cols == rid)[0]
    del rid
    if b == 1:
        trace = float(np.sum(vals[diag]))
    else:
        blk = vals.reshape(nnz, b * b)[diag]
        trace = float(np.sum(blk[:, [i * b + i for i in range(b)]]))
    sr = sample_rows(m.nb_node)
    samples = {}
    for r in sr:
        lo, hi = int(rows[r]), int(rows[r + 1])
        samples[str(int(r))] = {"cols": cols[lo:hi].tolist(), "vals": vals[lo * b * b:hi * b * b].tolist()}
    out = {"dim": dim, "n": n, "op": int(op), "b": b, "nb_cell": int(m.nb_cell), "nb_node": int(m.nb_node), "nnz": nnz,
           "abs_sum": float(np.sum(np.abs(vals))), "trace": trace, "sum": float(np.sum(vals)),
           "params": None if params is None else [float(x) for x in params],
           "jitter": 0.2, "seed": 12345, "form": "FORM_COMPACT, per-block layout", "sample_rows": samples,
           "oracle_seconds": round(time.time() - t0, 1)}
    return out


def main():
    names = [ALIAS.get(a, a) for a in sys.argv[1:]] or list(CONFIGS)
    res = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            res = json.load(f)
    for nm in names:
        print(nm, "...", flush=True)
        res[nm] = digest(nm)
        print("  ", {k: v for k, v in res[nm].items() if k != "sample_rows"}, flush=True)
        with open(OUT, "w") as f:
            json.dump(res, f, indent=0, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
