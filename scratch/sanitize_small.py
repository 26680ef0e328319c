"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): every kernel of the steady-state step, all three
operators through the tiled executors, the chained scan, the connectivity-based BuildMatrix, RHS terms and the PCG."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from arcanefem_b200 import capi as A, mesh as M
ctx = A.Context(0)
for dim, n in ((3, 7), (2, 20)):
    m = M.box_mesh(dim, n)
    ctx.set_mesh(m.dim, m.coords, m.cells)
    for b, op, fmt in ((1, A.OP_POISSON, A.FORMAT_CSR), (dim, A.OP_ELASTICITY, A.FORMAT_BSR)) + (((2, A.OP_BILAPLACIAN, A.FORMAT_BSR),) if dim == 2 else ()):
        for algo in (A.SPARSITY_FROM_CELLS, A.SPARSITY_FROM_CONNECTIVITY):
            ctx.set_sparsity_algorithm(algo)
            for rep in range(3):
                ctx.build_pattern(b)
                ctx.assemble(op, params=[1.0e6, 8.0e5], fmt=fmt, variant=A.VARIANT_TILED_GATHER)
        if op == A.OP_ELASTICITY:  # both vector executors, both value layouts
            for ex in (A.VEC_EXEC_ROWS, A.VEC_EXEC_UNITS, A.VEC_EXEC_AUTO):
                ctx.set_vector_executor(ex)
                for layout in (A.LAYOUT_PER_BLOCK, A.LAYOUT_PER_ROW):
                    ctx.build_pattern(b)
                    ctx.assemble(op, params=[1.0e6, 8.0e5], fmt=fmt, variant=A.VARIANT_TILED_GATHER, layout=layout)
            ctx.build_pattern(b)
            ctx.assemble(op, params=[1.0e6, 8.0e5], fmt=fmt, variant=A.VARIANT_TILED_GATHER)
        ctx.set_sparsity_algorithm(A.SPARSITY_AUTO)
        if op != A.OP_BILAPLACIAN:
            ids = np.nonzero(m.coords[:, dim - 1] == 0.0)[0].astype(np.int32)
            dofs = (ids[:, None] * b + np.arange(b)[None, :]).ravel().astype(np.int32)
            ctx.rhs_reset()
            ctx.rhs_source([1.0] * b)
            ctx.dirichlet_penalty(dofs, np.zeros(dofs.size), 1e30)
            x, it, res = ctx.solve_pcg(rtol=1e-8, max_iter=2000)
            print(dim, b, op, "pcg iterations", it, "residual", res)
ctx.close()
print("sanitize pass done")
