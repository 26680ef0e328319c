"""torchrun script: exchange-only time (ranks synchronised before every exchange) of the p2p and nccl transports."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from arcanefem_b200 import capi as A, mesh as M
from arcanefem_b200.distributed import DistributedAssembly
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(round(120 * world ** (1 / 3)))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
for transport in ("p2p", "nccl"):
    ctx = A.Context(lr, stream=stream.cuda_stream)
    k_lo, k_hi = M.slab_layers(n, world, rank)
    info = ctx.generate_box(3, n, k_lo=k_lo, k_hi=k_hi, ghost_cell_layer=True)
    ctx.build_pattern(1)
    gid, owner_rel, nb_own, _, _ = M.box_slab_numbering(3, n, k_lo, k_hi, True)
    da = DistributedAssembly(ctx, rank, world, gid, (rank + owner_rel).astype(np.int32), nb_own, lr, transport=transport)
    for _ in range(3):
        ctx.build_pattern(1)
        da.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, mode="exchange")
    ts, ta = [], []
    for _ in range(10):
        ctx.build_pattern(1)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        ctx.assemble(A.OP_POISSON, variant=A.VARIANT_TILED_GATHER, flags=A.FLAG_OWN_CELLS_ONLY | A.FLAG_ALL_ROWS)
        e1.record(stream)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e1b = torch.cuda.Event(enable_timing=True); e1b.record(stream)
        da.plan.exchange()
        da.wait()
        e2.record(stream)
        torch.cuda.synchronize()
        ta.append(e0.elapsed_time(e1)); ts.append(e1b.elapsed_time(e2))
    t = torch.tensor([min(ts), min(ta)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{transport}: exchange-only {t[0].item()*1e3:.1f} us (max over ranks of min), assembly {t[1].item()*1e3:.1f} us, bytes {da.plan.bytes_per_exchange()}", flush=True)
    if transport == "p2p":
        assert ctx.p2p_status() == 0
        ctx.p2p_disconnect()
    dist.barrier()
    ctx.close()
dist.destroy_process_group()
