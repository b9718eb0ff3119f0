#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_peer_gather.py : the fused (peer-store) gather must deliver to rank 0 exactly
what each rank computes locally and what the NCCL gather delivers."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention
from liteattention_b200.dist import BatchParallelLiteAttention
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, S, H, D = 1, 3000, 8, 128
g = torch.Generator(device=dev).manual_seed(100 + rank)
q, k, v = (torch.randn(B, S, H, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
mk = lambda: LiteAttention(enable_skipping=True, threshold=-3.0, max_batch_size=B)
peer = BatchParallelLiteAttention(mk, num_heads=H, dst=0, peer_store=True)
nccl = BatchParallelLiteAttention(mk, num_heads=H, num_groups=2, dst=0, peer_store=False)
ok = True
for step in range(3):                       # three chained steps: the skip lists evolve identically on both drivers
    o_p, g_p = peer(q, k, v)
    o_n, g_n = nccl(q, k, v)
    torch.cuda.synchronize(); dist.barrier()
    local = torch.cat(o_n, dim=2)
    if rank == 0:
        for r in range(world):
            ref_r = torch.cat([g_n[gi][r] for gi in range(len(g_n))], dim=2)
            same = torch.equal(g_p[0][r], ref_r)
            ok &= same
            print(f"step {step} rank-{r} slab at dst: peer-store == nccl gather: {same}")
    else:
        pass
    # every rank: what I wrote into the peer slot is what I computed locally
    mine = torch.equal(o_p[0], local)
    ok &= mine
    print(f"step {step} rank {rank}: peer slot readback == local O: {mine}", flush=True)
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("PEER GATHER", "OK" if int(t) else "MISMATCH")
dist.destroy_process_group()
