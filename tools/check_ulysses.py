#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_ulysses.py : one prompt, sequence-sharded over N ranks, through
UlyssesLiteAttention (all_to_all in, peer-store scatter out) must equal the same attention computed on one GPU."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from liteattention_b200 import LiteAttention
from liteattention_b200.dist import UlyssesLiteAttention
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, S, H, D = 1, int(os.environ.get("S", 3000)), int(os.environ.get("H", 8)), 128
assert S % world == 0 and H % world == 0
g = torch.Generator(device=dev).manual_seed(7)               # same seed everywhere: every rank can build the full prompt
q, k, v = (torch.randn(B, S, H, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
sl = S // world
loc = slice(rank * sl, (rank + 1) * sl)
thr = -3.0
ul = UlyssesLiteAttention(lambda: LiteAttention(enable_skipping=True, threshold=thr, max_batch_size=B))
ref = LiteAttention(enable_skipping=True, threshold=thr, max_batch_size=B)       # single-GPU, all heads
ok = True
for step in range(3):
    o_loc = ul(q[:, loc].contiguous(), k[:, loc].contiguous(), v[:, loc].contiguous())
    o_ref = ref(q, k, v)
    torch.cuda.synchronize()
    same = torch.equal(o_loc, o_ref[:, loc])
    ok &= same
    print(f"step {step} rank {rank}: sequence-parallel O == single-GPU O on my tokens: {same} "
          f"(max diff {(o_loc.float() - o_ref[:, loc].float()).abs().max().item():.3g}); "
          f"list sparsity {ul.attn.last_sparsity(B):.3f}", flush=True)
    dist.barrier()
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("ULYSSES", "OK" if int(t) else "MISMATCH")
if os.environ.get("TIME"):
    Sb, Hb = 75600, 40
    qb, kb, vb = (torch.randn(B, Sb // world, Hb, D, device=dev).to(torch.bfloat16) for _ in range(3))
    ulb = UlyssesLiteAttention(lambda: LiteAttention(enable_skipping=False, max_batch_size=B))
    for _ in range(2): ulb(qb, kb, vb)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ulb(qb, kb, vb)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"Wan2.1-14B shape, one prompt over {world} GPUs, dense: {ms.item():.2f} ms per attention call")
dist.destroy_process_group()
