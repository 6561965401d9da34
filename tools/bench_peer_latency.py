"""A/B of the cross-GPU barrier inside the fused peer paths (torchrun, one process per GPU): PeerReducedVoxel single-call
latency and PeerCmax ms per evaluation with evk_peer_barrier vs the symmetric-memory handle's own barrier, same box, same run."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
from event_utils_b200 import parallel
n, B, H, W = 50_000_000, 5, 480, 640
g = torch.Generator(device=device).manual_seed(2024 + rank)
x = torch.rand(n, device=device, generator=g) * (W - 1); y = torch.rand(n, device=device, generator=g) * (H - 1)
t = (torch.sort(torch.rand(n, device=device, generator=g)).values + rank) / world
p = (torch.randint(0, 2, (n,), device=device, generator=g) * 2 - 1).float()
cx = torch.rand(n, device=device, generator=g) * 239.0; cy = torch.rand(n, device=device, generator=g) * 179.0
ct = ((torch.sort(torch.rand(n, device=device, generator=g) * (0.05 / world))[0] + rank * (0.05 / world)).double() - 0.05).float()


def timed(fn, iters=10):
    for _ in range(3):
        fn(0)
    best, tot = 1e9, 0.0
    for i in range(iters):
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(i); b.record(); torch.cuda.synchronize()
        el = torch.tensor([a.elapsed_time(b)], device=device)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        best = min(best, float(el)); tot += float(el)
    return best, tot / iters


for kind in ("evk", "torch", "evk", "torch"):
    os.environ["EVK_PEER_BARRIER"] = kind
    pv = parallel.PeerReducedVoxel(B, (H, W), device)
    pc = parallel.PeerCmax(device)
    v = timed(lambda i: pv(x, y, t, p, 0.0, 1.0))
    c = timed(lambda i: pc((45.0 + i, -20.0), cx, cy, ct, p, (180, 240), 1.0, ts_relative=True))
    if rank == 0:
        print("barrier=%-5s world=%d  voxel single call best/mean %.4f / %.4f ms   PeerCmax (f,g) best/mean %.4f / %.4f ms" % (kind, world, v[0], v[1], c[0], c[1]), flush=True)
    del pv, pc
dist.destroy_process_group()
