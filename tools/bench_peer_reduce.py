"""torchrun --nproc-per-node N tools/bench_peer_reduce.py: latency of ONE sharded voxel build (50 M events
per GPU -> 5x480x640), fused fold + peer all-reduce kernel (parallel.PeerReducedVoxel) against scatter +
fold + NCCL all-reduce (parallel.events_to_voxel_sharded).  CUDA-event time, max over ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_utils_b200.parallel import PeerReducedVoxel, events_to_voxel_sharded  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
N, B, H, W = int(os.environ.get("N", 50_000_000)), 5, 480, 640
g = torch.Generator(device=dev).manual_seed(2024 + rank)
x = torch.rand(N, device=dev, generator=g) * (W - 1)
y = torch.rand(N, device=dev, generator=g) * (H - 1)
t = torch.sort(torch.rand(N, device=dev, generator=g)).values / world + rank / world
p = (torch.randint(0, 2, (N,), device=dev, generator=g) * 2 - 1).float()
t0, dt = 0.0, 1.0
fused = PeerReducedVoxel(B, (H, W), dev)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(iters):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        el = torch.tensor([a.elapsed_time(b)], device=dev)
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        best = min(best, float(el))
    return best


nvls = PeerReducedVoxel(B, (H, W), dev, multicast=True)
ga = fused(x, y, t, p, t0, dt).clone()
gb = events_to_voxel_sharded(x, y, t, p, B, (H, W), t0=t0, dt=dt)
err = float((ga - gb).abs().max() / gb.abs().max())
ms_f = timed(lambda: fused(x, y, t, p, t0, dt))
ms_n = timed(lambda: events_to_voxel_sharded(x, y, t, p, B, (H, W), t0=t0, dt=dt))
if nvls.multicast:
    gc = nvls(x, y, t, p, t0, dt).clone()
    err_mc = float((gc - gb).abs().max() / gb.abs().max())
    ms_mc = timed(lambda: nvls(x, y, t, p, t0, dt))
    if rank == 0:
        print("world %d: NVLS multimem fold+all-reduce %.3f ms, max rel diff vs NCCL %.2e" % (world, ms_mc, err_mc), flush=True)
elif rank == 0:
    print("no multicast support reported by the symmetric-memory backend", flush=True)
if rank == 0:
    print("world %d, %d M events/GPU: fused fold+peer all-reduce %.3f ms | scatter+fold+NCCL all-reduce %.3f ms | max rel diff %.2e"
          % (world, N // 1000000, ms_f, ms_n, err), flush=True)
dist.destroy_process_group()
