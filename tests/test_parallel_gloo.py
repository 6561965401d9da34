"""world_size-2 gloo test of the sharded voxel path (host logic + collective), CPU only: the local
kernel is replaced by the oracle, everything else is the code the NCCL run executes."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from event_utils_b200.parallel import events_to_voxel_sharded, global_time_span, shard_bounds
    from oracle import evk_oracle as O
    rng = np.random.default_rng(5)
    n, B, H, W = 30001, 5, 20, 28
    x = (rng.random(n) * (W - 1)).astype(np.float32)
    y = (rng.random(n) * (H - 1)).astype(np.float32)
    t = np.sort(rng.random(n)).astype(np.float32) + 3.0
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
    lo, hi = shard_bounds(n, world, rank)

    def compute(xs, ys, ts, ps, t0, dt, B, H, W):
        if xs.numel() == 0:
            return torch.zeros((B, H, W))
        return torch.from_numpy(O.voxel_f32(xs.numpy(), ys.numpy(), ts.numpy(), ps.numpy(), B, (H, W), t0=t0, dt=dt))

    sh = [torch.from_numpy(a[lo:hi]) for a in (x, y, t, p)]
    t0, dt = global_time_span(sh[2])
    assert t0 == float(t[0]) and abs(dt - float(np.float32(t[-1]) - np.float32(t[0]))) == 0.0
    grid = events_to_voxel_sharded(*sh, B, (H, W), compute=compute)
    full = O.voxel_f32(x, y, t, p, B, (H, W))
    err = np.abs(grid.numpy() - full).max() / np.abs(full).max()
    # an empty shard on one rank must still work
    e = [torch.from_numpy(a[:0]) for a in (x, y, t, p)] if rank == 1 else [torch.from_numpy(a) for a in (x, y, t, p)]
    grid2 = events_to_voxel_sharded(*e, B, (H, W), compute=compute)
    err2 = np.abs(grid2.numpy() - full).max() / np.abs(full).max()
    np.save(os.path.join(out_dir, "err%d.npy" % rank), np.array([err, err2, lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_voxel_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    spans = []
    for r in range(world):
        err, err2, lo, hi = np.load(tmp_path / ("err%d.npy" % r))
        assert err <= 1e-6 and err2 <= 1e-6
        spans.append((int(lo), int(hi)))
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 30001


def test_shard_bounds_cover():
    sys.path.insert(0, ROOT)
    from event_utils_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 400_000_001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
