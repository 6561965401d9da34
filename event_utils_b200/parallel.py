"""Multi-GPU voxel build: one process per GPU, events partitioned across ranks, ONE sum
all-reduce of the (B,H,W) grid over NCCL / NVLink.

Every output of the hot path is a sum over events (index_put_(accumulate=True), reference
image.py:95), so any partition of the events works.  The only global quantities are the two
scalars voxel_grid.py:133-134 derives from the whole stream (first timestamp, duration): they
are agreed with two scalar all-reduces (MIN / MAX) -- or passed in when known a priori -- and
handed to the kernel as explicit t0 / dt.  With contiguous time shards each rank touches only
about (B-1)/G + 1 bin planes, but the reduce is over the whole grid (6.1 MB at 5x480x640, latency
bound on NVLink), one collective per call.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world_size, rank):
    """Contiguous shard [lo, hi) of n events for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def global_time_span(ts_local, group=None):
    """(t0, dt) of the WHOLE stream from each rank's (possibly empty) time-sorted shard."""
    dev = ts_local.device
    if ts_local.numel():
        lohi = torch.stack((ts_local[0], -ts_local[-1])).to(torch.float32)
    else:
        lohi = torch.tensor([float("inf"), float("inf")], dtype=torch.float32, device=dev)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(lohi, op=dist.ReduceOp.MIN, group=group)
    first, neg_last = lohi.tolist()
    import numpy as np
    t0 = np.float32(first)
    dt = np.float32(-neg_last) - t0      # f32 subtraction, like ts[-1]-ts[0] on f32 tensors
    return float(t0), float(dt)


def _voxel_local_cuda(xs, ys, ts, ps, t0, dt, B, H, W):
    from .representations.voxel_grid import _voxel_device
    if xs.numel() == 0:
        return torch.zeros((B, H, W), dtype=torch.float32, device=xs.device)
    return _voxel_device(xs, ys, ts, ps, t0, dt, B, H, W)


def events_to_voxel_sharded(xs, ys, ts, ps, B, sensor_size=(180, 240), group=None, t0=None, dt=None,
                            compute=None):
    """
    Voxel grid of a stream whose events are spread over the ranks of `group` (each rank passes
    ITS shard: contiguous f32 tensors on its own GPU).  Every rank returns the full (B,H,W) grid.
    @param t0, dt global first timestamp / duration if already known (skips the two scalar reduces)
    @param compute local kernel `(xs,ys,ts,ps,t0,dt,B,H,W) -> tensor`; defaults to the CUDA path
           (tests inject the CPU oracle to exercise the collective logic with gloo)
    """
    H, W = int(sensor_size[0]), int(sensor_size[1])
    if t0 is None or dt is None:
        t0, dt = global_time_span(ts, group)
    compute = compute or _voxel_local_cuda
    grid = compute(xs, ys, ts, ps, t0, dt, int(B), H, W)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(grid, op=dist.ReduceOp.SUM, group=group)
    return grid


class ShardedVoxelStream:
    """Back-to-back sharded voxel builds (a data loader voxelising window after window): the sum
    all-reduce of build k runs on a communication stream and overlaps the scatter kernel of build
    k+1, which the NVLink transfer (6 MB) and its launch latency would otherwise serialise with.
    `submit()` returns the grid buffer and the event that marks its all-reduce complete; buffers are
    recycled round-robin (`depth` in flight)."""

    def __init__(self, B, sensor_size, device, group=None, depth=2):
        self.B, self.H, self.W = int(B), int(sensor_size[0]), int(sensor_size[1])
        self.device, self.group = torch.device(device), group
        self.grids = [torch.empty((self.B, self.H, self.W), dtype=torch.float32, device=self.device) for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.comm = torch.cuda.Stream(self.device)
        self.k = 0
        self.multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def submit(self, xs, ys, ts, ps, t0, dt):
        from . import _lib
        from .representations.voxel_grid import _voxel_device
        i = self.k % len(self.grids)
        self.k += 1
        cur = torch.cuda.current_stream(self.device)
        if self.k > len(self.grids):
            cur.wait_event(self.done[i])            # the buffer's previous all-reduce must have finished
        _voxel_device(xs, ys, ts, ps, t0, dt, self.B, self.H, self.W, out=self.grids[i])
        if self.multi:
            ready = torch.cuda.Event()
            ready.record(cur)
            self.comm.wait_event(ready)
            with torch.cuda.stream(self.comm):
                dist.all_reduce(self.grids[i], op=dist.ReduceOp.SUM, group=self.group)
                self.done[i].record(self.comm)
        else:
            self.done[i].record(cur)
        return self.grids[i], self.done[i]

    def drain(self):
        """Make the current stream wait for every outstanding all-reduce."""
        cur = torch.cuda.current_stream(self.device)
        for ev in self.done[: min(self.k, len(self.done))]:
            cur.wait_event(ev)
