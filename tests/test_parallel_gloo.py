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


def _cmax_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from event_utils_b200.parallel import cmax_variance_sharded, global_last_timestamp, shard_bounds
    from oracle import evk_oracle as O
    rng = np.random.default_rng(11)
    n = 40001
    x = rng.uniform(0, 240, n)
    y = rng.uniform(0, 180, n)
    t = np.sort(rng.uniform(0, 0.05, n)) + 1.0
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.float64)
    params, img_size = (35.0, -22.0), (180, 240)
    lo, hi = shard_bounds(n, world, rank)

    def images(params, xs, ys, ts, ps, t_ref, img_size, want_grad, use_polarity):
        out = torch.zeros((3, 181, 241))
        if xs.numel():
            iwe, d = O.iwe_linvel(params, xs.numpy(), ys.numpy(), ts.numpy(), ps.numpy(), img_size, compute_gradient=want_grad,
                                  use_polarity=use_polarity, t_ref=t_ref)
            out[0] = torch.from_numpy(iwe)
            if want_grad:
                out[1:] = torch.from_numpy(d)
        return out, torch.zeros(1, dtype=torch.int64)

    def tail(images, blur_sigma, want_grad):
        a = images.numpy()
        return O.variance_f(a[0], blur_sigma), (O.variance_g(a[0], a[1:], blur_sigma) if want_grad else np.zeros(2))

    sh = [torch.from_numpy(a[lo:hi]) for a in (x, y, t, p)]
    assert global_last_timestamp(sh[2]) == float(t[-1])
    f, g = cmax_variance_sharded(params, *sh, img_size, 1.0, compute_images=images, compute_tail=tail)
    f_ref, g_ref = O.cmax_variance(params, x, y, t, p, img_size, 1.0)
    # an empty shard on rank 0
    e = [torch.from_numpy(a[:0] if rank == 0 else a) for a in (x, y, t, p)]
    f2, g2 = cmax_variance_sharded(params, *e, img_size, 1.0, compute_images=images, compute_tail=tail)
    np.save(os.path.join(out_dir, "cmax%d.npy" % rank), np.array([f, g[0], g[1], f_ref, g_ref[0], g_ref[1], f2, g2[0], g2[1]]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_cmax_world2(tmp_path):
    world = 2
    mp.spawn(_cmax_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = [np.load(tmp_path / ("cmax%d.npy" % r)) for r in range(world)]
    assert np.array_equal(rows[0], rows[1])           # every rank returns the same numbers
    f, g, f_ref, g_ref, f2, g2 = rows[0][0], rows[0][1:3], rows[0][3], rows[0][4:6], rows[0][6], rows[0][7:9]
    assert abs(f - f_ref) <= 1e-5 * abs(f_ref) and abs(f2 - f_ref) <= 1e-5 * abs(f_ref)
    scale = np.abs(g_ref).max()
    assert np.abs(g - g_ref).max() <= 1e-4 * scale and np.abs(g2 - g_ref).max() <= 1e-4 * scale


def _image_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from event_utils_b200.parallel import events_to_image_sharded, shard_bounds
    from oracle import evk_oracle as O
    rng = np.random.default_rng(11)
    n, H, W = 20001, 24, 40
    x = (rng.random(n) * (W - 1)).astype(np.float32)
    y = (rng.random(n) * (H - 1)).astype(np.float32)
    x[::3], y[::3] = 7.0, 5.0                                   # a hot pixel
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
    lo, hi = shard_bounds(n, world, rank)
    errs = []
    for interp in (None, 'bilinear'):
        def compute(xs, ys, ps, sensor_size, interpolation):
            return torch.from_numpy(O.image_torch_f32(xs.numpy(), ys.numpy(), ps.numpy(), sensor_size=sensor_size, interpolation=interpolation,
                                                      clip_out_of_range=interpolation == 'bilinear'))
        img = events_to_image_sharded(*(torch.from_numpy(a[lo:hi]) for a in (x, y, p)), (H, W), interp, compute=compute)
        full = O.image_torch_f32(x, y, p, sensor_size=(H, W), interpolation=interp, clip_out_of_range=interp == 'bilinear')
        errs.append(np.abs(img.numpy() - full).max() / np.abs(full).max())
    np.save(os.path.join(out_dir, "img%d.npy" % rank), np.array(errs))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_image_world2(tmp_path):
    """events_to_image_sharded: the sum over shards of the event image == the image of the whole stream (nearest: exact
    integers; bilinear: 1e-6), with the oracle as local kernel"""
    world = 2
    mp.spawn(_image_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        e = np.load(tmp_path / ("img%d.npy" % r))
        assert e[0] == 0.0 and e[1] <= 1e-6, e


def test_numa_binding_is_a_no_op_without_nvml_device():
    """bind_to_gpu_numa_node never raises and changes nothing where there is no GPU / NVML"""
    from event_utils_b200.parallel import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    got = bind_to_gpu_numa_node(0)
    if not torch.cuda.is_available():
        assert got is None and os.sched_getaffinity(0) == before


def test_shard_bounds_cover():
    sys.path.insert(0, ROOT)
    from event_utils_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 400_000_001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
