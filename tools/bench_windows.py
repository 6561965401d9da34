"""Batched-window voxelisation (row f2): one launch for all windows vs one call per window."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import event_utils_b200 as eu  # noqa: E402
from event_utils_b200.representations.voxel_grid import events_to_voxel_torch, voxel_grids_fixed_n_torch  # noqa: E402

eu.config.check_index_errors = False
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
N, H, W, B = 20_000_000, 180, 240, 5
x = torch.rand(N, device=dev, generator=g) * (W - 1)
y = torch.rand(N, device=dev, generator=g) * (H - 1)
t = torch.sort(torch.rand(N, device=dev, generator=g)).values
p = (torch.randint(0, 2, (N,), device=dev, generator=g) * 2 - 1).float()
for n in (10_000, 50_000, 200_000):
    for _ in range(2):
        grids = voxel_grids_fixed_n_torch(x, y, t, p, B, n, sensor_size=(H, W))
    torch.cuda.synchronize()
    s = time.perf_counter()
    grids = voxel_grids_fixed_n_torch(x, y, t, p, B, n, sensor_size=(H, W))
    torch.cuda.synchronize()
    tb = time.perf_counter() - s
    nw = len(grids)
    s = time.perf_counter()
    loop = [events_to_voxel_torch(x[i:i + n], y[i:i + n], t[i:i + n], p[i:i + n], B, sensor_size=(H, W)) for i in range(0, N - n, n)]
    torch.cuda.synchronize()
    tl = time.perf_counter() - s
    err = max(float((a - b).abs().max()) for a, b in zip(grids[:5], loop[:5]))
    print("windows of %7d events: %5d windows  batched %8.2f ms (%7.1f Mev/s)   per-window calls %8.2f ms (%7.1f Mev/s)  max diff %.2e"
          % (n, nw, tb * 1e3, nw * n / tb / 1e6, tl * 1e3, nw * n / tl / 1e6, err), flush=True)
    del grids, loop
