"""Pageable-source e2e (what the reference's callers hand over: ordinary numpy arrays) through evk_voxel_host_f32, and the
raw staging copy (evk_host_copy) beside it, for a few worker counts and with / without non-temporal stores.  Each
configuration runs in its own process (the settings are read once per process).

    python tools/bench_pageable.py            # prints one line per configuration
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import ctypes, sys, time
import numpy as np, torch
sys.path.insert(0, %r)
from event_utils_b200 import _lib, parallel
from event_utils_b200.representations.voxel_grid import events_to_voxel_torch
parallel.bind_to_gpu_numa_node(0)
L = _lib.load()
n = 50_000_000
rng = np.random.default_rng(1)
x = (rng.random(n, dtype=np.float32) * 639); y = (rng.random(n, dtype=np.float32) * 479)
t = np.sort(rng.random(n, dtype=np.float32)); p = (rng.integers(0, 2, n, dtype=np.int8) * 2 - 1).astype(np.float32)
# raw staging copy: 4 x 16 MB blocks into pinned memory, as the pipeline does per chunk
m = 4 << 20
pin = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(4)]
vp = ctypes.c_void_p
d = (vp * 4)(*[b.data_ptr() for b in pin])
best = 1e9
for rep in range(20):
    off = (rep %% 8) * m
    s = (vp * 4)(*[a[off:].ctypes.data for a in (x, y, t, p)])
    t0 = time.perf_counter(); L.evk_host_copy(d, s, 4, m * 4); best = min(best, time.perf_counter() - t0)
copy_gbs = 4 * m * 4 / best / 1e9
ev = [torch.from_numpy(a) for a in (x, y, t, p)]
for _ in range(2): events_to_voxel_torch(*ev, 5, sensor_size=(480, 640))
times = []
for _ in range(5):
    t0 = time.perf_counter(); out = events_to_voxel_torch(*ev, 5, sensor_size=(480, 640)); times.append(time.perf_counter() - t0)
med = sorted(times)[len(times) // 2]
print("copy %%6.1f GB/s (best of 20, 64 MB)   pageable e2e %%7.1f Mev/s (median of 5, %%.1f ms)" %% (copy_gbs, n / med / 1e6, med * 1e3))
''' % ROOT


def main():
    configs = [("default", {})]
    for th in (8, 16, 32):
        configs.append(("threads=%d stream" % th, {"EVK_HOST_THREADS": str(th)}))
    configs.append(("threads=16 memcpy", {"EVK_HOST_THREADS": "16", "EVK_HOST_COPY_STREAM": "0"}))
    configs.append(("threads=32 memcpy", {"EVK_HOST_THREADS": "32", "EVK_HOST_COPY_STREAM": "0"}))
    for name, env in configs:
        out = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=dict(os.environ, **env), timeout=600)
        line = out.stdout.strip().splitlines()[-1] if out.returncode == 0 and out.stdout.strip() else "FAILED: " + out.stderr[-400:]
        print("%-20s %s" % (name, line), flush=True)


if __name__ == "__main__":
    main()
