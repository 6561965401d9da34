"""Seeded generator of small adversarial event sets shared by the CPU (oracle vs reference) and GPU
(CUDA vs oracle) edge-case tests: tiny sensors, coordinates around and beyond the borders (negative
wrap, exact integers, out of range), repeated / identical timestamps, zero / fractional / NaN
polarities, single events."""
import numpy as np


def cases(seed, count):
    rng = np.random.default_rng(seed)
    for k in range(count):
        H, W = int(rng.integers(2, 18)), int(rng.integers(2, 22))
        B = int(rng.integers(1, 10))
        n = int(rng.integers(1, 260))
        spread = rng.choice([0.0, 0.0, 1.5, 3.0])            # how far beyond the border coordinates may go
        x = rng.uniform(-spread, W - 1 + spread, n)
        y = rng.uniform(-spread, H - 1 + spread, n)
        snap = rng.random(n) < 0.3                            # exact integers (incl. borders)
        x[snap], y[snap] = np.round(x[snap]), np.round(y[snap])
        t = np.sort(rng.random(n)) * rng.choice([1.0, 1e-3, 50.0])
        if rng.random() < 0.2:
            t[: n // 2] = t[0]                                # repeated stamps
        if rng.random() < 0.1:
            t[:] = t[0]                                       # dt == 0 -> NaN weights
        p = rng.choice([-1.0, 1.0, 0.0, 0.37, 2.5], n)
        yield dict(k=k, B=B, H=H, W=W, x=x.astype(np.float32), y=y.astype(np.float32), t=t.astype(np.float32),
                   p=p.astype(np.float32), clip=bool(rng.random() < 0.5), padding=bool(rng.random() < 0.5))
