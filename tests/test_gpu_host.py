"""Host -> device staging (evk_host_upload behind _lib.upload / representations._events.to_device): numpy inputs of every
drop-in function are ordinary pageable memory (the reference passes numpy arrays to each objective call,
lib/contrast_max/objectives.py:211, and loader arrays to the representations, base_dataset.py:446-453)."""
import time

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def test_upload_equals_torch_copy_for_every_size_dtype_and_memory_kind():
    from event_utils_b200 import _lib
    rng = np.random.default_rng(0)
    dev = torch.device("cuda", 0)
    piece = 16 << 20
    cases = []
    for nbytes in (1, 4095, (4 << 20) - 8, 4 << 20, (4 << 20) + 8, piece - 8, piece, piece + 8, 5 * piece + 24, 9 * piece + 8):
        cases.append(torch.from_numpy(rng.integers(0, 256, nbytes, dtype=np.uint8)))
    cases.append(torch.from_numpy(rng.random(3_000_001)))                       # f64, 24 MB
    cases.append(torch.from_numpy(rng.random((1_000_000, 4), dtype=np.float32)))  # 2-D, 16 MB
    cases.append(torch.from_numpy(rng.integers(-300, 300, 7_000_001).astype(np.int16)))
    cases.append(torch.from_numpy(rng.random(6_000_000, dtype=np.float32)).pin_memory())   # pinned: copied directly
    cases.append(torch.from_numpy(rng.random(12_000_000, dtype=np.float32))[::2])          # not contiguous -> torch path
    out = _lib.upload(cases, dev)
    for h, d in zip(cases, out):
        assert d.is_cuda and d.dtype == h.dtype and d.shape == h.shape
        assert torch.equal(d.cpu(), h)
    # again: the bounce slots are reused, and the arrays of one call share rounds
    out = _lib.upload(cases[::-1], dev)
    for h, d in zip(cases[::-1], out):
        assert torch.equal(d.cpu(), h)


def test_numpy_inputs_take_the_upload_path_and_results_do_not_change(oracle):
    """events_to_image_torch / the contrast-maximisation objective on numpy arrays large enough for the bounce path equal
    the oracle exactly as before."""
    from event_utils_b200.contrast_max import objectives as O
    from event_utils_b200.contrast_max.warps import linvel_warp
    from event_utils_b200.representations.image import events_to_image_torch
    rng = np.random.default_rng(5)
    n = 3_000_000                                    # 12 MB per f32 array, 24 MB per f64 array
    x = (rng.random(n) * 239).astype(np.float32)
    y = (rng.random(n) * 179).astype(np.float32)
    p = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
    ref = oracle.image_torch_f32(x, y, p, sensor_size=(180, 240), clip_out_of_range=False)
    out = events_to_image_torch(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(p), device="cuda", sensor_size=(180, 240),
                                clip_out_of_range=False)
    assert np.abs(out.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    xs, ys = rng.random(n) * 238 + 0.5, rng.random(n) * 178 + 0.5
    ts = np.sort(rng.random(n)) * 0.05
    ps = rng.integers(0, 2, n) * 2.0 - 1.0
    fo, go = oracle.cmax_variance((30.0, -20.0), xs, ys, ts, ps, blur_sigma=1.0)
    for prec in ("f64", "f32"):
        O.precision = prec
        O.clear_cache()
        obj = O.variance_objective()
        f = obj.evaluate_function((30.0, -20.0), xs, ys, ts, ps, linvel_warp(), (180, 240), 1.0)
        assert abs(f - fo) <= (1e-5 if prec == "f64" else 1e-3) * abs(fo), (prec, f, fo)
    O.precision = "f64"
    O.clear_cache()


def test_upload_is_faster_than_the_pageable_driver_copy():
    """Not a parity test: the reason the path exists.  1.6 GB of pageable f64 (a 50 M-event contrast-maximisation set)."""
    from event_utils_b200 import _lib
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    arrs = [torch.from_numpy(rng.random(50_000_000)) for _ in range(4)]
    _lib.upload(arrs, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); out = _lib.upload(arrs, dev); t_up = time.perf_counter() - t0
    del out
    t0 = time.perf_counter(); out = [a.to(dev) for a in arrs]; torch.cuda.synchronize(); t_torch = time.perf_counter() - t0
    gb = 4 * 50_000_000 * 8 / 1e9
    print("upload %.1f ms = %.1f GB/s; torch pageable copy %.1f ms = %.1f GB/s" % (t_up * 1e3, gb / t_up, t_torch * 1e3, gb / t_torch))
    assert t_up < t_torch
