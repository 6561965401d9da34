"""The oracle against the UNMODIFIED reference executed in this container (skipped where
/root/reference does not exist, i.e. on the GPU box).  Larger, randomised companions of the
committed golden vectors."""
import numpy as np
import pytest
import torch

from conftest import assert_close_to_max, make_events
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    return ref_loader.load()


@pytest.mark.parametrize("seed,n,B,H,W", [(1, 200000, 5, 260, 346), (2, 50000, 7, 180, 240), (3, 1000, 2, 10, 12)])
def test_voxel_torch_bit_exact(ref, oracle, seed, n, B, H, W):
    x, y, t, p = make_events(seed, n, H, W)
    v = ref.voxel_grid.events_to_voxel_torch(*(torch.from_numpy(a) for a in (x, y, t, p)), B, sensor_size=(H, W)).numpy()
    torch.set_num_threads(1)
    assert_close_to_max(oracle.voxel_f32(x, y, t, p, B, (H, W)), v, 1e-6)


def test_voxel_numpy(ref, oracle):
    rng = np.random.default_rng(1234)
    n = 100000
    xs, ys = rng.integers(0, 346, n), rng.integers(0, 260, n)
    ts = np.sort(rng.random(n))
    ps = rng.integers(0, 2, n) * 2 - 1.0
    v = ref.voxel_grid.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(260, 346))
    assert_close_to_max(oracle.voxel_f64(xs, ys, ts, ps, 5, (260, 346)), v, 1e-12)


@pytest.mark.parametrize("kw", [dict(), dict(interpolation='bilinear'), dict(padding=False),
                                dict(interpolation='bilinear', padding=False), dict(clip_out_of_range=False)])
def test_image_torch(ref, oracle, kw):
    x, y, t, p = make_events(5, 100000, 180, 240, pol="real")
    if kw.get("clip_out_of_range", True):
        x[::7] += 4
    a = ref.image.events_to_image_torch(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(p), **kw).numpy()
    assert_close_to_max(oracle.image_torch_f32(x, y, p, **kw), a, 2e-6)


def test_flow(ref, oracle):
    x, y, t, p = make_events(6, 50000, 180, 240)
    flow = torch.randn(2, 180, 240) * 30
    xw, yw = ref.optic_flow.warp_events_flow_torch(*(torch.from_numpy(a) for a in (x, y, t, p)), flow)
    xo, yo = oracle.warp_flow_f32(x, y, t, flow.numpy())
    assert_close_to_max(xo, xw.numpy(), 1e-6)
    assert_close_to_max(yo, yw.numpy(), 1e-6)


@pytest.mark.parametrize("sigma", [1.0, 0.0, 3.0])
def test_cmax(ref, oracle, sigma):
    x, y, t, p = make_events(7, 100000, 180, 240, dtype=np.float64)
    obj, warp = ref.objectives.variance_objective(), ref.warps.linvel_warp()
    for params in [(30.0, -20.0), (-400.0, 900.0)]:
        f = obj.evaluate_function(params, x, y, t, p, warp, (180, 240), sigma)
        g = obj.evaluate_gradient(params, x, y, t, p, warp, (180, 240), sigma)
        fo, go = oracle.cmax_variance(params, x, y, t, p, blur_sigma=sigma)
        assert abs(fo - f) <= 1e-6 * abs(f)
        iwe, d = ref.objectives.get_iwe(params, x, y, t, p, warp, (180, 240), compute_gradient=True)
        scale = np.sqrt(np.mean((2 * (iwe - iwe.mean())) ** 2) * np.mean(d ** 2))
        assert np.abs(go - g).max() <= 1e-5 * scale


def test_find_new_range_matches_reference(ref):
    """the grid-search driver's interval update (events_cmax.py:160-182), host logic only"""
    import importlib
    ref_cmax = importlib.import_module("lib.contrast_max.events_cmax")
    from event_utils_b200.contrast_max.events_cmax import find_new_range
    rng = np.random.default_rng(0)
    for _ in range(500):
        axis = np.sort(rng.uniform(-200, 200, rng.integers(3, 9)))
        value = rng.choice(axis) if rng.random() < 0.7 else rng.uniform(-250, 250)
        assert np.array_equal(np.asarray(ref_cmax.find_new_range(axis, value)), np.asarray(find_new_range(axis, value)))


def test_window_tables_match_reference_loader(ref):
    """k_events / t_seconds / fixed_frames / between_frames tables of BaseVoxelDataset
    (base_dataset.py:322-417), built here vectorised"""
    import importlib
    bd = importlib.import_module("lib.data_loaders.base_dataset")
    from event_utils_b200.data_loaders import windows as Wn
    rng = np.random.default_rng(3)
    ts = np.sort(rng.uniform(5.0, 7.5, 20000))
    ts[100:110] = ts[100]                                   # repeated stamps

    class Probe(bd.BaseVoxelDataset):                       # the table code only, no recording behind it
        def __init__(self):
            self.num_events, self.t0, self.tk = len(ts), ts[0], ts[-1]
            self.duration = self.tk - self.t0
            self.has_frames = False

        def find_ts_index(self, timestamp):
            return np.searchsorted(ts, timestamp)           # memmap_dataset.py:81-83

    d = Probe()
    for k, w in ((1000, 0), (1500, 500), (333, 33)):
        d.set_voxel_method({"method": "k_events", "k": k, "sliding_window_w": w})
        assert np.array_equal(np.asarray(d.event_indices), Wn.k_event_indices(len(ts), k, w))
    for t, w in ((0.1, 0.0), (0.25, 0.05), (0.013, 0.0)):
        d.set_voxel_method({"method": "t_seconds", "t": t, "sliding_window_t": w})
        assert np.array_equal(np.asarray(d.event_indices), Wn.timeblock_indices(ts, t, w))
    for nf in (1, 7, 40):
        d.set_voxel_method({"method": "fixed_frames", "num_frames": nf})
        assert np.array_equal(np.asarray(d.event_indices), Wn.fixed_frames_indices(ts, nf))
    d.frame_ts = np.concatenate((np.sort(rng.uniform(5.0, 7.5, 30)), [9.0]))
    d.num_frames = len(d.frame_ts)
    d.set_voxel_method({"method": "between_frames"})
    assert np.array_equal(np.asarray(d.event_indices), Wn.between_frame_indices(ts, d.frame_ts))
    with pytest.raises(Exception):
        Wn.check_event_indices(Wn.k_event_indices(1000, 300, 0).tolist() + [[900, 1200]], 1000)


@pytest.mark.parametrize("name", ["rms", "sos", "soe", "moa", "isoa", "sosa", "r1"])
def test_other_objectives_random_scene(ref, oracle, name):
    x, y, t, p = make_events(17, 60000, 180, 240, dtype=np.float64)
    obj, warp = getattr(ref.objectives, name + "_objective")(), ref.warps.linvel_warp()
    for params in [(25.0, -40.0), (-300.0, 120.0)]:
        for sigma in (None, 0.0, 1.7):
            f = obj.evaluate_function(params, x, y, t, p, warp, (180, 240), sigma)
            fo, go = oracle.cmax_objective(name, params, x, y, t, p, blur_sigma=sigma)
            assert abs(fo - f) <= 1e-6 * max(abs(f), 1e-12), (name, params, sigma)
            if name in ("rms", "soe", "isoa", "sosa"):
                g = obj.evaluate_gradient(params, x, y, t, p, warp, (180, 240), sigma)
                assert np.abs(go - g).max() <= 1e-5 * max(np.abs(g).max(), 1e-9), (name, params, sigma)
