"""The reference's OWN callers on top of this repo's kernels: the `sys.modules` aliasing that
INTEGRATION.md section 1 advertises, then the unmodified `optimize_contrast` (lib/contrast_max/events_cmax.py:313-346) and
`BaseVoxelDataset.get_voxel_grid` (lib/data_loaders/base_dataset.py:433-455) from the reference package
(/root/reference in the build container, the unmodified copy in baseline/_ref on the GPU box).  Runs in a subprocess so
that the aliased modules do not leak into the other tests."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, sys, types
import numpy as np, torch
sys.path.insert(0, %(root)r)
from oracle import ref_loader, evk_oracle
assert ref_loader.available(), "reference package not found"
ref_loader._install_stubs()                       # matplotlib / h5py are not in the image; nothing below plots or reads files
import event_utils_b200.representations.voxel_grid as vg, event_utils_b200.representations.image as im
import event_utils_b200.transforms.optic_flow as of
import event_utils_b200.contrast_max.objectives as ob, event_utils_b200.contrast_max.warps as wp
# --- INTEGRATION.md section 1: alias the modules before the reference's callers import them ---
sys.modules["lib.representations.voxel_grid"] = vg
sys.modules["lib.representations.image"] = im
sys.modules["lib.transforms.optic_flow"] = of
sys.modules["lib.contrast_max.objectives"] = ob
sys.modules["lib.contrast_max.warps"] = wp
sys.path.insert(0, ref_loader.REF_ROOT)
ec = importlib.import_module("lib.contrast_max.events_cmax")          # the reference's file, unmodified
bd = importlib.import_module("lib.data_loaders.base_dataset")         # the reference's file, unmodified
assert ec.__file__.startswith(ref_loader.REF_ROOT) and bd.__file__.startswith(ref_loader.REF_ROOT)
assert ec.variance_objective is ob.variance_objective and bd.events_to_voxel_torch is vg.events_to_voxel_torch

# --- the reference's optimize_contrast drives scipy's BFGS over OUR fused objective ---
rng = np.random.default_rng(5)
v_true = np.array([60.0, -35.0])
n_pts, per = 60, 1500
px, py = rng.random(n_pts) * 120 + 60, rng.random(n_pts) * 80 + 50
ts = np.sort(rng.random(n_pts * per) * 0.04)
k = rng.integers(0, n_pts, n_pts * per)
xs = px[k] + v_true[0] * (ts - ts[-1]) + rng.normal(0, 0.3, ts.shape)
ys = py[k] + v_true[1] * (ts - ts[-1]) + rng.normal(0, 0.3, ts.shape)
ps = np.ones_like(ts)
obj, warp = ec.variance_objective(), ec.linvel_warp()
x0 = v_true + np.array([6.0, -5.0])
argmax = ec.optimize_contrast(xs, ys, ts, ps, warp, obj, x0=x0, numeric_grads=False, blur_sigma=1.0, img_size=(180, 240))
f_found = obj.evaluate_function(argmax, xs, ys, ts, ps, warp, (180, 240), 1.0)
f_true, _ = evk_oracle.cmax_variance(tuple(v_true), xs, ys, ts, ps, blur_sigma=1.0)
f_start, _ = evk_oracle.cmax_variance(tuple(x0), xs, ys, ts, ps, blur_sigma=1.0)
f_check, _ = evk_oracle.cmax_variance(tuple(argmax), xs, ys, ts, ps, blur_sigma=1.0)
assert abs(f_found - f_check) <= 1e-5 * abs(f_check), (f_found, f_check)         # the value it reports is the oracle's
assert f_found <= f_start and f_found <= 0.9 * f_true, (f_found, f_start, f_true)  # and it is a (near-)optimal contrast
assert np.abs(np.asarray(argmax) - v_true).max() < 3.0, argmax
print("optimize_contrast ->", argmax, f_found)

# --- BaseVoxelDataset.get_voxel_grid, both channel layouts, on CUDA tensors (the loader's hot call) ---
n, H, W = 300000, 260, 346
ex = (rng.random(n) * (W - 1)).astype(np.float32); ey = (rng.random(n) * (H - 1)).astype(np.float32)
et = np.sort(rng.random(n)).astype(np.float32); ep = (rng.integers(0, 2, n) * 2 - 1).astype(np.float32)
stub = types.SimpleNamespace(num_bins=5, sensor_resolution=(H, W))
dev = [torch.from_numpy(a).cuda() for a in (ex, ey, et, ep)]
grid = bd.BaseVoxelDataset.get_voxel_grid(stub, *dev, combined_voxel_channels=True)
ref = evk_oracle.voxel_f32(ex, ey, et, ep, 5, (H, W))
assert grid.is_cuda and tuple(grid.shape) == (5, H, W)
assert np.abs(grid.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
grid2 = bd.BaseVoxelDataset.get_voxel_grid(stub, *dev, combined_voxel_channels=False)
pos = evk_oracle.voxel_f32(ex, ey, et, (ep > 0).astype(np.float32), 5, (H, W))
neg = evk_oracle.voxel_f32(ex, ey, et, (ep <= 0).astype(np.float32), 5, (H, W))
assert tuple(grid2.shape) == (10, H, W)
assert np.abs(grid2.cpu().numpy() - np.concatenate([pos, neg])).max() <= 1e-5 * max(np.abs(pos).max(), np.abs(neg).max())
print("get_voxel_grid ok")
'''


def test_reference_callers_run_on_the_swapped_modules():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference package present neither at /root/reference nor in baseline/_ref")
    out = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "get_voxel_grid ok" in out.stdout
