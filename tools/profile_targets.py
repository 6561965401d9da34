"""One launch of each hot kernel on its BASELINE-sized workload, in a fixed order, for
`ncu --set full -k regex:"scatter|hot|count" ...` (profiles/).  Not a benchmark."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_utils_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda")
N = int(os.environ.get("N", 50_000_000))
g = torch.Generator(device=dev).manual_seed(2024)
oob = torch.zeros(1, dtype=torch.int64, device=dev)

# 1/2: voxel 5x480x640, vector-red then scalar-red
B, H, W = 5, 480, 640
x = torch.rand(N, device=dev, generator=g) * (W - 1)
y = torch.rand(N, device=dev, generator=g) * (H - 1)
t = torch.sort(torch.rand(N, device=dev, generator=g)).values
p = (torch.randint(0, 2, (N,), device=dev, generator=g) * 2 - 1).float()
out = torch.empty((B, H, W), device=dev)
ws = torch.empty(L.evk_voxel_workspace_bytes(B, H, W, 0), dtype=torch.uint8, device=dev)
for v in (_lib.VARIANT_VECTOR_RED, _lib.VARIANT_GLOBAL_RED):
    _lib.check(L.evk_voxel_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), N, 0.0, 1.0, B, H, W, v, out.data_ptr(),
                               ws.data_ptr(), ws.numel(), oob.data_ptr(), None))
torch.cuda.synchronize()
del x, y, t, p

# 3/4/5: event image 720x1280 on a Zipf(1.0) stream: plain global reds, smem cache, and integer counts
Hi, Wi = 720, 1280
npx = Hi * Wi
w = 1.0 / torch.arange(1, npx + 1, device=dev, dtype=torch.float64)
cdf = torch.cumsum(w, 0) / w.sum()
ranks = torch.searchsorted(cdf, torch.rand(N, device=dev, generator=g, dtype=torch.float64)).clamp_(max=npx - 1)
pix = torch.randperm(npx, device=dev, generator=g)[ranks]
xi, yi, pi = (pix % Wi).float(), (pix // Wi).float(), torch.ones(N, device=dev)
del w, cdf, ranks, pix
img = torch.empty((Hi + 1, Wi + 1), device=dev)
wsi = torch.empty(L.evk_image_workspace_bytes(Hi + 1, Wi + 1, _lib.BILINEAR), dtype=torch.uint8, device=dev)
for v in (_lib.VARIANT_GLOBAL_RED, _lib.VARIANT_SMEM_TILE):
    _lib.check(L.evk_image_f32(xi.data_ptr(), yi.data_ptr(), pi.data_ptr(), N, Hi, Wi, 0.0, 0.0, v, 0.0, img.data_ptr(), None, 0,
                               oob.data_ptr(), None))
cnt = torch.empty((Hi, Wi), dtype=torch.int32, device=dev)
_lib.check(L.evk_count_u32(xi.data_ptr(), yi.data_ptr(), N, Hi, Wi, 0.0, 0.0, _lib.VARIANT_SMEM_TILE, cnt.data_ptr(), oob.data_ptr(), None))
# 6: bilinear, uniform stream, block vector-red
xb = torch.rand(N, device=dev, generator=g) * (Wi - 1)
yb = torch.rand(N, device=dev, generator=g) * (Hi - 1)
_lib.check(L.evk_image_f32(xb.data_ptr(), yb.data_ptr(), pi.data_ptr(), N, Hi + 1, Wi + 1, float(Wi), float(Hi),
                           _lib.VARIANT_VECTOR_RED | _lib.BILINEAR | _lib.CLIP, 0.0, img.data_ptr(), wsi.data_ptr(), wsi.numel(),
                           oob.data_ptr(), None))
torch.cuda.synchronize()
del xi, yi, pi, xb, yb

# 7/8: fused cmax, f64 parity mode, f+g then f
x64 = torch.rand(N, device=dev, generator=g, dtype=torch.float64) * 239
y64 = torch.rand(N, device=dev, generator=g, dtype=torch.float64) * 179
t64 = torch.sort(torch.rand(N, device=dev, generator=g, dtype=torch.float64)).values * 0.05
p64 = torch.ones(N, device=dev, dtype=torch.float64)
wsc = torch.empty(L.evk_cmax_workspace_bytes(180, 240), dtype=torch.uint8, device=dev)
res = torch.empty(8, dtype=torch.float64, device=dev)
for fl in (_lib.CMAX_WANT_GRAD, 0):
    _lib.check(L.evk_cmax_linvel_variance_f64(x64.data_ptr(), y64.data_ptr(), t64.data_ptr(), p64.data_ptr(), N, 1.0, 45.0, -20.0,
                                              float(t64[-1]), 180, 240, 180, 240, 1.0, fl, res.data_ptr(), None, None,
                                              wsc.data_ptr(), wsc.numel(), None))
torch.cuda.synchronize()
print("done", res.cpu().numpy()[:3])
