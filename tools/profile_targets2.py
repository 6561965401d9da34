"""One launch of each kernel of the rows next to the headline path (timestamp images, neg/pos voxels,
dense-flow warp and its objective, f32 cmax, RobustNorm) on BASELINE-sized inputs, for
`ncu --set full -k regex:"tsimg_scatter|warp_flow|voxel_scatter|cmax_scatter|select_hist|robust_norm"`.
Not a benchmark."""
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
B, H, W = 5, 480, 640
x = torch.rand(N, device=dev, generator=g) * (W - 1)
y = torch.rand(N, device=dev, generator=g) * (H - 1)
t = torch.sort(torch.rand(N, device=dev, generator=g)).values
p = (torch.randint(0, 2, (N,), device=dev, generator=g) * 2 - 1).float()

# 1: timestamp images (two v4 reductions per event)
tp, tn = torch.empty((H + 1, W + 1), device=dev), torch.empty((H + 1, W + 1), device=dev)
wst = torch.empty(L.evk_timestamp_image_workspace_bytes(H + 1, W + 1), dtype=torch.uint8, device=dev)
_lib.check(L.evk_timestamp_image_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), N, 0.0, 1.0, H + 1, W + 1, float(W), float(H),
                                     _lib.CLIP, tp.data_ptr(), tn.data_ptr(), wst.data_ptr(), wst.numel(), oob.data_ptr(), None))
# 2: neg/pos voxel grids in one pass
outnp = torch.empty((2, B, H, W), device=dev)
wsnp = torch.empty(2 * L.evk_voxel_workspace_bytes(B, H, W, 0), dtype=torch.uint8, device=dev)
_lib.check(L.evk_voxel_negpos_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), N, 0.0, 1.0, B, H, W, _lib.AUTO_SPAN,
                                  outnp.data_ptr(), wsnp.data_ptr(), wsnp.numel(), oob.data_ptr(), None))
# 3: dense-flow warp, interleaved gather
flow = torch.randn(2, H, W, device=dev) * 30
xw, yw = torch.empty_like(x), torch.empty_like(y)
wsf = torch.empty(L.evk_warp_flow_workspace_bytes(H, W), dtype=torch.uint8, device=dev)
_lib.check(L.evk_warp_flow_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), N, flow.data_ptr(), H, W, 1.0, xw.data_ptr(), yw.data_ptr(),
                               wsf.data_ptr(), wsf.numel(), None))
# 4: RobustNorm of the voxel grid
grid = outnp[0].contiguous()
normed = torch.empty_like(grid)
wsr = torch.empty(L.evk_robust_norm_workspace_bytes(), dtype=torch.uint8, device=dev)
nel = grid.numel()
_lib.check(L.evk_robust_norm_f32(grid.data_ptr(), nel, 1, 1 + round(0.95 * (nel - 1)), normed.data_ptr(), None, wsr.data_ptr(), wsr.numel(), None))
torch.cuda.synchronize()
del tp, tn, wst, outnp, wsnp, flow, xw, yw, wsf

# 5/6: f32 cmax f+g, and the dense-flow objective, on the 180x240 sensor
xs, ys = x * (239.0 / (W - 1)), y * (179.0 / (H - 1))
ts = (t - 1.0) * 0.05
wsc = torch.empty(L.evk_cmax_workspace_bytes(180, 240), dtype=torch.uint8, device=dev)
res = torch.empty(12, dtype=torch.float64, device=dev)
_lib.check(L.evk_cmax_linvel_variance_f32(xs.data_ptr(), ys.data_ptr(), ts.data_ptr(), p.data_ptr(), N, 1.0, 45.0, -20.0, 180, 240, 180, 240,
                                          1.0, _lib.CMAX_WANT_GRAD, res.data_ptr(), None, None, wsc.data_ptr(), wsc.numel(), None))
flow_c = torch.randn(2, 180, 240, device=dev) * 30
_lib.check(L.evk_cmax_flow_variance_f32(xs.data_ptr(), ys.data_ptr(), ts.data_ptr(), p.data_ptr(), N, flow_c.data_ptr(), 0.0, 180, 240, 1.0, 0,
                                        res.data_ptr(), None, wsc.data_ptr(), wsc.numel(), None))
torch.cuda.synchronize()
print("done", res.cpu().numpy()[:3])
