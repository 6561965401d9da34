import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_utils_b200 import _lib
L = _lib.lib()
B, H, W = 5, 480, 640
N = 50_000_000
g = torch.Generator(device="cuda").manual_seed(2024)
x = torch.rand(N, device="cuda", generator=g) * 639; y = torch.rand(N, device="cuda", generator=g) * 479
t = torch.sort(torch.rand(N, device="cuda", generator=g)).values; p = (torch.randint(0, 2, (N,), device="cuda", generator=g) * 2 - 1).float()
hm = torch.rand(N, device="cuda") < 0.10
x[hm] = torch.where(torch.rand(int(hm.sum()), device="cuda") < 0.5, 17.0, 400.0)
y[hm] = torch.where(x[hm] == 17.0, 33.0, 301.0)
out = torch.empty((B, H, W), device="cuda")
ws = torch.empty(L.evk_voxel_workspace_bytes(B, H, W, 0), dtype=torch.uint8, device="cuda")
oob = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(3):
    _lib.check(L.evk_voxel_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), N, 0.0, 1.0, B, H, W, _lib.VARIANT_SMEM_TILE | _lib.AUTO_SPAN,
                               out.data_ptr(), ws.data_ptr(), ws.numel(), oob.data_ptr(), None))
torch.cuda.synchronize()
print("sum", float(out.double().sum()))
