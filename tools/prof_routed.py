import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from event_utils_b200 import _lib
L = _lib.lib()
B, H, W = 5, 480, 640
N = int(os.environ.get("N", 50_000_000))
V = {"routed": _lib.VARIANT_ROUTED, "vector_red": _lib.VARIANT_VECTOR_RED}[os.environ.get("VARIANT", "routed")]
g = torch.Generator(device="cuda").manual_seed(2024)
x = torch.rand(N, device="cuda", generator=g) * 639; y = torch.rand(N, device="cuda", generator=g) * 479
t = torch.sort(torch.rand(N, device="cuda", generator=g)).values; p = (torch.randint(0, 2, (N,), device="cuda", generator=g) * 2 - 1).float()
out = torch.empty((B, H, W), device="cuda")
ws = torch.empty(L.evk_voxel_workspace_bytes(B, H, W, V), dtype=torch.uint8, device="cuda")
oob = torch.zeros(1, dtype=torch.int64, device="cuda")
for _ in range(int(os.environ.get("REPS", 3))):
    _lib.check(L.evk_voxel_f32(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), N, 0.0, 1.0, B, H, W, V | _lib.AUTO_SPAN,
                               out.data_ptr(), ws.data_ptr(), ws.numel(), oob.data_ptr(), None))
torch.cuda.synchronize()
print("sum", float(out.double().sum()))
