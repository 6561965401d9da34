"""Dense-flow event warp -- drop-in for the reference's lib/transforms/optic_flow.py."""
import torch

from .. import _lib
from ..representations import _events as E


def warp_events_flow_torch(xt, yt, tt, pt, flow_field, t0=None,
        batched=False, batch_indices=None):
    """
    Warp events by a per-pixel flow field: x' = x + u(x,y)*(t-t0), y' = y + v(x,y)*(t-t0).
    Drop-in for optic_flow.py:5-46: flow sampled bilinearly at the event position over integer
    pixel centres (grid_sample, align_corners=True), neighbours outside the field count as zero.
    @param xt, yt, tt, pt event components, shape (N,) or (N,1) (pt is unused, as in the reference)
    @param flow_field (2,H,W) or (1,2,H,W) tensor
    @param t0 reference time (default: the last timestamp)
    @returns warped_xt, warped_yt (shape (N,))
    """
    L = _lib.lib()
    xt, yt, tt = (E.as_tensor(a) for a in (xt, yt, tt))
    out_device = xt.device
    if len(xt.shape) > 1:
        xt, yt, tt = xt.squeeze(), yt.squeeze(), tt.squeeze()
    if t0 is None:
        t0 = tt[-1]
    t0 = float(t0)
    flow = E.as_tensor(flow_field)
    while flow.dim() < 4:
        flow = flow.unsqueeze(0)
    if flow.shape[0] != 1 or flow.shape[1] != 2:
        raise RuntimeError("flow_field must have shape (2,H,W) or (1,2,H,W), got %s" % (tuple(flow_field.shape),))
    dev = E.compute_device(xt, yt, tt, flow)
    with torch.cuda.device(dev):
        x, y = (a.reshape(-1).to(dev).to(torch.float32).contiguous() for a in (xt, yt))
        t = tt.reshape(-1).to(dev)
        if t.dtype != torch.float32:
            # the reference forms dt = tt - t0 in the INPUT dtype (optic_flow.py:33): absolute float64 / integer
            # stamps must not be rounded to float32 before the subtraction.  Pass the relative times, t0 = 0.
            t = (t - t0).to(torch.float32).contiguous()
            t0 = 0.0
        else:
            t = t.contiguous()
        f = flow[0].to(dev).to(torch.float32).contiguous()
        xw, yw = torch.empty_like(x), torch.empty_like(y)
        ws = _lib.scratch("flow_ws", L.evk_warp_flow_workspace_bytes(f.shape[1], f.shape[2]), dev)
        _lib.check(L.evk_warp_flow_f32(_lib.ptr(x), _lib.ptr(y), _lib.ptr(t), x.shape[0], _lib.ptr(f),
                                       f.shape[1], f.shape[2], t0, _lib.ptr(xw), _lib.ptr(yw), _lib.ptr(ws), ws.numel(),
                                       _lib.stream()))
    if xw.device != out_device:
        xw, yw = xw.to(out_device), yw.to(out_device)
    return xw, yw
