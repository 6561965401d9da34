"""Tensor normalisation applied to the voxel grid by the data loaders -- drop-in for the
`RobustNorm` transform of the reference's lib/data_loaders/data_augmentation.py:75-136."""
import torch

from .. import _lib
from ..representations import _events as E


class RobustNorm(object):
    """
    Robustly normalise a tensor: clamp it between its `low_perc`-th and `top_perc`-th percentile
    (exact order statistics, no interpolation) and rescale.  Same constructor, `percentile` and
    call semantics as the reference; both order statistics come from ONE 3-pass radix select on
    the GPU (csrc/evk_norm.cu) instead of two full sorts (kthvalue).
    """

    def __init__(self, low_perc=0, top_perc=95):
        self.top_perc = top_perc
        self.low_perc = low_perc

    @staticmethod
    def _rank(numel, q):
        # data_augmentation.py:100 -- python's round() (banker's rounding), one-based
        return 1 + round(.01 * float(q) * (numel - 1))

    @staticmethod
    def _run(t, k_low, k_top):
        L = _lib.lib()
        dev = E.compute_device(t)
        with torch.cuda.device(dev):
            x = t.detach().to(dev).to(torch.float32).contiguous()
            out = torch.empty_like(x)
            stats = torch.empty(2, dtype=torch.float32, device=dev)
            ws = _lib.scratch("norm_ws", L.evk_robust_norm_workspace_bytes(), dev)
            _lib.check(L.evk_robust_norm_f32(_lib.ptr(x), x.numel(), int(k_low), int(k_top), _lib.ptr(out), _lib.ptr(stats),
                                             _lib.ptr(ws), ws.numel(), _lib.stream()))
        return out, stats

    @staticmethod
    def percentile(t, q):
        """The q-th percentile of the flattened tensor (value of rank 1 + round(.01 q (numel-1)))."""
        k = RobustNorm._rank(t.numel(), q)
        _, stats = RobustNorm._run(t, k, k)
        return float(stats[0].item())

    def __call__(self, x, is_flow=False):
        out, stats = self._run(x, self._rank(x.numel(), self.low_perc), self._rank(x.numel(), self.top_perc))
        t_min, t_max = stats.tolist()
        if t_max == 0 and t_min == 0:
            return x                            # data_augmentation.py:122-123: the input object itself
        out = out.reshape(x.shape)
        if out.dtype != x.dtype and x.dtype.is_floating_point:
            out = out.to(x.dtype)               # the reference's clamp / divide keep the input dtype
        return out if out.device == x.device else out.to(x.device)

    def __repr__(self):
        return self.__class__.__name__ + '(top_perc={:.2f}, low_perc={:.2f})'.format(self.top_perc, self.low_perc)
