"""Motion models for contrast maximisation -- drop-in for the reference's lib/contrast_max/warps.py
(`warp_function` ABC :6-42, `linvel_warp` :44-61, the two empty stubs :63-83).

An objective from `event_utils_b200.contrast_max.objectives` that is handed a `linvel_warp` never
calls `warp()`: it recognises the model through `fused_kind` and evaluates warp, bounds mask, image
formation, objective and gradient in one fused GPU pass over the device-resident events
(csrc/evk_cmax.cu).  `warp()` is still provided, with the reference's semantics, for user code that
calls it directly and for the generic (non-fused) path taken by other `warp_function` subclasses.
"""
from abc import ABC, abstractmethod

import numpy as np


class warp_function(ABC):
    """A parametrised, differentiable motion model that transports events to a reference time.
    `name` identifies the model, `dims` is its number of degrees of freedom (used by the grid
    search drivers, events_cmax.py:282)."""

    fused_kind = None   # set by models the fused kernels implement natively

    def __init__(self, name, dims):
        self.name, self.dims = name, dims
        super().__init__()

    @abstractmethod
    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        """Transport the events (xs, ys, ts, ps) to time t0 under `params`.
        Returns (xs_warped, ys_warped, jacobian_x, jacobian_y); the Jacobians have shape
        (dims, N) -- d x'/d params and d y'/d params per event -- or are None without
        compute_grad."""


class linvel_warp(warp_function):
    """Global optic flow: every event moves with one image-plane velocity (vx, vy) = params,
        x' = x - (t - t0) vx,      y' = y - (t - t0) vy,
    so d x'/d vx = d y'/d vy = -(t - t0) and the cross terms vanish (reference warps.py:51-61)."""

    fused_kind = "linvel"

    def __init__(self):
        super().__init__('linvel_warp', 2)

    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        lag = ts - t0
        moved = (xs - lag * params[0], ys - lag * params[1])
        if not compute_grad:
            return moved[0], moved[1], None, None
        nothing = np.zeros_like(lag, dtype=np.float64)
        return moved[0], moved[1], np.stack((-lag, nothing)), np.stack((nothing, -lag))


class _placeholder_warp(warp_function):
    """The reference declares these models but leaves `warp` empty (it returns None)."""

    def __init__(self, name, dims):
        super().__init__(name, dims)

    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        return None


class xyztheta_warp(_placeholder_warp):
    """4-DoF x, y, z, rotation model (Mitrokhin et al.); an empty stub in the reference (warps.py:63-72)."""

    def __init__(self):
        super().__init__('xyztheta_warp', 4)


class pure_rotation_warp(_placeholder_warp):
    """Rotation about a centre (x, y) with angular velocity theta; an empty stub in the reference
    (warps.py:74-83), which also registers it with dims=4."""

    def __init__(self):
        super().__init__('pure_rotation_warp', 4)
