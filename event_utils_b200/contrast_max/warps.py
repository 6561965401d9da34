"""Warp functions -- drop-in for the reference's lib/contrast_max/warps.py.

`linvel_warp` is recognised by the objective functions in this package: events warped with it
never leave the GPU (warp, mask, IWE, objective and gradient are one fused kernel).  Its
`warp()` method still exists with the reference's semantics for code that calls it directly.
"""
from abc import ABC, abstractmethod

import numpy as np


class warp_function(ABC):
    """
    Base class of parametrised, differentiable motion models that move events to a reference
    time (reference: warps.py:6-42).
    """
    def __init__(self, name, dims):
        self.name = name
        self.dims = dims
        super().__init__()

    @abstractmethod
    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        """
        @returns xs_warped, ys_warped, xs_jacobian, ys_jacobian (jacobians (dims,N) or None)
        """
        pass


class linvel_warp(warp_function):
    """
    Linear velocity (global optic flow) warp, reference warps.py:44-61:
    x' = x - (t-t0)*vx, y' = y - (t-t0)*vy; d x'/d vx = d y'/d vy = -(t-t0).
    """
    fused_kind = "linvel"  # lets get_iwe / the objectives take the fused GPU path

    def __init__(self):
        warp_function.__init__(self, 'linvel_warp', 2)

    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        dt = ts - t0
        x_prime = xs - dt * params[0]
        y_prime = ys - dt * params[1]
        jacobian_x, jacobian_y = None, None
        if compute_grad:
            n = len(x_prime)
            jacobian_x = np.zeros((2, n))
            jacobian_y = np.zeros((2, n))
            jacobian_x[0, :] = -dt
            jacobian_y[1, :] = -dt
        return x_prime, y_prime, jacobian_x, jacobian_y


class xyztheta_warp(warp_function):
    """4-DoF x,y,z,rotation warp: an empty stub in the reference (warps.py:63-72), kept as one."""
    def __init__(self):
        warp_function.__init__(self, 'xyztheta_warp', 4)

    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        pass


class pure_rotation_warp(warp_function):
    """Pure rotation warp: an empty stub in the reference (warps.py:74-83), kept as one."""
    def __init__(self):
        warp_function.__init__(self, 'pure_rotation_warp', 4)

    def warp(self, xs, ys, ts, ps, t0, params, compute_grad=False):
        pass
