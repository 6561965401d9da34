"""The helpers of lib/util/event_util.py that the hot path and its drivers use."""
import numpy as np


def infer_resolution(xs, ys):
    """Sensor size guessed from the largest coordinates, [max(y)+1, max(x)+1] (event_util.py:5-13)."""
    return [np.max(ys) + 1, np.max(xs) + 1]


def events_bounds_mask(xs, ys, x_min, x_max, y_min, y_max):
    """
    1.0 where x_min < x <= x_max and y_min < y <= y_max, else 0.0 (strict lower, inclusive
    upper bound).  Drop-in for event_util.py:15-28.  Host-side numpy: inside the fused
    contrast-maximisation kernel the same test is evaluated per event on the device
    (csrc/evk_cmax.cu); this stand-alone form exists for callers that want the mask itself.
    """
    xs, ys = np.asarray(xs), np.asarray(ys)
    inside_x = ~((xs <= x_min) | (xs > x_max))
    inside_y = ~((ys <= y_min) | (ys > y_max))
    return (inside_x & inside_y).astype(np.float64)
