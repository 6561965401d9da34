"""Optimiser drivers around the fused objective -- the callers of the hot path in the reference's
lib/contrast_max/events_cmax.py (`grid_search_initial` :241-311, `optimize_contrast` :313-346,
`optimize` :348-368).  They are thin: scipy's BFGS stays on the host and calls the fused GPU
evaluation; the grid search evaluates ALL its sample points with one pass over the events per 32
candidates (objectives.evaluate_candidates) instead of one full evaluation per point.
"""
import copy

import numpy as np
import scipy.optimize as opt

from .objectives import evaluate_candidates


def grid_search_initial(xs, ys, ts, ps, warp_function, objective_function, img_size, param_ranges=None,
        log_scale=True, num_samples_per_param=5):
    """
    Sample the objective on a (log- or linearly spaced) grid over the warp parameters; same inputs
    and returned dict as events_cmax.py:241-311: 'params', 'eval', 'search_axes', 'min_params',
    'min_func_eval'.  (As in the reference the blur is fixed at 1.0 and a point only becomes the
    optimum if its value is below 0.)
    """
    assert num_samples_per_param % 2 == 1
    half = int(num_samples_per_param / 2.0) + 1
    if log_scale:
        scale = np.logspace(0, 2.0, half)[1:]
        scale /= scale[-1]
    else:
        scale = np.linspace(0, 1.0, half)[1:]
    if param_ranges is None:
        param_ranges = [[-150, 150] for _ in range(warp_function.dims)]
    axes = []
    for lo, hi in param_ranges:
        span = hi - lo
        mid = lo + span / 2.0
        axes.append(np.concatenate((np.array(mid - scale * (span / 2.0))[::-1], np.array([mid]), np.array(mid + scale * (span / 2.0)))))
    coords = np.vstack([np.ravel(g) for g in np.meshgrid(*axes)])
    points = list(zip(*coords))
    evals, _ = evaluate_candidates(objective_function, points, xs, ys, ts, ps, warp_function, img_size, blur_sigma=1.0)
    output = {"params": [], "eval": [], "search_axes": axes}
    best_eval, best_params = 0, None
    for params, f_eval in zip(points, evals):
        output["params"].append(params)
        output["eval"].append(float(f_eval))
        if f_eval < best_eval:
            best_eval, best_params = float(f_eval), params
    output["min_params"] = best_params
    output["min_func_eval"] = best_eval
    return output


def optimize_contrast(xs, ys, ts, ps, warp_function, objective, optimizer=opt.fmin_bfgs, x0=None,
        numeric_grads=False, blur_sigma=None, img_size=(180, 240), grid_search_init=False, minimum_events=200):
    """
    Gradient-based contrast maximisation; same signature as events_cmax.py:313-346.  With
    grid_search_init the start point comes from grid_search_initial (the reference calls an
    undefined `recursive_search` there, events_cmax.py:336).
    @returns the maximising warp parameters
    """
    if grid_search_init and x0 is None:
        init_obj = copy.deepcopy(objective)
        init_obj.adaptive_lifespan = False
        x0 = grid_search_initial(xs, ys, ts, ps, warp_function, init_obj, img_size, log_scale=False)["min_params"]
        x0 = np.array([0, 0]) if x0 is None else np.array(x0)
    elif x0 is None:
        x0 = np.array([0, 0])
    objective.iter_update(x0)
    args = (xs, ys, ts, ps, warp_function, img_size, blur_sigma)
    if numeric_grads:
        return optimizer(objective.evaluate_function, x0, args=args, epsilon=1, disp=False, callback=objective.iter_update)
    return optimizer(objective.evaluate_function, x0, fprime=objective.evaluate_gradient, args=args, disp=False,
                     callback=objective.iter_update)


def optimize(xs, ys, ts, ps, warp, obj, numeric_grads=True, img_size=(180, 240)):
    """events_cmax.py:348-368: optimize_contrast with blur 1.0, numeric gradients unless the objective has analytic ones."""
    numeric_grads = numeric_grads if obj.has_derivative else True
    return optimize_contrast(xs, ys, ts, ps, warp, obj, numeric_grads=numeric_grads, blur_sigma=1.0, img_size=img_size)
