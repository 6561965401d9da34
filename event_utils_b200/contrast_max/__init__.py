"""Contrast maximisation: motion models (`warps`), objectives evaluated by the fused GPU pass
(`objectives`) and the optimiser / grid-search drivers around them (`events_cmax`).

Everything public in the three modules is re-exported here, so `from event_utils_b200.contrast_max
import linvel_warp, variance_objective, optimize_contrast` works the way the same import from the
reference's `lib.contrast_max` package does.
"""
from . import events_cmax as _drivers
from . import objectives as _objectives
from . import warps as _warps

for _module in (_drivers, _warps, _objectives):
    globals().update({_name: getattr(_module, _name) for _name in dir(_module) if not _name.startswith("_")})
del _module
