# same star-exports as the reference's lib/contrast_max/__init__.py
from .events_cmax import *   # noqa: F401,F403
from .warps import *         # noqa: F401,F403
from .objectives import *    # noqa: F401,F403
