# the reference's lib/util/__init__.py star-exports event_util (and util, which is not on the path)
from .event_util import *    # noqa: F401,F403
