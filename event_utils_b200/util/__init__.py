"""Helpers shared by the hot path and its drivers; the public names of `event_util` are re-exported,
as the reference's `lib.util` package does for its own event_util module."""
from . import event_util as _event_util

globals().update({_name: getattr(_event_util, _name) for _name in dir(_event_util) if not _name.startswith("_")})
