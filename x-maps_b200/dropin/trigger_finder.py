"""Drop-in for the reference's ``trigger_finder`` module (python/trigger_finder.py)."""
from xmaps_b200.trigger_finder import MIN_EVENTS_PER_FRAME, RobustTriggerFinder  # noqa: F401
