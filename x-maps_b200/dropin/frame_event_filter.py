"""Drop-in for the reference's ``frame_event_filter`` module (python/frame_event_filter.py)."""
from xmaps_b200.frame_event_filter import (  # noqa: F401
    FirstEventPerXYFilter,
    FirstEventPerYTFilter,
    FrameEventFilter,
    FrameEventFilterProcessor,
    LastEventPerXYFilter,
    MeanFirstLastEventPerXYFilter,
    NoFilter,
)
