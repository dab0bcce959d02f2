"""xmaps_b200 — B200-native implementation of the X-maps per-event depth path.

Layout (only what the hot path needs):

* ``csrc/``          hand-written sm_100a CUDA kernels + the C-ABI (``include/xmaps_b200.h``)
* ``_native.py``     ctypes binding of the C-ABI shared library (fails loudly if it is missing)
* ``engine.py``      ``DepthEngine``: one device context per GPU, torch tensors in / out
* ``calibration.py``, ``time_map.py``, ``disparity.py``, ``depth.py``, ``frame_event_filter.py``,
  ``trigger_finder.py``, ``pipeline.py``  host-side mirror of the reference's Python call surface
  (the reference's own ``depth_reprojection_pipe.py`` / ``_processor.py`` run unchanged on top of it)
* ``lazy.py``, ``events.py``, ``host_stream.py``  lazy device handles, the EventCD record on the device,
  pinned-host streaming
* ``dropin/``        modules carrying the reference's own module names, for PYTHONPATH drop-in
* ``sharding.py``    round-robin frame sharding across GPUs + the NCCL gather of depth frames
"""
__version__ = "0.1.0"

EVENT_RECORD_BYTES = 16  # Metavision EventCD: x:u16 y:u16 p:i16 pad:u16 t:i64
